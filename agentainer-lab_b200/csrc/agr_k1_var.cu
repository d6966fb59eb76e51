// agr_k1_var.cu — K1 for VARIABLE-LENGTH records (BASELINE config 5: 128 B - 4 KB bodies).
//
// A variable-length record = the same 96 B header as agr_record + its payload (path | headers | body) rounded up to
// 16 B; a batch is one contiguous blob plus an offsets array (n + 1 entries).  Work is balanced by BYTES, not by
// records: the blob is cut into 8 KiB tiles and a tile OWNS the records that START in it (k1v_tile_index), so a warp
// always moves about the same number of bytes whether they hold 60 small requests or two 4 KB ones, and no record is
// ever split between warps.  Per tile: one 1-D bulk TMA copy (UBLKCP) of the owned records into the warp's stage behind
// an mbarrier; a warp-cooperative, position-weighted checksum per record (lane = 16 B chunk, REDUX per record); then
// the decisions thread-per-record with the same staged functions as the fixed-stride kernel (agr_device.cuh).
// Algorithmic bytes per record: its stored length (96 + 16-rounded payload) + 8.
#include "agr_device.cuh"

#define VT_TILE 8192u                 // bytes of blob per tile
#define VT_WARPS 16
#define VT_MAXREC 8192u               // longest legal record (header + payload)
#define VT_TAIL 4608u                 // stage room behind the tile for the last owned record (a 4 KB body fits); a record that
                                      // ends beyond the stage (> 4.5 KiB and badly placed: rare) is read from global memory
#define VT_STAGE (VT_TILE + VT_TAIL)  // 12800 B = 100 x 128
#define VT_MAXCNT (VT_TILE / 96u + 2u)   // at most this many records can start in one tile (87: three offsets per lane)

__device__ __forceinline__ uint32_t v_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 v_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// tile_first[t] = index of the first record whose start offset is >= t * VT_TILE, for t = 0 .. ntiles (tile_first[ntiles]
// = n).  Record i is that record for every boundary t * VT_TILE in (start(i-1), start(i)]; the sentinel i == n takes the
// boundaries after the last start.  Tile t then owns records [tile_first[t], tile_first[t+1]).
__global__ void __launch_bounds__(256) k1v_tile_index(const uint32_t* __restrict__ off, const uint32_t n, uint32_t* __restrict__ tile_first,
                                                      const uint32_t ntiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t lo = (i == 0) ? 0u : off[i - 1] / VT_TILE + 1u;
    const uint32_t hi = (i == n) ? ntiles : off[i] / VT_TILE;
    for (uint32_t t = lo; t <= hi; ++t) tile_first[t] = i;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k1_ingest_var(const agr_dev d, const uint8_t* __restrict__ blob, const uint32_t* __restrict__ off, const uint32_t* __restrict__ tile_first,
              const uint32_t ntiles, const uint32_t first_rid, const unsigned long long blob_base /* byte offset of blob in the slab */) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[WARPS];
    __shared__ uint32_t s_off[WARPS][2][VT_MAXCNT + 1];       // record offsets of the current / the next tile
    __shared__ unsigned long long s_ck[WARPS][VT_MAXCNT];
    __shared__ uint32_t s_ctr[K1_NLC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t smem_base = (v_smem_u32(smem_raw) + 127u) & ~127u;
    if (threadIdx.x < K1_NLC) s_ctr[threadIdx.x] = 0;
    if (threadIdx.x < WARPS) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(v_smem_u32(&bars[threadIdx.x])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t lc[K1_NLC];
#pragma unroll
    for (int c = 0; c < K1_NLC; ++c) lc[c] = 0;
    const uint32_t stage = smem_base + (uint32_t)warp * VT_STAGE;
    const uint32_t bar = v_smem_u32(&bars[warp]);
    const uint32_t tstride = gridDim.x * WARPS;

    // Tile metadata is software-pipelined two tiles deep so that no global-load latency sits on the warp's critical path:
    // while tile T is processed, the record offsets of T+1 are in flight into registers (their range [a1, b1) was read one
    // iteration earlier) and the range of T+2 is being read.  A tile inside one long record owns nothing (cnt == 0).
    auto issue = [&](int buf, uint32_t cnt) {
        if (lane == 0) {
            const uint32_t start = s_off[warp][buf][0];
            const uint32_t bytes = min(s_off[warp][buf][cnt] - start, VT_STAGE);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(stage), "l"(blob + start), "r"(bytes), "r"(bar) : "memory");
        }
    };
    auto range_of = [&](uint32_t t, uint32_t& ra, uint32_t& rb) {
        ra = rb = 0;
        if (t < ntiles) { ra = __ldg(&tile_first[t]); rb = __ldg(&tile_first[t + 1]); }
    };

    int cur = 0;
    uint32_t phase = 0;
    uint32_t tile = blockIdx.x * WARPS + warp;
    uint32_t a, b1v, a1, b1;
    range_of(tile, a, b1v);
    uint32_t cnt = b1v - a;
    if (cnt) {
        for (uint32_t k = lane; k <= cnt; k += 32) s_off[warp][cur][k] = __ldg(&off[a + k]);
        __syncwarp();
        issue(cur, cnt);
    }
    range_of(tile + tstride, a1, b1);
    while (tile < ntiles) {
        // offsets of the next tile -> registers (at most 3 per lane), range of the one after it -> registers
        const uint32_t ncnt = b1 - a1;
        uint32_t o0 = 0, o1 = 0, o2 = 0, a2, b2;
        if (ncnt) {
            if ((uint32_t)lane <= ncnt) o0 = __ldg(&off[a1 + lane]);
            if ((uint32_t)lane + 32u <= ncnt) o1 = __ldg(&off[a1 + lane + 32u]);
            if ((uint32_t)lane + 64u <= ncnt) o2 = __ldg(&off[a1 + lane + 64u]);
        }
        range_of(tile + 2u * tstride, a2, b2);
        auto stage_next = [&]() {                                 // called once this tile's stage has been consumed
            if (ncnt) {
                uint32_t* nso = s_off[warp][cur ^ 1];
                if ((uint32_t)lane <= ncnt) nso[lane] = o0;
                if ((uint32_t)lane + 32u <= ncnt) nso[lane + 32u] = o1;
                if ((uint32_t)lane + 64u <= ncnt) nso[lane + 64u] = o2;
                __syncwarp();
                issue(cur ^ 1, ncnt);
            }
        };
        if (cnt == 0) {                                           // nothing starts in this tile
            stage_next();
            tile += tstride; a = a1; cnt = ncnt; a1 = a2; b1 = b2; cur ^= 1;
            continue;
        }
        const uint32_t* so = s_off[warp][cur];
        const uint32_t start = so[0], bytes = min(so[cnt] - start, VT_STAGE);
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     ::"r"(bar), "r"(phase) : "memory");
        phase ^= 1u;
        // ---- checksum: one record at a time, lanes sweep its 16 B chunks (conflict-free), REDUX per record
        for (uint32_t r = 0; r < cnt; ++r) {
            const uint32_t lo = so[r] - start, chunks = (so[r + 1] - so[r]) >> 4;
            const bool in_stage = (so[r + 1] - start) <= bytes;                    // warp-uniform
            const uint8_t* g = blob + so[r];
            uint32_t c0 = 0, c1 = 0;
            for (uint32_t c = lane; c < chunks; c += 32) {
                const uint4 v = in_stage ? v_lds128(stage + lo + (c << 4)) : ldg_nc_v4(g + ((size_t)c << 4));
                c0 += v.x + v.y + v.z + v.w;
                c1 += (4 * c + 1) * v.x + (4 * c + 2) * v.y + (4 * c + 3) * v.z + (4 * c + 4) * v.w;
            }
            c0 = __reduce_add_sync(FULL, c0);
            c1 = __reduce_add_sync(FULL, c1);
            if (lane == 0) s_ck[warp][r] = agr_cksum_pack(c0, c1);
        }
        __syncwarp();
        // ---- decisions, one record per lane.  Round 0 (records 0..31 of the tile) reads its headers, THEN the stage is
        // handed back to the TMA unit for the next tile, and only then does the long-latency decision chain run.
        auto load_header = [&](uint32_t r, uint4& h0, uint4& h1, uint4& h2, uint4& h3, uint4& h4, uint4& h5) {
            const uint32_t hb = stage + (so[r] - start);
            if ((so[r] - start) + 96u <= bytes) {
                h0 = v_lds128(hb); h1 = v_lds128(hb + 16); h2 = v_lds128(hb + 32); h3 = v_lds128(hb + 48); h4 = v_lds128(hb + 64); h5 = v_lds128(hb + 80);
            } else {
                const uint8_t* g = blob + so[r];
                h0 = ldg_nc_v4(g); h1 = ldg_nc_v4(g + 16); h2 = ldg_nc_v4(g + 32); h3 = ldg_nc_v4(g + 48); h4 = ldg_nc_v4(g + 64); h5 = ldg_nc_v4(g + 80);
            }
        };
        auto decide = [&](uint32_t r, const uint4& h0, const uint4& h1, const uint4& h2, const uint4& h3, const uint4& h4, const uint4& h5) {
            const uint32_t rid = first_rid + a + r;
            k1_ctx cx;
            k1_begin(d, k1_agent_issue(d, h2, h3), h0, h2, h3, h4, cx);
            const k1_result res = k1_finish(d, rid, h1, h5, cx, lc);
            d.state[rid] = res.state;
            d.route[rid] = res.route;
            d.cksum[rid] = s_ck[warp][r];
            d.voff[rid] = blob_base + so[r];
            d.vlen[rid] = so[r + 1] - so[r];
        };
        for (uint32_t r = 32 + lane; r < cnt; r += 32) {          // rounds 1.. (tiles of many small records) first
            uint4 h0, h1, h2, h3, h4, h5;
            load_header(r, h0, h1, h2, h3, h4, h5);
            decide(r, h0, h1, h2, h3, h4, h5);
        }
        uint4 h0 = make_uint4(0, 0, 0, 0), h1 = h0, h2 = h0, h3 = h0, h4 = h0, h5 = h0;
        const bool mine = (uint32_t)lane < cnt;
        if (mine) load_header(lane, h0, h1, h2, h3, h4, h5);
        __syncwarp();                                             // the stage has been consumed
        stage_next();
        if (mine) decide(lane, h0, h1, h2, h3, h4, h5);
        __syncwarp();
        tile += tstride; a = a1; cnt = ncnt; a1 = a2; b1 = b2; cur ^= 1;
    }
    k1_flush_counters(d, lc, s_ctr);
}

// gather of variable-length rows: lens pass, then copy pass (offsets are scanned on the host: a per-tick operation)
__global__ void __launch_bounds__(256) k_var_lens(const agr_dev d, const uint32_t* __restrict__ rids, const uint32_t n, uint32_t* __restrict__ lens) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) lens[j] = d.vlen[rids[j]];
}
__global__ void __launch_bounds__(256) k_var_copy(const agr_dev d, const uint32_t* __restrict__ rids, const uint32_t n,
                                                  const unsigned long long* __restrict__ out_off, uint8_t* __restrict__ out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t rid = rids[w];
    const uint8_t* src = d.slab + d.voff[rid];
    uint8_t* dst = out + out_off[w];
    const uint32_t chunks = d.vlen[rid] >> 4;
    for (uint32_t c = lane; c < chunks; c += 32) {
        uint4 v = ldg_nc_v4(src + (size_t)c * 16);
        if (c == 0 && (d.cfg_flags & AGR_CFG_MINT_IDS)) {
            unsigned long long lo, hi;
            agr_mint_id(row_logical(d, rid), d.shard_id, d.id_gen, d.id_secret, lo, hi);
            v = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        }
        if (c == 5) {   // bytes 80..95: body_len | status,retry,max,err | resp_status  (live state patched in)
            const uint32_t st = d.state[rid], aux = d.aux[rid];
            v.y = st_status(st) | (st_retry(st) << 8) | (st_max(st) << 16) | (((aux >> AUX_ERR_SHIFT) & 0xffu) << 24);
            v.z = (v.z & 0xffff0000u) | (aux & 0xffffu);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)c * 16) = v;
    }
}

cudaError_t agr_launch_k1_var(const agr_dev& d, const uint8_t* blob, const uint32_t* off, uint32_t n, unsigned long long blob_bytes,
                              uint32_t* tile_first, uint32_t first_rid, unsigned long long blob_base, int sm_count, cudaStream_t st) {
    constexpr int WARPS = VT_WARPS;
    const size_t smem = (size_t)WARPS * VT_STAGE + 128;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k1_ingest_var<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const uint32_t ntiles = (uint32_t)((blob_bytes + VT_TILE - 1) / VT_TILE);
    k1v_tile_index<<<(n + 1 + 255u) / 256u, 256, 0, st>>>(off, n, tile_first, ntiles);
    uint32_t blocks = (uint32_t)sm_count;
    const uint32_t need = (ntiles + WARPS - 1) / WARPS;
    if (blocks > need) blocks = need;
    if (blocks == 0) blocks = 1;
    k1_ingest_var<WARPS><<<blocks, WARPS * 32, smem, st>>>(d, blob, off, tile_first, ntiles, first_rid, blob_base);
    return cudaGetLastError();
}
void agr_launch_var_lens(const agr_dev& d, const uint32_t* rids, uint32_t n, uint32_t* lens, cudaStream_t st) {
    if (n) k_var_lens<<<(n + 255u) / 256u, 256, 0, st>>>(d, rids, n, lens);
}
void agr_launch_var_copy(const agr_dev& d, const uint32_t* rids, uint32_t n, const unsigned long long* out_off, uint8_t* out, cudaStream_t st) {
    if (n) k_var_copy<<<(n * 32u + 255u) / 256u, 256, 0, st>>>(d, rids, n, out_off, out);
}
