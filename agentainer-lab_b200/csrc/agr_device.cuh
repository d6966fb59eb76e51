// agr_device.cuh — device helpers shared by the K1 variants and K2/K3: cache-hinted loads, the 128-bit CAS, the
// agent-table and dedupe-index probes, and the per-record decision of proxyToAgentHandler + StoreRequest.
#pragma once
#include "agr_kernels.cuh"
#include "../../include/agentainer_gpu.h"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// 128-bit compare-and-swap on a dedupe-index slot's id.  The result stays an opaque 128-bit value (no unpacking
// at the issue point), so nothing depends on the atomic's return until the caller actually inspects it.
typedef unsigned __int128 u128;
__device__ __forceinline__ u128 make_u128(unsigned long long lo, unsigned long long hi) { return ((u128)hi << 64) | lo; }
__device__ __forceinline__ u128 cas128(void* addr, u128 cmp, u128 val) {
    u128 old;
    asm volatile("atom.relaxed.gpu.global.cas.b128 %0, [%1], %2, %3;" : "=q"(old) : "l"(addr), "q"(cmp), "q"(val) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long pack64(uint32_t lo, uint32_t hi) {
    return ((unsigned long long)hi << 32) | lo;
}

// dedupe-index lookup: returns slot index or ~0ULL
__device__ __forceinline__ unsigned long long table_find(const agr_dev& d, unsigned long long lo, unsigned long long hi) {
    unsigned long long idx = agr_hash_id(lo, hi) & d.table_mask;
    for (unsigned long long probe = 0; probe <= d.table_mask; ++probe) {
        const agr_slot* s = d.table + idx;
        const uint4 k = __ldcg(reinterpret_cast<const uint4*>(s));
        unsigned long long klo = pack64(k.x, k.y), khi = pack64(k.z, k.w);
        if ((klo | khi) == 0ULL) return ~0ULL;
        if (klo == lo && khi == hi) return idx;
        idx = (idx + 1) & d.table_mask;
    }
    return ~0ULL;
}

// logical <-> physical row (see agr_dev)
__device__ __forceinline__ unsigned long long row_logical(const agr_dev& d, uint32_t p) {
    if (!d.ring_rows) return p;
    return d.tail + (p >= d.tail_phys ? p - d.tail_phys : p + d.ring_rows - d.tail_phys);
}
__device__ __forceinline__ uint32_t row_physical(const agr_dev& d, unsigned long long l) {
    return d.ring_rows ? (uint32_t)(l % d.ring_rows) : (uint32_t)l;
}

// address of a row's record: fixed 512 B stride, or the byte offset kept per row in variable-length mode
__device__ __forceinline__ const uint8_t* rec_ptr(const agr_dev& d, uint32_t rid) {
    return d.voff ? d.slab + d.voff[rid] : d.slab + (size_t)rid * AGR_REC;
}

// (agent slot is checked by the callers) request id -> row holding the record stored under it, or AGR_RID_NONE.
// Hash mode: probe the dedupe index.  Mint mode: decode the row from the id and accept only an exact 128-bit match.
__device__ __forceinline__ uint32_t lookup_rid(const agr_dev& d, unsigned long long lo, unsigned long long hi) {
    if ((lo | hi) == 0ULL) return AGR_RID_NONE;
    if (d.cfg_flags & AGR_CFG_MINT_IDS) {
        unsigned long long rid; uint32_t shard, gen;
        if (!agr_unmint_id(lo, d.id_secret, rid, shard, gen)) return AGR_RID_NONE;
        if (rid < d.tail || rid >= d.head_l || shard != d.shard_id || gen != d.id_gen) return AGR_RID_NONE;   // not a live row
        unsigned long long mlo, mhi;
        agr_mint_id(rid, shard, gen, d.id_secret, mlo, mhi);
        return (mlo == lo && mhi == hi) ? row_physical(d, rid) : AGR_RID_NONE;
    }
    const unsigned long long idx = table_find(d, lo, hi);
    if (idx == ~0ULL) return AGR_RID_NONE;
    const uint32_t inv = __ldcg(&d.table[idx].inv_rid);
    return inv ? ~inv : AGR_RID_NONE;
}

// ------------------------------------------------------------------------------------------------ K1
// Decision + persistence for ONE record whose 96 B header is in registers: the sequential semantics of
// proxyToAgentHandler (server.go:498-541) with StoreRequest (requests.go:64-117) inlined, split in stages so a
// kernel can put independent work (the record checksum) between the long-latency global operations:
//   k1_agent_issue   -> loads of the first agent-table probe                      (GetAgent, server.go:498)
//   k1_begin         -> resolves the agent, classifies, issues the FIRST index CAS (StoreRequest's SET)
//   k1_finish        -> finishes probing, publishes the row id, forms verdict + state word
// The index insert is ONE returning atomic (CAS.128 on the id) plus a fire-and-forget RED.max of ~rid.  Who owns
// an id is therefore decided by the final value of inv_rid (lowest row wins = arrival order); a row that found its
// id already present is provisionally a duplicate and bumps dupfix, and only then does k1_post re-check owners.
struct k1_result { uint32_t state, route; };
struct ag_probe { uint4 a, b, t; uint32_t idx; };
struct k1_ctx {
    unsigned long long id_lo, id_hi, tidx;
    u128 old;
    uint32_t slot, astatus;
    bool replay, want_store, cas_issued, deferred;
};

__device__ __forceinline__ ag_probe agent_probe_load(const agr_dev& d, uint32_t idx) {
    ag_probe p;
    const agr_agent_key* e = d.akeys + idx;
    p.a = ldg_v4(&e->w[0]); p.b = ldg_v4(&e->w[2]); p.t = ldg_v4(&e->slot); p.idx = idx;
    return p;
}
__device__ __forceinline__ ag_probe k1_agent_issue(const agr_dev& d, const uint4& h2, const uint4& h3) {
    return agent_probe_load(d, (uint32_t)agr_hash_agent(pack64(h2.x, h2.y), pack64(h2.z, h2.w), pack64(h3.x, h3.y), pack64(h3.z, h3.w)) & d.amask);
}
// resolves (slot, status) from the first probe, continuing the linear probe if needed
__device__ __forceinline__ void agent_resolve(const agr_dev& d, ag_probe p, const uint4& h2, const uint4& h3, uint32_t& slot, uint32_t& status) {
    slot = RT_SLOT_NONE; status = AG_STATUS_REMOVED;
    if ((h2.x | h2.y | h2.z | h2.w | h3.x | h3.y | h3.z | h3.w) == 0u) return;
    for (uint32_t probe = 0; probe <= d.amask; ++probe) {
        if ((p.a.x | p.a.y | p.a.z | p.a.w | p.b.x | p.b.y | p.b.z | p.b.w) == 0u) return;          // empty: miss
        if (p.a.x == h2.x && p.a.y == h2.y && p.a.z == h2.z && p.a.w == h2.w && p.b.x == h3.x && p.b.y == h3.y &&
            p.b.z == h3.z && p.b.w == h3.w) { slot = p.t.x; status = p.t.y & 0xffu; return; }
        p = agent_probe_load(d, (p.idx + 1) & d.amask);
    }
}

__device__ __forceinline__ void k1_begin(const agr_dev& d, const ag_probe& ap, const uint4& h0, const uint4& h2, const uint4& h3,
                                         const uint4& h4, k1_ctx& c) {
    c.id_lo = pack64(h0.x, h0.y); c.id_hi = pack64(h0.z, h0.w);
    c.replay = (h4.z & AGR_F_REPLAY) != 0;                                                // server.go:506
    agent_resolve(d, ap, h2, h3, c.slot, c.astatus);                                      // server.go:498
    const bool found = c.slot != RT_SLOT_NONE && c.astatus != AG_STATUS_REMOVED;
    c.want_store = found && (d.cfg_flags & AGR_CFG_PERSISTENCE) && !c.replay;             // server.go:508
    // split mode: the insert is done by k1_index after the stream kernel; the row is provisionally "stored"
    // mint mode: the id is a function of the row, nothing to insert; the caller's request_id is ignored
    c.deferred = c.want_store && ((d.cfg_flags & AGR_CFG_MINT_IDS) || ((c.id_lo | c.id_hi) != 0ULL && (d.cfg_flags & AGR_CFGI_SPLIT_INDEX)));
    c.cas_issued = c.want_store && (c.id_lo | c.id_hi) != 0ULL && !c.deferred && !(d.cfg_flags & AGR_CFG_DIAG_NO_INDEX);
    c.old = 0;
    c.tidx = 0;
    if (c.cas_issued) {
        c.tidx = agr_hash_id(c.id_lo, c.id_hi) & d.table_mask;
        c.old = cas128(&d.table[c.tidx], 0, make_u128(c.id_lo, c.id_hi));
    }
}

__device__ __forceinline__ k1_result k1_finish(const agr_dev& d, uint32_t rid, const uint4& h1, const uint4& h5, k1_ctx& c,
                                               uint32_t* lc /*local counters*/) {
    k1_result out{0u, 0u};
    lc[C_INGESTED]++;
    if (c.slot == RT_SLOT_NONE || c.astatus == AG_STATUS_REMOVED) {                       // server.go:499-502
        lc[C_NOT_FOUND]++;
        out.route = RT_SLOT_NONE | (AGR_V_NOT_FOUND << RT_CODE_SHIFT);
        return out;
    }
    uint32_t vflags = 0;
    bool tracked = false;
    if (c.want_store) {
        // StoreRequest: SET rec (the row itself, already in the slab) + index insert + RPUSH pending (INQ bit)
        bool ok = c.cas_issued || c.deferred;
        if (c.cas_issued) {
            const u128 key = make_u128(c.id_lo, c.id_hi);
            for (;;) {
                if (c.old == 0) break;                                                    // claimed an empty slot
                if (c.old == key) {                                                       // id already present
                    ok = false;
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(d.dupfix) : "memory");
                    break;
                }
                c.tidx = (c.tidx + 1) & d.table_mask;
                c.old = cas128(&d.table[c.tidx], 0, key);
            }
            asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(&d.table[c.tidx].inv_rid), "r"(~rid) : "memory");
        }
        if (ok) {
            uint32_t maxr = (h5.y >> 16) & 0xffu;
            if (maxr == 0) maxr = 3;                                                      // requests.go:95
            out.state = AGR_ST_PENDING | ST_INQ | ST_STORED | (maxr << ST_MAX_SHIFT);     // requests.go:93-95,111
            vflags |= AGR_VF_STORED | AGR_VF_TRACKED;
            tracked = true;
            lc[C_STORED]++;
        } else {
            vflags |= AGR_VF_DUP_ID;                                                      // server.go:511-514 path
            lc[C_DUP_IDS]++;
        }
    } else if (c.replay) {                                                                // server.go:519-522
        vflags |= AGR_VF_REPLAY;
        lc[C_REPLAY]++;
        tracked = (h1.x | h1.y | h1.z | h1.w) != 0u;
        if (tracked) vflags |= AGR_VF_TRACKED;
    }
    uint32_t code;
    if (c.astatus != AGR_AGENT_RUNNING) {                                                 // server.go:525
        if ((d.cfg_flags & AGR_CFG_PERSISTENCE) && tracked) { code = AGR_V_QUEUED; lc[C_QUEUED]++; }   // :526-536
        else { code = AGR_V_UNAVAILABLE; lc[C_UNAVAILABLE]++; }                           // :539-540
    } else {
        code = AGR_V_FORWARD; lc[C_FORWARDED]++;                                          // :546-572
        if (out.state) out.state |= ST_INFLIGHT;
    }
    out.route = c.slot | (code << RT_CODE_SHIFT) | (vflags << RT_FLAG_SHIFT);
    return out;
}

#define K1_NLC 9   // counters C_INGESTED .. C_DUP_IDS are contiguous from 0

__device__ __forceinline__ void k1_flush_counters(const agr_dev& d, uint32_t* lc, uint32_t* s_ctr) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < K1_NLC; ++c) {
        uint32_t v = __reduce_add_sync(FULL, lc[c]);
        if (lane == 0 && v) atomicAdd(&s_ctr[c], v);
    }
    __syncthreads();
    if (threadIdx.x < K1_NLC && s_ctr[threadIdx.x]) atomicAdd(&d.ctr[threadIdx.x], (unsigned long long)s_ctr[threadIdx.x]);
    // per-batch note for k1_post: replay-flagged records need its "stored earlier?" pass
    if (threadIdx.x == C_REPLAY && s_ctr[C_REPLAY]) atomicAdd(d.dupfix + 2, s_ctr[C_REPLAY]);
}

