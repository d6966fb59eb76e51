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

struct u128 { unsigned long long lo, hi; };
__device__ __forceinline__ u128 cas128(void* addr, u128 cmp, u128 val) {
    u128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\t"
                 "mov.b128 c, {%2, %3};\n\t"
                 "mov.b128 v, {%4, %5};\n\t"
                 "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
                 "mov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr) : "memory");
    return old;
}
__device__ __forceinline__ unsigned long long pack64(uint32_t lo, uint32_t hi) {
    return ((unsigned long long)hi << 32) | lo;
}

// agent id -> (slot, status).  Table is tiny (48 B per entry) and read-only inside a kernel: L1 / L2 resident.
__device__ __forceinline__ uint32_t agent_lookup(const agr_dev& d, unsigned long long w0, unsigned long long w1,
                                                 unsigned long long w2, unsigned long long w3) {
    if ((w0 | w1 | w2 | w3) == 0ULL) return RT_SLOT_NONE;
    uint32_t idx = (uint32_t)agr_hash_agent(w0, w1, w2, w3) & d.amask;
    for (uint32_t probe = 0; probe <= d.amask; ++probe) {
        const agr_agent_key* e = d.akeys + idx;
        uint4 a = ldg_v4(&e->w[0]);
        uint4 b = ldg_v4(&e->w[2]);
        unsigned long long e0 = pack64(a.x, a.y), e1 = pack64(a.z, a.w), e2 = pack64(b.x, b.y), e3 = pack64(b.z, b.w);
        if ((e0 | e1 | e2 | e3) == 0ULL) return RT_SLOT_NONE;
        if (e0 == w0 && e1 == w1 && e2 == w2 && e3 == w3) return __ldg(&e->slot);
        idx = (idx + 1) & d.amask;
    }
    return RT_SLOT_NONE;
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

// ------------------------------------------------------------------------------------------------ K1
// Decision + persistence for ONE record whose 96 B header is in registers.  Sequential semantics of
// proxyToAgentHandler (server.go:498-541) with StoreRequest (requests.go:64-117) inlined.
struct k1_result { uint32_t state, route; };

__device__ __forceinline__ k1_result k1_decide(const agr_dev& d, uint32_t rid, uint4 h0, uint4 h1, uint4 h2, uint4 h3,
                                               uint4 h4, uint4 h5, uint32_t* lc /*local counters*/) {
    k1_result out{0u, 0u};
    const unsigned long long id_lo = pack64(h0.x, h0.y), id_hi = pack64(h0.z, h0.w);
    const uint32_t flags_in = h4.z;
    const bool replay = (flags_in & AGR_F_REPLAY) != 0;                                   // server.go:506
    const bool persistence = (d.cfg_flags & AGR_CFG_PERSISTENCE) != 0;
    // GetAgent (server.go:498, agent.go:372-390)
    uint32_t slot = agent_lookup(d, pack64(h2.x, h2.y), pack64(h2.z, h2.w), pack64(h3.x, h3.y), pack64(h3.z, h3.w));
    uint32_t astatus = AG_STATUS_REMOVED;
    if (slot != RT_SLOT_NONE) astatus = d.astatus[slot];
    lc[C_INGESTED]++;
    if (slot == RT_SLOT_NONE || astatus == AG_STATUS_REMOVED) {                           // server.go:499-502
        lc[C_NOT_FOUND]++;
        out.route = RT_SLOT_NONE | (AGR_V_NOT_FOUND << RT_CODE_SHIFT);
        return out;
    }
    uint32_t vflags = 0;
    bool tracked = false;
    if (persistence && !replay) {                                                         // server.go:508
        // StoreRequest: SET rec (the row itself, already in the slab) + index insert + RPUSH pending (INQ bit)
        bool ok = (id_lo | id_hi) != 0ULL;
        if (ok) {
            unsigned long long idx = agr_hash_id(id_lo, id_hi) & d.table_mask;
            const u128 zero{0ULL, 0ULL}, key{id_lo, id_hi};
            for (;;) {
                u128 old = cas128(&d.table[idx], zero, key);
                if ((old.lo | old.hi) == 0ULL || (old.lo == id_lo && old.hi == id_hi)) break;
                idx = (idx + 1) & d.table_mask;
            }
            const uint32_t inv = ~rid;
            uint32_t prev = atomicMax(&d.table[idx].inv_rid, inv);
            if (prev > inv) ok = false;                       // an EARLIER row owns this id: duplicate
            else if (prev != 0u) atomicAdd(d.dupfix, 1u);     // a LATER row raced ahead: k1_post demotes it
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
    } else if (replay) {                                                                  // server.go:519-522
        vflags |= AGR_VF_REPLAY;
        lc[C_REPLAY]++;
        tracked = (pack64(h1.x, h1.y) | pack64(h1.z, h1.w)) != 0ULL;
        if (tracked) vflags |= AGR_VF_TRACKED;
    }
    uint32_t code;
    if (astatus != AGR_AGENT_RUNNING) {                                                   // server.go:525
        if (persistence && tracked) { code = AGR_V_QUEUED; lc[C_QUEUED]++; }              // :526-536
        else { code = AGR_V_UNAVAILABLE; lc[C_UNAVAILABLE]++; }                           // :539-540
    } else {
        code = AGR_V_FORWARD; lc[C_FORWARDED]++;                                          // :546-572
        if (out.state) out.state |= ST_INFLIGHT;
    }
    out.route = slot | (code << RT_CODE_SHIFT) | (vflags << RT_FLAG_SHIFT);
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
}

