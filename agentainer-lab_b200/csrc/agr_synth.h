// agr_synth.h — counter-based generator of BASELINE.json's synthetic request streams.
// Integer-only and a pure function of (seed, record index), so the host build (gcc) and the device build (nvcc)
// produce byte-identical records; tests/test_synth.py checks that on the GPU box.
//
// Stream shape follows SURVEY.md 8(d): 512 B records, POST /agent/<id>/chat (examples/gpt-agent/app.py:70-79 for
// the body shape), agent ids "agent-<unixnano>" (internal/agent/agent.go:594-596), uniform or Zipf agent choice,
// and replay-flagged duplicates (internal/api/server.go:506-522) that name an EARLIER fresh record of the same
// agent in replay_of.
#pragma once
#include "agr_common.h"

struct agr_synth_dev {
    unsigned long long seed;
    uint32_t n_agents;
    uint32_t dup_permille;
    unsigned long long agent_nanos0;
    const unsigned long long* cdf;   // n_agents cumulative thresholds (pick smallest k with u <= cdf[k]); NULL = uniform
    // engine-minted ids (AGR_CFG_MINT_IDS): a duplicate names its target by the id the ENGINE minted for the target's row
    uint32_t mint, mint_shard, mint_gen, pad;
    unsigned long long mint_base_rid, mint_secret;   // stream index j lives in row mint_base_rid + j
};

AGR_HD unsigned long long agr_splitmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ULL;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
AGR_HD unsigned long long agr_rnd(unsigned long long seed, unsigned long long i, uint32_t k) {
    return agr_splitmix64(seed ^ agr_splitmix64(i * 64ULL + k));
}
AGR_HD bool agr_synth_raw_replay(const agr_synth_dev& s, unsigned long long i) {
    return s.dup_permille != 0 && i > 0 && (agr_rnd(s.seed, i, 0) % 1000ULL) < s.dup_permille;
}
AGR_HD uint32_t agr_synth_fresh_agent(const agr_synth_dev& s, unsigned long long i) {
    unsigned long long u = agr_rnd(s.seed, i, 1);
    if (s.cdf == nullptr) return (uint32_t)(u % s.n_agents);
    uint32_t lo = 0, hi = s.n_agents - 1;          // smallest k with u <= cdf[k]
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (u <= s.cdf[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}
AGR_HD void agr_synth_id(const agr_synth_dev& s, unsigned long long i, unsigned long long& lo, unsigned long long& hi) {
    lo = agr_rnd(s.seed, i, 2);
    hi = agr_rnd(s.seed, i, 3);
    // RFC 4122 v4: byte 6 high nibble = 4, byte 8 top bits = 10 (bytes are little-endian within lo / hi)
    lo = (lo & ~(0xf0ULL << 48)) | (0x40ULL << 48);
    hi = (hi & ~0xc0ULL) | 0x80ULL;
}
// decimal "agent-<nanos>" into a 32 B NUL padded buffer; returns the string length
AGR_HD uint32_t agr_synth_agent_name(unsigned long long nanos, char* out /*32*/) {
    for (int k = 0; k < 32; ++k) out[k] = 0;
    out[0] = 'a'; out[1] = 'g'; out[2] = 'e'; out[3] = 'n'; out[4] = 't'; out[5] = '-';
    char tmp[20]; int nd = 0;
    do { tmp[nd++] = (char)('0' + (nanos % 10ULL)); nanos /= 10ULL; } while (nanos && nd < 20);
    for (int k = 0; k < nd; ++k) out[6 + k] = tmp[nd - 1 - k];
    return 6u + (uint32_t)nd;
}

#define AGR_SYNTH_HDRS "Content-Type: application/json\nUser-Agent: agr-synth/1\n"
#define AGR_SYNTH_HDRS_LEN 55u

// Writes record i (512 B) to out.  out must be 8-byte aligned.
AGR_HD void agr_synth_record(const agr_synth_dev& s, unsigned long long i, unsigned char* out) {
    unsigned long long* o64 = (unsigned long long*)out;
    for (int k = 0; k < 64; ++k) o64[k] = 0;
    bool replay = false;
    unsigned long long target = 0;
    if (agr_synth_raw_replay(s, i)) {
        for (uint32_t t = 0; t < 8; ++t) {
            unsigned long long j = agr_rnd(s.seed, i, 8 + t) % i;
            if (!agr_synth_raw_replay(s, j)) { replay = true; target = j; break; }
        }
    }
    unsigned long long id_lo, id_hi;
    agr_synth_id(s, i, id_lo, id_hi);
    o64[0] = id_lo; o64[1] = id_hi;
    uint32_t agent;
    if (replay) {
        unsigned long long t_lo, t_hi;
        if (s.mint) agr_mint_id(s.mint_base_rid + target, s.mint_shard, s.mint_gen, s.mint_secret, t_lo, t_hi);
        else agr_synth_id(s, target, t_lo, t_hi);
        o64[2] = t_lo; o64[3] = t_hi;
        agent = agr_synth_fresh_agent(s, target);
    } else {
        agent = agr_synth_fresh_agent(s, i);
    }
    char* aid = (char*)(out + AGR_OFF_AGENT_ID);
    uint32_t alen = agr_synth_agent_name(s.agent_nanos0 + (unsigned long long)agent * 1000003ULL, aid);
    o64[AGR_OFF_SEQ / 8] = i + 1;
    uint32_t flags = (2u << 8) | (replay ? 1u : 0u);                 // POST
    *(uint32_t*)(out + AGR_OFF_FLAGS) = flags;
    unsigned char* p = out + AGR_OFF_PAYLOAD;
    uint32_t n = 0;
    const char pre[] = "/agent/";
    for (int k = 0; k < 7; ++k) p[n++] = (unsigned char)pre[k];
    for (uint32_t k = 0; k < alen; ++k) p[n++] = (unsigned char)aid[k];
    const char suf[] = "/chat";
    for (int k = 0; k < 5; ++k) p[n++] = (unsigned char)suf[k];
    uint32_t path_len = n;
    const char hdrs[] = AGR_SYNTH_HDRS;
    for (uint32_t k = 0; k < AGR_SYNTH_HDRS_LEN; ++k) p[n++] = (unsigned char)hdrs[k];
    uint32_t body_len = (AGR_REC - AGR_OFF_PAYLOAD) - n;              // fill the record: "512 B records"
    const char bpre[] = "{\"message\":\"";
    for (int k = 0; k < 12; ++k) p[n++] = (unsigned char)bpre[k];
    uint32_t fill = body_len - 14;
    const char alpha[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789 ,";
    uint32_t c = 0, kk = 16;
    while (c < fill) {
        unsigned long long r = agr_rnd(s.seed, i, kk++);
        for (int b = 0; b < 10 && c < fill; ++b, ++c) { p[n++] = (unsigned char)alpha[r & 63]; r >>= 6; }
    }
    p[n++] = '"'; p[n++] = '}';
    *(uint16_t*)(out + AGR_OFF_PATH_LEN) = (uint16_t)path_len;
    *(uint16_t*)(out + AGR_OFF_HDR_LEN) = (uint16_t)AGR_SYNTH_HDRS_LEN;
    *(uint32_t*)(out + AGR_OFF_BODY_LEN) = body_len;
    out[AGR_OFF_STATUS] = 1;        // pending
    out[AGR_OFF_RETRY] = 0;
    out[AGR_OFF_MAX_RETRIES] = 3;   // internal/requests/requests.go:95
}
