// agr_k5_json.cu — K5: the wire form of a stored record.
//
// The reference keeps every record in Redis as json.Marshal(requests.Request) (requests.go:101,170,265) and its
// management surface hands that JSON on: GET /agents/{id}/requests marshals []*Request (server.go:626-652),
// GET /agents/{id}/requests/{reqId} re-marshals one record (server.go:655-679), the CLI parses both
// (cmd/agentainer/main.go:1143-1174).  K5 produces those bytes from the binary rows: field order = struct order
// (requests.go:27-41,44-49), encoding/json's string escaping with HTML escaping on, map keys in sorted order,
// []byte as padded std base64, time.Time as RFC 3339 with nanoseconds (UTC), omitempty on processed_at / response /
// error.  Two passes of the same code: MEASURE gives every record's exact length, a scan turns lengths into
// offsets, EMIT writes.  One warp per record; output is staged in shared memory and leaves as aligned 16 B stores.
//
// Pure byte work: per record ~0.5 KB read, ~1 KB written; bounded by HBM bandwidth and, before that, by issue rate.
#include "agr_device.cuh"

#define K5_WARPS 8
#define K5_CHUNK 256u          // records per CTA (32 per warp)
#define K5_SB 1280u            // staging bytes per warp
#define K5_FLUSH 768u

// up to 32 literal bytes as four packed words, evaluated at compile time: no memory access, the lane picks its byte
__host__ __device__ constexpr unsigned long long jpack(const char* s, size_t n, size_t k) {
    unsigned long long w = 0;
    for (size_t i = 0; i < 8; ++i) if (8 * k + i < n) w |= (unsigned long long)(unsigned char)s[8 * k + i] << (8 * i);
    return w;
}
#define JLIT(W, S) do { constexpr unsigned long long _a = jpack(S, sizeof(S) - 1, 0), _b = jpack(S, sizeof(S) - 1, 1), \
                                                      _c = jpack(S, sizeof(S) - 1, 2), _d = jpack(S, sizeof(S) - 1, 3); \
                        static_assert(sizeof(S) - 1 <= 32, "literal too long"); \
                        (W).lit(_a, _b, _c, _d, (uint32_t)(sizeof(S) - 1)); } while (0)

namespace {

__device__ __forceinline__ uint32_t pick_byte(unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d, int i) {
    const unsigned long long w = (i & 16) ? ((i & 8) ? d : c) : ((i & 8) ? b : a);
    return (uint32_t)(w >> (8 * (i & 7))) & 0xffu;
}
// exclusive prefix of per-lane lengths that only take the values 0, 1, 2, 3, 6 — four ballots instead of a shuffle scan
__device__ __forceinline__ uint32_t len_scan(uint32_t el, int lane, uint32_t& total) {
    const uint32_t g1 = __ballot_sync(FULL, el >= 1u), g2 = __ballot_sync(FULL, el >= 2u), g3 = __ballot_sync(FULL, el >= 3u),
                   g6 = __ballot_sync(FULL, el == 6u);
    const uint32_t below = (1u << lane) - 1u;
    total = __popc(g1) + __popc(g2) + __popc(g3) + 3u * __popc(g6);
    return __popc(g1 & below) + __popc(g2 & below) + __popc(g3 & below) + 3u * __popc(g6 & below);
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}
__device__ __forceinline__ char hexc(uint32_t v) { return (char)(v < 10 ? '0' + v : 'a' + (v - 10)); }

// time.Time.MarshalJSON of time.Unix(0, ns).UTC(): RFC3339Nano, trailing zeros of the fraction dropped, no quotes.
// w0..w3 hold the 9-digit form "YYYY-MM-DDTHH:MM:SS.fffffffff"; the text is its first len-1 characters followed by 'Z'.
struct jtime { unsigned long long w0, w1, w2, w3; uint32_t len; };
__device__ __forceinline__ uint32_t time_len(unsigned long long ns) {
    uint32_t ft = (uint32_t)(ns % 1000000000ull), nd = 0;
    if (ft) { nd = 9; while (ft % 10u == 0u) { ft /= 10u; --nd; } }
    return 19u + (nd ? 1u + nd : 0u) + 1u;
}
__device__ __forceinline__ jtime fmt_time(unsigned long long ns) {
    const unsigned long long secs = ns / 1000000000ull;
    const uint32_t frac = (uint32_t)(ns - secs * 1000000000ull);
    const uint32_t days = (uint32_t)(secs / 86400ull);
    const uint32_t sod = (uint32_t)(secs - (unsigned long long)days * 86400ull);
    const uint32_t z = days + 719468u;                                // days since 0000-03-01 (civil-from-days)
    const uint32_t era = z / 146097u;
    const uint32_t doe = z - era * 146097u;
    const uint32_t yoe = (doe - doe / 1460u + doe / 36524u - doe / 146096u) / 365u;
    uint32_t y = yoe + era * 400u;
    const uint32_t doy = doe - (365u * yoe + yoe / 4u - yoe / 100u);
    const uint32_t mp = (5u * doy + 2u) / 153u;
    const uint32_t dd = doy - (153u * mp + 2u) / 5u + 1u;
    const uint32_t mm = mp < 10u ? mp + 3u : mp - 9u;
    if (mm <= 2u) ++y;
    uint32_t nd = 0, ft = frac;
    if (frac) { nd = 9; while (ft % 10u == 0u) { ft /= 10u; --nd; } }
    const uint32_t hh = sod / 3600u, mi = sod / 60u % 60u, ss = sod % 60u;
    const uint32_t fa = frac / 100000u, fb = frac % 100000u;          // 4 + 5 fraction digits
    jtime t;
#define D8(v, sh) ((unsigned long long)('0' + (v)) << (sh))
    t.w0 = D8(y / 1000u % 10u, 0) | D8(y / 100u % 10u, 8) | D8(y / 10u % 10u, 16) | D8(y % 10u, 24) |
           ((unsigned long long)'-' << 32) | D8(mm / 10u, 40) | D8(mm % 10u, 48) | ((unsigned long long)'-' << 56);
    t.w1 = D8(dd / 10u, 0) | D8(dd % 10u, 8) | ((unsigned long long)'T' << 16) | D8(hh / 10u, 24) | D8(hh % 10u, 32) |
           ((unsigned long long)':' << 40) | D8(mi / 10u, 48) | D8(mi % 10u, 56);
    t.w2 = ((unsigned long long)':' << 0) | D8(ss / 10u, 8) | D8(ss % 10u, 16) | ((unsigned long long)'.' << 24) |
           D8(fa / 1000u, 32) | D8(fa / 100u % 10u, 40) | D8(fa / 10u % 10u, 48) | D8(fa % 10u, 56);
    t.w3 = D8(fb / 10000u, 0) | D8(fb / 1000u % 10u, 8) | D8(fb / 100u % 10u, 16) | D8(fb / 10u % 10u, 24) | D8(fb % 10u, 32);
#undef D8
    t.len = 19u + (nd ? 1u + nd : 0u) + 1u;
    return t;
}
__device__ __forceinline__ jtime bcast_time(const jtime& t, int src) {
    jtime r;
    r.w0 = __shfl_sync(FULL, t.w0, src); r.w1 = __shfl_sync(FULL, t.w1, src);
    r.w2 = __shfl_sync(FULL, t.w2, src); r.w3 = __shfl_sync(FULL, t.w3, src);
    r.len = __shfl_sync(FULL, t.len, src);
    return r;
}

template <bool EMIT>
struct jwriter {
    uint8_t* sb;                 // this warp's staging buffer (16 B aligned)
    uint8_t* gout;
    unsigned long long gbase;    // global offset of sb[0]; 16-aligned
    uint32_t fill;               // bytes staged (including `head` foreign bytes of the first word)
    uint32_t head;
    unsigned long long total;    // MEASURE
    int lane;

    __device__ __forceinline__ void begin(uint8_t* out, unsigned long long off) {
        gout = out; gbase = off & ~15ull; head = (uint32_t)(off & 15ull); fill = head; total = 0;
    }
    __device__ __forceinline__ void flush(bool final) {
        __syncwarp();
        const uint32_t nwords = fill >> 4;
        uint32_t w0 = 0;
        if (head) {
            const uint32_t e = min(16u, fill);
            if ((uint32_t)lane >= head && (uint32_t)lane < e) gout[gbase + lane] = sb[lane];
            w0 = 1;
        }
        for (uint32_t w = w0 + lane; w < nwords; w += 32)
            *reinterpret_cast<uint4*>(gout + gbase + 16ull * w) = *reinterpret_cast<const uint4*>(sb + 16u * w);
        const uint32_t rem = fill & 15u;
        if (final) {
            if (nwords >= w0 && (uint32_t)lane < rem) gout[gbase + 16ull * nwords + lane] = sb[16u * nwords + lane];
        } else {
            uint8_t v = 0;
            if ((uint32_t)lane < rem) v = sb[16u * nwords + lane];
            __syncwarp();
            if ((uint32_t)lane < rem) sb[lane] = v;
            gbase += 16ull * nwords; fill = rem; head = 0;
            __syncwarp();
        }
    }
    // the caller has written k bytes at sb[fill ..).  Nothing flushes here: check() does, and is called at least every
    // K5_SB - K5_FLUSH staged bytes (every step of the variable-length loops, once per group of fixed fields)
    __device__ __forceinline__ void commit(uint32_t k) {
        if (EMIT) fill += k; else total += k;
    }
    __device__ __forceinline__ void check() {
        if (EMIT && fill >= K5_FLUSH) flush(false);
    }
    __device__ __forceinline__ void lit(unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d, uint32_t k) {
        if (EMIT && (uint32_t)lane < k) {                        // k is a compile-time constant at every call site
            uint32_t ch_;
            if (k <= 8) ch_ = (uint32_t)(a >> (8 * (lane & 7)));
            else if (k <= 16) ch_ = (uint32_t)(((lane & 8) ? b : a) >> (8 * (lane & 7)));
            else ch_ = pick_byte(a, b, c, d, lane);
            sb[fill + lane] = (uint8_t)ch_;
        }
        commit(k);
    }
    __device__ __forceinline__ void ch(char c) {
        if (EMIT && lane == 0) sb[fill] = (uint8_t)c;
        commit(1);
    }
    __device__ __forceinline__ void uint(uint32_t v) {           // v < 100000 (status codes, retry counters)
        if (v < 1000u) {                                         // the usual case: three digits at most
            const uint32_t d2 = v / 100u, r = v - 100u * d2, d1 = r / 10u, d0 = r - 10u * d1;
            const uint32_t nd = 1u + (v >= 10u) + (v >= 100u);
            if (EMIT && (uint32_t)lane < nd) {
                const uint32_t pos = nd - 1u - (uint32_t)lane;   // 0 = units
                sb[fill + lane] = (uint8_t)('0' + (pos == 0u ? d0 : pos == 1u ? d1 : d2));
            }
            commit(nd);
            return;
        }
        const uint32_t d4 = v / 10000u, d3 = v / 1000u % 10u, d2 = v / 100u % 10u, d1 = v / 10u % 10u, d0 = v % 10u;
        const uint32_t nd = v >= 10000u ? 5u : 4u;
        if (EMIT && (uint32_t)lane < nd) {
            const unsigned long long digs = (unsigned long long)d0 | ((unsigned long long)d1 << 8) | ((unsigned long long)d2 << 16) |
                                            ((unsigned long long)d3 << 24) | ((unsigned long long)d4 << 32);
            sb[fill + lane] = (uint8_t)('0' + ((digs >> (8 * (nd - 1 - lane))) & 0xffu));
        }
        commit(nd);
    }
    // uuid.UUID.String(): 8-4-4-4-12 lower-case hex
    __device__ __forceinline__ void uuid(unsigned long long lo, unsigned long long hi) {
        if (EMIT) {
            const int i = lane;                                      // nibble index
            const int b = i >> 1;
            const uint32_t byte = (uint32_t)((b < 8 ? lo >> (8 * b) : hi >> (8 * (b - 8))) & 0xffu);
            const uint32_t nib = (i & 1) ? (byte & 15u) : (byte >> 4);
            const int pos = i + (i >= 8) + (i >= 12) + (i >= 16) + (i >= 20);
            sb[fill + pos] = (uint8_t)hexc(nib);
            if (i < 4) sb[fill + 8 + 5 * i] = '-';
        }
        commit(36);
    }
    // a formatted time (jtime, below): lane i writes character i
    __device__ __forceinline__ void time(unsigned long long w0, unsigned long long w1, unsigned long long w2, unsigned long long w3, uint32_t len) {
        if (EMIT && (uint32_t)lane < len) {
            uint32_t c = pick_byte(w0, w1, w2, w3, lane);
            if ((uint32_t)lane == len - 1u) c = 'Z';
            sb[fill + lane] = (uint8_t)c;
        }
        commit(len);
    }
};

__device__ __forceinline__ uint32_t byte_at(const uint8_t* s, uint32_t len, long long p) {
    return (p >= 0 && p < (long long)len) ? s[p] : 0u;      // outside the string: a non-continuation byte
}
// utf8.DecodeRuneInString at a NON-continuation byte j: length of the valid sequence starting there, 0 if invalid
// (unicode/utf8 `first` / acceptRanges tables)
__device__ __forceinline__ uint32_t utf8_seq(const uint8_t* s, uint32_t len, long long j) {
    const uint32_t b0 = byte_at(s, len, j);
    if (b0 < 0x80u) return 1;
    uint32_t need, lo = 0x80u, hi = 0xbfu;
    if (b0 >= 0xc2u && b0 <= 0xdfu) need = 2;
    else if (b0 >= 0xe0u && b0 <= 0xefu) { need = 3; if (b0 == 0xe0u) lo = 0xa0u; else if (b0 == 0xedu) hi = 0x9fu; }
    else if (b0 >= 0xf0u && b0 <= 0xf4u) { need = 4; if (b0 == 0xf0u) lo = 0x90u; else if (b0 == 0xf4u) hi = 0x8fu; }
    else return 0;
    if (j + need > (long long)len) return 0;
    const uint32_t b1 = byte_at(s, len, j + 1);
    if (b1 < lo || b1 > hi) return 0;
    for (uint32_t k = 2; k < need; ++k) { const uint32_t b = byte_at(s, len, j + k); if (b < 0x80u || b > 0xbfu) return 0; }
    return need;
}

// encoding/json encodeState.string (escapeHTML = true) for the byte at position p of the string s[0..len).
// Returns how many output bytes this position contributes (0..6) and packs them little-endian into `chars`.
// `touched`: the record has been through json.Unmarshal + Marshal (StoreResponse / MarkRequestFailed), which turns an
// invalid byte into a REAL U+FFFD that the second Marshal copies through as EF BF BD instead of writing "�".
__device__ __forceinline__ uint32_t esc_byte(const uint8_t* s, uint32_t len, uint32_t p, uint32_t b, bool touched,
                                             unsigned long long& chars) {
    if (b < 0x80u) {
        if (b >= 0x20u && b != '"' && b != '\\' && b != '<' && b != '>' && b != '&') { chars = b; return 1; }
        char e = 0;
        switch (b) {
            case '"': e = '"'; break;   case '\\': e = '\\'; break;
            case '\n': e = 'n'; break;  case '\r': e = 'r'; break;  case '\t': e = 't'; break;
            case '\b': e = 'b'; break;  case '\f': e = 'f'; break;          // Go >= 1.22 (go.mod: go 1.23)
        }
        if (e) { chars = (unsigned long long)'\\' | ((unsigned long long)(uint8_t)e << 8); return 2; }
        chars = (unsigned long long)'\\' | ((unsigned long long)'u' << 8) | ((unsigned long long)'0' << 16) |
                ((unsigned long long)'0' << 24) | ((unsigned long long)(uint8_t)hexc(b >> 4) << 32) |
                ((unsigned long long)(uint8_t)hexc(b & 15u) << 40);
        return 6;
    }
    // multi-byte territory: find the sequence this byte belongs to
    long long j = p;
    if (b <= 0xbfu) {                                        // continuation byte: nearest non-continuation byte before it
        j = -1;
        for (int k = 1; k <= 3; ++k) {
            const long long q = (long long)p - k;
            if (q < 0) break;
            const uint32_t c = s[q];
            if (c < 0x80u || c > 0xbfu) { j = q; break; }
        }
    }
    uint32_t sl = 0;
    if (j >= 0) { sl = utf8_seq(s, len, j); if ((long long)p - j >= (long long)sl) sl = 0; }
    if (sl == 0) {                                           // RuneError, width 1
        if (touched) { chars = 0xefull | (0xbfull << 8) | (0xbdull << 16); return 3; }
        chars = (unsigned long long)'\\' | ((unsigned long long)'u' << 8) | ((unsigned long long)'f' << 16) |
                ((unsigned long long)'f' << 24) | ((unsigned long long)'f' << 32) | ((unsigned long long)'d' << 40);
        return 6;
    }
    // U+2028 / U+2029 are escaped (JSONP safety): E2 80 A8 / E2 80 A9
    if (sl == 3 && s[j] == 0xe2u && s[j + 1] == 0x80u && (s[j + 2] == 0xa8u || s[j + 2] == 0xa9u)) {
        if ((long long)p != j) return 0;
        chars = (unsigned long long)'\\' | ((unsigned long long)'u' << 8) | ((unsigned long long)'2' << 16) |
                ((unsigned long long)'0' << 24) | ((unsigned long long)'2' << 32) |
                ((unsigned long long)(s[j + 2] == 0xa8u ? '8' : '9') << 40);
        return 6;
    }
    chars = b;
    return 1;
}

template <bool EMIT>
__device__ __forceinline__ void put_packed(jwriter<EMIT>& w, uint32_t el, unsigned long long chars) {
    uint32_t tot;
    const uint32_t off = len_scan(el, w.lane, tot);
    if (EMIT) {
        uint8_t* dst = w.sb + w.fill + off;
        if (el >= 1u) dst[0] = (uint8_t)chars;
        if (el >= 2u) dst[1] = (uint8_t)(chars >> 8);
        if (el >= 3u) dst[2] = (uint8_t)(chars >> 16);
        if (el == 6u) { dst[3] = (uint8_t)(chars >> 24); dst[4] = (uint8_t)(chars >> 32); dst[5] = (uint8_t)(chars >> 40); }
    }
    w.commit(tot);
    w.check();
}

// a JSON string body (no quotes) from s[0..len)
template <bool EMIT>
__device__ __forceinline__ void put_escaped(jwriter<EMIT>& w, const uint8_t* s, uint32_t len, bool touched) {
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t p = base + w.lane;
        const uint32_t b = p < len ? s[p] : 0x20u;
        const bool special = b < 0x20u || b >= 0x80u || b == '"' || b == '\\' || b == '<' || b == '>' || b == '&';
        if (!__any_sync(FULL, special)) {                     // the common step: bytes copy through
            const uint32_t k = min(32u, len - base);
            if (EMIT && p < len) w.sb[w.fill + w.lane] = (uint8_t)b;
            w.commit(k);
            w.check();
            continue;
        }
        unsigned long long chars = 0; uint32_t el = 0;
        if (p < len) el = esc_byte(s, len, p, b, touched, chars);
        put_packed(w, el, chars);
    }
}

// map[string]string from its flattened form "Key: Value\n"... (keys already in sorted order, Q5): {"k":"v","k2":"v2"}
template <bool EMIT>
__device__ __forceinline__ void put_headers(jwriter<EMIT>& w, const uint8_t* s, uint32_t len, bool touched) {
    if (len == 0) { JLIT(w, "{}"); return; }
    JLIT(w, "{\"");
    bool in_val = false, prev_sep = false;                   // carried across 32-byte steps (warp-uniform)
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t p = base + w.lane;
        const uint32_t b = p < len ? s[p] : 0u;
        const uint32_t nl = __ballot_sync(FULL, p < len && b == '\n');
        const uint32_t co = __ballot_sync(FULL, p < len && b == ':');
        const uint32_t below = (1u << w.lane) - 1u;
        const uint32_t prev_nl = nl & below;
        const uint32_t start = prev_nl ? 32u - __clz(prev_nl) : 0u;
        const uint32_t same_line = below & ~((start >= 32u) ? 0xffffffffu : ((1u << start) - 1u));
        const bool inval = (co & same_line) != 0u || (prev_nl == 0u && in_val);
        const bool is_sep = p < len && b == ':' && !inval;
        const uint32_t sepm = __ballot_sync(FULL, is_sep);
        const bool after_sep = w.lane ? ((sepm >> (w.lane - 1)) & 1u) : prev_sep;
        unsigned long long chars = 0; uint32_t el = 0;
        if (p < len) {
            if (is_sep) { chars = (unsigned long long)'"' | ((unsigned long long)':' << 8) | ((unsigned long long)'"' << 16); el = 3; }
            else if (b == ' ' && after_sep) el = 0;
            else if (b == '\n') {
                if (p == len - 1) { chars = '"'; el = 1; }
                else { chars = (unsigned long long)'"' | ((unsigned long long)',' << 8) | ((unsigned long long)'"' << 16); el = 3; }
            } else el = esc_byte(s, len, p, b, touched, chars);
        }
        put_packed(w, el, chars);
        // carry: state after the last byte of this step
        const uint32_t last_nl = nl ? 32u - __clz(nl) : 0u;            // index after the last newline in the step
        const uint32_t tail = (last_nl >= 32u) ? 0u : ~((1u << last_nl) - 1u);
        in_val = (co & tail) != 0u || (nl == 0u && in_val);
        prev_sep = (sepm >> 31) & 1u;
    }
    w.ch('}');
}

// []byte -> base64.StdEncoding (padded), no quotes
template <bool EMIT>
__device__ __forceinline__ void put_base64(jwriter<EMIT>& w, const uint8_t* s, uint32_t len) {
    const uint32_t groups = (len + 2u) / 3u;
    if (!EMIT) w.commit(4u * groups);
    else for (uint32_t g0 = 0; g0 < groups; g0 += 32) {
        const uint32_t g = g0 + w.lane;
        if (g < groups) {
            const uint32_t p = 3u * g;
            const uint32_t b0 = s[p], b1 = p + 1 < len ? s[p + 1] : 0u, b2 = p + 2 < len ? s[p + 2] : 0u;
            const uint32_t v = (b0 << 16) | (b1 << 8) | b2;
            // the four 6-bit values, one per byte, mapped to their characters together (SWAR): a byte holds x < 64, so
            // x + (128 - t) sets bit 7 exactly when x >= t and never carries into the next byte; the character is
            // x + 'A', + 6 from 'a' on, - 75 from '0' on, - 15 for '+', + 3 for '/' — no byte borrows at any step.
            const uint32_t X = ((v >> 18) & 63u) | (((v >> 12) & 63u) << 8) | (((v >> 6) & 63u) << 16) | ((v & 63u) << 24);
            const uint32_t ge26 = ((X + 0x66666666u) >> 7) & 0x01010101u, ge52 = ((X + 0x4c4c4c4cu) >> 7) & 0x01010101u,
                           ge62 = ((X + 0x42424242u) >> 7) & 0x01010101u, ge63 = ((X + 0x41414141u) >> 7) & 0x01010101u;
            uint32_t out = X + 0x41414141u + 6u * ge26 + 3u * ge63 - 75u * ge52 - 15u * ge62;
            if (p + 1 >= len) out = (out & 0x0000ffffu) | ((uint32_t)'=' << 16) | ((uint32_t)'=' << 24);
            else if (p + 2 >= len) out = (out & 0x00ffffffu) | ((uint32_t)'=' << 24);
            uint8_t* dst = w.sb + w.fill + 4u * w.lane;
            dst[0] = (uint8_t)out; dst[1] = (uint8_t)(out >> 8); dst[2] = (uint8_t)(out >> 16); dst[3] = (uint8_t)(out >> 24);
        }
        w.commit(4u * min(32u, groups - g0));
        w.check();
    }
}

__device__ __forceinline__ uint32_t cstr_len32(const uint8_t* s, int lane) {      // strnlen(s, 32), one byte per lane
    const uint32_t z = __ballot_sync(FULL, s[lane] == 0);
    return z ? (uint32_t)__ffs(z) - 1u : 32u;
}

// json.Marshal(requests.Request) for the record in row `rid`
template <bool EMIT>
__device__ __forceinline__ void encode_record(jwriter<EMIT>& w, const agr_dev& d, const agr_k5_params& p, uint32_t rid,
                                              const uint8_t* rec /* header + payload, generic pointer */) {
    const uint32_t st = d.state[rid];
    if (!(st & ST_STORED)) { JLIT(w, "null"); return; }
    const uint32_t aux = d.aux[rid];
    const bool responded = (st & ST_RESPONDED) != 0u;
    const uint32_t retry = st_retry(st);
    const bool touched = responded || retry != 0u || p.roundtrip != 0u;
    unsigned long long lo, hi;
    if (d.cfg_flags & AGR_CFG_MINT_IDS) agr_mint_id(row_logical(d, rid), d.shard_id, d.id_gen, d.id_secret, lo, hi);
    else { lo = *reinterpret_cast<const unsigned long long*>(rec); hi = *reinterpret_cast<const unsigned long long*>(rec + 8); }
    const unsigned long long seq = *reinterpret_cast<const unsigned long long*>(rec + AGR_OFF_SEQ);
    const uint32_t flags = *reinterpret_cast<const uint32_t*>(rec + AGR_OFF_FLAGS);
    const uint32_t path_len = *reinterpret_cast<const uint16_t*>(rec + AGR_OFF_PATH_LEN);
    const uint32_t hdr_len = *reinterpret_cast<const uint16_t*>(rec + AGR_OFF_HDR_LEN);
    const uint32_t body_len = *reinterpret_cast<const uint32_t*>(rec + AGR_OFF_BODY_LEN);
    const uint8_t* pay = rec + AGR_OFF_PAYLOAD;

    JLIT(w, "{\"id\":\"");
    w.uuid(lo, hi);
    JLIT(w, "\",\"agent_id\":\"");
    w.check();
    put_escaped(w, rec + AGR_OFF_AGENT_ID, cstr_len32(rec + AGR_OFF_AGENT_ID, w.lane), touched);
    JLIT(w, "\",\"method\":\"");
    switch ((flags & AGR_F_METHOD_MASK) >> AGR_F_METHOD_SHIFT) {
        case AGR_M_GET: JLIT(w, "GET"); break;        case AGR_M_POST: JLIT(w, "POST"); break;
        case AGR_M_PUT: JLIT(w, "PUT"); break;        case AGR_M_DELETE: JLIT(w, "DELETE"); break;
        case AGR_M_PATCH: JLIT(w, "PATCH"); break;    case AGR_M_HEAD: JLIT(w, "HEAD"); break;
        case AGR_M_OPTIONS: JLIT(w, "OPTIONS"); break;
    }
    JLIT(w, "\",\"path\":\"");
    put_escaped(w, pay, path_len, touched);
    JLIT(w, "\",\"headers\":");
    put_headers(w, pay + path_len, hdr_len, touched);
    JLIT(w, ",\"body\":\"");
    put_base64(w, pay + path_len + hdr_len, body_len);
    JLIT(w, "\",\"status\":\"");
    switch (st_status(st)) {
        case AGR_ST_PENDING: JLIT(w, "pending"); break;      case AGR_ST_PROCESSING: JLIT(w, "processing"); break;
        case AGR_ST_COMPLETED: JLIT(w, "completed"); break;  case AGR_ST_FAILED: JLIT(w, "failed"); break;
    }
    JLIT(w, "\",\"retry_count\":");
    w.uint(retry);
    JLIT(w, ",\"max_retries\":");
    w.uint(st_max(st));
    JLIT(w, ",\"created_at\":\"");
    // the two timestamps of a record are formatted side by side (odd lanes: processed_at, even lanes: created_at)
    const unsigned long long pt = responded ? p.ptime[rid] : 0ull;
    jtime tc{}, tp{};
    if (EMIT) {
        const jtime mine = fmt_time((w.lane & 1) ? pt : seq);
        tc = bcast_time(mine, 0); tp = bcast_time(mine, 1);
    } else { tc.len = time_len(seq); tp.len = responded ? time_len(pt) : 0u; }
    w.time(tc.w0, tc.w1, tc.w2, tc.w3, tc.len);
    w.ch('"');
    w.check();
    if (responded) {                                         // requests.go:165-167
        JLIT(w, ",\"processed_at\":\"");
        w.time(tp.w0, tp.w1, tp.w2, tp.w3, tp.len);
        JLIT(w, "\",\"response\":{\"status_code\":");
        w.uint(aux & 0xffffu);
        JLIT(w, ",\"headers\":");
        w.check();
        const uint32_t rl = p.resp_len[rid], rh = min(p.resp_hlen[rid], rl);
        const uint8_t* rb = p.bytes + p.resp_off[rid];
        put_headers(w, rb, rh, (st & ST_RESP_RT) != 0u || p.roundtrip != 0u);
        JLIT(w, ",\"body\":\"");
        put_base64(w, rb + rh, rl - rh);
        JLIT(w, "\",\"received_at\":\"");
        w.time(tp.w0, tp.w1, tp.w2, tp.w3, tp.len);
        JLIT(w, "\"}");
        w.check();
    }
    if (retry != 0u) {                                       // requests.go:244 request.Error = err.Error()
        JLIT(w, ",\"error\":\"");
        const uint32_t el = p.err_len[rid];
        if (el) put_escaped(w, p.bytes + p.err_off[rid], el, (st & ST_ERR_RT) != 0u || p.roundtrip != 0u);
        else JLIT(w, "transport error");
        w.ch('"');
    }
    w.ch('}');
}

template <bool EMIT>
__global__ void __launch_bounds__(K5_WARPS * 32, EMIT ? 3 : 6) k5_json(const agr_dev d, const agr_k5_params p) {
    __shared__ __align__(16) uint8_t s_stage[EMIT ? K5_WARPS * K5_SB : 16];
    __shared__ __align__(16) uint4 s_rec[K5_WARPS][32];
    __shared__ unsigned long long s_off[K5_CHUNK];
    __shared__ unsigned long long s_wsum[K5_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk0 = blockIdx.x * K5_CHUNK;
    const uint32_t in_chunk = min(K5_CHUNK, p.n - chunk0);
    if (EMIT) {
        // exclusive scan of this chunk's record lengths on top of the chunk's base
        const uint32_t t = threadIdx.x;
        const uint32_t v = t < in_chunk ? p.len[chunk0 + t] : 0u;
        uint32_t tot;
        const uint32_t ex = warp_excl_scan(v, lane, tot);
        if (lane == 0) s_wsum[warp] = tot;
        __syncthreads();
        unsigned long long base = p.chunk_sum[blockIdx.x];
        for (int k = 0; k < warp; ++k) base += s_wsum[k];
        s_off[t] = base + ex;
        if (t < in_chunk) p.off[chunk0 + t] = base + ex;
        if (chunk0 + t + 1 == p.n) p.off[p.n] = base + ex + v;
        __syncthreads();
    }
    jwriter<EMIT> w;
    w.lane = lane;
    w.sb = EMIT ? s_stage + warp * K5_SB : nullptr;
    unsigned long long wsum = 0;
    for (uint32_t k = 0; k < 32; ++k) {
        const uint32_t r = k * K5_WARPS + warp;              // neighbouring warps write neighbouring output
        if (r >= in_chunk) break;
        const uint32_t i = chunk0 + r;
        const uint32_t rid = p.rids ? p.rids[i] : row_physical(d, p.first_l + i);
        const uint8_t* src = rec_ptr(d, rid);
        const uint8_t* rec = src;
        if (!d.voff) {                                       // fixed 512 B rows: stage the record in shared memory
            __syncwarp();
            s_rec[warp][lane] = ldg_nc_v4(src + lane * 16);
            __syncwarp();
            rec = reinterpret_cast<const uint8_t*>(&s_rec[warp][0]);
        }
        if (EMIT) w.begin(p.out, s_off[r]); else w.total = 0;
        if (p.array) w.ch(i == 0 ? '[' : ',');
        encode_record<EMIT>(w, d, p, rid, rec);
        if (p.array && i + 1 == p.n) w.ch(']');
        if (EMIT) w.flush(true);
        else { if (lane == 0) p.len[i] = (uint32_t)w.total; wsum += w.total; }
    }
    if (!EMIT) {
        if (lane == 0) s_wsum[warp] = wsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long s = 0;
            for (int k = 0; k < K5_WARPS; ++k) s += s_wsum[k];
            p.chunk_sum[blockIdx.x] = s;
        }
    }
}

// exclusive scan of the chunk sums in place; chunk_sum[nchunks] = grand total
__global__ void __launch_bounds__(1024) k5_scan(unsigned long long* chunk_sum, const uint32_t nchunks) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nchunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < nchunks ? chunk_sum[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (int k = 0; k < warp; ++k) pre += s_w[k];
        if (i < nchunks) chunk_sum[i] = pre + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sum[nchunks] = s_carry;
}

}  // namespace

uint32_t agr_k5_chunks(uint32_t n) { return (n + K5_CHUNK - 1) / K5_CHUNK; }
void agr_launch_k5_measure(const agr_dev& d, const agr_k5_params& p, cudaStream_t st) {
    if (!p.n) return;
    const uint32_t nch = agr_k5_chunks(p.n);
    k5_json<false><<<nch, K5_WARPS * 32, 0, st>>>(d, p);
    k5_scan<<<1, 1024, 0, st>>>(p.chunk_sum, nch);
}
void agr_launch_k5_emit(const agr_dev& d, const agr_k5_params& p, cudaStream_t st) {
    if (!p.n) return;
    k5_json<true><<<agr_k5_chunks(p.n), K5_WARPS * 32, 0, st>>>(d, p);
}
