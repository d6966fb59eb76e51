/*
 * oracle/cpu_ref.c — C restatement of Agentainer's request persistence / replay / proxy-decision path.
 *
 * TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline).  Never linked into or called by the product library.
 * PARITY UNPINNED: the reference ships no tests / golden vectors for this path and cannot be built here (no Go, no
 * Redis); this file follows the reference source line by line and is cross-checked against oracle/model.py and the
 * hand-derived KATs (tests/test_cpu_ref.py).
 *
 * What is restated (paths relative to the reference tree):
 *   internal/requests/requests.go:64-117   StoreRequest      -> cref_ingest (store part)
 *   internal/requests/requests.go:120-194  StoreResponse     -> cref_complete kind RESPONSE
 *   internal/requests/requests.go:197-225  GetPendingRequests-> cref_pending / cref_scan
 *   internal/requests/requests.go:228-275  MarkRequestFailed -> cref_complete kind ERROR
 *   internal/api/server.go:493-541         proxyToAgentHandler decision -> cref_ingest
 *   internal/api/server.go:583-615         interceptTransport.RoundTrip classification -> cref_complete kinds
 *   internal/requests/replay_worker.go:58-117,166-199  processAgents / isAgentRunning / skip rule -> cref_scan
 *   internal/agent/agent.go:372-390,343-367  GetAgent / Remove cleanup -> agent JSON docs, cref_drop_agent
 * Third-party semantics restated: Redis 7 (redis:7-alpine, docker-compose.yml:8) strings / lists / KEYS / DEL with
 * one keyspace; encoding/json of requests.Request (field order = struct order, map keys sorted, []byte -> base64,
 * HTML-escaped strings).  Records are kept as JSON text and marshalled / unmarshalled at exactly the points where
 * the Go code does (2 encodes + 2 decodes per proxied request, 1 agent-document decode per request), so the timed
 * baseline carries the reference's CPU costs; RESP/TCP round trips and HTTP parsing are NOT modelled, which makes
 * this an upper bound on the reference's speed.
 *
 * The call surface mirrors include/agentainer_gpu.h (same structs) so that one scenario driver exercises both.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/agentainer_gpu.h"

/* ------------------------------------------------------------------ mini Redis keyspace */
enum { T_STRING = 1, T_LIST = 2 };
typedef struct { char id[37]; } lid;           /* uuid text */
typedef struct {
    char* key; uint32_t klen; uint64_t hash; int type;
    char* val; uint32_t vlen, vcap;            /* T_STRING (val lives in the arena: never freed, reused in place) */
    lid* items; uint32_t n, cap;               /* T_LIST */
    uint64_t expires;                          /* SET ... EX: the server's clock value at which the key is gone; 0 = none */
} kent;
/* bump arena per keyspace: one Redis instance = one allocator, so shards on different threads never contend */
typedef struct achunk { struct achunk* next; size_t used, cap; } achunk;
typedef struct { kent* e; uint64_t cap, used, tomb; achunk* arena; uint64_t now; } keyspace;   /* now: the server's clock (set by the stream) */
static char* arena_alloc(keyspace* k, size_t n) {
    n = (n + 15) & ~(size_t)15;
    if (!k->arena || k->arena->used + n > k->arena->cap) {
        size_t cap = n > ((size_t)8 << 20) ? n : ((size_t)8 << 20);
        achunk* c = (achunk*)malloc(sizeof(achunk) + cap);
        c->next = k->arena; c->used = 0; c->cap = cap; k->arena = c;
    }
    char* p = (char*)(k->arena + 1) + k->arena->used;
    k->arena->used += n;
    return p;
}

static uint64_t fnv(const char* s, uint32_t n) { uint64_t h = 0xcbf29ce484222325ULL; for (uint32_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ULL; } return h; }
static void ks_init(keyspace* k, uint64_t cap) { k->cap = cap; k->used = k->tomb = 0; k->arena = NULL; k->now = 0; k->e = (kent*)calloc(cap, sizeof(kent)); }
static kent* ks_find(keyspace* k, const char* key, uint32_t klen, int create);
static void ks_grow(keyspace* k) {
    keyspace n; ks_init(&n, k->cap * 2); n.arena = k->arena; n.now = k->now;
    for (uint64_t i = 0; i < k->cap; ++i) if (k->e[i].key && k->e[i].type) {
        uint64_t j = k->e[i].hash & (n.cap - 1);
        while (n.e[j].key) j = (j + 1) & (n.cap - 1);
        n.e[j] = k->e[i]; n.used++;
    }
    free(k->e); *k = n;
}
static kent* ks_find(keyspace* k, const char* key, uint32_t klen, int create) {
    if (create && (k->used + k->tomb + 1) * 10 > k->cap * 7) ks_grow(k);
    uint64_t h = fnv(key, klen), j = h & (k->cap - 1);
    kent* firsttomb = NULL;
    for (;;) {
        kent* e = &k->e[j];
        if (!e->key) {
            if (!create) return NULL;
            if (firsttomb) { e = firsttomb; k->tomb--; }
            e->key = arena_alloc(k, klen + 1); memcpy(e->key, key, klen); e->key[klen] = 0;
            e->klen = klen; e->hash = h; e->type = 0; e->val = NULL; e->vlen = e->vcap = 0; e->items = NULL; e->n = e->cap = 0; e->expires = 0;
            k->used++;
            return e;
        }
        if (e->type == 0) { if (!firsttomb) firsttomb = e; }
        else if (e->hash == h && e->klen == klen && memcmp(e->key, key, klen) == 0) return e;
        j = (j + 1) & (k->cap - 1);
    }
}
static void ks_del_entry(keyspace* k, kent* e) {          /* DEL */
    if (!e || !e->type) return;
    free(e->items); e->val = NULL; e->items = NULL; e->n = e->cap = 0; e->vlen = e->vcap = 0;
    e->type = 0; e->expires = 0; k->used--; k->tomb++;
}
#define TTL_24H (24ULL * 3600ULL * 1000000000ULL)           /* requests.go:106,175,270; the stream's clock is in nanoseconds */
static void r_set_ex(keyspace* k, const char* key, uint32_t klen, const char* v, uint32_t vlen, uint64_t ttl) {   /* SET key val [EX ttl] */
    kent* e = ks_find(k, key, klen, 1);
    e->expires = ttl ? k->now + ttl : 0;                      /* every SET restarts the TTL (Q11) */
    if (e->type == T_LIST) { free(e->items); e->items = NULL; e->val = NULL; e->vcap = 0; }
    if (e->val == NULL || vlen + 1 > e->vcap) {              /* room for the response that StoreResponse adds later */
        e->vcap = vlen + vlen / 2 + 128;
        e->val = arena_alloc(k, e->vcap);
    }
    memcpy(e->val, v, vlen); e->val[vlen] = 0; e->vlen = vlen; e->type = T_STRING;
}
static void r_set(keyspace* k, const char* key, uint32_t klen, const char* v, uint32_t vlen) { r_set_ex(k, key, klen, v, vlen, TTL_24H); }   /* the record keys */
static kent* r_get(keyspace* k, const char* key, uint32_t klen) {
    kent* e = ks_find(k, key, klen, 0);
    if (e && e->type == T_STRING && e->expires && k->now >= e->expires) { ks_del_entry(k, e); return NULL; }   /* expired keys are gone */
    return (e && e->type == T_STRING) ? e : NULL;
}
static void r_rpush(keyspace* k, const char* key, uint32_t klen, const char* id) {
    kent* e = ks_find(k, key, klen, 1);
    if (e->type != T_LIST) { e->val = NULL; e->vcap = 0; e->type = T_LIST; e->n = 0; }
    if (e->n == e->cap) { e->cap = e->cap ? e->cap * 2 : 8; e->items = (lid*)realloc(e->items, e->cap * sizeof(lid)); }
    memcpy(e->items[e->n].id, id, 36); e->items[e->n].id[36] = 0; e->n++;
}
static int r_lrem1(keyspace* k, const char* key, uint32_t klen, const char* id) {   /* LREM key 1 id: head -> tail, O(len) */
    kent* e = ks_find(k, key, klen, 0);
    if (!e || e->type != T_LIST) return 0;
    for (uint32_t i = 0; i < e->n; ++i) if (memcmp(e->items[i].id, id, 36) == 0) {
        memmove(&e->items[i], &e->items[i + 1], (size_t)(e->n - i - 1) * sizeof(lid));
        if (--e->n == 0) ks_del_entry(k, e);               /* empty lists do not exist as keys */
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ JSON / base64 of requests.Request */
typedef struct { char* p; size_t n, cap; } sbuf;
static void sb_need(sbuf* b, size_t m) { if (b->n + m > b->cap) { b->cap = (b->n + m) * 2 + 64; b->p = (char*)realloc(b->p, b->cap); } }
static void sb_put(sbuf* b, const char* s, size_t m) { sb_need(b, m); memcpy(b->p + b->n, s, m); b->n += m; }
static void sb_str(sbuf* b, const char* s) { sb_put(b, s, strlen(s)); }
/* unicode/utf8.DecodeRune at s[i..m): width of the valid sequence starting there, 0 if the byte starts none */
static int utf8_width(const unsigned char* s, size_t i, size_t m) {
    unsigned char c = s[i]; int need; unsigned lo = 0x80, hi = 0xbf;
    if (c < 0x80) return 1;
    if (c >= 0xc2 && c <= 0xdf) need = 2;
    else if (c >= 0xe0 && c <= 0xef) { need = 3; if (c == 0xe0) lo = 0xa0; else if (c == 0xed) hi = 0x9f; }
    else if (c >= 0xf0 && c <= 0xf4) { need = 4; if (c == 0xf0) lo = 0x90; else if (c == 0xf4) hi = 0x8f; }
    else return 0;
    if (i + (size_t)need > m) return 0;
    if (s[i + 1] < lo || s[i + 1] > hi) return 0;
    for (int k = 2; k < need; ++k) if (s[i + k] < 0x80 || s[i + k] > 0xbf) return 0;
    return need;
}
static void sb_jstr(sbuf* b, const char* s_, size_t m) {   /* encoding/json encodeState.string, escapeHTML = true (Go 1.23) */
    static const char hex[] = "0123456789abcdef";
    const unsigned char* s = (const unsigned char*)s_;
    sb_need(b, m * 6 + 2); b->p[b->n++] = '"';
    for (size_t i = 0; i < m;) {
        unsigned char c = s[i];
        if (c < 0x80) {
            if (c == '"' || c == '\\') { b->p[b->n++] = '\\'; b->p[b->n++] = (char)c; }
            else if (c == '\n') { b->p[b->n++] = '\\'; b->p[b->n++] = 'n'; }
            else if (c == '\r') { b->p[b->n++] = '\\'; b->p[b->n++] = 'r'; }
            else if (c == '\t') { b->p[b->n++] = '\\'; b->p[b->n++] = 't'; }
            else if (c == '\b') { b->p[b->n++] = '\\'; b->p[b->n++] = 'b'; }
            else if (c == '\f') { b->p[b->n++] = '\\'; b->p[b->n++] = 'f'; }
            else if (c < 0x20 || c == '<' || c == '>' || c == '&') { memcpy(b->p + b->n, "\\u00", 4); b->n += 4; b->p[b->n++] = hex[c >> 4]; b->p[b->n++] = hex[c & 15]; }
            else b->p[b->n++] = (char)c;
            i++;
            continue;
        }
        int w = utf8_width(s, i, m);
        if (w == 0) { memcpy(b->p + b->n, "\\ufffd", 6); b->n += 6; i++; continue; }      /* RuneError, width 1 */
        if (w == 3 && c == 0xe2 && s[i + 1] == 0x80 && (s[i + 2] == 0xa8 || s[i + 2] == 0xa9)) {
            memcpy(b->p + b->n, s[i + 2] == 0xa8 ? "\\u2028" : "\\u2029", 6); b->n += 6; i += 3; continue;
        }
        memcpy(b->p + b->n, s + i, (size_t)w); b->n += (size_t)w; i += (size_t)w;
    }
    b->p[b->n++] = '"';
}
static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
static void sb_b64(sbuf* b, const uint8_t* s, size_t m) {
    sb_need(b, (m + 2) / 3 * 4 + 2); b->p[b->n++] = '"';
    size_t i = 0;
    for (; i + 2 < m; i += 3) { uint32_t v = (s[i] << 16) | (s[i + 1] << 8) | s[i + 2]; b->p[b->n++] = B64[v >> 18]; b->p[b->n++] = B64[(v >> 12) & 63]; b->p[b->n++] = B64[(v >> 6) & 63]; b->p[b->n++] = B64[v & 63]; }
    if (i + 1 == m) { uint32_t v = s[i] << 16; b->p[b->n++] = B64[v >> 18]; b->p[b->n++] = B64[(v >> 12) & 63]; b->p[b->n++] = '='; b->p[b->n++] = '='; }
    else if (i + 2 == m) { uint32_t v = (s[i] << 16) | (s[i + 1] << 8); b->p[b->n++] = B64[v >> 18]; b->p[b->n++] = B64[(v >> 12) & 63]; b->p[b->n++] = B64[(v >> 6) & 63]; b->p[b->n++] = '='; }
    b->p[b->n++] = '"';
}
static void uuid_text(const uint8_t id[16], char out[37]) {
    static const char hex[] = "0123456789abcdef"; int o = 0;
    for (int i = 0; i < 16; ++i) { if (i == 4 || i == 6 || i == 8 || i == 10) out[o++] = '-'; out[o++] = hex[id[i] >> 4]; out[o++] = hex[id[i] & 15]; }
    out[36] = 0;
}
static int hexv(char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }
static void uuid_parse(const char* t, uint8_t id[16]) { int o = 0; for (int i = 0; i < 16; ++i) { if (t[o] == '-') o++; id[i] = (uint8_t)((hexv(t[o]) << 4) | hexv(t[o + 1])); o += 2; } }
static char* put2(char* t, unsigned v) { t[0] = (char)('0' + v / 10); t[1] = (char)('0' + v % 10); return t + 2; }
static void sb_time(sbuf* b, uint64_t ns) {                 /* time.Unix(0, ns).UTC().MarshalJSON(): RFC3339Nano */
    uint64_t secs = ns / 1000000000ULL; unsigned frac = (unsigned)(ns % 1000000000ULL);
    uint64_t days = secs / 86400ULL; unsigned sod = (unsigned)(secs % 86400ULL);
    int64_t z = (int64_t)days + 719468; int64_t era = z / 146097; unsigned doe = (unsigned)(z - era * 146097);
    unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365; unsigned y = yoe + (unsigned)era * 400;
    unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100); unsigned mp = (5 * doy + 2) / 153;
    unsigned d = doy - (153 * mp + 2) / 5 + 1; unsigned mth = mp < 10 ? mp + 3 : mp - 9; if (mth <= 2) y++;
    char t[48]; char* q = t;
    *q++ = '"'; q = put2(q, y / 100); q = put2(q, y % 100); *q++ = '-'; q = put2(q, mth); *q++ = '-'; q = put2(q, d); *q++ = 'T';
    q = put2(q, sod / 3600); *q++ = ':'; q = put2(q, sod / 60 % 60); *q++ = ':'; q = put2(q, sod % 60);
    if (frac) {
        char f[9]; unsigned v = frac; for (int k = 8; k >= 0; --k) { f[k] = (char)('0' + v % 10); v /= 10; }
        int k = 9; while (k > 0 && f[k - 1] == '0') k--;
        *q++ = '.'; memcpy(q, f, (size_t)k); q += k;
    }
    *q++ = 'Z'; *q++ = '"';
    sb_put(b, t, (size_t)(q - t));
}
static const char* status_name(int s) { return s == AGR_ST_PENDING ? "pending" : s == AGR_ST_PROCESSING ? "processing" : s == AGR_ST_COMPLETED ? "completed" : s == AGR_ST_FAILED ? "failed" : ""; }
static const char* method_name(uint32_t f) { static const char* m[] = {"", "GET", "POST", "PUT", "DELETE", "PATCH", "HEAD", "OPTIONS"}; uint32_t c = (f & AGR_F_METHOD_MASK) >> AGR_F_METHOD_SHIFT; return c < 8 ? m[c] : ""; }
static uint32_t method_code(const char* s, size_t n) { static const char* m[] = {"", "GET", "POST", "PUT", "DELETE", "PATCH", "HEAD", "OPTIONS"}; for (uint32_t c = 1; c < 8; ++c) if (strlen(m[c]) == n && memcmp(m[c], s, n) == 0) return c; return 0; }

/* the decoded form (requests.Request) */
typedef struct {
    agr_record rec;            /* binary fields incl. payload (path | headers | body) */
    int has_response; uint16_t resp_status; uint64_t processed_at; uint64_t received_at;
    char error[64];
} reqdoc;

/* json.Marshal(request) — requests.go:101,170,265 */
static void marshal_request(sbuf* b, const reqdoc* d) {
    const agr_record* r = &d->rec; char idt[37];
    b->n = 0;
    uuid_text(r->request_id, idt);
    sb_str(b, "{\"id\":\""); sb_put(b, idt, 36); sb_str(b, "\",\"agent_id\":"); sb_jstr(b, r->agent_id, strnlen(r->agent_id, 32));
    sb_str(b, ",\"method\":"); sb_jstr(b, method_name(r->flags), strlen(method_name(r->flags)));
    sb_str(b, ",\"path\":"); sb_jstr(b, (const char*)r->payload, r->path_len);
    sb_str(b, ",\"headers\":{");
    const char* h = (const char*)r->payload + r->path_len; const char* hend = h + r->hdr_len; int first = 1;
    while (h < hend) {                                  /* flattened "Key: Value\n", already sorted by key */
        const char* nl = (const char*)memchr(h, '\n', (size_t)(hend - h)); if (!nl) nl = hend;
        const char* colon = (const char*)memchr(h, ':', (size_t)(nl - h));
        if (colon) { if (!first) sb_put(b, ",", 1); first = 0; sb_jstr(b, h, (size_t)(colon - h)); sb_put(b, ":", 1); const char* v = colon + 1; if (v < nl && *v == ' ') v++; sb_jstr(b, v, (size_t)(nl - v)); }
        h = nl + 1;
    }
    sb_str(b, "},\"body\":");
    sb_b64(b, r->payload + r->path_len + r->hdr_len, r->body_len);       /* io.ReadAll: a non-nil empty slice is "" */
    sb_str(b, ",\"status\":\""); sb_str(b, status_name(r->status));
    char t[96]; int m = snprintf(t, sizeof t, "\",\"retry_count\":%u,\"max_retries\":%u,\"created_at\":", r->retry_count, r->max_retries);
    sb_put(b, t, (size_t)m); sb_time(b, r->seq);
    if (d->processed_at || d->has_response) { sb_str(b, ",\"processed_at\":"); sb_time(b, d->processed_at); }
    if (d->has_response) {
        m = snprintf(t, sizeof t, ",\"response\":{\"status_code\":%u,\"headers\":{},\"body\":\"\",\"received_at\":", d->resp_status);
        sb_put(b, t, (size_t)m); sb_time(b, d->received_at); sb_put(b, "}", 1);
    }
    if (d->error[0]) { sb_str(b, ",\"error\":"); sb_jstr(b, d->error, strlen(d->error)); }
    sb_put(b, "}", 1);
}

/* json.Unmarshal(data, &request) — requests.go:159,216,238.  Small recursive-descent reader for the shape above. */
typedef struct { const char* p; const char* e; } jr;
static void j_ws(jr* j) { while (j->p < j->e && (*j->p == ' ' || *j->p == '\n' || *j->p == '\t' || *j->p == '\r')) j->p++; }
static size_t j_string(jr* j, char* out, size_t cap) {   /* decodes escapes; returns length */
    size_t n = 0; j->p++;
    while (j->p < j->e && *j->p != '"') {
        char c = *j->p++;
        if (c == '\\') {
            char x = *j->p++;
            if (x == 'n') c = '\n'; else if (x == 't') c = '\t'; else if (x == 'r') c = '\r'; else if (x == 'b') c = '\b'; else if (x == 'f') c = '\f';
            else if (x == 'u') {                                  /* \uXXXX -> UTF-8 (encoding/json writes no surrogate pairs here) */
                unsigned cp = (unsigned)((hexv(j->p[0]) << 12) | (hexv(j->p[1]) << 8) | (hexv(j->p[2]) << 4) | hexv(j->p[3])); j->p += 4;
                if (cp < 0x80) c = (char)cp;
                else if (cp < 0x800) { if (n < cap) out[n] = (char)(0xc0 | (cp >> 6)); n++; c = (char)(0x80 | (cp & 0x3f)); }
                else { if (n < cap) out[n] = (char)(0xe0 | (cp >> 12)); n++; if (n < cap) out[n] = (char)(0x80 | ((cp >> 6) & 0x3f)); n++; c = (char)(0x80 | (cp & 0x3f)); }
            }
            else c = x;
        }
        if (n < cap) out[n] = c;
        n++;
    }
    j->p++;
    return n;
}
static int b64v(char c) { if (c >= 'A' && c <= 'Z') return c - 'A'; if (c >= 'a' && c <= 'z') return c - 'a' + 26; if (c >= '0' && c <= '9') return c - '0' + 52; if (c == '+') return 62; if (c == '/') return 63; return -1; }
static size_t j_b64(jr* j, uint8_t* out, size_t cap) {
    size_t n = 0; j->p++; uint32_t acc = 0; int bits = 0;
    while (j->p < j->e && *j->p != '"') { int v = b64v(*j->p++); if (v < 0) continue; acc = (acc << 6) | (uint32_t)v; bits += 6; if (bits >= 8) { bits -= 8; if (n < cap) out[n] = (uint8_t)(acc >> bits); n++; } }
    j->p++;
    return n;
}
static uint64_t j_uint(jr* j) { uint64_t v = 0; while (j->p < j->e && *j->p >= '0' && *j->p <= '9') v = v * 10 + (uint64_t)(*j->p++ - '0'); return v; }
static unsigned get_n(const char* t, int n) { unsigned v = 0; for (int k = 0; k < n; ++k) v = v * 10 + (unsigned)(t[k] - '0'); return v; }
static uint64_t j_time(jr* j) {                             /* time.Time.UnmarshalJSON of the form sb_time writes */
    char t[48]; size_t n = j_string(j, t, sizeof t - 1); t[n < 47 ? n : 47] = 0;
    if (n < 20) return 0;
    unsigned y = get_n(t, 4), mo = get_n(t + 5, 2), d = get_n(t + 8, 2), hh = get_n(t + 11, 2), mi = get_n(t + 14, 2), ss = get_n(t + 17, 2);
    uint64_t frac = 0;
    if (t[19] == '.') { int k = 0; for (const char* q = t + 20; *q >= '0' && *q <= '9'; ++q, ++k) frac = frac * 10 + (uint64_t)(*q - '0'); for (; k < 9; ++k) frac *= 10; }
    int64_t yy = (int64_t)y - (mo <= 2); int64_t era = yy / 400; unsigned yoe = (unsigned)(yy - era * 400);
    unsigned doy = (153 * (mo > 2 ? mo - 3 : mo + 9) + 2) / 5 + d - 1; unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    int64_t days = era * 146097 + (int64_t)doe - 719468;
    return ((uint64_t)days * 86400ULL + hh * 3600ULL + mi * 60ULL + ss) * 1000000000ULL + frac;
}
static void j_skip(jr* j) {
    j_ws(j);
    if (*j->p == '"') { char d[1]; j_string(j, d, 0); return; }
    if (*j->p == '{' || *j->p == '[') { char o = *j->p, c = (o == '{') ? '}' : ']'; int depth = 0; do { if (*j->p == '"') { char d[1]; j_string(j, d, 0); continue; } if (*j->p == o) depth++; else if (*j->p == c) depth--; j->p++; } while (depth > 0 && j->p < j->e); return; }
    while (j->p < j->e && *j->p != ',' && *j->p != '}' && *j->p != ']') j->p++;
}
static int unmarshal_request(const char* s, size_t n, reqdoc* d) {
    memset(d, 0, sizeof *d);
    jr j = {s, s + n}; agr_record* r = &d->rec;
    char path[512]; size_t path_len = 0; char hdrs[512]; size_t hdr_len = 0; uint8_t body[512]; size_t body_len = 0;
    j_ws(&j); if (*j.p != '{') return -1; j.p++;
    for (;;) {
        j_ws(&j); if (*j.p == '}') break; if (*j.p == ',') { j.p++; continue; }
        char key[32]; size_t kl = j_string(&j, key, sizeof key - 1); key[kl < 31 ? kl : 31] = 0;
        j_ws(&j); j.p++; j_ws(&j);
        if (!strcmp(key, "id")) { char t[40]; j_string(&j, t, 39); uuid_parse(t, r->request_id); }
        else if (!strcmp(key, "agent_id")) { size_t m = j_string(&j, r->agent_id, 31); r->agent_id[m < 31 ? m : 31] = 0; }
        else if (!strcmp(key, "method")) { char t[16]; size_t m = j_string(&j, t, 15); r->flags |= method_code(t, m) << AGR_F_METHOD_SHIFT; }
        else if (!strcmp(key, "path")) path_len = j_string(&j, path, sizeof path);
        else if (!strcmp(key, "headers")) {
            j.p++;                                           /* { */
            for (;;) { j_ws(&j); if (*j.p == '}') { j.p++; break; } if (*j.p == ',') { j.p++; continue; }
                hdr_len += j_string(&j, hdrs + hdr_len, sizeof hdrs - hdr_len); hdrs[hdr_len++] = ':'; hdrs[hdr_len++] = ' ';
                j_ws(&j); j.p++; j_ws(&j);
                hdr_len += j_string(&j, hdrs + hdr_len, sizeof hdrs - hdr_len); hdrs[hdr_len++] = '\n'; }
        }
        else if (!strcmp(key, "body")) { if (*j.p == '"') body_len = j_b64(&j, body, sizeof body); else j_skip(&j); }
        else if (!strcmp(key, "status")) { char t[16]; size_t m = j_string(&j, t, 15); t[m] = 0; r->status = !strcmp(t, "pending") ? AGR_ST_PENDING : !strcmp(t, "completed") ? AGR_ST_COMPLETED : !strcmp(t, "failed") ? AGR_ST_FAILED : !strcmp(t, "processing") ? AGR_ST_PROCESSING : 0; }
        else if (!strcmp(key, "retry_count")) r->retry_count = (uint8_t)j_uint(&j);
        else if (!strcmp(key, "max_retries")) r->max_retries = (uint8_t)j_uint(&j);
        else if (!strcmp(key, "created_at")) r->seq = j_time(&j);
        else if (!strcmp(key, "processed_at")) d->processed_at = j_time(&j);
        else if (!strcmp(key, "error")) { size_t m = j_string(&j, d->error, sizeof d->error - 1); d->error[m < 63 ? m : 63] = 0; }
        else if (!strcmp(key, "response")) {
            d->has_response = 1; j.p++;
            for (;;) { j_ws(&j); if (*j.p == '}') { j.p++; break; } if (*j.p == ',') { j.p++; continue; }
                char k2[24]; size_t m = j_string(&j, k2, 23); k2[m < 23 ? m : 23] = 0; j_ws(&j); j.p++; j_ws(&j);
                if (!strcmp(k2, "status_code")) d->resp_status = (uint16_t)j_uint(&j);
                else if (!strcmp(k2, "received_at")) d->received_at = j_time(&j);
                else j_skip(&j); }
        }
        else j_skip(&j);
    }
    if (path_len + hdr_len + body_len > AGR_PAYLOAD_BYTES) return -1;
    memcpy(r->payload, path, path_len); memcpy(r->payload + path_len, hdrs, hdr_len); memcpy(r->payload + path_len + hdr_len, body, body_len);
    r->path_len = (uint16_t)path_len; r->hdr_len = (uint16_t)hdr_len; r->body_len = (uint32_t)body_len;
    r->resp_status = d->resp_status; r->error_code = d->error[0] ? AGR_OUT_ERROR : 0;
    return 0;
}

/* ------------------------------------------------------------------ the path */
typedef struct cref {
    keyspace ks;
    sbuf scratch;
    uint32_t flags;
    char (*agents)[32]; uint32_t n_agents, cap_agents;      /* registration order = canonical cross-agent order (Q9) */
    uint64_t stats_ingested, stats_stored, stats_completions, stats_failures;
} cref;

cref* cref_create(uint32_t flags) {
    cref* c = (cref*)calloc(1, sizeof(cref));
    ks_init(&c->ks, 1u << 16);
    c->flags = flags ? flags : AGR_CFG_PERSISTENCE;
    return c;
}
void cref_destroy(cref* c) {
    if (!c) return;
    for (uint64_t i = 0; i < c->ks.cap; ++i) free(c->ks.e[i].items);
    for (achunk* a = c->ks.arena; a;) { achunk* nx = a->next; free(a); a = nx; }
    free(c->ks.e); free(c->scratch.p); free(c->agents); free(c);
}
static const char* agent_status_name(uint8_t s) { static const char* n[] = {"created", "running", "stopped", "paused", "failed"}; return s < 5 ? n[s] : "unknown"; }
static int agent_index(cref* c, const char* id) { for (uint32_t i = 0; i < c->n_agents; ++i) if (!strncmp(c->agents[i], id, 32)) return (int)i; return -1; }

/* saveAgent (agent.go:510-530): SET agent:{id} <agent JSON>.  Only the fields the path reads are kept realistic. */
int cref_set_agent_state(cref* c, const char* agent_id, uint8_t status) {
    int idx = agent_index(c, agent_id);
    if (idx < 0) {
        if (c->n_agents == c->cap_agents) { c->cap_agents = c->cap_agents ? c->cap_agents * 2 : 64; c->agents = (char(*)[32])realloc(c->agents, (size_t)c->cap_agents * 32); }
        idx = (int)c->n_agents++; memset(c->agents[idx], 0, 32); strncpy(c->agents[idx], agent_id, 31);
    }
    char key[64], doc[640];
    int kl = snprintf(key, sizeof key, "agent:%s", agent_id);
    int dl = snprintf(doc, sizeof doc,
        "{\"id\":\"%s\",\"name\":\"synthetic-agent\",\"image\":\"agentainer/gpt-agent:latest\",\"container_id\":\"0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef\","
        "\"status\":\"%s\",\"env_vars\":{\"OPENAI_API_KEY\":\"sk-test\"},\"cpu_limit\":1000000000,\"memory_limit\":536870912,\"auto_restart\":true,"
        "\"token\":\"agentainer-default-token\",\"ports\":[],\"volumes\":[],\"created_at\":\"2025-01-01T00:00:00Z\",\"updated_at\":\"2025-01-01T00:00:00Z\"}",
        agent_id, agent_status_name(status));
    r_set_ex(&c->ks, key, (uint32_t)kl, doc, (uint32_t)dl, 0);      /* saveAgent: no TTL (agent.go:519-522) */
    return idx;
}
/* GetAgent (agent.go:372-390): GET agent:{id} + json.Unmarshal; returns status code or -1 */
static int get_agent_status(cref* c, const char* agent_id) {
    char key[64]; int kl = snprintf(key, sizeof key, "agent:%.31s", agent_id);
    kent* e = r_get(&c->ks, key, (uint32_t)kl);
    if (!e) return -1;
    jr j = {e->val, e->val + e->vlen}; int status = -1;     /* walk the whole document like Unmarshal does */
    j_ws(&j); j.p++;
    for (;;) { j_ws(&j); if (j.p >= j.e || *j.p == '}') break; if (*j.p == ',') { j.p++; continue; }
        char k[24]; size_t m = j_string(&j, k, 23); k[m < 23 ? m : 23] = 0; j_ws(&j); j.p++; j_ws(&j);
        if (!strcmp(k, "status")) { char t[16]; size_t n = j_string(&j, t, 15); t[n < 15 ? n : 15] = 0; for (int s = 0; s < 5; ++s) if (!strcmp(t, agent_status_name((uint8_t)s))) status = s; }
        else j_skip(&j); }
    return status;
}
/* agent.Manager.Remove cleanup (agent.go:343-359) */
int cref_drop_agent(cref* c, const char* agent_id) {
    char key[96]; int kl = snprintf(key, sizeof key, "agent:%s", agent_id);
    ks_del_entry(&c->ks, ks_find(&c->ks, key, (uint32_t)kl, 0));
    const char* q[] = {"pending", "completed", "failed"};
    for (int i = 0; i < 3; ++i) { kl = snprintf(key, sizeof key, "agent:%s:requests:%s", agent_id, q[i]); ks_del_entry(&c->ks, ks_find(&c->ks, key, (uint32_t)kl, 0)); }
    return 0;
}

static int rec_key(char* key, size_t cap, const char* agent_id, const char* idt) { return snprintf(key, cap, "agent:%.31s:requests:%s", agent_id, idt); }
static int list_key(char* key, size_t cap, const char* agent_id, const char* q) { return snprintf(key, cap, "agent:%.31s:requests:%s", agent_id, q); }

/* proxyToAgentHandler decision (server.go:493-541) with StoreRequest (requests.go:64-117) */
int cref_ingest(cref* c, const agr_record* recs, uint32_t n, agr_verdict* out) {
    char key[128], idt[37];
    for (uint32_t i = 0; i < n; ++i) {
        const agr_record* r = &recs[i];
        agr_verdict v; memset(&v, 0, sizeof v);
        c->stats_ingested++;
        int astatus = get_agent_status(c, r->agent_id);                               /* server.go:498 */
        if (astatus < 0) { v.code = AGR_V_NOT_FOUND; v.http_status = 404; v.agent_slot = 0x00ffffffu; out[i] = v; continue; }   /* :499-502 */
        v.agent_slot = (uint32_t)agent_index(c, r->agent_id);
        int is_replay = (r->flags & AGR_F_REPLAY) != 0;                                /* :506 */
        int tracked = 0;
        if ((c->flags & AGR_CFG_PERSISTENCE) && !is_replay) {                           /* :508 */
            reqdoc d; memset(&d, 0, sizeof d);
            d.rec = *r; d.rec.status = AGR_ST_PENDING; d.rec.retry_count = 0; d.rec.max_retries = r->max_retries ? r->max_retries : 3;   /* requests.go:86-97 */
            d.rec.flags &= ~AGR_F_REPLAY; memset(d.rec.replay_of, 0, 16);
            marshal_request(&c->scratch, &d);                                          /* :101 */
            uuid_text(r->request_id, idt);
            int kl = rec_key(key, sizeof key, r->agent_id, idt);
            r_set(&c->ks, key, (uint32_t)kl, c->scratch.p, (uint32_t)c->scratch.n);  /* :106 */
            kl = list_key(key, sizeof key, r->agent_id, "pending");
            r_rpush(&c->ks, key, (uint32_t)kl, idt);                                   /* :112 */
            tracked = 1; v.flags |= AGR_VF_STORED | AGR_VF_TRACKED; c->stats_stored++;
        } else if (is_replay) {                                                         /* :519-522 */
            v.flags |= AGR_VF_REPLAY;
            static const uint8_t z[16] = {0};
            tracked = memcmp(r->replay_of, z, 16) != 0;
            if (tracked) v.flags |= AGR_VF_TRACKED;
        }
        if (astatus != AGR_AGENT_RUNNING) {                                            /* :525 */
            if ((c->flags & AGR_CFG_PERSISTENCE) && tracked) { v.code = AGR_V_QUEUED; v.http_status = 202; }
            else { v.code = AGR_V_UNAVAILABLE; v.http_status = 503; }
        } else v.code = AGR_V_FORWARD;
        out[i] = v;
    }
    return 0;
}

/* StoreResponse (requests.go:120-194) / MarkRequestFailed (:228-275), selected like RoundTrip does (server.go:588-611) */
int cref_complete(cref* c, const agr_outcome* outs, uint32_t n, int32_t* results) {
    char key[128], idt[37]; static const uint8_t z[16] = {0};
    for (uint32_t j = 0; j < n; ++j) {
        const agr_outcome* o = &outs[j];
        if (results) results[j] = 0;
        if (memcmp(o->request_id, z, 16) == 0) continue;                               /* t.requestID == "" */
        if (o->kind == AGR_OUT_DIAL_ERR) continue;                                     /* server.go:600-605 */
        if (o->kind != AGR_OUT_RESPONSE && o->kind != AGR_OUT_ERROR) continue;
        uuid_text(o->request_id, idt);
        int kl = rec_key(key, sizeof key, o->agent_id, idt);
        kent* e = r_get(&c->ks, key, (uint32_t)kl);                                    /* :153 / :232 */
        if (!e) { if (results) results[j] = AGR_ENOTFOUND; continue; }
        reqdoc d;
        if (unmarshal_request(e->val, e->vlen, &d) != 0) { if (results) results[j] = AGR_ENOTFOUND; continue; }   /* :159 / :238 */
        char lk[128]; int lkl;
        if (o->kind == AGR_OUT_RESPONSE) {
            d.has_response = 1; d.resp_status = o->http_status; d.received_at = o->seq;   /* :142-147,165 */
            d.rec.status = AGR_ST_COMPLETED; d.processed_at = o->seq;                    /* :166-167 */
            marshal_request(&c->scratch, &d);                                            /* :170 */
            r_set(&c->ks, key, (uint32_t)kl, c->scratch.p, (uint32_t)c->scratch.n);    /* :175 */
            lkl = list_key(lk, sizeof lk, o->agent_id, "pending"); r_lrem1(&c->ks, lk, (uint32_t)lkl, idt);      /* :180-184 */
            lkl = list_key(lk, sizeof lk, o->agent_id, "completed"); r_rpush(&c->ks, lk, (uint32_t)lkl, idt);    /* :187-191 */
            c->stats_completions++;
        } else {
            d.rec.status = AGR_ST_FAILED; strcpy(d.error, "transport error");            /* :243-244 */
            if (d.rec.retry_count < 255) d.rec.retry_count++;                            /* :245 */
            if (d.rec.retry_count < d.rec.max_retries) d.rec.status = AGR_ST_PENDING;    /* :248-249 */
            else {
                lkl = list_key(lk, sizeof lk, o->agent_id, "failed"); r_rpush(&c->ks, lk, (uint32_t)lkl, idt);   /* :252-255 */
                lkl = list_key(lk, sizeof lk, o->agent_id, "pending"); r_lrem1(&c->ks, lk, (uint32_t)lkl, idt);  /* :258-261 */
            }
            marshal_request(&c->scratch, &d);                                            /* :265 */
            r_set(&c->ks, key, (uint32_t)kl, c->scratch.p, (uint32_t)c->scratch.n);    /* :270 */
            c->stats_failures++;
        }
    }
    return 0;
}

/* GetPendingRequests (requests.go:197-225): LRANGE + GET + Unmarshal each; missing / invalid silently skipped */
static uint32_t pending_of(cref* c, const char* agent_id, reqdoc** out) {
    char key[128]; int kl = list_key(key, sizeof key, agent_id, "pending");
    kent* l = ks_find(&c->ks, key, (uint32_t)kl, 0);
    if (!l || l->type != T_LIST || l->n == 0) { *out = NULL; return 0; }
    uint32_t n = l->n; lid* ids = (lid*)malloc((size_t)n * sizeof(lid)); memcpy(ids, l->items, (size_t)n * sizeof(lid));   /* LRANGE 0 -1 */
    reqdoc* docs = (reqdoc*)malloc((size_t)n * sizeof(reqdoc)); uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) {
        kl = rec_key(key, sizeof key, agent_id, ids[i].id);
        kent* e = r_get(&c->ks, key, (uint32_t)kl);
        if (!e) continue;                                                               /* :210-213 */
        if (unmarshal_request(e->val, e->vlen, &docs[m]) != 0) continue;                /* :216-219 */
        m++;
    }
    free(ids); *out = docs; return m;
}
int cref_pending(cref* c, const char* agent_id, agr_record* out, uint32_t cap, uint32_t* n) {
    reqdoc* docs; uint32_t m = pending_of(c, agent_id, &docs);
    *n = m;
    if (m > cap) { free(docs); return AGR_ECAP; }
    for (uint32_t i = 0; i < m; ++i) out[i] = docs[i].rec;
    free(docs); return 0;
}
/* processAgents (replay_worker.go:58-87) up to the HTTP call: KEYS scan over the WHOLE keyspace, running check,
 * FIFO snapshot, skip rule (:101).  Cross-agent order canonicalised to registration order (Q9). */
int cref_scan(cref* c, agr_dispatch* out, agr_record* recs, uint32_t cap, uint32_t* n) {
    uint32_t total = 0; int rc = 0;
    uint8_t* has = (uint8_t*)calloc(c->n_agents + 1, 1);
    for (uint64_t i = 0; i < c->ks.cap; ++i) {                                          /* KEYS agent:*:requests:pending */
        kent* e = &c->ks.e[i];
        if (!e->key || !e->type) continue;
        size_t kl = e->klen;
        if (kl > 23 && !memcmp(e->key, "agent:", 6) && !memcmp(e->key + kl - 17, ":requests:pending", 17)) {
            char aid[32]; size_t al = kl - 23; if (al > 31) continue; memcpy(aid, e->key + 6, al); aid[al] = 0;   /* extractAgentID */
            int idx = agent_index(c, aid); if (idx >= 0) has[idx] = 1;
        }
    }
    for (uint32_t a = 0; a < c->n_agents; ++a) {
        if (!has[a]) continue;
        if (get_agent_status(c, c->agents[a]) != AGR_AGENT_RUNNING) continue;           /* :76-81,166-189 */
        reqdoc* docs; uint32_t m = pending_of(c, c->agents[a], &docs);                   /* :91 */
        for (uint32_t i = 0; i < m; ++i) {
            if (docs[i].rec.status == AGR_ST_PROCESSING || docs[i].rec.retry_count >= docs[i].rec.max_retries) continue;   /* :101 */
            if (total < cap) {
                if (out) { memset(&out[total], 0, sizeof out[total]); out[total].agent_slot = a; memcpy(out[total].request_id, docs[i].rec.request_id, 16); }
                if (recs) recs[total] = docs[i].rec;
            } else rc = AGR_ECAP;
            total++;
        }
        free(docs);
    }
    free(has); *n = total; return rc;
}
int cref_get_record(cref* c, const char* agent_id, const uint8_t request_id[16], agr_record* out) {
    char key[128], idt[37]; uuid_text(request_id, idt);
    int kl = rec_key(key, sizeof key, agent_id, idt);
    kent* e = r_get(&c->ks, key, (uint32_t)kl);
    if (!e) return AGR_ENOTFOUND;
    reqdoc d; if (unmarshal_request(e->val, e->vlen, &d) != 0) return AGR_ENOTFOUND;
    *out = d.rec; return 0;
}
int cref_list(cref* c, const char* agent_id, int which, uint8_t (*ids)[16], uint32_t cap, uint32_t* n) {
    const char* q[] = {"pending", "completed", "failed"};
    char key[128]; int kl = list_key(key, sizeof key, agent_id, q[which]);
    kent* l = ks_find(&c->ks, key, (uint32_t)kl, 0);
    uint32_t m = (l && l->type == T_LIST) ? l->n : 0;
    *n = m;
    if (m > cap) return AGR_ECAP;
    for (uint32_t i = 0; i < m; ++i) uuid_parse(l->items[i].id, ids[i]);
    return 0;
}
uint64_t cref_keys(cref* c) { return c->ks.used; }
void cref_set_now(cref* c, uint64_t now) { c->ks.now = now; }   /* the Redis server's clock, same unit as the records' times */
/* json.Marshal(GetPendingRequests(agent)) — the "pending" member of GET /agents/{id}/requests (server.go:638-650): every
 * record is unmarshalled by GetPendingRequests and marshalled again by the handler; a nil slice is null */
int cref_pending_json(cref* c, const char* agent_id, char* out, uint32_t cap, uint32_t* len) {
    reqdoc* docs; uint32_t m = pending_of(c, agent_id, &docs);
    sbuf all = {0, 0, 0}, one = {0, 0, 0};
    if (m == 0) sb_str(&all, "null");
    else {
        sb_put(&all, "[", 1);
        for (uint32_t i = 0; i < m; ++i) { if (i) sb_put(&all, ",", 1); marshal_request(&one, &docs[i]); sb_put(&all, one.p, one.n); }
        sb_put(&all, "]", 1);
    }
    free(docs);
    *len = (uint32_t)all.n;
    int rc = 0;
    if (all.n > cap) rc = AGR_ECAP; else memcpy(out, all.p, all.n);
    free(all.p); free(one.p);
    return rc;
}
/* the value of agent:{a}:requests:{r} as it sits in the keyspace (what storage.Get returns, server.go:661-662) */
int cref_get_json(cref* c, const char* agent_id, const uint8_t request_id[16], char* out, uint32_t cap, uint32_t* len) {
    char key[128], idt[37];
    uuid_text(request_id, idt);
    int kl = rec_key(key, sizeof key, agent_id, idt);
    kent* e = r_get(&c->ks, key, (uint32_t)kl);
    if (!e) return AGR_ENOTFOUND;
    *len = e->vlen;
    if (e->vlen > cap) return AGR_ECAP;
    memcpy(out, e->val, e->vlen);
    return 0;
}
