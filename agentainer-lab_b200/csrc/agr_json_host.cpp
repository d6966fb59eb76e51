// agr_json_host.cpp — host-side reader of the wire form: json.Marshal(requests.Request) -> binary record.
//
// The other direction of K5 (agr_k5_json.cu).  The reference keeps every record in Redis as this JSON
// (requests.go:101,170,265) and reads it back with json.Unmarshal (requests.go:159,216,238; server.go:669,695); a host that
// migrates an existing Redis keyspace, or that checks what K5 produced, needs the same reader.  Pure host code: no CUDA
// call, no handle.  Accepts what encoding/json accepts for this shape: members in any order, unknown members skipped,
// all string escapes incl. surrogate pairs, null for absent maps / slices / pointers, RFC 3339 times with any offset.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/agentainer_gpu.h"

namespace {

struct reader {
    const uint8_t* p; const uint8_t* e; bool ok = true;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < e && *p == (uint8_t)c) { ++p; return true; } return false; }
    bool peek(char c) { ws(); return p < e && *p == (uint8_t)c; }
    bool lit(const char* s) { size_t n = strlen(s); ws(); if ((size_t)(e - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; } return false; }
};
int hexv(uint8_t c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }
void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) { out.push_back((char)(0xc0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3f))); }
    else if (cp < 0x10000) { out.push_back((char)(0xe0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3f))); out.push_back((char)(0x80 | (cp & 0x3f))); }
    else { out.push_back((char)(0xf0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3f))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3f))); out.push_back((char)(0x80 | (cp & 0x3f))); }
}
bool read_u16(reader& r, uint32_t& v) {
    if (r.e - r.p < 4) return false;
    v = 0;
    for (int k = 0; k < 4; ++k) { int h = hexv(r.p[k]); if (h < 0) return false; v = (v << 4) | (uint32_t)h; }
    r.p += 4;
    return true;
}
// a JSON string -> bytes (encoding/json's unquote: escapes decoded, lone surrogates become U+FFFD)
bool read_string(reader& r, std::string& out) {
    out.clear();
    if (!r.eat('"')) return false;
    while (r.p < r.e) {
        uint8_t c = *r.p++;
        if (c == '"') return true;
        if (c < 0x20) return false;
        if (c != '\\') { out.push_back((char)c); continue; }
        if (r.p >= r.e) return false;
        uint8_t x = *r.p++;
        switch (x) {
            case '"': out.push_back('"'); break;   case '\\': out.push_back('\\'); break;  case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;  case 'f': out.push_back('\f'); break;   case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;  case 't': out.push_back('\t'); break;
            case 'u': {
                uint32_t cp;
                if (!read_u16(r, cp)) return false;
                if (cp >= 0xd800 && cp < 0xdc00) {                       // high surrogate: needs \uDC00..DFFF right behind it
                    uint32_t lo = 0;
                    const uint8_t* save = r.p;
                    if (r.e - r.p >= 6 && r.p[0] == '\\' && r.p[1] == 'u' && (r.p += 2, read_u16(r, lo)) && lo >= 0xdc00 && lo < 0xe000)
                        cp = 0x10000 + ((cp - 0xd800) << 10) + (lo - 0xdc00);
                    else { r.p = save; cp = 0xfffd; }
                } else if (cp >= 0xdc00 && cp < 0xe000) cp = 0xfffd;
                put_utf8(out, cp);
                break;
            }
            default: return false;
        }
    }
    return false;
}
// encoding/json refuses documents nested deeper than 10000 ("exceeded max depth"); unknown members are skipped recursively, so
// the same limit keeps a hostile document from overflowing the host stack
#define JSON_MAX_DEPTH 10000
bool skip_value(reader& r, int depth);
bool skip_container(reader& r, char open, char close, int depth) {
    if (depth > JSON_MAX_DEPTH) return false;
    if (!r.eat(open)) return false;
    if (r.eat(close)) return true;
    for (;;) {
        if (open == '{') { std::string k; if (!read_string(r, k) || !r.eat(':')) return false; }
        if (!skip_value(r, depth)) return false;
        if (r.eat(',')) continue;
        return r.eat(close);
    }
}
bool skip_value(reader& r, int depth = 0) {
    r.ws();
    if (r.p >= r.e) return false;
    if (*r.p == '"') { std::string s; return read_string(r, s); }
    if (*r.p == '{') return skip_container(r, '{', '}', depth + 1);
    if (*r.p == '[') return skip_container(r, '[', ']', depth + 1);
    if (r.lit("null") || r.lit("true") || r.lit("false")) return true;
    const uint8_t* s = r.p;
    while (r.p < r.e && (*r.p == '-' || *r.p == '+' || *r.p == '.' || *r.p == 'e' || *r.p == 'E' || (*r.p >= '0' && *r.p <= '9'))) ++r.p;
    return r.p > s;
}
bool read_uint(reader& r, uint64_t& v) {
    r.ws();
    const uint8_t* s = r.p; v = 0;
    while (r.p < r.e && *r.p >= '0' && *r.p <= '9') v = v * 10 + (uint64_t)(*r.p++ - '0');
    return r.p > s;
}
// map[string]string -> "Key: Value\n" lines sorted by key bytes (the form agr_record.payload keeps); null -> empty
bool read_headers(reader& r, std::string& flat) {
    flat.clear();
    if (r.lit("null")) return true;
    if (!r.eat('{')) return false;
    std::vector<std::pair<std::string, std::string>> kv;
    if (!r.eat('}')) {
        for (;;) {
            std::string k, v;
            if (!read_string(r, k) || !r.eat(':') || !read_string(r, v)) return false;
            // the flattened form frames a header as "Key: Value\n": a line feed in either part, or a colon in the key, cannot be
            // represented (net/http never produces such a header; a document that holds one is refused, not mangled)
            if (k.find('\n') != std::string::npos || v.find('\n') != std::string::npos || k.find(':') != std::string::npos) return false;
            bool dup = false;
            for (auto& e : kv) if (e.first == k) { e.second = v; dup = true; }   // the last duplicate wins, like a Go map
            if (!dup) kv.emplace_back(std::move(k), std::move(v));
            if (r.eat(',')) continue;
            if (!r.eat('}')) return false;
            break;
        }
    }
    std::sort(kv.begin(), kv.end());
    for (auto& e : kv) { flat += e.first; flat += ": "; flat += e.second; flat += '\n'; }
    return true;
}
int b64v(uint8_t c) { return c >= 'A' && c <= 'Z' ? c - 'A' : c >= 'a' && c <= 'z' ? c - 'a' + 26 : c >= '0' && c <= '9' ? c - '0' + 52 : c == '+' ? 62 : c == '/' ? 63 : -1; }
bool read_bytes(reader& r, std::string& out) {              // []byte: base64.StdEncoding in a string, or null
    out.clear();
    if (r.lit("null")) return true;
    std::string s;
    if (!read_string(r, s)) return false;
    uint32_t acc = 0; int bits = 0;
    for (uint8_t c : s) {
        if (c == '=') break;
        const int v = b64v(c);
        if (v < 0) return false;
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((char)(uint8_t)(acc >> bits)); }
    }
    return true;
}
// time.Time.UnmarshalJSON: RFC 3339, optional fraction, Z or a numeric offset -> Unix nanoseconds
bool read_time(reader& r, uint64_t& ns) {
    std::string s;
    if (!read_string(r, s) || s.size() < 20) return false;
    auto num = [&](size_t at, size_t n, uint32_t& v) { v = 0; for (size_t k = 0; k < n; ++k) { const char c = s[at + k]; if (c < '0' || c > '9') return false; v = v * 10 + (uint32_t)(c - '0'); } return true; };
    uint32_t y, mo, d, hh, mi, ss;
    if (!num(0, 4, y) || s[4] != '-' || !num(5, 2, mo) || s[7] != '-' || !num(8, 2, d) || (s[10] != 'T' && s[10] != 't') ||
        !num(11, 2, hh) || s[13] != ':' || !num(14, 2, mi) || s[16] != ':' || !num(17, 2, ss)) return false;
    size_t at = 19; uint64_t frac = 0;
    if (at < s.size() && s[at] == '.') {
        ++at; int k = 0;
        while (at < s.size() && s[at] >= '0' && s[at] <= '9') { if (k < 9) { frac = frac * 10 + (uint64_t)(s[at] - '0'); ++k; } ++at; }
        for (; k < 9; ++k) frac *= 10;
    }
    long long off = 0;
    if (at < s.size() && (s[at] == 'Z' || s[at] == 'z')) ++at;
    else if (at + 6 <= s.size() && (s[at] == '+' || s[at] == '-')) {
        uint32_t oh, om;
        if (!num(at + 1, 2, oh) || s[at + 3] != ':' || !num(at + 4, 2, om)) return false;
        off = (long long)(oh * 3600 + om * 60) * (s[at] == '-' ? -1 : 1);
        at += 6;
    } else return false;
    if (at != s.size() || mo < 1 || mo > 12 || d < 1 || d > 31) return false;
    const long long yy = (long long)y - (mo <= 2);
    const long long era = (yy >= 0 ? yy : yy - 399) / 400;
    const uint32_t yoe = (uint32_t)(yy - era * 400);
    const uint32_t doy = (153 * (mo > 2 ? mo - 3 : mo + 9) + 2) / 5 + d - 1;
    const uint32_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    const long long days = era * 146097 + (long long)doe - 719468;
    const long long secs = days * 86400 + (long long)(hh * 3600 + mi * 60 + ss) - off;
    if (secs < 0) return false;
    ns = (uint64_t)secs * 1000000000ULL + frac;
    return true;
}
bool parse_uuid(const std::string& s, uint8_t id[16]) {
    if (s.size() != 36) return false;
    size_t o = 0;
    for (int i = 0; i < 16; ++i) {
        if (o == 8 || o == 13 || o == 18 || o == 23) { if (s[o] != '-') return false; ++o; }
        const int h = hexv((uint8_t)s[o]), l = hexv((uint8_t)s[o + 1]);
        if (h < 0 || l < 0) return false;
        id[i] = (uint8_t)((h << 4) | l); o += 2;
    }
    return true;
}

}  // namespace

extern "C" int agr_json_decode(const uint8_t* json, uint32_t len, uint8_t* record, uint32_t record_cap, uint8_t* resp, uint32_t resp_cap,
                               char* error, uint32_t error_cap, agr_decoded* out) {
    if (!json || !out || (record_cap && !record) || (resp_cap && !resp) || (error_cap && !error)) return AGR_EINVAL;
    memset(out, 0, sizeof *out);
    reader r{json, json + len};
    std::string id, agent, method, path, hdrs, body, status, err, rhdrs, rbody, key;
    uint64_t retry = 0, maxr = 0, rstatus = 0;
    if (!r.eat('{')) return AGR_EINVAL;
    if (!r.eat('}')) {
        for (;;) {
            if (!read_string(r, key) || !r.eat(':')) return AGR_EINVAL;
            bool ok;
            if (key == "id") ok = read_string(r, id);
            else if (key == "agent_id") ok = read_string(r, agent);
            else if (key == "method") ok = read_string(r, method);
            else if (key == "path") ok = read_string(r, path);
            else if (key == "headers") ok = read_headers(r, hdrs);
            else if (key == "body") ok = read_bytes(r, body);
            else if (key == "status") ok = read_string(r, status);
            else if (key == "retry_count") ok = read_uint(r, retry);
            else if (key == "max_retries") ok = read_uint(r, maxr);
            else if (key == "created_at") ok = read_time(r, out->created_at);
            else if (key == "processed_at") ok = r.lit("null") || read_time(r, out->processed_at);
            else if (key == "error") ok = read_string(r, err);
            else if (key == "response") {
                if (r.lit("null")) ok = true;
                else {
                    ok = r.eat('{');
                    out->has_response = 1;
                    if (ok && !r.eat('}')) {
                        for (;;) {
                            std::string k2;
                            if (!read_string(r, k2) || !r.eat(':')) return AGR_EINVAL;
                            bool ok2;
                            if (k2 == "status_code") ok2 = read_uint(r, rstatus);
                            else if (k2 == "headers") ok2 = read_headers(r, rhdrs);
                            else if (k2 == "body") ok2 = read_bytes(r, rbody);
                            else if (k2 == "received_at") ok2 = read_time(r, out->received_at);
                            else ok2 = skip_value(r);
                            if (!ok2) return AGR_EINVAL;
                            if (r.eat(',')) continue;
                            if (!r.eat('}')) return AGR_EINVAL;
                            break;
                        }
                    }
                }
            }
            else ok = skip_value(r);
            if (!ok) return AGR_EINVAL;
            if (r.eat(',')) continue;
            if (!r.eat('}')) return AGR_EINVAL;
            break;
        }
    }
    r.ws();
    if (r.p != r.e) return AGR_EINVAL;
    // ---- the record: 96 B header + payload (path | flattened headers | body) rounded up to 16 B
    const size_t pay = path.size() + hdrs.size() + body.size();
    const size_t rec_len = AGR_HEADER_BYTES + ((pay + 15) & ~(size_t)15);
    out->record_len = (uint32_t)rec_len;
    out->resp_hdr_len = (uint32_t)rhdrs.size(); out->resp_body_len = (uint32_t)rbody.size(); out->error_len = (uint32_t)err.size();
    out->status = status == "pending" ? AGR_ST_PENDING : status == "processing" ? AGR_ST_PROCESSING : status == "completed" ? AGR_ST_COMPLETED
                  : status == "failed" ? AGR_ST_FAILED : AGR_ST_NONE;
    out->retry_count = (uint8_t)std::min<uint64_t>(retry, 255); out->max_retries = (uint8_t)std::min<uint64_t>(maxr, 255);
    out->resp_status = (uint16_t)std::min<uint64_t>(rstatus, 65535);
    if (rec_len > AGR_VAR_MAX_RECORD || path.size() > 0xffff || hdrs.size() > 0xffff || agent.size() >= AGR_AGENT_ID_BYTES) return AGR_EINVAL;
    if (rec_len > record_cap || rhdrs.size() + rbody.size() > resp_cap || err.size() > error_cap) return AGR_ECAP;
    memset(record, 0, rec_len);
    agr_record* h = (agr_record*)record;                       // only the 96-byte header part is addressed through the struct
    if (!id.empty() && !parse_uuid(id, h->request_id)) return AGR_EINVAL;
    memcpy(h->agent_id, agent.data(), agent.size());
    h->seq = out->created_at;
    static const char* methods[] = {"", "GET", "POST", "PUT", "DELETE", "PATCH", "HEAD", "OPTIONS"};
    uint32_t mcode = 0;
    for (uint32_t c = 1; c < 8; ++c) if (method == methods[c]) mcode = c;
    h->flags = mcode << AGR_F_METHOD_SHIFT;
    h->path_len = (uint16_t)path.size(); h->hdr_len = (uint16_t)hdrs.size(); h->body_len = (uint32_t)body.size();
    h->status = out->status; h->retry_count = out->retry_count; h->max_retries = out->max_retries;
    h->error_code = err.empty() ? 0 : AGR_OUT_ERROR; h->resp_status = out->resp_status;
    uint8_t* p = record + AGR_HEADER_BYTES;
    memcpy(p, path.data(), path.size()); p += path.size();
    memcpy(p, hdrs.data(), hdrs.size()); p += hdrs.size();
    memcpy(p, body.data(), body.size());
    if (!rhdrs.empty()) memcpy(resp, rhdrs.data(), rhdrs.size());
    if (!rbody.empty()) memcpy(resp + rhdrs.size(), rbody.data(), rbody.size());
    if (!err.empty()) memcpy(error, err.data(), err.size());
    return 0;
}
