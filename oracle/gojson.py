"""
oracle/gojson.py — CPU restatement of Go's encoding/json.Marshal for requests.Request / requests.Response.

TEST INFRASTRUCTURE ONLY (same rule as oracle/model.py): imported by tests/ only, as the checker of the K5 kernels.

PARITY UNPINNED.  No Go toolchain exists in the build image, so these bytes cannot be compared with a run of
encoding/json.  The authority is the Go 1.23 standard library's documented behaviour (go.mod:3-5 of the reference
pins go 1.23), restated below; tests/golden/gojson_kats.json holds hand-derived known answers labelled as such.

What the reference marshals (all paths relative to the reference tree):
  * requests.Request, internal/requests/requests.go:27-41 — field order is struct order; `omitempty` on
    processed_at (*time.Time), response (*Response) and error (string);
  * requests.Response, requests.go:44-49;
  * json.Marshal call sites: requests.go:101 (StoreRequest), :170 (StoreResponse, after Unmarshal of the stored value),
    :265 (MarkRequestFailed, after Unmarshal), server.go:646-650 / 674-678 (management reads re-marshal).

encoding/json rules used (Go 1.23, src/encoding/json/encode.go):
  * strings: `"` and `\\` backslash-escaped; \\n \\r \\t \\b \\f short forms (\\b, \\f since Go 1.22); other bytes < 0x20 as
    \\u00XX (lower-case hex); `<`, `>`, `&` as \\u003c \\u003e \\u0026 (Marshal always escapes HTML); U+2028 / U+2029 as
    \\u2028 / \\u2029; each byte that does not start a valid UTF-8 sequence (utf8.DecodeRuneInString returns
    RuneError, width 1) as \\ufffd; everything else copied through;
  * map[string]string: keys sorted bytewise, `{}` when empty, `null` when nil (never nil here: requests.go:78,135 `make`);
  * []byte: base64.StdEncoding with padding, in quotes; `null` when nil (never nil on the HTTP path: io.ReadAll returns
    a non-nil empty slice, and the json round trip keeps it non-nil);
  * int: decimal;
  * time.Time: RFC 3339 with up to nine fractional digits, trailing zeros dropped; `Z` for a zero offset.  The reference
    runs in a container without TZ data configured, so time.Now() carries offset 0; the event stream supplies the
    instant as Unix nanoseconds.

The round trip matters: StoreResponse and MarkRequestFailed Unmarshal the stored JSON and Marshal it again.  Unmarshal
turns an escaped \\ufffd into a real U+FFFD, which the second Marshal copies through as the three bytes EF BF BD, so a
string that held invalid UTF-8 is written differently before and after its first round trip.  Types carry that here:
a `bytes` string is fresh from the wire (may hold invalid UTF-8), a `str` has been through Unmarshal (`unmarshal_strings`
is what oracle/model.py applies where the Go code calls json.Unmarshal).
"""
from __future__ import annotations

import base64
from typing import Dict, Optional

_HEX = "0123456789abcdef"


def _decode_rune(b: bytes, i: int):
    """unicode/utf8.DecodeRune at b[i:]: (is_valid, width)."""
    n = len(b) - i
    b0 = b[i]
    if b0 < 0x80:
        return True, 1
    lo, hi = 0x80, 0xBF
    if 0xC2 <= b0 <= 0xDF:
        need = 2
    elif 0xE0 <= b0 <= 0xEF:
        need = 3
        if b0 == 0xE0:
            lo = 0xA0
        elif b0 == 0xED:
            hi = 0x9F
    elif 0xF0 <= b0 <= 0xF4:
        need = 4
        if b0 == 0xF0:
            lo = 0x90
        elif b0 == 0xF4:
            hi = 0x8F
    else:
        return False, 1
    if n < need:
        return False, 1
    if not (lo <= b[i + 1] <= hi):
        return False, 1
    for k in range(2, need):
        if not (0x80 <= b[i + k] <= 0xBF):
            return False, 1
    return True, need


def go_decode(b: bytes) -> str:
    """What json.Unmarshal leaves in a Go string that was marshalled from b: invalid bytes have become U+FFFD."""
    out, i = [], 0
    while i < len(b):
        ok, w = _decode_rune(b, i)
        out.append(b[i:i + w].decode("utf-8") if ok else "\ufffd")
        i += w
    return "".join(out)


def go_string(s) -> bytes:
    """encodeState.string(s, escapeHTML=true), quotes included.  s: bytes (fresh) or str (already round-tripped)."""
    if isinstance(s, str):
        s = s.encode("utf-8")
    out = bytearray(b'"')
    i = 0
    while i < len(s):
        c = s[i]
        if c < 0x80:
            if c >= 0x20 and c not in (0x22, 0x5C, 0x3C, 0x3E, 0x26):
                out.append(c)
            elif c in (0x22, 0x5C):
                out += b"\\" + bytes([c])
            elif c == 0x0A:
                out += b"\\n"
            elif c == 0x0D:
                out += b"\\r"
            elif c == 0x09:
                out += b"\\t"
            elif c == 0x08:
                out += b"\\b"
            elif c == 0x0C:
                out += b"\\f"
            else:
                out += b"\\u00" + _HEX[c >> 4].encode() + _HEX[c & 15].encode()
            i += 1
            continue
        ok, w = _decode_rune(s, i)
        if not ok:
            out += b"\\ufffd"
            i += 1
            continue
        if s[i:i + w] in (b"\xe2\x80\xa8", b"\xe2\x80\xa9"):
            out += b"\\u202" + (b"8" if s[i + 2] == 0xA8 else b"9")
        else:
            out += s[i:i + w]
        i += w
    out += b'"'
    return bytes(out)


def _raw(x) -> bytes:
    return bytes(x) if isinstance(x, (bytes, bytearray)) else x.encode("utf-8")


def go_map(m: Optional[Dict]) -> bytes:
    if m is None:
        return b"null"
    items = sorted(m.items(), key=lambda kv: _raw(kv[0]))      # sorted by the key's bytes, before escaping
    return b"{" + b",".join(go_string(k) + b":" + go_string(v) for k, v in items) + b"}"


def go_bytes(b: Optional[bytes]) -> bytes:
    if b is None:
        return b"null"
    return b'"' + base64.b64encode(bytes(b)) + b'"'


def go_time(ns: int) -> bytes:
    """time.Unix(0, ns).UTC().MarshalJSON()"""
    secs, frac = divmod(int(ns), 1_000_000_000)
    days, sod = divmod(secs, 86400)
    # civil from days (proleptic Gregorian)
    z = days + 719468
    era = z // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + 3 if mp < 10 else mp - 9
    if m <= 2:
        y += 1
    s = "%04d-%02d-%02dT%02d:%02d:%02d" % (y, m, d, sod // 3600, sod // 60 % 60, sod % 60)
    if frac:
        s += "." + ("%09d" % frac).rstrip("0")
    return ('"' + s + 'Z"').encode()


_STATUS = {"pending", "processing", "completed", "failed"}


def marshal_response(r: dict) -> bytes:
    return (b'{"status_code":' + str(int(r["status_code"])).encode() + b',"headers":' + go_map(r["headers"]) +
            b',"body":' + go_bytes(r["body"]) + b',"received_at":' + go_time(r["received_at"]) + b"}")


def marshal_request(r: dict) -> bytes:
    """json.Marshal(requests.Request) for a record dict of oracle/model.py."""
    out = bytearray()
    out += b'{"id":' + go_string(r["id"])
    out += b',"agent_id":' + go_string(r["agent_id"])
    out += b',"method":' + go_string(r["method"])
    out += b',"path":' + go_string(r["path"])
    out += b',"headers":' + go_map(r["headers"])
    out += b',"body":' + go_bytes(r["body"])
    out += b',"status":' + go_string(r["status"])
    out += b',"retry_count":' + str(int(r["retry_count"])).encode()
    out += b',"max_retries":' + str(int(r["max_retries"])).encode()
    out += b',"created_at":' + go_time(r["created_at"])
    if r.get("processed_at") is not None:
        out += b',"processed_at":' + go_time(r["processed_at"])
    if r.get("response") is not None:
        out += b',"response":' + marshal_response(r["response"])
    if r.get("error"):
        out += b',"error":' + go_string(r["error"])
    out += b"}"
    return bytes(out)


def _us(x):
    return go_decode(bytes(x)) if isinstance(x, (bytes, bytearray)) else x


def unmarshal_strings(r: dict) -> dict:
    """The effect of json.Unmarshal(json.Marshal(r)) on the string-typed fields of a record dict (in place)."""
    for k in ("id", "agent_id", "method", "path", "status", "error"):
        r[k] = _us(r[k])
    r["headers"] = {_us(k): _us(v) for k, v in r["headers"].items()}
    if r.get("response") is not None:
        r["response"]["headers"] = {_us(k): _us(v) for k, v in r["response"]["headers"].items()}
    return r


def marshal_list(rs) -> bytes:
    """json.Marshal([]*Request): a nil slice is null."""
    if not rs:
        return b"null"
    return b"[" + b",".join(marshal_request(r) for r in rs) + b"]"


def format_uuid(raw16: bytes) -> str:
    h = raw16.hex()
    return f"{h[0:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:32]}"


def _parse_time(s: str) -> int:
    """RFC 3339 (what time.Time.MarshalJSON writes) -> Unix nanoseconds; any numeric offset is accepted."""
    import calendar
    import re
    m = re.fullmatch(r"(\d{4})-(\d{2})-(\d{2})T(\d{2}):(\d{2}):(\d{2})(?:\.(\d{1,9}))?(Z|[+-]\d{2}:\d{2})", s)
    if not m:
        raise ValueError(f"not an RFC 3339 time: {s!r}")
    y, mo, d, h, mi, sec = (int(m.group(i)) for i in range(1, 7))
    frac = int((m.group(7) or "0").ljust(9, "0"))
    secs = calendar.timegm((y, mo, d, h, mi, sec, 0, 0, 0))
    z = m.group(8)
    if z != "Z":
        off = (int(z[1:3]) * 60 + int(z[4:6])) * 60
        secs -= off if z[0] == "+" else -off
    return secs * 1_000_000_000 + frac


def unmarshal_request(text: bytes) -> dict:
    """json.Unmarshal of a stored record into the dict form marshal_request takes (used to check Redis values written by the
    REFERENCE itself, tests/golden/from_reference: marshal_request(unmarshal_request(v)) == v pins the field order, the
    escaping, base64, time and omitempty rules of this restatement to encoding/json's output)."""
    import json as _json
    d = _json.loads(text.decode("utf-8"))
    body = None if d.get("body") is None else base64.b64decode(d["body"])
    r = {"id": d["id"], "agent_id": d["agent_id"], "method": d["method"], "path": d["path"], "headers": d.get("headers"),
         "body": body, "status": d["status"], "retry_count": d["retry_count"], "max_retries": d["max_retries"],
         "created_at": _parse_time(d["created_at"]), "processed_at": None, "response": None, "error": d.get("error", "")}
    if d.get("processed_at") is not None:
        r["processed_at"] = _parse_time(d["processed_at"])
    if d.get("response") is not None:
        q = d["response"]
        r["response"] = {"status_code": q["status_code"], "headers": q.get("headers"),
                         "body": None if q.get("body") is None else base64.b64decode(q["body"]), "received_at": _parse_time(q["received_at"])}
    return r
