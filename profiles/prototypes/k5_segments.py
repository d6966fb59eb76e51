"""Prototype (not product code) of the emit-by-output-position design for K5: a per-record table of
segments, and every 16-byte output chunk generated INDEPENDENTLY from that table — what one lane would do.  Checked here
against oracle/gojson.py on records that need no escaping (the design's fast path), so that the CUDA version starts from a
segment vocabulary and position arithmetic that are known to be right.     usage: python profiles/prototypes/k5_segments.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import gojson as G  # noqa: E402

B64 = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"
LIT, RAW, BASE64 = 0, 1, 2          # segment kinds; uuid / time / numbers are short LIT segments built once per record


def plain(b: bytes) -> bool:
    return all(0x20 <= c < 0x7F and c not in b'"\\<>&' for c in b)


def segments(rec):
    """[(kind, out_off, out_len, data)] for json.Marshal(rec); None if some string needs escaping (slow path)."""
    segs, off = [], 0

    def add(kind, data, out_len=None):
        nonlocal off
        n = len(data) if out_len is None else out_len
        if n:
            segs.append((kind, off, n, data))
            off += n

    strs = [rec["id"], rec["agent_id"], rec["method"], rec["path"], rec["status"], rec.get("error") or ""]
    hdrs = [(k, v) for m in (rec["headers"], (rec.get("response") or {}).get("headers", {})) for k, v in m.items()]
    raw = lambda x: x if isinstance(x, (bytes, bytearray)) else x.encode()
    if not all(plain(raw(x)) for x in strs) or not all(plain(raw(k)) and plain(raw(v)) for k, v in hdrs):
        return None, 0

    def string(x):
        add(LIT, b'"'); add(RAW, raw(x)); add(LIT, b'"')

    def header_map(m):
        add(LIT, b"{")
        for i, (k, v) in enumerate(sorted(m.items(), key=lambda kv: raw(kv[0]))):
            add(LIT, b"," if i else b""); string(k); add(LIT, b":"); string(v)
        add(LIT, b"}")

    def b64(body):
        add(LIT, b'"'); add(BASE64, bytes(body), 4 * ((len(body) + 2) // 3)); add(LIT, b'"')

    add(LIT, b'{"id":'); string(rec["id"])
    add(LIT, b',"agent_id":'); string(rec["agent_id"])
    add(LIT, b',"method":'); string(rec["method"])
    add(LIT, b',"path":'); string(rec["path"])
    add(LIT, b',"headers":'); header_map(rec["headers"])
    add(LIT, b',"body":'); b64(rec["body"])
    add(LIT, b',"status":'); string(rec["status"])
    add(LIT, b',"retry_count":' + str(rec["retry_count"]).encode() + b',"max_retries":' + str(rec["max_retries"]).encode())
    add(LIT, b',"created_at":' + G.go_time(rec["created_at"]))
    if rec.get("processed_at") is not None:
        add(LIT, b',"processed_at":' + G.go_time(rec["processed_at"]))
    r = rec.get("response")
    if r is not None:
        add(LIT, b',"response":{"status_code":' + str(r["status_code"]).encode() + b',"headers":'); header_map(r["headers"])
        add(LIT, b',"body":'); b64(r["body"])
        add(LIT, b',"received_at":' + G.go_time(r["received_at"]) + b"}")
    if rec.get("error"):
        add(LIT, b',"error":'); string(rec["error"])
    add(LIT, b"}")
    return segs, off


def byte_at(seg, k):
    """Output byte k of a segment, computed from k alone."""
    kind, _, _, data = seg
    if kind != BASE64:
        return data[k]
    g, j = divmod(k, 4)                                    # group of 3 input bytes, character inside the group
    b = data[3 * g: 3 * g + 3]
    if j > len(b):
        return ord("=")
    v = int.from_bytes(b + bytes(3 - len(b)), "big")
    return B64[(v >> (18 - 6 * j)) & 63]


def emit_chunk(segs, total, c):
    """Bytes [16c, 16c + 16) of the record's JSON: binary search for the first segment, then walk."""
    lo, hi, pos = 0, len(segs) - 1, 16 * c
    while lo < hi:                                          # last segment whose out_off <= pos
        mid = (lo + hi + 1) // 2
        if segs[mid][1] <= pos:
            lo = mid
        else:
            hi = mid - 1
    out, s = bytearray(), lo
    while len(out) < 16 and pos < total:
        seg = segs[s]
        if pos >= seg[1] + seg[2]:
            s += 1
            continue
        out.append(byte_at(seg, pos - seg[1]))
        pos += 1
    return bytes(out)


def main():
    from jsoncase import make_requests, make_script, run_model
    agents = ["agent-1700000000000000001", "agent-1700000000000000002"]
    import random
    rng = random.Random(2)
    reqs = make_requests(41, 600, agents)
    for i, r in enumerate(reqs):                             # two thirds realistic (plain ASCII), one third adversarial as generated
        if i % 3:
            r.path = f"/agent/{r.agent_id}/chat/{rng.randrange(10**6)}".encode()
            r.headers = {k: v for k, v in ((b"Content-Type", b"application/json"), (b"User-Agent", b"curl/8.5.0"), (b"Accept", b"*/*"))
                         if rng.random() < 0.7}
            r.body = r.body[: max(0, 416 - len(r.path) - len(r.flat_headers()))]
    script = [op for op in make_script(41, len(reqs), 700)
              if not (reqs[op[1]].path.startswith(b"/agent/") and b"/chat/" in reqs[op[1]].path) or
              (op[0] == "resp" and all(plain(k) and plain(v) for k, v in op[3].items())) or (op[0] == "err" and plain(op[2]))]
    redis, _ = run_model(reqs, script)
    fast = slow = 0
    max_segs = 0
    for r in reqs:
        rec = redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}")
        segs, total = segments(rec)
        if segs is None:
            slow += 1
            continue
        want = G.marshal_request(rec)
        got = b"".join(emit_chunk(segs, total, c) for c in range((total + 15) // 16))
        assert got == want, (got, want)
        fast += 1
        max_segs = max(max_segs, len(segs))
    print(f"{fast} records emitted chunk by chunk == oracle; {slow} need the escaping path; at most {max_segs} segments per record")


if __name__ == "__main__":
    main()
