"""Seeded record streams for the JSON wire-form tests: a list of requests with awkward strings, followed by a script of
outcomes, driven into oracle/model.py's Manager (the literal Go restatement) and, on the GPU, into the engine."""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

METHODS = ["GET", "POST", "PUT", "DELETE", "PATCH", "HEAD", "OPTIONS"]
# strings that exercise encoding/json's escaper
AWKWARD = [b"plain", b'q"uote', b"back\\slash", b"<script>&amp;</script>", b"tab\there", b"ctl\x01\x1f\x7f", b"bs\x08ff\x0c",
           "café".encode(), " line sep".encode(), "\U0001f600 emoji".encode(), "� real".encode(),
           b"bad\xff\xfe", b"trunc\xe2\x82", b"overlong\xc0\xaf", b"surrogate\xed\xa0\x80", b"lonecont\x80\xbf", b"f5\xf5\x80\x80\x80",
           b"e0\xe0\x80\x80", b"f4\xf4\x90\x80\x80", b"ok\xf4\x8f\xbf\xbf", b"mix\xe2\x80\xa8\xe2\x80", b"a:b:c", b"sp  ace", b""]


@dataclass
class JReq:
    agent_id: str
    method: str
    path: bytes
    headers: Dict[bytes, bytes]
    body: bytes
    now: int                     # Unix nanoseconds
    rid: bytes = b""             # 16 raw bytes (hash-id mode: supplied; mint mode: filled in after ingest)

    def flat_headers(self) -> bytes:
        return b"".join(k + b": " + v + b"\n" for k, v in sorted(self.headers.items()))


def _hval(rng) -> bytes:
    v = rng.choice(AWKWARD)
    return v.replace(b"\n", b" ")


def _times(rng) -> int:
    base = 1_700_000_000 + rng.randrange(0, 400_000_000)
    frac = rng.choice([0, 1, 10, 123_000_000, 999_999_999, 500, 120_000, rng.randrange(0, 1_000_000_000)])
    return base * 1_000_000_000 + frac


def make_requests(seed: int, n: int, agents: List[str], max_payload: int = 416, big_bodies: bool = False) -> List[JReq]:
    rng = random.Random(seed)
    out = []
    for i in range(n):
        a = rng.choice(agents)
        path = f"/agent/{a}/".encode() + rng.choice(AWKWARD)
        nh = rng.choice([0, 1, 1, 2, 3, 5])
        headers: Dict[bytes, bytes] = {}
        for _ in range(nh):
            k = rng.choice([b"Content-Type", b"Accept", b"X-Custom", b"User-Agent", b"Host", b"X-Trace-Id", b"Authorization", b"X-A", b"x-lower"])
            headers[k] = _hval(rng)
        if big_bodies and rng.random() < 0.3:
            blen = rng.randrange(400, 6000)
        else:
            blen = rng.choice([0, 1, 2, 3, 4, 5, 31, 32, 33, 95, 96, 97, rng.randrange(0, 200)])
        r = JReq(a, rng.choice(METHODS), path, headers, b"", _times(rng))
        room = max_payload - len(r.path) - len(r.flat_headers())
        while room < 0:                                        # keep the fixed 416 B payload budget
            headers.pop(next(iter(headers)))
            room = max_payload - len(r.path) - len(r.flat_headers())
        blen = min(blen, room)
        r.body = bytes(rng.randrange(256) for _ in range(blen))
        r.rid = bytes(rng.randrange(256) for _ in range(16))
        out.append(r)
    return out


def make_script(seed: int, n_reqs: int, n_ops: int) -> List[Tuple]:
    """Outcome script: ("resp", i, code, headers, body, now) | ("err", i, text)."""
    rng = random.Random(seed * 7919 + 1)
    ops = []
    for _ in range(n_ops):
        i = rng.randrange(n_reqs)
        if rng.random() < 0.55:
            hdrs = {}
            for _ in range(rng.choice([0, 1, 2, 3])):
                hdrs[rng.choice([b"Content-Type", b"Server", b"X-Resp", b"Date", b"Set-Cookie"])] = _hval(rng)
            body = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 2, 3, 50, 300, rng.randrange(0, 1500)])))
            ops.append(("resp", i, rng.choice([200, 201, 204, 404, 500, 503]), hdrs, body, _times(rng)))
        else:
            ops.append(("err", i, rng.choice([b"EOF", b"read tcp 10.0.0.1:1->10.0.0.2:8000: connection reset by peer",
                                              b'net/http: "quoted" <err> & more', b"bad\xffbytes", b"context deadline exceeded"])))
    return ops


def run_model(reqs: List[JReq], script: List[Tuple]):
    """oracle/model.py's Manager driven like StoreRequest / StoreResponse / MarkRequestFailed are; returns the model."""
    from oracle import model as M, gojson as G
    redis = M.MiniRedis()
    mgr = M.Manager(redis)
    for r in reqs:
        mgr.store_request(r.agent_id, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.now))
    for op in script:
        r = reqs[op[1]]
        rid = G.format_uuid(r.rid)
        if op[0] == "resp":
            mgr.store_response(r.agent_id, rid, M.HttpResponse(op[2], dict(op[3]), op[4], now=op[5]))
        else:
            mgr.mark_request_failed(r.agent_id, rid, op[2])
    return redis, mgr


def records_array(reqs: List[JReq]) -> np.ndarray:
    from agentainer_lab_b200 import record_dtype, constants as K
    recs = np.zeros(len(reqs), dtype=record_dtype)
    for i, r in enumerate(reqs):
        recs[i]["request_id"] = np.frombuffer(r.rid, dtype=np.uint8)
        recs[i]["agent_id"] = r.agent_id.encode()
        recs[i]["seq"] = r.now
        recs[i]["flags"] = K.METHOD_CODES[r.method] << K.AGR_F_METHOD_SHIFT
        hdrs = r.flat_headers()
        blob = r.path + hdrs + r.body
        assert len(blob) <= 416
        recs[i]["path_len"], recs[i]["hdr_len"], recs[i]["body_len"] = len(r.path), len(hdrs), len(r.body)
        recs[i]["status"], recs[i]["max_retries"] = K.AGR_ST_PENDING, 3
        recs[i]["payload"][: len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    return recs


def var_batch(reqs: List[JReq]):
    from agentainer_lab_b200 import header_dtype, constants as K
    parts, offsets = [], [0]
    for r in reqs:
        hdrs = r.flat_headers()
        payload = r.path + hdrs + r.body
        h = np.zeros(1, dtype=header_dtype)
        h["request_id"] = np.frombuffer(r.rid, dtype=np.uint8)
        h["agent_id"] = r.agent_id.encode()
        h["seq"] = r.now
        h["flags"] = K.METHOD_CODES[r.method] << K.AGR_F_METHOD_SHIFT
        h["path_len"], h["hdr_len"], h["body_len"] = len(r.path), len(hdrs), len(r.body)
        h["status"], h["max_retries"] = K.AGR_ST_PENDING, 3
        parts.append(h.tobytes() + payload + bytes((-len(payload)) % 16))
        offsets.append(offsets[-1] + len(parts[-1]))
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy(), np.array(offsets, dtype=np.uint32)
