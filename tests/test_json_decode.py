"""agr_json_decode — the host-side reader of the wire form (json.Unmarshal of requests.Request into the binary record).
Pure host code: these tests need no GPU.  Checked against oracle/gojson.py: decoding what the oracle marshals gives back
the fields Go's decoder would hold, for fresh and for round-tripped records."""
import base64

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K
from oracle import gojson as G
from jsoncase import make_requests, make_script, run_model, METHODS

AGENTS = ["agent-1700000000000000001", "agent-1700000000000000002", "agent-1700000000000000003"]
ST = {"pending": K.AGR_ST_PENDING, "processing": K.AGR_ST_PROCESSING, "completed": K.AGR_ST_COMPLETED, "failed": K.AGR_ST_FAILED}


def flat(m):
    items = sorted((G.go_decode(k) if isinstance(k, bytes) else k, G.go_decode(v) if isinstance(v, bytes) else v) for k, v in m.items())
    return "".join(f"{k}: {v}\n" for k, v in sorted(items, key=lambda kv: kv[0].encode())).encode()


def s_(x):
    return (G.go_decode(x) if isinstance(x, (bytes, bytearray)) else x).encode()


def test_decode_inverts_the_oracles_marshal():
    reqs = make_requests(17, 300, AGENTS)
    redis, mgr = run_model(reqs, make_script(17, len(reqs), 420))
    seen_resp = seen_err = 0
    for r in reqs:
        rec = redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}")
        rc, d = A.json_decode(G.marshal_request(rec))
        assert rc == 0
        h, info = d["header"], d["info"]
        assert bytes(h["request_id"]) == r.rid and h["agent_id"].decode() == r.agent_id
        assert (int(h["flags"]) >> K.AGR_F_METHOD_SHIFT) & 0xff == K.METHOD_CODES[r.method]
        assert d["path"] == s_(rec["path"]) and d["body"] == r.body and d["headers"] == flat(rec["headers"])
        assert info.status == ST[rec["status"]] == h["status"] and info.retry_count == rec["retry_count"] == h["retry_count"]
        assert info.max_retries == 3 and info.created_at == r.now == int(h["seq"])
        assert info.has_response == (rec["response"] is not None)
        if rec["response"] is not None:
            seen_resp += 1
            assert info.resp_status == rec["response"]["status_code"] and info.received_at == rec["response"]["received_at"]
            assert info.processed_at == rec["processed_at"]
            assert d["resp_headers"] == flat(rec["response"]["headers"]) and d["resp_body"] == rec["response"]["body"]
        else:
            assert info.processed_at == 0 and info.received_at == 0 and d["resp_headers"] == b"" and d["resp_body"] == b""
        assert d["error"] == (s_(rec["error"]) if rec["error"] else b"")
        seen_err += bool(rec["error"])
        assert info.record_len == 96 + (len(d["path"]) + len(d["headers"]) + len(d["body"]) + 15) // 16 * 16
    assert seen_resp > 50 and seen_err > 50


def test_decode_accepts_what_encoding_json_accepts():
    js = (b' { "unknown" : [1, {"a": "b"}, null], "status":"failed", "path":"/agent/a/\\u00e9\\ud83d\\ude00\\ud800x\\/", "id":"00010203-0405-0607-0809-0a0b0c0d0e0f",'
          b'"agent_id":"a","method":"DELETE","headers":{"b":"2","a":"1","b":"3"},"body":null,"retry_count": 3 ,"max_retries":3,'
          b'"created_at":"2023-11-14T23:13:20.5+01:00","processed_at":null,"response":null,"error":"x\\ty"} ')
    rc, d = A.json_decode(js)
    assert rc == 0
    assert d["path"] == "/agent/a/é\U0001f600�x/".encode()          # surrogate pair joined, lone surrogate -> U+FFFD, \/ -> /
    assert d["headers"] == b"a: 1\nb: 3\n" and d["body"] == b""          # keys sorted; the last duplicate wins; null slice
    assert d["info"].status == K.AGR_ST_FAILED and d["info"].retry_count == 3 and d["error"] == b"x\ty"
    assert d["info"].created_at == 1_700_000_000_500_000_000 and d["info"].has_response == 0   # +01:00 -> UTC
    assert bytes(d["header"]["request_id"]) == bytes(range(16))
    assert (int(d["header"]["flags"]) >> 8) & 0xff == K.AGR_M_DELETE


def test_decode_rejects_malformed_input():
    good = G.marshal_request({"id": "00010203-0405-0607-0809-0a0b0c0d0e0f", "agent_id": "a", "method": "GET", "path": "/agent/a/", "headers": {},
                              "body": b"", "status": "pending", "retry_count": 0, "max_retries": 3, "created_at": 1, "processed_at": None,
                              "response": None, "error": ""})
    assert A.json_decode(good)[0] == 0
    for bad in (good[:-1], good + b"x", good.replace(b'"GET"', b'"GET'), good.replace(b"1970-01-01", b"1970-13-01"),
                good.replace(b'"id":"0001', b'"id":"zz01'), b"", b"[]", good.replace(b'"path":"', b'"path":"\\q')):
        assert A.json_decode(bad)[0] == K.AGR_EINVAL, bad
    long_agent = good.replace(b'"agent_id":"a"', b'"agent_id":"' + b"a" * 40 + b'"')
    assert A.json_decode(long_agent)[0] == K.AGR_EINVAL
    big = good.replace(b'"body":""', b'"body":"' + base64.b64encode(bytes(9000)) + b'"')
    assert A.json_decode(big)[0] == K.AGR_EINVAL                          # longer than AGR_VAR_MAX_RECORD


def test_decode_then_marshal_is_gos_round_trip():
    """Unmarshal + Marshal of a stored value (what StoreResponse does first): decoding the oracle's bytes and marshalling
    the decoded fields again gives the oracle's own round-tripped form."""
    import copy
    reqs = make_requests(23, 200, AGENTS)
    redis, _ = run_model(reqs, make_script(23, len(reqs), 150))
    for r in reqs:
        rec = redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}")
        rc, d = A.json_decode(G.marshal_request(rec))
        assert rc == 0
        info = d["info"]
        def unflat(b):
            return {k: v for k, v in (ln.split(b": ", 1) for ln in b.split(b"\n") if ln)}
        back = {"id": G.format_uuid(bytes(d["header"]["request_id"])), "agent_id": d["header"]["agent_id"].decode(),
                "method": METHODS[((int(d["header"]["flags"]) >> 8) & 0xff) - 1], "path": d["path"].decode(),
                "headers": {k.decode(): v.decode() for k, v in unflat(d["headers"]).items()}, "body": d["body"],
                "status": {v: k for k, v in ST.items()}[info.status], "retry_count": info.retry_count, "max_retries": info.max_retries,
                "created_at": info.created_at, "processed_at": info.processed_at if info.has_response else None,
                "response": ({"status_code": info.resp_status, "headers": {k.decode(): v.decode() for k, v in unflat(d["resp_headers"]).items()},
                              "body": d["resp_body"], "received_at": info.received_at} if info.has_response else None),
                "error": d["error"].decode()}
        assert G.marshal_request(back) == G.marshal_request(G.unmarshal_strings(copy.deepcopy(rec)))


def test_decoder_survives_mutated_input_under_asan(tmp_path):
    """The reader takes untrusted bytes (a Redis dump): 100 k mutations of real documents through an ASAN + UBSAN build of
    the decoder, with exact-size input buffers and deliberately small output buffers."""
    import os, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        import pytest
        pytest.skip("no g++")
    reqs = make_requests(3, 40, AGENTS[:2])
    redis, _ = run_model(reqs, make_script(3, 40, 60))
    seeds = tmp_path / "seeds.bin"
    with open(seeds, "wb") as f:
        for r in reqs:
            j = G.marshal_request(redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"))
            f.write(len(j).to_bytes(4, "little") + j)
    exe = tmp_path / "fuzz"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", os.path.join(root, "include"), os.path.join(root, "agentainer-lab_b200", "csrc", "agr_json_host.cpp"),
                            os.path.join(root, "tests", "fuzz_json_decode.cpp"), "-o", str(exe)], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        import pytest
        pytest.skip("sanitizers not available in this toolchain")
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe), str(seeds), "100000"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-3000:]
    ok, ecap, einval = (int(x) for x in run.stdout.split()[1::2])
    assert ok > 1000 and ecap > 1000 and einval > 1000


def test_decoder_on_the_hand_derived_kats():
    """tests/golden/gojson_kats.json was written by hand from encoding/json's rules; the product's reader must turn each
    document back into the record it was derived from."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gojson_kats.json")) as f:
        kats = json.load(f)["records"]
    assert len(kats) >= 6
    for k in kats:
        rec = k["record"]
        rc, d = A.json_decode(k["json"].encode())
        assert rc == 0, k["name"]
        h, info = d["header"], d["info"]
        assert G.format_uuid(bytes(h["request_id"])) == rec["id"] and h["agent_id"].decode() == rec["agent_id"], k["name"]
        assert METHODS[((int(h["flags"]) >> 8) & 0xff) - 1] == rec["method"] and d["path"].decode() == rec["path"]
        assert d["headers"] == "".join(f"{a}: {b}\n" for a, b in sorted(rec["headers"].items(), key=lambda kv: kv[0].encode())).encode()
        assert d["body"] == bytes.fromhex(rec["body_hex"])
        assert info.status == ST[rec["status"]] and info.retry_count == rec["retry_count"] and info.max_retries == rec["max_retries"]
        assert info.created_at == rec["created_at"] and d["error"].decode() == rec["error"]
        if rec["response"]:
            r = rec["response"]
            assert info.has_response and info.resp_status == r["status_code"] and info.received_at == r["received_at"]
            assert info.processed_at == rec["processed_at"] and d["resp_body"] == bytes.fromhex(r["body_hex"])
            assert d["resp_headers"] == "".join(f"{a}: {b}\n" for a, b in sorted(r["headers"].items())).encode()
        else:
            assert not info.has_response


def test_decoder_refuses_what_it_cannot_represent():
    """Hostile or unrepresentable documents: nesting deeper than encoding/json's 10000 levels in a skipped member (the reader
    recurses there: without a limit a few hundred thousand brackets overflow the host stack), and header keys / values that
    would break the flattened "Key: Value\\n" framing.  All of them are AGR_EINVAL, none of them crashes."""
    base = ('{"id":"00000000-0000-4000-8000-000000000001","agent_id":"agent-1","method":"POST","path":"/agent/agent-1/x",'
            '"headers":%s,"body":"aGk=","status":"pending","retry_count":0,"max_retries":3,"created_at":"2023-11-14T22:13:20Z"%s}')
    ok = base % ('{"A":"b"}', "")
    assert A.json_decode(ok.encode())[0] == 0
    deep_ok = base % ('{"A":"b"}', ',"x":' + "[" * 5000 + "]" * 5000)
    assert A.json_decode(deep_ok.encode())[0] == 0                                   # encoding/json accepts this depth too
    deep = base % ('{"A":"b"}', ',"x":' + "[" * 400000 + "]" * 400000)
    assert A.json_decode(deep.encode())[0] == K.AGR_EINVAL
    deep_obj = base % ('{"A":"b"}', ',"x":' + '{"a":' * 300000 + "1" + "}" * 300000)
    assert A.json_decode(deep_obj.encode())[0] == K.AGR_EINVAL
    for hdr in ('{"A\\nB":"v"}', '{"A":"v\\nX: y"}', '{"A:B":"v"}'):
        assert A.json_decode((base % (hdr, "")).encode())[0] == K.AGR_EINVAL, hdr
