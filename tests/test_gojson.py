"""The Go-JSON restatement (oracle/gojson.py) against hand-derived known answers (tests/golden/gojson_kats.json) and
its own structural properties.  PARITY UNPINNED: no Go toolchain in the image; the KATs are derived by hand from the
documented behaviour of encoding/json (Go 1.23) and say so."""
import base64
import json
import os

from oracle import gojson as G
from jsoncase import make_requests, make_script, run_model

HERE = os.path.dirname(os.path.abspath(__file__))


def _kats():
    with open(os.path.join(HERE, "golden", "gojson_kats.json")) as f:
        return json.load(f)


def test_string_kats():
    for k in _kats()["strings"]:
        assert G.go_string(bytes.fromhex(k["in_hex"])) == k["out"].encode("latin-1"), k["name"]


def test_time_kats():
    for k in _kats()["times"]:
        assert G.go_time(k["unix_nanos"]) == ('"' + k["out"] + '"').encode(), k


def test_record_kats():
    for k in _kats()["records"]:
        r = dict(k["record"])
        r["body"] = bytes.fromhex(r.pop("body_hex"))
        if r.get("response"):
            r["response"] = dict(r["response"]); r["response"]["body"] = bytes.fromhex(r["response"].pop("body_hex"))
        assert G.marshal_request(r).decode() == k["json"], k["name"]


def test_nil_slice_is_null():
    assert G.marshal_list([]) == b"null" and G.marshal_list(None) == b"null"


def test_output_is_valid_json_and_round_trips():
    """Whatever went in, the output parses as JSON, and parsing gives back the fields with Go's substitutions applied."""
    agents = ["agent-1700000000000000001", "agent-1700000000000000002"]
    reqs = make_requests(11, 300, agents)
    redis, mgr = run_model(reqs, make_script(11, len(reqs), 500))
    for r in reqs:
        rec = redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}")
        js = G.marshal_request(rec)
        js.decode("utf-8")                                       # Marshal never emits invalid UTF-8
        d = json.loads(js)
        assert list(d)[:10] == ["id", "agent_id", "method", "path", "headers", "body", "status", "retry_count", "max_retries", "created_at"]
        assert base64.b64decode(d["body"]) == r.body
        assert d["path"] == G.go_decode(r.path)
        assert d["headers"] == {G.go_decode(k): G.go_decode(v) for k, v in r.headers.items()}
        assert list(d["headers"]) == [G.go_decode(k) for k in sorted(r.headers)]
        assert ("response" in d) == (rec["response"] is not None) == ("processed_at" in d)
        assert ("error" in d) == (rec["retry_count"] > 0)
        # a second round trip is a fixed point (what StoreResponse's Unmarshal + Marshal relies on)
        assert G.marshal_request(G.unmarshal_strings(dict(rec, headers=dict(rec["headers"])))) == G.marshal_request(G.unmarshal_strings(rec))
    for a in agents:
        js = G.marshal_list(mgr.get_pending_requests(a))
        assert js == b"null" or len(json.loads(js)) == len(redis.lrange_all(f"agent:{a}:requests:pending"))


def test_invalid_utf8_changes_form_after_a_round_trip():
    from oracle import model as M
    redis = M.MiniRedis(); mgr = M.Manager(redis)
    mgr.store_request("a", M.HttpRequest("GET", b"/agent/a/\xff", {}, b"", new_id="i", now=0))
    assert b'"path":"/agent/a/\\ufffd"' in G.marshal_request(redis.get("agent:a:requests:i"))
    mgr.mark_request_failed("a", "i", b"e\xff")
    js = G.marshal_request(redis.get("agent:a:requests:i"))
    assert b'"path":"/agent/a/\xef\xbf\xbd"' in js and b'"error":"e\\ufffd"' in js
    mgr.store_response("a", "i", M.HttpResponse(200, {b"K": b"\xff"}, b"", now=0))
    js = G.marshal_request(redis.get("agent:a:requests:i"))
    assert b'"error":"e\xef\xbf\xbd"' in js and b'"headers":{"K":"\\ufffd"}' in js


def test_oracle_reproduces_the_committed_stream():
    """tests/golden/json_stream.json pins the oracle's bytes for the seeded stream (regenerate with make_json_golden.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_json_golden", os.path.join(HERE, "golden", "make_json_golden.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    with open(os.path.join(HERE, "golden", "json_stream.json")) as f:
        want = json.load(f)
    assert mod.build(want["seed"], want["n"], want["n_ops"]) == want
