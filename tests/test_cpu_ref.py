"""The C restatement (oracle/cpu_ref.c: the large-scale checker and the timed CPU baseline) against the Python
oracle and the hand-derived KATs, driven through the same scenario driver the GPU engine uses.  CPU only."""
import numpy as np
import pytest

import agentainer_lab_b200 as A
from kats import SCENARIOS, load_golden
from oracle.cpu_ref import CRef
from scenario import run_oracle, run_engine, assert_same, random_scenario, Req, rid_of, make_records

GOLD, _ = load_golden()


@pytest.mark.parametrize("kat", sorted(SCENARIOS))
def test_kats(kat):
    with CRef() as c:
        got = run_engine(c, SCENARIOS[kat])
    exp = GOLD[kat]
    assert got.verdicts == exp["verdicts"] and got.ticks == exp["ticks"] and got.records == exp["records"]
    for a, qs in exp["lists"].items():
        assert got.lists[a] == qs
    assert_same(run_oracle(SCENARIOS[kat]), got)


@pytest.mark.parametrize("seed", range(8))
def test_random_streams(seed):
    ev = random_scenario(seed, n_events=400, n_agents=2 + seed % 6)
    with CRef() as c:
        assert_same(run_oracle(ev), run_engine(c, ev, rng=np.random.default_rng(seed)))


def test_json_round_trip_preserves_the_record():
    """SET stores json.Marshal(request); GET + Unmarshal must give the same request back (body via base64,
    headers via the JSON map, HTML-escaped characters)."""
    r = Req("agent-1", rid_of(1), 123456789, body=b'{"message":"<b>&\\"q\\"\n\x01 \xc3\xa9"}',
            headers={"Content-Type": "application/json", "X-Odd": 'a"b<c>&d'})
    recs = make_records([r])
    with CRef() as c:
        c.set_agent_state("agent-1", "stopped")
        c.ingest(recs)
        back = c.get_record("agent-1", rid_of(1))
    for f in ("request_id", "agent_id", "seq", "path_len", "hdr_len", "body_len", "max_retries"):
        assert np.array_equal(back[f], recs[0][f]), f
    assert bytes(back["payload"]) == bytes(recs[0]["payload"])
    assert int(back["status"]) == 1 and int(back["flags"]) >> 8 == 2


def test_config1_stream_c_port_equals_python_oracle():
    """BASELINE config 1: 10 k synthetic 512 B POST /agent/<id>/chat records, 16 agent ids, half the agents stopped for
    the first 5 000 records then started, one tick: dedupe + replay order.  C port == Python oracle."""
    from scenario import config1_events, synth_to_req
    ev, recs = config1_events()
    ref = run_oracle(ev)
    assert sum(len(t) for t in ref.ticks) > 1000
    with CRef() as c:
        got = run_engine(c, ev, max_batch=4096)
    assert_same(ref, got)
    # the scenario's records are byte-identical to the synthetic stream's
    assert make_records([synth_to_req(recs[7])]).tobytes() == recs[7:8].tobytes()


def test_two_restatements_of_the_wire_form_agree():
    """oracle/cpu_ref.c keeps every record as JSON TEXT in its keyspace and really parses and re-marshals it on each update
    (like the Go code); oracle/model.py + oracle/gojson.py keep dicts and carry the round trip in their types.  On a stream
    with awkward strings (quotes, HTML characters, control bytes, invalid UTF-8, U+2028) and several updates per record the two
    must produce the same bytes."""
    import numpy as np
    from jsoncase import make_requests, records_array
    from oracle import gojson as G, model as M
    from oracle.cpu_ref import CRef, record_dtype
    from agentainer_lab_b200 import constants as K, outcome_dtype
    agents = ["agent-1700000000000000001", "agent-1700000000000000002"]
    reqs = make_requests(31, 300, agents, max_payload=330)         # room for invalid bytes growing to EF BF BD after a round trip
    for r in reqs:                                                 # the C port parses times back: keep them in its supported range
        r.now = 1_700_000_000_000_000_000 + (r.now % 10**15)
    c = CRef()
    redis = M.MiniRedis(); mgr = M.Manager(redis)
    for a in agents:
        c.set_agent_state(a, "running")
    recs = records_array(reqs).astype(record_dtype)
    c.ingest(recs)
    for r in reqs:
        mgr.store_request(r.agent_id, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.now))
    import random
    rng = random.Random(5)
    ops = [(rng.randrange(len(reqs)), rng.random() < 0.6, 1_700_000_000_000_000_000 + rng.randrange(10**12) * 1000 + rng.choice([0, 0, 1, 500]))
           for _ in range(500)]
    for i, is_resp, t in ops:
        r = reqs[i]
        out = np.zeros(1, dtype=outcome_dtype)
        out[0]["request_id"] = np.frombuffer(r.rid, dtype=np.uint8); out[0]["agent_id"] = r.agent_id.encode()
        out[0]["kind"], out[0]["http_status"], out[0]["seq"] = (K.AGR_OUT_RESPONSE, 200, t) if is_resp else (K.AGR_OUT_ERROR, 0, t)
        c.complete(out)
        if is_resp:
            mgr.store_response(r.agent_id, G.format_uuid(r.rid), M.HttpResponse(200, {}, b"", now=t))
        else:
            mgr.mark_request_failed(r.agent_id, G.format_uuid(r.rid), "transport error")
    checked = 0
    for r in reqs:
        want = G.marshal_request(redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"))
        assert c.get_json(r.agent_id, r.rid) == want, (r, c.get_json(r.agent_id, r.rid), want)
        checked += 1
    assert checked == 300
    for a in agents + ["agent-nobody"]:
        assert c.pending_json(a) == G.marshal_list(mgr.get_pending_requests(a))
    c.close()


def test_two_restatements_agree_on_the_key_ttl():
    """SET ... EX 24h in both restatements: records vanish 24 h after their LAST SET, their ids stay in the lists, updates of
    a vanished record fail with "failed to get request", GetPendingRequests skips them."""
    import random
    import numpy as np
    from jsoncase import make_requests, records_array
    from oracle import gojson as G, model as M
    from oracle.cpu_ref import CRef, record_dtype
    from agentainer_lab_b200 import constants as K, outcome_dtype
    SEC = 1_000_000_000; HOUR = 3600 * SEC; T0 = 1_700_000_000 * SEC
    agents = ["agent-1700000000000000001", "agent-1700000000000000002"]
    c = CRef(); redis = M.MiniRedis(); mgr = M.Manager(redis)
    c.set_agent_state(agents[0], "running"); c.set_agent_state(agents[1], "stopped")
    rng = random.Random(9)
    known = []
    for rnd in range(16):                                    # a batch every 4 hours, over 64 hours
        now = T0 + rnd * 4 * HOUR
        c.set_now(now); redis.now = now
        reqs = make_requests(200 + rnd, 40, agents, max_payload=330)
        for i, r in enumerate(reqs):
            r.now = now + i
            mgr.store_request(r.agent_id, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.now))
        c.ingest(records_array(reqs).astype(record_dtype))
        known += reqs
        t = now + HOUR
        c.set_now(t); redis.now = t
        for r in rng.sample(known, 30):                      # updates hit fresh and long-gone records alike
            out = np.zeros(1, dtype=outcome_dtype)
            out[0]["request_id"] = np.frombuffer(r.rid, dtype=np.uint8); out[0]["agent_id"] = r.agent_id.encode()
            is_resp = rng.random() < 0.5
            out[0]["kind"], out[0]["http_status"], out[0]["seq"] = (K.AGR_OUT_RESPONSE, 200, t) if is_resp else (K.AGR_OUT_ERROR, 0, t)
            res = c.complete(out)
            try:
                if is_resp:
                    mgr.store_response(r.agent_id, G.format_uuid(r.rid), M.HttpResponse(200, {}, b"", now=t))
                else:
                    mgr.mark_request_failed(r.agent_id, G.format_uuid(r.rid), "transport error")
                assert res[0] == 0
            except KeyError:
                assert res[0] == K.AGR_ENOTFOUND               # "failed to get request"
        gone = 0
        for r in known:
            try:
                want = G.marshal_request(redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"))
            except M.RedisNil:
                want = None; gone += 1
            assert c.get_json(r.agent_id, r.rid) == want
        for a in agents:
            assert c.pending_json(a) == G.marshal_list(mgr.get_pending_requests(a))
            for which, q in ((0, "pending"), (1, "completed"), (2, "failed")):
                assert [G.format_uuid(bytes(i)) for i in c.list(a, which, cap=1 << 12)] == redis.lrange_all(f"agent:{a}:requests:{q}")
    assert gone > 200
    c.close()
