"""the 24 h TTL of the record keys (SET ... EX 24h, requests.go:106,175,270).  The model is oracle/model.py's
MiniRedis with a clock; the engine sweeps with agr_expire.  After expiry a record is gone for every reader while its id
stays in the lists (Q10), and every SET restarts the TTL (Q11)."""
import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K
from oracle import model as M, gojson as G
from jsoncase import make_requests, records_array, var_batch

pytestmark = pytest.mark.gpu
MODES = {"hash": K.AGR_CFG_PERSISTENCE, "mint": K.AGR_CFG_PERSISTENCE | K.AGR_CFG_MINT_IDS,
         "var": K.AGR_CFG_PERSISTENCE | K.AGR_CFG_VARLEN | K.AGR_CFG_MINT_IDS}
AGENTS = ["agent-1700000000000000001", "agent-1700000000000000002"]
SEC = 1_000_000_000
HOUR = 3600 * SEC
TTL = 24 * HOUR
T0 = 1_700_000_000 * SEC


def outcomes(rows):
    outs = np.zeros(len(rows), dtype=A.outcome_dtype)
    for j, (rid, agent, kind, http, seq) in enumerate(rows):
        outs[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8)
        outs[j]["agent_id"] = agent.encode()
        outs[j]["kind"], outs[j]["http_status"], outs[j]["seq"] = kind, http, seq
    return outs


def check_same(eng, redis, mgr, reqs):
    for r in reqs:
        key = f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"
        try:
            want = G.marshal_request(redis.get(key))
        except M.RedisNil:
            want = None
        assert eng.get_record_json(r.agent_id, r.rid) == want
        rec = eng.get_record_var(r.agent_id, r.rid) if eng.varlen else eng.get_record(r.agent_id, r.rid)
        assert (rec is None) == (want is None)
    for a in AGENTS:
        got, cnt = eng.pending_json(a)
        live = mgr.get_pending_requests(a)                          # skips the expired ones (requests.go:210-213)
        assert got == G.marshal_list(live) and cnt == len(live)
        for which, q in ((K.AGR_LIST_PENDING, "pending"), (K.AGR_LIST_COMPLETED, "completed"), (K.AGR_LIST_FAILED, "failed")):
            ids = [G.format_uuid(bytes(i)) for i in eng.list(a, which)]
            assert ids == redis.lrange_all(f"agent:{a}:requests:{q}"), q   # LRANGE still names the expired records


@pytest.mark.parametrize("mode", sorted(MODES))
def test_ttl_expiry_matches_the_model(mode, tmp_path):
    n = 60
    reqs = make_requests(21, n, AGENTS)
    for i, r in enumerate(reqs):
        r.now = T0 + i * SEC
    kw = dict(slab_rows=1 << 10, max_agents=8, flags=MODES[mode], vslab_bytes=4 << 20)
    eng = A.Engine(**kw)
    try:
        eng.set_agent_state(AGENTS[0], K.AGR_AGENT_RUNNING); eng.set_agent_state(AGENTS[1], K.AGR_AGENT_STOPPED)
        if eng.varlen:
            _, ids, _ = eng.ingest_var(*var_batch(reqs))
        else:
            out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
            eng.ingest_ex(records_array(reqs), out, ids)
        for r, i in zip(reqs, ids):
            r.rid = bytes(i)
        redis = M.MiniRedis(); mgr = M.Manager(redis)
        for r in reqs:
            redis.now = r.now
            mgr.store_request(r.agent_id, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.now))
        # SETs that restart the TTL: every 3rd record answered one hour in, every 5th failed two hours in
        ops = []
        for i, r in enumerate(reqs):
            if i % 3 == 0:
                ops.append((r, "resp", T0 + HOUR + i))
            elif i % 5 == 0:
                ops.append((r, "err", T0 + 2 * HOUR + i))
        for r, kind, t in ops:
            redis.now = t
            if kind == "resp":
                mgr.store_response(r.agent_id, G.format_uuid(r.rid), M.HttpResponse(200, {}, b"", now=t))
            else:
                mgr.mark_request_failed(r.agent_id, G.format_uuid(r.rid), "transport error")
        eng.complete(outcomes([(r.rid, r.agent_id, K.AGR_OUT_RESPONSE if kind == "resp" else K.AGR_OUT_ERROR, 200 if kind == "resp" else 0, t)
                               for r, kind, t in ops]))
        # nothing has expired one second before the first deadline
        redis.now = T0 + TTL - 1
        assert eng.expire(redis.now, TTL) == 0
        check_same(eng, redis, mgr, reqs)
        # 24 h + 25 s after the first arrival: the untouched records among the first 26 are gone
        redis.now = T0 + TTL + 25 * SEC
        gone = eng.expire(redis.now, TTL)
        assert gone == sum(1 for i in range(26) if i % 3 and i % 5)
        check_same(eng, redis, mgr, reqs)
        # an outcome for an expired record: "failed to get request" on both sides
        victim = reqs[1]
        with pytest.raises(KeyError):
            mgr.store_response(victim.agent_id, G.format_uuid(victim.rid), M.HttpResponse(200, now=redis.now))
        res = eng.complete(outcomes([(victim.rid, victim.agent_id, K.AGR_OUT_RESPONSE, 200, redis.now)]))
        assert res[0] == K.AGR_ENOTFOUND
        # the replay scan skips expired records of the running agent
        disp = eng.replay_scan_var()[0] if eng.varlen else eng.replay_scan(with_records=False)[0]
        want = [G.format_uuid(r.rid) for r in reqs if r.agent_id == AGENTS[0]]
        live = {q["id"] for q in mgr.get_pending_requests(AGENTS[0]) if q["retry_count"] < q["max_retries"]}
        assert [G.format_uuid(bytes(d["request_id"])) for d in disp] == [i for i in want if i in live]
        # survive a restart, then the records answered at T0 + 1 h expire exactly 24 h after that SET
        path = str(tmp_path / "ttl.snap")
        eng.snapshot(path); eng.close()
        eng = A.Engine(restore_from=path, **kw)
        redis.now = T0 + HOUR + TTL + n
        assert eng.expire(redis.now, TTL) > 0
        check_same(eng, redis, mgr, reqs)
        assert eng.verify()[1] == 0
    finally:
        eng.close()
