"""AGR_CFG_RING: the slab as a ring.  A shard with 4096 rows takes three times that many requests while agr_expire +
agr_reclaim release the rows at the tail; every observable stays equal to oracle/model.py's (with its key TTL), FIFO order
holds across the wrap, ids of released rows stop resolving, and a restart in the middle is invisible."""
import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K
from oracle import model as M, gojson as G
from jsoncase import make_requests, records_array, var_batch

pytestmark = pytest.mark.gpu
AGENTS = ["agent-1700000000000000001", "agent-1700000000000000002", "agent-1700000000000000003"]
SEC = 1_000_000_000
HOUR = 3600 * SEC
T0 = 1_700_000_000 * SEC
STEP = 3 * HOUR          # one round of traffic every three hours
TTL = 24 * HOUR          # the reference's key TTL (requests.go:106); oracle/model.py uses the same
FLAGS = K.AGR_CFG_PERSISTENCE | K.AGR_CFG_MINT_IDS | K.AGR_CFG_RING


def outcomes(rows):
    outs = np.zeros(len(rows), dtype=A.outcome_dtype)
    for j, (rid, agent, kind, http, seq) in enumerate(rows):
        outs[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8)
        outs[j]["agent_id"] = agent.encode()
        outs[j]["kind"], outs[j]["http_status"], outs[j]["seq"] = kind, http, seq
    return outs


@pytest.mark.parametrize("mode", ["fixed", "var", "hash"])
def test_ring_runs_past_its_capacity(mode, tmp_path):
    R, per_round, rounds = 4096, 300, 42
    var = mode == "var"
    # variable-length records: the byte slab (6 MiB) is a ring too and laps about as often as the rows do
    RESP = 64 << 10                                            # the response / error byte slab is a ring too: 64 KiB, lapped twice
    kw = dict(slab_rows=R, max_agents=8, flags=(FLAGS & ~K.AGR_CFG_MINT_IDS if mode == "hash" else FLAGS) | (K.AGR_CFG_VARLEN if var else 0),
              vslab_bytes=6 << 20, resp_bytes=RESP if mode == "fixed" else 0)
    eng = A.Engine(**kw)
    redis = M.MiniRedis(); mgr = M.Manager(redis)
    agents = M.AgentStore(redis); proxy = M.Proxy(redis, agents); worker = M.ReplayWorker(redis, agents, proxy)
    status = {AGENTS[0]: "running", AGENTS[1]: "stopped", AGENTS[2]: "stopped"}
    for a, s in status.items():
        eng.set_agent_state(a, s); agents.save(a, s)
    known = []                                    # (request, logical row)
    rng = np.random.default_rng(7)
    wrapped_with_backlog = 0
    bytes_in = 0
    stored_bytes = 0
    try:
        for rnd in range(rounds):
            now = T0 + rnd * STEP
            redis.now = now
            reqs = make_requests(1000 + rnd, per_round, AGENTS, max_payload=7000 if var else 416, big_bodies=var)
            for i, r in enumerate(reqs):
                r.now = now + i
            if var:
                blob, offs = var_batch(reqs)
                bytes_in += int(offs[-1])
                _, ids, first = eng.ingest_var(blob, offs)
            else:
                out = np.zeros(per_round, dtype=A.verdict_dtype); ids = np.zeros((per_round, 16), dtype=np.uint8)
                first = eng.ingest_ex(records_array(reqs), out, ids)
            for i, (r, rid) in enumerate(zip(reqs, ids)):
                r.rid = bytes(rid)
                known.append((r, first + i))
                mgr.store_request(r.agent_id, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.now))
            # the running agent answers most of its requests of this round; a few fail
            ops = []
            for r in reqs:
                if r.agent_id == AGENTS[0]:
                    u = rng.random()
                    if u < 0.8:
                        ops.append((r, K.AGR_OUT_RESPONSE, 200))
                    elif u < 0.9:
                        ops.append((r, K.AGR_OUT_ERROR, 0))
            t = now + 10 * SEC
            redis.now = t
            with_bytes = mode == "fixed"                       # responses and error texts carry bytes in this mode
            texts = {}
            for r, kind, code in ops:
                if kind == K.AGR_OUT_RESPONSE:
                    body = bytes(rng.integers(0, 256, int(rng.integers(0, 48)), dtype=np.uint8)) if with_bytes else b""
                    hdrs = {b"Server": b"x<y>"} if with_bytes and rng.random() < 0.5 else {}
                    texts[r.rid] = (hdrs, body)
                    mgr.store_response(r.agent_id, G.format_uuid(r.rid), M.HttpResponse(code, dict(hdrs), body, now=t))
                else:
                    text = (b"EOF", b"read tcp: connection reset by peer")[int(rng.integers(0, 2))] if with_bytes else b"transport error"
                    texts[r.rid] = text
                    mgr.mark_request_failed(r.agent_id, G.format_uuid(r.rid), text if with_bytes else "transport error")
            res = eng.complete(outcomes([(r.rid, r.agent_id, kind, code, t) for r, kind, code in ops]))
            assert (res == 0).all()
            if with_bytes:
                for r, kind, code in ops:
                    if kind == K.AGR_OUT_RESPONSE:
                        hdrs, body = texts[r.rid]
                        flat = b"".join(k + b": " + v + b"\n" for k, v in sorted(hdrs.items()))
                        assert eng.store_response(r.agent_id, r.rid, flat, body)
                        stored_bytes += (len(flat) + len(body) + 15) // 16 * 16
                    else:
                        assert eng.store_error_text(r.agent_id, r.rid, texts[r.rid])
                        stored_bytes += (len(texts[r.rid]) + 15) // 16 * 16
            # agent 3 comes up every 5th round for one tick: its backlog (possibly lying across the wrap) replays in FIFO order
            if rnd % 5 == 4:
                eng.set_agent_state(AGENTS[2], "running"); agents.save(AGENTS[2], "running")
                disp = eng.replay_scan_var()[0] if var else eng.replay_scan(with_records=False)[0]
                want = worker.process_agents(lambda a, q: ("response", 200), now=t)
                got = [(AGENTS[int(d["agent_slot"])], G.format_uuid(bytes(d["request_id"]))) for d in disp]
                assert got == want and len(got) > 50
                rows = [int(d["rid"]) for d in disp if int(d["agent_slot"]) == 2]
                assert rows == sorted(rows)                                   # arrival order == ascending logical row
                if (rows[0] % R) > (rows[-1] % R):
                    wrapped_with_backlog += 1                                 # the backlog straddled the physical wrap
                # what the Go side does with a dispatched request (replay_worker.go:120-163): through the proxy, then StoreResponse
                outs = []
                for a, rid_s in got:
                    raw = bytes.fromhex(rid_s.replace("-", ""))
                    outs.append((raw, a, K.AGR_OUT_RESPONSE, 200, t)); outs.append((raw, a, K.AGR_OUT_RESPONSE, 200, t))
                assert (eng.complete(outcomes(outs)) == 0).all()
                eng.set_agent_state(AGENTS[2], "stopped"); agents.save(AGENTS[2], "stopped")
            # TTL, then hand the dead rows at the tail back
            redis.now = now + 30 * SEC
            eng.expire(redis.now, TTL)
            eng.reclaim()
            st = eng.stats()
            tail = st["rows_tail"]
            assert st["rows_used"] - tail <= R
            if rnd == 20:                                                     # a restart in the middle
                path = str(tmp_path / "ring.snap")
                eng.snapshot(path); eng.close()
                eng = A.Engine(restore_from=path, **kw)
                assert eng.stats()["rows_tail"] == tail
            if rnd % 3 == 2 or rnd == rounds - 1:
                # records: live ones read the same, released / expired ones are gone on both sides
                sample = known[-900:] + known[:: max(1, len(known) // 200)]
                for r, lrow in sample:
                    try:
                        want = G.marshal_request(redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"))
                    except M.RedisNil:
                        want = None
                    assert eng.get_record_json(r.agent_id, r.rid) == want
                    if lrow < tail:
                        assert want is None
                live_ids = {G.format_uuid(r.rid) for r, lrow in known if lrow >= tail}
                for a in AGENTS:
                    got, cnt = eng.pending_json(a)
                    assert got == G.marshal_list(mgr.get_pending_requests(a))
                    for which, q in ((K.AGR_LIST_PENDING, "pending"), (K.AGR_LIST_COMPLETED, "completed"), (K.AGR_LIST_FAILED, "failed")):
                        ids_e = [G.format_uuid(bytes(i)) for i in eng.list(a, which, cap=1 << 14)]
                        assert ids_e == [i for i in redis.lrange_all(f"agent:{a}:requests:{q}") if i in live_ids], (rnd, a, q)
        st = eng.stats()
        assert st["rows_used"] > 3 * R - 2 * per_round and st["rows_tail"] > 2 * R
        assert wrapped_with_backlog >= 1
        assert not var or bytes_in > 2 * (6 << 20)                            # the byte ring lapped too
        assert mode != "fixed" or stored_bytes > 2 * RESP                     # and so did the response / error byte ring
        assert eng.verify()[1] == 0
    finally:
        eng.close()


def test_ring_full_then_reclaimed():
    R = 1024
    with A.Engine(slab_rows=R, max_agents=4, flags=FLAGS) as eng:
        eng.set_agent_state(AGENTS[0], "stopped")
        reqs = make_requests(3, 256, AGENTS[:1])
        for i, r in enumerate(reqs):
            r.now = T0 + i
        recs = records_array(reqs)
        out = np.zeros(256, dtype=A.verdict_dtype); ids = np.zeros((256, 16), dtype=np.uint8)
        firsts = [eng.ingest_ex(recs, out, ids) for _ in range(4)]
        assert firsts == [0, 256, 512, 768]
        with pytest.raises(A.AgrError) as e:
            eng.ingest_ex(recs, out, ids)
        assert e.value.code == K.AGR_ENOSPC
        old = bytes(eng.mint_ids(0, 1)[0])
        assert eng.get_record_json(AGENTS[0], old) is not None
        assert eng.reclaim() == 0                                             # everything still holds a record
        assert eng.expire(T0 + 10 * HOUR, HOUR) == 1024
        assert eng.reclaim() == 1024
        assert eng.get_record_json(AGENTS[0], old) is None                    # ids of released rows never resolve again
        assert eng.ingest_ex(recs, out, ids) == 1024                          # row ids keep counting
        assert eng.get_record_json(AGENTS[0], bytes(ids[5])) is not None
        assert eng.pending_json(AGENTS[0])[1] == 256
        # row ranges are addressed by row id wherever the rows physically are (1024.. live in physical rows 0..)
        blob, offs = eng.rows_json(1024, 256)
        for i in (0, 1, 100, 255):
            assert blob[int(offs[i]):int(offs[i + 1])] == eng.get_record_json(AGENTS[0], bytes(ids[i]))
        st_words = eng.debug_read("state", 1024, 256)
        assert (st_words & 0x20).all()                                        # ST_STORED
        # a batch that would wrap skips to the start of the slab
        small = recs[:200]
        o2 = np.zeros(200, dtype=A.verdict_dtype); i2 = np.zeros((200, 16), dtype=np.uint8)
        f = [eng.ingest_ex(small, o2, i2) for _ in range(3)]                  # 1280, 1480, 1680 -> next would end at 2080 > 2048
        assert f == [1280, 1480, 1680]
        assert eng.expire(T0 + 100 * HOUR, HOUR) > 0 and eng.reclaim() > 0
        assert eng.ingest_ex(small, o2, i2) == 2 * R                          # 168 rows skipped at the end of the lap
        assert eng.ingest_ex(small, o2, i2) == 2 * R + 200
        st = eng.stats()
        assert st["rows_used"] == 2 * R + 400 and st["rows_tail"] == 1880


def test_ring_with_caller_ids_forgets_released_ids():
    """hash-id mode: an id whose row was released leaves the dedupe index (rebuilt at every release) and may be stored again."""
    with A.Engine(slab_rows=1024, max_agents=4, flags=K.AGR_CFG_PERSISTENCE | K.AGR_CFG_RING) as eng:
        eng.set_agent_state(AGENTS[0], "stopped")
        reqs = make_requests(8, 100, AGENTS[:1])
        for i, r in enumerate(reqs):
            r.now = T0 + i
        recs = records_array(reqs)
        out = np.zeros(100, dtype=A.verdict_dtype); ids = np.zeros((100, 16), dtype=np.uint8)
        eng.ingest_ex(recs, out, ids)
        assert (out["flags"] & K.AGR_VF_STORED).all()
        eng.ingest_ex(recs, out, ids)
        assert (out["flags"] & K.AGR_VF_DUP_ID).all()                         # same ids again: refused while the first are live
        assert eng.expire(T0 + 10 * HOUR, HOUR) == 100 and eng.reclaim() == 200
        assert eng.get_record_json(AGENTS[0], reqs[0].rid) is None
        eng.ingest_ex(recs, out, ids)
        assert (out["flags"] & K.AGR_VF_STORED).all()
        assert eng.get_record_json(AGENTS[0], reqs[0].rid) is not None
