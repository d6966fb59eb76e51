"""Round-2 behaviour through the C-ABI: malformed lengths, reserved-but-unfilled rows, log overflow, long outcome chains on
one record inside one agr_complete batch (against oracle/model.py), the dedupe index after the ring has wrapped, and the
single-request front end (lock-free ring + resident service kernel) under threads, in both id modes."""
import os
import threading

import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K
from oracle import model as M, gojson as G
from scenario import Req, rid_of, make_records

pytestmark = pytest.mark.gpu
MINT = K.AGR_CFG_PERSISTENCE | K.AGR_CFG_MINT_IDS
HASH = K.AGR_CFG_PERSISTENCE


def outcomes(rows):
    outs = np.zeros(len(rows), dtype=A.outcome_dtype)
    for j, (rid, agent, kind, http, seq) in enumerate(rows):
        outs[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8)
        outs[j]["agent_id"] = agent.encode()
        outs[j]["kind"], outs[j]["http_status"], outs[j]["seq"] = kind, http, seq
    return outs


@pytest.mark.parametrize("flags", [MINT, HASH])
def test_malformed_lengths_are_rejected_not_trusted(flags):
    """A record whose path_len + hdr_len + body_len exceeds its payload is never persisted (AGR_VF_BAD_LEN): the request
    still gets its verdict (StoreRequest's error path, server.go:511-514), nothing reads past the record, and the JSON
    view of the batch's rows encodes such rows as null."""
    rng = np.random.default_rng(3)
    with A.Engine(slab_rows=1 << 12, max_agents=8, flags=flags) as eng:
        eng.set_agent_state("agent-1", "running"); eng.set_agent_state("agent-2", "stopped")
        n = 512
        recs = make_records([Req("agent-1" if i % 2 else "agent-2", rid_of(i + 1), i + 1) for i in range(n)])
        bad = np.zeros(n, dtype=bool)
        for i in range(0, n, 3):
            which = int(rng.integers(0, 4))
            if which == 0:
                recs[i]["body_len"] = int(rng.integers(417, 1 << 31))
            elif which == 1:
                recs[i]["path_len"] = 0xffff
            elif which == 2:
                recs[i]["hdr_len"] = int(rng.integers(400, 0x10000))
            else:
                recs[i]["body_len"] = 0xffffffff
            bad[i] = int(recs[i]["path_len"]) + int(recs[i]["hdr_len"]) + int(recs[i]["body_len"]) > 416
        assert bad.sum() > 100
        out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
        first = eng.ingest_ex(recs, out, ids)
        flagged = (out["flags"] & K.AGR_VF_BAD_LEN) != 0
        assert (flagged == bad).all()
        assert ((out["flags"][bad] & (K.AGR_VF_STORED | K.AGR_VF_TRACKED)) == 0).all()
        assert ((out["flags"][~bad] & K.AGR_VF_STORED) != 0).all()
        running = np.array([i % 2 == 1 for i in range(n)])
        assert (out["code"][bad & running] == K.AGR_V_FORWARD).all()          # forwarded untracked
        assert (out["code"][bad & ~running] == K.AGR_V_UNAVAILABLE).all()     # 503: nothing was queued
        s = eng.stats()
        assert s["malformed"] == int(bad.sum()) and s["stored"] == int((~bad).sum())
        import json
        docs = json.loads(eng.rows_json(first, n, as_array=True)[0])
        assert [d is None for d in docs] == bad.tolist()
        assert len(eng.list("agent-2", K.AGR_LIST_PENDING)) == int((~bad & ~running).sum())


def test_reserved_rows_are_not_skipped_by_a_scan_or_a_reclaim():
    """agr_reserve_rows, then a replay tick and a reclaim BEFORE agr_ingest_rows: the rows filled in later must still be
    replayed (the scan's low-water mark and the ring's tail stop at the first reserved-but-unfilled row)."""
    flags = MINT | K.AGR_CFG_RING
    with A.Engine(slab_rows=4096, max_agents=4, max_batch=512, flags=flags) as eng:
        names = [A.synth_agent_id(k) for k in range(2)]
        eng.set_agent_state(names[0], "stopped"); eng.set_agent_state(names[1], "stopped")
        first = eng.reserve_rows(256)
        eng.synth_fill_rows(0, first, 256, seed=5, n_agents=2)
        assert len(eng.replay_scan(with_records=False)[0]) == 0              # nothing ingested yet
        assert eng.reclaim() == 0                                            # reserved rows are not "dead"
        v = eng.ingest_rows(first, 256)
        assert (v["code"] == K.AGR_V_QUEUED).all()
        eng.set_agent_state(names[0], "running"); eng.set_agent_state(names[1], "running")
        disp, _ = eng.replay_scan(with_records=False)
        assert len(disp) == 256 and sorted(int(d["rid"]) for d in disp) == list(range(first, first + 256))


def test_log_overflow_is_reported():
    """completed-list pushes that do not fit agr_config.log_entries: agr_complete says AGR_ENOSPC (the transitions were applied)."""
    with A.Engine(slab_rows=256, max_agents=4, log_entries=64, flags=MINT) as eng:
        eng.set_agent_state("agent-1", "running")
        recs = make_records([Req("agent-1", rid_of(i + 1), i + 1) for i in range(100)])
        out = np.zeros(100, dtype=A.verdict_dtype); ids = np.zeros((100, 16), dtype=np.uint8)
        eng.ingest_ex(recs, out, ids)
        assert (eng.complete(outcomes([(bytes(ids[i]), "agent-1", K.AGR_OUT_RESPONSE, 200, 5) for i in range(60)])) == 0).all()
        with pytest.raises(A.AgrError) as e:
            eng.complete(outcomes([(bytes(ids[i]), "agent-1", K.AGR_OUT_RESPONSE, 200, 6) for i in range(60, 100)]))
        assert e.value.code == K.AGR_ENOSPC
        s = eng.stats()
        assert s["log_overflow"] == 1 and s["completed_log_len"] == 64 and s["completions"] == 100


@pytest.mark.parametrize("flags", [MINT, HASH])
def test_many_outcomes_of_one_record_in_one_batch_apply_in_call_order(flags):
    """K2 chains: 1, 2, 5, 9 and 40 outcomes on the same record inside ONE agr_complete (short chains are sorted in registers,
    long ones take the selection path) interleaved with outcomes of other records — final record states and the completed /
    failed lists equal the reference's calls made one after the other."""
    rng = np.random.default_rng(11)
    agent = "agent-1700000000000000001"
    with A.Engine(slab_rows=1 << 10, max_agents=4, flags=flags) as eng:
        redis = M.MiniRedis(); mgr = M.Manager(redis)
        eng.set_agent_state(agent, "running")
        n = 64
        reqs = [Req(agent, rid_of(i + 1), i + 1) for i in range(n)]
        out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
        eng.ingest_ex(make_records(reqs), out, ids)
        for r, rid in zip(reqs, ids):
            r.rid = bytes(rid)
            mgr.store_request(agent, M.HttpRequest(r.method, r.path, dict(r.headers), r.body, new_id=G.format_uuid(r.rid), now=r.seq))
        ops = []
        for target, k in ((3, 1), (7, 2), (11, 5), (19, 9), (23, 40)):
            for _ in range(k):
                ops.append((target, K.AGR_OUT_ERROR if rng.random() < 0.6 else K.AGR_OUT_RESPONSE))
        for i in range(30, 60):
            ops.append((i, K.AGR_OUT_RESPONSE))
        order = rng.permutation(len(ops))
        ops = [ops[i] for i in order]
        batch = []
        for t, (i, kind) in enumerate(ops):
            r = reqs[i]
            batch.append((r.rid, agent, kind, 200 if kind == K.AGR_OUT_RESPONSE else 0, 1000 + t))
            redis.now = 1000 + t
            if kind == K.AGR_OUT_RESPONSE:
                mgr.store_response(agent, G.format_uuid(r.rid), M.HttpResponse(200, {}, b"", now=1000 + t))
            else:
                mgr.mark_request_failed(agent, G.format_uuid(r.rid), "transport error")
        assert (eng.complete(outcomes(batch)) == 0).all()
        for name, which in (("pending", K.AGR_LIST_PENDING), ("completed", K.AGR_LIST_COMPLETED), ("failed", K.AGR_LIST_FAILED)):
            got = [G.format_uuid(bytes(x)) for x in eng.list(agent, which)]
            assert got == redis.lrange_all(f"agent:{agent}:requests:{name}"), name
        for r in reqs:
            rec = eng.get_record(agent, r.rid)
            want = redis.get(f"agent:{agent}:requests:{G.format_uuid(r.rid)}")
            assert K.STATUS_NAMES[int(rec["status"])] == want["status"] and int(rec["retry_count"]) == want["retry_count"]


def test_dedupe_index_orders_rows_by_arrival_after_the_ring_wrapped():
    """Caller-supplied ids on a ring that has wrapped: a LATER arrival sits at a LOWER physical row than the original.  The
    original must keep owning the id (the duplicate is refused), and a replay-flagged request naming the original is KNOWN."""
    flags = HASH | K.AGR_CFG_RING
    R = 1024
    agent = "agent-1700000000000000001"
    with A.Engine(slab_rows=R, max_agents=4, max_batch=256, flags=flags) as eng:
        eng.set_agent_state(agent, "stopped")
        T = 1000
        seq = 0
        # three batches of 256 expire and are released, so the tail moves to 768; 200 more land at rows 768..967
        for b in range(3):
            recs = make_records([Req(agent, rid_of(10_000 + b * 256 + i), seq + i) for i in range(256)])
            eng.ingest(recs); seq += 256
        eng.expire(seq + T, T)
        assert eng.reclaim() == 768
        orig = [Req(agent, rid_of(50_000 + i), 10 * T + i) for i in range(200)]
        v, first = eng.ingest(make_records(orig))
        assert first == 768 and ((v["flags"] & K.AGR_VF_STORED) != 0).all()
        # the next batch does not fit before the end of the slab: it wraps to physical row 0 (logical 1024)
        later = [Req(agent, orig[i].rid, 10 * T + 500 + i) for i in range(100)]                                  # same ids again
        later += [Req(agent, rid_of(60_000 + i), 10 * T + 600 + i, replay=True, replay_of=orig[i].rid) for i in range(100)]
        v2, first2 = eng.ingest(make_records(later))
        assert first2 == 1024 and first2 % R < first % R                                                        # wrapped
        assert ((v2["flags"][:100] & K.AGR_VF_DUP_ID) != 0).all() and ((v2["flags"][:100] & K.AGR_VF_STORED) == 0).all()
        assert ((v2["flags"][100:] & K.AGR_VF_KNOWN) != 0).all()
        # the originals still resolve to THEIR rows
        got = eng.get_record(agent, orig[5].rid)
        assert int(got["seq"]) == orig[5].seq
        # an in-batch duplicate pair straddling nothing special still resolves lowest-arrival-wins after the wrap
        pair = [Req(agent, rid_of(70_000), 20 * T), Req(agent, rid_of(70_000), 20 * T + 1)]
        v3, _ = eng.ingest(make_records(pair))
        assert [bool(int(x) & K.AGR_VF_STORED) for x in v3["flags"]] == [True, False]


@pytest.mark.parametrize("flags", [MINT, HASH])
def test_single_request_front_end_under_threads(flags):
    """AGR_CFG_COMBINE: 32 threads x 80 single-request calls (ingest, then complete / fail for the forwarded ones) through the
    lock-free ring and the resident service kernel.  Every request gets its own row; per-agent FIFO is the row order; the
    completed / failed lists and every record's final state equal the reference's calls replayed in row order."""
    nt, per = 32, 80
    agents = ["agent-%d" % k for k in range(4)]
    with A.Engine(slab_rows=1 << 14, max_agents=8, flags=flags | K.AGR_CFG_COMBINE) as eng:
        for k, a in enumerate(agents):
            eng.set_agent_state(a, "stopped" if k % 2 else "running")
        results = [[] for _ in range(nt)]
        errors = []

        def worker(t):
            try:
                for i in range(per):
                    a = agents[(t + i) % 4]
                    rec = make_records([Req(a, rid_of(1 + t * per + i), 1 + t * per + i)])
                    out, ids = np.zeros(1, dtype=A.verdict_dtype), np.zeros((1, 16), dtype=np.uint8)
                    first = eng.ingest_ex(rec, out, ids)
                    kind = None
                    if int(out[0]["code"]) == K.AGR_V_FORWARD:
                        kind = K.AGR_OUT_ERROR if (t + i) % 5 == 0 else K.AGR_OUT_RESPONSE
                        o = outcomes([(bytes(ids[0]), a, kind, 200, 7)])
                        assert list(eng.complete(o)) == [0]
                        if (t + i) % 7 == 0:                                   # unknown id: a miss, not an error
                            assert list(eng.complete(outcomes([(rid_of(10 ** 9 + t * per + i), a, K.AGR_OUT_RESPONSE, 200, 7)]))) == [K.AGR_ENOTFOUND]
                    results[t].append((a, int(out[0]["code"]), first, bytes(ids[0]), kind))
            except Exception as e:                                               # noqa: BLE001
                errors.append(repr(e))

        ths = [threading.Thread(target=worker, args=(t,)) for t in range(nt)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        assert not errors, errors[:3]
        flat = [x for r in results for x in r]
        assert len({x[2] for x in flat}) == nt * per                           # every request got its own row
        assert all((code == K.AGR_V_QUEUED) == (agents.index(a) % 2 == 1) for a, code, _, _, _ in flat)
        if flags & K.AGR_CFG_MINT_IDS:
            assert all(bytes(eng.mint_ids(first, 1)[0]) == rid for _, _, first, rid, _ in flat)
        else:
            assert all(rid == rid_of(1 + t * per + i) for t in range(nt) for i, (_, _, _, rid, _) in enumerate(results[t]))
        for k, a in enumerate(agents):
            mine = sorted((first, rid, kind) for aa, _, first, rid, kind in flat if aa == a)
            pend = [bytes(x) for x in eng.list(a, K.AGR_LIST_PENDING)]
            comp = [bytes(x) for x in eng.list(a, K.AGR_LIST_COMPLETED)]
            if k % 2:
                assert pend == [rid for _, rid, _ in mine] and comp == []       # FIFO == row order
            else:
                assert pend == [rid for _, rid, kind in mine if kind == K.AGR_OUT_ERROR]          # failed once: back in place (Q11)
                assert sorted(comp) == sorted(rid for _, rid, kind in mine if kind == K.AGR_OUT_RESPONSE)
                for _, rid, kind in mine[::13]:
                    rec = eng.get_record(a, rid)
                    assert (int(rec["status"]), int(rec["retry_count"])) == ((K.AGR_ST_PENDING, 1) if kind == K.AGR_OUT_ERROR else (K.AGR_ST_COMPLETED, 0))
        s = eng.stats()
        assert s["stored"] == nt * per and s["svc_ops"] >= nt * per and 0 < s["svc_batches"] <= s["svc_ops"]
        # the engine's other entry points keep working next to the resident kernel (they stop it, the dispatcher restarts it)
        big = make_records([Req(agents[0], rid_of(10 ** 6 + i), 10 ** 6 + i) for i in range(100)])
        v, _ = eng.ingest(big)
        assert (v["code"] == K.AGR_V_FORWARD).all()
        out, ids = np.zeros(1, dtype=A.verdict_dtype), np.zeros((1, 16), dtype=np.uint8)
        eng.ingest_ex(make_records([Req(agents[1], rid_of(2 * 10 ** 6), 2 * 10 ** 6)]), out, ids)
        assert int(out[0]["code"]) == K.AGR_V_QUEUED


@pytest.mark.parametrize("flags", [MINT, HASH, MINT | K.AGR_CFG_VARLEN])
def test_random_streams_with_the_manual_replay_handler(flags):
    """Random event streams that also call POST /agents/{id}/requests/{reqId}/replay (server.go:681-751: the agent is called
    directly, every client error reaches MarkRequestFailed, misses answer 404 / 503 without touching anything)."""
    from scenario import run_oracle, run_engine, assert_same, random_scenario
    for seed in (21, 22, 23):
        ev = random_scenario(seed, n_events=400, n_agents=4, p_manual=0.08)
        assert sum(1 for e in ev if e[0] == "manual") > 10
        with A.Engine(slab_rows=1 << 12, max_agents=16, flags=flags) as eng:
            o, g = run_oracle(ev), run_engine(eng, ev)
            assert_same(o, g)
            assert {200, 502} <= set(o.manual) and (404 in o.manual or 503 in o.manual)


@pytest.mark.parametrize("flags", [MINT, MINT | K.AGR_CFG_VARLEN])
def test_reclaim_async_keeps_a_ring_running_one_step_behind(flags):
    """agr_expire(now, ttl, NULL) + agr_reclaim_async every step: the release lags one step (it applies the previous call's scan) and
    never waits; the ring still runs far past its capacity, live ids resolve, released ones do not."""
    from jsoncase import make_requests, records_array, var_batch
    R, per, ttl = 4096, 256, 3 * 256
    var = bool(flags & K.AGR_CFG_VARLEN)
    agent = "agent-1700000000000000001"
    with A.Engine(slab_rows=R, max_agents=4, max_batch=per, vslab_bytes=4 << 20, flags=flags | K.AGR_CFG_RING) as eng:
        eng.set_agent_state(agent, "stopped")
        ids_by_step = []
        released = 0
        for step in range(80):                                   # 80 x 256 = 5 laps of the ring
            reqs = make_requests(100 + step, per, [agent], max_payload=1500 if var else 416, big_bodies=var)
            for i, r in enumerate(reqs):
                r.now = step * per + i
            if var:
                blob, offs = var_batch(reqs)
                _, ids, first = eng.ingest_var(blob, offs)
            else:
                out = np.zeros(per, dtype=A.verdict_dtype); ids = np.zeros((per, 16), dtype=np.uint8)
                first = eng.ingest_ex(records_array(reqs), out, ids)
            assert first >= step * per
            ids_by_step.append(ids.copy())
            eng.expire((step + 1) * per, ttl, want_count=False)
            released += eng.reclaim_async()
            st = eng.stats()
            assert st["rows_used"] - st["rows_tail"] <= R
        released += eng.reclaim()                                # drains the scan left pending and scans once more
        st = eng.stats()
        assert st["rows_tail"] == released and st["rows_used"] - st["rows_tail"] <= 4 * per + 64
        get = (lambda rid: eng.get_record_var(agent, rid)) if var else (lambda rid: eng.get_record(agent, rid))
        assert get(bytes(ids_by_step[-1][7])) is not None        # the last step's records are alive
        assert get(bytes(ids_by_step[10][7])) is None            # long gone: expired, released, row reused
        assert len(eng.list(agent, K.AGR_LIST_PENDING, cap=1 << 14)) <= 4 * per


def test_more_tickets_held_than_ring_slots_never_deadlocks():
    """Callers that together hold MORE uncollected tickets than the request ring has slots (8 threads x 4096 in flight against
    16384 slots): a submit that draws a slot whose answer nobody has collected must come back with AGR_EAGAIN instead of waiting
    for that ticket's holder (who may be waiting for ours).  The run has to end, without errors, and make progress."""
    import json
    import subprocess
    exe = os.path.join(os.path.dirname(A.build_host()), "bench_callers")
    for threads, inflight in ((8, 4096), (12, 256), (8, 512)):
        res = subprocess.run([exe, str(threads), "1.0", "0", "mint", "64", str(inflight)], capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, (threads, inflight, res.stdout[-400:], res.stderr[-400:])
        d = json.loads(res.stdout.strip().splitlines()[-1])
        assert d["errors"] == 0 and d["round_trips"] > 0, d
