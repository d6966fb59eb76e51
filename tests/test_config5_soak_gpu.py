"""BASELINE config 5 in the small: a sustained stream of variable-length records (128 B - 4 KB bodies) through the ring —
rows and bytes lap several times — with injected agent crashes: for one batch 1 % of the agents answer nothing (dial
errors, the records stay pending, Q12), then a replay tick dispatches exactly those records, grouped by agent, in arrival
order (Q14), and the double completion of a replay follows (Q7).  No oracle at this size: the checks are the
size-independent properties of the path."""
import os
import sys

import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def outcomes(ids, agents, kinds, seq):
    outs = np.zeros(len(ids), dtype=A.outcome_dtype)
    outs["request_id"] = ids
    outs["agent_id"] = agents
    outs["kind"] = kinds
    outs["http_status"] = np.where(kinds == K.AGR_OUT_RESPONSE, 200, 0)
    outs["seq"] = seq
    return outs


def test_config5_sustained_with_crash_replay():
    import bench
    n, na, steps = 1 << 15, 256, 40
    R, VB = 1 << 18, 512 << 20
    rng = np.random.default_rng(5)
    flags = K.AGR_CFG_PERSISTENCE | K.AGR_CFG_VARLEN | K.AGR_CFG_MINT_IDS | K.AGR_CFG_RING
    names = [A.synth_agent_id(k) for k in range(na)]
    with A.Engine(slab_rows=R, max_agents=512, max_batch=n, vslab_bytes=VB, flags=flags) as eng:
        slots = eng.set_agent_states(names, ["running"] * na)
        slot_of = {nm.encode(): int(s) for nm, s in zip(names, slots)}
        total_bytes = crashed_records = dispatched_total = 0
        pending_dial = None                                    # (ids, agents) of the records that got no answer last step
        for s in range(steps):
            blob, offs, nbytes = bench.make_var_blob(A, n, s * n, 5, na, 0, rng)
            agents = A.synth_fill_host(s * n, n, seed=5, n_agents=na, agent_nanos0=0)["agent_id"]
            total_bytes += nbytes
            v, ids, first = eng.ingest_var(blob, offs)
            assert (v["code"] == K.AGR_V_FORWARD).all() and first == s * n   # R is a multiple of n: no rows skipped
            now = (s + 1) * n
            kinds = np.full(n, K.AGR_OUT_RESPONSE, dtype=np.uint8)
            if s % 10 == 5:                                    # 1 % of the agents crash for this batch
                down = {names[k].encode() for k in range(na) if k % 100 == (s // 10) % 100}
                hit = np.array([a in down for a in agents])
                kinds[hit] = K.AGR_OUT_DIAL_ERR
                pending_dial = (ids[hit].copy(), agents[hit].copy())
                crashed_records += int(hit.sum())
            assert (eng.complete(outcomes(ids, agents, kinds, now)) == 0).all()
            if s % 10 == 6:                                    # the agents are back: one replay tick
                disp = eng.replay_scan_var()[0]
                want_ids, want_agents = pending_dial
                order = np.argsort(np.array([slot_of[a] for a in want_agents]), kind="stable")   # grouped by agent, FIFO inside
                assert len(disp) == len(want_ids) > 0
                assert (disp["request_id"] == want_ids[order]).all()
                rows = disp["rid"].astype(np.int64)
                for sl in np.unique(disp["agent_slot"]):
                    r = rows[disp["agent_slot"] == sl]
                    assert (np.diff(r) > 0).all()                # arrival order inside an agent
                twice = np.repeat(np.arange(len(disp)), 2)      # proxy-side completion, then the worker's (Q7)
                assert (eng.complete(outcomes(disp["request_id"][twice], want_agents[order][twice],
                                              np.full(len(twice), K.AGR_OUT_RESPONSE, dtype=np.uint8), now)) == 0).all()
                dispatched_total += len(disp)
                pending_dial = None
                lst = eng.list(want_agents[order][0].decode(), K.AGR_LIST_COMPLETED, cap=1 << 16)
                assert (lst[-2] == lst[-1]).all()                # the replayed id closes the list twice
            eng.expire(now, 4 * n)
            eng.reclaim()
        st = eng.stats()
        assert st["stored"] == steps * n and st["forwarded"] == steps * n
        assert st["dial_errors"] == crashed_records == dispatched_total and crashed_records > 500
        assert st["completions"] == steps * n - crashed_records + 2 * dispatched_total and st["completion_misses"] == 0
        assert st["rows_used"] == steps * n and st["rows_used"] - st["rows_tail"] <= 5 * n
        assert total_bytes > 3 * VB                                # the byte ring lapped three times, the row ring five
        for k in (0, 1, 57):
            assert eng.pending_json(names[k])[1] == 0
        assert eng.verify()[1] == 0
