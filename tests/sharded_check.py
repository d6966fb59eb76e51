"""Result invariance of the sharded path, as a function: every rank runs its engine through agr_ingest_sharded /
agr_complete_sharded and compares, against oracle/cpu_ref.c fed the owner's merge order (own host's records first, then
the peers' by rank), (1) the verdict of every record of its batch wherever it was decided and (2) the owner's per-agent
pending / completed / failed lists.  Used by tests/test_sharded_gpu.py and, on the driver's multi-GPU box, by bench.py (the
pytest is skipped there: the driver's test box has one GPU).  `dist` is an initialised torch.distributed (any backend)."""
import numpy as np


def check_sharded(A, K, dist, rank: int, world: int, device: int, steps=(3000, 1, 0, 5000), per_rank_agents: int = 8) -> dict:
    from oracle.cpu_ref import CRef
    from agentainer_lab_b200.sharding import owned_agents, make_rank_batch
    uid = [A.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    own = owned_agents(world, per_rank_agents)
    status = lambda a: "running" if int(a[-1]) % 3 else "stopped"
    eng = A.Engine(device=device, slab_rows=1 << 16, max_agents=64, max_batch=1 << 14)
    eng.comm_init(uid[0], rank, world)
    ref = CRef()
    for a in own[rank]:
        eng.set_agent_state(a, status(a)); ref.set_agent_state(a, status(a))
    ok, checked = True, 0
    for step, n in enumerate(steps):
        n_here = n if (n != 1 or rank == 0) else 0          # ragged: one rank sends a single record, the others nothing
        batch = np.zeros(0, dtype=A.record_dtype)
        if n_here:
            batch = make_rank_batch(rank, world, own, n_here, seed=11 + step, p_cross_replay=0.05, p_missteer=0.05, first_index=step * 10000)
        v, info = eng.ingest_sharded(batch)
        allb = [None] * world
        dist.all_gather_object(allb, batch.tobytes())
        batches = [np.frombuffer(b, dtype=A.record_dtype) for b in allb]
        # expected at this owner: own host's records first, then the other ranks' in rank order
        order = [rank] + [p for p in range(world) if p != rank]
        mine = [b[np.array([A.agent_shard(a.decode(), world) == rank for a in b["agent_id"]], dtype=bool)] if len(b) else b for b in (batches[p] for p in order)]
        mine = np.ascontiguousarray(np.concatenate(mine)) if sum(len(m) for m in mine) else np.zeros(0, dtype=A.record_dtype)
        ev, _ = ref.ingest(mine) if len(mine) else (np.zeros(0, dtype=A.verdict_dtype), 0)
        ok &= info.n_local + info.n_sent == n_here and info.n_received == len(mine) - info.n_local
        # verdicts of MY batch, wherever each record was decided: gather every owner's expected verdict by request id
        exp = {}
        alle = [None] * world
        dist.all_gather_object(alle, (mine["request_id"].tobytes(), ev["code"].tobytes(), (ev["flags"] & 0x7).tobytes()))
        for ids, codes, flags in alle:
            ids = np.frombuffer(ids, dtype=np.uint8).reshape(-1, 16)
            for i, c, f in zip(ids, np.frombuffer(codes, dtype=np.uint8), np.frombuffer(flags, dtype=np.uint8)):
                exp[bytes(i)] = (int(c), int(f))
        for rec, got in zip(batch, v if v is not None else []):
            ok &= (int(got["code"]), int(got["flags"]) & 0x7) == exp[bytes(rec["request_id"])]
            checked += 1
    # ---- outcomes reported at the WRONG shard travel to the owner (agr_complete_sharded): every rank completes a slice of
    # EVERY agent's pending records, own or not; the owner applies them (own host first, then peers by rank)
    pend_all = [None] * world
    dist.all_gather_object(pend_all, {a: [bytes(x) for x in eng.list(a, 0)] for a in own[rank]})
    outs = []
    for r in range(world):
        for a, ids in sorted(pend_all[r].items()):
            for j, rid in enumerate(ids[:40]):
                if j % world == rank:
                    outs.append((rid, a, K.AGR_OUT_RESPONSE if j % 3 else K.AGR_OUT_ERROR))
    outs.append((bytes(range(16)), own[(rank + 1) % world][0], K.AGR_OUT_RESPONSE))          # unknown id at a foreign owner
    arr = np.zeros(len(outs), dtype=A.outcome_dtype)
    for j, (rid, a, kind) in enumerate(outs):
        arr[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8); arr[j]["agent_id"] = a.encode(); arr[j]["kind"] = kind; arr[j]["http_status"] = 200
    res, cinfo = eng.complete_sharded(arr)
    ok &= list(res[:-1]) == [0] * (len(outs) - 1) and res[-1] == K.AGR_ENOTFOUND
    ok &= cinfo.n_sent > 0 and cinfo.n_received > 0
    allo = [None] * world
    dist.all_gather_object(allo, arr.tobytes())
    order = [rank] + [p for p in range(world) if p != rank]
    for src in order:                                  # the owner's merge order
        o = np.frombuffer(allo[src], dtype=A.outcome_dtype)
        mine_o = o[np.array([A.agent_shard(a.decode(), world) == rank for a in o["agent_id"]], dtype=bool)]
        if len(mine_o):
            ref.complete(np.ascontiguousarray(mine_o))
    lists = 0
    for a in own[rank]:
        for w in (0, 1, 2):
            ok &= [bytes(x).hex() for x in eng.list(a, w)] == [bytes(x).hex() for x in ref.list(a, w)]
            lists += 1
    s = eng.stats()
    eng.close()
    ref.close()
    return {"ok": bool(ok), "verdicts_checked": checked, "agent_lists_checked": lists, "k4_launches": s["k4_launches"], "stored": s["stored"]}
