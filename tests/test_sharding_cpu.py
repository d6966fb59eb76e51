"""N > 1 host-side logic on CPU (world_size 2, gloo): shard ownership, routing of cross-shard records to their owner,
and result invariance — per-agent results of the sharded run equal the single-shard oracle."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import agentainer_lab_b200 as A
    from oracle.cpu_ref import CRef
    from agentainer_lab_b200.sharding import owned_agents, make_rank_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    own = owned_agents(world, 6)
    status = lambda a: "running" if int(a[-1]) % 3 else "stopped"
    batch = make_rank_batch(rank, world, own, 2000, seed=7, p_cross_replay=0.05, p_missteer=0.05)
    owner = np.array([A.agent_shard(a.decode(), world) for a in batch["agent_id"]])
    # all-to-all of the foreign records (what K4 + NCCL do on the GPUs), local first then by source rank
    send = [batch[owner == p].tobytes() for p in range(world)]
    recv = [None] * world
    dist.all_to_all_object_list(recv, send) if hasattr(dist, "all_to_all_object_list") else None
    if recv[0] is None:                      # gloo without all_to_all_object_list: gather everything, pick ours
        allsend = [None] * world
        dist.all_gather_object(allsend, send)
        recv = [allsend[src][rank] for src in range(world)]
    order = [rank] + [p for p in range(world) if p != rank]
    mine = np.concatenate([np.frombuffer(recv[p], dtype=A.record_dtype) for p in order])
    c = CRef()
    for a in own[rank]:
        c.set_agent_state(a, status(a))
    v, _ = c.ingest(np.ascontiguousarray(mine))
    lists = {a: [bytes(x).hex() for x in c.list(a, 0)] for a in own[rank]}
    allb = [None] * world
    dist.all_gather_object(allb, batch.tobytes())
    q.put((rank, lists, allb if rank == 0 else None, [int(x) for x in v["code"]], mine["request_id"].tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_results_equal_single_shard_oracle():
    import agentainer_lab_b200 as A
    from oracle.cpu_ref import CRef
    from agentainer_lab_b200.sharding import owned_agents
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    batches = [np.frombuffer(b, dtype=A.record_dtype) for b in res[0][2]]
    own = owned_agents(world, 6)
    status = lambda a: "running" if int(a[-1]) % 3 else "stopped"
    # single-shard oracle over the union stream, ordered the way the owner merges it: for every owner, its own
    # host's records first, then the other hosts' by rank — per-agent order is all that matters (agents are independent)
    single = CRef()
    for r in range(world):
        for a in own[r]:
            single.set_agent_state(a, status(a))
    for r in range(world):
        order = [r] + [p for p in range(world) if p != r]
        for src in order:
            b = batches[src]
            sel = np.array([A.agent_shard(a.decode(), world) == r for a in b["agent_id"]])
            if sel.any():
                single.ingest(np.ascontiguousarray(b[sel]))
    for r in range(world):
        for a in own[r]:
            assert res[r][1][a] == [bytes(x).hex() for x in single.list(a, 0)], (r, a)
    # every rank saw records of both kinds, and ownership is a partition
    assert all(len(set(A.agent_shard(a, world) for a in own[r])) == 1 for r in range(world))
    assert sum(len(v) for v in res[0][1].values()) > 100
