"""K4 + NCCL exchange on two GPUs: agr_ingest_sharded routes records that reached a non-owner shard to their owner,
runs K1 there and brings the verdicts back in the caller's order.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import agentainer_lab_b200 as A
    from agentainer_lab_b200 import constants as K
    from sharded_check import check_sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = check_sharded(A, K, dist, rank, world, device=rank)
    q.put((rank, r["ok"], r["k4_launches"], r["stored"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_ingest_two_gpus():
    import torch.multiprocessing as mp
    world, port = 2, 29544
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, k4, stored in res:
        assert ok, f"rank {rank}: sharded ingest differs from the oracle"
        assert k4 > 0 and stored > 0
