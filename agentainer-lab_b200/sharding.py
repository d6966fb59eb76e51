"""Harness helpers for the sharded (multi-GPU) tests and bench: who owns an agent, and how a host builds the batch
that arrives at one shard (BASELINE config 4).  Not a product surface."""
import numpy as np

from . import binding as A
from . import constants as K


def owned_agents(world: int, per_rank: int, nanos0: int = 1700000000000000000):
    """First `per_rank` synthetic agent ids owned by each rank under owner = FNV-1a64(id) mod world."""
    own = [[] for _ in range(world)]
    k = 0
    while min(len(o) for o in own) < per_rank:
        a = A.synth_agent_id(k, agent_nanos0=nanos0)
        r = A.agent_shard(a, world)
        if len(own[r]) < per_rank:
            own[r].append(a)
        k += 1
    return own


def make_rank_batch(rank: int, world: int, own, n: int, seed: int, p_cross_replay=0.05, p_missteer=0.0, first_index=0):
    """n records arriving at `rank`'s host: fresh records for its own agents, plus replay-flagged records (and optionally
    mis-steered fresh ones) whose agent is owned by ANOTHER shard (BASELINE config 4)."""
    rng = np.random.default_rng(seed * 1000 + rank)
    recs = A.synth_fill_host(first_index, n, seed=seed * 100 + rank, n_agents=len(own[rank]))
    ids = np.array([a.encode() for a in own[rank]], dtype="S32")
    base = A.synth_agent_id(0).encode()[:6]
    assert base == b"agent-"
    # synthetic agent index -> this rank's own agent ids
    idx = np.array([int(x[6:]) for x in recs["agent_id"]]) - 1700000000000000000
    idx //= 1000003
    recs["agent_id"] = ids[idx]
    if world > 1:
        u = rng.random(n)
        cross = u < p_cross_replay
        mis = (u >= p_cross_replay) & (u < p_cross_replay + p_missteer)
        other = (rank + 1 + rng.integers(0, world - 1, n)) % world
        pick = rng.integers(0, len(own[0]), n)
        for i in np.nonzero(cross | mis)[0]:
            recs["agent_id"][i] = own[other[i]][pick[i]].encode()
        recs["flags"][cross] |= K.AGR_F_REPLAY
        recs["replay_of"][cross] = recs["request_id"][cross]          # any non-zero id: tracked
    return recs
