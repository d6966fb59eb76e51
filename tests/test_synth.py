"""Synthetic stream generator (BASELINE.json configs): host-side properties on CPU, host == device on the GPU."""
import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K


def test_record_shape_and_determinism():
    a = A.synth_fill_host(0, 2000, seed=7, n_agents=16)
    b = A.synth_fill_host(0, 2000, seed=7, n_agents=16)
    c = A.synth_fill_host(1000, 1000, seed=7, n_agents=16)
    assert a.tobytes() == b.tobytes()
    assert a[1000:].tobytes() == c.tobytes()          # counter based: any window reproduces
    assert (a["seq"] == np.arange(1, 2001)).all()
    assert (a["path_len"].astype(int) + a["hdr_len"] + a["body_len"] == K.AGR_PAYLOAD_BYTES).all()
    r = a[5]
    path = bytes(r["payload"][: r["path_len"]]).decode()
    assert path == "/agent/" + r["agent_id"].decode() + "/chat"
    body = bytes(r["payload"][r["path_len"] + r["hdr_len"]:]).decode()
    assert body.startswith('{"message":"') and body.endswith('"}')
    assert (a["request_id"][:, 6] >> 4 == 4).all() and (a["request_id"][:, 8] >> 6 == 2).all()   # UUIDv4 bits
    assert len({bytes(x) for x in a["request_id"]}) == 2000
    assert (a["flags"] == (K.AGR_M_POST << 8)).all() and (a["max_retries"] == 3).all()
    assert {x.decode() for x in np.unique(a["agent_id"])} <= {A.synth_agent_id(k) for k in range(16)}


def test_duplicates_name_an_earlier_fresh_record_of_the_same_agent():
    a = A.synth_fill_host(0, 20000, seed=3, n_agents=64, dup_permille=100)
    rep = (a["flags"] & 1) == 1
    assert 0.08 < rep.mean() < 0.12
    first = {}
    for i, r in enumerate(a):
        if not rep[i]:
            first[bytes(r["request_id"])] = (i, bytes(r["agent_id"]))
    for i in np.nonzero(rep)[0]:
        j, agent = first[bytes(a[i]["replay_of"])]
        assert j < i and agent == bytes(a[i]["agent_id"])
    assert (a["replay_of"][~rep] == 0).all()


def test_zipf_is_skewed_by_rank():
    a = A.synth_fill_host(0, 200000, seed=5, n_agents=256, zipf_milli=1200)
    ids = [A.synth_agent_id(k).encode() for k in range(256)]
    counts = np.array([(a["agent_id"] == i).sum() for i in ids[:8]])
    frac = counts / len(a)
    assert 0.22 < frac[0] < 0.29                     # top agent ~25 % at s = 1.2 over 256 ids
    assert (np.diff(counts[:6]) < 0).all()
    # zipf weight ratio rank1/rank2 = 2^1.2 = 2.297
    assert 2.1 < counts[0] / counts[1] < 2.5


@pytest.mark.gpu
def test_device_generator_is_byte_identical_to_host():
    n = 4096
    with A.Engine(slab_rows=3 * n, max_agents=512) as eng:
        for kw in (dict(seed=11, n_agents=16), dict(seed=12, n_agents=256, zipf_milli=1200, dup_permille=100)):
            first = eng.reserve_rows(n)
            eng.synth_fill_rows(1000, first, n, **kw)
            import torch
            host = A.synth_fill_host(1000, n, **kw)
            dev = np.zeros(n, dtype=A.record_dtype)
            import ctypes as C
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(dev.ctypes.data), C.c_void_p(eng.slab_ptr(first)), C.c_size_t(n * 512), 2)
            assert dev.tobytes() == host.tobytes()
