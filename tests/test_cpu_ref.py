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
