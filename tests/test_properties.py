"""Property tests: for ANY event stream the implementation equals the oracle on per-request
verdicts, per-record (status, retry, response), per-agent pending / completed / failed id sequences and per-tick replay
dispatch order.  CPU: the C restatement against the Python oracle.  GPU: the CUDA path against the Python oracle."""
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from scenario import Req, rid_of, run_oracle, run_engine, assert_same, ZERO16

AGENTS = ["agent-17000000000000000%02d" % k for k in range(4)] + ["agent-ghost"]
STATUSES = ["running", "stopped", "paused", "failed", "created"]
BACKENDS = [("response", 200), ("response", 500), ("dial",), ("error",)]


@st.composite
def streams(draw):
    n = draw(st.integers(5, 60))
    events = [("agent", a, draw(st.sampled_from(["running", "stopped"]))) for a in AGENTS[:4]]
    fresh = {a: [] for a in AGENTS}
    removed = set()
    ctr = 0
    for _ in range(n):
        kind = draw(st.sampled_from(["req"] * 6 + ["agent", "tick", "replayreq", "remove"]))
        if kind == "agent":
            a = draw(st.sampled_from(AGENTS[:4]))
            if a not in removed:
                events.append(("agent", a, draw(st.sampled_from(STATUSES))))
        elif kind == "remove":
            a = draw(st.sampled_from(AGENTS[:4]))
            if a not in removed and len(removed) < 2:
                removed.add(a)
                events.append(("remove", a))
        elif kind == "tick":
            backends = {}
            for a in AGENTS[:4]:
                for rid in fresh[a]:
                    b = draw(st.sampled_from([None, None, ("dial",), ("error",), ("client",), ("response", 503)]))
                    if b is not None:
                        backends[rid.hex()] = b
            events.append(("tick", backends, None))
        else:
            ctr += 1
            a = draw(st.sampled_from(AGENTS))
            backend = draw(st.sampled_from(BACKENDS))
            if kind == "replayreq":
                pool = [r for x in AGENTS for r in fresh[x]]
                target = draw(st.sampled_from(pool + [rid_of(900000 + ctr), ZERO16])) if pool else ZERO16
                events.append(("req", Req(a, rid_of(ctr), ctr, replay=True, replay_of=target), backend))
            else:
                body = draw(st.binary(min_size=0, max_size=120))
                r = Req(a, rid_of(ctr), ctr, body=body, method=draw(st.sampled_from(["GET", "POST", "PUT", "DELETE"])),
                        subpath=draw(st.sampled_from(["/chat", "", "/", "/history"])))
                if a != "agent-ghost" and a not in removed:
                    fresh[a].append(r.rid)
                events.append(("req", r, backend))
    events.append(("tick", {}, None))
    return events


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(streams())
def test_c_port_equals_python_oracle_on_any_stream(events):
    from oracle.cpu_ref import CRef
    with CRef() as c:
        assert_same(run_oracle(events), run_engine(c, events))


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 9, 25])      # hash ids / engine-minted ids / minted + variable-length
def test_cuda_path_equals_python_oracle_on_any_stream(flags):
    import agentainer_lab_b200 as A

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
    @given(streams())
    def run(events):
        with A.Engine(slab_rows=1 << 12, max_agents=32, flags=flags, vslab_bytes=8 << 20) as eng:
            assert_same(run_oracle(events), run_engine(eng, events))
    run()
