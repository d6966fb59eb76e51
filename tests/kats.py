"""Known-answer scenarios KAT-A..H (+ edge cases) of SURVEY.md section 3.4.  The expected observables live in
tests/golden/kats.json and were derived BY HAND from the reference source (the reference ships no tests), using
symbolic ids r1, r2, ... = scenario.rid_of(1), rid_of(2), ...; agent "A" = AGENT_A."""
import json
import os

from scenario import Req, rid_of, ZERO16

AGENT_A = "agent-1700000000000000001"
AGENT_B = "agent-1700000000000000002"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kats.json")
# written by oracle/go/requests_kat_test.go from a run of the UNMODIFIED reference; preferred when present (parity "pinned")
FROM_REFERENCE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "from_reference", "kats.json")
PINNED = os.path.exists(FROM_REFERENCE)
OK = ("response", 200)


def R(i, agent=AGENT_A, **kw):
    return Req(agent, rid_of(i), i, **kw)


SCENARIOS = {
    # agent running, 200 OK
    "KAT-A": [("agent", AGENT_A, "running"), ("req", R(1), OK)],
    # agent stopped: three 202s, FIFO pending
    "KAT-B": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("req", R(2), OK), ("req", R(3), OK)],
    # start A, one tick: FIFO replay, every id completed twice (Q7)
    "KAT-C": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("req", R(2), OK), ("req", R(3), OK),
              ("agent", AGENT_A, "running"), ("tick", {}, None)],
    # status running but container dead: dial error leaves the record pending, retry untouched (Q12)
    "KAT-D": [("agent", AGENT_A, "running"), ("req", R(1), ("dial",))],
    # non-dial transport error: retry 1, still pending in place (Q11)
    "KAT-E1": [("agent", AGENT_A, "running"), ("req", R(1), ("error",)), ("req", R(2), ("dial",))],
    # three MarkRequestFailed in total -> dead letter (Q13); a further tick no longer sees it
    "KAT-E3": [("agent", AGENT_A, "running"), ("req", R(1), ("error",)),
               ("tick", {rid_of(1).hex(): ("client",)}, None), ("tick", {rid_of(1).hex(): ("client",)}, None),
               ("tick", {}, None)],
    # replay while the agent flips to stopped: 202 stored as the response, record completed, request lost (Q8)
    "KAT-F": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("req", R(2), OK), ("agent", AGENT_A, "running"),
              ("tick", {}, (1, AGENT_A, "stopped"))],
    # client-sent duplicate = replay-flagged request naming a completed id: not stored, completed again
    "KAT-G": [("agent", AGENT_A, "running"), ("req", R(1), OK),
              ("req", R(2, replay=True, replay_of=rid_of(1)), OK)],
    # replay through the proxy hits a non-dial error: retry 1 (server side) then completed with 502 (worker side)
    "KAT-H": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("agent", AGENT_A, "running"),
              ("tick", {rid_of(1).hex(): ("error",)}, None)],
    # replay hits a dial error: server leaves it pending, worker stores the 502 -> completed once
    "KAT-H2": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("agent", AGENT_A, "running"),
               ("tick", {rid_of(1).hex(): ("dial",)}, None)],
    # replay-flag edge cases: unknown id (forwarded, completion misses); empty id (untracked; 503 when down)
    "EDGE-REPLAY": [("agent", AGENT_A, "running"), ("agent", AGENT_B, "stopped"),
                    ("req", R(1, replay=True, replay_of=rid_of(99)), OK),
                    ("req", R(2, replay=True, replay_of=ZERO16), OK),
                    ("req", R(3, agent=AGENT_B, replay=True, replay_of=ZERO16), OK),
                    ("req", R(4, agent=AGENT_B, replay=True, replay_of=rid_of(98)), OK),
                    ("req", R(5, agent="agent-404"), OK)],
    # an id stored under A, replay-flagged at B: the key is agent:{B}:requests:{r} -> miss, A's record untouched
    "EDGE-CROSS-AGENT": [("agent", AGENT_A, "stopped"), ("agent", AGENT_B, "running"), ("req", R(1), OK),
                         ("req", R(2, agent=AGENT_B, replay=True, replay_of=rid_of(1)), OK)],
    # agent.Remove: lists deleted, record orphaned (Q17), later traffic 404
    "KAT-REMOVE": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("remove", AGENT_A), ("req", R(2), OK)],
    # POST /agents/{id}/requests/{reqId}/replay (server.go:681-751): direct call to the agent, ANY client error counts (Q12 note)
    "KAT-MANUAL": [("agent", AGENT_A, "stopped"), ("req", R(1), OK), ("req", R(2), OK),
                   ("manual", AGENT_A, rid_of(1), OK),
                   ("agent", AGENT_A, "running"),
                   ("manual", AGENT_A, rid_of(1), OK), ("manual", AGENT_A, rid_of(2), ("dial",)), ("manual", AGENT_A, rid_of(99), OK),
                   ("tick", {}, None)],
    # FIFO across a failed-in-place entry: r1 errors (retry 1, keeps its position), r2 dial -> pending [r1, r2];
    # stop/start and tick: replay order r1 then r2
    "KAT-ORDER": [("agent", AGENT_A, "running"), ("req", R(1), ("error",)), ("req", R(2), ("dial",)),
                  ("req", R(3), OK), ("tick", {}, None)],
}


def load_golden(path=None):
    """Expected observables per scenario: the reference's own output when tests/golden/from_reference/kats.json exists
    (PINNED), else the hand-derived tests/golden/kats.json."""
    with open(GOLDEN) as f:
        g = json.load(f)
    if path is None and PINNED:
        with open(FROM_REFERENCE) as f:
            ref = json.load(f)
        g = dict(g, kats=ref["kats"], _provenance=ref.get("_provenance", "reference run"))
    elif path is not None:
        with open(path) as f:
            g = dict(g, kats=json.load(f)["kats"])
    sym = {f"r{i}": rid_of(i).hex() for i in range(1, 100)}
    names = {"A": AGENT_A, "B": AGENT_B}

    def ag(a):
        return names.get(a, a)

    out = {}
    for kat, exp in g["kats"].items():
        out[kat] = {
            "verdicts": [tuple(v) for v in exp["verdicts"]],
            "ticks": [[(ag(a), sym[r]) for a, r in t] for t in exp["ticks"]],
            "lists": {ag(a): {q: [sym[r] for r in ids] for q, ids in qs.items()} for a, qs in exp["lists"].items()},
            "records": {(ag(k.split("/")[0]), sym[k.split("/")[1]]): tuple(v) for k, v in exp["records"].items()},
            "manual": list(exp.get("manual", [])),
        }
    return out, g
