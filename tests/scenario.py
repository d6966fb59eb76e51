"""Event-stream scenarios and their two interpreters.

A scenario is a list of events in a total order (the order commands would reach the single Redis server):

  ("agent",  agent_id, status)                 saveAgent / state sync writes Agent.Status
  ("remove", agent_id)                         agent.Manager.Remove
  ("req",    Req, backend)                     one HTTP call to /agent/{id}/...; backend = what the agent side does
                                               if the proxy forwards it: ("response", code) | ("dial",) | ("error",)
  ("manual", agent_id, rid, backend)           POST /agents/{id}/requests/{reqId}/replay (server.go:681-751): the agent is
                                               called directly; backend as for "req" (every error kind counts as a failure)
  ("tick",   backends, flip)                   one ReplayWorker tick; backends maps request-id hex -> backend for
                                               that replay (default ("response", 200), also ("client",)); flip =
                                               None or (k, agent_id, status): status write right after the k-th
                                               replay of the tick (only for the agent being replayed, see below)

run_oracle() executes the literal restatement of the Go code (oracle/model.py).  run_engine() does what the Go
glue does with the C-ABI: batch consecutive requests into agr_ingest, report outcomes through agr_complete, and on
a tick call agr_replay_scan, re-inject each dispatched record replay-flagged, and report the server-side and
worker-side completions (replay_worker.go:120-163) in the order the reference performs them.

Tick linearisation: the reference checks isAgentRunning and snapshots the pending list per agent when its turn
comes; the engine snapshots all agents at the scan.  The two agree whenever no OTHER agent's status changes in the
middle of a tick, which is what these scenarios guarantee (a flip only ever targets the agent being replayed —
KAT-F / Q8).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import model as M  # noqa: E402

ZERO16 = bytes(16)


@dataclass
class Req:
    agent_id: str
    rid: bytes                     # the id StoreRequest would mint (uuid.New) — 16 raw bytes
    seq: int
    replay: bool = False           # X-Agentainer-Replay: true
    replay_of: bytes = ZERO16      # X-Agentainer-Request-ID
    method: str = "POST"
    subpath: str = "/chat"
    body: bytes = b'{"message":"hi"}'
    headers: Dict[str, str] = field(default_factory=lambda: {"Content-Type": "application/json"})

    @property
    def path(self) -> str:
        return f"/agent/{self.agent_id}{self.subpath}"


@dataclass
class Observed:
    verdicts: List[Tuple] = field(default_factory=list)        # per "req": (code, http, stored, tracked)
    ticks: List[List[Tuple[str, str]]] = field(default_factory=list)   # per tick: [(agent_id, id hex)] dispatch order
    manual: List[int] = field(default_factory=list)            # per "manual": HTTP status of the management call
    lists: Dict[str, Dict[str, List[str]]] = field(default_factory=dict)   # agent -> pending/completed/failed id hex
    records: Dict[Tuple[str, str], Tuple] = field(default_factory=dict)   # (agent, id hex) -> (status, retry, resp)

    def per_agent_ticks(self):
        out = []
        for t in self.ticks:
            d: Dict[str, List[str]] = {}
            for a, r in t:
                d.setdefault(a, []).append(r)
            out.append(d)
        return out


def all_agents(events) -> List[str]:
    seen: List[str] = []
    for e in events:
        a = e[1].agent_id if e[0] == "req" else (e[1] if e[0] in ("agent", "remove") else None)
        if a is not None and a not in seen:
            seen.append(a)
    return seen


def all_fresh(events) -> List[Tuple[str, bytes]]:
    return [(e[1].agent_id, e[1].rid) for e in events if e[0] == "req" and not e[1].replay]


# ----------------------------------------------------------------------------------------------- oracle side
def run_oracle(events, persistence: bool = True) -> Observed:
    ref = M.ReferencePath(persistence)
    obs = Observed()
    for e in events:
        if e[0] == "agent":
            ref.set_agent(e[1], e[2])
        elif e[0] == "remove":
            ref.remove_agent(e[1])
        elif e[0] == "req":
            r: Req = e[1]
            headers = dict(r.headers)
            if r.replay:
                headers["X-Agentainer-Replay"] = "true"
                if r.replay_of != ZERO16:
                    headers["X-Agentainer-Request-ID"] = r.replay_of.hex()
            req = M.HttpRequest(r.method, r.path, headers, r.body, new_id=r.rid.hex(), now=r.seq)
            v, _ = ref.request(r.agent_id, req, e[2])
            obs.verdicts.append((v.code, v.http_status, v.stored, v.request_id != ""))
        elif e[0] == "tick":
            backends, flip = e[1], e[2]

            def backend_for(agent_id, req, _b=backends):
                return _b.get(req["id"], ("response", 200))

            def on_replay(agent_id, request_id, k, _f=flip):
                if _f is not None and k == _f[0]:
                    ref.set_agent(_f[1], _f[2])

            obs.ticks.append(ref.tick(backend_for, now=0, on_replay=on_replay))
        elif e[0] == "manual":
            obs.manual.append(ref.manual_replay(e[1], e[2].hex(), e[3], now=0))
        else:
            raise ValueError(e[0])
    for a in all_agents(events):
        obs.lists[a] = ref.lists(a)
    for a, rid in all_fresh(events):
        st = ref.record_state(a, rid.hex())
        if st is not None:
            obs.records[(a, rid.hex())] = st
    return obs


# ----------------------------------------------------------------------------------------------- engine side
def make_records(reqs: List[Req]) -> np.ndarray:
    from agentainer_lab_b200 import record_dtype, constants as K
    recs = np.zeros(len(reqs), dtype=record_dtype)
    for i, r in enumerate(reqs):
        recs[i]["request_id"] = np.frombuffer(r.rid, dtype=np.uint8)
        recs[i]["replay_of"] = np.frombuffer(r.replay_of if r.replay else ZERO16, dtype=np.uint8)
        recs[i]["agent_id"] = r.agent_id.encode()
        recs[i]["seq"] = r.seq
        recs[i]["flags"] = (K.AGR_F_REPLAY if r.replay else 0) | (K.METHOD_CODES[r.method] << K.AGR_F_METHOD_SHIFT)
        path = r.path.encode()
        hdrs = "".join(f"{k}: {v}\n" for k, v in sorted(r.headers.items())).encode()
        blob = path + hdrs + r.body
        assert len(blob) <= 416
        recs[i]["path_len"], recs[i]["hdr_len"], recs[i]["body_len"] = len(path), len(hdrs), len(r.body)
        recs[i]["status"], recs[i]["max_retries"] = K.AGR_ST_PENDING, 3
        recs[i]["payload"][: len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    return recs


def make_var_batch(reqs: List[Req]):
    """Variable-length wire form (AGR_CFG_VARLEN): 96 B header + payload padded to 16 B per record, one blob + offsets."""
    from agentainer_lab_b200 import header_dtype, constants as K
    parts, offsets = [], [0]
    for r in reqs:
        path = r.path.encode()
        hdrs = "".join(f"{k}: {v}\n" for k, v in sorted(r.headers.items())).encode()
        payload = path + hdrs + r.body
        h = np.zeros(1, dtype=header_dtype)
        h["request_id"] = np.frombuffer(r.rid, dtype=np.uint8)
        h["replay_of"] = np.frombuffer(r.replay_of if r.replay else ZERO16, dtype=np.uint8)
        h["agent_id"] = r.agent_id.encode()
        h["seq"] = r.seq
        h["flags"] = (K.AGR_F_REPLAY if r.replay else 0) | (K.METHOD_CODES[r.method] << K.AGR_F_METHOD_SHIFT)
        h["path_len"], h["hdr_len"], h["body_len"] = len(path), len(hdrs), len(r.body)
        h["status"], h["max_retries"] = K.AGR_ST_PENDING, 3
        pad = (-len(payload)) % 16
        parts.append(h.tobytes() + payload + bytes(pad))
        offsets.append(offsets[-1] + len(parts[-1]))
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8).copy() if parts else np.zeros(0, dtype=np.uint8)
    return blob, np.array(offsets, dtype=np.uint32)


def _outcome(outs: list, rid: bytes, agent_id: str, kind: int, http: int = 0, seq: int = 0):
    outs.append((rid, agent_id, kind, http, seq))


def _outcomes_array(outs: list) -> np.ndarray:
    from agentainer_lab_b200 import outcome_dtype
    arr = np.zeros(len(outs), dtype=outcome_dtype)
    for j, (rid, agent_id, kind, http, seq) in enumerate(outs):
        arr[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8)
        arr[j]["agent_id"] = agent_id.encode()
        arr[j]["kind"], arr[j]["http_status"], arr[j]["seq"] = kind, http, seq
    return arr


def run_engine(eng, events, max_batch: int = 1 << 30, rng=None) -> Observed:
    """Drive the C-ABI like the Go glue.  max_batch / rng vary how consecutive requests are batched (results must
    not depend on it)."""
    from agentainer_lab_b200 import constants as K
    obs = Observed()
    slot_name: Dict[int, str] = {}
    agent_status: Dict[str, str] = {}                 # the Go side's own view (agentMgr.GetAgent), for the manual replay handler
    pend_reqs: List[Tuple[Req, Tuple]] = []
    # AGR_CFG_MINT_IDS: the engine mints the ids (like StoreRequest does); the scenario's symbolic ids are mapped to
    # them so that the oracle (which takes its ids from the stream) and the engine can be compared id for id
    mint = bool(getattr(eng, "mint", False))
    varlen = bool(getattr(eng, "varlen", False))
    s2e: Dict[bytes, bytes] = {}
    e2s: Dict[bytes, bytes] = {}

    def tin(b: bytes) -> bytes:
        return s2e.get(b, b)

    def tout(b: bytes) -> bytes:
        return e2s.get(b, b)

    def flush():
        nonlocal pend_reqs
        while pend_reqs:
            take = len(pend_reqs) if rng is None else int(rng.integers(1, len(pend_reqs) + 1))
            take = min(take, max_batch)
            chunk, pend_reqs = pend_reqs[:take], pend_reqs[take:]
            if varlen:
                import dataclasses
                reqs = [dataclasses.replace(r, replay_of=tin(r.replay_of)) if (mint and r.replay) else r for r, _ in chunk]
                blob, offs = make_var_batch(reqs)
                verdicts, ids, first_rid = eng.ingest_var(blob, offs)
            else:
                recs = make_records([r for r, _ in chunk])
                if mint:
                    for i, (r, _) in enumerate(chunk):
                        if r.replay:
                            recs[i]["replay_of"] = np.frombuffer(tin(r.replay_of), dtype=np.uint8)
                verdicts, first_rid = eng.ingest(recs)
                if mint:
                    ids = eng.mint_ids(first_rid, len(chunk))
            if mint:
                for i, ((r, _), v) in enumerate(zip(chunk, verdicts)):
                    if int(v["flags"]) & K.AGR_VF_STORED:
                        s2e[r.rid] = bytes(ids[i]); e2s[bytes(ids[i])] = r.rid
            outs: list = []
            for (r, backend), v in zip(chunk, verdicts):
                code, flags = int(v["code"]), int(v["flags"])
                tracked = bool(flags & K.AGR_VF_TRACKED)
                obs.verdicts.append((code, int(v["http_status"]), bool(flags & K.AGR_VF_STORED), tracked))
                if code != K.AGR_V_FORWARD or not tracked:
                    continue                                 # interceptTransport: requestID == "" records nothing
                rid = tin(r.replay_of if r.replay else r.rid)
                if backend[0] == "response":
                    _outcome(outs, rid, r.agent_id, K.AGR_OUT_RESPONSE, backend[1], r.seq)
                elif backend[0] == "dial":
                    _outcome(outs, rid, r.agent_id, K.AGR_OUT_DIAL_ERR, 0, r.seq)
                else:
                    _outcome(outs, rid, r.agent_id, K.AGR_OUT_ERROR, 0, r.seq)
            if outs:
                eng.complete(_outcomes_array(outs))

    for e in events:
        if e[0] == "req":
            pend_reqs.append((e[1], e[2]))
            continue
        flush()
        if e[0] == "agent":
            slot_name[eng.set_agent_state(e[1], e[2])] = e[1]
            agent_status[e[1]] = e[2]
        elif e[0] == "remove":
            eng.drop_agent(e[1])
            agent_status.pop(e[1], None)
        elif e[0] == "manual":
            # what the Go handler does with the C-ABI: storage.Get -> agr_get_record, GetAgent stays Go (host-side status),
            # client.Do stays Go, and the outcome goes through agr_complete: StoreResponse, or MarkRequestFailed for ANY error
            agent_id, rid, backend = e[1], tin(e[2]), e[3]
            known = (not mint) or (e[2] in s2e)
            rec = None
            if known:
                rec = (eng.get_record_var(agent_id, rid) if varlen else eng.get_record(agent_id, rid))
            if rec is None:
                obs.manual.append(404)
            elif agent_id not in agent_status:
                obs.manual.append(404)
            elif agent_status[agent_id] != "running":
                obs.manual.append(503)
            else:
                outs = []
                if backend[0] == "response":
                    _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, backend[1]); obs.manual.append(200)
                else:
                    _outcome(outs, rid, agent_id, K.AGR_OUT_ERROR); obs.manual.append(502)
                eng.complete(_outcomes_array(outs))
        elif e[0] == "tick":
            backends, flip = e[1], e[2]
            if varlen:
                from agentainer_lab_b200 import header_dtype
                disp, vblob, voffs = eng.replay_scan_var()
                vblob = vblob.copy()
                for j in range(len(disp)):          # replayRequest: same record, replay-flagged, ID in the tracking header
                    hv = vblob[int(voffs[j]): int(voffs[j]) + 96].view(header_dtype)
                    hv["flags"] |= K.AGR_F_REPLAY
                    hv["replay_of"] = hv["request_id"]
                recs = disp                          # only its length / request ids are used below
            else:
                disp, recs = eng.replay_scan(with_records=True)
                # replayRequest (replay_worker.go:120-163): same record, replay-flagged, ID in the tracking header
                recs = recs.copy()
                recs["flags"] |= K.AGR_F_REPLAY
                recs["replay_of"] = recs["request_id"]
            order = [(slot_name[int(d["agent_slot"])], tout(bytes(d["request_id"])).hex()) for d in disp]
            obs.ticks.append(order)
            cuts = [0, len(recs)]
            if flip is not None and 0 < flip[0] < len(recs):
                cuts = [0, flip[0], len(recs)]
            for a, b in zip(cuts[:-1], cuts[1:]):
                if a > 0 and flip is not None:
                    eng.set_agent_state(flip[1], flip[2]); agent_status[flip[1]] = flip[2]
                seg = np.ascontiguousarray(recs[a:b])
                if len(seg) == 0:
                    continue
                if varlen:
                    lo = int(voffs[a])
                    verdicts, _, _ = eng.ingest_var(np.ascontiguousarray(vblob[lo: int(voffs[b])]), (voffs[a: b + 1] - voffs[a]).astype(np.uint32))
                else:
                    verdicts, _ = eng.ingest(seg)
                outs = []
                for rec, v, (agent_id, rid_hex) in zip(seg, verdicts, order[a:b]):
                    rid = bytes(rec["request_id"])
                    code = int(v["code"])
                    backend = backends.get(rid_hex, ("response", 200))
                    if code == K.AGR_V_FORWARD:
                        if backend[0] == "response":        # server side (server.go:588-594) then worker (:158)
                            _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, backend[1])
                            _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, backend[1])
                        elif backend[0] == "dial":          # server: stays pending; proxy answers 502 -> worker completes
                            _outcome(outs, rid, agent_id, K.AGR_OUT_DIAL_ERR)
                            _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, 502)
                        elif backend[0] == "error":         # server: MarkRequestFailed; 502 -> worker completes (KAT-H)
                            _outcome(outs, rid, agent_id, K.AGR_OUT_ERROR)
                            _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, 502)
                        else:                               # worker's own client failed (replay_worker.go:109-112)
                            _outcome(outs, rid, agent_id, K.AGR_OUT_ERROR)
                    else:                                   # proxy answered itself (202 / 503 / 404): worker stores it (Q8)
                        _outcome(outs, rid, agent_id, K.AGR_OUT_RESPONSE, int(v["http_status"]))
                if outs:
                    eng.complete(_outcomes_array(outs))
            if flip is not None and flip[0] >= len(recs):
                eng.set_agent_state(flip[1], flip[2]); agent_status[flip[1]] = flip[2]
        else:
            raise ValueError(e[0])
    flush()
    names = {0: "pending", 1: "completed", 2: "failed"}
    for a in all_agents(events):
        obs.lists[a] = {names[w]: [tout(bytes(x)).hex() for x in eng.list(a, w)] for w in (0, 1, 2)}
    for a, rid in all_fresh(events):
        if varlen:
            from agentainer_lab_b200 import header_dtype
            raw = eng.get_record_var(a, tin(rid)) if (not mint or rid in s2e) else None
            rec = raw[:96].view(header_dtype)[0] if raw is not None else None
        else:
            rec = eng.get_record(a, tin(rid)) if (not mint or rid in s2e) else None
        if rec is not None:
            obs.records[(a, rid.hex())] = (K.STATUS_NAMES[int(rec["status"])], int(rec["retry_count"]), int(rec["resp_status"]))
    return obs


def assert_same(o: Observed, g: Observed) -> None:
    assert o.verdicts == g.verdicts, _first_diff(o.verdicts, g.verdicts, "verdict")
    assert o.per_agent_ticks() == g.per_agent_ticks(), "replay dispatch order differs"
    assert o.ticks == g.ticks, "cross-agent dispatch order differs from the canonical (registration) order"
    assert o.manual == g.manual, f"manual replay statuses differ: oracle {o.manual} != engine {g.manual}"
    assert o.lists == g.lists, _lists_diff(o.lists, g.lists)
    assert o.records == g.records, _first_diff(sorted(o.records.items()), sorted(g.records.items()), "record")


def _first_diff(a, b, what):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return f"{what} {i}: oracle {x} != engine {y}"
    return f"{what} count: oracle {len(a)} != engine {len(b)}"


def _lists_diff(a, b):
    for ag in a:
        for q in a[ag]:
            if a[ag][q] != b.get(ag, {}).get(q):
                return f"list {ag}/{q}: oracle {a[ag][q][:6]}.. (n={len(a[ag][q])}) != engine {b.get(ag, {}).get(q, [])[:6]}.. (n={len(b.get(ag, {}).get(q, []))})"
    return "lists differ"


# ----------------------------------------------------------------------------------------------- generators
def rid_of(i: int) -> bytes:
    """Deterministic UUIDv4-shaped id for scenario record i."""
    import hashlib
    b = bytearray(hashlib.sha256(b"agr-scenario-%d" % i).digest()[:16])
    b[6] = (b[6] & 0x0F) | 0x40
    b[8] = (b[8] & 0x3F) | 0x80
    return bytes(b)


def random_scenario(seed: int, n_events: int = 300, n_agents: int = 5, p_replay: float = 0.1, p_manual: float = 0.0) -> list:
    """Mixed stream: status flips, fresh requests with all backend kinds, client-sent replay-flagged duplicates,
    ticks with mixed replay outcomes and KAT-F style mid-tick flips; with p_manual > 0 also calls of the manual replay
    handler (server.go:681-751) on known and unknown ids."""
    rng = np.random.default_rng(seed)
    agents = [f"agent-{1700000000000000000 + 1000003 * k}" for k in range(n_agents)]
    events: list = []
    status = {}
    for a in agents:
        status[a] = "running" if rng.random() < 0.6 else "stopped"
        events.append(("agent", a, status[a]))
    fresh: Dict[str, List[bytes]] = {a: [] for a in agents}
    ctr = 0
    for _ in range(n_events):
        u = rng.random()
        if p_manual and rng.random() < p_manual:
            a = agents[int(rng.integers(n_agents))]
            w = rng.random()
            target = fresh[a][int(rng.integers(len(fresh[a])))] if (fresh[a] and w < 0.85) else rid_of(20_000_000 + ctr)
            v = rng.random()
            events.append(("manual", a, target, ("response", int(rng.choice([200, 500]))) if v < 0.6 else (("dial",) if v < 0.8 else ("error",))))
        elif u < 0.07:
            a = agents[int(rng.integers(n_agents))]
            status[a] = str(rng.choice(["running", "stopped", "paused", "failed", "created"], p=[0.5, 0.3, 0.1, 0.05, 0.05]))
            events.append(("agent", a, status[a]))
        elif u < 0.12:
            backends = {}
            for a in agents:
                for rid in fresh[a]:
                    v = rng.random()
                    if v < 0.12:
                        backends[rid.hex()] = ("dial",)
                    elif v < 0.24:
                        backends[rid.hex()] = ("error",)
                    elif v < 0.30:
                        backends[rid.hex()] = ("client",)
                    elif v < 0.40:
                        backends[rid.hex()] = ("response", int(rng.choice([404, 500, 503])))
            events.append(("tick", backends, None))
        else:
            ctr += 1
            a = agents[int(rng.integers(n_agents))] if rng.random() > 0.03 else "agent-unknown"
            v = rng.random()
            backend = ("response", 200) if v < 0.7 else (("dial",) if v < 0.8 else (("error",) if v < 0.92 else ("response", 500)))
            if a != "agent-unknown" and fresh[a] and rng.random() < p_replay:
                w = rng.random()
                if w < 0.8:
                    target = fresh[a][int(rng.integers(len(fresh[a])))]
                elif w < 0.9:
                    target = rid_of(10_000_000 + ctr)          # unknown id
                else:
                    target = ZERO16                             # header absent
                events.append(("req", Req(a, rid_of(ctr), ctr, replay=True, replay_of=target), backend))
            else:
                r = Req(a, rid_of(ctr), ctr, body=b'{"message":"m%d"}' % ctr)
                if a != "agent-unknown":
                    fresh[a].append(r.rid)
                events.append(("req", r, backend))
    events.append(("tick", {}, None))
    return events


def synth_to_req(r) -> Req:
    """A synthetic-stream record (agr_synth) as a scenario request (byte-identical when packed again)."""
    rep = bool(r["flags"] & 1)
    o = int(r["path_len"]) + int(r["hdr_len"])
    return Req(r["agent_id"].decode(), bytes(r["request_id"]), int(r["seq"]), replay=rep, replay_of=bytes(r["replay_of"]),
               body=bytes(r["payload"][o:o + int(r["body_len"])]),
               headers={"Content-Type": "application/json", "User-Agent": "agr-synth/1"})


def config1_events(n=10_000, na=16, seed=1, dup_permille=50):
    """BASELINE config 1: 10 k synthetic 512 B POST /agent/<id>/chat records, 16 agent ids; half the agents are stopped for
    the first half of the stream, then started; one tick (dedupe + replay order)."""
    import agentainer_lab_b200 as A
    recs = A.synth_fill_host(0, n, seed=seed, n_agents=na, dup_permille=dup_permille)
    agents = [A.synth_agent_id(k) for k in range(na)]
    ev = [("agent", a, "stopped" if k % 2 else "running") for k, a in enumerate(agents)]
    for i in range(n):
        if i == n // 2:
            ev += [("agent", a, "running") for a in agents]
        ev.append(("req", synth_to_req(recs[i]), ("response", 200)))
    ev.append(("tick", {}, None))
    return ev, recs
