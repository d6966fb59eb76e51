"""the JSON wire form of stored records (K5) against the Go-JSON restatement, byte for byte.

The model side is oracle/model.py's Manager doing what StoreRequest / StoreResponse / MarkRequestFailed do (with the
json.Unmarshal + Marshal round trips those functions perform); the engine side is agr_ingest + agr_complete +
agr_store_response / agr_store_error_text, read back through agr_get_record_json / agr_pending_json / agr_rows_json."""
import json

import numpy as np
import pytest

import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K
from oracle import gojson as G
from jsoncase import make_requests, make_script, run_model, records_array, var_batch

pytestmark = pytest.mark.gpu
MODES = {"hash": K.AGR_CFG_PERSISTENCE, "mint": K.AGR_CFG_PERSISTENCE | K.AGR_CFG_MINT_IDS,
         "var": K.AGR_CFG_PERSISTENCE | K.AGR_CFG_VARLEN | K.AGR_CFG_MINT_IDS}
AGENTS = ["agent-1700000000000000001", "agent-1700000000000000002", "agent-1700000000000000003"]


def drive_engine(eng, reqs, script, batch=64):
    """Ingest, then apply the outcome script; fills r.rid with the ids the engine knows the records by."""
    for k, a in enumerate(AGENTS):
        eng.set_agent_state(a, K.AGR_AGENT_RUNNING if k != 1 else K.AGR_AGENT_STOPPED)
    first = None
    for lo in range(0, len(reqs), batch):
        chunk = reqs[lo:lo + batch]
        if eng.varlen:
            blob, offs = var_batch(chunk)
            _, ids, rid0 = eng.ingest_var(blob, offs)
        else:
            recs = records_array(chunk)
            out = np.zeros(len(chunk), dtype=A.verdict_dtype)
            ids = np.zeros((len(chunk), 16), dtype=np.uint8)
            rid0 = eng.ingest_ex(recs, out, ids)
        first = rid0 if first is None else first
        for r, i in zip(chunk, ids):
            r.rid = bytes(i)
    pending_ops = []

    def flush():
        if pending_ops:
            outs = np.zeros(len(pending_ops), dtype=A.outcome_dtype)
            for j, (rid, agent, kind, http, seq) in enumerate(pending_ops):
                outs[j]["request_id"] = np.frombuffer(rid, dtype=np.uint8)
                outs[j]["agent_id"] = agent.encode()
                outs[j]["kind"], outs[j]["http_status"], outs[j]["seq"] = kind, http, seq
            res = eng.complete(outs)
            assert (res == 0).all()
            pending_ops.clear()

    for n, op in enumerate(script):
        r = reqs[op[1]]
        if op[0] == "resp":
            pending_ops.append((r.rid, r.agent_id, K.AGR_OUT_RESPONSE, op[2], op[5]))
        else:
            pending_ops.append((r.rid, r.agent_id, K.AGR_OUT_ERROR, 0, 0))
        if n % 7 == 6:                                          # several outcomes (often of one record) per K2 batch
            flush()
    flush()
    # the bytes travel off the hot path; the latest store per record wins, like the reference's SET
    for op in script:
        r = reqs[op[1]]
        if op[0] == "resp":
            hdr = b"".join(k + b": " + v + b"\n" for k, v in sorted(op[3].items()))
            assert eng.store_response(r.agent_id, r.rid, hdr, op[4])
        else:
            assert eng.store_error_text(r.agent_id, r.rid, op[2])
    return first


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_wire_form_is_byte_identical(mode, seed):
    reqs = make_requests(seed, 260, AGENTS, big_bodies=False)
    if mode == "var":
        reqs = make_requests(seed, 260, AGENTS, max_payload=7000, big_bodies=True)
    script = make_script(seed, len(reqs), 330)
    with A.Engine(slab_rows=1 << 12, max_agents=64, flags=MODES[mode], vslab_bytes=8 << 20, resp_bytes=4 << 20) as eng:
        first = drive_engine(eng, reqs, script)
        redis, mgr = run_model(reqs, script)                      # after the drive: mint mode filled in the ids
        for r in reqs:
            want = G.marshal_request(redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}"))
            got = eng.get_record_json(r.agent_id, r.rid)
            assert got == want, (r, got, want)
        assert eng.get_record_json(AGENTS[0], bytes(range(16))) is None
        n_pending = 0
        for a in AGENTS + ["agent-unknown"]:
            want = G.marshal_list(mgr.get_pending_requests(a))
            got, cnt = eng.pending_json(a)
            assert got == want
            assert cnt == len(redis.lrange_all(f"agent:{a}:requests:pending"))
            n_pending += cnt
        assert n_pending > 20
        # all rows at once: array form and back-to-back form with offsets
        every = [redis.get(f"agent:{r.agent_id}:requests:{G.format_uuid(r.rid)}") for r in reqs]
        got, offs = eng.rows_json(first, len(reqs), as_array=True)
        assert got == G.marshal_list(every)                              # the stored values, as an array
        assert len(json.loads(got)) == len(reqs)
        got, offs = eng.rows_json(first, len(reqs), as_array=False)
        for i, rec in enumerate(every):
            assert got[int(offs[i]):int(offs[i + 1])] == G.marshal_request(rec)
        import copy
        got, offs = eng.rows_json(first, len(reqs), as_array=True, roundtrip=True)      # after a json.Unmarshal of every record
        assert got == G.marshal_list([G.unmarshal_strings(copy.deepcopy(r)) for r in every])
        assert eng.rows_json(first, 0, as_array=True)[0] == b"null"
        assert eng.stats()["k5_launches"] > 0


def test_default_error_text_and_unstored_rows():
    """Without agr_store_error_text the wire form says "transport error" (what oracle/model.py's proxy passes to
    MarkRequestFailed); rows that hold no stored record (replay-flagged arrivals) encode as null."""
    reqs = make_requests(5, 4, AGENTS[:1])
    with A.Engine(slab_rows=1 << 10, max_agents=8, flags=MODES["hash"]) as eng:
        eng.set_agent_state(AGENTS[0], K.AGR_AGENT_RUNNING)
        recs = records_array(reqs)
        recs[3]["flags"] |= K.AGR_F_REPLAY                        # not stored (server.go:508)
        recs[3]["replay_of"] = recs[0]["request_id"]
        out = np.zeros(4, dtype=A.verdict_dtype); ids = np.zeros((4, 16), dtype=np.uint8)
        first = eng.ingest_ex(recs, out, ids)
        outs = np.zeros(1, dtype=A.outcome_dtype)
        outs[0]["request_id"] = recs[1]["request_id"]; outs[0]["agent_id"] = AGENTS[0].encode(); outs[0]["kind"] = K.AGR_OUT_ERROR
        eng.complete(outs)
        d = json.loads(eng.get_record_json(AGENTS[0], reqs[1].rid))
        assert d["error"] == "transport error" and d["retry_count"] == 1 and d["status"] == "pending"
        got, offs = eng.rows_json(first, 4, as_array=True)
        assert json.loads(got)[3] is None


def test_large_pending_list():
    """A pending list far larger than one CTA's chunk: offsets across chunk boundaries, the array framing, the count."""
    n = 5000
    reqs = make_requests(9, n, AGENTS[:1])
    with A.Engine(slab_rows=1 << 13, max_agents=8, flags=MODES["mint"]) as eng:
        eng.set_agent_state(AGENTS[0], K.AGR_AGENT_STOPPED)
        out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
        eng.ingest_ex(records_array(reqs), out, ids)
        for r, i in zip(reqs, ids):
            r.rid = bytes(i)
        redis, mgr = run_model(reqs, [])
        got, cnt = eng.pending_json(AGENTS[0])
        assert cnt == n
        assert got == G.marshal_list(mgr.get_pending_requests(AGENTS[0]))


def test_engine_matches_the_committed_stream():
    """The same seeded stream as tests/golden/json_stream.json (caller-supplied ids, so the bytes are reproducible): the
    engine's output hashes to the committed values."""
    import hashlib, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "json_stream.json")) as f:
        want = json.load(f)
    reqs = make_requests(want["seed"], want["n"], AGENTS)
    script = make_script(want["seed"], want["n"], want["n_ops"])
    with A.Engine(slab_rows=1 << 12, max_agents=64, flags=MODES["hash"], resp_bytes=4 << 20) as eng:
        drive_engine(eng, reqs, script)
        records = [eng.get_record_json(r.agent_id, r.rid) for r in reqs]
        assert hashlib.sha256(b"\n".join(records)).hexdigest() == want["records_sha256"]
        assert [records[i].decode("latin-1") for i in (0, 7, 19, 42, 100)] == want["examples_latin1"]
        for a in AGENTS:
            got, cnt = eng.pending_json(a)
            assert hashlib.sha256(got).hexdigest() == want["pending_sha256"][a] and cnt == want["pending_counts"][a]


@pytest.mark.parametrize("code", [7, 99, 100, 999, 1000, 9999, 10000, 65535])
def test_status_code_digits(code):
    """Response.StatusCode is an int in the reference; every digit count the 16-bit field can hold."""
    reqs = make_requests(6, 1, AGENTS[:1])
    with A.Engine(slab_rows=1 << 10, max_agents=8, flags=MODES["mint"]) as eng:
        eng.set_agent_state(AGENTS[0], K.AGR_AGENT_RUNNING)
        out = np.zeros(1, dtype=A.verdict_dtype); ids = np.zeros((1, 16), dtype=np.uint8)
        eng.ingest_ex(records_array(reqs), out, ids)
        outs = np.zeros(1, dtype=A.outcome_dtype)
        outs[0]["request_id"] = ids[0]; outs[0]["agent_id"] = AGENTS[0].encode()
        outs[0]["kind"], outs[0]["http_status"], outs[0]["seq"] = K.AGR_OUT_RESPONSE, code, 1_700_000_000_000_000_000
        assert eng.complete(outs)[0] == 0
        d = json.loads(eng.get_record_json(AGENTS[0], bytes(ids[0])))
        assert d["response"]["status_code"] == code and d["status"] == "completed"
        assert d["processed_at"] == d["response"]["received_at"] == "2023-11-14T22:13:20Z"


def test_wire_form_against_the_c_restatement_at_scale():
    """100 k records of the synthetic stream, half of them answered or failed: every record's wire form from the engine
    (one agr_rows_json call, 391 chunks of offsets) equals the value the C restatement holds in its keyspace."""
    from oracle.cpu_ref import CRef, record_dtype as cref_record
    n, na = 100_000, 64
    recs = A.synth_fill_host(0, n, seed=9, n_agents=na)
    names = [A.synth_agent_id(k) for k in range(na)]
    rng = np.random.default_rng(3)
    with A.Engine(slab_rows=1 << 17, max_agents=128, flags=MODES["hash"]) as eng:
        c = CRef()
        eng.set_agent_states(names, ["running"] * na)
        for nm in names:
            c.set_agent_state(nm, "running")
        out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
        first = eng.ingest_ex(recs, out, ids)
        c.ingest(recs.astype(cref_record))
        pick = rng.permutation(n)[: n // 2]
        outs = np.zeros(len(pick), dtype=A.outcome_dtype)
        outs["request_id"] = recs["request_id"][pick]; outs["agent_id"] = recs["agent_id"][pick]
        is_resp = rng.random(len(pick)) < 0.7
        outs["kind"] = np.where(is_resp, K.AGR_OUT_RESPONSE, K.AGR_OUT_ERROR)
        outs["http_status"] = np.where(is_resp, rng.choice([200, 201, 404, 500], len(pick)), 0)
        outs["seq"] = 1_000_000 + np.arange(len(pick))
        assert (eng.complete(outs) == 0).all()
        c.complete(outs)
        blob, offs = eng.rows_json(first, n)
        bad = 0
        for i in range(n):
            want = c.get_json(recs["agent_id"][i].decode(), recs["request_id"][i].tobytes())
            if blob[int(offs[i]):int(offs[i + 1])] != want:
                bad += 1
                if bad < 3:
                    print(i, blob[int(offs[i]):int(offs[i + 1])], want)
        assert bad == 0
        c.close()


def test_encode_then_decode_gives_the_records_back():
    """Size-independent property, no oracle involved: what K5 writes, the host-side reader (agr_json_decode) turns back into
    the fields that were ingested and the state the outcomes left."""
    n, na = 20_000, 32
    recs = A.synth_fill_host(0, n, seed=21, n_agents=na)
    names = [A.synth_agent_id(k) for k in range(na)]
    rng = np.random.default_rng(8)
    with A.Engine(slab_rows=1 << 15, max_agents=64, flags=MODES["mint"]) as eng:
        eng.set_agent_states(names, ["running"] * na)
        out = np.zeros(n, dtype=A.verdict_dtype); ids = np.zeros((n, 16), dtype=np.uint8)
        first = eng.ingest_ex(recs, out, ids)
        pick = rng.permutation(n)[: n // 2]
        outs = np.zeros(len(pick), dtype=A.outcome_dtype)
        outs["request_id"] = ids[pick]; outs["agent_id"] = recs["agent_id"][pick]
        is_resp = rng.random(len(pick)) < 0.6
        outs["kind"] = np.where(is_resp, K.AGR_OUT_RESPONSE, K.AGR_OUT_ERROR)
        outs["http_status"] = np.where(is_resp, 200, 0)
        outs["seq"] = 1_700_000_000_000_000_000 + np.arange(len(pick)) * 1000
        assert (eng.complete(outs) == 0).all()
        state = {int(i): (bool(r), int(t)) for i, r, t in zip(pick, is_resp, outs["seq"])}
        blob, offs = eng.rows_json(first, n)
        for i in range(n):
            rc, d = A.json_decode(blob[int(offs[i]):int(offs[i + 1])])
            assert rc == 0
            h, info, r = d["header"], d["info"], recs[i]
            pl, hl, bl = int(r["path_len"]), int(r["hdr_len"]), int(r["body_len"])
            pay = r["payload"].tobytes()
            assert bytes(h["request_id"]) == bytes(ids[i]) and h["agent_id"] == r["agent_id"] and int(h["seq"]) == int(r["seq"])
            assert d["path"] == pay[:pl] and d["headers"] == pay[pl:pl + hl] and d["body"] == pay[pl + hl:pl + hl + bl]
            assert (int(h["flags"]) >> 8) & 0xff == (int(r["flags"]) >> 8) & 0xff
            if i not in state:
                assert info.status == K.AGR_ST_PENDING and info.retry_count == 0 and not info.has_response and d["error"] == b""
            elif state[i][0]:
                assert info.status == K.AGR_ST_COMPLETED and info.has_response and info.resp_status == 200
                assert info.processed_at == info.received_at == state[i][1]
            else:
                assert info.status == K.AGR_ST_PENDING and info.retry_count == 1 and d["error"] == b"transport error"


def test_engine_reproduces_the_hand_derived_kats():
    """The hand-written documents of tests/golden/gojson_kats.json, byte for byte, from the engine itself: each record is
    ingested with its id, driven to its state with agr_complete / agr_store_*, and read back.  (A KAT whose received_at
    differs from its processed_at is skipped: the engine keeps one time per StoreResponse)"""
    import os
    from agentainer_lab_b200 import record_dtype
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gojson_kats.json")) as f:
        kats = json.load(f)["records"]
    done = 0
    for k in kats:
        rec = k["record"]
        resp = rec["response"]
        if resp and resp["received_at"] != rec["processed_at"]:
            continue
        rid = bytes.fromhex(rec["id"].replace("-", ""))
        path = rec["path"].encode()
        hdrs = "".join(f"{a}: {b}\n" for a, b in sorted(rec["headers"].items(), key=lambda kv: kv[0].encode())).encode()
        body = bytes.fromhex(rec["body_hex"])
        r = np.zeros(1, dtype=record_dtype)
        r[0]["request_id"] = np.frombuffer(rid, dtype=np.uint8); r[0]["agent_id"] = rec["agent_id"].encode(); r[0]["seq"] = rec["created_at"]
        r[0]["flags"] = K.METHOD_CODES[rec["method"]] << K.AGR_F_METHOD_SHIFT
        r[0]["path_len"], r[0]["hdr_len"], r[0]["body_len"] = len(path), len(hdrs), len(body)
        r[0]["status"], r[0]["max_retries"] = K.AGR_ST_PENDING, rec["max_retries"]
        blob = path + hdrs + body
        r[0]["payload"][: len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        with A.Engine(slab_rows=1 << 10, max_agents=8, flags=MODES["hash"]) as eng:
            eng.set_agent_state(rec["agent_id"], K.AGR_AGENT_STOPPED)
            out = np.zeros(1, dtype=A.verdict_dtype); ids = np.zeros((1, 16), dtype=np.uint8)
            eng.ingest_ex(r, out, ids)
            outs = np.zeros(1, dtype=A.outcome_dtype)
            outs[0]["request_id"] = np.frombuffer(rid, dtype=np.uint8); outs[0]["agent_id"] = rec["agent_id"].encode()
            for _ in range(rec["retry_count"]):
                outs[0]["kind"], outs[0]["http_status"], outs[0]["seq"] = K.AGR_OUT_ERROR, 0, rec["created_at"] + 1
                assert eng.complete(outs)[0] == 0
            if rec["error"]:
                assert eng.store_error_text(rec["agent_id"], rid, rec["error"].encode())
            if resp:
                outs[0]["kind"], outs[0]["http_status"], outs[0]["seq"] = K.AGR_OUT_RESPONSE, resp["status_code"], rec["processed_at"]
                assert eng.complete(outs)[0] == 0
                rh = "".join(f"{a}: {b}\n" for a, b in sorted(resp["headers"].items(), key=lambda kv: kv[0].encode())).encode()
                assert eng.store_response(rec["agent_id"], rid, rh, bytes.fromhex(resp["body_hex"]))
            assert eng.get_record_json(rec["agent_id"], rid).decode() == k["json"], k["name"]
        done += 1
    assert done >= 5
