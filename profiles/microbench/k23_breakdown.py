"""Per-kernel times of K2 / K3 / K5 at size (not a bench: run under `ncu --metrics gpu__time_duration.sum`).
  python profiles/microbench/k23_breakdown.py [steps of 1 M records, default 64]
Fills `steps` M rows (C3 synthetic workload, mint ids), completes the last batch (K2), runs one replay tick over everything
with 1/16 of the agents running (K3) and encodes the last batch as JSON (K5); prints the engine's own CUDA-event times."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import agentainer_lab_b200 as A
from agentainer_lab_b200 import constants as K

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = 1 << 20
eng = A.Engine(device=0, slab_rows=S * B, max_agents=1024, max_batch=B, flags=K.AGR_CFG_PERSISTENCE | K.AGR_CFG_TIMING | K.AGR_CFG_MINT_IDS)
nanos0 = 1700000000000000000
for k in range(256):
    eng.set_agent_state(A.synth_agent_id(k, agent_nanos0=nanos0), "running")
synth = dict(seed=2, n_agents=256, zipf_milli=1200, dup_permille=100, agent_nanos0=nanos0)
first = eng.reserve_rows(B)
for s in range(1, S):
    eng.reserve_rows(B)
for s in range(S):
    eng.synth_fill_rows(s * B, first + s * B, B, mint_base=first, **synth)
for s in range(S):
    eng.ingest_rows_async(first + s * B, B)
eng.sync()
host = A.synth_fill_host((S - 1) * B, B, mint=(eng, first), **synth)
outs = eng.pinned(B, A.outcome_dtype)
outs.array["request_id"] = eng.mint_ids(first + (S - 1) * B, B)
outs.array["agent_id"] = host["agent_id"]
outs.array["kind"] = K.AGR_OUT_RESPONSE
outs.array["http_status"] = 200
outs.array["seq"] = (S + 8) * B
rep = (host["flags"] & 1) != 0
outs.array["request_id"][rep] = host["replay_of"][rep]
eng.complete(outs.array, want_results=False)
k2 = eng.op_time(0)
for k in range(256):
    if k % 16:
        eng.set_agent_state(A.synth_agent_id(k, agent_nanos0=nanos0), "stopped")
for rep_ in range(2):
    disp, _ = eng.replay_scan(with_records=False, cap=1 << 22)
    k3 = eng.op_time(1)
eng.rows_json(first + (S - 1) * B, B, as_array=True, fetch=False)
jb = eng.rows_json(first + (S - 1) * B, B, as_array=True, fetch=False)
k5 = eng.op_time(2)
print({"rows": S * B, "k2_ms": k2, "k3_ms": k3, "k3_rows_per_s": S * B / (k3 * 1e-3), "dispatched": len(disp), "k5_ms": k5, "json_bytes": int(jb)})
eng.close()
