"""ctypes wrapper of oracle/libcpu_ref.so with the same method names as the product's Engine harness, so one
scenario driver (tests/scenario.run_engine) exercises both.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

record_dtype = np.dtype([
    ("request_id", "u1", 16), ("replay_of", "u1", 16), ("agent_id", "S32"), ("seq", "<u8"), ("flags", "<u4"),
    ("path_len", "<u2"), ("hdr_len", "<u2"), ("body_len", "<u4"), ("status", "u1"), ("retry_count", "u1"),
    ("max_retries", "u1"), ("error_code", "u1"), ("resp_status", "<u2"), ("reserved0", "<u2"), ("reserved1", "<u4"),
    ("payload", "u1", 416)])
verdict_dtype = np.dtype([("code", "u1"), ("flags", "u1"), ("http_status", "<u2"), ("agent_slot", "<u4")])
dispatch_dtype = np.dtype([("rid", "<u8"), ("agent_slot", "<u4"), ("reserved", "<u4"), ("request_id", "u1", 16)])
AGR_ECAP, AGR_ENOTFOUND = -7, -5
AGENT_STATUS_CODES = {"created": 0, "running": 1, "stopped": 2, "paused": 3, "failed": 4}


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcpu_ref.so")
        src = os.path.join(HERE, "cpu_ref.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = C.CDLL(so)
        vp, u32 = C.c_void_p, C.c_uint32
        lib.cref_create.restype, lib.cref_create.argtypes = vp, [u32]
        lib.cref_destroy.restype, lib.cref_destroy.argtypes = None, [vp]
        lib.cref_set_agent_state.restype, lib.cref_set_agent_state.argtypes = C.c_int, [vp, C.c_char_p, C.c_uint8]
        lib.cref_drop_agent.restype, lib.cref_drop_agent.argtypes = C.c_int, [vp, C.c_char_p]
        lib.cref_ingest.restype, lib.cref_ingest.argtypes = C.c_int, [vp, vp, u32, vp]
        lib.cref_complete.restype, lib.cref_complete.argtypes = C.c_int, [vp, vp, u32, vp]
        lib.cref_pending.restype, lib.cref_pending.argtypes = C.c_int, [vp, C.c_char_p, vp, u32, C.POINTER(u32)]
        lib.cref_scan.restype, lib.cref_scan.argtypes = C.c_int, [vp, vp, vp, u32, C.POINTER(u32)]
        lib.cref_get_record.restype, lib.cref_get_record.argtypes = C.c_int, [vp, C.c_char_p, vp, vp]
        lib.cref_list.restype, lib.cref_list.argtypes = C.c_int, [vp, C.c_char_p, C.c_int, vp, u32, C.POINTER(u32)]
        lib.cref_keys.restype, lib.cref_keys.argtypes = C.c_uint64, [vp]
        lib.cref_get_json.restype, lib.cref_get_json.argtypes = C.c_int, [vp, C.c_char_p, vp, vp, u32, C.POINTER(u32)]
        lib.cref_set_now.restype, lib.cref_set_now.argtypes = None, [vp, C.c_uint64]
        lib.cref_pending_json.restype, lib.cref_pending_json.argtypes = C.c_int, [vp, C.c_char_p, vp, u32, C.POINTER(u32)]
        _lib = lib
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class CRef:
    def __init__(self, flags=0):
        self.lib = load()
        self.h = C.c_void_p(self.lib.cref_create(flags))

    def close(self):
        if self.h:
            self.lib.cref_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_agent_state(self, agent_id, status):
        code = AGENT_STATUS_CODES[status] if isinstance(status, str) else int(status)
        return self.lib.cref_set_agent_state(self.h, agent_id.encode(), code)

    def drop_agent(self, agent_id):
        self.lib.cref_drop_agent(self.h, agent_id.encode())

    def ingest(self, recs, want_verdicts=True):
        out = np.zeros(len(recs), dtype=verdict_dtype)
        self.lib.cref_ingest(self.h, _p(recs), len(recs), _p(out))
        return out, 0

    def complete(self, outs, want_results=True):
        res = np.zeros(len(outs), dtype=np.int32)
        self.lib.cref_complete(self.h, _p(outs), len(outs), _p(res))
        return res

    def replay_scan(self, with_records=True, cap=1 << 12):
        while True:
            disp = np.zeros(cap, dtype=dispatch_dtype)
            recs = np.zeros(cap, dtype=record_dtype)
            n = C.c_uint32()
            rc = self.lib.cref_scan(self.h, _p(disp), _p(recs), cap, C.byref(n))
            if rc == AGR_ECAP:
                cap = n.value
                continue
            return disp[: n.value], recs[: n.value]

    def pending(self, agent_id, cap=1 << 12):
        while True:
            out = np.zeros(cap, dtype=record_dtype)
            n = C.c_uint32()
            rc = self.lib.cref_pending(self.h, agent_id.encode(), _p(out), cap, C.byref(n))
            if rc == AGR_ECAP:
                cap = n.value
                continue
            return out[: n.value]

    def get_record(self, agent_id, request_id):
        out = np.zeros(1, dtype=record_dtype)
        rid = np.frombuffer(request_id, dtype=np.uint8).copy()
        rc = self.lib.cref_get_record(self.h, agent_id.encode(), _p(rid), _p(out))
        return None if rc == AGR_ENOTFOUND else out[0]

    def set_now(self, now):
        """The Redis server's clock (key TTLs); same unit as the records' times."""
        self.lib.cref_set_now(self.h, int(now))

    def get_json(self, agent_id, request_id):
        """The stored value of agent:{a}:requests:{r}: the C port's own json.Marshal(request), after every round trip."""
        out = np.zeros(1 << 14, dtype=np.uint8)
        rid = np.frombuffer(request_id, dtype=np.uint8).copy()
        n = C.c_uint32()
        rc = self.lib.cref_get_json(self.h, agent_id.encode(), _p(rid), _p(out), out.size, C.byref(n))
        return None if rc == AGR_ENOTFOUND else out[: n.value].tobytes()

    def pending_json(self, agent_id):
        cap = 1 << 20
        while True:
            out = np.zeros(cap, dtype=np.uint8)
            n = C.c_uint32()
            rc = self.lib.cref_pending_json(self.h, agent_id.encode(), _p(out), cap, C.byref(n))
            if rc == AGR_ECAP:
                cap = int(n.value)
                continue
            return out[: n.value].tobytes()

    def list(self, agent_id, which, cap=1 << 12):
        while True:
            ids = np.zeros((cap, 16), dtype=np.uint8)
            n = C.c_uint32()
            rc = self.lib.cref_list(self.h, agent_id.encode(), which, _p(ids), cap, C.byref(n))
            if rc == AGR_ECAP:
                cap = n.value
                continue
            return ids[: n.value]
