"""
oracle/model.py — CPU restatement of Agentainer's request persistence / replay / proxy-decision path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker.

PARITY UNPINNED.  The reference (oso95/Agentainer-lab @ 42c3607) ships no tests, golden vectors or fixtures for
this path, and neither Go nor Redis exists in the build image, so this restatement cannot be checked against a
run of the reference.  Its authority is the reference source text, followed function by function below, plus
the documented semantics of the six Redis commands the path uses.  The known-answer traces in
tests/golden/kats.json are hand-derived from the same source and are labelled as such.

Third-party semantics restated here (sources absent from the reference tree):
  * Redis server, image redis:7-alpine, minor version unpinned (docker-compose.yml:8): SET key val EX ttl,
    GET, RPUSH (append at tail), LREM key 1 val (remove first match scanning from the head), LRANGE key 0 -1,
    KEYS pattern (order undefined), DEL.  An empty list does not exist as a key.
  * github.com/go-redis/redis/v8 v8.11.5 (go.mod:9): thin client, GET miss -> redis.Nil error.
  * github.com/google/uuid v1.6.0 (go.mod:10): random v4 IDs -> supplied by the event stream here.
  * time.Now() (requests.go:96,146,164) -> logical sequence numbers supplied by the event stream.
  * net/http/httputil.ReverseProxy (Go 1.23): a RoundTrip error is answered with 502 Bad Gateway.

All paths cited are relative to the reference tree.
"""
from __future__ import annotations

import copy
import fnmatch
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from . import gojson

# RequestStatus, internal/requests/requests.go:19-24
STATUS_PENDING = "pending"
STATUS_PROCESSING = "processing"   # declared, never assigned anywhere (Q1)
STATUS_COMPLETED = "completed"
STATUS_FAILED = "failed"

# agent.Status, internal/agent/agent.go:23-29
AGENT_CREATED, AGENT_RUNNING, AGENT_STOPPED, AGENT_PAUSED, AGENT_FAILED = (
    "created", "running", "stopped", "paused", "failed")


class RedisNil(Exception):
    """go-redis redis.Nil: key does not exist."""


class MiniRedis:
    """The six commands of the path (+DEL), single-threaded like the server they model."""

    def __init__(self) -> None:
        self.strings: Dict[str, object] = {}
        self.lists: Dict[str, List[str]] = {}
        self.ops: List[Tuple] = []          # command trace, used by KAT-A
        self.trace = False
        # key expiry (SET ... EX): the server's clock is set by the event stream (same unit as the stream's times,
        # nanoseconds by default); it stays 0 in streams that never advance it, so nothing expires there
        self.now = 0
        self.ticks_per_second = 1_000_000_000
        self.expires_at: Dict[str, int] = {}

    def _t(self, *op) -> None:
        if self.trace:
            self.ops.append(op)

    def set(self, key: str, value, ttl_s: int = 0) -> None:   # SET key value EX ttl (TTL reset on every SET, Q11)
        self._t("SET", key)
        self.strings[key] = value
        if ttl_s:
            self.expires_at[key] = self.now + ttl_s * self.ticks_per_second
        else:
            self.expires_at.pop(key, None)

    def get(self, key: str):
        self._t("GET", key)
        if key in self.expires_at and self.now >= self.expires_at[key]:    # expired keys are gone
            del self.strings[key], self.expires_at[key]
        if key not in self.strings:
            raise RedisNil(key)
        return self.strings[key]

    def rpush(self, key: str, value: str) -> int:
        self._t("RPUSH", key, value)
        self.lists.setdefault(key, []).append(value)
        return len(self.lists[key])

    def lrem(self, key: str, count: int, value: str) -> int:
        """LREM key 1 value: remove the first occurrence scanning head -> tail.  O(len)."""
        assert count == 1
        self._t("LREM", key, value)
        lst = self.lists.get(key)
        if not lst:
            return 0
        try:
            lst.remove(value)
        except ValueError:
            return 0
        if not lst:
            del self.lists[key]             # empty lists do not exist as keys
        return 1

    def lrange_all(self, key: str) -> List[str]:               # LRANGE key 0 -1
        self._t("LRANGE", key)
        return list(self.lists.get(key, []))

    def keys(self, pattern: str) -> List[str]:                 # KEYS pattern; order is undefined in Redis
        self._t("KEYS", pattern)
        ks = [k for k in list(self.strings) + list(self.lists) if fnmatch.fnmatchcase(k, pattern)]
        return ks

    def delete(self, key: str) -> None:
        self._t("DEL", key)
        self.strings.pop(key, None)
        self.expires_at.pop(key, None)
        self.lists.pop(key, None)


# ----------------------------------------------------------------------------------------------------------
# internal/requests/requests.go
# ----------------------------------------------------------------------------------------------------------
@dataclass
class HttpRequest:
    """What StoreRequest reads from *http.Request (requests.go:64-97) plus the stream-supplied ID / time."""
    method: str
    path: str                     # URL.Path, still carrying /agent/{id} (Q3); query dropped (Q4)
    headers: Dict[str, str]       # first value per key (Q5)
    body: bytes
    new_id: str                   # what uuid.New().String() returns for this call (requests.go:87)
    now: int                      # what time.Now() returns (requests.go:96)


@dataclass
class HttpResponse:
    status_code: int
    headers: Dict[str, str] = field(default_factory=dict)
    body: bytes = b""
    now: int = 0                  # time.Now() at requests.go:146,164


class Manager:
    """requests.Manager (requests.go:52-61): stateless wrapper over the Redis client."""

    def __init__(self, redis: MiniRedis) -> None:
        self.redis = redis

    # requests.go:64-117
    def store_request(self, agent_id: str, req: HttpRequest) -> dict:
        request = {
            "id": req.new_id,                 # :87
            "agent_id": agent_id,             # :88
            "method": req.method,             # :89
            "path": req.path,                 # :90
            "headers": dict(req.headers),     # :78-83,91
            "body": bytes(req.body),          # :66-75,92
            "status": STATUS_PENDING,         # :93
            "retry_count": 0,                 # :94
            "max_retries": 3,                 # :95
            "created_at": req.now,            # :96
            "processed_at": None,
            "response": None,
            "error": "",
        }
        key = f"agent:{agent_id}:requests:{request['id']}"           # :100
        self.redis.set(key, copy.deepcopy(request), 24 * 3600)       # :101-108
        self.redis.rpush(f"agent:{agent_id}:requests:pending", request["id"])   # :111-114
        return request

    # requests.go:120-194
    def store_response(self, agent_id: str, request_id: str, resp: HttpResponse) -> None:
        response = {"status_code": resp.status_code, "headers": dict(resp.headers),
                    "body": bytes(resp.body), "received_at": resp.now}          # :142-147
        key = f"agent:{agent_id}:requests:{request_id}"                          # :150
        try:
            request = copy.deepcopy(self.redis.get(key))                         # :153
        except RedisNil as e:
            raise KeyError(f"failed to get request: {e}")                        # :154-156
        gojson.unmarshal_strings(request)                                        # :158-161 json.Unmarshal
        request["response"] = response                                           # :165
        request["status"] = STATUS_COMPLETED                                     # :166
        request["processed_at"] = resp.now                                       # :167
        self.redis.set(key, request, 24 * 3600)                                  # :175
        self.redis.lrem(f"agent:{agent_id}:requests:pending", 1, request_id)     # :180-184 (error only logged)
        self.redis.rpush(f"agent:{agent_id}:requests:completed", request_id)     # :187-191

    # requests.go:197-225
    def get_pending_requests(self, agent_id: str) -> List[dict]:
        ids = self.redis.lrange_all(f"agent:{agent_id}:requests:pending")        # :201
        out = []
        for rid in ids:
            try:
                data = self.redis.get(f"agent:{agent_id}:requests:{rid}")        # :208-209
            except RedisNil:
                continue                                                         # :210-213 (Q10: stays in list)
            out.append(gojson.unmarshal_strings(copy.deepcopy(data)))            # :215-221 json.Unmarshal
        return out

    # requests.go:228-275
    def mark_request_failed(self, agent_id: str, request_id: str, err: str) -> None:
        key = f"agent:{agent_id}:requests:{request_id}"                          # :229
        try:
            request = copy.deepcopy(self.redis.get(key))                         # :232
        except RedisNil as e:
            raise KeyError(f"failed to get request: {e}")                        # :233-235
        gojson.unmarshal_strings(request)                                        # :237-240 json.Unmarshal
        request["status"] = STATUS_FAILED                                        # :243
        request["error"] = err                                                   # :244
        request["retry_count"] += 1                                              # :245
        if request["retry_count"] < request["max_retries"]:                      # :248
            request["status"] = STATUS_PENDING                                   # :249 (keeps queue position, Q11)
        else:
            self.redis.rpush(f"agent:{agent_id}:requests:failed", request_id)    # :252-255
            self.redis.lrem(f"agent:{agent_id}:requests:pending", 1, request_id) # :258-261
        self.redis.set(key, request, 24 * 3600)                                  # :270


# ----------------------------------------------------------------------------------------------------------
# internal/agent/agent.go (only what the path reads / what writes the path's keys)
# ----------------------------------------------------------------------------------------------------------
class AgentStore:
    """agent:{id} JSON documents: saveAgent (agent.go:510-530), GetAgent (:372-390), Remove cleanup (:343-367)."""

    def __init__(self, redis: MiniRedis) -> None:
        self.redis = redis
        self.order: List[str] = []     # registration order; the canonical cross-agent order (Q9)

    def save(self, agent_id: str, status: str) -> None:
        if agent_id not in self.order:
            self.order.append(agent_id)
        self.redis.set(f"agent:{agent_id}", {"id": agent_id, "status": status})

    def get_agent(self, agent_id: str) -> dict:                 # agent.go:372-390
        try:
            return self.redis.get(f"agent:{agent_id}")
        except RedisNil:
            raise KeyError("agent not found")

    def remove(self, agent_id: str) -> None:                    # agent.go:343-367
        self.redis.delete(f"agent:{agent_id}")                  # :344
        for q in ("pending", "completed", "failed"):            # :349-359
            self.redis.delete(f"agent:{agent_id}:requests:{q}")
        # :361-367 scans request:{id}:* which never matches agent:{id}:requests:{r} -> records orphaned (Q17)


# ----------------------------------------------------------------------------------------------------------
# internal/api/server.go:493-615 — decision part of the proxy
# ----------------------------------------------------------------------------------------------------------
# what the agent side does with a forwarded request; chosen by the event stream
BACKEND_RESPONSE = "response"      # any HTTP response (status code attached)
BACKEND_DIAL_ERR = "dial"          # "connection refused" / "no such host" / "dial tcp"
BACKEND_ERROR = "error"            # any other transport error (EOF, reset, ...)
BACKEND_CLIENT_ERR = "client"      # replay only: the worker's own http.Client fails (30 s timeout); Q23 modelled
                                   # as the worker-side MarkRequestFailed alone

V_FORWARD, V_QUEUED, V_UNAVAILABLE, V_NOT_FOUND = 1, 2, 3, 4


@dataclass
class Verdict:
    code: int
    http_status: int           # status the CLIENT of the proxy sees when the proxy itself answers (0 = forwarded)
    request_id: str            # requestID variable of proxyToAgentHandler
    stored: bool               # StoreRequest ran successfully
    replay: bool


class Proxy:
    def __init__(self, redis: MiniRedis, agents: AgentStore, persistence: bool = True) -> None:
        self.redis = redis
        self.agents = agents
        self.request_mgr = Manager(redis)          # server.go:62 (the server's own Manager, Q19)
        self.persistence = persistence             # cfg.Features.RequestPersistence (config.go:70)

    # server.go:493-541 — everything before the reverse proxy is constructed
    def decide(self, agent_id: str, req: HttpRequest) -> Verdict:
        try:
            agent_obj = self.agents.get_agent(agent_id)                         # :498
        except KeyError:
            return Verdict(V_NOT_FOUND, 404, "", False, False)                  # :499-502
        request_id = ""                                                         # :505
        is_replay = req.headers.get("X-Agentainer-Replay", "") == "true"        # :506
        stored = False
        if self.persistence and not is_replay:                                  # :508
            stored_req = self.request_mgr.store_request(agent_id, req)          # :510
            request_id = stored_req["id"]                                       # :515
            stored = True
            req.headers["X-Agentainer-Request-ID"] = request_id                 # :517 (after the store: not persisted)
        elif is_replay:
            request_id = req.headers.get("X-Agentainer-Request-ID", "")         # :519-522
        if agent_obj["status"] != AGENT_RUNNING:                                # :525
            if self.persistence and request_id != "":                           # :526
                return Verdict(V_QUEUED, 202, request_id, stored, is_replay)    # :528-536
            return Verdict(V_UNAVAILABLE, 503, request_id, stored, is_replay)   # :539-540
        return Verdict(V_FORWARD, 0, request_id, stored, is_replay)             # :546-572

    # interceptTransport.RoundTrip, server.go:583-615.  Returns the status the proxy's client sees.
    def round_trip(self, agent_id: str, request_id: str, backend: Tuple, now: int) -> int:
        kind = backend[0]
        if kind == BACKEND_RESPONSE:
            if request_id != "":                                                # :588
                try:
                    self.request_mgr.store_response(agent_id, request_id,
                                                    HttpResponse(backend[1], now=now))   # :590
                except KeyError:
                    pass                                                        # :591-593 only logs
            return backend[1]
        # err != nil
        if request_id != "":                                                    # :597
            if kind == BACKEND_DIAL_ERR:
                pass                                                            # :600-605 stays pending
            else:
                try:
                    self.request_mgr.mark_request_failed(agent_id, request_id, "transport error")  # :608
                except KeyError:
                    pass                                                        # :609-611 only logs
        return 502                                                              # ReverseProxy default ErrorHandler

    def handle(self, agent_id: str, req: HttpRequest, backend: Tuple) -> Tuple[Verdict, int]:
        """A full pass of proxyToAgentHandler.  Returns (verdict, status seen by the caller)."""
        v = self.decide(agent_id, req)
        if v.code != V_FORWARD:
            return v, v.http_status
        if backend[0] == BACKEND_CLIENT_ERR:
            # the caller gave up before any response; server side sees a cancelled context (Q23).  Modelled as
            # no server-side effect; the worker's MarkRequestFailed is applied by the caller.
            return v, -1
        return v, self.round_trip(agent_id, v.request_id, backend, req.now)


# ----------------------------------------------------------------------------------------------------------
# internal/requests/replay_worker.go
# ----------------------------------------------------------------------------------------------------------
class ReplayWorker:
    def __init__(self, redis: MiniRedis, agents: AgentStore, proxy: Proxy) -> None:
        self.redis = redis
        self.agents = agents
        self.proxy = proxy                      # replay goes back through http://localhost:8081/agent/... (:133)
        self.manager = Manager(redis)           # main.go:335 (the worker's own Manager, Q19)

    @staticmethod
    def extract_agent_id(key: str) -> str:      # :192-199
        parts = key.split(":")
        return parts[1] if len(parts) >= 2 else ""

    def is_agent_running(self, agent_id: str) -> bool:          # :166-189
        try:
            data = self.redis.get(f"agent:{agent_id}")
        except RedisNil:
            return False
        return data.get("status") == AGENT_RUNNING

    # :58-87.  backend_for(agent_id, request) -> backend tuple decides what happens to each replayed request;
    # on_replay(agent_id, request_id, k) is called after the k-th replay of the tick (mid-tick status flips).
    def process_agents(self, backend_for: Callable[[str, dict], Tuple], now: int,
                       on_replay: Optional[Callable[[str, str, int], None]] = None) -> List[Tuple[str, str]]:
        keys = self.redis.keys("agent:*:requests:pending")                       # :60
        # Redis does not define the KEYS order (Q9): canonicalise to agent registration order
        rank = {a: i for i, a in enumerate(self.agents.order)}
        keys.sort(key=lambda k: rank.get(self.extract_agent_id(k), 1 << 60))
        dispatched: List[Tuple[str, str]] = []
        count = [0]
        for key in keys:                                                          # :68
            agent_id = self.extract_agent_id(key)                                 # :70
            if agent_id == "":
                continue
            if not self.is_agent_running(agent_id):                               # :76-81
                continue
            self.process_pending_requests(agent_id, backend_for, now, dispatched, on_replay, count)  # :85
        return dispatched

    # :90-117
    def process_pending_requests(self, agent_id, backend_for, now, dispatched, on_replay, count) -> None:
        reqs = self.manager.get_pending_requests(agent_id)                        # :91 (snapshot)
        for req in reqs:                                                          # :99
            if req["status"] == STATUS_PROCESSING or req["retry_count"] >= req["max_retries"]:   # :101
                continue
            dispatched.append((agent_id, req["id"]))
            err = self.replay_request(agent_id, req, backend_for(agent_id, req), now)   # :109
            if err is not None:
                try:
                    self.manager.mark_request_failed(agent_id, req["id"], err)    # :112 (return value ignored)
                except KeyError:
                    pass
            count[0] += 1
            if on_replay is not None:
                on_replay(agent_id, req["id"], count[0])

    # :120-163
    def replay_request(self, agent_id: str, req: dict, backend: Tuple, now: int) -> Optional[str]:
        path = req["path"]                                                        # :123
        prefix = f"/agent/{agent_id}"                                             # :124
        if path.startswith(prefix):                                               # :125
            path = path[len(prefix):]                                             # :126
            if path == "":
                path = "/"                                                        # :127-129
        headers = dict(req["headers"])                                            # :142-144
        headers["X-Agentainer-Request-ID"] = req["id"]                            # :147
        headers["X-Agentainer-Replay"] = "true"                                   # :148
        http_req = HttpRequest(req["method"], f"/agent/{agent_id}{path}", headers, req["body"],
                               new_id="", now=now)                                # :133-139
        _, status = self.proxy.handle(agent_id, http_req, backend)                # :151 (loops back through L4)
        if status < 0:
            return "request failed: client error"                                 # :152-154
        try:
            self.manager.store_response(agent_id, req["id"], HttpResponse(status, now=now))   # :158 (Q7: 2nd completion)
        except KeyError:
            pass                                                                  # :159-160 only logs
        return None


# ----------------------------------------------------------------------------------------------------------
# The whole path as one object + the observables parity is defined on
# ----------------------------------------------------------------------------------------------------------
class ReferencePath:
    """Server + worker wired like main.runServer (cmd/agentainer/main.go:284-356), driven by an event stream."""

    def __init__(self, persistence: bool = True) -> None:
        self.redis = MiniRedis()
        self.agents = AgentStore(self.redis)
        self.proxy = Proxy(self.redis, self.agents, persistence)
        self.worker = ReplayWorker(self.redis, self.agents, self.proxy)
        self.manager = self.proxy.request_mgr

    # ---- events
    def set_agent(self, agent_id: str, status: str) -> None:
        self.agents.save(agent_id, status)

    def remove_agent(self, agent_id: str) -> None:
        self.agents.remove(agent_id)

    def request(self, agent_id: str, req: HttpRequest, backend: Tuple) -> Tuple[Verdict, int]:
        return self.proxy.handle(agent_id, req, backend)

    def tick(self, backend_for, now: int, on_replay=None) -> List[Tuple[str, str]]:
        return self.worker.process_agents(backend_for, now, on_replay)

    def manual_replay(self, agent_id: str, request_id: str, backend: Tuple, now: int) -> int:
        """POST /agents/{id}/requests/{reqId}/replay — replayRequestHandler, internal/api/server.go:681-751.  Unlike the
        worker it talks to the agent DIRECTLY (http://{id}:8000 + the stored path, prefix and all, Q3): no proxy, no
        X-Agentainer-* headers, so ANY client error — a refused connection included — reaches MarkRequestFailed
        (:728-733; through the proxy a dial error would not count, Q12).  Returns the status the caller of the
        management API sees."""
        key = f"agent:{agent_id}:requests:{request_id}"                          # :687
        try:
            self.redis.get(key)                                                 # :688 storage.Get
        except RedisNil:
            return 404                                                          # :689-692 "Request not found"
        try:
            agent_obj = self.agents.get_agent(agent_id)                         # :702
        except KeyError:
            return 404                                                          # :703-706 "Agent not found"
        if agent_obj["status"] != AGENT_RUNNING:                                # :708
            return 503                                                          # :709-711 "Agent is not running"
        if backend[0] != BACKEND_RESPONSE:                                      # client.Do failed (:726)
            try:
                self.manager.mark_request_failed(agent_id, request_id, "transport error")   # :730
            except KeyError:
                pass
            return 502                                                          # :731
        try:
            self.manager.store_response(agent_id, request_id, HttpResponse(backend[1], now=now))   # :739
        except KeyError:
            pass                                                                # :740-742 only warns
        return 200                                                              # :744-751

    # ---- observables
    def lists(self, agent_id: str) -> Dict[str, List[str]]:
        return {q: self.redis.lrange_all(f"agent:{agent_id}:requests:{q}") for q in ("pending", "completed", "failed")}

    def record(self, agent_id: str, request_id: str) -> Optional[dict]:
        try:
            return self.redis.get(f"agent:{agent_id}:requests:{request_id}")
        except RedisNil:
            return None

    def record_state(self, agent_id: str, request_id: str) -> Optional[Tuple[str, int, int]]:
        r = self.record(agent_id, request_id)
        if r is None:
            return None
        resp = r["response"]["status_code"] if r["response"] else 0
        return (r["status"], r["retry_count"], resp)
