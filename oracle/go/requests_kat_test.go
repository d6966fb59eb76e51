// requests_kat_test.go — pins tests/golden/kats.json (hand-derived) to the UNMODIFIED reference.
//
// NOT part of this repository's build (there is no Go toolchain in the build image): a maintainer runs it inside a checkout of
// oso95/Agentainer-lab @ 42c3607 (oracle/go/README.md).  It lives in package `requests` so that it can call the reference's
// own unexported ReplayWorker.processAgents; Redis is github.com/alicebob/miniredis/v2 (go.mod in this directory pins it).
//
// What is the reference's own code here, and what is test glue:
//   reference  requests.Manager.StoreRequest / StoreResponse / GetPendingRequests / MarkRequestFailed  (requests.go:64-275)
//              requests.ReplayWorker.processAgents -> processPendingRequests -> replayRequest           (replay_worker.go:58-163)
//              encoding/json of requests.Request (the stored wire form), go-redis, and the Redis list / TTL semantics
//   glue       a handler on localhost:8081 that makes the calls of proxyToAgentHandler + interceptTransport.RoundTrip in their
//              order (internal/api/server.go:493-615, cited line by line below).  The real handler needs agent.Manager, i.e. a
//              Docker daemon; the glue reads the agent's status from the same Redis key the reference reads (agent:{id},
//              agent.go:372-390) and fakes only the agent container's answer.  The manual replay handler (server.go:681-751)
//              and agent.Manager.Remove's queue cleanup (agent.go:349-359) are restated the same way.
// The run writes tests/golden/from_reference/kats.json in the schema of tests/golden/kats.json; tests/kats.py prefers that
// file when it exists and the CPU / GPU parity tests then say PINNED.
//
//   go test ./internal/requests/ -run TestAgrKATs -args -scenarios=<repo>/tests/golden/kat_scenarios.json -out=<repo>/tests/golden/from_reference/kats.json
package requests

import (
	"bytes"
	"context"
	"encoding/json"
	"errors"
	"flag"
	"fmt"
	"io"
	"net"
	"net/http"
	"os"
	"sort"
	"strings"
	"sync"
	"testing"

	"github.com/alicebob/miniredis/v2"
	"github.com/go-redis/redis/v8"
)

var (
	flagScenarios = flag.String("scenarios", "kat_scenarios.json", "tests/golden/kat_scenarios.json of the B200 repo")
	flagOut       = flag.String("out", "kats_from_reference.json", "where to write the observables")
)

type katEvent struct {
	Op       string                   `json:"op"`
	Agent    string                   `json:"agent"`
	Status   string                   `json:"status"`
	Rid      string                   `json:"rid"`
	Replay   bool                     `json:"replay"`
	ReplayOf string                   `json:"replay_of"`
	Backend  []interface{}            `json:"backend"`
	Backends map[string][]interface{} `json:"backends"`
	Flip     []interface{}            `json:"flip"`
}

type katFile struct {
	Agents    map[string]string     `json:"agents"`
	Scenarios map[string][]katEvent `json:"scenarios"`
}

type katResult struct {
	Verdicts [][]interface{}                `json:"verdicts"`
	Manual   []int                          `json:"manual"`
	Ticks    [][][]string                   `json:"ticks"`
	Lists    map[string]map[string][]string `json:"lists"`
	Records  map[string][]interface{}       `json:"records"`
	RawJSON  map[string]string              `json:"stored_json"` // the Redis value of every record, ids and times as stored
}

// one scenario's world: miniredis + the reference Manager / ReplayWorker + the proxy glue on :8081
type world struct {
	t        *testing.T
	mr       *miniredis.Miniredis
	rdb      *redis.Client
	mgr      *Manager
	worker   *ReplayWorker
	names    map[string]string // "A" -> agent-17...
	sym      map[string]string // minted uuid -> r1, r2, ...
	real     map[string]string // r1 -> minted uuid
	mu       sync.Mutex
	backend  func(agent, sym string) []interface{} // what the agent container does for a forwarded request
	dispatch [][]string                             // current tick: [agent symbol, request symbol] in the order the worker sent them
	flip     []interface{}
	nReplay  int
	res      *katResult
}

func (w *world) agentName(a string) string {
	if n, ok := w.names[a]; ok {
		return n
	}
	return a
}
func (w *world) agentSym(name string) string {
	for k, v := range w.names {
		if v == name {
			return k
		}
	}
	return name
}

// saveAgent (agent.go:510-530) writes the agent document the path reads back; only `status` matters to it
func (w *world) setAgent(a, status string) {
	name := w.agentName(a)
	doc, _ := json.Marshal(map[string]interface{}{"id": name, "name": a, "status": status})
	if err := w.rdb.Set(context.Background(), "agent:"+name, doc, 0).Err(); err != nil {
		w.t.Fatal(err)
	}
}

// agent.Manager.Remove, the part that touches the path (agent.go:343-359): DEL agent:{id} and the three lists
func (w *world) removeAgent(a string) {
	name := w.agentName(a)
	ctx := context.Background()
	w.rdb.Del(ctx, "agent:"+name)
	for _, q := range []string{"pending", "completed", "failed"} {
		w.rdb.Del(ctx, fmt.Sprintf("agent:%s:requests:%s", name, q))
	}
}

func (w *world) agentStatus(name string) (string, bool) {
	data, err := w.rdb.Get(context.Background(), "agent:"+name).Result() // GetAgent, agent.go:376-377
	if err != nil {
		return "", false
	}
	var doc map[string]interface{}
	if json.Unmarshal([]byte(data), &doc) != nil {
		return "", false
	}
	s, _ := doc["status"].(string)
	return s, true
}

var errTransport = errors.New("EOF")
var errDial = errors.New("dial tcp 172.18.0.2:8000: connect: connection refused")

// the agent container: answers, refuses the connection, or breaks the transport
func (w *world) callAgent(b []interface{}) (*http.Response, error) {
	switch b[0].(string) {
	case "response":
		code := int(b[1].(float64))
		return &http.Response{StatusCode: code, Header: http.Header{}, Body: io.NopCloser(bytes.NewReader(nil))}, nil
	case "dial":
		return nil, errDial
	default:
		return nil, errTransport
	}
}

// proxyToAgentHandler (server.go:493-573) + interceptTransport.RoundTrip (:583-615), as an http.Handler on :8081
func (w *world) ServeHTTP(rw http.ResponseWriter, r *http.Request) {
	parts := strings.SplitN(strings.TrimPrefix(r.URL.Path, "/agent/"), "/", 2)
	agentID := parts[0]
	ctx := r.Context()
	status, ok := w.agentStatus(agentID) // :498
	if !ok {
		w.note(r, 4, 404, false, "")
		http.Error(rw, "Agent not found", http.StatusNotFound) // :499-502
		return
	}
	requestID := ""                                          // :505
	isReplay := r.Header.Get("X-Agentainer-Replay") == "true" // :506
	stored := false
	if !isReplay { // :508 (features.request_persistence is on)
		storedReq, err := w.mgr.StoreRequest(ctx, agentID, r) // :510
		if err == nil {
			requestID = storedReq.ID // :515
			stored = true
			w.learn(storedReq.ID, r.Header.Get("X-Kat-Symbol"))
			r.Header.Set("X-Agentainer-Request-ID", requestID) // :517
		}
	} else {
		requestID = r.Header.Get("X-Agentainer-Request-ID") // :519-522
	}
	if status != "running" { // :525
		if requestID != "" { // :526
			w.note(r, 2, 202, stored, requestID)
			rw.WriteHeader(http.StatusAccepted) // :527-536
			return
		}
		w.note(r, 3, 503, stored, requestID)
		http.Error(rw, "Agent is not running", http.StatusServiceUnavailable) // :539-540
		return
	}
	w.note(r, 1, 0, stored, requestID)
	b := w.backend(agentID, w.symOf(requestID, r))
	if b[0].(string) == "client" { // the caller gives up before any response: no server-side effect (Q23 modelled as one increment, by the worker)
		hj, _ := rw.(http.Hijacker)
		conn, _, _ := hj.Hijack()
		conn.Close()
		return
	}
	resp, err := w.callAgent(b)
	if err == nil { // RoundTrip :588-594
		if requestID != "" {
			_ = w.mgr.StoreResponse(ctx, agentID, requestID, resp)
		}
		rw.WriteHeader(resp.StatusCode)
		w.afterReplay(r)
		return
	}
	if requestID != "" { // :597
		msg := err.Error()
		if strings.Contains(msg, "connection refused") || strings.Contains(msg, "no such host") || strings.Contains(msg, "dial tcp") { // :600-605
			// stays pending
		} else {
			_ = w.mgr.MarkRequestFailed(ctx, agentID, requestID, err) // :606-611
		}
	}
	rw.WriteHeader(http.StatusBadGateway) // ReverseProxy's default ErrorHandler
	w.afterReplay(r)
}

func (w *world) learn(id, sym string) {
	w.mu.Lock()
	defer w.mu.Unlock()
	if sym != "" {
		w.sym[id] = sym
		w.real[sym] = id
	}
}
func (w *world) symOf(id string, r *http.Request) string {
	w.mu.Lock()
	defer w.mu.Unlock()
	if s, ok := w.sym[id]; ok {
		return s
	}
	return r.Header.Get("X-Kat-Symbol")
}

// verdict of a client request (replays sent by the worker are recorded as the tick's dispatch order instead)
func (w *world) note(r *http.Request, code, httpStatus int, stored bool, requestID string) {
	if r.Header.Get("X-Kat-Client") == "1" {
		w.res.Verdicts = append(w.res.Verdicts, []interface{}{code, httpStatus, stored, requestID != ""})
		return
	}
	if r.Header.Get("X-Agentainer-Replay") == "true" {
		w.mu.Lock()
		w.dispatch = append(w.dispatch, []string{w.agentSym(strings.SplitN(strings.TrimPrefix(r.URL.Path, "/agent/"), "/", 2)[0]), w.sym[r.Header.Get("X-Agentainer-Request-ID")]})
		w.mu.Unlock()
	}
}

// a mid-tick status write right after the k-th replay (KAT-F)
func (w *world) afterReplay(r *http.Request) {
	if r.Header.Get("X-Kat-Client") == "1" || w.flip == nil {
		return
	}
	w.nReplay++
	if w.nReplay == int(w.flip[0].(float64)) {
		w.setAgent(w.flip[1].(string), w.flip[2].(string))
	}
}

// replayRequestHandler (server.go:681-751)
func (w *world) manualReplay(a, sym string, b []interface{}) int {
	ctx := context.Background()
	agentID := w.agentName(a)
	requestID, ok := w.real[sym]
	if !ok {
		requestID = "00000000-0000-4000-8000-0000000000" + fmt.Sprintf("%02d", len(sym)) // an id nobody stored
	}
	if _, err := w.rdb.Get(ctx, fmt.Sprintf("agent:%s:requests:%s", agentID, requestID)).Result(); err != nil { // :687-692
		return 404
	}
	status, found := w.agentStatus(agentID) // :702-706
	if !found {
		return 404
	}
	if status != "running" { // :708-711
		return 503
	}
	resp, err := w.callAgent(b)
	if err != nil || b[0].(string) == "client" { // :726-733: ANY client error, a refused connection included
		if err == nil {
			err = errTransport
		}
		_ = w.mgr.MarkRequestFailed(ctx, agentID, requestID, err)
		return 502
	}
	_ = w.mgr.StoreResponse(ctx, agentID, requestID, resp) // :739
	return 200
}

func (w *world) clientRequest(e katEvent) {
	name := w.agentName(e.Agent)
	req, _ := http.NewRequest("POST", "http://localhost:8081/agent/"+name+"/chat", bytes.NewReader([]byte(`{"message":"hi"}`)))
	req.Header.Set("Content-Type", "application/json")
	req.Header.Set("X-Kat-Client", "1")
	req.Header.Set("X-Kat-Symbol", e.Rid)
	if e.Replay {
		req.Header.Set("X-Agentainer-Replay", "true")
		if e.ReplayOf != "" {
			id, ok := w.real[e.ReplayOf]
			if !ok {
				id = "00000000-0000-4000-8000-0000000000" + fmt.Sprintf("%02d", len(e.ReplayOf))
			}
			req.Header.Set("X-Agentainer-Request-ID", id)
		}
	}
	b := e.Backend
	w.backend = func(string, string) []interface{} { return b }
	resp, err := http.DefaultClient.Do(req)
	if err == nil {
		resp.Body.Close()
	}
}

func TestAgrKATs(t *testing.T) {
	raw, err := os.ReadFile(*flagScenarios)
	if err != nil {
		t.Skipf("no scenario file (%v): pass -args -scenarios=<repo>/tests/golden/kat_scenarios.json", err)
	}
	var kf katFile
	if err := json.Unmarshal(raw, &kf); err != nil {
		t.Fatal(err)
	}
	out := map[string]*katResult{}
	names := make([]string, 0, len(kf.Scenarios))
	for k := range kf.Scenarios {
		names = append(names, k)
	}
	sort.Strings(names)
	for _, name := range names {
		mr := miniredis.RunT(t)
		rdb := redis.NewClient(&redis.Options{Addr: mr.Addr()})
		w := &world{t: t, mr: mr, rdb: rdb, names: kf.Agents, sym: map[string]string{}, real: map[string]string{},
			res: &katResult{Lists: map[string]map[string][]string{}, Records: map[string][]interface{}{}, RawJSON: map[string]string{}, Ticks: [][][]string{}, Manual: []int{}, Verdicts: [][]interface{}{}}}
		w.mgr = NewManager(rdb)
		w.worker = NewReplayWorker(w.mgr, rdb) // a SECOND stateless manager in the reference (main.go:335): the same Redis
		ln, err := net.Listen("tcp", "127.0.0.1:8081") // replay_worker.go:133 hard-codes localhost:8081 (Q18)
		if err != nil {
			t.Fatalf("port 8081 must be free: %v", err)
		}
		srv := &http.Server{Handler: w}
		go srv.Serve(ln)
		seen := map[string]bool{}
		for _, e := range kf.Scenarios[name] {
			switch e.Op {
			case "agent":
				w.setAgent(e.Agent, e.Status)
				seen[e.Agent] = true
			case "remove":
				w.removeAgent(e.Agent)
			case "req":
				seen[e.Agent] = true
				w.clientRequest(e)
			case "manual":
				w.res.Manual = append(w.res.Manual, w.manualReplay(e.Agent, e.Rid, e.Backend))
			case "tick":
				backends := e.Backends
				w.backend = func(_ string, sym string) []interface{} {
					if b, ok := backends[sym]; ok {
						return b
					}
					return []interface{}{"response", float64(200)}
				}
				w.dispatch, w.flip, w.nReplay = nil, e.Flip, 0
				w.worker.processAgents(context.Background()) // the reference's own tick
				// KEYS order is undefined (Q9): canonicalise to agent registration order, FIFO inside an agent is the worker's
				sort.SliceStable(w.dispatch, func(i, j int) bool { return w.dispatch[i][0] < w.dispatch[j][0] })
				if w.dispatch == nil {
					w.dispatch = [][]string{}
				}
				w.res.Ticks = append(w.res.Ticks, w.dispatch)
				if e.Flip != nil && w.nReplay < int(e.Flip[0].(float64)) {
					w.setAgent(e.Flip[1].(string), e.Flip[2].(string))
				}
				w.flip = nil
			}
		}
		ctx := context.Background()
		for a := range seen {
			agentID := w.agentName(a)
			lists := map[string][]string{}
			for _, q := range []string{"pending", "completed", "failed"} {
				ids, _ := rdb.LRange(ctx, fmt.Sprintf("agent:%s:requests:%s", agentID, q), 0, -1).Result()
				syms := []string{}
				for _, id := range ids {
					syms = append(syms, w.sym[id])
				}
				lists[q] = syms
			}
			w.res.Lists[a] = lists
		}
		for id, sym := range w.sym {
			keys, _ := rdb.Keys(ctx, "agent:*:requests:"+id).Result()
			for _, key := range keys {
				data, err := rdb.Get(ctx, key).Result()
				if err != nil {
					continue
				}
				var rec Request
				if json.Unmarshal([]byte(data), &rec) != nil {
					continue
				}
				code := 0
				if rec.Response != nil {
					code = rec.Response.StatusCode
				}
				k := w.agentSym(rec.AgentID) + "/" + sym
				w.res.Records[k] = []interface{}{string(rec.Status), rec.RetryCount, code}
				w.res.RawJSON[k] = data
			}
		}
		out[name] = w.res
		srv.Close()
		rdb.Close()
		mr.Close()
	}
	doc := map[string]interface{}{
		"_provenance": "OUTPUT OF THE REFERENCE: oso95/Agentainer-lab @ 42c3607 internal/requests (Manager, ReplayWorker) driven by oracle/go/requests_kat_test.go over miniredis; proxy / manual-replay call order restated from internal/api/server.go (see the test's header)",
		"kats":        out,
	}
	buf, _ := json.MarshalIndent(doc, "", " ")
	if err := os.WriteFile(*flagOut, buf, 0o644); err != nil {
		t.Fatal(err)
	}
	t.Logf("wrote %d scenarios to %s", len(out), *flagOut)
}
