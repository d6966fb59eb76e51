// test_host.cpp — drives the C++ mirror of requests.Manager / ReplayWorker exactly as the Go server would:
// KAT-A, KAT-B, KAT-C through the mirrored method names, then 8 threads calling Decide/StoreResponse concurrently
// (one goroutine per HTTP request in the reference).  Exit code 0 == all checks passed.  Needs a B200.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "requests.hpp"
using namespace agentainer::requests;

#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<std::string> list_ids(agr_handle* h, const char* agent, int which) {
    uint8_t ids[4096][16]; uint32_t n = 0;
    if (agr_list(h, agent, which, ids, 4096, &n) < 0) return {};
    std::vector<std::string> out;
    for (uint32_t i = 0; i < n; ++i) out.push_back(FormatUUID(ids[i]));
    return out;
}

static int run(bool mint, bool combine) {
    agr_config cfg; memset(&cfg, 0, sizeof cfg); cfg.device = 0; cfg.slab_rows = 1 << 18; cfg.max_agents = 64; cfg.max_batch = 1 << 15;
    cfg.flags = AGR_CFG_PERSISTENCE | (mint ? AGR_CFG_MINT_IDS : 0u) | (combine ? AGR_CFG_COMBINE : 0u);
    agr_handle* h = nullptr;
    if (agr_create(&cfg, &h) < 0) { printf("agr_create: %s\n", agr_last_error()); return 2; }
    Manager mgr(h, mint);
    {   // a14: the reference's second manager (main.go:335) wraps the SAME handle; a different one is a wiring bug
        Manager second(h, mint);
        CHECK(Manager::ProcessHandle() == h);
        bool threw = false;
        try { Manager wrong(reinterpret_cast<agr_handle*>(&cfg), mint); } catch (const std::logic_error&) { threw = true; }
        CHECK(threw);
    }
    const char* A = "agent-1700000000000000001";
    HttpRequest post; post.Method = "POST"; post.Path = std::string("/agent/") + A + "/chat";
    post.Header["Content-Type"] = "application/json"; post.Body = {'{', '}'};

    // KAT-A: running agent, 200 OK
    CHECK(agr_set_agent_state(h, A, AGR_AGENT_RUNNING) == 0);
    Verdict v; CHECK(mgr.Decide(A, post, &v).empty());
    CHECK(v.Code == AGR_V_FORWARD && v.Stored && !v.RequestID.empty());
    Response ok; ok.StatusCode = 200;
    CHECK(mgr.StoreResponse(A, v.RequestID, ok).empty());
    CHECK(list_ids(h, A, AGR_LIST_PENDING).empty());
    CHECK(list_ids(h, A, AGR_LIST_COMPLETED) == std::vector<std::string>{v.RequestID});
    CHECK(!mgr.StoreResponse(A, "00000000-0000-4000-8000-000000000001", ok).empty());   // "failed to get request"

    // KAT-B / KAT-C: stopped agent queues three, start + one tick replays FIFO, completed holds every id twice (Q7)
    const char* B = "agent-1700000000000000002";
    CHECK(agr_set_agent_state(h, B, AGR_AGENT_STOPPED) == 1);
    std::vector<std::string> ids;
    for (int i = 0; i < 3; ++i) {
        HttpRequest r = post; r.Path = std::string("/agent/") + B + "/chat"; r.Body.push_back((uint8_t)('0' + i));
        CHECK(mgr.Decide(B, r, &v).empty());
        CHECK(v.Code == AGR_V_QUEUED && v.HTTPStatus == 202);
        ids.push_back(v.RequestID);
    }
    std::vector<Request> pend; CHECK(mgr.GetPendingRequests(B, &pend).empty());
    CHECK(pend.size() == 3 && pend[0].ID == ids[0] && pend[2].ID == ids[2] && pend[1].Body.back() == '1');
    CHECK(pend[0].Path == std::string("/agent/") + B + "/chat" && pend[0].Headers.at("Content-Type") == "application/json");
    CHECK(agr_set_agent_state(h, B, AGR_AGENT_RUNNING) == 1);
    std::vector<std::string> seen;
    ReplayWorker worker(&mgr, [&](const std::string& agent, const Request& req) {
        seen.push_back(req.ID);
        HttpRequest rr; rr.Method = req.Method; rr.Path = req.Path; rr.Header = req.Headers; rr.Body = req.Body;
        rr.Header["X-Agentainer-Request-ID"] = req.ID; rr.Header["X-Agentainer-Replay"] = "true";   // replay_worker.go:147-148
        Verdict pv; mgr.Decide(agent, rr, &pv);                              // loops back through the proxy
        if (pv.Code != AGR_V_FORWARD) return pv.HTTPStatus;
        Response r200; r200.StatusCode = 200;
        mgr.StoreResponse(agent, pv.RequestID, r200);                        // interceptTransport (server.go:588-594)
        return 200;
    });
    CHECK(worker.ProcessAgents() == 3);
    CHECK(seen == ids);
    CHECK(list_ids(h, B, AGR_LIST_PENDING).empty());
    CHECK((list_ids(h, B, AGR_LIST_COMPLETED) == std::vector<std::string>{ids[0], ids[0], ids[1], ids[1], ids[2], ids[2]}));

    // wire form + TTL through the mirror: a failed-then-answered record in the reference's JSON, then expiry
    {
        const char* D = "agent-1700000000000000004";
        CHECK(agr_set_agent_state(h, D, AGR_AGENT_STOPPED) == 2);
        uint64_t clock = 1700000000ull * 1000000000ull;
        mgr.SetClock([&] { return clock; });
        HttpRequest r = post; r.Path = std::string("/agent/") + D + "/x<y>"; r.Body = {'h', 'i'};
        Verdict jv; CHECK(mgr.Decide(D, r, &jv).empty() && jv.Code == AGR_V_QUEUED);
        std::string js; CHECK(mgr.GetRequestJSON(D, jv.RequestID, &js).empty());
        const std::string want0 = std::string("{\"id\":\"") + jv.RequestID + "\",\"agent_id\":\"" + D + "\",\"method\":\"POST\",\"path\":\"/agent/" + D +
            "/x\\u003cy\\u003e\",\"headers\":{\"Content-Type\":\"application/json\"},\"body\":\"aGk=\",\"status\":\"pending\",\"retry_count\":0,"
            "\"max_retries\":3,\"created_at\":\"2023-11-14T22:13:20Z\"}";
        CHECK(js == want0);
        clock += 1500000000ull;
        CHECK(mgr.MarkRequestFailed(D, jv.RequestID, "EOF").empty());
        clock += 1000000000ull;
        Response rr; rr.StatusCode = 201; rr.Headers["Server"] = "x"; rr.Body = {'o', 'k'};
        CHECK(mgr.StoreResponse(D, jv.RequestID, rr).empty());
        CHECK(mgr.GetRequestJSON(D, jv.RequestID, &js).empty());
        CHECK(js.find("\"status\":\"completed\",\"retry_count\":1,") != std::string::npos);
        CHECK(js.find("\"processed_at\":\"2023-11-14T22:13:22.5Z\",\"response\":{\"status_code\":201,\"headers\":{\"Server\":\"x\"},\"body\":\"b2s=\","
                      "\"received_at\":\"2023-11-14T22:13:22.5Z\"},\"error\":\"EOF\"}") != std::string::npos);
        Verdict jv2; CHECK(mgr.Decide(D, r, &jv2).empty());
        size_t cnt = 0; CHECK(mgr.GetPendingRequestsJSON(D, &js, &cnt).empty());
        CHECK(cnt == 1 && js.front() == '[' && js.back() == ']' && js.find(jv2.RequestID) != std::string::npos);
        uint64_t gone = 0;
        CHECK(mgr.Expire(clock + 24ull * 3600 * 1000000000ull, 24ull * 3600 * 1000000000ull, &gone).empty());
        CHECK(gone >= 2);                                                    // everything of this run is a day old by then
        CHECK(!mgr.GetRequestJSON(D, jv.RequestID, &js).empty());            // redis: nil
        CHECK(mgr.GetPendingRequestsJSON(D, &js, &cnt).empty() && js == "null" && cnt == 0);
        mgr.SetClock(nullptr);
        // the rest of the driver continues on the logical clock; rows ingested before stay expired, which the counters below ignore
    }

    // concurrent replay: six stopped agents queue 40 requests each, all start, one tick on four workers;
    // every agent's requests are replayed in its arrival order and its completed list holds each id twice, in order (Q7)
    {
        std::vector<std::string> names; std::vector<std::vector<std::string>> want(6);
        char ids[6][AGR_AGENT_ID_BYTES]; uint8_t sts[6]; int32_t slots[6];
        for (int a = 0; a < 6; ++a) {
            names.push_back("agent-17000000000000001" + std::to_string(10 + a));
            memset(ids[a], 0, sizeof ids[a]); strncpy(ids[a], names[a].c_str(), AGR_AGENT_ID_BYTES - 1); sts[a] = AGR_AGENT_STOPPED;
        }
        CHECK(agr_set_agent_states(h, ids, sts, 6, slots) == 0 && slots[5] == slots[0] + 5);
        for (int i = 0; i < 40; ++i) for (int a = 0; a < 6; ++a) {
            HttpRequest r = post; r.Path = "/agent/" + names[a] + "/chat"; r.Body.push_back((uint8_t)i);
            Verdict qv; CHECK(mgr.Decide(names[a], r, &qv).empty() && qv.Code == AGR_V_QUEUED);
            want[a].push_back(qv.RequestID);
        }
        for (int a = 0; a < 6; ++a) sts[a] = AGR_AGENT_RUNNING;
        CHECK(agr_set_agent_states(h, ids, sts, 6, nullptr) == 0);
        std::mutex smu; std::map<std::string, std::vector<std::string>> seen_by;
        ReplayWorker cw(&mgr, [&](const std::string& agent, const Request& req) {
            { std::lock_guard<std::mutex> g(smu); seen_by[agent].push_back(req.ID); }
            HttpRequest rr; rr.Method = req.Method; rr.Path = req.Path; rr.Header = req.Headers; rr.Body = req.Body;
            rr.Header["X-Agentainer-Request-ID"] = req.ID; rr.Header["X-Agentainer-Replay"] = "true";
            Verdict pv; mgr.Decide(agent, rr, &pv);
            if (pv.Code != AGR_V_FORWARD) return pv.HTTPStatus;
            Response r200; r200.StatusCode = 200;
            mgr.StoreResponse(agent, pv.RequestID, r200);
            return 200;
        });
        CHECK(cw.ProcessAgentsConcurrent(4) == 240);
        for (int a = 0; a < 6; ++a) {
            CHECK(seen_by[names[a]] == want[a]);
            std::vector<std::string> twice;
            for (const auto& id : want[a]) { twice.push_back(id); twice.push_back(id); }
            CHECK(list_ids(h, names[a].c_str(), AGR_LIST_COMPLETED) == twice);
            CHECK(list_ids(h, names[a].c_str(), AGR_LIST_PENDING).empty());
        }
    }

    // concurrency: NT threads x 500 requests against a running agent, each completed by its own thread
    const char* C = "agent-1700000000000000003";
    CHECK(agr_set_agent_state(h, C, AGR_AGENT_RUNNING) == 9);
    const int NT = combine ? 64 : 8;
    std::vector<std::thread> ths; std::atomic<int> bad{0};
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < NT; ++t) ths.emplace_back([&, t] {
        for (int i = 0; i < 500; ++i) {
            HttpRequest r = post; r.Path = std::string("/agent/") + C + "/chat";
            Verdict tv; Response r200; r200.StatusCode = 200;
            if (!mgr.Decide(C, r, &tv).empty() || tv.Code != AGR_V_FORWARD) { bad++; continue; }
            if (!mgr.StoreResponse(C, tv.RequestID, r200).empty()) bad++;
        }
        (void)t;
    });
    for (auto& th : ths) th.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CHECK(bad == 0);
    CHECK(list_ids(h, C, AGR_LIST_PENDING).empty());
    agr_stats st; CHECK(agr_stats_get(h, &st) == 0);
    CHECK(st.completions == (uint64_t)(1 + 6 + 1 + 480 + NT * 500) && st.completion_misses == 1);
    printf("host mirror OK (%s ids%s): KAT-A/B/C + %d concurrent single-request Decide+StoreResponse round trips from %d threads, %.0f req/s, %llu K1 launches\n",
           mint ? "engine-minted" : "caller-supplied", combine ? ", flat-combined ingest" : "", NT * 500, NT, NT * 500 / secs,
           (unsigned long long)st.k1_launches);
    agr_destroy(h);
    return 0;
}

int main() {
    int rc = run(false, false);
    if (rc) return rc;
    rc = run(true, false);
    if (rc) return rc;
    return run(true, true);
}
