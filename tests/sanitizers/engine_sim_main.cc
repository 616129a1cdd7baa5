// Drives the REAL scheduler (csrc/engine.cc behind the C ABI) over the fake Model from many threads:
// mixed prompt lengths with shared prefixes (agent preamble + per-task tail), a KV pool too small
// for everything (admission queueing + prefix-cache eviction), cancellations, wait and poll used
// concurrently.  Every 200 response must equal fakemodel::generate(prompt) — computed without any
// cache — and at the end every page must be free or cached.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "acp_host.h"
#include "acp_infer.h"
#include "fake_model.h"
#include "json.h"

using acp::Json;

static std::vector<int> make_prompt(uint64_t seed, int agent, int task) {
  std::vector<int> p = {128000};
  uint64_t h = fakemodel::mix(1000 + (uint64_t)agent);
  const int pre = 96 + 64 * (agent % 3);                       // shared "system prompt + tools" of the agent
  for (int i = 0; i < pre; ++i) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
  h = fakemodel::mix(seed * 7919 + (uint64_t)task);
  const int tail = 5 + (int)(h % 150);
  for (int i = 0; i < tail; ++i) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
  return p;
}

int main() {
  acp_engine* e = nullptr;
  const bool no_cache = getenv("ACP_SIM_NO_CACHE") != nullptr;   // second leg: the same load with the prefix cache off
  const bool replicas = getenv("ACP_SIM_REPLICAS") != nullptr;   // third leg: 4 data-parallel engines behind ONE handle
  const char* cfg = replicas ? "{\"model\": \"sim\", \"replicas\": 4, \"max_batch\": 8, \"kv_pages\": 120, \"max_tokens_per_step\": 512, "
                               "\"max_pages_per_seq\": 16, \"prefix_cache\": true}"
                    : no_cache ? "{\"model\": \"sim\", \"max_batch\": 24, \"kv_pages\": 300, \"max_tokens_per_step\": 512, "
                               "\"max_pages_per_seq\": 16, \"prefix_cache\": false, \"decode_interleave\": 2}"
                             : "{\"model\": \"sim\", \"max_batch\": 24, \"kv_pages\": 300, \"max_tokens_per_step\": 512, "
                               "\"max_pages_per_seq\": 16, \"prefix_cache\": true}";
  if (acp_infer_init(cfg, &e) != 0) { fprintf(stderr, "init failed\n"); return 1; }
  const uint64_t sim_seed = getenv("ACP_SIM_SEED") ? strtoull(getenv("ACP_SIM_SEED"), nullptr, 10) : 17;
  const int kThreads = 24, kPerThread = 60, kMaxTokens = 24;
  std::atomic<int> bad{0}, ok{0}, cancelled{0};
  auto worker = [&](int tid) {
    for (int i = 0; i < kPerThread; ++i) {
      // every 7th request repeats a prompt other threads send too (twins: the publish-time switch to
      // an already cached page), the rest share only their agent's preamble
      const std::vector<int> prompt = (i % 7 == 3) ? make_prompt(sim_seed, i % 5, 424242 + i % 3) : make_prompt(sim_seed, (tid + i) % 5, tid * 1000 + i);
      Json ids = Json::array();
      for (int t : prompt) ids.push(Json(t));
      Json acp = Json::object();
      acp.set("prompt_token_ids", ids);
      Json req = Json::object();
      req.set("model", Json("sim"));
      req.set("max_tokens", Json(kMaxTokens));
      req.set("acp", acp);
      if (i % 2 == 0) req.set("user", Json("task-" + std::to_string((tid + i) % 5)));   // sticky routing key (replica leg)
      const std::string body = req.dump();
      uint64_t ticket = 0;
      if (acp_infer_submit(e, body.c_str(), body.size(), &ticket) != 0) { ++bad; continue; }
      const bool cancel = (tid * 31 + i) % 11 == 0;
      if (cancel) acp_infer_cancel(e, ticket);
      if (i % 2) {                                             // half the waiters poll-spin like the Go poller would
        while (acp_infer_wait(e, ticket, 0) == ACP_ERR_TIMEOUT) {
          uint64_t got[8];
          acp_infer_poll(e, got, 8, 1);
        }
      } else if (acp_infer_wait(e, ticket, -1) != 0) { ++bad; continue; }
      char* out = nullptr;
      size_t len = 0;
      int status = 0;
      if (acp_infer_result(e, ticket, &out, &len, &status) != 0) { ++bad; continue; }
      const std::string resp(out, len);
      acp_infer_free(out);
      if (status == 499 && cancel) { ++cancelled; continue; }
      if (status == 422) {   // empty completion (terminal 4xx): right iff the very first token is a stop token
        const std::vector<int> want = fakemodel::generate(prompt, kMaxTokens);
        if (want.size() == 1 && want[0] == 128009) ++ok; else { fprintf(stderr, "unexpected 422\n"); ++bad; }
        continue;
      }
      Json j;
      std::string err;
      if (status != 200 || !Json::parse(resp, &j, &err)) { fprintf(stderr, "status %d: %s\n", status, resp.c_str()); ++bad; continue; }
      std::vector<int> got;
      for (const Json& t : j.get("acp").get("token_ids").items()) got.push_back((int)t.as_int());
      const std::vector<int> want = fakemodel::generate(prompt, kMaxTokens);
      if (got != want) {
        fprintf(stderr, "MISMATCH thread %d req %d: prompt %zu tokens, got %zu want %zu tokens\n", tid, i, prompt.size(), got.size(), want.size());
        ++bad;
      } else {
        ++ok;
      }
    }
  };
  // phase 0: a cold burst — 48 Tasks of ONE agent (same 288-token preamble, different tails) submitted
  // back to back before anything is cached: the first computes the preamble, the rest must wait for its
  // pages (in-flight dedup) instead of prefilling 47 copies
  {
    std::vector<std::string> bodies;
    std::vector<std::vector<int>> prompts;
    for (int i = 0; i < 48; ++i) {
      std::vector<int> p = {128000};
      uint64_t h = 99;
      for (int k = 0; k < 287; ++k) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
      h = fakemodel::mix(5000 + (uint64_t)i);
      for (int k = 0; k < 20 + i % 30; ++k) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
      Json ids = Json::array();
      for (int t : p) ids.push(Json(t));
      Json acp = Json::object();
      acp.set("prompt_token_ids", ids);
      Json req = Json::object();
      req.set("model", Json("sim"));
      req.set("max_tokens", Json(kMaxTokens));
      req.set("acp", acp);
      bodies.push_back(req.dump());
      prompts.push_back(p);
    }
    {  // the blocker: an unrelated 420-token prompt whose prefill step keeps the scheduler busy (fake_model.cc)
      std::vector<int> p = {128000};
      uint64_t h = 7;
      for (int k = 0; k < 419; ++k) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
      Json ids = Json::array();
      for (int t : p) ids.push(Json(t));
      Json acp = Json::object();
      acp.set("prompt_token_ids", ids);
      Json req = Json::object();
      req.set("model", Json("sim"));
      req.set("max_tokens", Json(4));
      req.set("acp", acp);
      bodies.push_back(req.dump());
      prompts.push_back(p);
      std::swap(bodies.front(), bodies.back());
      std::swap(prompts.front(), prompts.back());
    }
    std::vector<uint64_t> tickets(bodies.size());
    acp_infer_submit(e, bodies[0].c_str(), bodies[0].size(), &tickets[0]);
    std::this_thread::sleep_for(std::chrono::milliseconds(5));      // the blocker's prefill step is now running
    for (size_t i = 1; i < bodies.size(); ++i) acp_infer_submit(e, bodies[i].c_str(), bodies[i].size(), &tickets[i]);
    for (size_t i = 0; i < bodies.size(); ++i) {
      acp_infer_wait(e, tickets[i], -1);
      char* out = nullptr; size_t len = 0; int status = 0;
      acp_infer_result(e, tickets[i], &out, &len, &status);
      const std::string resp(out ? out : "", len);
      acp_infer_free(out);
      const std::vector<int> want = fakemodel::generate(prompts[i], i == 0 ? 4 : kMaxTokens);
      std::vector<int> got;
      Json j; std::string err;
      if (status == 200 && Json::parse(resp, &j, &err))
        for (const Json& t : j.get("acp").get("token_ids").items()) got.push_back((int)t.as_int());
      const bool empty_ok = status == 422 && want.size() == 1 && want[0] == 128009;
      if (!(empty_ok || (status == 200 && got == want))) { fprintf(stderr, "cold burst: request %zu wrong (status %d)\n", i, status); ++bad; }
    }
    char* sj0 = nullptr;
    acp_infer_stats(e, &sj0);
    Json s0; std::string err0;
    Json::parse(std::string(sj0), &s0, &err0);
    acp_infer_free(sj0);
    const long long pf = s0.get("prefill_tokens").as_int(), df = s0.get("prefix_deferrals").as_int();
    printf("cold burst: prefill_tokens=%lld (48 cold copies would be %d) deferrals=%lld hits=%lld\n", pf, 48 * 288, df,
           (long long)s0.get("prefix_hits").as_int());
    if (!no_cache && !replicas && (df < 1 || pf > 420 + 48 * 288 / 3)) { fprintf(stderr, "in-flight dedup did not engage\n"); ++bad; }
  }
  std::vector<std::thread> th;
  for (int t = 0; t < kThreads; ++t) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
  char* sj = nullptr;
  acp_infer_stats(e, &sj);
  Json s;
  std::string err;
  Json::parse(std::string(sj), &s, &err);
  acp_infer_free(sj);
  const long long free_pages = s.get("kv_pages_free").as_int(), cached = s.get("prefix_cache_pages").as_int(),
                  total = s.get("kv_pages_total").as_int(), hits = s.get("prefix_hits").as_int();
  printf("engine-sim: ok=%d cancelled=%d bad=%d prefix_hits=%lld reused=%lld deferrals=%lld prefill_tokens=%lld pages free=%lld cached=%lld total=%lld running=%lld\n",
         ok.load(), cancelled.load(), bad.load(), hits, (long long)s.get("prefix_tokens_reused").as_int(),
         (long long)s.get("prefix_deferrals").as_int(), (long long)s.get("prefill_tokens").as_int(), free_pages, cached, total,
         (long long)s.get("running").as_int());
  int rc = 0;
  if (bad != 0) rc = 1;
  if (ok + cancelled != kThreads * kPerThread) rc = 1;
  if (free_pages + cached != total) { fprintf(stderr, "page leak: %lld + %lld != %lld\n", free_pages, cached, total); rc = 1; }
  if (no_cache && (hits != 0 || cached != 0)) { fprintf(stderr, "cache is off but was used\n"); rc = 1; }
  if (replicas && (s.get("replica_count").as_int() != 4 || s.get("replicas").size() != 4)) { fprintf(stderr, "replica stats missing\n"); rc = 1; }
  if (replicas)
    for (const Json& r : s.get("replicas").items())
      if (r.get("requests_done").as_int() + r.get("requests_failed").as_int() < 100) { fprintf(stderr, "a replica got almost no work\n"); rc = 1; }
  if (!no_cache && hits < kThreads * kPerThread / 2) { fprintf(stderr, "too few prefix hits: %lld\n", hits); rc = 1; }
  acp_infer_shutdown(e);
  return rc;
}
