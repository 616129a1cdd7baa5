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
  const char* cfg = "{\"model\": \"sim\", \"max_batch\": 24, \"kv_pages\": 300, \"max_tokens_per_step\": 512, "
                    "\"max_pages_per_seq\": 16, \"prefix_cache\": true}";
  if (acp_infer_init(cfg, &e) != 0) { fprintf(stderr, "init failed\n"); return 1; }
  const int kThreads = 24, kPerThread = 60, kMaxTokens = 24;
  std::atomic<int> bad{0}, ok{0}, cancelled{0};
  auto worker = [&](int tid) {
    for (int i = 0; i < kPerThread; ++i) {
      const std::vector<int> prompt = make_prompt(17, (tid + i) % 5, tid * 1000 + i);
      Json ids = Json::array();
      for (int t : prompt) ids.push(Json(t));
      Json acp = Json::object();
      acp.set("prompt_token_ids", ids);
      Json req = Json::object();
      req.set("model", Json("sim"));
      req.set("max_tokens", Json(kMaxTokens));
      req.set("acp", acp);
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
  printf("engine-sim: ok=%d cancelled=%d bad=%d prefix_hits=%lld reused=%lld pages free=%lld cached=%lld total=%lld running=%lld\n",
         ok.load(), cancelled.load(), bad.load(), hits, (long long)s.get("prefix_tokens_reused").as_int(), free_pages, cached, total,
         (long long)s.get("running").as_int());
  int rc = 0;
  if (bad != 0) rc = 1;
  if (ok + cancelled != kThreads * kPerThread) rc = 1;
  if (free_pages + cached != total) { fprintf(stderr, "page leak: %lld + %lld != %lld\n", free_pages, cached, total); rc = 1; }
  if (hits < kThreads * kPerThread / 2) { fprintf(stderr, "too few prefix hits: %lld\n", hits); rc = 1; }
  acp_infer_shutdown(e);
  return rc;
}
