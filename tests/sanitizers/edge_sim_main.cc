// Edge cases of the real scheduler over the fake Model (single thread of submits, deterministic):
// prompts of 1, 2, 31, 32, 33, 63, 64, 65 ... tokens (page boundaries), prefill chunked into 48-row
// steps (a prompt spans many steps, pages are published chunk by chunk), twins and prefix-of-prefix
// prompts back to back, max_tokens 1, a prompt that does not fit the context limit (400), and a pool
// so small that sequences queue.  Every 200 must equal the cache-free reference.
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "acp_infer.h"
#include "fake_model.h"
#include "json.h"

using acp::Json;

static std::vector<int> prompt_of(int n, uint64_t seed) {
  std::vector<int> p = {128000};
  uint64_t h = seed;
  for (int i = 1; i < n; ++i) { h = fakemodel::mix(h); p.push_back(32 + (int)(h % 90)); }
  return p;
}

int main() {
  acp_engine* e = nullptr;
  if (acp_infer_init("{\"model\": \"sim\", \"max_batch\": 6, \"kv_pages\": 64, \"max_tokens_per_step\": 48, "
                     "\"max_pages_per_seq\": 12, \"prefix_cache\": true, \"decode_interleave\": 1}", &e) != 0) return 1;
  struct Case { std::vector<int> prompt; int max_tokens; int want_status; };
  std::vector<Case> cases;
  const int lens[] = {1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 128, 129, 200, 255, 256, 257, 300};
  for (int n : lens) cases.push_back({prompt_of(n, 5), 9, 200});          // every one is a prefix of the next
  for (int n : lens) cases.push_back({prompt_of(n, 5), 1, 200});          // again, warm, one token each
  for (int i = 0; i < 6; ++i) cases.push_back({prompt_of(160, 77), 12, 200});   // sextuplets submitted together
  cases.push_back({prompt_of(380, 9), 24, 400});                          // 380 + 24 > 12 pages * 32
  cases.push_back({prompt_of(360, 9), 24, 200});                          // exactly fits
  std::vector<uint64_t> tickets(cases.size());
  for (size_t i = 0; i < cases.size(); ++i) {
    Json ids = Json::array();
    for (int t : cases[i].prompt) ids.push(Json(t));
    Json acp = Json::object();
    acp.set("prompt_token_ids", ids);
    Json req = Json::object();
    req.set("model", Json("sim"));
    req.set("max_tokens", Json(cases[i].max_tokens));
    req.set("acp", acp);
    const std::string body = req.dump();
    if (acp_infer_submit(e, body.c_str(), body.size(), &tickets[i]) != 0) return 1;
  }
  int bad = 0;
  for (size_t i = 0; i < cases.size(); ++i) {
    if (acp_infer_wait(e, tickets[i], 60000) != 0) { fprintf(stderr, "case %zu timed out\n", i); return 1; }
    char* out = nullptr; size_t len = 0; int status = 0;
    acp_infer_result(e, tickets[i], &out, &len, &status);
    const std::string resp(out ? out : "", len);
    acp_infer_free(out);
    const std::vector<int> want = fakemodel::generate(cases[i].prompt, cases[i].max_tokens);
    const bool empty = want.size() == 1 && want[0] == 128009;
    const int expect = cases[i].want_status == 200 && empty ? 422 : cases[i].want_status;
    if (status != expect) { fprintf(stderr, "case %zu (prompt %zu): status %d, expected %d\n", i, cases[i].prompt.size(), status, expect); ++bad; continue; }
    if (status != 200) continue;
    Json j; std::string err;
    std::vector<int> got;
    if (Json::parse(resp, &j, &err)) for (const Json& t : j.get("acp").get("token_ids").items()) got.push_back((int)t.as_int());
    if (got != want) { fprintf(stderr, "case %zu (prompt %zu tokens): wrong tokens\n", i, cases[i].prompt.size()); ++bad; }
  }
  char* sj = nullptr;
  acp_infer_stats(e, &sj);
  Json s; std::string err;
  Json::parse(std::string(sj), &s, &err);
  acp_infer_free(sj);
  printf("edge-sim: cases=%zu bad=%d prefill_steps=%lld prefix_hits=%lld deferrals=%lld pages free=%lld cached=%lld total=%lld\n", cases.size(), bad,
         (long long)s.get("prefill_steps").as_int(), (long long)s.get("prefix_hits").as_int(), (long long)s.get("prefix_deferrals").as_int(),
         (long long)s.get("kv_pages_free").as_int(), (long long)s.get("prefix_cache_pages").as_int(), (long long)s.get("kv_pages_total").as_int());
  if (s.get("kv_pages_free").as_int() + s.get("prefix_cache_pages").as_int() != s.get("kv_pages_total").as_int()) { fprintf(stderr, "page leak\n"); ++bad; }
  acp_infer_shutdown(e);
  // request_timeout_ms: a one-sequence engine is held by a long prefill step (420 rows = the fake model's 40 ms
  // blocker) while three more requests wait: they must end with a transient 504, the blocker itself with 200
  {
    acp_engine* t = nullptr;
    if (acp_infer_init("{\"model\": \"sim\", \"max_batch\": 1, \"kv_pages\": 64, \"max_tokens_per_step\": 512, "
                       "\"max_pages_per_seq\": 16, \"prefix_cache\": false, \"request_timeout_ms\": 15}", &t) != 0) return 1;
    std::vector<uint64_t> tk(4);
    for (int i = 0; i < 4; ++i) {
      Json ids = Json::array();
      for (int tok : prompt_of(i == 0 ? 420 : 40, 100 + (uint64_t)i)) ids.push(Json(tok));
      Json acp = Json::object();
      acp.set("prompt_token_ids", ids);
      Json req = Json::object();
      req.set("model", Json("sim"));
      req.set("max_tokens", Json(i == 0 ? 1 : 8));
      req.set("acp", acp);
      const std::string body = req.dump();
      acp_infer_submit(t, body.c_str(), body.size(), &tk[(size_t)i]);
    }
    int n504 = 0, n200 = 0;
    for (int i = 0; i < 4; ++i) {
      acp_infer_wait(t, tk[(size_t)i], 60000);
      char* out = nullptr; size_t len = 0; int status = 0;
      acp_infer_result(t, tk[(size_t)i], &out, &len, &status);
      acp_infer_free(out);
      if (status == 504) ++n504;
      if (status == 200 || status == 422) ++n200;
    }
    printf("timeout-sim: 504=%d completed=%d\n", n504, n200);
    if (n504 != 3 || n200 != 1) { fprintf(stderr, "request_timeout_ms did not fire as expected\n"); ++bad; }
    acp_infer_shutdown(t);
  }
  return bad ? 1 : 0;
}
