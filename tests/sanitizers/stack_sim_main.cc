// The whole host stack in one ThreadSanitizer binary: reconcile workers (csrc/host/hostsim.cc, task.cc)
// -> LocalClient (llmclient.cc) -> C ABI (c_api.cc) -> the REAL scheduler (engine.cc) -> the fake
// Model (fake_model.cc).  Runs BASELINE config 3's shape on a CPU: Tasks with two tool schemas, a
// scripted tool call on the first LLM step (force_tokens through the real engine), ToolCall CRs,
// fold-back, second LLM step served from the shared prefix cache.
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "acp_host.h"
#include "acp_infer.h"
#include "json.h"

using acp::Json;

static Json stats(acp_engine* e) {
  char* sj = nullptr;
  acp_infer_stats(e, &sj);
  Json s;
  std::string err;
  Json::parse(std::string(sj), &s, &err);
  acp_infer_free(sj);
  return s;
}

int main() {
  acp_engine* e = nullptr;
  // ACP_SIM_REPLICAS: the same run over 4 data-parallel engines behind one handle — the second LLM step of a
  // Task must find its first step's pages, i.e. the router must be sticky without any Task id in the request
  const bool replicas = getenv("ACP_SIM_REPLICAS") != nullptr;
  if (acp_infer_init(replicas ? "{\"model\": \"sim\", \"replicas\": 4, \"max_batch\": 32, \"kv_pages\": 1500, \"max_tokens_per_step\": 2048, "
                                "\"max_pages_per_seq\": 40, \"prefix_cache\": true}"
                              : "{\"model\": \"sim\", \"max_batch\": 32, \"kv_pages\": 1500, \"max_tokens_per_step\": 2048, "
                                "\"max_pages_per_seq\": 40, \"prefix_cache\": true}", &e) != 0) return 1;
  int bad = 0;
  long long hits_before = 0, reused_before = 0;
  for (int round = 0; round < 3; ++round) {
    const std::string cfg = "{\"tasks\": 48, \"workers\": 48, \"provider\": \"local\", \"model\": \"sim\", \"max_tokens\": 32, "
                            "\"prompt_tokens\": 512, \"tools\": 2, \"tool_loop\": true, \"seed\": " + std::to_string(round + 1) + "}";
    char* out = nullptr;
    if (acp_hostsim_run(e, cfg.c_str(), &out) != 0 || !out) { fprintf(stderr, "hostsim failed\n"); return 1; }
    Json r;
    std::string err;
    Json::parse(std::string(out), &r, &err);
    acp_infer_free(out);
    const Json s = stats(e);
    const long long hits = s.get("prefix_hits").as_int();
    // the fake model may end a turn with an empty completion (first token = EOT, 1 in 23): that Task is
    // terminally Failed (422 -> LLMRequestError), exactly like the reference's 4xx arm
    long long final_answers = r.get("final_phases").get("FinalAnswer").as_int(0), failed = r.get("final_phases").get("Failed").as_int(0);
    printf("round %d: reconciles=%lld FinalAnswer=%lld Failed=%lld prefix_hits+=%lld deferrals=%lld cache_pages=%lld\n", round,
           (long long)r.get("reconciles").as_int(), final_answers, failed, hits - hits_before,
           (long long)s.get("prefix_deferrals").as_int(), (long long)s.get("prefix_cache_pages").as_int());
    if (final_answers + failed != 48) { fprintf(stderr, "tasks lost\n"); ++bad; }
    if (r.get("reconciles").as_int() < 48 + final_answers) { fprintf(stderr, "a Task reached FinalAnswer without two LLM steps\n"); ++bad; }
    // second turns reuse >= 21 pages of their own first turn: only possible on the SAME replica
    if (s.get("prefix_tokens_reused").as_int() - reused_before < (r.get("reconciles").as_int() - 48) * 640) { fprintf(stderr, "second turns did not find their K/V\n"); ++bad; }
    reused_before = s.get("prefix_tokens_reused").as_int();
    if (round > 0 && hits - hits_before < r.get("reconciles").as_int() - 2) { fprintf(stderr, "warm rounds must hit the shared prefix\n"); ++bad; }
    if (s.get("kv_pages_free").as_int() + s.get("prefix_cache_pages").as_int() != s.get("kv_pages_total").as_int()) { fprintf(stderr, "page leak\n"); ++bad; }
    hits_before = hits;
  }
  // BASELINE config 4's shape: open-loop Poisson arrivals and sub-agent delegation chains of depth 2 —
  // root Task: delegate tool call -> child Task (sub-agent-1): delegate -> grandchild (sub-agent-2): answer ->
  // results fold back up: 5 LLM steps and 2 ToolCall CRs per root Task, every child created like
  // executeDelegateToAgent does (toolcall/executor.go:176-242)
  {
    const std::string cfg = "{\"tasks\": 40, \"workers\": 40, \"provider\": \"local\", \"model\": \"sim\", \"max_tokens\": 24, "
                            "\"prompt_tokens\": 256, \"seed\": 9, \"arrival_rate\": 400.0, \"delegation_depth\": 2}";
    char* out = nullptr;
    if (acp_hostsim_run(e, cfg.c_str(), &out) != 0 || !out) { fprintf(stderr, "hostsim (delegation) failed\n"); return 1; }
    Json r;
    std::string err;
    Json::parse(std::string(out), &r, &err);
    acp_infer_free(out);
    const long long fa = r.get("final_phases").get("FinalAnswer").as_int(0), failed = r.get("final_phases").get("Failed").as_int(0);
    printf("delegation: reconciles=%lld FinalAnswer=%lld Failed=%lld task_ms_p50=%.2f wall=%.3f\n", (long long)r.get("reconciles").as_int(), fa, failed,
           r.get("task_ms_p50").as_double(), r.get("wall_s").as_double());
    if (fa + failed != 40) { fprintf(stderr, "root tasks lost\n"); ++bad; }
    // a chain that completes takes 5 LLM steps; an empty completion (1 in 23 per step) ends its Task as Failed earlier
    if (r.get("reconciles").as_int() < 3 * fa || r.get("reconciles").as_int() > 5 * 40) { fprintf(stderr, "unexpected LLM step count\n"); ++bad; }
    if (fa < 20) { fprintf(stderr, "too few delegation chains completed\n"); ++bad; }
    if (r.get("wall_s").as_double() < 40 / 400.0 * 0.5) { fprintf(stderr, "arrivals were not spread out\n"); ++bad; }
  }
  acp_infer_shutdown(e);
  return bad ? 1 : 0;
}
