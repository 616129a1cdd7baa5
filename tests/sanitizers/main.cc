// ThreadSanitizer driver for the host side: the reference's reconcile loop restated in C++
// (csrc/host/*) under many concurrent reconcile workers, (1) over HTTP/1.1 keep-alive to the loopback
// stub completion server, (2) through LocalClient -> C ABI -> (stub) engine, (3) the tool loop
// (ToolCall creation + fold-back).  Exit code 0 and no "ThreadSanitizer" report = clean.
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "acp_host.h"
#include "acp_infer.h"

static int run(acp_engine* e, const std::string& cfg, const char* what, const char* expect) {
  char* out = nullptr;
  const int rc = acp_hostsim_run(e, cfg.c_str(), &out);
  if (rc != 0 || !out) { fprintf(stderr, "%s: acp_hostsim_run rc=%d\n", what, rc); return 1; }
  const std::string s(out);
  acp_infer_free(out);
  printf("%s: %s\n", what, s.c_str());
  if (s.find(expect) == std::string::npos) { fprintf(stderr, "%s: expected %s\n", what, expect); return 1; }
  return 0;
}

int main() {
  int port = 0;
  const int srv = acp_host_stub_server_start(nullptr, &port);
  if (srv < 0) { fprintf(stderr, "stub server failed\n"); return 1; }
  int bad = 0;
  bad += run(nullptr, "{\"tasks\": 600, \"workers\": 16, \"provider\": \"openai\", \"model\": \"gpt-4o\", \"baseURL\": \"http://127.0.0.1:" +
                          std::to_string(port) + "/v1\", \"prompt_tokens\": 256}",
             "openai/http", "\"FinalAnswer\":600");
  acp_host_stub_server_stop(srv);
  acp_engine* e = nullptr;
  if (acp_infer_init("{}", &e) != 0) return 1;
  bad += run(e, "{\"tasks\": 512, \"workers\": 64, \"provider\": \"local\", \"model\": \"tiny\", \"prompt_tokens\": 256}",
             "local/abi", "\"FinalAnswer\":512");
  bad += run(e, "{\"tasks\": 128, \"workers\": 32, \"provider\": \"local\", \"model\": \"tiny\", \"prompt_tokens\": 0, \"tools\": 2, "
                "\"tool_loop\": true}",
             "local/tool-loop", "\"reconciles\":256");
  acp_infer_shutdown(e);
  return bad ? 1 : 0;
}
