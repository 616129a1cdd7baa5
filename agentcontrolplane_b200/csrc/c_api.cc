// c_api.cc — extern "C" surface declared in include/acp_infer.h.  Nothing throws across it.
#include "acp_infer.h"
#include <new>
#include <stdlib.h>
#include <string.h>
#include "engine.h"

struct acp_engine {
  acp::Engine engine;
};

static char* dup_string(const std::string& s, size_t* len) {
  char* p = (char*)malloc(s.size() + 1);
  if (!p) return nullptr;
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  if (len) *len = s.size();
  return p;
}

extern "C" {

const char* acp_infer_version(void) { return "acp_infer 0.1.0 sm_100a"; }

int acp_infer_init(const char* config_json, acp_engine** out) {
  if (!out) return ACP_ERR_INVALID;
  *out = nullptr;
  try {
    acp_engine* e = new (std::nothrow) acp_engine();
    if (!e) return ACP_ERR_NOMEM;
    int rc = e->engine.init(config_json);
    if (rc != 0) { delete e; return rc; }
    *out = e;
    return ACP_OK;
  } catch (const std::bad_alloc&) {
    return ACP_ERR_NOMEM;
  } catch (...) {
    return ACP_ERR_INVALID;
  }
}

int acp_infer_submit(acp_engine* e, const char* chat_request_json, size_t len, uint64_t* ticket) {
  if (!e || !chat_request_json || !ticket) return ACP_ERR_INVALID;
  try { return e->engine.submit(chat_request_json, len, ticket); } catch (...) { return ACP_ERR_NOMEM; }
}

int acp_infer_wait(acp_engine* e, uint64_t ticket, int timeout_ms) {
  if (!e) return ACP_ERR_INVALID;
  try { return e->engine.wait(ticket, timeout_ms); } catch (...) { return ACP_ERR_INVALID; }
}

int acp_infer_poll(acp_engine* e, uint64_t* tickets, int max, int timeout_ms) {
  if (!e) return ACP_ERR_INVALID;
  try { return e->engine.poll(tickets, max, timeout_ms); } catch (...) { return ACP_ERR_INVALID; }
}

int acp_infer_result(acp_engine* e, uint64_t ticket, char** chat_response_json, size_t* len,
                     int* http_like_status) {
  if (!e || !chat_response_json || !http_like_status) return ACP_ERR_INVALID;
  try {
    std::string body;
    int rc = e->engine.result(ticket, &body, http_like_status);
    if (rc != 0) return rc;
    *chat_response_json = dup_string(body, len);
    return *chat_response_json ? ACP_OK : ACP_ERR_NOMEM;
  } catch (...) {
    return ACP_ERR_NOMEM;
  }
}

int acp_infer_result_logits(acp_engine* e, uint64_t ticket, float* out, int max_positions) {
  if (!e) return ACP_ERR_INVALID;
  try { return e->engine.result_logits(ticket, out, max_positions); } catch (...) { return ACP_ERR_INVALID; }
}

void acp_infer_cancel(acp_engine* e, uint64_t ticket) {
  if (!e) return;
  try { e->engine.cancel(ticket); } catch (...) {}
}

int acp_infer_stats(acp_engine* e, char** json) {
  if (!e || !json) return ACP_ERR_INVALID;
  try {
    *json = dup_string(e->engine.stats_json(), nullptr);
    return *json ? ACP_OK : ACP_ERR_NOMEM;
  } catch (...) {
    return ACP_ERR_NOMEM;
  }
}

void acp_infer_stats_reset(acp_engine* e) {
  if (e) e->engine.stats_reset();
}

void acp_infer_free(void* p) { free(p); }

void acp_infer_shutdown(acp_engine* e) {
  if (!e) return;
  try { e->engine.shutdown(); } catch (...) {}
  delete e;
}

}  // extern "C"
