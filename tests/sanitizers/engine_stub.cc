// Host-only stand-in for the CUDA engine behind include/acp_infer.h, for the ThreadSanitizer build
// of the host side (tests/test_sanitizers_cpu.py).  It answers every submitted request from a worker
// thread after a short delay, so LocalClient's submit / wait / result hand-off, the Task state
// machine, the object store and the HTTP client/server are exercised under real concurrency with
// no GPU.  TEST INFRASTRUCTURE: never linked into libacp_infer.so.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

#include "acp_infer.h"
#include "json.h"
#include "model.h"

struct acp_engine {
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<std::pair<uint64_t, std::string>> queue;
  std::unordered_map<uint64_t, std::string> done;
  std::deque<uint64_t> unreported;
  uint64_t next = 1;
  bool stop = false;
  std::thread worker;
};

static std::string answer(uint64_t ticket, const std::string& req) {
  using acp::Json;
  Json msg = Json::object();
  msg.set("role", Json("assistant"));
  if (req.find("\"force_tokens\"") != std::string::npos) {  // scripted tool-call step of the tool loop
    Json fn = Json::object();
    fn.set("name", Json("fetch__tool_0"));
    fn.set("arguments", Json("{\"url\": \"https://api.example.com/data\"}"));
    Json tc = Json::object();
    tc.set("id", Json("call_" + std::to_string(ticket) + "_0"));
    tc.set("type", Json("function"));
    tc.set("function", fn);
    Json arr = Json::array();
    arr.push(tc);
    msg.set("content", Json(""));
    msg.set("tool_calls", arr);
  } else {
    msg.set("content", Json("answer " + std::to_string(ticket)));
  }
  Json choice = Json::object();
  choice.set("index", Json(0));
  choice.set("message", msg);
  choice.set("finish_reason", Json("stop"));
  Json choices = Json::array();
  choices.push(choice);
  Json root = Json::object();
  root.set("id", Json("stub-" + std::to_string(ticket)));
  root.set("object", Json("chat.completion"));
  root.set("choices", choices);
  return root.dump();
}

extern "C" {

int acp_infer_init(const char*, acp_engine** out) {
  if (!out) return ACP_ERR_INVALID;
  acp_engine* e = new acp_engine();
  e->worker = std::thread([e] {
    std::unique_lock<std::mutex> lk(e->mu);
    while (true) {
      e->cv_work.wait(lk, [&] { return e->stop || !e->queue.empty(); });
      if (e->stop) return;
      auto batch = std::move(e->queue);
      e->queue.clear();
      lk.unlock();
      std::this_thread::sleep_for(std::chrono::microseconds(200));  // "one engine step"
      lk.lock();
      for (auto& kv : batch) { e->done[kv.first] = answer(kv.first, kv.second); e->unreported.push_back(kv.first); }
      e->cv_done.notify_all();
    }
  });
  *out = e;
  return ACP_OK;
}

int acp_infer_submit(acp_engine* e, const char* json, size_t len, uint64_t* ticket) {
  if (!e || !json || !ticket) return ACP_ERR_INVALID;
  std::lock_guard<std::mutex> lk(e->mu);
  *ticket = e->next++;
  e->queue.emplace_back(*ticket, std::string(json, len));
  e->cv_work.notify_one();
  return ACP_OK;
}

int acp_infer_wait(acp_engine* e, uint64_t ticket, int timeout_ms) {
  std::unique_lock<std::mutex> lk(e->mu);
  auto ready = [&] { return e->done.count(ticket) > 0; };
  if (timeout_ms < 0) { e->cv_done.wait(lk, ready); return ACP_OK; }
  return e->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready) ? ACP_OK : ACP_ERR_TIMEOUT;
}

int acp_infer_poll(acp_engine* e, uint64_t* tickets, int max, int timeout_ms) {
  std::unique_lock<std::mutex> lk(e->mu);
  if (e->unreported.empty() && timeout_ms != 0)
    e->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms < 0 ? 1000 : timeout_ms), [&] { return !e->unreported.empty(); });
  int n = 0;
  while (n < max && !e->unreported.empty()) { tickets[n++] = e->unreported.front(); e->unreported.pop_front(); }
  return n;
}

int acp_infer_result(acp_engine* e, uint64_t ticket, char** body, size_t* len, int* status) {
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->done.find(ticket);
  if (it == e->done.end()) return ACP_ERR_NOT_FOUND;
  *body = (char*)malloc(it->second.size() + 1);
  memcpy(*body, it->second.c_str(), it->second.size() + 1);
  if (len) *len = it->second.size();
  if (status) *status = 200;
  e->done.erase(it);
  return ACP_OK;
}

int acp_infer_result_logits(acp_engine*, uint64_t, float*, int) { return 0; }
void acp_infer_cancel(acp_engine*, uint64_t) {}
int acp_infer_stats(acp_engine*, char** json) { *json = strdup("{}"); return ACP_OK; }
void acp_infer_stats_reset(acp_engine*) {}
void acp_infer_free(void* p) { free(p); }
void acp_infer_shutdown(acp_engine* e) {
  if (!e) return;
  { std::lock_guard<std::mutex> lk(e->mu); e->stop = true; }
  e->cv_work.notify_all();
  e->worker.join();
  delete e;
}
const char* acp_infer_version(void) { return "acp_infer host-only stub"; }

}  // extern "C"

