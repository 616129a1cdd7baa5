// host/llmclient.cc — see llmclient.h.
#include "llmclient.h"
#include <dlfcn.h>
#include <mutex>
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

namespace acp {
namespace llmclient {

std::string Error::Error_() const {
  if (is_request_error)  // LLMRequestError.Error(), llm_client.go:24-26
    return "LLM request failed with status " + std::to_string(StatusCode) + ": " + Message;
  return Message;
}

Json message_to_crd_json(const Message& m) {
  Json j = Json::object();
  j.set("role", Json(m.Role));
  j.set("content", Json(m.Content));
  if (!m.ToolCalls.empty()) {
    Json tcs = Json::array();
    for (const auto& tc : m.ToolCalls) {
      Json fn = Json::object();
      fn.set("name", Json(tc.Function.Name));
      fn.set("arguments", Json(tc.Function.Arguments));
      Json o = Json::object();
      o.set("id", Json(tc.ID));
      o.set("function", fn);
      o.set("type", Json(tc.Type));
      tcs.push(o);
    }
    j.set("toolCalls", tcs);
  }
  if (!m.ToolCallID.empty()) j.set("toolCallId", Json(m.ToolCallID));
  if (!m.Name.empty()) j.set("name", Json(m.Name));
  return j;
}

bool message_from_crd_json(const Json& j, Message* m) {
  if (!j.is_object()) return false;
  m->Role = j.get("role").as_string();
  m->Content = j.get("content").as_string();
  m->ToolCallID = j.get("toolCallId").as_string();
  m->Name = j.get("name").as_string();
  m->ToolCalls.clear();
  for (const Json& tc : j.get("toolCalls").items()) {
    MessageToolCall t;
    t.ID = tc.get("id").as_string();
    t.Type = tc.get("type").as_string();
    t.Function.Name = tc.get("function").get("name").as_string();
    t.Function.Arguments = tc.get("function").get("arguments").as_string();
    m->ToolCalls.push_back(std::move(t));
  }
  return true;
}

// convertToLangchainMessages (langchaingo_client.go:118-185) + langchaingo's openai wire mapping
static Json messages_to_wire(const std::vector<Message>& messages) {
  Json arr = Json::array();
  for (const Message& m : messages) {
    std::string role = m.Role;
    if (role != "system" && role != "user" && role != "assistant" && role != "tool") role = "user";  // :136-137
    Json o = Json::object();
    o.set("role", Json(role));
    if (role == "tool" && !m.ToolCallID.empty()) {  // :164-171: only the ToolCallResponse part
      o.set("content", Json(m.Content));
      o.set("tool_call_id", Json(m.ToolCallID));
    } else {
      o.set("content", Json(m.Content));
      if (!m.ToolCalls.empty()) {
        Json tcs = Json::array();
        for (const auto& tc : m.ToolCalls) {
          Json fn = Json::object();
          fn.set("name", Json(tc.Function.Name));
          fn.set("arguments", Json(tc.Function.Arguments));
          Json t = Json::object();
          t.set("id", Json(tc.ID));
          t.set("type", Json(tc.Type));
          t.set("function", fn);
          tcs.push(t);
        }
        o.set("tool_calls", tcs);
      }
      if (!m.ToolCallID.empty()) o.set("tool_call_id", Json(m.ToolCallID));
    }
    arr.push(o);
  }
  return arr;
}

// convertToLangchainTools (langchaingo_client.go:188-203)
static Json tools_to_wire(const std::vector<Tool>& tools) {
  Json arr = Json::array();
  for (const Tool& t : tools) {
    Json fn = Json::object();
    fn.set("name", Json(t.Function.Name));
    fn.set("description", Json(t.Function.Description));
    fn.set("parameters", t.Function.Parameters.is_null() ? Json::object() : t.Function.Parameters);
    Json o = Json::object();
    o.set("type", Json(t.Type));
    o.set("function", fn);
    arr.push(o);
  }
  return arr;
}

BaseConfig base_config_from_json(const Json& p) {
  BaseConfig bc;
  bc.Model = p.get("model").as_string();
  bc.BaseURL = p.get("baseUrl").as_string();
  if (bc.BaseURL.empty()) bc.BaseURL = p.get("baseURL").as_string();   // spelling used by this repo's test hooks
  bc.Temperature = p.get("temperature").as_string();
  bc.TopP = p.get("topP").as_string();
  bc.MaxTokens = (int)p.get("maxTokens").as_int(0);
  bc.TopK = (int)p.get("topK").as_int(0);
  return bc;
}

std::string build_chat_request_json(const std::string& model, const std::vector<Message>& messages,
                                    const std::vector<Tool>& tools, int max_tokens,
                                    const Json* acp_ext, const BaseConfig* sampling) {
  Json body = Json::object();
  body.set("model", Json(model));
  body.set("messages", messages_to_wire(messages));
  // ChatRequest.Temperature has no omitempty and ACP's SendRequest sets no option: the reference
  // sends 0 (greedy).  The local provider forwards LLM.spec.parameters when they are set.
  double temperature = 0.0;
  if (sampling && !sampling->Temperature.empty()) temperature = atof(sampling->Temperature.c_str());
  if (temperature == 0.0) body.set("temperature", Json(0));
  else body.set("temperature", Json(temperature));
  if (sampling && !sampling->TopP.empty()) body.set("top_p", Json(atof(sampling->TopP.c_str())));
  if (sampling && sampling->TopK > 0) body.set("top_k", Json(sampling->TopK));
  if (max_tokens > 0) body.set("max_tokens", Json(max_tokens));
  if (!tools.empty()) body.set("tools", tools_to_wire(tools));  // :93-99 only when present
  if (acp_ext && acp_ext->is_object()) body.set("acp", *acp_ext);
  return body.dump();
}

bool convert_from_response_json(const std::string& body, Message* out, std::string* err) {
  Json root;
  if (!Json::parse(body, &root, err)) return false;
  Message msg;
  msg.Role = "assistant";
  const auto& choices = root.get("choices").items();
  bool has_content = false;
  std::string content;
  for (const Json& ch : choices) {
    const Json& m = ch.get("message");
    const std::string c = m.get("content").as_string();
    if (!has_content && !c.empty()) { content = c; has_content = true; }
    for (const Json& tc : m.get("tool_calls").items()) {
      MessageToolCall t;
      t.ID = tc.get("id").as_string();
      t.Type = tc.get("type").as_string();
      t.Function.Name = tc.get("function").get("name").as_string();
      t.Function.Arguments = tc.get("function").get("arguments").as_string();
      msg.ToolCalls.push_back(std::move(t));
    }
  }
  if (msg.ToolCalls.empty() && has_content) msg.Content = content;  // tool calls win, clear content
  *out = std::move(msg);
  return true;
}

// ---------------------------------------------------------------------------------
// provider: local
// ---------------------------------------------------------------------------------
// libacp_host.so has NO link-time dependency on the product library: the engine's C ABI
// (include/acp_infer.h) is looked up in the process image the first time provider "local" is used —
// libacp_infer.so must have been loaded with RTLD_GLOBAL (agentcontrolplane_b200/_lib.py), or, in the
// sanitizer harnesses, the stand-in engine is linked into the executable (-rdynamic).
namespace {
struct InferApi {
  int (*submit)(acp_engine*, const char*, size_t, uint64_t*) = nullptr;
  int (*wait)(acp_engine*, uint64_t, int) = nullptr;
  int (*result)(acp_engine*, uint64_t, char**, size_t*, int*) = nullptr;
  void (*cancel)(acp_engine*, uint64_t) = nullptr;
  void (*free_)(void*) = nullptr;
  bool ok() const { return submit && wait && result && cancel && free_; }
};
const InferApi* infer_api() {
  static std::mutex mu;
  static InferApi api;
  std::lock_guard<std::mutex> lk(mu);
  if (!api.ok()) {   // only success is cached: the product may be loaded after the first (failed) look-up
    api.submit = (decltype(api.submit))dlsym(RTLD_DEFAULT, "acp_infer_submit");
    api.wait = (decltype(api.wait))dlsym(RTLD_DEFAULT, "acp_infer_wait");
    api.result = (decltype(api.result))dlsym(RTLD_DEFAULT, "acp_infer_result");
    api.cancel = (decltype(api.cancel))dlsym(RTLD_DEFAULT, "acp_infer_cancel");
    api.free_ = (decltype(api.free_))dlsym(RTLD_DEFAULT, "acp_infer_free");
  }
  return api.ok() ? &api : nullptr;
}
}  // namespace

bool LocalClient::SendRequest(const Context& ctx, const std::vector<Message>& messages,
                              const std::vector<Tool>& tools, Message* out, Error* err) {
  if (!engine_) {
    err->Message = "model API call failed: provider local: engine not initialised";
    return false;
  }
  const std::string body = build_chat_request_json(cfg_.Model, messages, tools, cfg_.MaxTokens,
                                                   has_ext_ ? &ext_ : nullptr, &cfg_);
  const InferApi* api = infer_api();
  if (!api) {
    err->Message = "model API call failed: provider local: libacp_infer.so is not loaded in this process";
    return false;
  }
  uint64_t ticket = 0;
  int rc = api->submit(engine_, body.data(), body.size(), &ticket);
  if (rc != ACP_OK) {
    err->Message = "model API call failed: acp_infer_submit error " + std::to_string(rc);
    return false;
  }
  // Blocking wait that honours ctx.Done() (manager shutdown): poll the flag every 50 ms.
  while (true) {
    rc = api->wait(engine_, ticket, 50);
    if (rc == ACP_OK) break;
    if (rc != ACP_ERR_TIMEOUT) {
      err->Message = "model API call failed: acp_infer_wait error " + std::to_string(rc);
      return false;
    }
    if (ctx.done()) api->cancel(engine_, ticket);
  }
  char* resp = nullptr;
  size_t len = 0;
  int status = 0;
  rc = api->result(engine_, ticket, &resp, &len, &status);
  if (rc != ACP_OK) {
    err->Message = "model API call failed: acp_infer_result error " + std::to_string(rc);
    return false;
  }
  last_response_.assign(resp, len);
  api->free_(resp);
  if (status != 200) {
    Json e;
    std::string perr;
    std::string msg = last_response_;
    if (Json::parse(last_response_, &e, &perr)) msg = e.get("error").get("message").as_string();
    if (status >= 400 && status < 500 && status != 499) {
      // typed error: terminal Failed in handleLLMError (state_machine.go:738-756)
      err->is_request_error = true;
      err->StatusCode = status;
      err->Message = msg;
    } else {
      err->Message = "model API call failed: " + msg;  // langchaingo_client.go:103-105 wrapping
    }
    return false;
  }
  std::string perr;
  if (!convert_from_response_json(last_response_, out, &perr)) {
    err->Message = "model API call failed: bad response JSON: " + perr;
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------
// provider: openai over loopback HTTP (CPU baseline)
// ---------------------------------------------------------------------------------
static bool parse_url(const std::string& url, std::string* host, int* port, std::string* path) {
  std::string u = url;
  const std::string pfx = "http://";
  if (u.compare(0, pfx.size(), pfx) != 0) return false;
  u = u.substr(pfx.size());
  size_t slash = u.find('/');
  std::string hp = slash == std::string::npos ? u : u.substr(0, slash);
  *path = slash == std::string::npos ? "" : u.substr(slash);
  size_t colon = hp.find(':');
  *host = colon == std::string::npos ? hp : hp.substr(0, colon);
  *port = colon == std::string::npos ? 80 : atoi(hp.c_str() + colon + 1);
  while (!path->empty() && path->back() == '/') path->pop_back();
  return !host->empty();
}

// One persistent connection per worker thread and endpoint, like Go's http.Transport keep-alive
// pool that langchaingo's openai client uses.
namespace {
struct Conn {
  int fd = -1;
  std::string key;
  std::string acc;
  ~Conn() { if (fd >= 0) close(fd); }
};
thread_local Conn t_conn;

bool conn_open(Conn& c, const std::string& host, int port) {
  const std::string key = host + ":" + std::to_string(port);
  if (c.fd >= 0 && c.key == key) return true;
  if (c.fd >= 0) { close(c.fd); c.fd = -1; }
  c.acc.clear();
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return false;
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  sockaddr_in addr;
  memset(&addr, 0, sizeof addr);
  addr.sin_family = AF_INET;
  addr.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host == "localhost" ? "127.0.0.1" : host.c_str(), &addr.sin_addr) != 1 ||
      connect(fd, (sockaddr*)&addr, sizeof addr) != 0) {
    close(fd);
    return false;
  }
  c.fd = fd;
  c.key = key;
  return true;
}

// returns false on transport failure; *status / *body filled on success
bool round_trip(Conn& c, const std::string& req, int* status, std::string* body) {
  size_t off = 0;
  while (off < req.size()) {
    ssize_t n = send(c.fd, req.data() + off, req.size() - off, MSG_NOSIGNAL);
    if (n <= 0) return false;
    off += (size_t)n;
  }
  char buf[16384];
  size_t he;
  while ((he = c.acc.find("\r\n\r\n")) == std::string::npos) {
    ssize_t n = recv(c.fd, buf, sizeof buf, 0);
    if (n <= 0) return false;
    c.acc.append(buf, (size_t)n);
  }
  if (c.acc.size() < 12) return false;
  *status = atoi(c.acc.c_str() + 9);
  size_t cl = c.acc.find("Content-Length:");
  const size_t len = (cl == std::string::npos || cl > he) ? 0 : (size_t)atoll(c.acc.c_str() + cl + 15);
  const size_t need = he + 4 + len;
  while (c.acc.size() < need) {
    ssize_t n = recv(c.fd, buf, sizeof buf, 0);
    if (n <= 0) return false;
    c.acc.append(buf, (size_t)n);
  }
  *body = c.acc.substr(he + 4, len);
  c.acc.erase(0, need);
  return true;
}
}  // namespace

bool HTTPClient::SendRequest(const Context&, const std::vector<Message>& messages,
                             const std::vector<Tool>& tools, Message* out, Error* err) {
  std::string host, path;
  int port = 0;
  if (!parse_url(cfg_.BaseURL, &host, &port, &path)) {
    err->Message = "model API call failed: unsupported base URL " + cfg_.BaseURL;
    return false;
  }
  const std::string body = build_chat_request_json(cfg_.Model, messages, tools, 0, nullptr);
  const std::string req = "POST " + path + "/chat/completions HTTP/1.1\r\nHost: " + host + ":" + std::to_string(port) +
                          "\r\nContent-Type: application/json\r\nAuthorization: Bearer " + api_key_ +
                          "\r\nContent-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
  int status = 0;
  std::string rbody;
  bool ok = false;
  for (int attempt = 0; attempt < 2 && !ok; ++attempt) {  // one reconnect if a pooled conn went stale
    if (!conn_open(t_conn, host, port)) {
      err->Message = "model API call failed: connect " + host + ":" + std::to_string(port) + " failed";
      return false;
    }
    ok = round_trip(t_conn, req, &status, &rbody);
    if (!ok) { close(t_conn.fd); t_conn.fd = -1; }
  }
  if (!ok) { err->Message = "model API call failed: connection reset"; return false; }
  if (status != 200) {
    // NOTE: the reference wraps every provider failure as a plain error (langchaingo_client.go:
    // 103-105) and never builds an LLMRequestError itself; restated faithfully.
    err->Message = "model API call failed: API returned unexpected status code: " + std::to_string(status);
    return false;
  }
  std::string perr;
  if (!convert_from_response_json(rbody, out, &perr)) {
    err->Message = "model API call failed: " + perr;
    return false;
  }
  return true;
}

std::unique_ptr<LLMClient> NewLLMClient(const std::string& provider, const std::string& api_key,
                                        const BaseConfig& cfg, acp_engine* engine, std::string* err) {
  if (provider == "local") {
    if (!engine) { *err = "failed to initialize local client: engine not initialised"; return nullptr; }
    return std::unique_ptr<LLMClient>(new LocalClient(engine, cfg));
  }
  if (provider == "openai") return std::unique_ptr<LLMClient>(new HTTPClient(api_key, cfg));
  // langchaingo_client.go:71-72, with `local` added to the list
  *err = "unsupported provider: " + provider +
         ". Supported providers are: openai, anthropic, mistral, google, vertex, local";
  return nullptr;
}

}  // namespace llmclient
}  // namespace acp
