// host/task.cc — see task.h.
#include "task.h"
#include <chrono>
#include <random>
#include <stdio.h>

namespace acp {
namespace task {

// ---------------------------------------------------------------------------------
// (de)serialisation — CRD JSON field names
// ---------------------------------------------------------------------------------
Json task_to_json(const Task& t) {
  Json meta = Json::object();
  meta.set("name", Json(t.Name));
  meta.set("namespace", Json(t.Namespace));
  meta.set("uid", Json(t.UID));
  if (!t.Labels.empty()) {
    Json l = Json::object();
    for (auto& kv : t.Labels) l.set(kv.first, Json(kv.second));
    meta.set("labels", l);
  }
  Json spec = Json::object();
  Json ar = Json::object();
  ar.set("name", Json(t.AgentName));
  spec.set("agentRef", ar);
  if (!t.UserMessage.empty()) spec.set("userMessage", Json(t.UserMessage));
  if (!t.SpecContextWindow.empty()) {
    Json scw = Json::array();
    for (const Message& m : t.SpecContextWindow) scw.push(llmclient::message_to_crd_json(m));
    spec.set("contextWindow", scw);
  }
  Json st = Json::object();
  st.set("ready", Json(t.Status.Ready));
  st.set("status", Json(t.Status.Status));
  st.set("statusDetail", Json(t.Status.StatusDetail));
  st.set("phase", Json(t.Status.Phase));
  if (!t.Status.Output.empty()) st.set("output", Json(t.Status.Output));
  if (!t.Status.Error.empty()) st.set("error", Json(t.Status.Error));
  if (!t.Status.ToolCallRequestID.empty()) st.set("toolCallRequestId", Json(t.Status.ToolCallRequestID));
  Json cw = Json::array();
  for (const Message& m : t.Status.ContextWindow) cw.push(llmclient::message_to_crd_json(m));
  st.set("contextWindow", cw);
  Json root = Json::object();
  root.set("apiVersion", Json("acp.humanlayer.dev/v1alpha1"));
  root.set("kind", Json("Task"));
  root.set("metadata", meta);
  root.set("spec", spec);
  root.set("status", st);
  return root;
}

bool task_from_json(const Json& j, Task* t) {
  if (!j.is_object()) return false;
  const Json& meta = j.get("metadata");
  t->Name = meta.get("name").as_string();
  t->Namespace = meta.get("namespace").as_string();
  if (t->Namespace.empty()) t->Namespace = "default";
  t->UID = meta.get("uid").as_string();
  t->Labels.clear();
  for (auto& kv : meta.get("labels").members()) t->Labels[kv.first] = kv.second.as_string();
  t->AgentName = j.get("spec").get("agentRef").get("name").as_string();
  t->UserMessage = j.get("spec").get("userMessage").as_string();
  t->SpecContextWindow.clear();
  for (const Json& m : j.get("spec").get("contextWindow").items()) {
    Message msg;
    llmclient::message_from_crd_json(m, &msg);
    t->SpecContextWindow.push_back(std::move(msg));
  }
  const Json& st = j.get("status");
  t->Status = TaskStatus();
  t->Status.Ready = st.get("ready").as_bool(false);
  t->Status.Status = st.get("status").as_string();
  t->Status.StatusDetail = st.get("statusDetail").as_string();
  t->Status.Phase = st.get("phase").as_string();
  t->Status.Output = st.get("output").as_string();
  t->Status.Error = st.get("error").as_string();
  t->Status.ToolCallRequestID = st.get("toolCallRequestId").as_string();
  for (const Json& m : st.get("contextWindow").items()) {
    Message msg;
    llmclient::message_from_crd_json(m, &msg);
    t->Status.ContextWindow.push_back(std::move(msg));
  }
  return true;
}

Json toolcall_to_json(const ToolCall& tc) {
  Json meta = Json::object();
  meta.set("name", Json(tc.Name));
  meta.set("namespace", Json(tc.Namespace));
  Json l = Json::object();
  for (auto& kv : tc.Labels) l.set(kv.first, Json(kv.second));
  meta.set("labels", l);
  Json owner = Json::object();
  owner.set("apiVersion", Json("acp.humanlayer.dev/v1alpha1"));
  owner.set("kind", Json("Task"));
  owner.set("name", Json(tc.OwnerName));
  owner.set("uid", Json(tc.OwnerUID));
  owner.set("controller", Json(true));
  Json owners = Json::array();
  owners.push(owner);
  meta.set("ownerReferences", owners);
  Json spec = Json::object();
  spec.set("toolCallId", Json(tc.ToolCallID));
  Json tr = Json::object(); tr.set("name", Json(tc.TaskRef)); spec.set("taskRef", tr);
  Json tf = Json::object(); tf.set("name", Json(tc.ToolRef)); spec.set("toolRef", tf);
  spec.set("toolType", Json(tc.ToolType));
  spec.set("arguments", Json(tc.Arguments));
  Json st = Json::object();
  if (!tc.StatusStatus.empty()) st.set("status", Json(tc.StatusStatus));
  if (!tc.StatusResult.empty()) st.set("result", Json(tc.StatusResult));
  Json root = Json::object();
  root.set("apiVersion", Json("acp.humanlayer.dev/v1alpha1"));
  root.set("kind", Json("ToolCall"));
  root.set("metadata", meta);
  root.set("spec", spec);
  root.set("status", st);
  return root;
}

bool toolcall_from_json(const Json& j, ToolCall* tc) {
  if (!j.is_object()) return false;
  const Json& meta = j.get("metadata");
  tc->Name = meta.get("name").as_string();
  tc->Namespace = meta.get("namespace").as_string();
  tc->Labels.clear();
  for (auto& kv : meta.get("labels").members()) tc->Labels[kv.first] = kv.second.as_string();
  const auto& owners = meta.get("ownerReferences").items();
  if (!owners.empty()) { tc->OwnerName = owners[0].get("name").as_string(); tc->OwnerUID = owners[0].get("uid").as_string(); }
  const Json& spec = j.get("spec");
  tc->ToolCallID = spec.get("toolCallId").as_string();
  tc->TaskRef = spec.get("taskRef").get("name").as_string();
  tc->ToolRef = spec.get("toolRef").get("name").as_string();
  tc->ToolType = spec.get("toolType").as_string();
  tc->Arguments = spec.get("arguments").as_string();
  tc->StatusStatus = j.get("status").get("status").as_string();
  tc->StatusResult = j.get("status").get("result").as_string();
  return true;
}

// ---------------------------------------------------------------------------------
// object store
// ---------------------------------------------------------------------------------
bool ObjectStore::Get(const std::string& kind, const std::string& name, Json* out) {
  std::string text;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ++reads_;
    auto it = objs_.find(kind + "/" + name);
    if (it == objs_.end()) return false;
    text = it->second;
  }
  std::string err;
  return Json::parse(text, out, &err);  // deserialise outside the lock, like a client would
}
void ObjectStore::Put(const std::string& kind, const std::string& name, const Json& obj) {
  std::string text = obj.dump();
  std::lock_guard<std::mutex> lk(mu_);
  ++writes_;
  objs_[kind + "/" + name] = std::move(text);
}
bool ObjectStore::Create(const std::string& kind, const std::string& name, const Json& obj) {
  std::string text = obj.dump();
  std::lock_guard<std::mutex> lk(mu_);
  ++writes_;   // the request reaches the API server either way
  return objs_.emplace(kind + "/" + name, std::move(text)).second;
}
bool ObjectStore::Delete(const std::string& kind, const std::string& name) {
  std::lock_guard<std::mutex> lk(mu_);
  ++writes_;
  return objs_.erase(kind + "/" + name) > 0;
}
std::vector<Json> ObjectStore::ListToolCalls(const std::string& task, const std::string& request_id) {
  std::vector<std::string> texts;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ++reads_;
    const std::string pfx = "ToolCall/";
    for (auto it = objs_.lower_bound(pfx); it != objs_.end() && it->first.compare(0, pfx.size(), pfx) == 0; ++it)
      texts.push_back(it->second);
  }
  std::vector<Json> out;
  for (auto& t : texts) {
    Json j;
    std::string err;
    if (!Json::parse(t, &j, &err)) continue;
    const Json& labels = j.get("metadata").get("labels");
    if (labels.get("acp.humanlayer.dev/task").as_string() == task &&
        labels.get("acp.humanlayer.dev/toolcallrequest").as_string() == request_id)
      out.push_back(std::move(j));
  }
  return out;
}

// ---------------------------------------------------------------------------------
// pure helpers
// ---------------------------------------------------------------------------------
std::vector<Message> buildInitialContextWindow(const std::vector<Message>& contextWindow,
                                               const std::string& systemPrompt,
                                               const std::string& userMessage) {
  std::vector<Message> out;
  if (!contextWindow.empty()) {
    out = contextWindow;
    bool has_system = false;
    for (const Message& m : out) if (m.Role == "system") { has_system = true; break; }
    if (!has_system) {
      Message s; s.Role = "system"; s.Content = systemPrompt;
      out.insert(out.begin(), s);
    }
  } else {
    Message s; s.Role = "system"; s.Content = systemPrompt;
    Message u; u.Role = "user"; u.Content = userMessage;
    out.push_back(s);
    out.push_back(u);
  }
  return out;
}

std::map<std::string, std::string> buildToolTypeMap(const std::vector<Tool>& tools) {
  std::map<std::string, std::string> m;
  for (const Tool& t : tools) m[t.Function.Name] = t.ACPToolType;
  return m;
}

std::string ValidateTaskMessageInput(const std::string& userMessage, const std::vector<Message>& cw) {
  if (!userMessage.empty() && !cw.empty()) return "only one of userMessage or contextWindow can be provided";
  if (userMessage.empty() && cw.empty()) return "one of userMessage or contextWindow must be provided";
  if (!cw.empty()) {
    bool has_user = false;
    for (const Message& m : cw) {
      if (m.Role != "system" && m.Role != "user" && m.Role != "assistant" && m.Role != "tool")
        return "invalid role in contextWindow: " + m.Role;
      if (m.Role == "user") has_user = true;
    }
    if (!has_user) return "contextWindow must contain at least one user message";
  }
  return "";
}

std::string GetUserMessagePreview(const std::string& userMessage, const std::vector<Message>& cw) {
  std::string preview;
  if (!userMessage.empty()) preview = userMessage;
  else
    for (auto it = cw.rbegin(); it != cw.rend(); ++it)
      if (it->Role == "user") { preview = it->Content; break; }
  if (preview.size() > 50) preview = preview.substr(0, 47) + "...";
  return preview;
}

std::string GenerateK8sRandomString(int n) {
  if (n < 1 || n > 8) n = 6;
  static const char letters[] = "abcdefghijklmnopqrstuvwxyz";
  static const char alnum[] = "abcdefghijklmnopqrstuvwxyz0123456789";
  thread_local std::random_device rd;  // crypto/rand in the reference
  std::string s;
  s.push_back(letters[rd() % 26]);
  for (int i = 1; i < n; ++i) s.push_back(alnum[rd() % 36]);
  return s;
}

static Json message_param_schema() {
  Json msg = Json::object(); msg.set("type", Json("string"));
  Json props = Json::object(); props.set("message", msg);
  Json req = Json::array(); req.push(Json("message"));
  Json p = Json::object();
  p.set("type", Json("object"));
  p.set("properties", props);
  p.set("required", req);
  return p;
}

std::vector<Tool> ConvertSubAgents(const std::vector<std::pair<std::string, std::string>>& agents) {
  std::vector<Tool> out;
  for (auto& a : agents) {
    Tool t;
    t.Type = "function";
    t.Function.Name = "delegate_to_agent__" + a.first;
    t.Function.Description = a.second;
    t.Function.Parameters = message_param_schema();
    t.ACPToolType = "DelegateToAgent";
    out.push_back(std::move(t));
  }
  return out;
}

std::vector<Tool> ConvertMCPTools(const std::vector<Json>& mcpTools, const std::string& serverName) {
  std::vector<Tool> out;
  for (const Json& mt : mcpTools) {
    Tool t;
    t.Type = "function";
    t.Function.Name = serverName + "__" + mt.get("name").as_string();
    t.Function.Description = mt.get("description").as_string();
    const Json& schema = mt.get("inputSchema");
    if (schema.is_object()) t.Function.Parameters = schema;
    else {
      Json p = Json::object();
      p.set("type", Json("object"));
      p.set("properties", Json::object());
      t.Function.Parameters = p;
    }
    t.ACPToolType = "MCP";
    out.push_back(std::move(t));
  }
  return out;
}

Tool ToolFromContactChannel(const Json& channel) {
  Tool t;
  t.Type = "function";
  t.Function.Parameters = message_param_schema();
  const std::string name = channel.get("metadata").get("name").as_string();
  const Json& spec = channel.get("spec");
  const std::string type = spec.get("type").as_string();
  if (type == "email") {
    t.Function.Name = name + "__human_contact_email";
    t.Function.Description = spec.get("email").get("contextAboutUser").as_string();
    if (t.Function.Description.empty()) t.Function.Description = "Contact a human via email";
  } else if (type == "slack") {
    t.Function.Name = name + "__human_contact_slack";
    t.Function.Description = spec.get("slack").get("contextAboutChannelOrUser").as_string();
    if (t.Function.Description.empty()) t.Function.Description = "Contact a human via Slack";
  } else {
    t.Function.Name = name + "__human_contact";
    t.Function.Description = "Contact a human via " + type + " channel";
  }
  t.ACPToolType = "HumanContact";
  return t;
}

// ---------------------------------------------------------------------------------
// state machine
// ---------------------------------------------------------------------------------
StateMachine::StateMachine(ObjectStore* store, Recorder* recorder) : store_(store), recorder_(recorder) {
  now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
}

// getTaskMutex (state_machine.go:1147-1170): one mutex per task name, created on first use
std::mutex* StateMachine::getTaskMutex(const std::string& taskName) {
  std::lock_guard<std::mutex> lk(mutex_map_lock_);
  auto& slot = task_mutexes_[taskName];
  if (!slot) slot.reset(new std::mutex());
  return slot.get();
}

// canAcquireLease (:1121-1132): we already hold it, or it has no renew time, or it has expired
bool StateMachine::canAcquireLease(const Json& lease) const {
  const Json& spec = lease.get("spec");
  if (spec.get("holderIdentity").as_string() == podName) return true;
  if (!spec.get("renewTime").is_number()) return true;
  return now() > spec.get("renewTime").as_double() + leaseDurationSeconds;
}

// acquireTaskLease (:1069-1118): Create; on AlreadyExists read it and take it over if allowed
int StateMachine::acquireTaskLease(const std::string& taskName) {
  const std::string leaseName = "task-llm-" + taskName;
  const double t = now();
  Json spec = Json::object();
  spec.set("holderIdentity", Json(podName));
  spec.set("leaseDurationSeconds", Json((int)leaseDurationSeconds));
  spec.set("acquireTime", Json(t));
  spec.set("renewTime", Json(t));
  Json meta = Json::object();
  meta.set("name", Json(leaseName));
  Json lease = Json::object();
  lease.set("metadata", meta);
  lease.set("spec", spec);
  if (store_->Create("Lease", leaseName, lease)) return 0;
  Json existing;
  if (!store_->Get("Lease", leaseName, &existing)) return -1;   // deleted in between: the reference returns the Get error
  if (!canAcquireLease(existing)) return 1;
  store_->Put("Lease", leaseName, lease);                       // Update with us as holder
  return 0;
}

void StateMachine::releaseTaskLease(const std::string& taskName) { store_->Delete("Lease", "task-llm-" + taskName); }

// apierrors.NewNotFound(...).Error() for a GET of `kind` `name` (what the reference stores in Status.Error)
static std::string not_found_error(const std::string& resource, const std::string& name) {
  return resource + ".acp.humanlayer.dev \"" + name + "\" not found";
}

Result StateMachine::validateTaskAndAgent(Task* task, Task* statusUpdate, Json* agent, bool* ok, std::string* err) {
  (void)err;
  *ok = false;
  TaskStatus& st = statusUpdate->Status;
  if (!store_->Get("Agent", task->AgentName, agent)) {   // apierrors.IsNotFound (:384-391)
    st.Ready = false;
    st.Status = "Pending";
    st.Phase = "Pending";
    st.StatusDetail = "Waiting for Agent to exist";
    st.Error.clear();
    recorder_->Emit("Normal", "Waiting", "Waiting for Agent to exist");
    store_->Put("Task", task->Name, task_to_json(*statusUpdate));
    *task = *statusUpdate;
    Result r; r.RequeueAfter = DefaultRequeueDelay;
    return r;
  }
  if (!agent->get("status").get("ready").as_bool(false)) {  // :408-420
    const std::string msg = "Waiting for agent \"" + agent->get("metadata").get("name").as_string() + "\" to become ready";
    st.Ready = false;
    st.Status = "Pending";
    st.Phase = "Pending";
    st.StatusDetail = msg;
    st.Error.clear();
    recorder_->Emit("Normal", "Waiting", msg);
    store_->Put("Task", task->Name, task_to_json(*statusUpdate));
    *task = *statusUpdate;
    Result r; r.RequeueAfter = DefaultRequeueDelay;
    return r;
  }
  *ok = true;
  return Result();
}

bool StateMachine::getLLMAndCredentials(const Json& agent, Task* task, Task* statusUpdate, Json* llm,
                                        std::string* apiKey, std::string* err) {
  TaskStatus& st = statusUpdate->Status;
  auto fail = [&](const std::string& detail, const std::string& error, const std::string& reason, const std::string& event) {
    st.Ready = false;
    st.Status = "Error";
    st.Phase = "Failed";
    st.StatusDetail = detail;
    st.Error = error;
    recorder_->Emit("Warning", reason, event);
    store_->Put("Task", task->Name, task_to_json(*statusUpdate));
    *task = *statusUpdate;
    *err = error;
    return false;
  };
  const std::string llm_name = agent.get("spec").get("llmRef").get("name").as_string();
  if (!store_->Get("LLM", llm_name, llm)) {   // :485-497
    const std::string e = not_found_error("llms", llm_name);
    return fail("Failed to get LLM: " + e, e, "LLMFetchFailed", e);
  }
  const Json& spec = llm->get("spec");
  const Json& key_from = spec.get("apiKeyFrom");
  apiKey->clear();
  if (spec.get("provider").as_string() == "local" && !key_from.is_object())
    return true;   // no credentials: the engine lives in this process (the reference would nil-deref at :504)
  const std::string secret_name = key_from.get("secretKeyRef").get("name").as_string();
  Json secret;
  if (!store_->Get("Secret", secret_name, &secret)) {   // :501-517
    // core/v1 objects have no API group in the NotFound text
    const std::string e_core = "secrets \"" + secret_name + "\" not found";
    return fail("Failed to get API key secret: " + e_core, e_core, "APIKeySecretFetchFailed", e_core);
  }
  *apiKey = secret.get("data").get(key_from.get("secretKeyRef").get("key").as_string()).as_string();
  if (apiKey->empty())   // :520-535
    return fail("API key is empty", "API key is empty", "EmptyAPIKey", "API key is empty");
  return true;
}

std::vector<Tool> StateMachine::collectTools(const Json& agent, const MCPToolsByServer& mcp) {
  std::vector<Tool> tools;
  const Json& spec = agent.get("spec");
  for (const Json& ref : spec.get("mcpServers").items()) {          // :545-554
    auto it = mcp.find(ref.get("name").as_string());
    if (it == mcp.end()) continue;
    for (Tool& t : ConvertMCPTools(it->second, it->first)) tools.push_back(std::move(t));
  }
  for (const Json& ref : agent.get("status").get("validHumanContactChannels").items()) {   // :557-567
    Json ch;
    if (!store_->Get("ContactChannel", ref.get("name").as_string(), &ch)) continue;
    tools.push_back(ToolFromContactChannel(ch));
  }
  std::vector<std::pair<std::string, std::string>> subs;               // :570-580
  for (const Json& ref : spec.get("subAgents").items()) {
    Json sub;
    if (!store_->Get("Agent", ref.get("name").as_string(), &sub)) continue;
    subs.emplace_back(sub.get("metadata").get("name").as_string(), sub.get("spec").get("description").as_string());
  }
  for (Tool& t : ConvertSubAgents(subs)) tools.push_back(std::move(t));
  return tools;
}

// mutex + Lease around an LLM step (state_machine.go:166-181); `body` runs while both are held
Result StateMachine::withTaskLock(const std::string& taskName, const std::function<Result()>& body) {
  std::lock_guard<std::mutex> task_lock(*getTaskMutex(taskName));   // :167-169
  if (emulate_lease) {        // acquireTaskLease -> API write #1 (:172-181)
    const int rc = acquireTaskLease(taskName);
    if (rc < 0) { Result r; r.RequeueAfter = 2.0; return r; }        // :174-176
    if (rc > 0) { Result r; r.RequeueAfter = 5.0; return r; }        // held by another pod (:177-180)
  }
  struct LeaseGuard {
    StateMachine* sm; std::string name; bool on;
    ~LeaseGuard() { if (on) sm->releaseTaskLease(name); }  // defer releaseTaskLease (:181)
  } guard{this, taskName, emulate_lease};
  return body();
}

// sendLLMRequest in the reference's own order (state_machine.go:162-288): mutex, Lease,
// validateTaskAndAgent, getLLMAndCredentials, CreateClient (provider switch on the LLM CR), collectTools,
// then the step.
Result StateMachine::sendLLMRequestFromCluster(const llmclient::Context& ctx, Task* task, const MCPToolsByServer& mcp,
                                               acp_engine* engine, std::string* err) {
  err->clear();
  return withTaskLock(task->Name, [&]() -> Result {
    Task statusUpdate = *task;
    Json agent;
    bool ok = false;
    Result r = validateTaskAndAgent(task, &statusUpdate, &agent, &ok, err);   // :183-186
    if (!ok) return r;
    Json llm;
    std::string apiKey;
    if (!getLLMAndCredentials(agent, task, &statusUpdate, &llm, &apiKey, err)) return Result();   // :188-191
    const std::string provider = llm.get("spec").get("provider").as_string();
    const llmclient::BaseConfig bc = llmclient::base_config_from_json(llm.get("spec").get("parameters"));
    ClientFactory factory = [&](std::string* cerr) {
      auto c = llmclient::NewLLMClient(provider, apiKey, bc, engine, cerr);
      if (c && client_hook) client_hook(c.get());
      return c;
    };
    const std::vector<Tool> tools = collectTools(agent, mcp);   // :220 (after CreateClient in the reference; no observable difference)
    return llmStepLocked(ctx, task, tools, factory, err);
  });
}

Result StateMachine::sendLLMRequest(const llmclient::Context& ctx, Task* task, const std::vector<Tool>& tools,
                                    const ClientFactory& factory, std::string* err) {
  err->clear();
  return withTaskLock(task->Name, [&]() -> Result { return llmStepLocked(ctx, task, tools, factory, err); });
}

Result StateMachine::llmStepLocked(const llmclient::Context& ctx, Task* task, const std::vector<Tool>& tools,
                                   const ClientFactory& factory, std::string* err) {
  Task statusUpdate = *task;  // task.DeepCopy()  (:164)
  std::string cerr;
  std::unique_ptr<llmclient::LLMClient> client = factory(&cerr);  // CreateClient (:195)
  if (!client) {
    statusUpdate.Status.Ready = false;
    statusUpdate.Status.Status = "Error";
    statusUpdate.Status.Phase = "Failed";
    statusUpdate.Status.StatusDetail = "Failed to create LLM client: " + cerr;
    statusUpdate.Status.Error = cerr;
    recorder_->Emit("Warning", "LLMClientCreationFailed", cerr);
    store_->Put("Task", task->Name, task_to_json(statusUpdate));
    *task = statusUpdate;
    return Result();
  }
  if (task->Status.Phase != "ReadyForLLM" || statusUpdate.Status.StatusDetail != "Sending request to LLM") {
    recorder_->Emit("Normal", "SendingContextWindowToLLM", "Sending context window to LLM");  // :223-230
    statusUpdate.Status.StatusDetail = "Sending request to LLM";
    store_->Put("Task", task->Name, task_to_json(statusUpdate));  // API write #2
  }
  Message output;
  llmclient::Error e;
  if (!client->SendRequest(ctx, task->Status.ContextWindow, tools, &output, &e)) {  // :238
    Result r = handleLLMError(&statusUpdate, e, err);
    *task = statusUpdate;
    return r;
  }
  Result r = processLLMResponse(output, task, &statusUpdate, tools, err);
  if (!err->empty()) {
    statusUpdate.Status.Ready = false;
    statusUpdate.Status.Status = "Error";
    statusUpdate.Status.Phase = "Failed";
    statusUpdate.Status.StatusDetail = "Failed to process LLM response: " + *err;
    statusUpdate.Status.Error = *err;
    recorder_->Emit("Warning", "LLMResponseProcessingFailed", *err);
    store_->Put("Task", task->Name, task_to_json(statusUpdate));
    *task = statusUpdate;
    err->clear();
    return Result();
  }
  if (!r.IsZero()) { *task = statusUpdate; return r; }
  store_->Put("Task", task->Name, task_to_json(statusUpdate));  // final status write (:280)
  *task = statusUpdate;
  return Result();
}

Result StateMachine::processLLMResponse(const Message& output, Task* task, Task* statusUpdate,
                                        const std::vector<Tool>& tools, std::string* err) {
  TaskStatus& st = statusUpdate->Status;
  if (!output.Content.empty()) {
    st.Output = output.Content;
    st.Phase = "FinalAnswer";
    st.Ready = true;
    Message m; m.Role = "assistant"; m.Content = output.Content;
    st.ContextWindow.push_back(m);
    st.Status = "Ready";
    st.StatusDetail = "LLM final response received";
    st.Error.clear();
    if (task->Status.Phase != "FinalAnswer")
      recorder_->Emit("Normal", "LLMFinalAnswer", "LLM response received successfully");
    return Result();
  }
  const std::string reqId = GenerateK8sRandomString(7);
  st.Output.clear();
  st.Phase = "ToolCallsPending";
  st.ToolCallRequestID = reqId;
  Message m; m.Role = "assistant"; m.ToolCalls = output.ToolCalls;
  st.ContextWindow.push_back(m);
  st.Ready = true;
  st.Status = "Ready";
  st.StatusDetail = "LLM response received, tool calls pending";
  st.Error.clear();
  recorder_->Emit("Normal", "ToolCallsPending", "LLM response received, tool calls pending");
  store_->Put("Task", statusUpdate->Name, task_to_json(*statusUpdate));  // :659-663
  return createToolCalls(task, statusUpdate, output.ToolCalls, tools, err);
}

Result StateMachine::createToolCalls(Task* task, Task* statusUpdate,
                                     const std::vector<llmclient::MessageToolCall>& toolCalls,
                                     const std::vector<Tool>& tools, std::string* err) {
  (void)task;
  if (statusUpdate->Status.ToolCallRequestID.empty()) {
    *err = "no ToolCallRequestID found in statusUpdate, cannot create tool calls";
    return Result();
  }
  const auto typeMap = buildToolTypeMap(tools);
  for (size_t i = 0; i < toolCalls.size(); ++i) {
    char suffix[16];
    snprintf(suffix, sizeof suffix, "-tc-%02d", (int)(i + 1));
    ToolCall tc;
    tc.Name = statusUpdate->Name + "-" + statusUpdate->Status.ToolCallRequestID + suffix;  // :690
    tc.Namespace = statusUpdate->Namespace;
    tc.Labels["acp.humanlayer.dev/task"] = statusUpdate->Name;
    tc.Labels["acp.humanlayer.dev/toolcallrequest"] = statusUpdate->Status.ToolCallRequestID;
    tc.OwnerName = statusUpdate->Name;
    tc.OwnerUID = statusUpdate->UID;
    tc.ToolCallID = toolCalls[i].ID;
    tc.TaskRef = statusUpdate->Name;
    tc.ToolRef = toolCalls[i].Function.Name;
    auto it = typeMap.find(toolCalls[i].Function.Name);
    tc.ToolType = it == typeMap.end() ? "" : it->second;
    tc.Arguments = toolCalls[i].Function.Arguments;  // verbatim
    store_->Put("ToolCall", tc.Name, toolcall_to_json(tc));
    recorder_->Emit("Normal", "ToolCallCreated", "Created ToolCall " + tc.Name);
  }
  Result r;
  r.RequeueAfter = DefaultRequeueDelay;
  return r;
}

Result StateMachine::handleLLMError(Task* statusUpdate, const llmclient::Error& e, std::string* err) {
  const bool is4xx = e.is_request_error && e.StatusCode >= 400 && e.StatusCode < 500;
  TaskStatus& st = statusUpdate->Status;
  const std::string text = e.Error_();
  st.Ready = false;
  st.Status = "Error";
  st.StatusDetail = "LLM request failed: " + text;
  st.Error = text;
  if (is4xx) {
    st.Phase = "Failed";
    recorder_->Emit("Warning", "LLMRequestFailed4xx",
                    "LLM request failed with status " + std::to_string(e.StatusCode) + ": " + e.Message);
  } else {
    recorder_->Emit("Warning", "LLMRequestFailed", text);  // phase preserved (will retry)
  }
  store_->Put("Task", statusUpdate->Name, task_to_json(*statusUpdate));
  if (is4xx) return Result();
  *err = text;
  Result r;
  r.RequeueAfter = DefaultRequeueDelay;
  return r;
}

Result StateMachine::checkToolCalls(Task* task, std::string* err) {
  err->clear();
  std::vector<Json> items = store_->ListToolCalls(task->Name, task->Status.ToolCallRequestID);
  for (const Json& j : items) {
    const std::string s = j.get("status").get("status").as_string();
    if (s != "Succeeded" && s != "Error") {
      Result r;
      r.RequeueAfter = DefaultRequeueDelay;
      return r;
    }
  }
  for (const Json& j : items) {
    Message m;
    m.Role = "tool";
    m.Content = j.get("status").get("result").as_string();
    m.ToolCallID = j.get("spec").get("toolCallId").as_string();
    task->Status.ContextWindow.push_back(std::move(m));
  }
  task->Status.Phase = "ReadyForLLM";
  task->Status.Status = "Ready";
  task->Status.StatusDetail = "All tool calls completed, ready to send tool results to LLM";
  task->Status.Error.clear();
  recorder_->Emit("Normal", "AllToolCallsCompleted", "All tool calls completed");
  store_->Put("Task", task->Name, task_to_json(*task));
  Result r;
  r.Requeue = true;
  return r;
}

// ---------------------------------------------------------------------------------
// Process / Reconcile
// ---------------------------------------------------------------------------------
Result StateMachine::initialize(Task* task, std::string* err) {
  err->clear();
  task->Status.Phase = "Initializing";
  task->Status.Status = "Pending";
  task->Status.StatusDetail = "Initializing Task";
  store_->Put("Task", task->Name, task_to_json(*task));   // the root span context is out of scope (OTel)
  Result r;
  r.Requeue = true;
  return r;
}

Result StateMachine::prepareForLLM(Task* task, Task* statusUpdate, const Json& agent, std::string* err) {
  TaskStatus& st = statusUpdate->Status;
  if (st.Phase != "Initializing" && st.Phase != "Pending") return Result();
  const std::string verr = ValidateTaskMessageInput(task->UserMessage, task->SpecContextWindow);
  if (!verr.empty()) {   // setValidationError (:462-476): returns the error to the controller
    st.Ready = false;
    st.Status = "Error";
    st.Phase = "Failed";
    st.StatusDetail = verr;
    st.Error = verr;
    recorder_->Emit("Warning", "ValidationFailed", verr);
    store_->Put("Task", task->Name, task_to_json(*statusUpdate));
    *task = *statusUpdate;
    *err = verr;
    return Result();
  }
  // ValidateContactChannelRef (:434) belongs to the HumanLayer v1beta3 path: out of scope
  st.ContextWindow = buildInitialContextWindow(task->SpecContextWindow, agent.get("spec").get("system").as_string(), task->UserMessage);
  st.Phase = "ReadyForLLM";
  st.Ready = true;
  st.Status = "Ready";
  st.StatusDetail = "Ready to send to LLM";
  st.Error.clear();
  if (task->Status.Phase != "ReadyForLLM") recorder_->Emit("Normal", "ValidationSucceeded", "Task validation succeeded");
  store_->Put("Task", task->Name, task_to_json(*statusUpdate));
  *task = *statusUpdate;
  Result r;
  r.Requeue = true;
  return r;
}

Result StateMachine::validateAgent(Task* task, std::string* err) {
  err->clear();
  Task statusUpdate = *task;
  Json agent;
  bool ok = false;
  Result r = validateTaskAndAgent(task, &statusUpdate, &agent, &ok, err);
  if (!ok) return r;
  return prepareForLLM(task, &statusUpdate, agent, err);
}

Result StateMachine::Process(const llmclient::Context& ctx, Task* task, const MCPToolsByServer& mcp, acp_engine* engine,
                             std::string* err) {
  err->clear();
  const std::string& phase = task->Status.Phase;
  if (phase == "FinalAnswer" || phase == "Failed") return Result();   // handleTerminal: only ends the trace
  if (phase.empty()) return initialize(task, err);
  if (phase == "Initializing" || phase == "Pending") return validateAgent(task, err);
  if (phase == "ReadyForLLM") return sendLLMRequestFromCluster(ctx, task, mcp, engine, err);
  if (phase == "ToolCallsPending") return checkToolCalls(task, err);
  return Result();   // handleUnknownPhase
}

Result TaskReconciler::Reconcile(const llmclient::Context& ctx, const std::string& taskName, const MCPToolsByServer& mcp,
                                 acp_engine* engine, std::string* err) {
  err->clear();
  Json tj;
  if (!store_->Get("Task", taskName, &tj)) return Result();   // client.IgnoreNotFound (:217-219)
  Task t;
  if (!task_from_json(tj, &t)) return Result();
  return sm_.Process(ctx, &t, mcp, engine, err);
}

}  // namespace task
}  // namespace acp
