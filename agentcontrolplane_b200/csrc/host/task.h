// host/task.h — C++ mirror of the Task reconciler's LLM step (the caller side of the boundary):
//   StateMachine.sendLLMRequest   acp/internal/controller/task/state_machine.go:162-288
//   processLLMResponse            :605-674      createToolCalls   :676-731
//   handleLLMError                :733-790      checkToolCalls    :291-341
//   buildInitialContextWindow     task_helpers.go:13-44    buildToolTypeMap :48-54
//   validation.*                  acp/internal/validation/task_validation.go:16-87
// Kubernetes itself is out of scope: the API server is emulated by an in-memory
// object store that JSON-(de)serialises every object on Get / Update / Create, so the per-step
// API traffic of the reference (>= 4 writes + M ToolCall creates) stays on the clock.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "llmclient.h"

namespace acp {
namespace task {

using llmclient::Message;
using llmclient::Tool;

constexpr double DefaultRequeueDelay = 5.0;  // task_controller.go:23

struct Result {  // ctrl.Result
  bool Requeue = false;
  double RequeueAfter = 0;
  bool IsZero() const { return !Requeue && RequeueAfter == 0; }
};

struct Event { std::string Type, Reason, Message; };
struct Recorder {
  std::mutex mu;
  std::vector<Event> events;
  void Emit(const std::string& type, const std::string& reason, const std::string& message) {
    std::lock_guard<std::mutex> lk(mu);
    events.push_back(Event{type, reason, message});
  }
};

struct TaskStatus {  // acp.TaskStatus (acp/api/v1alpha1/task_types.go:108-158), fields on this path
  bool Ready = false;
  std::string Status, StatusDetail, Phase, Output, Error, ToolCallRequestID;
  std::vector<Message> ContextWindow;
};
struct Task {
  std::string Name, Namespace = "default", UID, AgentName, UserMessage;
  std::vector<Message> SpecContextWindow;   // task.Spec.ContextWindow (alternative to UserMessage)
  std::map<std::string, std::string> Labels;
  TaskStatus Status;
};
struct ToolCall {  // acp.ToolCall (toolcall_types.go:26-45) as created by createToolCalls
  std::string Name, Namespace;
  std::map<std::string, std::string> Labels;
  std::string OwnerName, OwnerUID;
  std::string ToolCallID, TaskRef, ToolRef, ToolType, Arguments;
  std::string StatusStatus, StatusResult;  // filled by the (out of scope) ToolCall controller
};

Json task_to_json(const Task& t);
bool task_from_json(const Json& j, Task* t);
Json toolcall_to_json(const ToolCall& tc);
bool toolcall_from_json(const Json& j, ToolCall* tc);

// In-memory stand-in for the kube-apiserver: objects live as serialised JSON strings.
class ObjectStore {
 public:
  bool Get(const std::string& kind, const std::string& name, Json* out);
  void Put(const std::string& kind, const std::string& name, const Json& obj);  // create or update
  bool Create(const std::string& kind, const std::string& name, const Json& obj);  // false = AlreadyExists
  bool Delete(const std::string& kind, const std::string& name);
  std::vector<Json> ListToolCalls(const std::string& task, const std::string& request_id);
  long long writes() const { return writes_; }
  long long reads() const { return reads_; }

 private:
  std::mutex mu_;
  std::map<std::string, std::string> objs_;  // "kind/name" -> JSON text
  long long writes_ = 0, reads_ = 0;
};

// pure helpers
std::vector<Message> buildInitialContextWindow(const std::vector<Message>& contextWindow,
                                               const std::string& systemPrompt,
                                               const std::string& userMessage);
std::map<std::string, std::string> buildToolTypeMap(const std::vector<Tool>& tools);
std::string ValidateTaskMessageInput(const std::string& userMessage, const std::vector<Message>& cw);
std::string GetUserMessagePreview(const std::string& userMessage, const std::vector<Message>& cw);
std::string GenerateK8sRandomString(int n);
std::vector<Tool> ConvertSubAgents(const std::vector<std::pair<std::string, std::string>>& agents);
std::vector<Tool> ConvertMCPTools(const std::vector<Json>& mcpTools, const std::string& serverName);
// llmclient.ToolFromContactChannel (acp/internal/llmclient/llm_client.go:53-99); `channel` is the
// ContactChannel CR as JSON (metadata.name, spec.type, spec.email.contextAboutUser, spec.slack.contextAboutChannelOrUser)
Tool ToolFromContactChannel(const Json& channel);

// mcpManager.GetTools (out of scope: the MCP servers themselves): server name -> its tool list
using MCPToolsByServer = std::map<std::string, std::vector<Json>>;

using ClientFactory = std::function<std::unique_ptr<llmclient::LLMClient>(std::string* err)>;

class StateMachine {
 public:
  StateMachine(ObjectStore* store, Recorder* recorder);
  // ---- StateMachine.Process (state_machine.go:84-114): terminal -> handleTerminal; no phase ->
  // initialize; Initializing / Pending -> validateAgent (validateTaskAndAgent + prepareForLLM);
  // ReadyForLLM -> sendLLMRequest; ToolCallsPending -> checkToolCalls; anything else -> no-op.
  Result Process(const llmclient::Context& ctx, Task* task, const MCPToolsByServer& mcp, acp_engine* engine,
                 std::string* err);
  Result initialize(Task* task, std::string* err);                                    // :119-146
  Result validateAgent(Task* task, std::string* err);                                 // :149-160
  Result prepareForLLM(Task* task, Task* statusUpdate, const Json& agent, std::string* err);   // :426-460
  // ---- per-task mutex + distributed Lease (state_machine.go:166-181, 1069-1145) ----
  std::mutex* getTaskMutex(const std::string& taskName);
  // 0 = acquired, 1 = held by another pod (requeue 5 s), -1 = API error (requeue 2 s)
  int acquireTaskLease(const std::string& taskName);
  bool canAcquireLease(const Json& lease) const;
  void releaseTaskLease(const std::string& taskName);
  std::string podName = "acp-controller-manager-test";   // POD_NAME (:66-70)
  double leaseDurationSeconds = 30.0;                     // :81
  std::function<double()> now;                            // seconds; injectable clock for the lease tests
  // called on every client CreateClient returns (hostsim scripts forced tool calls on LocalClient)
  std::function<void(llmclient::LLMClient*)> client_hook;
  // The LLM step.  `err` receives the returned Go error text ("" = nil).
  Result sendLLMRequest(const llmclient::Context& ctx, Task* task, const std::vector<Tool>& tools,
                        const ClientFactory& factory, std::string* err);
  Result processLLMResponse(const Message& output, Task* task, Task* statusUpdate,
                            const std::vector<Tool>& tools, std::string* err);
  Result createToolCalls(Task* task, Task* statusUpdate, const std::vector<llmclient::MessageToolCall>& toolCalls,
                         const std::vector<Tool>& tools, std::string* err);
  Result handleLLMError(Task* statusUpdate, const llmclient::Error& e, std::string* err);
  Result checkToolCalls(Task* task, std::string* err);
  // ---- the whole ReadyForLLM arm of the reference, Kubernetes lookups included (emulated store) ----
  // validateTaskAndAgent (state_machine.go:379-424): Agent exists and Status.Ready; *agent = the CR.
  // Returns a non-zero Result (RequeueAfter 5 s) when the Task has to wait; *ok tells them apart.
  Result validateTaskAndAgent(Task* task, Task* statusUpdate, Json* agent, bool* ok, std::string* err);
  // getLLMAndCredentials (state_machine.go:480-538).  provider `local` needs no Secret: the
  // reference dereferences llm.Spec.APIKeyFrom unconditionally (:504) and rejects an empty key
  // (:520-535); both are skipped for `local` when apiKeyFrom is absent.
  bool getLLMAndCredentials(const Json& agent, Task* task, Task* statusUpdate, Json* llm, std::string* apiKey,
                            std::string* err);
  // collectTools (state_machine.go:540-583): MCP server tools, contact channels, sub-agents, in that order
  std::vector<Tool> collectTools(const Json& agent, const MCPToolsByServer& mcp);
  // sendLLMRequest exactly as the reference runs it (state_machine.go:162-288): lease, a5, a6,
  // CreateClient from the LLM CR (provider switch), a8, then the LLM step above.
  Result sendLLMRequestFromCluster(const llmclient::Context& ctx, Task* task, const MCPToolsByServer& mcp,
                                   acp_engine* engine, std::string* err);
  Result withTaskLock(const std::string& taskName, const std::function<Result()>& body);
  Result llmStepLocked(const llmclient::Context& ctx, Task* task, const std::vector<Tool>& tools,
                       const ClientFactory& factory, std::string* err);
  // when true the per-step Lease of acquireTaskLease/releaseTaskLease (state_machine.go:1069-1145)
  // is taken (reference behaviour: one Create and one Delete per LLM step); false = the `local`
  // provider's hand-off (the per-task mutex already serialises a Task's steps)
  bool emulate_lease = true;

 private:
  ObjectStore* store_;
  Recorder* recorder_;
  std::mutex mutex_map_lock_;
  std::map<std::string, std::unique_ptr<std::mutex>> task_mutexes_;
};

// TaskReconciler.Reconcile (task_controller.go:215-235): fetch the Task (NotFound is ignored),
// delegate to StateMachine.Process, persist nothing itself.
class TaskReconciler {
 public:
  TaskReconciler(ObjectStore* store, Recorder* recorder) : store_(store), sm_(store, recorder) {}
  Result Reconcile(const llmclient::Context& ctx, const std::string& taskName, const MCPToolsByServer& mcp,
                   acp_engine* engine, std::string* err);
  StateMachine& stateMachine() { return sm_; }

 private:
  ObjectStore* store_;
  StateMachine sm_;
};

}  // namespace task
}  // namespace acp
