// host/llmclient.h — C++ mirror of the reference's Go package acp/internal/llmclient, written
// above the C ABI because this image has no Go toolchain.  Same names, argument
// meaning and error behaviour as the Go interface so the parity tests read like the reference's:
//
//   type LLMClient interface { SendRequest(ctx, []acp.Message, []Tool) (*acp.Message, error) }
//                                                    acp/internal/llmclient/llm_client.go:11-14
//   type LLMRequestError struct { StatusCode int; Message string; Err error }        :18-30
//   func NewLLMClient(ctx, llm acp.LLM, apiKey string) (LLMClient, error)    factory.go:10-12
//   func NewLangchainClient(ctx, provider, apiKey, modelConfig)   langchaingo_client.go:27-80
//
// The Go files a maintainer would add (cgo shim + `case "local"`) are in INTEGRATION.md and
// integration/go/; they are thin wrappers over exactly the calls LocalClient makes here.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <vector>
#include "../json.h"
#include "acp_infer.h"

namespace acp {
namespace llmclient {

// acp.Message / MessageToolCall / ToolCallFunction (acp/api/v1alpha1/task_types.go:57-97)
struct ToolCallFunction { std::string Name, Arguments; };
struct MessageToolCall { std::string ID; ToolCallFunction Function; std::string Type; };
struct Message {
  std::string Role, Content;
  std::vector<MessageToolCall> ToolCalls;
  std::string ToolCallID, Name;
};
Json message_to_crd_json(const Message& m);        // CRD field names: toolCalls, toolCallId
bool message_from_crd_json(const Json& j, Message* m);

// llmclient.Tool (llm_client.go:33-50); ACPToolType is json:"-" and never leaves the process
struct ToolFunction { std::string Name, Description; Json Parameters; };
struct Tool { std::string Type; ToolFunction Function; std::string ACPToolType; };

// Go's `error` return: ok | *LLMRequestError | plain error
struct Error {
  bool is_request_error = false;  // errors.As(err, &llmErr) in handleLLMError
  int StatusCode = 0;
  std::string Message;            // LLMRequestError.Message, or the plain error text
  std::string Error_() const;     // what err.Error() prints
};

// context.Context: only cancellation matters on this path (ctx carries no deadline)
struct Context {
  std::shared_ptr<std::atomic<bool>> cancelled = std::make_shared<std::atomic<bool>>(false);
  bool done() const { return cancelled->load(); }
};

class LLMClient {
 public:
  virtual ~LLMClient() {}
  // returns true and fills *out, or false and fills *err
  virtual bool SendRequest(const Context& ctx, const std::vector<Message>& messages,
                           const std::vector<Tool>& tools, Message* out, Error* err) = 0;
};

// acp.BaseConfig subset the hot path passes on (langchaingo_client.go:33-39): Model, BaseURL
// acp.BaseConfig (acp/api/v1alpha1/llm_types.go:41-71).  Temperature / TopP are STRINGS in the CRD
// (pattern-validated decimals); MaxTokens / TopK are optional ints (0 = unset).  The reference's
// langchaingo path reads only Model and BaseURL (langchaingo_client.go:31-73); the local provider
// honours all of them.
struct BaseConfig {
  std::string Model, BaseURL, Temperature, TopP;
  int MaxTokens = 0, TopK = 0;
};
// BaseConfig from the JSON of LLM.spec.parameters (field names of the CRD)
BaseConfig base_config_from_json(const Json& params);

// wire conversion shared by both clients
std::string build_chat_request_json(const std::string& model, const std::vector<Message>& messages,
                                    const std::vector<Tool>& tools, int max_tokens,
                                    const Json* acp_ext, const BaseConfig* sampling = nullptr);
// convertFromLangchainResponse (langchaingo_client.go:208-282) on an OpenAI chat.completion body
bool convert_from_response_json(const std::string& body, Message* out, std::string* err);

// provider "local": a zero-cost handle onto the process-wide engine (created per reconcile like
// the reference does at state_machine.go:195, costs nothing).
class LocalClient : public LLMClient {
 public:
  LocalClient(acp_engine* engine, const BaseConfig& cfg) : engine_(engine), cfg_(cfg) {}
  bool SendRequest(const Context& ctx, const std::vector<Message>& messages,
                   const std::vector<Tool>& tools, Message* out, Error* err) override;
  // test / bench hook: extra "acp" block merged into the request (forced tokens, raw prompt ids)
  void set_extension(const Json& ext) { ext_ = ext; has_ext_ = true; }
  void set_max_tokens(int n) { cfg_.MaxTokens = n; }
  const std::string& last_response_json() const { return last_response_; }

 private:
  acp_engine* engine_;
  BaseConfig cfg_;
  Json ext_;
  bool has_ext_ = false;
  std::string last_response_;
};

// provider "openai" restated: HTTP/1.1 POST {BaseURL}/chat/completions over a TCP socket — what
// langchaingo's openai client does at langchaingo_client.go:102.  Used for the CPU baseline
// against the local stub completion server.
class HTTPClient : public LLMClient {
 public:
  HTTPClient(const std::string& api_key, const BaseConfig& cfg) : api_key_(api_key), cfg_(cfg) {}
  bool SendRequest(const Context& ctx, const std::vector<Message>& messages,
                   const std::vector<Tool>& tools, Message* out, Error* err) override;

 private:
  std::string api_key_;
  BaseConfig cfg_;
};

// NewLangchainClient's provider switch with the new `local` arm.  `engine` is the process-wide
// singleton (nullptr when the process did not start one).
std::unique_ptr<LLMClient> NewLLMClient(const std::string& provider, const std::string& api_key,
                                        const BaseConfig& cfg, acp_engine* engine, std::string* err);

}  // namespace llmclient
}  // namespace acp
