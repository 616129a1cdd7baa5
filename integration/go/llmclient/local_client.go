// local_client.go would live at acp/internal/llmclient/local_client.go: the `local` provider
// behind the UNCHANGED LLMClient interface (acp/internal/llmclient/llm_client.go:11-14).
//
// NOT COMPILED HERE (no Go toolchain in the build image); the C++ twin that is compiled and
// tested is agentcontrolplane_b200/csrc/host/llmclient.cc.
package llmclient

import (
	"context"
	"encoding/json"
	"fmt"
	"strconv"

	acp "github.com/humanlayer/agentcontrolplane/acp/api/v1alpha1"
	"github.com/humanlayer/agentcontrolplane/acp/internal/inference"
)

var _ LLMClient = &LocalClient{}

// LocalClient is a zero-cost handle on the process-wide engine: the Task controller builds a new
// client on every reconcile (controller/task/state_machine.go:195), so construction must be O(1).
type LocalClient struct {
	engine *inference.Engine
	cfg    acp.BaseConfig
}

func NewLocalClient(cfg acp.BaseConfig) (LLMClient, error) {
	e := inference.Get()
	if e == nil {
		return nil, fmt.Errorf("failed to initialize local client: engine not initialised")
	}
	return &LocalClient{engine: e, cfg: cfg}, nil
}

// wire types: the OpenAI chat-completions body langchaingo's openai provider would have POSTed
type wireFunction struct {
	Name      string `json:"name"`
	Arguments string `json:"arguments"`
}
type wireToolCall struct {
	ID       string       `json:"id"`
	Type     string       `json:"type"`
	Function wireFunction `json:"function"`
}
type wireMessage struct {
	Role       string         `json:"role"`
	Content    string         `json:"content"`
	ToolCalls  []wireToolCall `json:"tool_calls,omitempty"`
	ToolCallID string         `json:"tool_call_id,omitempty"`
}
type wireRequest struct {
	Model       string        `json:"model"`
	Messages    []wireMessage `json:"messages"`
	Temperature float64       `json:"temperature"` // no omitempty, like langchaingo's ChatRequest
	TopP        *float64      `json:"top_p,omitempty"`
	TopK        *int          `json:"top_k,omitempty"`
	MaxTokens   int           `json:"max_tokens,omitempty"`
	Tools       []Tool        `json:"tools,omitempty"` // ACPToolType is json:"-" (llm_client.go:38)
}
type wireResponse struct {
	Choices []struct {
		Message struct {
			Content   *string        `json:"content"`
			ToolCalls []wireToolCall `json:"tool_calls"`
		} `json:"message"`
	} `json:"choices"`
	Error *struct {
		Message string `json:"message"`
	} `json:"error"`
}

// SendRequest implements LLMClient.  Same conversions as convertToLangchainMessages /
// convertFromLangchainResponse (langchaingo_client.go:118-185, 208-282).
func (c *LocalClient) SendRequest(ctx context.Context, messages []acp.Message, tools []Tool) (*acp.Message, error) {
	// LLM.spec.parameters (acp/api/v1alpha1/llm_types.go:41-71).  The reference's langchaingo path
	// reads only Model and BaseURL and therefore always sends temperature 0; the local provider
	// honours temperature / topP / topK / maxTokens when they are set (strings in the CRD).
	req := wireRequest{Model: c.cfg.Model, Temperature: 0, Tools: tools}
	if c.cfg.Temperature != "" {
		t, err := strconv.ParseFloat(c.cfg.Temperature, 64)
		if err != nil {
			return nil, &LLMRequestError{StatusCode: 400, Message: "parameters.temperature: " + err.Error()}
		}
		req.Temperature = t
	}
	if c.cfg.TopP != "" {
		p, err := strconv.ParseFloat(c.cfg.TopP, 64)
		if err != nil {
			return nil, &LLMRequestError{StatusCode: 400, Message: "parameters.topP: " + err.Error()}
		}
		req.TopP = &p
	}
	if c.cfg.TopK != nil {
		req.TopK = c.cfg.TopK
	}
	if c.cfg.MaxTokens != nil {
		req.MaxTokens = *c.cfg.MaxTokens
	}
	for _, m := range messages {
		role := m.Role
		switch role {
		case "system", "user", "assistant", "tool":
		default:
			role = "user" // langchaingo_client.go:136-137
		}
		wm := wireMessage{Role: role, Content: m.Content, ToolCallID: m.ToolCallID}
		if !(role == "tool" && m.ToolCallID != "") {
			for _, tc := range m.ToolCalls {
				wm.ToolCalls = append(wm.ToolCalls, wireToolCall{ID: tc.ID, Type: tc.Type,
					Function: wireFunction{Name: tc.Function.Name, Arguments: tc.Function.Arguments}})
			}
		}
		req.Messages = append(req.Messages, wm)
	}
	body, err := json.Marshal(req)
	if err != nil {
		return nil, fmt.Errorf("model API call failed: %w", err)
	}
	status, respBytes, err := c.engine.Complete(ctx, body)
	if err != nil {
		return nil, fmt.Errorf("model API call failed: %w", err)
	}
	var resp wireResponse
	if err := json.Unmarshal(respBytes, &resp); err != nil {
		return nil, fmt.Errorf("model API call failed: %w", err)
	}
	if status != 200 {
		msg := string(respBytes)
		if resp.Error != nil {
			msg = resp.Error.Message
		}
		if status >= 400 && status < 500 && status != 499 {
			// typed: handleLLMError marks the Task Failed (state_machine.go:738-756)
			return nil, &LLMRequestError{StatusCode: status, Message: msg}
		}
		return nil, fmt.Errorf("model API call failed: %s", msg)
	}
	out := &acp.Message{Role: "assistant"}
	var content string
	hasContent := false
	for _, ch := range resp.Choices {
		if !hasContent && ch.Message.Content != nil && *ch.Message.Content != "" {
			content, hasContent = *ch.Message.Content, true
		}
		for _, tc := range ch.Message.ToolCalls {
			out.ToolCalls = append(out.ToolCalls, acp.MessageToolCall{ID: tc.ID, Type: tc.Type,
				Function: acp.ToolCallFunction{Name: tc.Function.Name, Arguments: tc.Function.Arguments}})
		}
	}
	if len(out.ToolCalls) == 0 && hasContent {
		out.Content = content // tool calls win and clear content (langchaingo_client.go:255-267)
	}
	return out, nil
}
