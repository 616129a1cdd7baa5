// local_provider.go would live at acp/internal/controller/llm/local_provider.go: how the LLM
// controller brings an LLM with provider "local" to Ready (SURVEY.md §8f rank 3).  Called from
// validateProviderConfig (integration/patches/provider-local.patch; reference
// acp/internal/controller/llm/state_machine.go:185-404).
//
// NOT COMPILED HERE (no Go toolchain in the build image).  What it does is exercised through the
// C ABI by tests/test_llmclient_gpu.py (engine start + a 1-token completion through LocalClient).
package llm

import (
	"context"
	"encoding/json"
	"fmt"

	acp "github.com/humanlayer/agentcontrolplane/acp/api/v1alpha1"
	"github.com/humanlayer/agentcontrolplane/acp/internal/inference"
	"github.com/humanlayer/agentcontrolplane/acp/internal/llmclient"
)

// localEngineConfig is the config_json of acp_infer_init: LLMSpec.Local verbatim (its JSON tags are
// the engine's keys) with the model name defaulted from parameters.model.
func localEngineConfig(llm *acp.LLM) (string, error) {
	cfg := map[string]interface{}{}
	if llm.Spec.Local != nil {
		raw, err := json.Marshal(llm.Spec.Local)
		if err != nil {
			return "", err
		}
		if err := json.Unmarshal(raw, &cfg); err != nil {
			return "", err
		}
	}
	if _, ok := cfg["model"]; !ok && llm.Spec.Parameters.Model != "" {
		cfg["model"] = llm.Spec.Parameters.Model
	}
	if _, ok := cfg["weights"]; !ok {
		return "", fmt.Errorf("local.weights is required for provider local (a checkpoint directory, or \"synthetic\")")
	}
	out, err := json.Marshal(cfg)
	return string(out), err
}

// validateLocalProvider replaces the reference's HTTP test call (state_machine.go:391-401): start
// the process-wide engine (idempotent: inference.Init runs acp_infer_init once) and run the same
// 1-token validation request, through the same LLMClient the Task controller will use.
func validateLocalProvider(ctx context.Context, llm *acp.LLM) error {
	cfgJSON, err := localEngineConfig(llm)
	if err != nil {
		return fmt.Errorf("failed to initialize local client: %w", err)
	}
	if _, err := inference.Init(cfgJSON); err != nil {
		return fmt.Errorf("failed to initialize local client: %w", err)
	}
	one := 1
	params := llm.Spec.Parameters
	params.MaxTokens = &one
	params.Temperature = "0"
	client, err := llmclient.NewLocalClient(params)
	if err != nil {
		return fmt.Errorf("failed to initialize local client: %w", err)
	}
	if _, err := client.SendRequest(ctx, []acp.Message{{Role: "user", Content: "test"}}, nil); err != nil {
		return fmt.Errorf("local API validation failed: %w", err)
	}
	return nil
}
