// local_provider.go would live at acp/internal/controller/task/local_provider.go (SURVEY.md §8f
// ranks 2 and 4; used by integration/patches/provider-local.patch).
//
// NOT COMPILED HERE (no Go toolchain).  The C++ twin of the step it guards is
// agentcontrolplane_b200/csrc/host/task.cc (StateMachine::sendLLMRequestFromCluster; the lease is
// the `emulate_lease` switch there, and bench.py's CPU baseline reports the API-write count).
package task

import (
	"context"

	"sigs.k8s.io/controller-runtime/pkg/client"

	acp "github.com/humanlayer/agentcontrolplane/acp/api/v1alpha1"
)

// taskUsesLocalProvider reports whether the Task's Agent points at an LLM with provider "local".
// Both GETs hit the controller-runtime cache (no API round trip).  Any lookup error answers false:
// the caller then takes the reference's path (Lease) and the error surfaces where the reference
// reports it (validateTaskAndAgent / getLLMAndCredentials).
func (sm *StateMachine) taskUsesLocalProvider(ctx context.Context, task *acp.Task) bool {
	var agent acp.Agent
	if err := sm.client.Get(ctx, client.ObjectKey{Namespace: task.Namespace, Name: task.Spec.AgentRef.Name}, &agent); err != nil {
		return false
	}
	var llm acp.LLM
	if err := sm.client.Get(ctx, client.ObjectKey{Namespace: task.Namespace, Name: agent.Spec.LLMRef.Name}, &llm); err != nil {
		return false
	}
	return llm.Spec.Provider == "local"
}
