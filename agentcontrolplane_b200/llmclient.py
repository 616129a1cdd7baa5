"""Python face of the reference's `llmclient` interface for provider `local`
(acp/internal/llmclient/llm_client.go:11-30): same names, argument meaning and error behaviour,
implemented as calls through the C ABI (no compute in Python).

    class LLMClient:  send_request(messages, tools) -> Message      (SendRequest)
    class LLMRequestError(status_code, message)                      (LLMRequestError)
    new_llm_client(provider, api_key, model_config, engine)          (NewLangchainClient switch)

Messages and tools are CRD-shaped dicts (acp/api/v1alpha1/task_types.go:57-97): role, content,
toolCalls[{id, type, function{name, arguments}}], toolCallId.
"""
from __future__ import annotations

from typing import Any

from . import host
from .engine import Engine


class LLMRequestError(Exception):
    """4xx from the provider: terminal for the Task (handleLLMError, state_machine.go:733-790)."""

    def __init__(self, status_code: int, message: str):
        self.status_code, self.message = status_code, message
        super().__init__(f"LLM request failed with status {status_code}: {message}")


class LLMClient:
    def send_request(self, messages: list[dict], tools: list[dict]) -> dict:  # pragma: no cover
        raise NotImplementedError


class LocalLLMClient(LLMClient):
    """`provider: local` — a zero-cost handle on the process-wide engine."""

    def __init__(self, engine: Engine, model: str, max_tokens: int = 0, acp_ext: dict | None = None):
        self.engine, self.model, self.max_tokens, self.acp_ext = engine, model, max_tokens, acp_ext

    def send_request(self, messages: list[dict], tools: list[dict]) -> dict:
        body: dict[str, Any] = host.build_chat_request(self.model, messages, tools)
        if self.max_tokens:
            body["max_tokens"] = self.max_tokens
        if self.acp_ext:
            body["acp"] = self.acp_ext
        status, resp = self.engine.complete(body)
        self.last_response = resp
        if status != 200:
            msg = (resp.get("error") or {}).get("message", str(resp))
            if 400 <= status < 500 and status != 499:
                raise LLMRequestError(status, msg)
            raise RuntimeError("model API call failed: " + msg)
        return host.convert_response(resp)


def new_llm_client(provider: str, api_key: str, model_config: dict, engine: Engine | None = None) -> LLMClient:
    """NewLangchainClient's provider switch (langchaingo_client.go:31-73) with the `local` arm."""
    if provider == "local":
        if engine is None:
            raise RuntimeError("failed to initialize local client: engine not initialised")
        return LocalLLMClient(engine, model_config.get("model", ""), int(model_config.get("maxTokens", 0) or 0))
    raise ValueError(f"unsupported provider: {provider}. Supported providers are: openai, anthropic, "
                     "mistral, google, vertex, local")
