"""CPU restatement of the reference's boundary logic for the Task -> LLM step (pure Python).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference lines it
follows (paths relative to /root/reference).  Pinned by the reference's own golden behaviours
G1..G14 (G12-G14 added in round 2 for a5/a6/a8) in tests/test_oracle_boundary.py.

Data model: plain dicts shaped like the CRD JSON (acp/api/v1alpha1/task_types.go:57-97):
    Message  = {"role", "content", "toolCalls": [{"id","type","function":{"name","arguments"}}],
                "toolCallId", "name"}                      (optional keys omitted when empty)
    Tool     = {"type", "function": {"name","description","parameters"}, "acpToolType"}
"""
from __future__ import annotations

import copy
import re
import secrets
from dataclasses import dataclass, field

DEFAULT_REQUEUE_DELAY = 5.0  # acp/internal/controller/task/task_controller.go:23

# phases / statuses (acp/api/v1alpha1/task_types.go:160-200)
PHASE_INITIALIZING = "Initializing"
PHASE_PENDING = "Pending"
PHASE_READY_FOR_LLM = "ReadyForLLM"
PHASE_TOOL_CALLS_PENDING = "ToolCallsPending"
PHASE_FINAL_ANSWER = "FinalAnswer"
PHASE_FAILED = "Failed"
STATUS_READY, STATUS_ERROR, STATUS_PENDING = "Ready", "Error", "Pending"

TOOL_TYPE_MCP = "MCP"
TOOL_TYPE_HUMAN_CONTACT = "HumanContact"
TOOL_TYPE_DELEGATE = "DelegateToAgent"

VALID_ROLES = {"system", "user", "assistant", "tool"}


class LLMRequestError(Exception):
    """acp/internal/llmclient/llm_client.go:18-30"""

    def __init__(self, status_code: int, message: str, err: Exception | None = None):
        self.status_code, self.message, self.err = status_code, message, err
        super().__init__(f"LLM request failed with status {status_code}: {message}")


@dataclass
class Result:
    """ctrl.Result"""
    requeue: bool = False
    requeue_after: float = 0.0

    def is_zero(self) -> bool:
        return not self.requeue and self.requeue_after == 0.0


@dataclass
class Recorder:
    events: list = field(default_factory=list)

    def event(self, etype: str, reason: str, message: str):
        self.events.append((etype, reason, message))


# ------------------------------------------------------------------------------------------
# llmclient: message / tool conversion and response flattening
# ------------------------------------------------------------------------------------------
def convert_to_openai_messages(messages: list[dict]) -> list[dict]:
    """convertToLangchainMessages (acp/internal/llmclient/langchaingo_client.go:118-185) composed
    with langchaingo v0.1.13's openai wire mapping (un-vendored; restated from its published
    behaviour): System->"system", Human->"user", AI->"assistant", Tool->"tool"; an AI message's
    ToolCall parts become `tool_calls`; a Tool message's ToolCallResponse becomes
    {role: tool, tool_call_id, content}.  Unknown roles map to Human (:136-137)."""
    out = []
    for m in messages:
        role = m.get("role", "")
        wire_role = role if role in ("system", "user", "assistant", "tool") else "user"
        o: dict = {"role": wire_role}
        content = m.get("content", "")
        tool_calls = m.get("toolCalls") or []
        tcid = m.get("toolCallId", "")
        if wire_role == "tool" and tcid:
            o["content"] = content          # :164-171 — only the ToolCallResponse part survives
            o["tool_call_id"] = tcid
        else:
            o["content"] = content
            if tool_calls:
                o["tool_calls"] = [{"id": tc.get("id", ""), "type": tc.get("type", ""),
                                    "function": {"name": tc["function"]["name"],
                                                 "arguments": tc["function"]["arguments"]}}
                                   for tc in tool_calls]
            if tcid:
                o["tool_call_id"] = tcid
        out.append(o)
    return out


def convert_to_openai_tools(tools: list[dict]) -> list[dict]:
    """convertToLangchainTools (langchaingo_client.go:188-203); ACPToolType is json:"-"
    (llm_client.go:38) and never reaches the wire."""
    return [{"type": t["type"], "function": {"name": t["function"]["name"],
                                             "description": t["function"].get("description", ""),
                                             "parameters": t["function"].get("parameters", {})}}
            for t in tools]


def build_chat_request(model: str, messages: list[dict], tools: list[dict]) -> dict:
    """The body SendRequest puts on the wire (langchaingo_client.go:83-102): model + messages,
    tools only when present (:93-99), temperature 0 (ChatRequest.Temperature has no omitempty and
    ACP sets no option — SURVEY.md §8c)."""
    body = {"model": model, "messages": convert_to_openai_messages(messages), "temperature": 0}
    if tools:
        body["tools"] = convert_to_openai_tools(tools)
    return body


def convert_from_response(response: dict) -> dict:
    """convertFromLangchainResponse (langchaingo_client.go:208-282) applied to an OpenAI
    chat.completion body: first non-empty content across choices; ALL tool calls across ALL
    choices; tool calls win and clear content; no choices -> empty assistant message."""
    msg = {"role": "assistant", "content": ""}
    choices = response.get("choices") or []
    if not choices:
        return msg
    tool_calls, content, has_content = [], "", False
    for ch in choices:
        m = ch.get("message") or {}
        c = m.get("content") or ""
        if not has_content and c != "":
            content, has_content = c, True
        for tc in m.get("tool_calls") or []:
            fn = tc.get("function") or {}
            tool_calls.append({"id": tc.get("id", ""), "type": tc.get("type", ""),
                               "function": {"name": fn.get("name", ""),
                                            "arguments": fn.get("arguments", "")}})
    if tool_calls:
        msg["toolCalls"] = tool_calls
        msg["content"] = ""
        return msg
    if has_content:
        msg["content"] = content
    return msg


# ------------------------------------------------------------------------------------------
# tool list construction
# ------------------------------------------------------------------------------------------
def convert_mcp_tools(mcp_tools: list[dict], server_name: str) -> list[dict]:
    """adapters.ConvertMCPToolsToLLMClientTools (acp/internal/adapters/mcp_adapter.go:12-51)"""
    out = []
    for t in mcp_tools:
        params = t.get("inputSchema")
        if not isinstance(params, dict):
            params = {"type": "object", "properties": {}}
        out.append({"type": "function",
                    "function": {"name": f"{server_name}__{t['name']}",
                                 "description": t.get("description", ""), "parameters": params},
                    "acpToolType": TOOL_TYPE_MCP})
    return out


def tool_from_contact_channel(channel: dict) -> dict:
    """llmclient.ToolFromContactChannel (acp/internal/llmclient/llm_client.go:53-99)"""
    params = {"type": "object", "properties": {"message": {"type": "string"}}, "required": ["message"]}
    ctype, name = channel["spec"]["type"], channel["name"]
    if ctype == "email":
        tname = f"{name}__human_contact_email"
        desc = (channel["spec"].get("email") or {}).get("contextAboutUser", "") or "Contact a human via email"
    elif ctype == "slack":
        tname = f"{name}__human_contact_slack"
        desc = (channel["spec"].get("slack") or {}).get("contextAboutChannelOrUser", "") or "Contact a human via Slack"
    else:
        tname = f"{name}__human_contact"
        desc = f"Contact a human via {ctype} channel"
    return {"type": "function", "function": {"name": tname, "description": desc, "parameters": params},
            "acpToolType": TOOL_TYPE_HUMAN_CONTACT}


def convert_sub_agents(agents: list[dict]) -> list[dict]:
    """defaultToolAdapter.ConvertSubAgents (controller/task/task_controller.go:94-117)"""
    return [{"type": "function",
             "function": {"name": "delegate_to_agent__" + a["name"],
                          "description": a.get("description", ""),
                          "parameters": {"type": "object",
                                         "properties": {"message": {"type": "string"}},
                                         "required": ["message"]}},
             "acpToolType": TOOL_TYPE_DELEGATE} for a in agents]


def build_tool_type_map(tools: list[dict]) -> dict:
    """buildToolTypeMap (controller/task/task_helpers.go:48-54)"""
    return {t["function"]["name"]: t.get("acpToolType", "") for t in tools}


# ------------------------------------------------------------------------------------------
# validation helpers
# ------------------------------------------------------------------------------------------
def validate_task_message_input(user_message: str, context_window: list[dict]) -> str | None:
    """validation.ValidateTaskMessageInput (acp/internal/validation/task_validation.go:16-40);
    returns the error string or None."""
    if user_message != "" and len(context_window) > 0:
        return "only one of userMessage or contextWindow can be provided"
    if user_message == "" and len(context_window) == 0:
        return "one of userMessage or contextWindow must be provided"
    if context_window:
        has_user = False
        for m in context_window:
            if m.get("role") not in VALID_ROLES:
                return f"invalid role in contextWindow: {m.get('role')}"
            if m.get("role") == "user":
                has_user = True
        if not has_user:
            return "contextWindow must contain at least one user message"
    return None


def get_user_message_preview(user_message: str, context_window: list[dict]) -> str:
    """validation.GetUserMessagePreview (task_validation.go:43-58) — byte-length truncation."""
    preview = ""
    if user_message != "":
        preview = user_message
    elif context_window:
        for m in reversed(context_window):
            if m.get("role") == "user":
                preview = m.get("content", "")
                break
    raw = preview.encode()
    if len(raw) > 50:
        preview = raw[:47].decode(errors="ignore") + "..."
    return preview


def generate_k8s_random_string(n: int) -> str:
    """validation.GenerateK8sRandomString (task_validation.go:61-87)"""
    if n < 1 or n > 8:
        n = 6
    letters = "abcdefghijklmnopqrstuvwxyz"
    alnum = letters + "0123456789"
    return secrets.choice(letters) + "".join(secrets.choice(alnum) for _ in range(n - 1))


K8S_RANDOM_RE = re.compile(r"^[a-z][a-z0-9]*$")


def build_initial_context_window(context_window: list[dict], system_prompt: str,
                                 user_message: str) -> list[dict]:
    """buildInitialContextWindow (controller/task/task_helpers.go:13-44)"""
    if context_window:
        out = [copy.deepcopy(m) for m in context_window]
        if not any(m.get("role") == "system" for m in out):
            out.insert(0, {"role": "system", "content": system_prompt})
        return out
    return [{"role": "system", "content": system_prompt}, {"role": "user", "content": user_message}]


# ------------------------------------------------------------------------------------------
# Task state machine: the LLM step and its neighbours
# ------------------------------------------------------------------------------------------
def process_llm_response(output: dict, task: dict, tools: list[dict], recorder: Recorder,
                         toolcalls_out: list, id_gen=generate_k8s_random_string) -> Result:
    """processLLMResponse + createToolCalls (controller/task/state_machine.go:605-731), minus the
    v1beta3 respond_to_human branch (:610-612) and the HumanLayer notify (:637-639).
    Mutates task["status"]; appends created ToolCall objects to toolcalls_out."""
    st = task["status"]
    if output.get("content", "") != "":
        prev_phase = st.get("phase", "")
        st["output"] = output["content"]
        st["phase"] = PHASE_FINAL_ANSWER
        st["ready"] = True
        st.setdefault("contextWindow", []).append({"role": "assistant", "content": output["content"]})
        st["status"] = STATUS_READY
        st["statusDetail"] = "LLM final response received"
        st["error"] = ""
        if prev_phase != PHASE_FINAL_ANSWER:
            recorder.event("Normal", "LLMFinalAnswer", "LLM response received successfully")
        return Result()
    req_id = id_gen(7)
    st["output"] = ""
    st["phase"] = PHASE_TOOL_CALLS_PENDING
    st["toolCallRequestId"] = req_id
    st.setdefault("contextWindow", []).append({"role": "assistant", "content": "",
                                               "toolCalls": copy.deepcopy(output.get("toolCalls") or [])})
    st["ready"] = True
    st["status"] = STATUS_READY
    st["statusDetail"] = "LLM response received, tool calls pending"
    st["error"] = ""
    recorder.event("Normal", "ToolCallsPending", "LLM response received, tool calls pending")
    return create_tool_calls(task, output.get("toolCalls") or [], tools, recorder, toolcalls_out)


def create_tool_calls(task: dict, tool_calls: list[dict], tools: list[dict], recorder: Recorder,
                      toolcalls_out: list) -> Result:
    """createToolCalls (state_machine.go:676-731)"""
    st = task["status"]
    if not st.get("toolCallRequestId"):
        raise RuntimeError("no ToolCallRequestID found in statusUpdate, cannot create tool calls")
    type_map = build_tool_type_map(tools)
    name = task["metadata"]["name"]
    for i, tc in enumerate(tool_calls):
        new_name = "%s-%s-tc-%02d" % (name, st["toolCallRequestId"], i + 1)
        toolcalls_out.append({
            "metadata": {"name": new_name, "namespace": task["metadata"].get("namespace", "default"),
                         "labels": {"acp.humanlayer.dev/task": name,
                                    "acp.humanlayer.dev/toolcallrequest": st["toolCallRequestId"]},
                         "ownerReferences": [{"apiVersion": "acp.humanlayer.dev/v1alpha1", "kind": "Task",
                                              "name": name, "uid": task["metadata"].get("uid", ""),
                                              "controller": True}]},
            "spec": {"toolCallId": tc.get("id", ""), "taskRef": {"name": name},
                     "toolRef": {"name": tc["function"]["name"]},
                     "toolType": type_map.get(tc["function"]["name"], ""),
                     "arguments": tc["function"]["arguments"]},
            "status": {}})
        recorder.event("Normal", "ToolCallCreated", "Created ToolCall " + new_name)
    return Result(requeue_after=DEFAULT_REQUEUE_DELAY)


def handle_llm_error(task: dict, err: Exception, recorder: Recorder):
    """handleLLMError (state_machine.go:733-790).  Returns (Result, error-or-None)."""
    st = task["status"]
    is4xx = isinstance(err, LLMRequestError) and 400 <= err.status_code < 500
    st["ready"] = False
    st["status"] = STATUS_ERROR
    st["statusDetail"] = f"LLM request failed: {err}"
    st["error"] = str(err)
    if is4xx:
        st["phase"] = PHASE_FAILED
        recorder.event("Warning", "LLMRequestFailed4xx",
                       f"LLM request failed with status {err.status_code}: {err.message}")
        return Result(), None
    recorder.event("Warning", "LLMRequestFailed", str(err))
    return Result(requeue_after=DEFAULT_REQUEUE_DELAY), err


def send_llm_request(task: dict, tools: list[dict], llm_client, recorder: Recorder,
                     toolcalls_out: list, id_gen=generate_k8s_random_string):
    """sendLLMRequest (state_machine.go:162-288) with the Kubernetes plumbing (mutex, lease, Agent /
    LLM / Secret lookups, OTel span) removed: event + status detail, SendRequest, then
    processLLMResponse / handleLLMError.  `llm_client.send_request(messages, tools)` returns a
    Message dict or raises."""
    st = task["status"]
    if st.get("phase") != PHASE_READY_FOR_LLM or st.get("statusDetail") != "Sending request to LLM":
        recorder.event("Normal", "SendingContextWindowToLLM", "Sending context window to LLM")
        st["statusDetail"] = "Sending request to LLM"
    try:
        output = llm_client.send_request(st.get("contextWindow", []), tools)
    except Exception as e:  # noqa: BLE001 — mirrors `if err != nil`
        return handle_llm_error(task, e, recorder)
    return process_llm_response(output, task, tools, recorder, toolcalls_out, id_gen), None


def validate_task_and_agent(task: dict, agent: dict | None, recorder: Recorder) -> Result | None:
    """validateTaskAndAgent (state_machine.go:379-424); `agent` is the Agent CR or None when the GET
    returned NotFound.  Returns the (non-zero) Result when the Task has to wait, None when it may go on."""
    st = task["status"]
    if agent is None:
        st.update(ready=False, status=STATUS_PENDING, phase=PHASE_PENDING,
                  statusDetail="Waiting for Agent to exist", error="")
        recorder.event("Normal", "Waiting", "Waiting for Agent to exist")
        return Result(requeue_after=DEFAULT_REQUEUE_DELAY)
    if not agent.get("status", {}).get("ready", False):
        msg = f"Waiting for agent \"{agent['metadata']['name']}\" to become ready"
        st.update(ready=False, status=STATUS_PENDING, phase=PHASE_PENDING, statusDetail=msg, error="")
        recorder.event("Normal", "Waiting", msg)
        return Result(requeue_after=DEFAULT_REQUEUE_DELAY)
    return None


def get_llm_and_credentials(task: dict, llm: dict | None, llm_name: str, secrets_by_name: dict,
                            recorder: Recorder):
    """getLLMAndCredentials (state_machine.go:480-538) -> (api_key, error string or None).
    The reference dereferences llm.Spec.APIKeyFrom unconditionally (:504); with the `local`
    provider and no apiKeyFrom that lookup is skipped and the key is ""."""
    st = task["status"]

    def fail(detail, error, reason):
        st.update(ready=False, status=STATUS_ERROR, phase=PHASE_FAILED, statusDetail=detail, error=error)
        recorder.event("Warning", reason, error)
        return None, error

    if llm is None:
        e = f'llms.acp.humanlayer.dev "{llm_name}" not found'
        return fail(f"Failed to get LLM: {e}", e, "LLMFetchFailed")
    spec = llm.get("spec", {})
    key_from = spec.get("apiKeyFrom")
    if spec.get("provider") == "local" and not key_from:
        return "", None
    ref = (key_from or {}).get("secretKeyRef", {})
    secret = secrets_by_name.get(ref.get("name", ""))
    if secret is None:
        e = f'secrets "{ref.get("name", "")}" not found'
        return fail(f"Failed to get API key secret: {e}", e, "APIKeySecretFetchFailed")
    api_key = secret.get("data", {}).get(ref.get("key", ""), "")
    if api_key == "":
        return fail("API key is empty", "API key is empty", "EmptyAPIKey")
    return api_key, None


def collect_tools(agent: dict, mcp_tools_by_server: dict, channels_by_name: dict, agents_by_name: dict) -> list[dict]:
    """collectTools (state_machine.go:540-583): MCP server tools (in agent.spec.mcpServers order; servers
    the manager does not know are skipped), then the agent's valid contact channels, then its sub-agents
    (objects that cannot be fetched are skipped, like the `continue` at :563 / :576)."""
    tools: list[dict] = []
    for ref in agent.get("spec", {}).get("mcpServers", []):
        if ref["name"] in mcp_tools_by_server:
            tools += convert_mcp_tools(mcp_tools_by_server[ref["name"]], ref["name"])
    for ref in agent.get("status", {}).get("validHumanContactChannels", []):
        ch = channels_by_name.get(ref["name"])
        if ch is not None:
            tools.append(tool_from_contact_channel({"name": ch["metadata"]["name"], "spec": ch["spec"]}))
    subs = []
    for ref in agent.get("spec", {}).get("subAgents", []):
        sub = agents_by_name.get(ref["name"])
        if sub is not None:
            subs.append({"name": sub["metadata"]["name"], "description": sub.get("spec", {}).get("description", "")})
    return tools + convert_sub_agents(subs)


def check_tool_calls(task: dict, toolcalls: list[dict], recorder: Recorder) -> Result:
    """checkToolCalls (state_machine.go:291-341): `toolcalls` is the label-selected list in list
    order."""
    st = task["status"]
    for tc in toolcalls:
        if tc.get("status", {}).get("status") not in ("Succeeded", "Error"):
            return Result(requeue_after=DEFAULT_REQUEUE_DELAY)
    for tc in toolcalls:
        st.setdefault("contextWindow", []).append({"role": "tool",
                                                   "content": tc.get("status", {}).get("result", ""),
                                                   "toolCallId": tc["spec"]["toolCallId"]})
    st["phase"] = PHASE_READY_FOR_LLM
    st["status"] = STATUS_READY
    st["statusDetail"] = "All tool calls completed, ready to send tool results to LLM"
    st["error"] = ""
    recorder.event("Normal", "AllToolCallsCompleted", "All tool calls completed")
    return Result(requeue=True)
