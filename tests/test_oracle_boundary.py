"""Pins oracle/boundary.py (the CPU restatement of the reference's Task -> LLM step) against the
reference's own golden behaviours G1..G11 (tests/golden/reference_goldens.json, each entry cites
the reference test it was taken from)."""
import copy
import json
import os
import re

from oracle import boundary as B

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
FX = G["fixtures"]


def _task(phase=B.PHASE_READY_FOR_LLM, window=None):
    return {"metadata": {"name": FX["task_name"], "namespace": "default", "uid": "uid-1"},
            "spec": {"agentRef": {"name": FX["agent_name"]}, "userMessage": FX["user_message"]},
            "status": {"phase": phase, "contextWindow": copy.deepcopy(window) if window is not None else [
                {"role": "system", "content": FX["system_prompt"]},
                {"role": "user", "content": FX["user_message"]}]}}


class _Client:
    def __init__(self, out=None, err=None):
        self.out, self.err, self.calls = out, err, []

    def send_request(self, messages, tools):
        self.calls.append((copy.deepcopy(messages), copy.deepcopy(tools)))
        if self.err is not None:
            raise self.err
        return copy.deepcopy(self.out)


def _reasons(rec):
    return [r for (_, r, _) in rec.events]


def test_G1_final_answer():
    g = G["G1_final_answer"]
    task, rec, tcs = _task(), B.Recorder(), []
    res, err = B.send_llm_request(task, [], _Client(g["llm_output"]), rec, tcs)
    e = g["expect"]
    assert err is None and res.requeue is False and res.requeue_after == 0
    st = task["status"]
    assert st["phase"] == e["phase"] and e["statusDetail_contains"] in st["statusDetail"]
    assert st["output"] == g["llm_output"]["content"]
    assert len(st["contextWindow"]) == e["window_len"] and st["contextWindow"][2]["role"] == e["last_role"]
    assert g["llm_output"]["content"] in st["contextWindow"][2]["content"]
    for r in e["events"]:
        assert r in _reasons(rec)
    assert tcs == []


def test_G2_tool_call_arguments_are_byte_identical():
    g = G["G2_tool_call"]
    tools = B.convert_mcp_tools([{"name": "fetch", "description": "fetch a url"}], "fetch")
    task, rec, tcs = _task(window=[]), B.Recorder(), []
    res, err = B.send_llm_request(task, tools, _Client(g["llm_output"]), rec, tcs)
    e = g["expect"]
    assert err is None and res.requeue_after == e["requeue_after"]
    st = task["status"]
    assert st["phase"] == e["phase"] and e["statusDetail_contains"] in st["statusDetail"]
    assert len(tcs) == e["n_toolcalls"]
    tc = tcs[0]
    assert tc["spec"]["toolRef"]["name"] == e["toolRef"]
    assert tc["spec"]["arguments"] == e["arguments"]          # byte-equal
    assert tc["spec"]["toolType"] == "MCP"
    assert tc["metadata"]["labels"]["acp.humanlayer.dev/toolcallrequest"] == st["toolCallRequestId"]
    assert tc["metadata"]["labels"]["acp.humanlayer.dev/task"] == FX["task_name"]
    # G10 / G11
    assert re.match(G["G10_id_format"]["regex"], st["toolCallRequestId"]) and len(st["toolCallRequestId"]) == 7
    assert tc["metadata"]["name"] == "%s-%s-tc-%02d" % (FX["task_name"], st["toolCallRequestId"], 1)
    for r in e["events"]:
        assert r in _reasons(rec)
    assert st["contextWindow"][-1]["toolCalls"][0]["id"] == "1"


def test_G3_generic_error_keeps_phase_and_returns_error():
    g = G["G3_generic_error"]
    task, rec = _task(), B.Recorder()
    res, err = B.send_llm_request(task, [], _Client(err=RuntimeError(g["error"])), rec, [])
    assert err is not None and res.requeue_after == 5
    st = task["status"]
    assert st["status"] == "Error" and st["phase"] == "ReadyForLLM" and st["error"] == g["error"]
    assert "LLMRequestFailed" in _reasons(rec)


def test_G4_4xx_is_terminal():
    g = G["G4_4xx_error"]
    task, rec = _task(), B.Recorder()
    e = B.LLMRequestError(g["status_code"], g["message"], RuntimeError("LLM API request failed"))
    res, err = B.send_llm_request(task, [], _Client(err=e), rec, [])
    assert err is None and res.is_zero()
    st = task["status"]
    assert st["status"] == "Error" and st["phase"] == "Failed"
    assert g["expect"]["error_contains"] in st["error"]
    assert "LLMRequestFailed4xx" in _reasons(rec)


def test_G5_tool_results_fold_back_in_list_order():
    g = G["G5_tool_results_fold_back"]
    window = [{"role": "system", "content": FX["system_prompt"]}, {"role": "user", "content": FX["user_message"]},
              {"role": "assistant", "content": "", "toolCalls": [
                  {"id": "1", "type": "function", "function": {"name": "fetch__fetch", "arguments": "{}"}},
                  {"id": "2", "type": "function", "function": {"name": "fetch__fetch", "arguments": "{}"}}]}]
    task, rec = _task(phase=B.PHASE_TOOL_CALLS_PENDING, window=window), B.Recorder()
    pending = copy.deepcopy(g["toolcalls"])
    pending[1]["status"]["status"] = "Running"
    assert B.check_tool_calls(task, pending, rec).requeue_after == 5     # :603-638
    res = B.check_tool_calls(task, g["toolcalls"], rec)
    assert res.requeue is True
    st = task["status"]
    assert st["phase"] == "ReadyForLLM" and len(st["contextWindow"]) == g["expect"]["window_len"]
    for tc, msg in zip(g["toolcalls"], st["contextWindow"][3:]):
        assert msg["role"] == "tool" and msg["content"] == tc["status"]["result"]
        assert msg["toolCallId"] == tc["spec"]["toolCallId"]


def test_G6_initial_window_and_validation():
    for c in G["G6_initial_window"]["cases"]:
        out = B.build_initial_context_window(c["contextWindow"], c["systemPrompt"], c["userMessage"])
        assert len(out) == c["len"] and out[0]["content"] == c["first"] and out[1]["content"] == c["second"]
        assert out[0]["role"] == "system"
    v = G["G6_initial_window"]["validation_errors"]
    assert B.validate_task_message_input("x", [{"role": "user", "content": "y"}]) == v["both"]
    assert B.validate_task_message_input("", []) == v["neither"]
    assert B.validate_task_message_input("", [{"role": "system", "content": "y"}]) == v["no_user"]
    assert B.validate_task_message_input("hello", []) is None
    assert B.get_user_message_preview("a" * 60, []) == "a" * 47 + "..."


def test_G7_delegate_tool():
    g = G["G7_delegate_tool"]
    t = B.convert_sub_agents([g["agent"]])[0]
    assert t["function"]["name"] == g["expect"]["name"] and t["acpToolType"] == g["expect"]["type"]
    assert t["function"]["description"] == g["agent"]["description"]
    assert B.build_tool_type_map([t]) == {g["expect"]["name"]: "DelegateToAgent"}


def test_G8_wire_fixtures_flatten_like_the_reference():
    g = G["G8_wire_fixtures"]
    m = B.convert_from_response(json.loads(g["content_body"]["body"]))
    assert m == {"role": "assistant", "content": g["content_body"]["expect_content"]}
    m = B.convert_from_response(json.loads(g["tool_body"]["body"]))
    e = g["tool_body"]["expect_tool"]
    assert m["content"] == "" and m["toolCalls"] == [
        {"id": e["id"], "type": e["type"], "function": {"name": e["name"], "arguments": e["arguments"]}}]
    # tool calls win over content; all choices are scanned (langchaingo_client.go:229-268)
    both = {"choices": [{"message": {"content": "hi"}},
                        {"message": {"content": "x", "tool_calls": [{"id": "a", "type": "function", "function": {"name": "n", "arguments": "{}"}}]}}]}
    m = B.convert_from_response(both)
    assert m["content"] == "" and len(m["toolCalls"]) == 1
    assert B.convert_from_response({"choices": []}) == {"role": "assistant", "content": ""}


def test_G9_request_body_shape():
    window = [{"role": "system", "content": "s"}, {"role": "user", "content": "u"},
              {"role": "assistant", "content": "", "toolCalls": [{"id": "1", "type": "function", "function": {"name": "f__g", "arguments": "{\"a\": 1}"}}]},
              {"role": "tool", "content": "42", "toolCallId": "1"}, {"role": "weird", "content": "w"}]
    tools = B.convert_mcp_tools([{"name": "g", "description": "d", "inputSchema": {"type": "object", "properties": {"a": {"type": "number"}}}}], "f")
    body = B.build_chat_request("m", window, tools)
    assert [m["role"] for m in body["messages"]] == ["system", "user", "assistant", "tool", "user"]
    assert body["messages"][2]["tool_calls"][0]["function"]["arguments"] == "{\"a\": 1}"
    assert body["messages"][3] == {"role": "tool", "content": "42", "tool_call_id": "1"}
    assert "acpToolType" not in json.dumps(body)
    assert body["tools"][0]["function"]["name"] == "f__g" and body["temperature"] == 0
    assert "tools" not in B.build_chat_request("m", window, [])


def test_G10_G11_formats():
    for n in (1, 6, 7, 8):
        s = B.generate_k8s_random_string(n)
        assert len(s) == n and B.K8S_RANDOM_RE.match(s)
    assert len(B.generate_k8s_random_string(0)) == 6 and len(B.generate_k8s_random_string(9)) == 6
    assert G["G11_toolcall_name"]["format"] % ("fetch-task", "2fe18aa", 1) == G["G11_toolcall_name"]["example"]


def test_contact_channel_tools():
    e = B.tool_from_contact_channel({"name": "ops", "spec": {"type": "email", "email": {"contextAboutUser": ""}}})
    assert e["function"]["name"] == "ops__human_contact_email" and e["function"]["description"] == "Contact a human via email"
    s = B.tool_from_contact_channel({"name": "ops", "spec": {"type": "slack", "slack": {"contextAboutChannelOrUser": "the ops channel"}}})
    assert s["function"]["name"] == "ops__human_contact_slack" and s["function"]["description"] == "the ops channel"
    assert s["acpToolType"] == "HumanContact"


def test_G12_G13_validate_task_and_agent():
    for gname, agent in (("G12_agent_missing", None),
                         ("G13_agent_not_ready", {"metadata": {"name": FX["agent_name"]}, "status": {"ready": False}})):
        exp = G[gname]["expect"]
        task = {"metadata": {"name": FX["task_name"]}, "status": {"phase": "Initializing", "status": "Pending", "error": "stale"}}
        rec = B.Recorder()
        res = B.validate_task_and_agent(task, agent, rec)
        st = task["status"]
        assert st["phase"] == exp["phase"] and st["status"] == exp["status"] and st["error"] == ""
        assert exp["statusDetail_contains"] in st["statusDetail"] and res.requeue_after == exp["requeue_after"]
        for frag in exp["events_containing"]:
            assert any(frag in r or frag in m for (_, r, m) in rec.events)
    ready = {"metadata": {"name": FX["agent_name"]}, "status": {"ready": True}}
    assert B.validate_task_and_agent({"status": {}}, ready, B.Recorder()) is None


def test_G14_contact_channel_goldens():
    g = G["G14_contact_channel_tools"]
    for case in g["cases"]:
        t = B.tool_from_contact_channel({"name": case["channel"]["metadata"]["name"], "spec": case["channel"]["spec"]})
        assert t["function"]["name"] == case["name"] and t["function"]["description"] == case["description"]
        assert t["function"]["parameters"] == g["parameters"] and t["acpToolType"] == g["acpToolType"]
