"""Python restatement of the text side of the local provider: chat template,
synthetic byte-level tokenizer, tool-call extraction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference has none of this — tokenizer,
template and tool-call grammar all live behind the hosted provider
(acp/internal/llmclient/langchaingo_client.go:102), so this is "parity unpinned" against the
reference: it is pinned instead against the published Llama-3 chat format
(header / eot special tokens at their real ids, Llama-3.1 JSON tool calling) and is the single
written definition both the C++ engine (csrc/chat.cc) and the tests follow.
"""
from __future__ import annotations

import json

BOT, EOT_TEXT, START_H, END_H, EOM, EOT, PYTAG = 128000, 128001, 128006, 128007, 128008, 128009, 128010
STOP_TOKENS = (EOT_TEXT, EOM, EOT)

TOOL_PREAMBLE = ("Given the following functions, please respond with a JSON for a function call with its "
                 "proper arguments that best answers the given prompt.\n\nRespond in the format {\"name\": "
                 "function name, \"parameters\": dictionary of argument name and its value}. Do not use "
                 "variables.\n\n")


def _dump(o) -> str:
    # compact separators + raw UTF-8, like csrc/json.h
    return json.dumps(o, separators=(",", ":"), ensure_ascii=False)


def encode_text(text: str | bytes) -> list[int]:
    raw = text if isinstance(text, bytes) else text.encode()
    return list(raw)


def decode_tokens(ids) -> bytes:
    out = bytearray()
    for t in ids:
        if t < 0:
            continue
        if t < 256:
            out.append(t)
        elif t < BOT:
            out += b" "
            n = t - 256
            while True:
                out.append(ord("a") + n % 26)
                n //= 26
                if n == 0:
                    break
    return bytes(out)


def tool_json(t: dict) -> str:
    fn = t["function"]
    return _dump({"type": t.get("type") or "function",
                  "function": {"name": fn["name"], "description": fn.get("description", ""),
                               "parameters": fn.get("parameters") if fn.get("parameters") is not None else {}}})


def render(messages: list[dict], tools: list[dict]):
    """messages / tools in OpenAI wire form.  Returns (token ids, spelled-out text)."""
    ids, text = [], []

    def special(i, s):
        ids.append(i)
        text.append(s)

    def raw(s):
        ids.extend(encode_text(s))
        text.append(s)

    def header(role):
        special(START_H, "<|start_header_id|>")
        raw(role)
        special(END_H, "<|end_header_id|>")
        raw("\n\n")

    special(BOT, "<|begin_of_text|>")
    idx, sys = 0, ""
    if messages and messages[0].get("role") == "system":
        sys, idx = messages[0].get("content") or "", 1
    if sys or tools:
        header("system")
        raw(("Environment: ipython\n\n" if tools else "") + sys)
        special(EOT, "<|eot_id|>")
    first_user = True
    for m in messages[idx:]:
        role, content = m.get("role"), m.get("content") or ""
        if role == "assistant":
            header("assistant")
            tcs = m.get("tool_calls") or []
            if tcs:
                raw("\n".join('{"name": ' + json.dumps(tc["function"]["name"], ensure_ascii=False) +
                              ', "parameters": ' + (tc["function"].get("arguments") or "{}") + "}" for tc in tcs))
            else:
                raw(content)
        elif role == "tool":
            header("ipython")
            raw(content)
        elif role == "system":
            header("system")
            raw(content)
        else:
            header("user")
            if tools and first_user:
                raw(TOOL_PREAMBLE + "".join(tool_json(t) + "\n\n" for t in tools) + content)
            else:
                raw(content)
            first_user = False
        special(EOT, "<|eot_id|>")
    header("assistant")
    return ids, "".join(text)


def parse_completion(text: str, tools: list[dict], prefix: str = "call_") -> dict:
    """One JSON object per tool call: {"name": <known tool>, "parameters": {...}}; `arguments` is
    the verbatim substring that spells the parameters object; anything else is content."""
    names = {t["function"]["name"] for t in tools or []}
    if names:
        dec, pos, calls, ok = json.JSONDecoder(), 0, [], True
        n = len(text)
        while True:
            while pos < n and text[pos] in " \n\t\r":
                pos += 1
            if pos >= n:
                break
            if text[pos] != "{":
                ok = False
                break
            try:
                obj, end = dec.raw_decode(text, pos)
            except ValueError:
                ok = False
                break
            key = "parameters" if isinstance(obj, dict) and "parameters" in obj else "arguments"
            if not isinstance(obj, dict) or obj.get("name") not in names or not isinstance(obj.get(key), dict):
                ok = False
                break
            span = _member_span(text, pos, key)
            calls.append({"id": f"{prefix}{len(calls)}", "type": "function",
                          "function": {"name": obj["name"], "arguments": span}})
            pos = end
        if ok and calls:
            return {"tool_calls": calls}
    return {"content": text}


def _member_span(text: str, obj_start: int, key: str) -> str:
    """Verbatim source text of top-level member `key` of the JSON object starting at obj_start."""
    dec = json.JSONDecoder()
    i = obj_start + 1
    last = ""
    while True:
        while text[i] in " \n\t\r,":
            i += 1
        if text[i] == "}":
            return last
        k, i = dec.raw_decode(text, i)
        while text[i] in " \n\t\r":
            i += 1
        assert text[i] == ":"
        i += 1
        while text[i] in " \n\t\r":
            i += 1
        _, j = dec.raw_decode(text, i)
        if k == key:
            last = text[i:j]
        i = j
