// AddressSanitizer + UBSan smoke-fuzzer for the code that parses UNTRUSTED bytes on the hot path's
// host side: the chat-completions request body (message content comes from Task CRs and tool
// results), the completion text -> tool-call extraction, the tokenizers, and the checkpoint
// reader.  Deterministic mutation of valid seeds; any sanitizer report fails
// tests/test_sanitizers_cpu.py.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "chat.h"
#include "safetensors.h"
#include "tokenizer.h"

using namespace acp;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
  rng_state ^= rng_state << 7;
  rng_state ^= rng_state >> 9;
  return (uint32_t)(rng_state >> 16);
}

static std::string mutate(const std::string& seed) {
  std::string s = seed;
  const int edits = 1 + (int)(rnd() % 4);
  for (int e = 0; e < edits && !s.empty(); ++e) {
    const size_t pos = rnd() % s.size();
    switch (rnd() % 6) {
      case 0: s[pos] = (char)(rnd() & 0xff); break;
      case 1: s.erase(pos, 1 + rnd() % 8); break;
      case 2: s.insert(pos, 1 + rnd() % 4, (char)(rnd() & 0xff)); break;
      case 3: s.insert(pos, s.substr(rnd() % s.size(), rnd() % 24)); break;
      case 4: s.insert(pos, "\\ud83d"); break;   // lone surrogate escapes
      default: s.resize(pos); break;
    }
  }
  return s;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  const char* tok_path = argc > 2 ? argv[2] : nullptr;
  const std::vector<std::string> requests = {
      "{\"model\":\"m\",\"messages\":[{\"role\":\"system\",\"content\":\"You are a helpful assistant.\"},"
      "{\"role\":\"user\",\"content\":\"What's at https://x.y/z?\\n\"}],\"max_tokens\":16,\"temperature\":0}",
      "{\"model\":\"m\",\"tools\":[{\"type\":\"function\",\"function\":{\"name\":\"fetch__fetch\",\"description\":\"Fetch\","
      "\"parameters\":{\"type\":\"object\",\"properties\":{\"url\":{\"type\":\"string\"}}}}}],\"messages\":[{\"role\":\"user\","
      "\"content\":\"go\"},{\"role\":\"assistant\",\"content\":\"\",\"tool_calls\":[{\"id\":\"1\",\"type\":\"function\","
      "\"function\":{\"name\":\"fetch__fetch\",\"arguments\":\"{\\\"url\\\": \\\"u\\\"}\"}}]},{\"role\":\"tool\","
      "\"tool_call_id\":\"1\",\"content\":\"{\\\"data\\\": [1,2,3]}\"}],\"acp\":{\"force_tokens\":[1,2,3],\"return_logits\":1}}",
      "{\"model\":\"m\",\"messages\":[{\"role\":\"user\",\"content\":\"\xc3\xa9\xf0\x9f\x99\x82 caf\xc3\xa9   <|eot_id|> \\u00e9\\ud83d\\ude42\"}],"
      "\"acp\":{\"prompt_token_ids\":[128000,1,2,3]}}"};
  const std::vector<std::string> completions = {
      "{\"name\": \"fetch__fetch\", \"parameters\": {\"url\": \"https://api.example.com/data\"}}",
      "<|python_tag|>{\"name\": \"fetch__fetch\", \"arguments\": {\"a\": {\"b\": [1, {\"c\": \"}\"}]}}}\n"
      "{\"name\": \"fetch__fetch\", \"parameters\": {}}",
      "plain answer with { braces } and \"quotes\" and \xe6\xb1\x89\xe5\xad\x97 123 'll"};
  std::unique_ptr<Tokenizer> bpe;
  if (tok_path && *tok_path) {
    std::string err;
    bpe = load_tokenizer_json(tok_path, &err);
    if (!bpe) { fprintf(stderr, "tokenizer: %s\n", err.c_str()); return 2; }
  }
  const Tokenizer& tok = bpe ? *bpe : synthetic_tokenizer();
  std::vector<ToolDef> tools(1);
  tools[0].type = "function";
  tools[0].name = "fetch__fetch";
  size_t sink = 0;
  for (int i = 0; i < iters; ++i) {
    {
      const std::string body = mutate(requests[(size_t)i % requests.size()]);
      ChatRequest req;
      std::string err;
      if (parse_chat_request(body.data(), body.size(), &req, &err) == 0) {
        std::vector<int> ids;
        render_prompt(req, &ids, tok);
        sink += ids.size() + render_prompt_text(req).size();
      }
    }
    {
      const std::string text = mutate(completions[(size_t)i % completions.size()]);
      ParsedCompletion pc = parse_completion(text, tools, "call_");
      sink += pc.content.size() + pc.tool_calls.size();
      std::vector<int> ids;
      tok.encode(text, &ids);
      std::vector<int> noisy = ids;
      noisy.push_back((int)rnd());
      noisy.push_back(-(int)(rnd() % 1000));
      sink += tok.decode(noisy).size();
      std::vector<std::string> pieces;
      llama3_pretokenize(text, &pieces);
      sink += pieces.size();
    }
  }
  // file reader: truncated / corrupted copies of a valid file must be refused, never crash
  if (argc > 3) {
    FILE* f = fopen(argv[3], "rb");
    std::string data;
    char buf[65536];
    size_t n;
    while (f && (n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
    if (f) fclose(f);
    const std::string tmp = std::string(argv[3]) + ".fuzz";
    for (int i = 0; i < 300 && !data.empty(); ++i) {
      std::string m = data;
      const size_t lim = m.size() < 4096 ? m.size() : 4096;     // headers live at the front
      for (int e = 0; e < 3; ++e) m[rnd() % lim] = (char)(rnd() & 0xff);
      if (i % 3 == 0) m.resize(rnd() % m.size());
      FILE* o = fopen(tmp.c_str(), "wb");
      if (!o) break;
      fwrite(m.data(), 1, m.size(), o);
      fclose(o);
      Checkpoint ck;
      std::string err;
      if (ck.open(tmp, &err)) sink += ck.tensor_count() + ck.index_json().size();
    }
    remove(tmp.c_str());
  }
  // tokenizer.json reader: corrupted vocabularies / merges must be refused or load consistently
  if (tok_path && *tok_path) {
    FILE* f = fopen(tok_path, "rb");
    std::string data;
    char buf[65536];
    size_t n;
    while (f && (n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
    if (f) fclose(f);
    const std::string tmp = std::string(argc > 3 ? argv[3] : "/tmp/acp_fuzz") + ".tokenizer.json";
    for (int i = 0; i < 200 && !data.empty(); ++i) {
      std::string m = mutate(data);
      FILE* o = fopen(tmp.c_str(), "wb");
      if (!o) break;
      fwrite(m.data(), 1, m.size(), o);
      fclose(o);
      std::string err;
      std::unique_ptr<Tokenizer> t = load_tokenizer_json(tmp, &err);
      if (t) {
        std::vector<int> ids;
        t->encode("Hello, world! 123 it's \n\n  x", &ids);
        sink += t->decode(ids).size();
      }
    }
    remove(tmp.c_str());
  }
  printf("fuzz ok: %d iterations, sink %zu\n", iters, sink);
  return 0;
}
