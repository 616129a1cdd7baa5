// tokenizer.h — the vocabulary side of the local provider.  The reference never tokenises (the
// hosted provider behind acp/internal/llmclient/langchaingo_client.go:102 does); a local engine
// must.  Two implementations behind one interface:
//   * the synthetic byte-level vocabulary used with the seeded weights, and
//   * a HuggingFace `tokenizer.json` byte-level BPE (Llama-3 family): GPT-2 byte<->unicode table,
//     rank-ordered merges, `ignore_merges`, and the Llama-3 pre-tokenizer regex implemented as a
//     hand-written matcher over code points (no regex engine in this image's C++ toolchain).
#pragma once
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace acp {

struct SpecialIds {
  int begin_of_text = 128000, end_of_text = 128001, start_header = 128006, end_header = 128007,
      eom = 128008, eot = 128009, python_tag = 128010;
};

class Tokenizer {
 public:
  virtual ~Tokenizer() {}
  // ordinary text -> ids (special-token spellings inside `text` are NOT recognised: user content
  // can never inject control tokens)
  virtual void encode(const std::string& text, std::vector<int>* ids) const = 0;
  // ids -> bytes; special / out-of-range ids decode to nothing
  virtual std::string decode(const std::vector<int>& ids) const = 0;
  virtual int vocab_size() const = 0;   // ids are in [0, vocab_size)
  virtual const char* kind() const = 0;
  const SpecialIds& special() const { return sp_; }
  bool is_stop(int id) const { return id == sp_.eot || id == sp_.eom || id == sp_.end_of_text; }

 protected:
  SpecialIds sp_;
};

// ids 0..255 = bytes, 256..127999 = " " + base-26 letters, Llama-3 specials at their real ids
const Tokenizer& synthetic_tokenizer();

// HuggingFace tokenizer.json (`path` = the file, or a directory containing it).  nullptr + *err on
// anything this implementation does not cover (non-BPE model, unknown pre-tokenizer, missing
// Llama-3 chat special tokens).
std::unique_ptr<Tokenizer> load_tokenizer_json(const std::string& path, std::string* err);

// Splits text the way the Llama-3 pre-tokenizer regex does (exposed for tests):
// (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
void llama3_pretokenize(const std::string& text, std::vector<std::string>* pieces);

}  // namespace acp
