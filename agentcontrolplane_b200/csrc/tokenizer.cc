// tokenizer.cc — see tokenizer.h.
#include "tokenizer.h"

#include <stdint.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <queue>

#include "json.h"
#include "unicode_tables.h"

namespace acp {

// ---------------------------------------------------------------------------------
// synthetic vocabulary
// ---------------------------------------------------------------------------------
namespace {

class SyntheticTokenizer : public Tokenizer {
 public:
  void encode(const std::string& text, std::vector<int>* ids) const override {
    for (unsigned char c : text) ids->push_back((int)c);
  }
  std::string decode(const std::vector<int>& ids) const override {
    std::string out;
    for (int id : ids) {
      if (id < 0) continue;
      if (id < 256) out.push_back((char)(unsigned char)id);
      else if (id < sp_.begin_of_text) {
        // synthetic "word" tokens: a space followed by the id in base-26 letters (LSB first)
        out.push_back(' ');
        unsigned n = (unsigned)(id - 256);
        do { out.push_back((char)('a' + n % 26)); n /= 26; } while (n > 0);
      }
      // special tokens (>= 128000) decode to nothing
    }
    return out;
  }
  int vocab_size() const override { return 128256; }
  const char* kind() const override { return "synthetic-bytes"; }
};

// ---------------------------------------------------------------------------------
// code points
// ---------------------------------------------------------------------------------
struct Cp { uint32_t cp; uint32_t off; };  // code point and its byte offset

// Lenient UTF-8 decoding: an invalid byte becomes one "symbol" code point of its own.
void decode_utf8(const std::string& s, std::vector<Cp>* out) {
  const unsigned char* p = (const unsigned char*)s.data();
  const size_t n = s.size();
  size_t i = 0;
  while (i < n) {
    const unsigned char c = p[i];
    uint32_t cp = 0xFFFD;
    size_t len = 1;
    if (c < 0x80) { cp = c; }
    else if (c >= 0xC2 && c <= 0xDF && i + 1 < n && (p[i + 1] & 0xC0) == 0x80) {
      cp = ((uint32_t)(c & 0x1F) << 6) | (p[i + 1] & 0x3F); len = 2;
    } else if (c >= 0xE0 && c <= 0xEF && i + 2 < n && (p[i + 1] & 0xC0) == 0x80 && (p[i + 2] & 0xC0) == 0x80) {
      const uint32_t v = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(p[i + 1] & 0x3F) << 6) | (p[i + 2] & 0x3F);
      if (v >= 0x800 && !(v >= 0xD800 && v <= 0xDFFF)) { cp = v; len = 3; }
    } else if (c >= 0xF0 && c <= 0xF4 && i + 3 < n && (p[i + 1] & 0xC0) == 0x80 && (p[i + 2] & 0xC0) == 0x80 &&
               (p[i + 3] & 0xC0) == 0x80) {
      const uint32_t v = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(p[i + 1] & 0x3F) << 12) |
                         ((uint32_t)(p[i + 2] & 0x3F) << 6) | (p[i + 3] & 0x3F);
      if (v >= 0x10000 && v <= 0x10FFFF) { cp = v; len = 4; }
    }
    out->push_back({cp, (uint32_t)i});
    i += len;
  }
}

bool in_ranges(uint32_t cp, const UnicodeRange* r, int n) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < r[mid].lo) hi = mid - 1;
    else if (cp > r[mid].hi) lo = mid + 1;
    else return true;
  }
  return false;
}
inline bool is_letter(uint32_t c) {
  if (c < 0x80) return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
  return in_ranges(c, kUnicodeLetter, kUnicodeLetterCount);
}
inline bool is_number(uint32_t c) {
  if (c < 0x80) return c >= '0' && c <= '9';
  return in_ranges(c, kUnicodeNumber, kUnicodeNumberCount);
}
inline bool is_space(uint32_t c) {
  if (c < 0x80) return c == ' ' || (c >= 0x09 && c <= 0x0D);
  return in_ranges(c, kUnicodeSpace, kUnicodeSpaceCount);
}
inline bool is_newline(uint32_t c) { return c == '\r' || c == '\n'; }
inline bool is_symbol(uint32_t c) { return !is_space(c) && !is_letter(c) && !is_number(c); }
inline uint32_t ascii_lower(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

// length (in code points) of the regex match that starts at i; always >= 1
size_t llama3_match(const std::vector<Cp>& s, size_t i) {
  const size_t n = s.size();
  const uint32_t c = s[i].cp;
  // (?i:'s|'t|'re|'ve|'m|'ll|'d)
  if (c == '\'' && i + 1 < n) {
    const uint32_t a = ascii_lower(s[i + 1].cp);
    if (a == 's' || a == 't' || a == 'm' || a == 'd') return 2;
    if (i + 2 < n) {
      const uint32_t b = ascii_lower(s[i + 2].cp);
      if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) return 3;
    }
  }
  // [^\r\n\p{L}\p{N}]?\p{L}+
  {
    size_t j = i;
    if (!is_newline(c) && !is_letter(c) && !is_number(c) && i + 1 < n && is_letter(s[i + 1].cp)) j = i + 1;
    if (is_letter(s[j].cp)) {
      while (j < n && is_letter(s[j].cp)) ++j;
      return j - i;
    }
  }
  // \p{N}{1,3}
  if (is_number(c)) {
    size_t j = i + 1;
    while (j < n && j < i + 3 && is_number(s[j].cp)) ++j;
    return j - i;
  }
  // ` ?[^\s\p{L}\p{N}]+[\r\n]*`
  {
    size_t j = i;
    if (c == ' ' && i + 1 < n && is_symbol(s[i + 1].cp)) j = i + 1;
    if (is_symbol(s[j].cp)) {
      while (j < n && is_symbol(s[j].cp)) ++j;
      while (j < n && is_newline(s[j].cp)) ++j;
      return j - i;
    }
  }
  // whitespace: \s*[\r\n]+ | \s+(?!\S) | \s+
  size_t e = i;
  while (e < n && is_space(s[e].cp)) ++e;
  if (e == i) return 1;  // unreachable: every code point is a letter, number, space or symbol
  for (size_t k = e; k > i; --k)
    if (is_newline(s[k - 1].cp)) return k - i;           // up to and including the run's last newline
  if (e == n) return e - i;                              // run reaches the end of the text
  if (e - i >= 2) return e - i - 1;                      // leave one space to lead the next piece
  return e - i;
}

// ---------------------------------------------------------------------------------
// byte-level BPE
// ---------------------------------------------------------------------------------
class BpeTokenizer : public Tokenizer {
 public:
  bool load(const Json& root, std::string* err);
  void encode(const std::string& text, std::vector<int>* ids) const override;
  std::string decode(const std::vector<int>& ids) const override;
  int vocab_size() const override { return vocab_size_; }
  const char* kind() const override { return "byte-level-bpe"; }

 private:
  void encode_piece(const std::string& piece, std::vector<int>* ids) const;
  std::unordered_map<std::string, int> bytes2id_;      // raw bytes of a token -> id
  std::vector<std::string> id2bytes_;                  // id -> raw bytes ("" for specials / holes)
  std::unordered_map<uint64_t, std::pair<int, int>> merges_;  // (left id, right id) -> (rank, merged id)
  int byte_id_[256];
  bool ignore_merges_ = false;
  int vocab_size_ = 0;
};

// GPT-2's bytes_to_unicode(): printable bytes map to themselves, the rest to U+0100...
void build_byte_unicode(uint32_t byte2cp[256]) {
  bool direct[256] = {false};
  for (int b = 33; b <= 126; ++b) direct[b] = true;
  for (int b = 161; b <= 172; ++b) direct[b] = true;
  for (int b = 174; b <= 255; ++b) direct[b] = true;
  uint32_t next = 256;
  for (int b = 0; b < 256; ++b) byte2cp[b] = direct[b] ? (uint32_t)b : next++;
}

// token spelled in the byte-level alphabet -> raw bytes; false when a code point is outside it
bool unicode_to_bytes(const std::string& tok, const std::unordered_map<uint32_t, int>& cp2byte, std::string* out) {
  std::vector<Cp> cps;
  decode_utf8(tok, &cps);
  out->clear();
  for (const Cp& c : cps) {
    auto it = cp2byte.find(c.cp);
    if (it == cp2byte.end()) return false;
    out->push_back((char)(unsigned char)it->second);
  }
  return true;
}

bool pretokenizer_is_llama3(const Json& pt) {
  // accepted: Sequence[ Split(Regex llama3, isolated), ByteLevel(use_regex=false) ], or the same Split alone
  // followed by ByteLevel; anything else is refused rather than silently mis-tokenised.
  if (!pt.is_object()) return false;
  std::vector<const Json*> stages;
  if (pt.get("type").as_string() == "Sequence") for (const Json& j : pt.get("pretokenizers").items()) stages.push_back(&j);
  else stages.push_back(&pt);
  bool split_ok = false, bytelevel_ok = false;
  for (const Json* st : stages) {
    const std::string type = st->get("type").as_string();
    if (type == "Split") {
      const std::string rx = st->get("pattern").get("Regex").as_string();
      const std::string beh = st->get("behavior").as_string();
      if (rx.find("(?i:'s|'t|'re|'ve|'m|'ll|'d)") != 0 || rx.find("\\p{N}{1,3}") == std::string::npos ||
          rx.find("\\s+(?!\\S)|\\s+") == std::string::npos || st->get("invert").as_bool(false) || beh != "Isolated")
        return false;
      split_ok = true;
    } else if (type == "ByteLevel") {
      if (st->get("use_regex").as_bool(true) || st->get("add_prefix_space").as_bool(false)) return false;
      bytelevel_ok = true;
    } else {
      return false;
    }
  }
  return split_ok && bytelevel_ok;
}

bool BpeTokenizer::load(const Json& root, std::string* err) {
  const Json& model = root.get("model");
  const std::string mtype = model.get("type").as_string();
  if (!mtype.empty() && mtype != "BPE") { *err = "tokenizer.json: model type " + mtype + " is not supported (byte-level BPE only)"; return false; }
  if (!model.get("vocab").is_object() || !model.get("merges").is_array()) { *err = "tokenizer.json: model.vocab / model.merges missing"; return false; }
  if (!root.get("normalizer").is_null()) { *err = "tokenizer.json: a normalizer is configured (not supported)"; return false; }
  if (!pretokenizer_is_llama3(root.get("pre_tokenizer"))) { *err = "tokenizer.json: pre_tokenizer is not the Llama-3 Split + ByteLevel pipeline"; return false; }
  if (model.get("byte_fallback").as_bool(false)) { *err = "tokenizer.json: byte_fallback is not supported"; return false; }
  ignore_merges_ = model.get("ignore_merges").as_bool(false);

  uint32_t byte2cp[256];
  build_byte_unicode(byte2cp);
  std::unordered_map<uint32_t, int> cp2byte;
  for (int b = 0; b < 256; ++b) cp2byte[byte2cp[b]] = b;

  std::unordered_map<std::string, int> str2id;  // vocabulary spelling -> id (merges refer to spellings)
  int max_id = -1;
  for (const auto& kv : model.get("vocab").members()) {
    const int id = (int)kv.second.as_int(-1);
    if (id < 0) { *err = "tokenizer.json: negative token id"; return false; }
    str2id[kv.first] = id;
    if (id > max_id) max_id = id;
  }
  for (const Json& at : root.get("added_tokens").items()) {
    const int id = (int)at.get("id").as_int(-1);
    if (id > max_id) max_id = id;
  }
  vocab_size_ = max_id + 1;
  id2bytes_.assign((size_t)vocab_size_, std::string());
  std::string raw;
  for (const auto& kv : str2id) {
    if (!unicode_to_bytes(kv.first, cp2byte, &raw)) continue;  // not a byte-level token (e.g. a special in vocab)
    id2bytes_[(size_t)kv.second] = raw;
    bytes2id_[raw] = kv.second;
  }
  for (int b = 0; b < 256; ++b) {
    auto it = bytes2id_.find(std::string(1, (char)(unsigned char)b));
    if (it == bytes2id_.end()) { *err = "tokenizer.json: vocabulary lacks a single-byte token (not byte-level BPE)"; return false; }
    byte_id_[b] = it->second;
  }
  int rank = 0;
  for (const Json& m : model.get("merges").items()) {
    std::string a, b;
    if (m.is_string()) {
      const std::string& s = m.as_string();
      const size_t sp = s.find(' ');
      if (sp == std::string::npos) { *err = "tokenizer.json: malformed merge entry"; return false; }
      a = s.substr(0, sp); b = s.substr(sp + 1);
    } else if (m.is_array() && m.size() == 2) {
      a = m.items()[0].as_string(); b = m.items()[1].as_string();
    } else { *err = "tokenizer.json: malformed merge entry"; return false; }
    auto ia = str2id.find(a), ib = str2id.find(b), im = str2id.find(a + b);
    if (ia != str2id.end() && ib != str2id.end() && im != str2id.end())
      merges_.emplace(((uint64_t)(uint32_t)ia->second << 32) | (uint32_t)ib->second, std::make_pair(rank, im->second));
    ++rank;
  }
  // chat-control tokens by spelling
  std::unordered_map<std::string, int> added;
  for (const Json& at : root.get("added_tokens").items()) {
    const int id = (int)at.get("id").as_int(-1);
    added[at.get("content").as_string()] = id;
    if (id >= 0 && id < vocab_size_) {  // special ids never decode to text and never come out of encode()
      auto it = bytes2id_.find(id2bytes_[(size_t)id]);
      if (it != bytes2id_.end() && it->second == id && at.get("special").as_bool(false)) bytes2id_.erase(it);
      if (at.get("special").as_bool(false)) id2bytes_[(size_t)id].clear();
    }
  }
  auto sid = [&](const char* name) { auto it = added.find(name); return it == added.end() ? -1 : it->second; };
  sp_.begin_of_text = sid("<|begin_of_text|>");
  sp_.end_of_text = sid("<|end_of_text|>");
  sp_.start_header = sid("<|start_header_id|>");
  sp_.end_header = sid("<|end_header_id|>");
  sp_.eom = sid("<|eom_id|>");
  sp_.eot = sid("<|eot_id|>");
  sp_.python_tag = sid("<|python_tag|>");
  if (sp_.begin_of_text < 0 || sp_.start_header < 0 || sp_.end_header < 0 || sp_.eot < 0) {
    *err = "tokenizer.json: the Llama-3 chat special tokens (<|begin_of_text|>, <|start_header_id|>, <|end_header_id|>, <|eot_id|>) are missing";
    return false;
  }
  return true;
}

void BpeTokenizer::encode_piece(const std::string& piece, std::vector<int>* ids) const {
  if (piece.empty()) return;
  if (ignore_merges_ || piece.size() == 1) {
    auto it = bytes2id_.find(piece);
    if (it != bytes2id_.end()) { ids->push_back(it->second); return; }
  }
  // linked list of symbols + a heap of candidate merges ordered by (rank, position): the
  // lowest-ranked, leftmost pair merges first, exactly one occurrence at a time
  struct Sym { int id, prev, next; };
  const int n = (int)piece.size();
  std::vector<Sym> sym((size_t)n);
  for (int i = 0; i < n; ++i) sym[(size_t)i] = {byte_id_[(unsigned char)piece[(size_t)i]], i - 1, i + 1 < n ? i + 1 : -1};
  struct Cand { int rank, pos, left, right, merged; };
  auto worse = [](const Cand& a, const Cand& b) { return a.rank != b.rank ? a.rank > b.rank : a.pos > b.pos; };
  std::priority_queue<Cand, std::vector<Cand>, decltype(worse)> heap(worse);
  auto push = [&](int pos) {
    if (pos < 0) return;
    const int nx = sym[(size_t)pos].next;
    if (nx < 0) return;
    auto it = merges_.find(((uint64_t)(uint32_t)sym[(size_t)pos].id << 32) | (uint32_t)sym[(size_t)nx].id);
    if (it != merges_.end()) heap.push({it->second.first, pos, sym[(size_t)pos].id, sym[(size_t)nx].id, it->second.second});
  };
  for (int i = 0; i + 1 < n; ++i) push(i);
  while (!heap.empty()) {
    const Cand c = heap.top();
    heap.pop();
    Sym& s = sym[(size_t)c.pos];
    if (s.id != c.left || s.next < 0 || sym[(size_t)s.next].id != c.right) continue;  // stale
    const int dead = s.next;
    s.id = c.merged;
    s.next = sym[(size_t)dead].next;
    sym[(size_t)dead].id = -1;
    if (s.next >= 0) sym[(size_t)s.next].prev = c.pos;
    push(s.prev);
    push(c.pos);
  }
  for (int i = 0; i >= 0; i = sym[(size_t)i].next) ids->push_back(sym[(size_t)i].id);
}

void BpeTokenizer::encode(const std::string& text, std::vector<int>* ids) const {
  std::vector<std::string> pieces;
  llama3_pretokenize(text, &pieces);
  for (const std::string& p : pieces) encode_piece(p, ids);
}

std::string BpeTokenizer::decode(const std::vector<int>& ids) const {
  std::string out;
  for (int id : ids)
    if (id >= 0 && id < vocab_size_) out += id2bytes_[(size_t)id];
  return out;
}

}  // namespace

const Tokenizer& synthetic_tokenizer() {
  static const SyntheticTokenizer t;
  return t;
}

void llama3_pretokenize(const std::string& text, std::vector<std::string>* pieces) {
  std::vector<Cp> cps;
  decode_utf8(text, &cps);
  size_t i = 0;
  while (i < cps.size()) {
    const size_t len = llama3_match(cps, i);
    const size_t b0 = cps[i].off, b1 = i + len < cps.size() ? cps[i + len].off : text.size();
    pieces->push_back(text.substr(b0, b1 - b0));
    i += len;
  }
}

std::unique_ptr<Tokenizer> load_tokenizer_json(const std::string& path, std::string* err) {
  std::string e;
  if (!err) err = &e;
  std::string file = path;
  struct stat st;
  if (stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) file = path + "/tokenizer.json";
  FILE* f = fopen(file.c_str(), "rb");
  if (!f) { *err = "cannot open " + file; return nullptr; }
  std::string text;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  Json root;
  std::string perr;
  if (!Json::parse(text, &root, &perr) || !root.is_object()) { *err = file + ": " + perr; return nullptr; }
  std::unique_ptr<BpeTokenizer> t(new BpeTokenizer());
  if (!t->load(root, err)) return nullptr;
  return std::unique_ptr<Tokenizer>(t.release());
}

}  // namespace acp
