// json.h — small self-contained JSON value / parser / serialiser for the C-ABI boundary.
// Objects keep insertion order (the OpenAI wire bodies are order-sensitive for byte-exact tests);
// strings are UTF-8, \uXXXX escapes (incl. surrogate pairs) are decoded on parse; invalid UTF-8
// is replaced by U+FFFD on serialisation so that every response is valid JSON text.
#pragma once
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace acp {

class Json {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Json() : type_(Null) {}
  Json(bool b) : type_(Bool), b_(b) {}
  Json(double d) : type_(Number), d_(d), is_int_(false) {}
  Json(int i) : type_(Number), d_(i), i_(i), is_int_(true) {}
  Json(long long i) : type_(Number), d_((double)i), i_(i), is_int_(true) {}
  Json(const char* s) : type_(String), s_(s) {}
  Json(const std::string& s) : type_(String), s_(s) {}
  static Json array() { Json j; j.type_ = Array; return j; }
  static Json object() { Json j; j.type_ = Object; return j; }

  Type type() const { return type_; }
  bool is_null() const { return type_ == Null; }
  bool is_bool() const { return type_ == Bool; }
  bool is_number() const { return type_ == Number; }
  bool is_string() const { return type_ == String; }
  bool is_array() const { return type_ == Array; }
  bool is_object() const { return type_ == Object; }
  bool as_bool(bool def = false) const { return type_ == Bool ? b_ : def; }
  double as_double(double def = 0) const { return type_ == Number ? d_ : def; }
  long long as_int(long long def = 0) const {
    if (type_ != Number) return def;
    if (is_int_) return i_;
    // out-of-range / NaN doubles (untrusted request bodies): the cast would be undefined behaviour
    if (!(d_ == d_)) return def;
    if (d_ >= 9.2e18) return 9200000000000000000LL;
    if (d_ <= -9.2e18) return -9200000000000000000LL;
    return (long long)d_;
  }
  const std::string& as_string() const { static const std::string e; return type_ == String ? s_ : e; }
  const std::vector<Json>& items() const { return a_; }
  std::vector<Json>& items() { return a_; }
  const std::vector<std::pair<std::string, Json>>& members() const { return o_; }
  size_t size() const { return type_ == Array ? a_.size() : type_ == Object ? o_.size() : 0; }

  const Json* find(const std::string& key) const {
    if (type_ != Object) return nullptr;
    for (auto& kv : o_) if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const Json& get(const std::string& key) const {
    static const Json null_json;
    const Json* p = find(key);
    return p ? *p : null_json;
  }
  Json& set(const std::string& key, Json v) {
    for (auto& kv : o_) if (kv.first == key) { kv.second = std::move(v); return kv.second; }
    o_.emplace_back(key, std::move(v));
    return o_.back().second;
  }
  void push(Json v) { a_.push_back(std::move(v)); }

  // ---- serialisation ----
  static void escape_to(const std::string& s, std::string& out) {
    out.push_back('"');
    const unsigned char* p = (const unsigned char*)s.data();
    const size_t n = s.size();
    size_t i = 0;
    char buf[8];
    while (i < n) {
      unsigned char c = p[i];
      if (c == '"') { out += "\\\""; ++i; }
      else if (c == '\\') { out += "\\\\"; ++i; }
      else if (c == '\n') { out += "\\n"; ++i; }
      else if (c == '\r') { out += "\\r"; ++i; }
      else if (c == '\t') { out += "\\t"; ++i; }
      else if (c < 0x20) { snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; ++i; }
      else if (c < 0x80) { out.push_back((char)c); ++i; }
      else {
        // validate one UTF-8 sequence; replace with U+FFFD if malformed
        int len = (c >= 0xF0 && c <= 0xF4) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2 && c < 0xE0) ? 2 : 0;
        bool ok = len != 0 && i + len <= n;
        if (ok) {
          for (int k = 1; k < len; ++k) if ((p[i + k] & 0xC0) != 0x80) ok = false;
          if (ok && len == 3) {
            unsigned cp = ((c & 0x0F) << 12) | ((p[i + 1] & 0x3F) << 6) | (p[i + 2] & 0x3F);
            if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) ok = false;
          }
          if (ok && len == 4) {
            unsigned cp = ((c & 0x07) << 18) | ((p[i + 1] & 0x3F) << 12) | ((p[i + 2] & 0x3F) << 6) | (p[i + 3] & 0x3F);
            if (cp < 0x10000 || cp > 0x10FFFF) ok = false;
          }
        }
        if (ok) { out.append((const char*)p + i, len); i += len; }
        else { out += "\xEF\xBF\xBD"; ++i; }
      }
    }
    out.push_back('"');
  }
  void dump_to(std::string& out) const {
    switch (type_) {
      case Null: out += "null"; break;
      case Bool: out += b_ ? "true" : "false"; break;
      case Number: {
        char buf[40];
        if (is_int_) snprintf(buf, sizeof buf, "%lld", i_);
        else if (std::isfinite(d_)) snprintf(buf, sizeof buf, "%.17g", d_);
        else snprintf(buf, sizeof buf, "null");
        out += buf;
        break;
      }
      case String: escape_to(s_, out); break;
      case Array:
        out.push_back('[');
        for (size_t i = 0; i < a_.size(); ++i) { if (i) out.push_back(','); a_[i].dump_to(out); }
        out.push_back(']');
        break;
      case Object:
        out.push_back('{');
        for (size_t i = 0; i < o_.size(); ++i) {
          if (i) out.push_back(',');
          escape_to(o_[i].first, out);
          out.push_back(':');
          o_[i].second.dump_to(out);
        }
        out.push_back('}');
        break;
    }
  }
  std::string dump() const { std::string s; dump_to(s); return s; }

  // ---- parsing ----
  // Returns true on success.  *end (optional) receives the offset just past the parsed value, so a
  // caller can parse a JSON value embedded in longer text (tool-call extraction).
  struct MemberSpan { std::string key; size_t begin, end; };  // value bytes [begin, end)
  static bool parse(const char* text, size_t len, Json* out, std::string* err, size_t* end = nullptr,
                    bool allow_trailing = false, std::vector<MemberSpan>* top_spans = nullptr) {
    Parser p{text, len, 0, err, top_spans};
    p.skip_ws();
    if (!p.value(out, 0)) return false;
    if (end) *end = p.pos;
    if (!allow_trailing) {
      p.skip_ws();
      if (p.pos != len) { p.fail("trailing characters after JSON value"); return false; }
    }
    return true;
  }
  static bool parse(const std::string& s, Json* out, std::string* err) {
    return parse(s.data(), s.size(), out, err);
  }

 private:
  struct Parser {
    const char* t;
    size_t n, pos;
    std::string* err;
    std::vector<MemberSpan>* spans;  // members of the outermost object, if requested
    bool fail(const char* m) {
      if (err && err->empty()) {
        char buf[96];
        snprintf(buf, sizeof buf, "%s at offset %zu", m, pos);
        *err = buf;
      }
      return false;
    }
    void skip_ws() { while (pos < n && (t[pos] == ' ' || t[pos] == '\n' || t[pos] == '\t' || t[pos] == '\r')) ++pos; }
    bool lit(const char* s) {
      size_t l = strlen(s);
      if (pos + l <= n && memcmp(t + pos, s, l) == 0) { pos += l; return true; }
      return false;
    }
    static void put_utf8(unsigned cp, std::string& o) {
      if (cp < 0x80) o.push_back((char)cp);
      else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
      else if (cp < 0x10000) {
        o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
        o.push_back((char)(0x80 | (cp & 0x3F)));
      } else {
        o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
        o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F)));
      }
    }
    bool hex4(unsigned* v) {
      if (pos + 4 > n) return false;
      unsigned r = 0;
      for (int i = 0; i < 4; ++i) {
        char c = t[pos + i];
        r <<= 4;
        if (c >= '0' && c <= '9') r |= c - '0';
        else if (c >= 'a' && c <= 'f') r |= c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') r |= c - 'A' + 10;
        else return false;
      }
      pos += 4;
      *v = r;
      return true;
    }
    bool string(std::string* out) {
      if (pos >= n || t[pos] != '"') return fail("expected string");
      ++pos;
      out->clear();
      while (pos < n) {
        unsigned char c = (unsigned char)t[pos];
        if (c == '"') { ++pos; return true; }
        if (c == '\\') {
          if (++pos >= n) break;
          char e = t[pos++];
          switch (e) {
            case '"': out->push_back('"'); break;
            case '\\': out->push_back('\\'); break;
            case '/': out->push_back('/'); break;
            case 'b': out->push_back('\b'); break;
            case 'f': out->push_back('\f'); break;
            case 'n': out->push_back('\n'); break;
            case 'r': out->push_back('\r'); break;
            case 't': out->push_back('\t'); break;
            case 'u': {
              unsigned cp;
              if (!hex4(&cp)) return fail("bad \\u escape");
              if (cp >= 0xD800 && cp <= 0xDBFF && pos + 6 <= n && t[pos] == '\\' && t[pos + 1] == 'u') {
                size_t save = pos;
                pos += 2;
                unsigned lo;
                if (hex4(&lo) && lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                else { pos = save; cp = 0xFFFD; }
              } else if (cp >= 0xD800 && cp <= 0xDFFF) cp = 0xFFFD;
              put_utf8(cp, *out);
              break;
            }
            default: return fail("bad escape");
          }
        } else if (c < 0x20) {
          return fail("control character in string");
        } else { out->push_back((char)c); ++pos; }
      }
      return fail("unterminated string");
    }
    bool number(Json* out) {
      size_t s = pos;
      bool is_int = true;
      if (pos < n && t[pos] == '-') ++pos;
      if (pos >= n || !(t[pos] >= '0' && t[pos] <= '9')) return fail("bad number");
      while (pos < n && t[pos] >= '0' && t[pos] <= '9') ++pos;
      if (pos < n && t[pos] == '.') { is_int = false; ++pos; while (pos < n && t[pos] >= '0' && t[pos] <= '9') ++pos; }
      if (pos < n && (t[pos] == 'e' || t[pos] == 'E')) {
        is_int = false; ++pos;
        if (pos < n && (t[pos] == '+' || t[pos] == '-')) ++pos;
        while (pos < n && t[pos] >= '0' && t[pos] <= '9') ++pos;
      }
      std::string tok(t + s, pos - s);
      // every integer that fits int64 stays exact (request `seed`s are 63-bit); larger ones become doubles
      bool exact = false;
      if (is_int && tok.size() <= 20) {
        errno = 0;
        const long long v = strtoll(tok.c_str(), nullptr, 10);
        if (errno == 0) { *out = Json(v); exact = true; }
      }
      if (!exact) *out = Json(strtod(tok.c_str(), nullptr));
      return true;
    }
    bool value(Json* out, int depth) {
      if (depth > 200) return fail("nesting too deep");
      skip_ws();
      if (pos >= n) return fail("unexpected end of input");
      char c = t[pos];
      if (c == '{') {
        ++pos;
        *out = Json::object();
        skip_ws();
        if (pos < n && t[pos] == '}') { ++pos; return true; }
        while (true) {
          skip_ws();
          std::string k;
          if (!string(&k)) return false;
          skip_ws();
          if (pos >= n || t[pos] != ':') return fail("expected ':'");
          ++pos;
          Json v;
          skip_ws();
          const size_t vbegin = pos;
          if (!value(&v, depth + 1)) return false;
          if (depth == 0 && spans) spans->push_back(MemberSpan{k, vbegin, pos});
          out->o_.emplace_back(std::move(k), std::move(v));
          skip_ws();
          if (pos < n && t[pos] == ',') { ++pos; continue; }
          if (pos < n && t[pos] == '}') { ++pos; return true; }
          return fail("expected ',' or '}'");
        }
      }
      if (c == '[') {
        ++pos;
        *out = Json::array();
        skip_ws();
        if (pos < n && t[pos] == ']') { ++pos; return true; }
        while (true) {
          Json v;
          if (!value(&v, depth + 1)) return false;
          out->a_.push_back(std::move(v));
          skip_ws();
          if (pos < n && t[pos] == ',') { ++pos; continue; }
          if (pos < n && t[pos] == ']') { ++pos; return true; }
          return fail("expected ',' or ']'");
        }
      }
      if (c == '"') { std::string s; if (!string(&s)) return false; *out = Json(s); return true; }
      if (lit("true")) { *out = Json(true); return true; }
      if (lit("false")) { *out = Json(false); return true; }
      if (lit("null")) { *out = Json(); return true; }
      if (c == '-' || (c >= '0' && c <= '9')) return number(out);
      return fail("unexpected character");
    }
  };

  Type type_;
  bool b_ = false;
  double d_ = 0;
  long long i_ = 0;
  bool is_int_ = false;
  std::string s_;
  std::vector<Json> a_;
  std::vector<std::pair<std::string, Json>> o_;
};

}  // namespace acp
