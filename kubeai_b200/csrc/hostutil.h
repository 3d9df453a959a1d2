// Host-side helpers shared by the serving shell (server.cc) and the load generator (harness.cc):
// a small JSON DOM (parse + serialise), UTF-8 rune handling (Go's []rune semantics for
// api/openai/v1/utils.go:5-8 firstNChars), and the synthetic tokenizer / chat template
// (no Llama-3 tokenizer offline).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <time.h>

#include <string>
#include <utility>
#include <vector>

namespace b200 {

// ------------------------------------------------------------------ JSON
struct JVal {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;

  const JVal* get(const char* key) const {
    if (type != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool is_null() const { return type == Null; }
};

class JParser {
 public:
  JParser(const char* s, size_t n) : p_(s), end_(s + n) {}
  bool parse(JVal* out, std::string* err) {
    ws();
    if (!value(out, 0)) {
      if (err) *err = err_.empty() ? "invalid JSON" : err_;
      return false;
    }
    ws();
    if (p_ != end_) {
      if (err) *err = "invalid character after top-level value";
      return false;
    }
    return true;
  }

 private:
  const char* p_;
  const char* end_;
  std::string err_;
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
  }
  bool fail(const char* m) {
    if (err_.empty()) err_ = m;
    return false;
  }
  static void put_utf8(std::string& s, uint32_t c) {
    if (c < 0x80) s += static_cast<char>(c);
    else if (c < 0x800) { s += static_cast<char>(0xC0 | (c >> 6)); s += static_cast<char>(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { s += static_cast<char>(0xE0 | (c >> 12)); s += static_cast<char>(0x80 | ((c >> 6) & 0x3F)); s += static_cast<char>(0x80 | (c & 0x3F)); }
    else { s += static_cast<char>(0xF0 | (c >> 18)); s += static_cast<char>(0x80 | ((c >> 12) & 0x3F)); s += static_cast<char>(0x80 | ((c >> 6) & 0x3F)); s += static_cast<char>(0x80 | (c & 0x3F)); }
  }
  bool hex4(uint32_t* v) {
    if (end_ - p_ < 4) return false;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
      char c = p_[i];
      r <<= 4;
      if (c >= '0' && c <= '9') r |= c - '0';
      else if (c >= 'a' && c <= 'f') r |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') r |= c - 'A' + 10;
      else return false;
    }
    p_ += 4;
    *v = r;
    return true;
  }
  bool string(std::string* out) {
    if (p_ >= end_ || *p_ != '"') return fail("expected string");
    ++p_;
    out->clear();
    while (p_ < end_) {
      unsigned char c = static_cast<unsigned char>(*p_++);
      if (c == '"') return true;
      if (c < 0x20) return fail("invalid control character in string");
      if (c != '\\') { *out += static_cast<char>(c); continue; }
      if (p_ >= end_) break;
      char e = *p_++;
      switch (e) {
        case '"': *out += '"'; break;
        case '\\': *out += '\\'; break;
        case '/': *out += '/'; break;
        case 'b': *out += '\b'; break;
        case 'f': *out += '\f'; break;
        case 'n': *out += '\n'; break;
        case 'r': *out += '\r'; break;
        case 't': *out += '\t'; break;
        case 'u': {
          uint32_t u;
          if (!hex4(&u)) return fail("invalid \\u escape");
          if (u >= 0xD800 && u < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
            const char* save = p_;
            p_ += 2;
            uint32_t lo;
            if (hex4(&lo) && lo >= 0xDC00 && lo < 0xE000) u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00);
            else { p_ = save; u = 0xFFFD; }
          } else if (u >= 0xD800 && u < 0xE000) {
            u = 0xFFFD;
          }
          put_utf8(*out, u);
          break;
        }
        default: return fail("invalid escape sequence");
      }
    }
    return fail("unexpected end of JSON input");
  }
  bool value(JVal* v, int depth) {
    if (depth > 200) return fail("JSON nested too deeply");
    if (p_ >= end_) return fail("unexpected end of JSON input");
    char c = *p_;
    if (c == '{') {
      ++p_;
      v->type = JVal::Obj;
      ws();
      if (p_ < end_ && *p_ == '}') { ++p_; return true; }
      for (;;) {
        ws();
        std::string k;
        if (!string(&k)) return false;
        ws();
        if (p_ >= end_ || *p_ != ':') return fail("expected ':' after object key");
        ++p_;
        ws();
        JVal child;
        if (!value(&child, depth + 1)) return false;
        v->obj.emplace_back(std::move(k), std::move(child));
        ws();
        if (p_ < end_ && *p_ == ',') { ++p_; continue; }
        if (p_ < end_ && *p_ == '}') { ++p_; return true; }
        return fail("expected ',' or '}' in object");
      }
    }
    if (c == '[') {
      ++p_;
      v->type = JVal::Arr;
      ws();
      if (p_ < end_ && *p_ == ']') { ++p_; return true; }
      for (;;) {
        ws();
        JVal child;
        if (!value(&child, depth + 1)) return false;
        v->arr.push_back(std::move(child));
        ws();
        if (p_ < end_ && *p_ == ',') { ++p_; continue; }
        if (p_ < end_ && *p_ == ']') { ++p_; return true; }
        return fail("expected ',' or ']' in array");
      }
    }
    if (c == '"') {
      v->type = JVal::Str;
      return string(&v->str);
    }
    if (c == 't' && end_ - p_ >= 4 && !memcmp(p_, "true", 4)) { p_ += 4; v->type = JVal::Bool; v->b = true; return true; }
    if (c == 'f' && end_ - p_ >= 5 && !memcmp(p_, "false", 5)) { p_ += 5; v->type = JVal::Bool; v->b = false; return true; }
    if (c == 'n' && end_ - p_ >= 4 && !memcmp(p_, "null", 4)) { p_ += 4; v->type = JVal::Null; return true; }
    if (c == '-' || (c >= '0' && c <= '9')) {
      const char* s = p_;
      bool isint = true;
      if (*p_ == '-') ++p_;
      if (p_ >= end_ || *p_ < '0' || *p_ > '9') return fail("invalid number");
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
      if (p_ < end_ && *p_ == '.') { isint = false; ++p_; while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
      if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
        isint = false; ++p_;
        if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
        while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
      }
      v->type = JVal::Num;
      v->is_int = isint;
      v->num = strtod(std::string(s, p_).c_str(), nullptr);
      return true;
    }
    return fail("invalid character looking for beginning of value");
  }
};

inline void json_escape(std::string& out, const std::string& s) {
  out += '"';
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
        else out += static_cast<char>(c);
    }
  }
  out += '"';
}
inline std::string json_str(const std::string& s) {
  std::string o;
  json_escape(o, s);
  return o;
}

// ------------------------------------------------------------------ UTF-8 runes
// First n runes of s (Go: string([]rune(s)[:min(n, len)])); invalid bytes count as one rune each.
inline std::string first_n_runes(const std::string& s, int n) {
  size_t i = 0;
  int runes = 0;
  while (i < s.size() && runes < n) {
    unsigned char c = static_cast<unsigned char>(s[i]);
    size_t len = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
    if (i + len > s.size()) len = 1;
    for (size_t k = 1; k < len; ++k)
      if ((static_cast<unsigned char>(s[i + k]) & 0xC0) != 0x80) { len = 1; break; }
    i += len;
    ++runes;
  }
  return s.substr(0, i);
}

// ------------------------------------------------------------------ synthetic tokenizer
// ids 0..255 are raw bytes; every id t also has the 5-byte spelling " wxyz" (space + 4 lowercase
// letters, base 26) which tokenises back to exactly t — so generated text echoed into the next
// turn's prompt reproduces the generated ids and the KV prefix cache hits like it does with a
// real tokenizer.  Chat template: ChatML, as benchmarks/multi-turn-chat-go/hack/chat-template.jinja.
struct Tokenizer {
  int vocab = 128256;
  int im_start() const { return vocab - 2; }
  int im_end() const { return vocab - 1; }

  std::string piece(int t) const {
    char b[6] = {' ', 0, 0, 0, 0, 0};
    int v = t;
    for (int i = 4; i >= 1; --i) { b[i] = static_cast<char>('a' + v % 26); v /= 26; }
    return std::string(b, 5);
  }
  void encode(const std::string& s, std::vector<int32_t>* out) const {
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
      if (s[i] == ' ' && i + 5 <= n) {
        int v = 0;
        bool ok = true;
        for (int k = 1; k <= 4; ++k) {
          char c = s[i + k];
          if (c < 'a' || c > 'z') { ok = false; break; }
          v = v * 26 + (c - 'a');
        }
        if (ok && (i + 5 == n || s[i + 5] < 'a' || s[i + 5] > 'z') && v < vocab) {
          out->push_back(v);
          i += 5;
          continue;
        }
      }
      out->push_back(static_cast<unsigned char>(s[i]));
      ++i;
    }
  }
  // <|im_start|>role\ncontent<|im_end|>\n ... <|im_start|>assistant\n
  void chat_prompt(const std::vector<std::pair<std::string, std::string>>& msgs, std::vector<int32_t>* out) const {
    for (auto& m : msgs) {
      out->push_back(im_start());
      encode(m.first, out);
      out->push_back('\n');
      encode(m.second, out);
      out->push_back(im_end());
      out->push_back('\n');
    }
    if (msgs.empty() || msgs.back().first != "assistant") {
      out->push_back(im_start());
      encode("assistant", out);
      out->push_back('\n');
    }
  }
};

inline double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return static_cast<double>(ts.tv_sec) + 1e-9 * static_cast<double>(ts.tv_nsec);
}

}  // namespace b200
