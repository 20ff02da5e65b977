// Tiny JSON value / parser / writer for the serde wire format of the proof structs (serde_json layout:
// structs are objects with fields in declaration order, enums are externally tagged, u8/usize are numbers,
// BigInts are strings).  Objects keep insertion order so round trips are byte-identical.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace zkhost {

struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  int64_t num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  static Json string(std::string s) { Json j; j.kind = Str; j.str = std::move(s); return j; }
  static Json number(int64_t v) { Json j; j.kind = Num; j.num = v; return j; }
  static Json boolean(bool v) { Json j; j.kind = Bool; j.b = v; return j; }
  static Json array() { Json j; j.kind = Arr; return j; }
  static Json object() { Json j; j.kind = Obj; return j; }
  Json& set(const std::string& k, Json v) { obj.emplace_back(k, std::move(v)); return *this; }
  Json& push(Json v) { arr.push_back(std::move(v)); return *this; }
  const Json* find(const std::string& k) const {
    for (auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Json& at(const std::string& k) const {
    const Json* p = find(k);
    if (!p) throw std::runtime_error("missing field `" + k + "`");
    return *p;
  }
  const std::string& as_str() const { if (kind != Str) throw std::runtime_error("expected a string"); return str; }
  int64_t as_num() const { if (kind != Num) throw std::runtime_error("expected a number"); return num; }

  void write(std::string& out) const {
    switch (kind) {
      case Null: out += "null"; break;
      case Bool: out += b ? "true" : "false"; break;
      case Num: out += std::to_string(num); break;
      case Str:
        out.push_back('"');
        for (char c : str) {
          if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back(c); }
          else if (c == '\n') out += "\\n";
          else out.push_back(c);
        }
        out.push_back('"');
        break;
      case Arr:
        out.push_back('[');
        for (size_t i = 0; i < arr.size(); ++i) { if (i) out.push_back(','); arr[i].write(out); }
        out.push_back(']');
        break;
      case Obj:
        out.push_back('{');
        for (size_t i = 0; i < obj.size(); ++i) {
          if (i) out.push_back(',');
          Json::string(obj[i].first).write(out);
          out.push_back(':');
          obj[i].second.write(out);
        }
        out.push_back('}');
        break;
    }
  }
  std::string dump() const { std::string s; write(s); return s; }

  static Json parse(const std::string& s) {
    size_t i = 0;
    Json j = parse_value(s, i, 0);
    skip(s, i);
    if (i != s.size()) throw std::runtime_error("trailing characters after JSON value");
    return j;
  }

 private:
  static void skip(const std::string& s, size_t& i) { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
  static constexpr int kMaxDepth = 64;  // proofs nest four levels deep; untrusted input must not recurse the stack away
  static Json parse_value(const std::string& s, size_t& i, int depth) {
    if (depth > kMaxDepth) throw std::runtime_error("JSON nested deeper than 64 levels");
    skip(s, i);
    if (i >= s.size()) throw std::runtime_error("unexpected end of JSON");
    char c = s[i];
    if (c == '{') {
      Json j = object();
      ++i; skip(s, i);
      if (i < s.size() && s[i] == '}') { ++i; return j; }
      for (;;) {
        skip(s, i);
        Json k = parse_value(s, i, depth + 1);
        if (k.kind != Str) throw std::runtime_error("object key must be a string");
        skip(s, i);
        if (i >= s.size() || s[i] != ':') throw std::runtime_error("expected ':'");
        ++i;
        j.obj.emplace_back(k.str, parse_value(s, i, depth + 1));
        skip(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == '}') { ++i; return j; }
        throw std::runtime_error("expected ',' or '}'");
      }
    }
    if (c == '[') {
      Json j = array();
      ++i; skip(s, i);
      if (i < s.size() && s[i] == ']') { ++i; return j; }
      for (;;) {
        j.arr.push_back(parse_value(s, i, depth + 1));
        skip(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == ']') { ++i; return j; }
        throw std::runtime_error("expected ',' or ']'");
      }
    }
    if (c == '"') {
      Json j; j.kind = Str;
      ++i;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) { ++i; j.str.push_back(s[i] == 'n' ? '\n' : s[i]); }
        else j.str.push_back(s[i]);
        ++i;
      }
      if (i >= s.size()) throw std::runtime_error("unterminated string");
      ++i;
      return j;
    }
    if (s.compare(i, 4, "true") == 0) { i += 4; return boolean(true); }
    if (s.compare(i, 5, "false") == 0) { i += 5; return boolean(false); }
    if (s.compare(i, 4, "null") == 0) { i += 4; return Json(); }
    size_t st = i;
    if (c == '-') ++i;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') ++i;
    if (i == st) throw std::runtime_error("unexpected character in JSON");
    return number(std::stoll(s.substr(st, i - st)));
  }
};

}  // namespace zkhost
