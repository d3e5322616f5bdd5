// Minimal JSON-with-comments reader for the reference's cfg/*.json files (the reference reads
// them with jsoncpp, /root/reference/common/utils.cpp LoadJson).  Supports // and /* */
// comments, objects, arrays, numbers, strings, true/false/null.  Host-only.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace xb {

class Json {
 public:
  enum Type { Null, Bool, Num, Str, Arr, Obj };
  Type type = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::map<std::string, Json> obj;

  static Json parse(const std::string& text) {
    size_t i = 0;
    Json v = parse_value(text, i);
    skip(text, i);
    if (i != text.size()) throw std::runtime_error("json: trailing characters");
    return v;
  }
  bool isNull() const { return type == Null; }
  bool isMember(const std::string& k) const { return type == Obj && obj.count(k); }
  const Json& operator[](const std::string& k) const {
    static const Json null_;
    if (type != Obj) return null_;
    auto it = obj.find(k);
    return it == obj.end() ? null_ : it->second;
  }
  const Json& operator[](size_t i) const { return arr.at(i); }
  size_t size() const { return type == Arr ? arr.size() : (type == Obj ? obj.size() : 0); }
  double asDouble() const {
    if (type == Num) return num;
    if (type == Bool) return b ? 1.0 : 0.0;
    throw std::runtime_error("json: value is not a number");
  }
  int asInt() const { return (int)std::llround(asDouble()); }
  bool asBool() const {
    if (type == Bool) return b;
    if (type == Num) return num != 0;
    throw std::runtime_error("json: value is not a bool");
  }
  const std::string& asString() const {
    if (type != Str) throw std::runtime_error("json: value is not a string");
    return str;
  }
  // jsoncpp-style get-with-default
  double get(const std::string& k, double d) const { return isMember(k) ? (*this)[k].asDouble() : d; }
  int get(const std::string& k, int d) const { return isMember(k) ? (*this)[k].asInt() : d; }
  bool get(const std::string& k, bool d) const { return isMember(k) ? (*this)[k].asBool() : d; }
  std::string get(const std::string& k, const char* d) const { return isMember(k) ? (*this)[k].asString() : std::string(d); }
  std::vector<double> vec(const std::string& k) const {
    const Json& v = (*this)[k];
    std::vector<double> out;
    if (v.type == Num) {
      out.push_back(v.num);
      return out;
    }
    if (v.type != Arr) throw std::runtime_error("json: '" + k + "' is not an array");
    for (const auto& e : v.arr) {
      if (e.type == Arr)
        for (const auto& f : e.arr) out.push_back(f.asDouble());
      else
        out.push_back(e.asDouble());
    }
    return out;
  }

 private:
  static void skip(const std::string& s, size_t& i) {
    for (;;) {
      while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
      if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
        while (i < s.size() && s[i] != '\n') ++i;
      } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
        i += 2;
        while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) ++i;
        i += 2;
      } else {
        return;
      }
    }
  }
  static Json parse_value(const std::string& s, size_t& i) {
    skip(s, i);
    if (i >= s.size()) throw std::runtime_error("json: unexpected end");
    Json v;
    char c = s[i];
    if (c == '{') {
      v.type = Obj;
      ++i;
      skip(s, i);
      if (i < s.size() && s[i] == '}') { ++i; return v; }
      for (;;) {
        skip(s, i);
        if (i >= s.size() || s[i] != '"') throw std::runtime_error("json: expected key");
        std::string k = parse_string(s, i);
        skip(s, i);
        if (i >= s.size() || s[i] != ':') throw std::runtime_error("json: expected ':'");
        ++i;
        v.obj[k] = parse_value(s, i);
        skip(s, i);
        if (i < s.size() && s[i] == ',') {
          ++i;
          skip(s, i);
          if (i < s.size() && s[i] == '}') { ++i; return v; }  // tolerate trailing comma
          continue;
        }
        if (i < s.size() && s[i] == '}') { ++i; return v; }
        throw std::runtime_error("json: expected ',' or '}'");
      }
    }
    if (c == '[') {
      v.type = Arr;
      ++i;
      skip(s, i);
      if (i < s.size() && s[i] == ']') { ++i; return v; }
      for (;;) {
        v.arr.push_back(parse_value(s, i));
        skip(s, i);
        if (i < s.size() && s[i] == ',') {
          ++i;
          skip(s, i);
          if (i < s.size() && s[i] == ']') { ++i; return v; }
          continue;
        }
        if (i < s.size() && s[i] == ']') { ++i; return v; }
        throw std::runtime_error("json: expected ',' or ']'");
      }
    }
    if (c == '"') {
      v.type = Str;
      v.str = parse_string(s, i);
      return v;
    }
    if (s.compare(i, 4, "true") == 0) { v.type = Bool; v.b = true; i += 4; return v; }
    if (s.compare(i, 5, "false") == 0) { v.type = Bool; v.b = false; i += 5; return v; }
    if (s.compare(i, 4, "null") == 0) { i += 4; return v; }
    char* end = nullptr;
    v.num = std::strtod(s.c_str() + i, &end);
    if (end == s.c_str() + i) throw std::runtime_error("json: bad token near '" + s.substr(i, 12) + "'");
    v.type = Num;
    i = end - s.c_str();
    return v;
  }
  static std::string parse_string(const std::string& s, size_t& i) {
    std::string out;
    ++i;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        ++i;
        switch (s[i]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          default: out += s[i];
        }
      } else {
        out += s[i];
      }
      ++i;
    }
    if (i >= s.size()) throw std::runtime_error("json: unterminated string");
    ++i;
    return out;
  }
};

}  // namespace xb
