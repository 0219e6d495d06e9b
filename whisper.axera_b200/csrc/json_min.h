// Minimal parser for the flat JSON object the reference writes as {type}_config.json
// (/root/reference/model_convert/export_onnx.py:592-629: string and integer values only, json.dump(indent=4)).
#pragma once
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

namespace b200w {

struct JsonFlat {
  std::map<std::string, std::string> kv;  // numbers kept as their literal text
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  const std::string& get_str(const std::string& k) const {
    auto it = kv.find(k);
    if (it == kv.end()) throw std::runtime_error("config: missing key '" + k + "'");
    return it->second;
  }
  int get_int(const std::string& k) const {
    const std::string& s = get_str(k);
    char* end = nullptr;
    long v = strtol(s.c_str(), &end, 10);
    if (end == s.c_str()) throw std::runtime_error("config: key '" + k + "' is not an integer");
    return (int)v;
  }
};

inline JsonFlat parse_flat_json(const std::string& text) {
  JsonFlat out;
  size_t i = 0;
  const size_t n = text.size();
  auto skip_ws = [&]() {
    while (i < n && (text[i] == ' ' || text[i] == '\n' || text[i] == '\r' || text[i] == '\t')) ++i;
  };
  auto parse_string = [&]() {
    std::string s;
    if (text[i] != '"') throw std::runtime_error("config: expected string");
    ++i;
    while (i < n && text[i] != '"') {
      if (text[i] == '\\' && i + 1 < n) {
        ++i;
        switch (text[i]) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'u':  // keep \uXXXX escapes verbatim; no key we read contains them
            s += "\\u";
            break;
          default: s += text[i];
        }
      } else {
        s += text[i];
      }
      ++i;
    }
    if (i >= n) throw std::runtime_error("config: unterminated string");
    ++i;
    return s;
  };
  skip_ws();
  if (i >= n || text[i] != '{') throw std::runtime_error("config: expected '{'");
  ++i;
  while (true) {
    skip_ws();
    if (i < n && text[i] == '}') break;
    std::string key = parse_string();
    skip_ws();
    if (i >= n || text[i] != ':') throw std::runtime_error("config: expected ':'");
    ++i;
    skip_ws();
    if (i >= n) throw std::runtime_error("config: truncated");
    std::string val;
    if (text[i] == '"') {
      val = parse_string();
    } else if (text[i] == '{' || text[i] == '[') {
      throw std::runtime_error("config: nested values are not part of the reference's config format");
    } else {
      while (i < n && text[i] != ',' && text[i] != '}' && text[i] != '\n' && text[i] != ' ') val += text[i++];
    }
    out.kv[key] = val;
    skip_ws();
    if (i < n && text[i] == ',') {
      ++i;
      continue;
    }
    skip_ws();
    if (i < n && text[i] == '}') break;
    throw std::runtime_error("config: expected ',' or '}'");
  }
  return out;
}

}  // namespace b200w
