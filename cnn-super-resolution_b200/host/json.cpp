#include "json.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <ios>
#include <sstream>

namespace cnn_sr {
namespace json {

const Value* Value::find(const std::string& key) const {
  for (const auto& kv : object)
    if (kv.first == key) return &kv.second;
  return nullptr;
}

namespace {

struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& text) : s(text) {}

  [[noreturn]] void error(const char* what) const {
    std::ostringstream os;
    os << "Json parsing error: " << what << " in: '" << s.substr(i, 20) << "'";
    throw std::ios_base::failure(os.str());
  }

  void skip_ws() {
    while (i < s.size()) {
      const char c = s[i];
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
        ++i;
      } else {
        break;
      }
    }
  }

  Value value() {
    skip_ws();
    if (i >= s.size()) error("unexpected end of input");
    const char c = s[i];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      Value v;
      v.type = Type::String;
      v.string = string();
      return v;
    }
    if (c == '-' || std::isdigit((unsigned char)c)) return number();
    if (s.compare(i, 4, "true") == 0) {
      i += 4;
      Value v;
      v.type = Type::Bool;
      v.boolean = true;
      return v;
    }
    if (s.compare(i, 5, "false") == 0) {
      i += 5;
      Value v;
      v.type = Type::Bool;
      return v;
    }
    if (s.compare(i, 4, "null") == 0) {
      i += 4;
      return Value();
    }
    error("unexpected character");
  }

  Value number() {
    const char* start = s.c_str() + i;
    char* end = nullptr;
    const double d = std::strtod(start, &end);
    if (end == start) error("bad number");
    i += (size_t)(end - start);
    Value v;
    v.type = Type::Number;
    v.number = d;
    return v;
  }

  std::string string() {
    std::string out;
    ++i;  // opening quote
    while (true) {
      if (i >= s.size()) error("unterminated string");
      const char c = s[i++];
      if (c == '"') break;
      if (c == '\\') {
        if (i >= s.size()) error("unterminated escape");
        const char e = s[i++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u':
            if (i + 4 > s.size()) error("bad \\u escape");
            out += (char)std::strtol(s.substr(i, 4).c_str(), nullptr, 16);
            i += 4;
            break;
          default: out += e;  // \" \\ \/
        }
      } else {
        out += c;
      }
    }
    return out;
  }

  Value array() {
    Value v;
    v.type = Type::Array;
    ++i;
    skip_ws();
    if (i < s.size() && s[i] == ']') {
      ++i;
      return v;
    }
    while (true) {
      v.array.push_back(value());
      skip_ws();
      if (i >= s.size()) error("unterminated array");
      if (s[i] == ',') {
        ++i;
        continue;
      }
      if (s[i] == ']') {
        ++i;
        break;
      }
      error("expected ',' or ']'");
    }
    return v;
  }

  Value object() {
    Value v;
    v.type = Type::Object;
    ++i;
    skip_ws();
    if (i < s.size() && s[i] == '}') {
      ++i;
      return v;
    }
    while (true) {
      skip_ws();
      if (i >= s.size() || s[i] != '"') error("expected a key");
      std::string key = string();
      skip_ws();
      if (i >= s.size() || s[i] != ':') error("expected ':'");
      ++i;
      v.object.emplace_back(std::move(key), value());
      skip_ws();
      if (i >= s.size()) error("unterminated object");
      if (s[i] == ',') {
        ++i;
        continue;
      }
      if (s[i] == '}') {
        ++i;
        break;
      }
      error("expected ',' or '}'");
    }
    return v;
  }
};

}  // namespace

Value parse(const std::string& text) {
  Parser p(text);
  Value v = p.value();
  p.skip_ws();
  if (p.i != text.size()) p.error("trailing characters");
  return v;
}

Value parse_file(const char* path) {
  if (std::strlen(path) > 250) throw std::ios_base::failure("Filepath is too long");
  std::ifstream f(path);
  if (!f.is_open()) throw std::ios_base::failure("File not found");
  std::stringstream ss;
  ss << f.rdbuf();
  return parse(ss.str());
}

}  // namespace json
}  // namespace cnn_sr
