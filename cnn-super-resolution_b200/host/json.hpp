// Minimal JSON reader for config.json / parameters.json (objects, arrays, numbers, strings,
// true/false/null).  The reference vendors gason (libs/include/json/gason.h); the formats are
// what must be kept, not the parser.
#ifndef CNN_SR_JSON_H
#define CNN_SR_JSON_H

#include <string>
#include <utility>
#include <vector>

namespace cnn_sr {
namespace json {

enum class Type { Null, Bool, Number, String, Array, Object };

struct Value {
  Type type = Type::Null;
  double number = 0.0;
  bool boolean = false;
  std::string string;
  std::vector<Value> array;
  std::vector<std::pair<std::string, Value>> object;  // insertion order kept

  bool is(Type t) const { return type == t; }
  const Value* find(const std::string& key) const;
};

/** Parses `text`; throws std::ios_base::failure (the reference's IOException) on a syntax
 * error, naming the offending position. */
Value parse(const std::string& text);

/** Reads and parses a file; throws std::ios_base::failure when it cannot be read/parsed. */
Value parse_file(const char* path);

}  // namespace json
}  // namespace cnn_sr
#endif
