// Host utilities shared by the pipeline classes and the CLI.  Mirrors the surface of the
// reference's src/pch.hpp:36-136 (require, dump_vector, closest_power_of_2, list_files, the
// hand-rolled Argparse) so callers written against the reference compile unchanged.
#ifndef CNN_SR_PCH_H
#define CNN_SR_PCH_H

#include <cstddef>
#include <ios>
#include <ostream>
#include <string>
#include <utility>
#include <vector>

namespace cnn_sr {

extern bool warn_about_blocking_operation;

namespace utils {

#define IOException std::ios_base::failure
#define STRINGIFY2(s) #s
#define STRINGIFY(s) STRINGIFY2(s)

/** throws std::runtime_error(msg) when the check fails (reference: src/pch.cpp:23-27) */
void require(bool check, const char* msg);

/** "a, b, c" with the stream's default precision = 6 significant digits (quirk Q6;
 * reference: src/pch.cpp:29-54) */
void dump_vector(std::ostream&, std::vector<float>&, const char* line_prefix = nullptr,
                 size_t per_line = 0, bool add_line_numbers = false);

template <typename T>
inline bool is_odd(T x) {
  return (x & 1) != 0;
}
template <typename T>
inline bool is_even(T x) {
  return !is_odd(x);
}

size_t closest_power_of_2(int);

/** names of all directory entries, like readdir (reference: src/pch.cpp:80-97) */
void list_files(const char* path, std::vector<std::string>& target);

/// Command line: positional words (`train`, `dry`, `profile`, `help`) are flags; options whose
/// first mnemonic starts with '-' take the next argv as their value; unknown arguments only
/// warn (reference: src/pch.cpp:183-299).
struct ArgOption {
  bool _required = false;
  std::string _name;
  std::string _help;
  std::vector<std::string> _mnemonics;
  ArgOption& help(const char*);
  ArgOption& required();
};

class Argparse {
 public:
  Argparse(const char* exec_name, const char* general_help);
  ArgOption& add_argument(const char*);
  ArgOption& add_argument(const char*, const char*);
  /** false when only help was requested; throws std::runtime_error when a required option
   * is missing */
  bool parse(size_t argc, char** argv);
  void print_help();
  bool has_arg(const char* name);
  /** nullptr when absent */
  const char* value(const char* name);
  /** leaves `target` untouched when absent */
  void value(const char* name, size_t& target);

 private:
  typedef std::pair<size_t, std::string> ArgValue;  // option index, value
  ArgOption& add(std::vector<std::string> mnemonics);
  const ArgValue* get(const char* name);
  std::string _general_help, _exec_name;
  std::vector<ArgOption> _options;
  std::vector<ArgValue> _values;
};

}  // namespace utils
}  // namespace cnn_sr
#endif
