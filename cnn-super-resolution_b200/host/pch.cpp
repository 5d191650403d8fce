#include "pch.hpp"

#include <dirent.h>

#include <cctype>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

namespace cnn_sr {

bool warn_about_blocking_operation = false;

namespace utils {

void require(bool check, const char* msg) {
  if (!check) throw std::runtime_error(msg);
}

void dump_vector(std::ostream& os, std::vector<float>& data, const char* line_prefix,
                 size_t per_line, bool add_line_numbers) {
  const size_t len = data.size();
  size_t lines = 1;
  if (per_line == 0)
    per_line = len;
  else
    lines = len / per_line;
  const std::string prefix = line_prefix ? line_prefix : "";
  for (size_t row = 0; row < lines; row++) {
    os << prefix;
    if (add_line_numbers) os << "[" << row << "] ";
    for (size_t col = 0; col < per_line; col++) {
      const size_t idx = row * per_line + col;
      if (idx < len) os << data[idx];
      if (col + 1 < per_line) os << ", ";
    }
    if (row + 1 < lines) os << std::endl;
  }
}

size_t closest_power_of_2(int x) {
  if (x < 0) return 0;
  size_t p = 1;
  while (p < (size_t)x) p <<= 1;
  return x == 0 ? 0 : p;
}

void list_files(const char* path, std::vector<std::string>& target) {
  DIR* d = opendir(path);
  if (!d) return;
  while (struct dirent* e = readdir(d)) target.push_back(e->d_name);
  closedir(d);
}

ArgOption& ArgOption::help(const char* text) {
  _help = text;
  return *this;
}

ArgOption& ArgOption::required() {
  _required = true;
  return *this;
}

Argparse::Argparse(const char* exec, const char* help) : _general_help(help), _exec_name(exec) {
  add_argument("help", "-h").help("Print this help");
}

ArgOption& Argparse::add_argument(const char* m) { return add({m}); }
ArgOption& Argparse::add_argument(const char* m1, const char* m2) { return add({m1, m2}); }

ArgOption& Argparse::add(std::vector<std::string> mnemonics) {
  ArgOption opt;
  for (const std::string& m : mnemonics) {
    bool valid = !m.empty();
    for (char c : m) valid = valid && (std::isalpha((unsigned char)c) || c == '-');
    if (!valid) {
      std::cout << "'" << m << "' is not valid mnemonic" << std::endl;
      continue;
    }
    if (m[0] != '-')
      opt._name = m;
    else if (m.size() > 2 && m[1] == '-')
      opt._name = m.substr(2);
    opt._mnemonics.push_back(m);
  }
  if (opt._mnemonics.empty()) throw std::runtime_error("Argument does not have valid mnemonic");
  if (opt._name.empty())
    throw std::runtime_error(
        "Argument does not have valid name (at least one mnemonic should: not have '-' prefix "
        "or start with '--')");
  _options.push_back(opt);
  return _options.back();
}

bool Argparse::parse(size_t argc, char** argv) {
  _values.clear();
  for (size_t i = 1; i < argc; i++) {
    const std::string arg = argv[i];
    size_t found = _options.size();
    for (size_t o = 0; o < _options.size() && found == _options.size(); o++)
      for (const std::string& m : _options[o]._mnemonics)
        if (m == arg) found = o;
    if (found == _options.size()) {
      std::cout << "Unrecognised argument: '" << arg << "'" << std::endl;
      continue;
    }
    const bool takes_value = _options[found]._mnemonics[0][0] == '-';
    if (!takes_value) {
      _values.emplace_back(found, "");
    } else if (i + 1 < argc) {
      _values.emplace_back(found, argv[++i]);
    } else {
      std::cout << "Expected value for: '" << _options[found]._name << "'" << std::endl;
    }
  }
  if (has_arg("help")) {
    print_help();
    return false;
  }
  for (size_t o = 0; o < _options.size(); o++) {
    if (!_options[o]._required) continue;
    bool provided = false;
    for (const ArgValue& v : _values) provided = provided || v.first == o;
    if (!provided)
      throw std::runtime_error("Value not provided for argument: '" + _options[o]._name + "'");
  }
  return true;
}

void Argparse::print_help() {
  std::cout << "Usage: " << _exec_name;
  for (const ArgOption& o : _options) {
    const bool flag = o._mnemonics[0][0] != '-';
    std::cout << " " << (o._required ? "" : "[") << o._mnemonics[0] << (flag ? "" : " VALUE")
              << (o._required ? "" : "]");
  }
  std::cout << std::endl << _general_help << std::endl << std::endl << "arguments:" << std::endl;
  for (const ArgOption& o : _options) {
    std::cout << "  ";
    for (size_t i = 0; i < o._mnemonics.size(); i++)
      std::cout << (i ? ", " : "") << o._mnemonics[i];
    std::cout << "  " << o._help << std::endl;
  }
}

const Argparse::ArgValue* Argparse::get(const char* name) {
  for (const ArgValue& v : _values)
    if (_options[v.first]._name == name) return &v;
  return nullptr;
}

bool Argparse::has_arg(const char* name) { return get(name) != nullptr; }

const char* Argparse::value(const char* name) {
  const ArgValue* v = get(name);
  return v ? v->second.c_str() : nullptr;
}

void Argparse::value(const char* name, size_t& target) {
  const ArgValue* v = get(name);
  if (v) target = (size_t)std::strtoull(v->second.c_str(), nullptr, 10);
}

}  // namespace utils
}  // namespace cnn_sr
