#include "Config.hpp"

#include <cmath>
#include <stdexcept>
#include <vector>

#include "json.hpp"
#include "pch.hpp"

namespace cnn_sr {

ParametersDistribution::ParametersDistribution(float mean_w_, float mean_b_, float sd_w_,
                                               float sd_b_)
    : mean_w(mean_w_), sd_w(sd_w_), mean_b(mean_b_), sd_b(sd_b_) {}

Config::Config(size_t n1_, size_t n2_, size_t f1_, size_t f2_, size_t f3_, float momentum_,
               float weight_decay, float* learning_rates, ParametersDistribution pd1,
               ParametersDistribution pd2, ParametersDistribution pd3,
               const char* const parameters_file_)
    : n1(n1_), n2(n2_), f1(f1_), f2(f2_), f3(f3_), momentum(momentum_),
      weight_decay_parameter(weight_decay),
      parameters_file(parameters_file_ ? parameters_file_ : ""),
      params_distr_1(pd1), params_distr_2(pd2), params_distr_3(pd3) {
  for (int i = 0; i < 3; i++) learning_rate[i] = learning_rates[i];
}

size_t Config::total_padding() const { return f1 + f2 + f3 - 3; }

void Config::validate(Config& c) {
  using utils::is_odd;
  using utils::require;
  require(is_odd(c.f1), "f1 should be odd");
  require(is_odd(c.f2), "f2 should be odd");
  require(is_odd(c.f3), "f3 should be odd");
  require(c.n1 > 0, "n1 should be >0");
  require(c.n2 > 0, "n2 should be >0");
  require(c.f1 > 0, "f1 should be >0");
  require(c.f2 > 0, "f2 should be >0");
  require(c.f3 > 0, "f3 should be >0");
  require(c.weight_decay_parameter >= 0, "weight_decay should be >0");
  require(c.learning_rate[0] > 0 && c.learning_rate[1] > 0 && c.learning_rate[2] > 0,
          "All learning rates should be >0");
  for (ParametersDistribution* pd : {&c.params_distr_1, &c.params_distr_2, &c.params_distr_3}) {
    require(pd->sd_w > 0, "std dev. for weights should be > 0");
    require(pd->sd_b >= 0, "std dev. for bias should be >= 0");
  }
}

namespace {

void read_distribution(const json::Value& node, ParametersDistribution& pd) {
  if (!node.is(json::Type::Object)) return;
  auto num = [&](const char* key, float& target) {
    const json::Value* v = node.find(key);
    if (v && v->is(json::Type::Number)) target = (float)v->number;
  };
  num("mean_w", pd.mean_w);
  num("mean_b", pd.mean_b);
  num("std_deviation_w", pd.sd_w);
  num("std_deviation_b", pd.sd_b);
  // the reference takes absolute values of all four (src/Config.cpp:87-92)
  pd.mean_w = std::fabs(pd.mean_w);
  pd.mean_b = std::fabs(pd.mean_b);
  pd.sd_w = std::fabs(pd.sd_w);
  pd.sd_b = std::fabs(pd.sd_b);
}

}  // namespace

Config ConfigReader::read(const char* const file) {
  const json::Value root = json::parse_file(file);
  if (!root.is(json::Type::Object))
    throw std::runtime_error("Expected root of JSON file had invalid type");

  size_t n1 = 0, n2 = 0, f1 = 0, f2 = 0, f3 = 0;
  float momentum = 0.f, weight_decay = 0.f;
  std::string parameters_file;
  std::vector<float> lrs;
  ParametersDistribution pd1, pd2, pd3;
  bool seen1 = false, seen2 = false, seen3 = false;

  // unknown keys are ignored, like the reference
  for (const auto& kv : root.object) {
    const std::string& key = kv.first;
    const json::Value& v = kv.second;
    const bool is_num = v.is(json::Type::Number);
    if (key == "n1" && is_num) n1 = (size_t)(unsigned int)v.number;
    else if (key == "n2" && is_num) n2 = (size_t)(unsigned int)v.number;
    else if (key == "f1" && is_num) f1 = (size_t)(unsigned int)v.number;
    else if (key == "f2" && is_num) f2 = (size_t)(unsigned int)v.number;
    else if (key == "f3" && is_num) f3 = (size_t)(unsigned int)v.number;
    else if (key == "momentum" && is_num) momentum = (float)v.number;
    else if (key == "weight_decay_parameter" && is_num) weight_decay = (float)v.number;
    else if (key == "parameters_file" && v.is(json::Type::String)) parameters_file = v.string;
    else if (key == "learning_rates" && v.is(json::Type::Array)) {
      for (const json::Value& e : v.array) lrs.push_back((float)e.number);
    } else if (key == "parameters_distribution_1") { read_distribution(v, pd1); seen1 = true; }
    else if (key == "parameters_distribution_2") { read_distribution(v, pd2); seen2 = true; }
    else if (key == "parameters_distribution_3") { read_distribution(v, pd3); seen3 = true; }
  }
  (void)seen1; (void)seen2; (void)seen3;
  utils::require(lrs.size() == 3, "Expected 3 learning rates (one per layer) to be provided");

  Config cfg(n1, n2, f1, f2, f3, momentum, weight_decay, lrs.data(), pd1, pd2, pd3,
             parameters_file.c_str());
  Config::validate(cfg);
  return cfg;
}

}  // namespace cnn_sr

static std::ostream& operator<<(std::ostream& os, const cnn_sr::ParametersDistribution& pd) {
  os << "{ weights(" << pd.mean_w << ", " << pd.sd_w << "), bias(" << pd.mean_b << ", "
     << pd.sd_b << ")}";
  return os;
}

std::ostream& operator<<(std::ostream& os, const cnn_sr::Config& cfg) {
  os << "Config {" << std::endl
     << "  parameters file: '" << cfg.parameters_file << "'" << std::endl
     << "  momentum: " << cfg.momentum << std::endl
     << "  learning rates: { " << cfg.learning_rate[0] << ", " << cfg.learning_rate[1] << ", "
     << cfg.learning_rate[2] << "}" << std::endl
     << "  layer 1: " << cfg.n1 << " filters, " << cfg.f1 << " spatial size" << std::endl
     << "  layer 2: " << cfg.n2 << " filters, " << cfg.f2 << " spatial size" << std::endl
     << "  layer 3: " << cfg.f3 << " spatial size" << std::endl
     << "  parameters dist. 1 " << cfg.params_distr_1 << std::endl
     << "  parameters dist. 2 " << cfg.params_distr_2 << std::endl
     << "  parameters dist. 3 " << cfg.params_distr_3 << "}" << std::endl;
  return os;
}
