// config.json reader/validator.  Same keys, defaults and failure behaviour as the reference's
// cnn_sr::Config / ConfigReader (src/Config.hpp:9-44, src/Config.cpp:46-147).
#ifndef CNN_SR_CONFIG_H
#define CNN_SR_CONFIG_H

#include <cstddef>
#include <ostream>
#include <string>

namespace cnn_sr {

struct ParametersDistribution {
  ParametersDistribution() {}
  ParametersDistribution(float mean_w, float mean_b, float sd_w, float sd_b);
  float mean_w = 0.01f, sd_w = 0.01f;
  float mean_b = 0.0f, sd_b = 0.0f;
};

struct Config {
  Config(size_t n1, size_t n2, size_t f1, size_t f2, size_t f3, float momentum,
         float weight_decay, float* learning_rates, ParametersDistribution,
         ParametersDistribution, ParametersDistribution, const char* const parameters_file = nullptr);

  /** throws std::runtime_error: odd f*, n* > 0, weight_decay >= 0, lr > 0, sd_w > 0, sd_b >= 0 */
  static void validate(Config&);

  /** f1 + f2 + f3 - 3: pixels lost by the three valid convolutions */
  size_t total_padding() const;

  const size_t n1, n2;
  const size_t f1, f2, f3;
  const float momentum, weight_decay_parameter;
  float learning_rate[3];
  std::string parameters_file = "";
  ParametersDistribution params_distr_1, params_distr_2, params_distr_3;
};

class ConfigReader {
 public:
  /** throws IOException (std::ios_base::failure) for a missing/unparsable file,
   * std::runtime_error for invalid values */
  Config read(const char* const file);
};

}  // namespace cnn_sr

std::ostream& operator<<(std::ostream&, const cnn_sr::Config&);
#endif
