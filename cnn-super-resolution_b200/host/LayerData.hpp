// Shape of one convolutional layer + host copies of its parameters.
// Same public surface as the reference's cnn_sr::LayerData (src/LayerData.hpp:31-59).
#ifndef CNN_SR_LAYER_DATA_H
#define CNN_SR_LAYER_DATA_H

#include <cstddef>
#include <ostream>
#include <vector>

namespace cnn_sr {

struct LayerData {
  /** n_prev_filter_cnt: 1/n1/n2, current_filter_count: n1/n2/1, f_spatial_size: f1/f2/f3 */
  LayerData(size_t n_prev_filter_cnt, size_t current_filter_count, size_t f_spatial_size);

  /** throws std::runtime_error when weights/bias hold fewer values than the shape needs
   * (reference: src/LayerData.cpp:21-45) */
  static void validate(const LayerData&);

  /** append weight_size() / bias_size() values (nullptr: no-op), src/LayerData.cpp:51-57 */
  void set_weights(float*);
  void set_bias(float*);

  size_t input_size(size_t w, size_t h) const;                 // w*h*n_prev_filter_cnt
  void get_output_dimensions(size_t* wh, size_t w, size_t h) const;  // valid conv: in - f + 1
  size_t weight_size() const;                                  // f*f*k*n
  size_t bias_size() const;                                    // n
  const float* weights_ptr() const { return weights.data(); }
  const float* bias_ptr() const { return bias.data(); }

  const size_t n_prev_filter_cnt;
  const size_t current_filter_count;
  const size_t f_spatial_size;

  /** host copies; stale once uploaded -- the device copy is authoritative */
  std::vector<float> weights;
  std::vector<float> bias;
};

}  // namespace cnn_sr

std::ostream& operator<<(std::ostream&, const cnn_sr::LayerData&);
#endif
