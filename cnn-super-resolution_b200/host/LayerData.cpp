#include "LayerData.hpp"

#include <sstream>
#include <stdexcept>

namespace cnn_sr {

LayerData::LayerData(size_t k, size_t n, size_t f)
    : n_prev_filter_cnt(k), current_filter_count(n), f_spatial_size(f) {
  // only capacity: validate() must still be able to catch vectors that were never filled
  weights.reserve(weight_size());
  bias.reserve(bias_size());
}

void LayerData::validate(const LayerData& d) {
  if (d.weights.size() < d.weight_size()) {
    std::ostringstream os;
    os << "Declared f_spatial_size(" << d.f_spatial_size << ")*f_spatial_size("
       << d.f_spatial_size << ")*n_prev_filter_cnt(" << d.n_prev_filter_cnt
       << ")*current_filter_count(" << d.current_filter_count << ")=" << d.weight_size()
       << " is bigger then weights array (" << d.weights.size()
       << " elements). Expected more elements in weights array. ";
    throw std::runtime_error(os.str());
  }
  if (d.bias.size() < d.bias_size()) {
    std::ostringstream os;
    os << "Bias array(size=" << d.bias.size() << ") should have equal size to "
       << "current_filter_count(" << d.bias_size() << ").";
    throw std::runtime_error(os.str());
  }
}

void LayerData::set_weights(float* x) {
  if (x) weights.insert(weights.end(), x, x + weight_size());
}

void LayerData::set_bias(float* x) {
  if (x) bias.insert(bias.end(), x, x + bias_size());
}

void LayerData::get_output_dimensions(size_t* wh, size_t w, size_t h) const {
  wh[0] = w - f_spatial_size + 1;
  wh[1] = h - f_spatial_size + 1;
}

size_t LayerData::weight_size() const {
  return f_spatial_size * f_spatial_size * n_prev_filter_cnt * current_filter_count;
}

size_t LayerData::bias_size() const { return current_filter_count; }

size_t LayerData::input_size(size_t w, size_t h) const { return w * h * n_prev_filter_cnt; }

}  // namespace cnn_sr

std::ostream& operator<<(std::ostream& os, const cnn_sr::LayerData& d) {
  os << "Layer { previous filters: " << d.n_prev_filter_cnt
     << ", current filters: " << d.current_filter_count
     << ", f_spatial_size: " << d.f_spatial_size << ", weighs.size: " << d.weights.size()
     << ", bias.size: " << d.bias.size() << "}";
  return os;
}
