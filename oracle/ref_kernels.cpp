// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// oracle/_ref/libsrcnn_ref.so : the reference's OWN OpenCL kernels, compiled
// unmodified by g++ through oracle/cl_shim.hpp, plus host loops that stand in
// for clEnqueueNDRangeKernel.  The .cl sources are #include'd from
// /root/reference/src/kernel at build time (oracle/Makefile passes -I); nothing
// from the reference is copied into this repository.
//
// What is "reference" and what is "restatement" here:
//   * kernel arithmetic (forward, deltas, backpropagate, last_layer_delta,
//     update_params, sub_from_all)            -> the reference's own source
//   * NDRange iteration                        -> ours (only in-range work items
//     are visited; the reference's idle out-of-range items return immediately,
//     layer_uber_kernel.cl:65-67 etc.)
//   * squared_err / sum                        -> restated as a double sum, as the
//     reference's own test does (test/specs/SquaredErrorTest.cpp:53-60); the
//     originals need __local memory + barriers (squared_error.cl:74-86)
//   * forward/backward/update sequencing       -> restated from
//     src/ConfigBasedDataPipeline.cpp:200-361
//
// Work-groups are spread over host cores with OpenMP.  To keep the result
// deterministic the sample dimension of `backpropagate` is serialised: the
// reference's plain `target_grad_w[id] += grad_w` (backpropagate.cl:110) races
// across samples on a real device; serial order is the race-free definition.
#include "cl_shim.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace clshim {
thread_local size_t gid[3] = {0, 0, 0};
// run-time stand-ins for the -D macros (generic / VLA instantiation)
thread_local size_t rt_n = 0, rt_k = 0, rt_f = 0;
}  // namespace clshim

typedef void (*FwdFn)(float*, float*, float*, float*, uint, uint);
typedef void (*DltFn)(float*, float*, float*, float*, uint, uint, uint, uint, uint);
struct FwdEntry {
  int k, n, f, skip;
  FwdFn fn;
};
struct DltEntry {
  int n;
  DltFn fn;
};

// ---- compile-time specialisations (generated) --------------------------------
#include "ref_inst.inc"

// ---- run-time (VLA) instantiations for any other shape -----------------------
#define CURRENT_FILTER_COUNT (clshim::rt_n)
#define PREVIOUS_FILTER_COUNT (clshim::rt_k)
#define F_SPATIAL_SIZE (clshim::rt_f)
namespace fwd_generic_relu {
#include "layer_uber_kernel.cl"
}
#define SKIP_RELU
namespace fwd_generic_lin {
#include "layer_uber_kernel.cl"
}
#undef SKIP_RELU
namespace dlt_generic {
#include "layer_deltas.cl"
}
#undef CURRENT_FILTER_COUNT
#undef PREVIOUS_FILTER_COUNT
#undef F_SPATIAL_SIZE

// ---- kernels without macros ---------------------------------------------------
namespace k_backprop {
#include "backpropagate.cl"
}
namespace k_lld {
#include "last_layer_delta.cl"
}
namespace k_update {
#include "update_parameters.cl"
}
namespace k_sub {
#include "subtract_from_all.cl"
}

extern "C" {

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// kernel `forward` (layer_uber_kernel.cl:36-96); host side
// DataPipeline::execute_layer (src/DataPipeline.cpp:358-410)
void ref_forward(const float* in, float* out, const float* W, const float* B,
                 int k, int n, int f, int skip_relu, int in_w, int in_h, int S) {
  FwdFn fn = nullptr;
  for (const FwdEntry& e : g_fwd_table)
    if (e.k == k && e.n == n && e.f == f && e.skip == (skip_relu ? 1 : 0)) fn = e.fn;
  const bool generic = (fn == nullptr);
  if (generic) fn = skip_relu ? &fwd_generic_lin::forward : &fwd_generic_relu::forward;
  const int ow = in_w - f + 1, oh = in_h - f + 1;
  if (ow <= 0 || oh <= 0) return;
#pragma omp parallel for collapse(2) schedule(static)
  for (int s = 0; s < S; s++) {
    for (int y = 0; y < oh; y++) {
      if (generic) {
        clshim::rt_n = n;
        clshim::rt_k = k;
        clshim::rt_f = f;
      }
      clshim::gid[2] = s;
      clshim::gid[1] = y;
      for (int x = 0; x < ow; x++) {
        clshim::gid[0] = x;
        fn(const_cast<float*>(in), out, const_cast<float*>(W), const_cast<float*>(B),
           (uint)in_w, (uint)in_h);
      }
    }
  }
}

// kernel `last_layer_delta` (last_layer_delta.cl:14-50)
void ref_last_layer_delta(const float* gt, const float* algo, float* target,
                          int gt_w, int gt_h, int algo_w, int algo_h, int S) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int s = 0; s < S; s++) {
    for (int y = 0; y < algo_h; y++) {
      clshim::gid[2] = s;
      clshim::gid[1] = y;
      for (int x = 0; x < algo_w; x++) {
        clshim::gid[0] = x;
        k_lld::last_layer_delta(const_cast<float*>(gt), const_cast<float*>(algo), target,
                                (uint)gt_w, (uint)gt_h, (uint)algo_w, (uint)algo_h);
      }
    }
  }
}

// squared_err (squared_error.cl:36-92), restated as a double-precision sum
double ref_squared_error(const float* gt, const float* algo, int gt_w, int gt_h,
                         int algo_w, int algo_h, int S) {
  const size_t pad = (size_t)(gt_w - algo_w) / 2;
  double acc = 0.0;
  for (int s = 0; s < S; s++)
    for (int y = 0; y < algo_h; y++)
      for (int x = 0; x < algo_w; x++) {
        float t = gt[(size_t)s * gt_w * gt_h + (y + pad) * gt_w + pad + x];
        float v = algo[(size_t)s * algo_w * algo_h + (size_t)y * algo_w + x];
        float d = v - t;
        acc += (double)(d * d);
      }
  return acc;
}

// kernel `deltas` (layer_deltas.cl:42-127); host side
// DataPipeline::calculate_deltas (src/DataPipeline.cpp:522-594)
void ref_deltas(const float* deltas_next, const float* layer_output, float* target,
                const float* W, int n_curr, int f_curr, int f_next, int n_next,
                int out_w, int out_h, int S) {
  DltFn fn = nullptr;
  for (const DltEntry& e : g_dlt_table)
    if (e.n == n_curr) fn = e.fn;
  const bool generic = (fn == nullptr);
  if (generic) fn = &dlt_generic::deltas;
#pragma omp parallel for collapse(2) schedule(static)
  for (int s = 0; s < S; s++) {
    for (int y = 0; y < out_h; y++) {
      if (generic) clshim::rt_n = n_curr;
      clshim::gid[2] = s;
      clshim::gid[1] = y;
      for (int x = 0; x < out_w; x++) {
        clshim::gid[0] = x;
        fn(const_cast<float*>(deltas_next), const_cast<float*>(layer_output), target,
           const_cast<float*>(W), (uint)f_curr, (uint)f_next, (uint)n_next, (uint)out_w,
           (uint)out_h);
      }
    }
  }
}

// kernel `backpropagate` (backpropagate.cl:56-114); host side
// DataPipeline::backpropagate (src/DataPipeline.cpp:596-663).  grad_w/grad_b
// ACCUMULATE (`+=`), exactly like the reference.
void ref_backpropagate(const float* deltas, const float* layer_input, float* grad_w,
                       float* grad_b, int n, int k, int f, int out_w, int out_h, int S) {
  const int wsize = f * f * k * n;
  for (int s = 0; s < S; s++) {  // serial: see header comment
#pragma omp parallel for schedule(static)
    for (int id = 0; id < wsize; id++) {
      clshim::gid[0] = id;
      clshim::gid[1] = s;
      k_backprop::backpropagate(const_cast<float*>(deltas), const_cast<float*>(layer_input),
                                grad_w, grad_b, (uint)n, (uint)k, (uint)f, (uint)out_w,
                                (uint)out_h);
    }
  }
}

// kernel `update_params` (update_parameters.cl:1-33)
void ref_update_params(float* w, float* b, const float* gw, const float* gb, float* pdw,
                       float* pdb, float momentum, float decay, float lr, unsigned batch,
                       unsigned wsize, unsigned bsize) {
  const unsigned total = wsize > bsize ? wsize : bsize;
  for (unsigned i = 0; i < total; i++) {
    clshim::gid[0] = i;
    k_update::update_params(w, b, const_cast<float*>(gw), const_cast<float*>(gb), pdw, pdb,
                            momentum, decay, lr, batch, wsize, bsize);
  }
}

// kernel `sub_from_all` (subtract_from_all.cl:1-8)
void ref_sub_from_all(float* data, float value, unsigned len) {
  for (unsigned i = 0; i < len; i++) {
    clshim::gid[0] = i;
    k_sub::sub_from_all(data, value, len);
  }
}

// sum / sum-of-squares (sum.cl:35-68), restated as a double sum
double ref_sum(const float* data, unsigned len, int squared) {
  double acc = 0.0;
  for (unsigned i = 0; i < len; i++) {
    float v = data[i];
    if (squared) v = v * v;
    acc += (double)v;
  }
  return acc;
}

// ------------------------------------------------------------------------------
// Sequencing, restated from src/ConfigBasedDataPipeline.cpp
// ------------------------------------------------------------------------------
struct RefNet {
  int n1, n2, f1, f2, f3;
  // parameters / gradient accumulators / momentum state, per layer
  float *w[3], *b[3], *gw[3], *gb[3], *pw[3], *pb[3];
};

static void layer_shape(const RefNet* net, int l, int* k, int* n, int* f) {
  // ConfigBasedDataPipeline ctor (src/ConfigBasedDataPipeline.cpp:24-30)
  if (l == 0) { *k = 1; *n = net->n1; *f = net->f1; }
  if (l == 1) { *k = net->n1; *n = net->n2; *f = net->f2; }
  if (l == 2) { *k = net->n2; *n = 1; *f = net->f3; }
}

// ConfigBasedDataPipeline::forward(w,h,S) (src/ConfigBasedDataPipeline.cpp:200-241)
void ref_net_forward(const RefNet* net, const float* in, int w, int h, int S, float* out1,
                     float* out2, float* out3) {
  int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  ref_forward(in, out1, net->w[0], net->b[0], 1, net->n1, net->f1, 0, w, h, S);
  ref_forward(out1, out2, net->w[1], net->b[1], net->n1, net->n2, net->f2, 0, w1, h1, S);
  ref_forward(out2, out3, net->w[2], net->b[2], net->n2, 1, net->f3, 1, w2, h2, S);
}

// ConfigBasedDataPipeline::backpropagate (src/ConfigBasedDataPipeline.cpp:243-323)
void ref_net_backward(RefNet* net, const float* in, const float* gt, int w, int h, int S,
                      const float* out1, const float* out2, const float* out3, float* d1,
                      float* d2, float* d3) {
  int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  int w3 = w2 - net->f3 + 1, h3 = h2 - net->f3 + 1;
  ref_last_layer_delta(gt, out3, d3, w, h, w3, h3, S);
  ref_deltas(d3, out2, d2, net->w[2], net->n2, net->f2, net->f3, 1, w2, h2, S);
  ref_deltas(d2, out1, d1, net->w[1], net->n1, net->f1, net->f2, net->n2, w1, h1, S);
  ref_backpropagate(d3, out2, net->gw[2], net->gb[2], 1, net->n2, net->f3, w3, h3, S);
  ref_backpropagate(d2, out1, net->gw[1], net->gb[1], net->n2, net->n1, net->f2, w2, h2, S);
  ref_backpropagate(d1, in, net->gw[0], net->gb[0], net->n1, 1, net->f1, w1, h1, S);
}

// ConfigBasedDataPipeline::update_parameters (src/ConfigBasedDataPipeline.cpp:325-361):
// layer 3, 2, 1 with learning_rate[2], [1], [0]; then zero the six accumulators.
void ref_net_update(RefNet* net, unsigned batch_size, float momentum, float decay,
                    const float* lr3) {
  for (int l = 2; l >= 0; l--) {
    int k, n, f;
    layer_shape(net, l, &k, &n, &f);
    unsigned ws = (unsigned)(f * f * k * n), bs = (unsigned)n;
    ref_update_params(net->w[l], net->b[l], net->gw[l], net->gb[l], net->pw[l], net->pb[l],
                      momentum, decay, lr3[l], batch_size, ws, bs);
    memset(net->gw[l], 0, sizeof(float) * ws);
    memset(net->gb[l], 0, sizeof(float) * bs);
  }
}

// One training epoch as driven by main() + execute_batch
// (src/Main_cl.cpp:161-170, src/ConfigBasedDataPipeline.cpp:128-195): the train
// set is processed in chunks of `mini_batch` samples, gradients accumulate over
// all chunks, ONE update with batch_size = |train set|.
// scratch must hold, for `mini_batch` samples: out1,out2,out3,d1,d2,d3.
void ref_net_train_epoch(RefNet* net, const float* inputs, const float* gts, int w, int h,
                         int n_samples, int mini_batch, float momentum, float decay,
                         const float* lr3, int do_update) {
  int w1 = w - net->f1 + 1, h1 = h - net->f1 + 1;
  int w2 = w1 - net->f2 + 1, h2 = h1 - net->f2 + 1;
  int w3 = w2 - net->f3 + 1, h3 = h2 - net->f3 + 1;
  size_t e1 = (size_t)w1 * h1 * net->n1, e2 = (size_t)w2 * h2 * net->n2, e3 = (size_t)w3 * h3;
  std::vector<float> o1(e1 * mini_batch), o2(e2 * mini_batch), o3(e3 * mini_batch);
  std::vector<float> d1(e1 * mini_batch), d2(e2 * mini_batch), d3(e3 * mini_batch);
  for (int i = 0; i < n_samples; i += mini_batch) {
    int S = n_samples - i < mini_batch ? n_samples - i : mini_batch;
    const float* in = inputs + (size_t)i * w * h;
    const float* gt = gts + (size_t)i * w * h;
    ref_net_forward(net, in, w, h, S, o1.data(), o2.data(), o3.data());
    ref_net_backward(net, in, gt, w, h, S, o1.data(), o2.data(), o3.data(), d1.data(),
                     d2.data(), d3.data());
  }
  if (do_update) ref_net_update(net, (unsigned)n_samples, momentum, decay, lr3);
}

}  // extern "C"
