// C++ spec runner for the host layer -- the counterpart of the reference's test/TestRunner.cpp
// + test/specs/*.cpp, written against the SAME DataPipeline signatures (with the sample_count
// argument the reference's stale specs lack).  Needs a GPU; driven by tests/test_host_cpp.py.
//
//   host_specs specs <golden_dir> <reference_test_data_dir>   run all specs; exit code = failures
//   host_specs chain <in.json> <out.json>                     ConfigBasedDataPipeline training
//        chain on float samples (execute_batch + update_parameters per epoch), dumps parameters
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <random>
#include <sstream>

#include "Config.hpp"
#include "ConfigBasedDataPipeline.hpp"
#include "Context.hpp"
#include "DataPipeline.hpp"
#include "LayerData.hpp"
#include "json.hpp"

using namespace cnn_sr;
using opencl::MemoryHandle;

namespace {

struct SpecFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// symmetric tolerance (the reference's assert_equals is one-sided, test/TestCase.cpp:48-63)
void expect_near(float expected, float got, float tol, const char* what, size_t idx = 0) {
  if (!(std::fabs(expected - got) <= tol)) {
    std::ostringstream os;
    os << what << "[" << idx << "]: expected " << expected << ", got " << got;
    throw SpecFailure(os.str());
  }
}

void expect_all_near(const std::vector<float>& exp, const std::vector<float>& got, float tol,
                     const char* what) {
  if (exp.size() != got.size()) throw SpecFailure(std::string(what) + ": size mismatch");
  for (size_t i = 0; i < exp.size(); i++) expect_near(exp[i], got[i], tol, what, i);
}

std::vector<float> floats(const json::Value& v) {
  std::vector<float> out;
  for (const json::Value& e : v.array) out.push_back((float)e.number);
  return out;
}

MemoryHandle upload(opencl::Context* ctx, const std::vector<float>& v) {
  MemoryHandle h = ctx->allocate(CL_MEM_READ_WRITE, sizeof(float) * v.size());
  ctx->write_buffer(h, (void*)v.data(), true);
  return h;
}

std::vector<float> download(opencl::Context* ctx, MemoryHandle h, size_t n) {
  std::vector<float> v(n);
  ctx->block();
  ctx->read_buffer(h, 0, sizeof(float) * n, v.data(), true);
  return v;
}

std::string golden_dir, ref_data_dir;

// ---- LayerTest (reference: test/specs/LayerTest.cpp:97-130) ------------------------------
void layer_test(DataPipeline* p) {
  const json::Value root = json::parse_file((golden_dir + "/layer_cases.json").c_str());
  const json::Value* cases = root.find("cases");
  for (const auto& kv : cases->object) {
    const json::Value& c = kv.second;
    LayerData data((size_t)c.find("n_prev_filter_cnt")->number,
                   (size_t)c.find("current_filter_count")->number,
                   (size_t)c.find("f_spatial_size")->number);
    std::vector<float> w = floats(*c.find("weights")), b = floats(*c.find("bias"));
    data.set_weights(w.data());
    data.set_bias(b.data());
    const size_t iw = (size_t)c.find("input_w")->number, ih = (size_t)c.find("input_h")->number;
    MemoryHandle in = upload(p->context(), floats(*c.find("input")));
    MemoryHandle out = gpu_nullptr;
    LayerAllocationPool pool;
    opencl::Kernel* kernel = p->create_layer_kernel(data, false);
    p->execute_layer(*kernel, data, pool, in, iw, ih, 1, out);
    const std::vector<float> exp = floats(*c.find("output"));
    expect_all_near(exp, download(p->context(), out, exp.size()), 5.1e-4f, kv.first.c_str());
  }
}

// ---- LayerDeltasTest (reference: test/specs/LayerDeltasTest.cpp:141-193) ------------------
void layer_deltas_test(DataPipeline* p) {
  const json::Value c = json::parse_file((golden_dir + "/layer_deltas_case.json").c_str());
  const size_t IGNORED = 10;
  LayerData prev_data(IGNORED, 2, IGNORED);
  LayerData curr_data(2, 3, 3);
  float bias[3] = {0, 0, 0};
  std::vector<float> w = floats(*c.find("weights"));
  curr_data.set_bias(bias);
  curr_data.set_weights(w.data());
  std::vector<float> output = floats(*c.find("input_x"));
  for (float& v : output) v = std::max(v, 0.0f);
  LayerAllocationPool curr_pool;
  MemoryHandle curr_deltas = upload(p->context(), floats(*c.find("deltas")));
  MemoryHandle prev_output = upload(p->context(), output);
  MemoryHandle prev_deltas = p->context()->allocate(CL_MEM_READ_WRITE, sizeof(float) * output.size());
  opencl::Kernel* kernel = p->create_deltas_kernel(prev_data);
  p->calculate_deltas(*kernel, prev_data, curr_data, curr_pool, prev_deltas, curr_deltas, 3, 3, 1,
                      prev_output);
  const std::vector<float> exp = floats(*c.find("expected_output"));
  expect_all_near(exp, download(p->context(), prev_deltas, exp.size()), 2e-6f, "deltas");
}

// ---- BackpropagationTest (reference: test/specs/BackpropagationTest.cpp:135-171) ----------
void backpropagation_test(DataPipeline* p) {
  const json::Value c = json::parse_file((golden_dir + "/backprop_case.json").c_str());
  LayerData data(2, 3, 3);
  float w[54] = {0}, bias[10] = {0};
  data.set_bias(bias);
  data.set_weights(w);
  LayerAllocationPool pool;
  MemoryHandle deltas = upload(p->context(), floats(*c.find("deltas")));
  MemoryHandle input = upload(p->context(), floats(*c.find("input")));
  pool.accumulating_grad_w = p->context()->allocate(CL_MEM_READ_WRITE, sizeof(float) * 54);
  p->context()->fill_float(pool.accumulating_grad_w, (float)c.find("grad_w_init")->number, true);
  p->backpropagate(data, input, deltas, pool, 3, 3, 1);
  expect_all_near(floats(*c.find("expected_weights")),
                  download(p->context(), pool.accumulating_grad_w, 54), 6e-5f, "grad_w");
  expect_all_near(floats(*c.find("expected_bias")),
                  download(p->context(), pool.accumulating_grad_b, 3), 1e-6f, "grad_b");
  // data set 2: big data must not crash (k=32 n=16 f=3 on 1024x1024)
  LayerData big(32, 16, 3);
  std::vector<float> bw(4608, 0.f), bb(16, 0.f);
  big.set_weights(bw.data());
  big.set_bias(bb.data());
  LayerAllocationPool big_pool;
  MemoryHandle bd = p->context()->allocate(CL_MEM_READ_WRITE, sizeof(float) * 1022 * 1022 * 16);
  MemoryHandle bi = p->context()->allocate(CL_MEM_READ_WRITE, sizeof(float) * 1024 * 1024 * 32);
  p->context()->zeros_float(bd, true);
  p->context()->zeros_float(bi, true);
  p->backpropagate(big, bi, bd, big_pool, 1022, 1022, 1);
  p->context()->block();
  p->context()->raw_memory(bd)->release();
  p->context()->raw_memory(bi)->release();
}

// ---- LastLayerDeltaTest / SquaredErrorTest (formula-pinned, poisoned border) --------------
void last_layer_delta_test(DataPipeline* p) {
  const size_t aw = 6, ah = 6, pad = 4, gw = aw + 2 * pad, gh = ah + 2 * pad;
  std::mt19937 gen(1234);
  std::vector<float> gt(gw * gh, 99999.0f), algo(aw * ah), exp(aw * ah);
  for (size_t i = 0; i < aw * ah; i++) {
    const size_t row = i / aw, col = i % aw;
    const float t = (gen() % 256) / 100.0f, x = (gen() % 2560) / 1000.0f - 1.0f;
    const float y = std::max(x, 0.0f);
    exp[i] = (y - t) * (x > 0.0f ? 1.0f : 0.0f);
    gt[(row + pad) * gw + pad + col] = t;
    algo[i] = y;
  }
  MemoryHandle g = upload(p->context(), gt), a = upload(p->context(), algo), out = gpu_nullptr;
  p->last_layer_delta(g, gw, gh, 1, a, out, 2 * pad);
  expect_all_near(exp, download(p->context(), out, exp.size()), 0.f, "last_layer_delta");
}

void squared_error_test(DataPipeline* p) {
  const size_t aw = 1000, ah = 2000, pad = 4, gw = aw + 2 * pad, gh = ah + 2 * pad;
  std::mt19937 gen(4321);
  std::vector<float> gt(gw * gh, 99999.0f), algo(aw * ah);
  double sum = 0.0;
  for (size_t i = 0; i < aw * ah; i++) {
    const size_t row = i / aw, col = i % aw, gi = (row + pad) * gw + pad + col;
    gt[gi] = (float)(gen() % 256);
    algo[i] = (gen() % 2560) / 10.0f;
    const float d = algo[i] - gt[gi];
    sum += (double)(d * d);
  }
  MemoryHandle g = upload(p->context(), gt), a = upload(p->context(), algo);
  float target = 0.f;
  p->squared_error(g, gw, gh, 1, a, gpu_nullptr, target, 2 * pad);
  p->context()->block();
  expect_near((float)sum, target, (float)(sum * 1e-6), "squared_error");
}

// ---- UpdateParametersTest (reference: test/specs/UpdateParametersTest.cpp:55-95) ----------
void update_parameters_test(DataPipeline* p) {
  LayerData layer(2, 400, 5);
  const size_t ws = layer.weight_size(), bs = layer.bias_size(), batch = 2;
  const float momentum = 0.8f, lr = 0.001f;
  std::mt19937 gen(99);
  LayerAllocationPool pool;
  auto make = [&](size_t n, MemoryHandle& cur, MemoryHandle& grad, MemoryHandle& prev,
                  std::vector<float>& exp, std::vector<float>& cur_v, std::vector<float>& delta) {
    std::vector<float> g(n), pd(n);
    cur_v.resize(n); exp.resize(n); delta.resize(n);
    for (size_t i = 0; i < n; i++) {
      cur_v[i] = (gen() % 2560) / 10.0f;
      g[i] = (gen() % 2560) / 100.0f;
      pd[i] = (gen() % 2560) / 10.0f;
      delta[i] = momentum * pd[i] + lr * g[i];
      exp[i] = cur_v[i] - delta[i] / batch;
    }
    cur = upload(p->context(), cur_v);
    grad = upload(p->context(), g);
    prev = upload(p->context(), pd);
  };
  std::vector<float> ew, cw, dw, eb, cb, db;
  make(ws, pool.weights, pool.accumulating_grad_w, pool.previous_batch_delta_w, ew, cw, dw);
  make(bs, pool.bias, pool.accumulating_grad_b, pool.previous_batch_delta_b, eb, cb, db);
  layer.set_weights(cw.data());
  layer.set_bias(cb.data());
  p->update_parameters(layer, pool, batch, momentum, 0.0f, lr);
  expect_all_near(ew, download(p->context(), pool.weights, ws), 1e-4f, "weights");
  expect_all_near(eb, download(p->context(), pool.bias, bs), 1e-4f, "bias");
  expect_all_near(dw, download(p->context(), pool.previous_batch_delta_w, ws), 1e-4f, "delta_w");
  expect_all_near(db, download(p->context(), pool.previous_batch_delta_b, bs), 1e-4f, "delta_b");
}

// ---- SumTest / SubtractFromAllTest ---------------------------------------------------------
void sum_test(DataPipeline* p) {
  std::vector<float> data(900);
  double s = 0, s2 = 0;
  for (size_t i = 0; i < 900; i++) {
    data[i] = (float)i;
    s += i;
    s2 += (double)i * i;
  }
  MemoryHandle h = upload(p->context(), data);
  expect_near((float)s, p->sum(h, false), 0.5f, "sum");
  expect_near((float)s2, p->sum(h, true), 20.f, "sum squared");
}

// ---- ExtractLumaTest / SwapLumaTest (reference image fixtures, decoded by its stb_image) ---
std::vector<unsigned char> read_bytes(const std::string& path, size_t expect) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) throw SpecFailure("cannot open fixture " + path);
  std::vector<unsigned char> v((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (v.size() != expect) throw SpecFailure("fixture " + path + " has the wrong size");
  return v;
}

// test/specs/ExtractLumaTest.cpp:24-28,50-74: 25 goldens on color_grid.png, both data sets
void extract_luma_test(DataPipeline* p) {
  const json::Value g = json::parse_file((golden_dir + "/luma_goldens.json").c_str());
  const std::vector<float> expected = floats(*g.find("extract_luma_normalized_5x5"));
  std::vector<unsigned char> rgba = read_bytes(golden_dir + "/color_grid_5x5.rgba", 100);
  for (int normalize = 1; normalize >= 0; normalize--) {
    opencl::utils::ImageData img(5, 5, 4, rgba.data());
    MemoryHandle raw = gpu_nullptr, luma = gpu_nullptr;
    p->extract_luma(img, raw, luma, normalize != 0);
    std::vector<float> exp = expected;
    if (!normalize) for (float& v : exp) v *= 255.f;
    expect_all_near(exp, download(p->context(), luma, 25), normalize ? 0.005f : 1.275f,
                    normalize ? "extract_luma normalized" : "extract_luma raw");
  }
}

// test/specs/SwapLumaTest.cpp:39-90: exact integer image
void swap_luma_test(DataPipeline* p) {
  std::vector<unsigned char> rgba = read_bytes(golden_dir + "/color_grid2_32x32.rgba", 32 * 32 * 4);
  const std::vector<unsigned char> expected =
      read_bytes(golden_dir + "/color_grid2_luma_swapped_32x32.rgb", 32 * 32 * 3);
  const size_t padding = 10, lw = 32 - 2 * padding, n = lw * lw;
  std::vector<float> new_luma(n);
  for (size_t i = 0; i < n; i++) new_luma[i] = i * 1.0f / n;
  opencl::utils::ImageData img(32, 32, 4, rgba.data());
  MemoryHandle raw = gpu_nullptr, target = gpu_nullptr;
  MemoryHandle luma = upload(p->context(), new_luma);
  p->swap_luma(img, raw, luma, target, lw, lw);
  std::vector<unsigned char> got(32 * 32 * 3);
  p->context()->read_buffer(target, got.data(), true);
  for (size_t i = 0; i < got.size(); i++)
    if (got[i] != expected[i])
      throw SpecFailure("swap_luma: [INT] Expected " + std::to_string((int)got[i]) + " to be " +
                        std::to_string((int)expected[i]) + " at byte " + std::to_string(i));
}

void subtract_from_all_test(DataPipeline* p) {
  std::vector<float> data(900), exp(900);
  for (size_t i = 0; i < 900; i++) {
    data[i] = (float)i;
    exp[i] = (float)i - 450.0f;
  }
  MemoryHandle h = upload(p->context(), data);
  p->subtract_from_all(h, 450.0f);
  expect_all_near(exp, download(p->context(), h, 900), 0.f, "sub_from_all");
  // quirk Q1: subtract_mean with an event pointer subtracts the mean of squares
  MemoryHandle h2 = upload(p->context(), data);
  cl_event ev = p->context()->ticket();
  float mean = 0.f;
  p->subtract_mean(h2, &mean, &ev);
  double s2 = 0;
  for (size_t i = 0; i < 900; i++) s2 += (double)i * i;
  expect_near((float)(s2 / 900.0), mean, 1.0f, "subtract_mean(Q1)");
  MemoryHandle h3 = upload(p->context(), data);
  p->subtract_mean(h3, &mean, nullptr);
  expect_near(449.5f, mean, 1e-3f, "subtract_mean(true mean)");
}

// ---- ConfigTest (reference: test/specs/ConfigTest.cpp:68-116) ------------------------------
void config_test(DataPipeline*) {
  if (ref_data_dir.empty()) return;
  ConfigReader reader;
  Config cfg = reader.read((ref_data_dir + "/config.json").c_str());
  if (cfg.n1 != 32 || cfg.n2 != 16 || cfg.f1 != 9 || cfg.f2 != 1 || cfg.f3 != 5)
    throw SpecFailure("config: wrong layer shape");
  expect_near(123.5f, cfg.momentum, 0.f, "momentum");
  expect_near(0.1f, cfg.weight_decay_parameter, 0.f, "weight_decay");
  expect_near(12.f, cfg.learning_rate[0], 0.f, "lr0");
  expect_near(56.f, cfg.learning_rate[2], 0.f, "lr2");
  expect_near(2.001f, cfg.params_distr_2.sd_w, 0.f, "pd2.sd_w");
  if (cfg.parameters_file != "cnn-parameters-a.json") throw SpecFailure("config: parameters_file");
  // "invalid value" data set of the reference: parses, but distribution 3 differs
  Config other = reader.read((ref_data_dir + "/config_invalid_val.json").c_str());
  expect_near(9999.f, other.params_distr_3.mean_w, 0.f, "invalid_val.pd3.mean_w");
  expect_near(0.001f, cfg.params_distr_3.mean_w, 0.f, "pd3.mean_w");
  bool threw = false;
  try { reader.read((ref_data_dir + "/config_non_parseable.json").c_str()); }
  catch (const IOException&) { threw = true; }
  if (!threw) throw SpecFailure("config: unparsable file must throw IOException");
  threw = false;
  try { reader.read((ref_data_dir + "/does_not_exist.json").c_str()); }
  catch (const IOException&) { threw = true; }
  if (!threw) throw SpecFailure("config: missing file must throw IOException");
}

void error_behaviour_test(DataPipeline* p) {
  // LayerData::validate must reject short weight vectors (src/LayerData.cpp:21-45)
  LayerData data(1, 4, 3);
  bool threw = false;
  try { LayerData::validate(data); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) throw SpecFailure("LayerData::validate accepted empty weights");
  // an existing but too small output allocation is an error, not a silent reallocation
  std::vector<float> w(36, 0.f), b(4, 0.f);
  data.set_weights(w.data());
  data.set_bias(b.data());
  MemoryHandle in = upload(p->context(), std::vector<float>(25, 0.f));
  MemoryHandle small = p->context()->allocate(CL_MEM_READ_WRITE, 8);
  LayerAllocationPool pool;
  opencl::Kernel* k = p->create_layer_kernel(data, false);
  threw = false;
  try { p->execute_layer(*k, data, pool, in, 5, 5, 1, small); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) throw SpecFailure("execute_layer accepted a too small output buffer");
}

// ---- host-only specs (no GPU needed): config, LayerData, Argparse, JSON ---------------------
void host_only_test() {
  config_test(nullptr);
  LayerData d(3, 2, 3);
  if (d.weight_size() != 54 || d.bias_size() != 2 || d.input_size(4, 5) != 60)
    throw SpecFailure("LayerData size math");
  size_t dims[2];
  d.get_output_dimensions(dims, 10, 7);
  if (dims[0] != 8 || dims[1] != 5) throw SpecFailure("LayerData output dims");
  bool threw = false;
  try { LayerData::validate(d); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) throw SpecFailure("LayerData::validate accepted empty weights");
  // CLI grammar (reference: src/pch.cpp:183-299)
  utils::Argparse ap("cnn", "help");
  ap.add_argument("train");
  ap.add_argument("dry");
  ap.add_argument("-c", "--config").required();
  ap.add_argument("-i", "--in").required();
  ap.add_argument("-e", "--epochs");
  const char* argv1[] = {"cnn", "train", "--config", "a.json", "-i", "dir", "bogus", "-e", "17"};
  if (!ap.parse(9, const_cast<char**>(argv1))) throw SpecFailure("argparse: parse failed");
  size_t epochs = 0;
  ap.value("epochs", epochs);
  if (!ap.has_arg("train") || ap.has_arg("dry") || epochs != 17 ||
      std::string(ap.value("config")) != "a.json" || std::string(ap.value("in")) != "dir")
    throw SpecFailure("argparse: wrong values");
  const char* argv2[] = {"cnn", "-i", "x"};
  threw = false;
  try { ap.parse(3, const_cast<char**>(argv2)); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) throw SpecFailure("argparse: missing required --config must throw");
  // JSON
  json::Value v = json::parse("{\"a\": [1, 2.5e0, -3], \"b\": {\"c\": \"x\\ny\"}, \"d\": true}");
  if (v.find("a")->array.size() != 3 || v.find("a")->array[1].number != 2.5 ||
      v.find("b")->find("c")->string != "x\ny" || !v.find("d")->boolean)
    throw SpecFailure("json parse");
  threw = false;
  try { json::parse("{\"a\": [1, 2"); } catch (const IOException&) { threw = true; }
  if (!threw) throw SpecFailure("json: truncated input must throw");
  if (utils::closest_power_of_2(5) != 8 || utils::closest_power_of_2(8) != 8 ||
      utils::closest_power_of_2(0) != 0)
    throw SpecFailure("closest_power_of_2");
}

int run_specs() {
  opencl::Context context;
  context.init();
  DataPipeline pipeline(&context);
  pipeline.init(DataPipeline::LOAD_KERNEL_MISC | DataPipeline::LOAD_KERNEL_BACKPROPAGATE |
                DataPipeline::LOAD_KERNEL_LUMA);
  struct Spec { const char* name; std::function<void(DataPipeline*)> fn; };
  const Spec specs[] = {
      {"Layer test", layer_test}, {"Layer deltas test", layer_deltas_test},
      {"Backpropagation test", backpropagation_test}, {"Last layer delta test", last_layer_delta_test},
      {"Squared error test", squared_error_test}, {"Update parameters test", update_parameters_test},
      {"Sum all test", sum_test}, {"Subtract from all test", subtract_from_all_test},
      {"Extract luma test", extract_luma_test}, {"Swap luma test", swap_luma_test},
      {"Config test", config_test}, {"Error behaviour test", error_behaviour_test},
  };
  int failures = 0;
  for (const Spec& s : specs) {
    bool ok = true;
    std::string msg;
    try { s.fn(&pipeline); } catch (const std::exception& e) { ok = false; msg = e.what(); }
    std::cout << (ok ? "  [+] " : "  [-] ") << s.name << (ok ? "" : " : " + msg) << std::endl;
    failures += ok ? 0 : 1;
  }
  std::cout << (failures ? "FAILED " : "PASSED ") << failures << " failure(s)" << std::endl;
  return failures;
}

// ---- training chain through ConfigBasedDataPipeline ----------------------------------------
int run_chain(const char* in_path, const char* out_path) {
  const json::Value in = json::parse_file(in_path);
  ConfigReader reader;
  Config cfg = reader.read(in.find("config_path")->string.c_str());
  const size_t w = (size_t)in.find("w")->number, h = (size_t)in.find("h")->number;
  const size_t epochs = (size_t)in.find("epochs")->number;
  const size_t chunk = (size_t)in.find("chunk")->number;
  opencl::Context context;
  context.init();
  ConfigBasedDataPipeline pipeline(cfg, &context);
  pipeline.init(DataPipeline::LOAD_KERNEL_ALL);
  GpuAllocationPool pool;
  const std::vector<float> xs = floats(*in.find("x")), gts = floats(*in.find("gt"));
  const size_t n = xs.size() / (w * h);
  for (size_t i = 0; i < n; i++) {
    SampleAllocationPool s;
    s.input_w = w;
    s.input_h = h;
    s.input_luma = upload(&context, std::vector<float>(xs.begin() + i * w * h, xs.begin() + (i + 1) * w * h));
    s.expected_luma = upload(&context, std::vector<float>(gts.begin() + i * w * h, gts.begin() + (i + 1) * w * h));
    pool.samples.push_back(s);
  }
  std::vector<SampleAllocationPool*> set;
  for (auto& s : pool.samples) set.push_back(&s);
  pipeline.set_mini_batch_size(chunk);
  for (size_t e = 0; e < epochs; e++) {
    pipeline.execute_batch(true, pool, set);
    pipeline.update_parameters(pool.layer_1, pool.layer_2, pool.layer_3, set.size());
  }
  const float sse = pipeline.execute_batch(false, pool, set);
  pipeline.write_params_to_file(out_path, pool.layer_1, pool.layer_2, pool.layer_3);
  std::cout << "validation_sse " << sse << std::endl;
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && std::strcmp(argv[1], "specs") == 0) {
      golden_dir = argv[2];
      if (argc >= 4) ref_data_dir = argv[3];
      return run_specs();
    }
    if (argc >= 4 && std::strcmp(argv[1], "chain") == 0) return run_chain(argv[2], argv[3]);
    if (argc >= 3 && std::strcmp(argv[1], "hostonly") == 0) {
      ref_data_dir = argv[2];
      host_only_test();
      std::cout << "PASSED host-only specs" << std::endl;
      return 0;
    }
  } catch (const std::exception& e) {
    std::cout << "[ERROR] " << e.what() << std::endl;
    return 100;
  }
  std::cout << "usage: host_specs specs <golden_dir> [ref_test_data_dir] | chain <in.json> <out.json>" << std::endl;
  return 2;
}
