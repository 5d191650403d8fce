// `cnn [train] [dry] [profile] --config CFG --in PATH [--out PATH] [--epochs N]`
// Same grammar, flow and console output as the reference CLI (src/Main_cl.cpp:40-318):
// forward mode = one image through the three layers, luma swapped back into the image;
// training mode = samples `*_large.*` (ground truth) / `*_small.*` (input) from a directory,
// 80/20 train/validation split re-drawn every epoch, ONE parameter update per epoch.
// Images are binary PPM/PGM (see Context.hpp); `_large.ppm/_small.ppm` as well as the
// reference's `.jpg` suffix pattern are recognised by name.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <random>
#include <stdexcept>
#include <unordered_map>
#include <utility>

#include "Config.hpp"
#include "ConfigBasedDataPipeline.hpp"
#include "Context.hpp"
#include "LayerData.hpp"
#include "pch.hpp"

using namespace opencl::utils;
using namespace cnn_sr;

typedef std::pair<std::string, std::string> TrainSampleFiles;  // (large, small)

static cl_event prepare_image(DataPipeline* pipeline, const char* file_path, ImageData& img,
                              opencl::MemoryHandle& gpu_data, opencl::MemoryHandle& gpu_luma) {
  load_image(file_path, img);
  return pipeline->extract_luma(img, gpu_data, gpu_luma, /*normalize=*/true);
}

static void execute_forward(ConfigBasedDataPipeline& pipeline, GpuAllocationPool& gpu_alloc,
                            const char* in_path, const char* out_path) {
  opencl::Context* context = pipeline.context();
  ImageData input_img;
  SampleAllocationPool sample;
  cl_event ev = prepare_image(&pipeline, in_path, input_img, sample.input_data, sample.input_luma);
  pipeline.subtract_mean(sample.input_luma, nullptr, &ev);  // (quirk Q1 applies)
  sample.input_w = (size_t)input_img.w;
  sample.input_h = (size_t)input_img.h;
  context->block();
  pipeline.forward(gpu_alloc.layer_1, gpu_alloc.layer_2, gpu_alloc.layer_3, sample);
  if (out_path) pipeline.write_result_image(out_path, input_img, sample);
}

static bool ends_with(const std::string& s, const std::string& suffix, size_t* stem_len) {
  if (s.size() < suffix.size() || s.compare(s.size() - suffix.size(), suffix.size(), suffix) != 0)
    return false;
  *stem_len = s.size() - suffix.size();
  return true;
}

static void get_training_samples(const std::string& dir, std::vector<TrainSampleFiles>& target) {
  std::vector<std::string> files;
  cnn_sr::utils::list_files(dir.c_str(), files);
  std::sort(files.begin(), files.end());  // deterministic order (readdir's is not)
  std::unordered_map<std::string, TrainSampleFiles> by_base;
  std::vector<std::string> order;
  for (const std::string& f : files) {
    size_t stem = 0;
    bool large = false, small = false;
    for (const char* ext : {".ppm", ".pgm", ".jpg"}) {
      if (ends_with(f, std::string("_large") + ext, &stem)) large = true;
      else if (ends_with(f, std::string("_small") + ext, &stem)) small = true;
      if (large || small) break;
    }
    if (!large && !small) {
      if (f != "." && f != "..")
        std::cout << "'" << f << "' is not a sample image. Skipping sample" << std::endl;
      continue;
    }
    const std::string base = f.substr(0, stem);
    if (!by_base.count(base)) order.push_back(base);
    (large ? by_base[base].first : by_base[base].second) = dir + "/" + f;
  }
  for (const std::string& base : order) {
    const TrainSampleFiles& p = by_base[base];
    if (p.first.empty() || p.second.empty())
      std::cout << "Only 1 image for pair with name '" << base << "'. Skipping sample" << std::endl;
    else
      target.push_back(p);
  }
}

static void divide_samples(size_t validation_set_size, GpuAllocationPool& pool,
                           std::vector<SampleAllocationPool*>& train_set,
                           std::vector<SampleAllocationPool*>& validation_set, std::mt19937& rng) {
  train_set.clear();
  validation_set.clear();
  std::shuffle(pool.samples.begin(), pool.samples.end(), rng);
  for (size_t i = 0; i < pool.samples.size(); i++)
    (i < validation_set_size ? validation_set : train_set).push_back(&pool.samples[i]);
}

// contiguous even split of a sample set across the ranks
static std::vector<SampleAllocationPool*> shard(const std::vector<SampleAllocationPool*>& set,
                                                int rank, int world) {
  const size_t base = set.size() / (size_t)world, rem = set.size() % (size_t)world;
  const size_t begin = (size_t)rank * base + std::min<size_t>((size_t)rank, rem);
  const size_t count = base + ((size_t)rank < rem ? 1 : 0);
  return std::vector<SampleAllocationPool*>(set.begin() + (long)begin, set.begin() + (long)(begin + count));
}

int main(int argc, char** argv) {
  cnn_sr::utils::Argparse argparse("cnn", "SRCNN super-resolution (B200 CUDA build)");
  argparse.add_argument("train").help("Train mode");
  argparse.add_argument("dry").help("Do not store result");
  argparse.add_argument("profile").help("Print kernel execution times");
  argparse.add_argument("-c", "--config").required().help("CNN configuration");
  argparse.add_argument("-i", "--in").required().help("Image during forward, samples directory during training");
  argparse.add_argument("-o", "--out").help("Output file path (either result image or new parameters)");
  argparse.add_argument("-e", "--epochs").help("Number of epochs during training");
  try {
    if (!argparse.parse((size_t)argc, argv)) return EXIT_SUCCESS;
  } catch (const std::exception& e) {
    std::cout << e.what() << std::endl;
    return EXIT_FAILURE;
  }

  const bool train = argparse.has_arg("train"), dry = argparse.has_arg("dry"),
             profile = argparse.has_arg("profile");
  const char* config_path = argparse.value("config");
  const char* in_path = argparse.value("in");
  const char* out_path = dry ? nullptr : argparse.value("out");
  size_t epochs = 0;
  argparse.value("epochs", epochs);
  if (!dry && !out_path) {
    std::cout << "Either provide out path or do the dry run" << std::endl;
    return EXIT_FAILURE;
  }
  if (profile)
    std::cout << "!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!" << std::endl
              << "!!! RUNNING IN PROFILING MODE !!!" << std::endl
              << "!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!" << std::endl;
  if (train)
    std::cout << "Training mode, epochs: " << epochs << std::endl
              << "Training samples directory: " << in_path << std::endl
              << "Output: " << (out_path ? out_path : "-") << std::endl;
  else
    std::cout << "Forward mode" << std::endl
              << "Input image: " << in_path << std::endl
              << "Output: " << (out_path ? out_path : "-") << std::endl;

  const size_t validation_set_percent = 20, mini_batch_count = 2;
  bool error = false;
  try {
    ConfigReader reader;
    Config cfg = reader.read(config_path);
    std::cout << cfg << std::endl;

    opencl::Context context;
    context.init(profile);
    ConfigBasedDataPipeline data_pipeline(cfg, &context);
    data_pipeline.init(DataPipeline::LOAD_KERNEL_ALL);
    GpuAllocationPool gpu_alloc;

    if (!train) {
      execute_forward(data_pipeline, gpu_alloc, in_path, out_path);
      context.block();
      return EXIT_SUCCESS;
    }

    std::vector<TrainSampleFiles> files;
    get_training_samples(in_path, files);
    if (files.empty()) throw std::runtime_error("No training samples found");
    const size_t validation_set_size = (size_t)(files.size() * validation_set_percent / 100.0f),
                 train_set_size = files.size() - validation_set_size;
    if (validation_set_size == 0)
      std::cout << "[WARNING] Validation set is empty" << std::endl;
    else
      std::cout << "validation_set_size: " << validation_set_size << "/" << files.size() << " = "
                << (validation_set_size * 100.0f / files.size()) << "%" << std::endl;
    data_pipeline.set_mini_batch_size((train_set_size / mini_batch_count) + mini_batch_count);

    for (auto& pair : files) {
      ImageData expected_img, input_img;
      SampleAllocationPool s;
      prepare_image(&data_pipeline, pair.first.c_str(), expected_img, s.expected_data, s.expected_luma);
      cl_event ev = prepare_image(&data_pipeline, pair.second.c_str(), input_img, s.input_data, s.input_luma);
      data_pipeline.subtract_mean(s.input_luma, nullptr, &ev);
      s.input_w = (size_t)input_img.w;
      s.input_h = (size_t)input_img.h;
      context.block();
      context.raw_memory(s.input_data)->release();     // RGBA copies are no longer needed
      context.raw_memory(s.expected_data)->release();
      gpu_alloc.samples.push_back(s);
    }
    const size_t per_sample_px = gpu_alloc.samples[0].input_w * gpu_alloc.samples[0].input_h;
    context.block();

    // the reference shuffles with an unseeded std::random_shuffle; CNN_SR_SEED pins it
    std::mt19937 rng(std::getenv("CNN_SR_SEED") ? (unsigned)std::atoi(std::getenv("CNN_SR_SEED")) : 5489u);
    std::vector<SampleAllocationPool*> train_set, validation_set;
    const int rank = context.rank(), world = context.world();
    if (world > 1 && !std::getenv("CNN_SR_SEED"))
      std::cout << "[WARNING] data parallel run without CNN_SR_SEED: the ranks share the default "
                   "shuffle seed" << std::endl;
    for (size_t epoch_id = 0; epoch_id < epochs; epoch_id++) {
      divide_samples(validation_set_size, gpu_alloc, train_set, validation_set, rng);
      // data parallel (CNN_SR_WORLD > 1): every rank draws the SAME split (same seed), trains its
      // contiguous share of it, and update_parameters sums the gradients over the ranks before
      // the one update with the GLOBAL batch size
      const size_t global_train = train_set.size(), global_val = validation_set.size();
      if (world > 1) {
        train_set = shard(train_set, rank, world);
        validation_set = shard(validation_set, rank, world);
      }
      if (!train_set.empty()) data_pipeline.execute_batch(true, gpu_alloc, train_set);
      data_pipeline.update_parameters(gpu_alloc.layer_1, gpu_alloc.layer_2, gpu_alloc.layer_3,
                                      global_train);
      if (((epoch_id % 25) == 0 || epoch_id == epochs - 1) && global_val != 0) {
        float sq_err = validation_set.empty()
                           ? 0.0f
                           : data_pipeline.execute_batch(false, gpu_alloc, validation_set);
        sq_err = context.allreduce_scalar(sq_err);   // 1-float all-reduce (no-op on one rank)
        if (std::isnan(sq_err)) {
          std::cout << "Error: squared error is NAN, after " << epoch_id << "/" << epochs
                    << " epochs" << std::endl;
          error = true;
          break;
        }
        const float mean_err = sq_err / global_val;
        if (rank == 0)
          std::cout << "[" << epoch_id << "] mean validation error: " << mean_err << " ("
                    << (mean_err / per_sample_px) << " per px)" << std::endl;
      }
      context.block();
    }
    if (out_path && rank == 0)   // replicas are identical: rank 0 writes
      data_pipeline.write_params_to_file(out_path, gpu_alloc.layer_1, gpu_alloc.layer_2, gpu_alloc.layer_3);
    context.block();
    std::cout << "DONE" << std::endl;
  } catch (const std::exception& e) {
    std::cout << "[ERROR] " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return error ? EXIT_FAILURE : EXIT_SUCCESS;
}
