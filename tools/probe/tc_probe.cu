// Stand-alone probe (not part of the library): validates the tcgen05 building blocks the
// tensor-core forward path uses, on the exact configuration it uses:
//   D[128][N] (TMEM, fp32) = A[128][K] * B[N][K]^T, kind::tf32, operands in shared memory in the
//   no-swizzle K-major canonical layout (8-row x 16-byte core matrices, LBO = 128 B between
//   K-chunks, SBO = 128*K/4 B between 8-row groups), error-compensated 3xTF32 split
//   (hi*hi + hi*lo + lo*hi), commit -> mbarrier, tcgen05.ld 32x32b back to registers.
// Prints max errors of plain TF32 and 3xTF32 against an fp64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu && ./tc_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"

using namespace srcnn::tc;

constexpr int M = 128, N = 64, K = 88;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A,
                                                    const float* __restrict__ B, float* D1,
                                                    float* D3) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* sAh = reinterpret_cast<float*>(smem);
  float* sAl = sAh + M * K;
  float* sBh = sAl + M * K;
  float* sBl = sBh + N * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid / 32;

  // operands -> canonical layout, split into tf32 hi / lo
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    float hi, lo;
    split_tf32(A[i], hi, lo);
    sAh[kmajor_offset(m, k, K)] = hi;
    sAl[kmajor_offset(m, k, K)] = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    float hi, lo;
    split_tf32(B[i], hi, lo);
    sBh[kmajor_offset(n, k, K)] = hi;
    sBl[kmajor_offset(n, k, K)] = lo;
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, 128);   // 2 accumulators of 64 columns
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N);
    const uint32_t sbo = 128 * (K / 4);
    for (int ks = 0; ks < K / 8; ks++) {
      const uint64_t ah = make_desc_kmajor(sAh, ks * 256, 128, sbo);
      const uint64_t al = make_desc_kmajor(sAl, ks * 256, 128, sbo);
      const uint64_t bh = make_desc_kmajor(sBh, ks * 256, 128, sbo);
      const uint64_t bl = make_desc_kmajor(sBl, ks * 256, 128, sbo);
      // accumulator 0: plain TF32; accumulator 1 (columns 64..127): 3xTF32
      mma_tf32(tmem + 0, ah, bh, idesc, ks > 0);
      mma_tf32(tmem + 64, al, bh, idesc, ks > 0);
      mma_tf32(tmem + 64, ah, bl, idesc, 1);
      mma_tf32(tmem + 64, ah, bh, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();

  // each warp reads its 32 TMEM lanes (rows), 8 columns at a time
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 8; j++) D1[row * N + c + j] = v[j];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c, v);
    for (int j = 0; j < 8; j++) D3[row * N + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// Same contraction with the A operand in TENSOR MEMORY (what layer 2 uses: A2 = relu(D1+b1) is
// written back to TMEM with tcgen05.st instead of going through shared memory).
constexpr int K2 = 64, N2 = 32;
__global__ void __launch_bounds__(128) probe_ts_kernel(const float* __restrict__ A,
                                                       const float* __restrict__ B, float* D3) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* sBh = reinterpret_cast<float*>(smem);
  float* sBl = sBh + N2 * K2;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < N2 * K2; i += 128) {
    const int n = i / K2, k = i % K2;
    float hi, lo;
    split_tf32(B[i], hi, lo);
    sBh[kmajor_offset(n, k, K2)] = hi;
    sBl[kmajor_offset(n, k, K2)] = lo;
  }
  if (warp == 0) tmem_alloc(&tmem_base_slot, 256);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;
  // A hi -> columns 64..127, A lo -> columns 128..191 ; D -> columns 0..31
  const int row = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < K2; c += 8) {
    float hi[8], lo[8];
    for (int j = 0; j < 8; j++) split_tf32(A[row * K2 + c + j], hi[j], lo[j]);
    tmem_st8(tmem + lane_base + 64 + c, hi);
    tmem_st8(tmem + lane_base + 128 + c, lo);
  }
  tmem_st_wait();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, N2);
    const uint32_t sbo = 128 * (K2 / 4);
    for (int ks = 0; ks < K2 / 8; ks++) {
      const uint64_t bh = make_desc_kmajor(sBh, ks * 256, 128, sbo);
      const uint64_t bl = make_desc_kmajor(sBl, ks * 256, 128, sbo);
      mma_tf32_ts(tmem, tmem + 128 + ks * 8, bh, idesc, ks > 0);
      mma_tf32_ts(tmem, tmem + 64 + ks * 8, bl, idesc, 1);
      mma_tf32_ts(tmem, tmem + 64 + ks * 8, bh, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c = 0; c < N2; c += 8) {
    float v[8];
    tmem_ld8(tmem + lane_base + c, v);
    for (int j = 0; j < 8; j++) D3[row * N2 + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static int run_ts() {
  std::vector<float> A(M * K2), B(N2 * K2);
  srand(11);
  for (auto& v : A) v = fmaxf((float)rand() / RAND_MAX - 0.3f, 0.f);
  for (auto& v : B) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, M * N2 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe_ts_kernel<<<1, 128, 2 * N2 * K2 * 4>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("TS PROBE FAIL: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> D(M * N2);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double e3 = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N2; n++) {
      double r = 0;
      for (int k = 0; k < K2; k++) r += (double)A[m * K2 + k] * B[n * K2 + k];
      e3 = fmax(e3, fabs(D[m * N2 + n] - r));
    }
  printf("TS (A in TMEM) 3xtf32 max err %.3e\n", e3);
  printf(e3 < 2e-5 ? "TS PROBE PASS\n" : "TS PROBE FAIL\n");
  return e3 < 2e-5 ? 0 : 1;
}

int main() {
  const int ts_rc = run_ts();
  std::vector<float> A(M * K), B(N * K);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX - 0.3f;
  for (auto& v : B) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
  float *dA, *dB, *dD1, *dD3;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD1, M * N * 4);
  cudaMalloc(&dD3, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD1, dD3);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PROBE FAIL: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> D1(M * N), D3(M * N);
  cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e3 = 0, ef = 0, mag = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      double r = 0;
      float rf = 0;
      for (int k = 0; k < K; k++) {
        r += (double)A[m * K + k] * B[n * K + k];
        rf = fmaf(A[m * K + k], B[n * K + k], rf);
      }
      e1 = fmax(e1, fabs(D1[m * N + n] - r));
      e3 = fmax(e3, fabs(D3[m * N + n] - r));
      ef = fmax(ef, fabs(rf - r));
      mag = fmax(mag, fabs(r));
    }
  printf("max|ref| %.4f  max err: tf32 %.3e  3xtf32 %.3e  fp32-fma %.3e\n", mag, e1, e3, ef);
  const bool ok = e1 < 5e-3 && e3 < 2e-5;
  printf(ok ? "PROBE PASS\n" : "PROBE FAIL\n");
  return (ok ? 0 : 1) | ts_rc;
}
