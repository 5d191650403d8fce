// Stand-alone probe: tcgen05.mma kind::tf32 with MN-MAJOR operands (the weight-gradient
// contraction runs over pixels, which are the slow index of every activation tensor).
// Operands are stored as "planes": element (m, k) at  (m/4)*GS + k*16 + (m%4)*4  bytes, i.e.
// 4 consecutive M elements per 16 bytes, consecutive K (pixels) 16 bytes apart.
// D[128][64] = sum_k A[m][k] * B[n][k], K = 64, checked against FP64 for both assignments of
// the two descriptor strides.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;
constexpr int M = 128, N = 64, K = 64;

__host__ __device__ inline uint32_t idesc_tf32_mn(int Mm, int Nn) {
  // c F32, a/b TF32, a_major = b_major = 1 (MN-major)
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Nn >> 3) << 17) |
         ((uint32_t)(Mm >> 4) << 24);
}
__global__ void __launch_bounds__(128) mn_kernel(const float* A, const float* B, float* D, int variant) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);            // [M/4][K][4]
  float* sB = sA + M * K;                                    // [N/4][K][4]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) { const int m = i / K, k = i % K; sA[(m / 4) * K * 4 + k * 4 + (m & 3)] = A[i]; }
  for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i % K; sB[(n / 4) * K * 4 + k * 4 + (n & 3)] = B[i]; }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32_mn(M, N);
    const uint32_t gs = K * 16;   // bytes between groups of 4 M (or N) elements
    for (int ks = 0; ks < K / 8; ks++) {
      // variant 0: LBO = 128 (next 8 K), SBO = group stride; variant 1: swapped
      const uint32_t lbo = variant == 0 ? 128 : gs, sbo = variant == 0 ? gs : 128;
      const uint64_t ad = make_desc_kmajor(sA, ks * 128, lbo, sbo);
      const uint64_t bd = make_desc_kmajor(sB, ks * 128, lbo, sbo);
      mma_tf32(tmem, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 8; j++) D[row * N + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}
int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(3);
  for (auto& v : A) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : B) v = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(M * K + N * K) * 4;
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rc = 1;
  for (int variant = 0; variant < 2; variant++) {
    cudaMemset(dD, 0, D.size() * 4);
    mn_kernel<<<1, 128, smem>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        double r = 0;
        for (int k = 0; k < K; k++) r += (double)A[m * K + k] * B[n * K + k];
        err = fmax(err, fabs(D[m * N + n] - r));
      }
    double mag = 0, dmag = 0; int nz = 0;
    for (int i = 0; i < M * N; i++) { dmag = fmax(dmag, fabs(D[i])); nz += D[i] != 0.f; }
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double r = 0; for (int k = 0; k < K; k++) r += (double)A[m*K+k]*B[n*K+k]; mag = fmax(mag, fabs(r)); }
    // hypotheses: (a) D = transposed roles, (b) per-element partial matches
    double e_first8 = 0;   // D vs sum over k<8 only (did only the first K group get read?)
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double r = 0; for (int k = 0; k < 8; k++) r += (double)A[m*K+k]*B[n*K+k]; e_first8 = fmax(e_first8, fabs(D[m*N+n]-r)); }
    printf("  max|ref| %.3f max|D| %.3f nonzero %d  err-vs-first-8-K %.3e  D[0][0..3] = %.4f %.4f %.4f %.4f\n", mag, dmag, nz, e_first8, D[0], D[1], D[2], D[3]);
    printf("MN-major tf32, variant %d (%s): max err %.3e %s\n", variant,
           variant == 0 ? "LBO=128 (K groups), SBO=group stride" : "LBO=group stride, SBO=128", err,
           err < 5e-3 ? "PASS" : "fail");
    if (err < 5e-3) rc = 0;
  }
  return rc;
}
