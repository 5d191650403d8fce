// Stand-alone probe: tcgen05.mma kind::f16 with MN-MAJOR (transposed) no-swizzle operands in the
// "plane" layout   element (m, k) at  (m/8)*GS + k*16 + (m%8)*2  bytes
// (8 consecutive M/N elements per 16 bytes, consecutive K = pixels 16 bytes apart), which is
// what a pixel-major activation tensor [pixel][channel] becomes when its channel groups of 8 are
// split into planes.  Questions answered (all checked against FP64):
//   1. which of the two descriptor strides is the group stride for MN-major INTERLEAVE
//   2. can an MN-major operand be combined with a K-major one
//   3. does a 16-byte shifted base address shift the K (pixel) index by one
//   4. where do the rows of an M = 64 accumulator live in tensor memory
//   5. may the M groups of an MN-major operand OVERLAP (SBO = 16: group g = the same one-channel
//      "oct" plane shifted by g entries -- the Hankel structure of a one-channel im2col matrix)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mn16_probe mn16_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

constexpr int KT = 64;        // K of the test GEMM (4 MMAs of K = 16)
constexpr int KP = KT + 16;   // plane length in K rows (room for shifted bases)

__host__ __device__ inline uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ inline void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major canonical offset in halves (core matrix 8 rows x 8 halves)
__host__ __device__ inline int kmaj(int r, int k, int K) {
  return (r >> 3) * (64 * (K >> 3)) + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
}

// mode bits: a_mn, b_mn, variant (0: LBO = 128 (K groups), SBO = group stride; 1: swapped), shift
__global__ void __launch_bounds__(128) probe_kernel(const __half* A, const __half* B, float* D, int M,
                                                    int N, int a_mn, int b_mn, int variant,
                                                    int shift) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __half* sA = reinterpret_cast<__half*>(smem_raw);
  __half* sB = sA + 128 * KP;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < 128 * KP; i += 128) { sA[i] = __float2half(0.f); sB[i] = __float2half(0.f); }
  __syncthreads();
  // A is given as [M][KT] row-major, element (m, k) stored at K row (k + shift) of the planes
  if (variant == 2) {
    // overlapping groups: ONE plane P[kk][e] = A[e][kk] (rows 0..7 of A are the plane); logical
    // A'[m = 8 g + e][k] = P[k + g][e]
    for (int i = tid; i < 8 * KP; i += 128) {
      const int kk = i / 8, e = i % 8;
      sA[kk * 8 + e] = kk < KT ? A[e * KT + kk] : __float2half(0.f);
    }
  } else
  for (int i = tid; i < M * KT; i += 128) {
    const int m = i / KT, k = i % KT;
    if (a_mn) sA[(m / 8) * (KP * 8) + (k + shift) * 8 + (m & 7)] = A[i];
    else sA[kmaj(m, k, KT)] = A[i];
  }
  for (int i = tid; i < N * KT; i += 128) {
    const int n = i / KT, k = i % KT;
    if (b_mn) sB[(n / 8) * (KP * 8) + k * 8 + (n & 7)] = B[i];
    else sB[kmaj(n, k, KT)] = B[i];
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(M, N, a_mn, b_mn);
    const uint32_t gs = KP * 16;   // bytes between groups of 8 M (or N) elements
    for (int ks = 0; ks < KT / 16; ks++) {
      const uint32_t lbo = variant == 1 ? gs : 128, sbo = variant == 0 ? gs : (variant == 1 ? 128 : 16);
      // MN-major: one K-step = 16 K rows = 256 bytes along the plane; K-major: 2 core matrices
      const uint64_t ad = a_mn ? make_desc_kmajor(sA, (ks * 16 + shift) * 16, lbo, sbo)
                               : make_desc_kmajor(sA, ks * 256, 128, 128 * (KT / 8));
      const uint64_t bd = b_mn ? make_desc_kmajor(sB, ks * 16 * 16, variant == 1 ? gs : 128, variant == 1 ? 128 : gs)
                               : make_desc_kmajor(sB, ks * 256, 128, 128 * (KT / 8));
      mma_f16(tmem, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;   // TMEM lane
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
  const int N = 64;
  std::vector<__half> A(128 * KT), B(N * KT);
  std::vector<float> Af(128 * KT), Bf(N * KT), D(128 * N);
  srand(5);
  for (size_t i = 0; i < A.size(); i++) { A[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Af[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); i++) { B[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Bf[i] = __half2float(B[i]); }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * 128 * KP) * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int fails = 0;
  struct Case { int M, a_mn, b_mn, variant, shift; const char* what; };
  const Case cases[] = {
      {128, 0, 0, 0, 0, "K-major A, K-major B (control)"},
      {128, 1, 1, 0, 0, "MN-major A and B, LBO=128 SBO=group"},
      {128, 1, 1, 1, 0, "MN-major A and B, LBO=group SBO=128"},
      {128, 1, 0, 0, 0, "MN-major A, K-major B, LBO=128 SBO=group"},
      {128, 1, 0, 1, 0, "MN-major A, K-major B, LBO=group SBO=128"},
      {128, 0, 1, 0, 0, "K-major A, MN-major B, LBO=128 SBO=group"},
      {128, 0, 1, 1, 0, "K-major A, MN-major B, LBO=group SBO=128"},
      {128, 1, 1, 0, 3, "MN-major, base shifted by 3 K rows (48 B), LBO=128 SBO=group"},
      {128, 1, 1, 1, 3, "MN-major, base shifted by 3 K rows (48 B), LBO=group SBO=128"},
      {64, 0, 0, 0, 0, "M=64 K-major (lane map)"},
      {64, 1, 1, 0, 0, "M=64 MN-major LBO=128 SBO=group (lane map)"},
      {64, 1, 1, 1, 0, "M=64 MN-major LBO=group SBO=128 (lane map)"},
      {64, 1, 1, 2, 0, "M=64 MN-major A with OVERLAPPING groups (SBO=16), MN-major B"},
      {64, 1, 0, 2, 0, "M=64 MN-major A with OVERLAPPING groups (SBO=16), K-major B"},
      {128, 1, 1, 2, 0, "M=128 MN-major A with OVERLAPPING groups (SBO=16), MN-major B"},
  };
  for (const Case& c : cases) {
    cudaMemset(dD, 0, D.size() * 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c.M, N, c.a_mn, c.b_mn, c.variant, c.shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<double> ref((size_t)c.M * N);
    for (int m = 0; m < c.M; m++)
      for (int n = 0; n < N; n++) {
        double r = 0;
        for (int k = 0; k < KT; k++) r += (double)Af[m * KT + k] * Bf[n * KT + k];
        ref[(size_t)m * N + n] = r;
      }
    if (c.variant == 2) {
      // reference: A'[8 g + e][k] = A[e][k + g] (zero beyond KT)
      double err = 0;
      int lanes_ok = 0;
      for (int m = 0; m < c.M; m++) {
        const int g = m / 8, e = m % 8;
        const int lane = c.M == 128 ? m : (m / 16) * 32 + m % 16;
        double rowerr = 0;
        for (int n = 0; n < N; n++) {
          double r = 0;
          for (int k = 0; k + g < KT && k < KT; k++) r += (double)Af[e * KT + k + g] * Bf[n * KT + k];
          rowerr = fmax(rowerr, fabs(D[lane * N + n] - r));
        }
        err = fmax(err, rowerr);
        lanes_ok += rowerr < 1e-3;
      }
      printf("%-70s max err %.3e (%d of %d rows match) %s\n", c.what, err, lanes_ok, c.M, err < 1e-3 ? "PASS" : "fail");
      continue;
    }
    if (c.M == 128) {
      double err = 0;
      for (int i = 0; i < 128 * N; i++) err = fmax(err, fabs(D[i] - ref[i]));
      printf("%-70s max err %.3e %s\n", c.what, err, err < 1e-3 ? "PASS" : "fail");
      if (err >= 1e-3 && (c.variant == 0 || !c.a_mn)) fails += (c.a_mn == 0 && c.b_mn == 0);
    } else {
      // for every TMEM lane find the logical row it holds (or none)
      printf("%-70s lanes:", c.what);
      int mapped = 0, prev = -2, start_lane = 0;
      for (int l = 0; l < 128; l++) {
        int found = -1;
        for (int m = 0; m < c.M && found < 0; m++) {
          double err = 0;
          for (int n = 0; n < N; n++) err = fmax(err, fabs(D[l * N + n] - ref[(size_t)m * N + n]));
          if (err < 1e-3) found = m;
        }
        if (found >= 0) mapped++;
        if (l == 0) { prev = found; start_lane = 0; }
        else if ((found < 0) != (prev < 0) || (found >= 0 && found != prev + 1)) {
          printf(" [%d..%d]=%s%d..", start_lane, l - 1, prev < 0 ? "none " : "row ", prev < 0 ? 0 : prev - (l - 1 - start_lane));
          start_lane = l;
        }
        prev = found;
      }
      printf(" [%d..127]=%s%d..  (%d lanes hold a row)\n", start_lane, prev < 0 ? "none " : "row ", prev < 0 ? 0 : prev - (127 - start_lane), mapped);
    }
  }
  return fails;
}
