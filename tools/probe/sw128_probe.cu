// Stand-alone probe: a K-major SWIZZLE_128B A operand (one 128-byte row = 64 halves per pixel)
// read at a base shifted by `dx` ROWS.  In the no-swizzle plane layout of conv5_tc.cuh a shift
// of one pixel is 16 bytes, every 128-byte core matrix of the shifted tile straddles two
// shared-memory lines and the fetch costs twice; with 128-byte rows a pixel shift is a whole
// line.  Questions: (1) is the result right for dx = 0..7 and which `base offset` (descriptor
// bits 49..51) does a start address that is not 1024-byte aligned need; (2) cycles per MMA,
// shifted and unshifted, against the no-swizzle tile at a 16-byte offset.
// B stays in the canonical no-swizzle layout (the packed weight images of the kernels).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o sw128_probe sw128_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

constexpr int KT = 64;     // K of the test GEMM = one 128-byte row
constexpr int ROWS = 144;  // plane rows (128 + room for shifts)
constexpr int N = 64;

__host__ __device__ inline uint32_t idesc_f16(int M, int Nn) {
  return (1u << 4) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__host__ __device__ inline int kmaj(int r, int k, int K) {
  return (r >> 3) * (64 * (K >> 3)) + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
}
// swizzled descriptor: SBO = 1024 (8 rows of 128 bytes), layout type 2 = SWIZZLE_128B
__device__ inline uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: swizzled A at row shift dx, base offset = bo;  mode 1: timing of R MMAs (same tile)
// mode 2: timing, no-swizzle A tile at byte offset 16 * dx (the conv5 plane situation)
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int dx, int bo,
                                             int mode, int R, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  uint8_t* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // 1024-aligned
  uint8_t* sA = smem_raw;                                   // ROWS x 128 bytes, swizzled
  __half* sB = reinterpret_cast<__half*>(smem_raw + ROWS * 128);
  __half* sP = sB + N * KT;                                 // no-swizzle planes for mode 2
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  if (smem_u32(smem_raw) & 1023u) __trap();
  for (int i = tid; i < ROWS * KT; i += 128) {
    const int r = i / KT, k = i % KT;
    const int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__half*>(sA + off) = A[i];
  }
  for (int i = tid; i < N * KT; i += 128) sB[kmaj(i / KT, i % KT, KT)] = B[i];
  for (int i = tid; i < 2 * ROWS * 8; i += 128) sP[i] = __float2half(1.f);
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_f16(128, N);
  if (warp == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      if (mode == 0) {
        for (int ks = 0; ks < KT / 16; ks++)
          mma_f16(tmem, desc_sw128(smem_u32(sA) + dx * 128 + ks * 32, bo),
                  make_desc_kmajor(sB, ks * 256, 128, 128 * (KT / 8)), idesc, ks > 0);
      } else if (mode == 1) {
        const uint64_t ad = desc_sw128(smem_u32(sA) + dx * 128, bo);
        const uint64_t bd = make_desc_kmajor(sB, 0, 128, 128 * (KT / 8));
        for (int r = 0; r < R; r++) mma_f16(tmem, ad, bd, idesc, 1);
      } else {   // two 8-channel planes of ROWS entries, tile at entry dx: SBO 128, LBO = plane
        const uint64_t ad = make_desc_kmajor(sP, dx * 16, ROWS * 16, 128);
        const uint64_t bd = make_desc_kmajor(sB, 0, 128, 128 * (KT / 8));
        for (int r = 0; r < R; r++) mma_f16(tmem, ad, bd, idesc, 1);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0 && cycles) *cycles = clock64() - t0;
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  if (mode == 0) {
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      for (int j = 0; j < 8; j++) D[row * N + c + j] = v[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> A(ROWS * KT), B(N * KT);
  std::vector<float> Af(ROWS * KT), Bf(N * KT), D(128 * N);
  srand(9);
  for (size_t i = 0; i < A.size(); i++) { A[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Af[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); i++) { B[i] = __float2half((float)rand() / RAND_MAX - 0.5f); Bf[i] = __half2float(B[i]); }
  __half *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ROWS * 128 + N * KT * 2 + 2 * ROWS * 16 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int dx = 0; dx <= 9; dx++)
    for (int bo : {0, dx & 7}) {
      if (bo == 0 && (dx & 7) == 0 && dx != 0 && false) continue;
      cudaMemset(dD, 0, D.size() * 4);
      probe<<<1, 128, smem>>>(dA, dB, dD, dx, bo, 0, 0, nullptr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("dx %d bo %d: %s\n", dx, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
          double ref = 0;
          for (int k = 0; k < KT; k++) ref += (double)Af[(m + dx) * KT + k] * Bf[n * KT + k];
          worst = fmax(worst, fabs(ref - D[m * N + n]));
        }
      printf("swizzle 128B, rows shifted by %d, base offset %d: max err %.3e %s\n", dx, bo, worst,
             worst < 1e-4 ? "PASS" : "fail");
      if ((dx & 7) == 0) break;   // both variants are the same descriptor
    }
  const int R = 96;
  for (int mode : {1, 2})
    for (int dx : {0, 1, 3, 4}) {
      long long best = 1LL << 60, c;
      for (int rep = 0; rep < 3; rep++) {
        probe<<<1, 128, smem>>>(dA, dB, dD, dx, dx & 7, mode, R, dC);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
        best = c < best ? c : best;
      }
      printf("%s tile, shift %d: %.1f cycles per MMA (M=128 N=64 K=16)\n",
             mode == 1 ? "swizzle-128B" : "no-swizzle plane", dx, (double)best / R);
    }
  printf("done\n");
  return 0;
}
