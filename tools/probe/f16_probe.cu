// Stand-alone probe for an FP16-split variant of the fused kernel (x = hi + lo * 2^-11 with hi, lo
// halves, scaled so that they stay in the normal range; products hi*whi, hi*wlo, lo*whi).
//  A. layer 1 through "oct planes" (8 vertically packed halves per 16 bytes) + a plane of 8
//     horizontally packed pixels, K = 16 per MMA: 6 K-steps instead of 11; stacked weights;
//     the lo-scaled products accumulate in the second half of the accumulator.
//  B. A operand in tensor memory as packed half pairs (layer 2 shape), same scheme.
//  C. cycles per tile of the resulting instruction mix with one issuer per layer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16_probe f16_probe.cu && ./f16_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

constexpr int M = 128, N1 = 64, F1 = 9, KS = 6, K1 = KS * 16;   // K in halves
constexpr int PW = 144;                                         // plane entries (16 bytes each)
constexpr int IN_W = PW + 8;

__host__ __device__ inline uint32_t make_idesc_f16(int Mm, int Nn) {
  return (1u << 4) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(Mm >> 4) << 24);   // A,B = F16, D = F32
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// K-major canonical layout for 16-bit elements: core matrix = 8 rows x 16 bytes (8 halves)
__host__ __device__ inline int kmajor16(int r, int k, int K) {   // offset in halves
  return (r >> 3) * (64 * (K >> 3)) + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
}
// x*s = hi + lo / 2048, hi and lo halves of comparable magnitude
__device__ __forceinline__ void split_h(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = __float2half_rn((xs - __half2float(hi)) * 2048.f);
}
// (K-step s, chunk j, element e) -> tap or -1: chunks = O dx 0..8, H8 dx' 0, 8
__host__ __device__ inline int tap16(int s, int j, int e) {
  const int c = 2 * s + j;                 // chunk index 0..11
  if (c < 9) return e * F1 + c;            // oct plane at dx = c: taps (dy = e, dx = c)
  if (c == 9) return 8 * F1 + e;           // H8 at dx' = 0: taps (8, e)
  if (c == 10) return e == 0 ? 8 * F1 + 8 : -1;   // H8 at dx' = 8: tap (8,8) + 7 pads
  return -1;                               // dummy chunk
}

__global__ void __launch_bounds__(128) conv16_kernel(const float* __restrict__ in,   // [9][IN_W]
                                                     const float* __restrict__ W,    // [81][64]
                                                     float sx, float sw, float* out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __half* sOh = reinterpret_cast<__half*>(smem_raw);   // oct plane rows 0..7
  __half* sOl = sOh + PW * 8;
  __half* sHh = sOl + PW * 8;                          // H8(8): above the oct planes
  __half* sHl = sHh + PW * 8;
  __half* sW = sHl + PW * 8;                           // [128][K1]: rows 0..63 hi, 64..127 lo
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int c = tid; c < PW; c += 128)
    for (int e = 0; e < 8; e++) {
      __half h, l;
      split_h(in[e * IN_W + c] * sx, h, l);
      sOh[c * 8 + e] = h; sOl[c * 8 + e] = l;
      split_h(in[8 * IN_W + c + e] * sx, h, l);
      sHh[c * 8 + e] = h; sHl[c * 8 + e] = l;
    }
  for (int i = tid; i < 2 * N1 * K1; i += 128) {
    const int n = i / K1, k = i % K1;
    const int t = tap16(k >> 4, (k >> 3) & 1, k & 7);
    __half h, l;
    split_h(t >= 0 ? W[t * N1 + (n & 63)] * sw : 0.f, h, l);
    sW[kmajor16(n, k, K1)] = n < N1 ? h : l;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t id128 = make_idesc_f16(M, 128), id64 = make_idesc_f16(M, 64);
    const uint32_t wsbo = 128 * (K1 / 8);
    for (int s = 0; s < KS; s++) {
      const __half *ph, *pl;
      uint32_t off, lbo;
      if (s < 4) { ph = sOh; pl = sOl; off = 2 * s * 16; lbo = 16; }
      else if (s == 4) { ph = sOh; pl = sOl; off = 8 * 16; lbo = (uint32_t)((sHh - sOh) * 2) - 8 * 16; }
      else { ph = sHh; pl = sHl; off = 8 * 16; lbo = 16; }
      const uint64_t ah = make_desc_kmajor(ph, off, lbo, 128), al = make_desc_kmajor(pl, off, lbo, 128);
      const uint64_t bw = make_desc_kmajor(sW, s * 256, 128, wsbo);
      mma_f16_ss(tmem, ah, bw, id128, s > 0);          // [0,64) += hi.whi ; [64,128) += hi.wlo
      mma_f16_ss(tmem + 64, al, bw, id64, 1);          // [64,128) += lo.whi
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  const float inv = 1.f / (sx * sw);
  for (int c = 0; c < N1; c += 8) {
    float v[8], w[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c, w);
    for (int j = 0; j < 8; j++) out[row * N1 + c + j] = (v[j] + w[j] * (1.f / 2048.f)) * inv;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// B: A[128][64] in TMEM as packed half pairs (hi: columns 64..95, lo: 96..127), W2 [32][64]
__global__ void __launch_bounds__(128) ts16_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   float sa, float sw, float* out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __half* sW = reinterpret_cast<__half*>(smem_raw);   // [64][64]: rows 0..31 hi, 32..63 lo
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  constexpr int K2 = 64, N2 = 32;
  for (int i = tid; i < 2 * N2 * K2; i += 128) {
    const int n = i / K2, k = i % K2;
    __half h, l;
    split_h(B[(n & 31) * K2 + k] * sw, h, l);
    sW[kmajor16(n, k, K2)] = n < N2 ? h : l;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const int row = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < K2 / 2; c += 8) {   // 8 columns = 16 K elements
    float hi[8], lo[8];
    for (int j = 0; j < 8; j++) {
      __half h0, l0, h1, l1;
      split_h(A[row * K2 + 2 * (c + j)] * sa, h0, l0);
      split_h(A[row * K2 + 2 * (c + j) + 1] * sa, h1, l1);
      hi[j] = __uint_as_float((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16));
      lo[j] = __uint_as_float((uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16));
    }
    tmem_st8(tmem + lane_base + 64 + c, hi);
    tmem_st8(tmem + lane_base + 96 + c, lo);
  }
  tmem_st_wait();
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t id64 = make_idesc_f16(M, 64), id32 = make_idesc_f16(M, 32);
    const uint64_t bw = make_desc_kmajor(sW, 0, 128, 128 * (K2 / 8));
    for (int ks = 0; ks < K2 / 16; ks++) {
      mma_f16_ts(tmem, tmem + 64 + ks * 8, bw + 16 * ks, id64, ks > 0);
      mma_f16_ts(tmem + 32, tmem + 96 + ks * 8, bw + 16 * ks, id32, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const float inv = 1.f / (sa * sw);
  for (int c = 0; c < N2; c += 8) {
    float v[8], w[8];
    tmem_ld8(tmem + lane_base + c, v);
    tmem_ld8(tmem + lane_base + 32 + c, w);
    for (int j = 0; j < 8; j++) out[row * N2 + c + j] = (v[j] + w[j] * (1.f / 2048.f)) * inv;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// C: instruction mix of one tile, one issuer warp per layer (bit 0: I1, 1: I2, 2: I3)
__global__ void __launch_bounds__(128) mix16_kernel(int mask, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* s = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 100 * 1024 / 4; i += 128) s[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if ((tid & 31) == 0 && ((mask >> warp) & 1) && warp < 3) {
    const uint32_t id32 = make_idesc_f16(M, 32), id64 = make_idesc_f16(M, 64), id128 = make_idesc_f16(M, 128);
    const uint64_t a0 = make_desc_kmajor(s, 0, 16, 128), b0 = make_desc_kmajor(s, 32768, 128, 128 * (K1 / 8));
    const uint64_t w2 = make_desc_kmajor(s, 65536, 128, 1024), w3 = make_desc_kmajor(s, 81920, 128, 512);
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (warp == 0) {
        const uint32_t d = tmem + 128 * (r & 1);
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          mma_f16_ss(d, a0 + 2 * ks, b0 + 16 * ks, id128, ks > 0);
          mma_f16_ss(d + 64, a0 + 512 + 2 * ks, b0 + 16 * ks, id64, 1);
        }
      } else if (warp == 1) {
        const uint32_t d = tmem + 256 + 64 * (r & 1), a2 = tmem + 128 * (r & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          mma_f16_ts(d, a2 + ks * 8, w2 + 16 * ks, id64, ks > 0);
          mma_f16_ts(d + 32, a2 + 32 + ks * 8, w2 + 16 * ks, id32, 1);
        }
      } else {
        const uint32_t d = tmem + 384 + 64 * (r & 1), a3 = tmem + 256 + 64 * (r & 1);
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          mma_f16_ts(d, a3 + ks * 8, w3 + 16 * ks, id64, ks > 0);
          mma_f16_ts(d + 32, a3 + 16 + ks * 8, w3 + 16 * ks, id32, 1);
        }
      }
    }
    mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    cycles[warp] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  int rc = 0;
  {  // A
    std::vector<float> in(9 * IN_W), W(81 * N1);
    srand(5);
    for (auto& v : in) v = (float)rand() / RAND_MAX - 0.4f;
    for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
    float *din, *dW, *dout;
    cudaMalloc(&din, in.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dout, M * N1 * 4);
    cudaMemcpy(din, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(4 * PW * 8 + 2 * N1 * K1) * 2;
    cudaFuncSetAttribute(conv16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv16_kernel<<<1, 128, smem>>>(din, dW, 1024.f, 65536.f, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("A FAIL: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(M * N1);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, mag = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N1; n++) {
        double r = 0;
        for (int dy = 0; dy < 9; dy++)
          for (int dx = 0; dx < 9; dx++) r += (double)in[dy * IN_W + m + dx] * W[(dy * 9 + dx) * N1 + n];
        err = fmax(err, fabs(out[m * N1 + n] - r));
        mag = fmax(mag, fabs(r));
      }
    printf("A. oct-plane f16-split conv: max|ref| %.4f  max err %.3e  %s\n", mag, err, err < 2e-5 ? "PASS" : "FAIL");
    rc |= err < 2e-5 ? 0 : 1;
  }
  {  // B
    constexpr int K2 = 64, N2 = 32;
    std::vector<float> A(M * K2), B(N2 * K2);
    srand(11);
    for (auto& v : A) v = fmaxf((float)rand() / RAND_MAX - 0.3f, 0.f);
    for (auto& v : B) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, M * N2 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    ts16_kernel<<<1, 128, 2 * N2 * K2 * 2>>>(dA, dB, 4096.f, 65536.f, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("B FAIL: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(M * N2);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N2; n++) {
        double r = 0;
        for (int k = 0; k < K2; k++) r += (double)A[m * K2 + k] * B[n * K2 + k];
        err = fmax(err, fabs(D[m * N2 + n] - r));
      }
    printf("B. A in TMEM (packed halves) f16-split: max err %.3e  %s\n", err, err < 2e-5 ? "PASS" : "FAIL");
    rc |= err < 2e-5 ? 0 : 1;
  }
  {  // C
    long long* d4;
    cudaMalloc(&d4, 32);
    cudaFuncSetAttribute(mix16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int mask : {1, 2, 4, 7}) {
      cudaMemset(d4, 0, 32);
      mix16_kernel<<<1, 128, 100 * 1024>>>(mask, 200, d4);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C FAIL: %s\n", cudaGetErrorString(e)); return 1; }
      long long cy[4];
      cudaMemcpy(cy, d4, 32, cudaMemcpyDeviceToHost);
      printf("C. f16 mix mask %d: per tile  I1 %7.1f  I2 %7.1f  I3 %7.1f\n", mask, cy[0] / 200.0, cy[1] / 200.0, cy[2] / 200.0);
    }
  }
  return rc;
}
