// Stand-alone probe (not part of the library) for the "plane" A operand of layer 1.
//
// Layer 1 is a 9x9 convolution of ONE input channel, so the im2col matrix A[pixel][tap] is a
// Hankel matrix: A[x][(dy,dx)] = in[y+dy][x+dx].  The tcgen05 shared-memory descriptor of the
// no-swizzle K-major layout addresses element (row m, k) at
//     base + (m/8)*SBO + (m%8)*16 + (k/4)*LBO + (k%4)*4
// so with SBO = 128 the 128 rows of a tile are consecutive 16-byte units, and a plane
//     Qd(s)[c] = float4(in[s][c], in[s+1][c], in[s+2][c], in[s+3][c])
// read at base = &Qd(s)[dx] IS the im2col block of taps (dy = s-y .. s-y+3, dx) for 128
// consecutive pixels -- no im2col copy is ever made.  The second 16-byte K chunk of an MMA
// (K = 8 TF32) is another (plane, dx) reached through LBO.  The last filter row (dy = 8) comes
// from a plane of horizontally packed pixels H(r)[c] = in[r][c..c+3].
//
// 1. correctness: one 128-pixel x 64-channel tile through 11 K-steps of 3xTF32 vs FP64.
// 2. timing: cycles per tcgen05.mma for the operand layouts under discussion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o plane_probe plane_probe.cu && ./plane_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"

using namespace srcnn::tc;

constexpr int M = 128, N1 = 64, F1 = 9, KS = 11, K1 = KS * 8;
constexpr int PW = 144;            // plane entries (float4 each)
constexpr int IN_W = PW + 4;       // input columns staged

// (k-step, chunk j, element e) -> filter tap, or -1
__host__ __device__ inline int tap_of(int s, int j, int e) {
  int dy, dx;
  if (s == 0) { dy = 8; dx = 4 * j + e; }
  else if (s == 1) { if (j == 0) { if (e) return -1; dy = 8; dx = 8; } else { dy = e; dx = 8; } }
  else if (s < 6) { dy = e; dx = 2 * (s - 2) + j; }
  else if (s < 10) { dy = 4 + e; dx = 2 * (s - 6) + j; }
  else { if (j) return -1; dy = 4 + e; dx = 8; }
  return dy * F1 + dx;
}

__global__ void __launch_bounds__(128) plane_conv_kernel(const float* __restrict__ in,  // [12][IN_W]
                                                         const float* __restrict__ W,   // [81][64]
                                                         float* out) {                  // [128][64]
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sHh = reinterpret_cast<float*>(smem_raw);   // H(8)
  float* sHl = sHh + PW * 4;
  float* sQ0h = sHl + PW * 4;                        // Qd(0)
  float* sQ0l = sQ0h + PW * 4;
  float* sQ4h = sQ0l + PW * 4;                       // Qd(4)
  float* sQ4l = sQ4h + PW * 4;
  float* sWh = sQ4l + PW * 4;
  float* sWl = sWh + N1 * K1;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;

  for (int c = tid; c < PW; c += 128) {
    for (int e = 0; e < 4; e++) {
      float hi, lo;
      split_tf32(in[8 * IN_W + c + e], hi, lo);
      sHh[c * 4 + e] = hi; sHl[c * 4 + e] = lo;
      split_tf32(in[e * IN_W + c], hi, lo);
      sQ0h[c * 4 + e] = hi; sQ0l[c * 4 + e] = lo;
      split_tf32(in[(4 + e) * IN_W + c], hi, lo);
      sQ4h[c * 4 + e] = hi; sQ4l[c * 4 + e] = lo;
    }
  }
  for (int i = tid; i < N1 * K1; i += 128) {
    const int n = i / K1, k = i % K1;
    const int t = tap_of(k >> 3, (k >> 2) & 1, k & 3);
    float hi, lo;
    split_tf32(t >= 0 ? W[t * N1 + n] : 0.f, hi, lo);
    sWh[kmajor_offset(n, k, K1)] = hi;
    sWl[kmajor_offset(n, k, K1)] = lo;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N1);
    const uint32_t wsbo = 128 * (K1 / 4);
    for (int s = 0; s < KS; s++) {
      const float *ph, *pl;
      uint32_t off, lbo;
      if (s == 0) { ph = sHh; pl = sHl; off = 0; lbo = 64; }
      else if (s == 1) { ph = sHh; pl = sHl; off = 8 * 16; lbo = (uint32_t)((sQ0h + 8 * 4) - (sHh + 8 * 4)) * 4; }
      else if (s < 6) { ph = sQ0h; pl = sQ0l; off = 2 * (s - 2) * 16; lbo = 16; }
      else if (s < 10) { ph = sQ4h; pl = sQ4l; off = 2 * (s - 6) * 16; lbo = 16; }
      else { ph = sQ4h; pl = sQ4l; off = 8 * 16; lbo = 16; }
      const uint64_t ah = make_desc_kmajor(ph, off, lbo, 128);
      const uint64_t al = make_desc_kmajor(pl, off, lbo, 128);
      const uint64_t bh = make_desc_kmajor(sWh, s * 256, 128, wsbo);
      const uint64_t bl = make_desc_kmajor(sWl, s * 256, 128, wsbo);
      mma_tf32(tmem, al, bh, idesc, s > 0);
      mma_tf32(tmem, ah, bl, idesc, 1);
      mma_tf32(tmem, ah, bh, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N1; c += 8) {
    float v[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 8; j++) out[row * N1 + c + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint32_t make_idesc_f16(int Mm, int Nn) {
  return (1u << 4) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(Mm >> 4) << 24);
}
// ------------------------------------------------------------------------------- timing
// mode 0: SS, A = explicit im2col [128][88] no-swizzle (SBO 2816, LBO 128), N = 64, 33 MMAs
// mode 1: SS, A = planes (SBO 128, LBO 16), N = 64, 33 MMAs
// mode 2: TS, A in TMEM, N = 64, 33 MMAs
// mode 3: TS, A in TMEM, N = 32, 24 MMAs (layer 2)
// mode 4: SS planes, hi pass with N = 128 (B = [Wh;Wl]) + lo pass N = 64: 22 MMAs
// mode 5: SS planes N=64 33 + TS N=32 24 + TS N=32 12 (one whole tile)
// mode 6: SS, A = explicit im2col, N = 256 (reference point: a "normal" GEMM shape), 11 MMAs
__global__ void __launch_bounds__(128) mma_time_kernel(int mode, int reps, int sync_each,
                                                       long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* s = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) s[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t id64 = make_idesc_tf32(M, 64), id32 = make_idesc_tf32(M, 32),
                   id128 = make_idesc_tf32(M, 128), id256 = make_idesc_tf32(M, 256);
    float* A = s;                     // 45 KB (or planes)
    float* B = s + 24 * 1024;         // at 96 KB: up to 256 x 88 x 4 = 90 KB
    const uint32_t sbo = 128 * (K1 / 4);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (mode == 0 || mode == 6) {
        for (int ks = 0; ks < KS; ks++)
          for (int p = 0; p < (mode == 0 ? 3 : 1); p++)
            mma_tf32(tmem, make_desc_kmajor(A, ks * 256, 128, sbo),
                     make_desc_kmajor(B, ks * 256, 128, sbo), mode == 0 ? id64 : id256,
                     (ks | p) > 0);
      } else if (mode == 1 || mode == 5) {
        for (int ks = 0; ks < KS; ks++)
          for (int p = 0; p < 3; p++)
            mma_tf32(tmem, make_desc_kmajor(A, (ks % 3) * 2304 + (ks & 7) * 16 + p * 8192, 16, 128),
                     make_desc_kmajor(B, ks * 256, 128, sbo), id64, (ks | p) > 0);
        if (mode == 5) {
          for (int ks = 0; ks < 8; ks++)
            for (int p = 0; p < 3; p++)
              mma_tf32_ts(tmem + 384, tmem + 128 + ks * 8, make_desc_kmajor(B, ks * 256, 128, 2048),
                          id32, (ks | p) > 0);
          for (int ks = 0; ks < 4; ks++)
            for (int p = 0; p < 3; p++)
              mma_tf32_ts(tmem + 416, tmem + 448 + ks * 8, make_desc_kmajor(B, ks * 256, 128, 1024),
                          id32, (ks | p) > 0);
        }
      } else if (mode == 2) {
        for (int ks = 0; ks < KS; ks++)
          for (int p = 0; p < 3; p++)
            mma_tf32_ts(tmem, tmem + 128 + ks * 8, make_desc_kmajor(B, ks * 256, 128, sbo), id64,
                        (ks | p) > 0);
      } else if (mode == 3) {
        for (int ks = 0; ks < 8; ks++)
          for (int p = 0; p < 3; p++)
            mma_tf32_ts(tmem, tmem + 128 + ks * 8, make_desc_kmajor(B, ks * 256, 128, 2048), id32,
                        (ks | p) > 0);
      } else if (mode == 4) {
        for (int ks = 0; ks < KS; ks++) {
          mma_tf32(tmem, make_desc_kmajor(A, (ks % 3) * 2304 + (ks & 7) * 16, 16, 128),
                   make_desc_kmajor(B, ks * 256, 128, sbo), id128, ks > 0);
          mma_tf32(tmem, make_desc_kmajor(A, (ks % 3) * 2304 + (ks & 7) * 16 + 8192, 16, 128),
                   make_desc_kmajor(B, ks * 256, 128, sbo), id64, 1);
        }
      }
      else if (mode == 7) {   // TS tf32 N=32 M=64
        for (int i = 0; i < 24; i++)
          mma_tf32_ts(tmem, tmem + 128 + (i & 7) * 8, make_desc_kmajor(B, (i & 7) * 256, 128, 2048),
                      make_idesc_tf32(64, 32), i > 0);
      } else if (mode == 8) {   // TS f16 N=32 M=128
        for (int i = 0; i < 24; i++)
          mma_f16_ts(tmem, tmem + 128 + (i & 7) * 8, make_desc_kmajor(B, (i & 7) * 256, 128, 2048),
                     make_idesc_f16(128, 32), i > 0);
      } else if (mode == 9) {   // SS f16 planes N=128
        for (int i = 0; i < 24; i++)
          mma_f16_ss(tmem, make_desc_kmajor(A, (i % 3) * 2304 + (i & 7) * 16, 16, 128),
                     make_desc_kmajor(B, (i & 7) * 256, 128, 2816), make_idesc_f16(128, 128), i > 0);
      } else if (mode == 10) {  // TS tf32 N=8
        for (int i = 0; i < 24; i++)
          mma_tf32_ts(tmem, tmem + 128 + (i & 7) * 8, make_desc_kmajor(B, (i & 7) * 256, 128, 2048),
                      make_idesc_tf32(128, 8), i > 0);
      } else if (mode == 11) {  // SS tf32 planes N=128
        for (int i = 0; i < 24; i++)
          mma_tf32(tmem, make_desc_kmajor(A, (i % 3) * 2304 + (i & 7) * 16, 16, 128),
                   make_desc_kmajor(B, (i & 7) * 256, 128, 2816), id128, i > 0);
      } else if (mode == 12) {  // SS tf32 planes N=256
        for (int i = 0; i < 24; i++)
          mma_tf32(tmem, make_desc_kmajor(A, (i % 3) * 2304 + (i & 7) * 16, 16, 128),
                   make_desc_kmajor(B, (i & 7) * 256, 128, 2816), id256, i > 0);
      } else if (mode == 13) {  // TS tf32 N=128
        for (int i = 0; i < 24; i++)
          mma_tf32_ts(tmem, tmem + 256 + (i & 7) * 8, make_desc_kmajor(B, (i & 7) * 256, 128, 2816),
                      id128, i > 0);
      } else if (mode == 14) {  // TS f16 N=64
        for (int i = 0; i < 24; i++)
          mma_f16_ts(tmem, tmem + 128 + (i & 7) * 8, make_desc_kmajor(B, (i & 7) * 256, 128, 2048),
                     make_idesc_f16(128, 64), i > 0);
      }
      if (sync_each || r == reps - 1) {
        mma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// two (or four) issuing threads in different warps, each issuing TS N=32 MMAs into its own accumulator
__global__ void __launch_bounds__(128) mma_time_multi_kernel(int n_issuers, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* s = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) s[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if ((tid & 31) == 0 && warp < n_issuers) {
    const uint32_t id32 = make_idesc_tf32(M, 32);
    for (int r = 0; r < reps; r++)
      for (int i = 0; i < 24 / n_issuers; i++)
        mma_tf32_ts(tmem + 32 * warp, tmem + 128 + 64 * warp + (i & 7) * 8,
                    make_desc_kmajor(s, (i & 7) * 256 + warp * 8192, 128, 2048), id32, i > 0);
    mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) *cycles = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// the MMA mix of one tile of the plane kernel (22 SS + 16 TS + 8 TS), one issuer warp per layer,
// accumulation chains as in the kernel, no epilogues / no data dependencies between layers.
// mask bit 0: I1, bit 1: I2, bit 2: I3;  variant 1: I1 alternates two accumulators (two tiles
// interleaved) instead of one dependent chain
__global__ void __launch_bounds__(128) mma_mix_kernel(int mask, int variant, int reps, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* s = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) s[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  float* A = s;                     // planes
  float* B = s + 16 * 1024;         // at 64 KB: 128 x 88 x 4 = 45 KB
  float* B2 = s + 30 * 1024;        // at 120 KB
  if ((tid & 31) == 0 && ((mask >> warp) & 1) && warp < 3) {
    const uint32_t id32 = make_idesc_tf32(M, 32), id64 = make_idesc_tf32(M, 64), id128 = make_idesc_tf32(M, 128);
    const uint64_t a0 = make_desc_kmajor(A, 0, 16, 128), b0 = make_desc_kmajor(B, 0, 128, 2816);
    const uint64_t w2 = make_desc_kmajor(B2, 0, 128, 2048), w3 = make_desc_kmajor(B2, 16384, 128, 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (warp == 0) {
        const uint32_t d = tmem + 128 * (r & 1);
        if (variant == 0) {
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            mma_tf32(d, a0 + 2 * ks, b0 + 16 * ks, id128, ks > 0);
            mma_tf32(d, a0 + 512 + 2 * ks, b0 + 16 * ks, id64, 1);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            mma_tf32(tmem, a0 + 2 * ks, b0 + 16 * ks, id128, ks > 0);
            mma_tf32(tmem + 128, a0 + 144 + 2 * ks, b0 + 16 * ks, id128, ks > 0);
            mma_tf32(tmem, a0 + 512 + 2 * ks, b0 + 16 * ks, id64, 1);
            mma_tf32(tmem + 128, a0 + 656 + 2 * ks, b0 + 16 * ks, id64, 1);
          }
          r++;
        }
      } else if (warp == 1) {
        const uint32_t d = tmem + 256 + 64 * (r & 1), a2 = tmem + 128 * (r & 1);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          mma_tf32_ts(d, a2 + ks * 8, w2 + 16 * ks, id64, ks > 0);
          mma_tf32_ts(d, a2 + 64 + ks * 8, w2 + 16 * ks, id32, 1);
        }
      } else {
        const uint32_t d = tmem + 384 + 64 * (r & 1), a3 = tmem + 256 + 64 * (r & 1);
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          mma_tf32_ts(d, a3 + ks * 8, w3 + 16 * ks, id64, ks > 0);
          mma_tf32_ts(d, a3 + 32 + ks * 8, w3 + 16 * ks, id32, 1);
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
  // ---------------- correctness of the plane descriptors
  std::vector<float> in(12 * IN_W), W(81 * N1);
  srand(5);
  for (auto& v : in) v = (float)rand() / RAND_MAX - 0.4f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
  float *din, *dW, *dout;
  cudaMalloc(&din, in.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dout, M * N1 * 4);
  cudaMemcpy(din, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(6 * PW * 4 + 2 * N1 * K1) * 4;
  cudaFuncSetAttribute(plane_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  plane_conv_kernel<<<1, 128, smem>>>(din, dW, dout);
  cudaError_t e = cudaDeviceSynchronize();
  int rc = 0;
  if (e != cudaSuccess) {
    printf("PLANE PROBE FAIL: %s\n", cudaGetErrorString(e));
    return 1;
  }
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
  printf("plane conv: max|ref| %.4f  max err %.3e\n", mag, err);
  printf(err < 2e-5 ? "PLANE PROBE PASS\n" : "PLANE PROBE FAIL\n");
  rc |= err < 2e-5 ? 0 : 1;

  // ---------------- timing
  long long* dcy;
  cudaMalloc(&dcy, 8);
  cudaFuncSetAttribute(mma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_mma[15] = {33, 33, 33, 24, 22, 69, 11, 24, 24, 24, 24, 24, 24, 24, 24};
  const char* names[15] = {"SS im2col N64 x33", "SS planes N64 x33", "TS N64 x33", "TS N32 K64 x24",
                          "SS planes N128+N64 x22", "tile: SS planes 33 + TS 24 + TS 12",
                          "SS im2col N256 x11", "TS tf32 M64 N32", "TS f16 M128 N32",
                          "SS f16 planes N128", "TS tf32 N8", "SS tf32 planes N128",
                          "SS tf32 planes N256", "TS tf32 N128", "TS f16 N64"};
  for (int grid : {1})
    for (int mode = 3; mode < 3; mode++)
      for (int sync_each = 0; sync_each < 2; sync_each++) {
        const int reps = 200;
        mma_time_kernel<<<grid, 128, 200 * 1024>>>(mode, reps, sync_each, dcy);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("TIMING FAIL mode %d: %s\n", mode, cudaGetErrorString(e));
          return 1;
        }
        long long cy;
        cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d  %-38s %s: %8.1f cyc/rep  %6.1f cyc/mma\n", grid, names[mode],
               sync_each ? "sync each rep" : "back to back ", (double)cy / reps,
               (double)cy / reps / n_mma[mode]);
      }
  cudaFuncSetAttribute(mma_time_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int ni : {1}) {
    mma_time_multi_kernel<<<1, 128, 64 * 1024>>>(ni, 200, dcy);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("MULTI FAIL: %s\n", cudaGetErrorString(e)); return 1; }
    long long cy;
    cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost);
    printf("TS tf32 N32, 24 MMAs per rep split over %d issuing warps: %8.1f cyc/rep  %6.1f cyc/mma\n", ni,
           (double)cy / 200, (double)cy / 200 / 24);
  }
  cudaFuncSetAttribute(mma_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  {
    long long* d4;
    cudaMalloc(&d4, 32);
    for (int variant = 0; variant < 2; variant++)
      for (int mask : {1, 2, 4, 3, 5, 6, 7}) {
        cudaMemset(d4, 0, 32);
        mma_mix_kernel<<<1, 128, 160 * 1024>>>(mask, variant, 200, d4);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("MIX FAIL: %s\n", cudaGetErrorString(e)); return 1; }
        long long cy[4];
        cudaMemcpy(cy, d4, 32, cudaMemcpyDeviceToHost);
        printf("mix variant %d mask %d: per tile  I1 %7.1f  I2 %7.1f  I3 %7.1f\n", variant, mask,
               cy[0] / 200.0, cy[1] / 200.0, cy[2] / 200.0);
      }
  }
  return rc;
}
