// Shape-specialised FP32 SIMT kernels for the training step of the SRCNN shapes.
//
// The any-shape GEMM core (kernels_generic.cuh) wastes 63/64 of its 64-wide N tile on the
// last layer (n = 1) and half of it on n = 32, and its gather costs integer divisions in the
// inner loop.  These kernels cover the shapes that dominate an epoch:
//   * the n = 1 layer (layer 3): forward, delta of the layer below, weight gradient -- lanes are
//     the k input channels, a warp reduces with shuffles (forward) or keeps per-tap sums (gW)
//   * weight gradients of the f = 1 layer (layer 2) and of the k = 1 layer (layer 1): each
//     lane owns an 8x8 register tile of the gradient and streams pixels straight from
//     global/L1 (64 FFMA per 4 LDG.128); warps hold disjoint pixel ranges ("slices"); partial
//     tiles go to scratch and are summed in a fixed order, so the result is deterministic
//     (the reference's `+=` on grad_w races across samples, backpropagate.cl:110).
#pragma once
#include <cuda_runtime.h>

#include "context.cuh"

namespace srcnn {
namespace train {

constexpr int NT = 256;

// ------------------------------------------------------------------ layer-3 forward ---------
// out[s][y][x] = b + sum_{dy,dx,c} W[dy][dx][c] * in[s][y+dy][x+dx][c]   (n = 1, skip_relu or not)
// reference: src/kernel/layer_uber_kernel.cl:36-96 with CURRENT_FILTER_COUNT = 1.
// A warp computes a run of P = 8 output pixels of one row; lane = input channel (c, c+32, ..).
template <int F, int CPL>   // CPL = channels per lane = ceil(k / 32)
__global__ void __launch_bounds__(NT) n1_forward_kernel(const float* __restrict__ in,
                                                        float* __restrict__ out,
                                                        const float* __restrict__ W,
                                                        const float* __restrict__ B, int k, int relu,
                                                        int iw, int ih, int ow, int oh, int S) {
  constexpr int P = 8;
  const int lane = threadIdx.x & 31;
  const int runs_per_row = (ow + P - 1) / P;
  const long long total_runs = (long long)S * oh * runs_per_row;
  float w[CPL][F][F];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int dy = 0; dy < F; dy++)
#pragma unroll
      for (int dx = 0; dx < F; dx++) w[j][dy][dx] = c < k ? __ldg(W + (dy * F + dx) * k + c) : 0.f;
  }
  const float bias = __ldg(B);
  for (long long run = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); run < total_runs;
       run += (long long)gridDim.x * (NT / 32)) {
    const int xr = (int)(run % runs_per_row);
    const long long t = run / runs_per_row;
    const int y = (int)(t % oh);
    const long long s = t / oh;
    const int x0 = xr * P;
    float acc[P];
#pragma unroll
    for (int p = 0; p < P; p++) acc[p] = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; j++) {
      const int c = lane + 32 * j;
      const int cc = c < k ? c : 0;
#pragma unroll
      for (int dy = 0; dy < F; dy++) {
        const float* row = in + ((s * ih + y + dy) * iw) * (long long)k + cc;
        float v[P + F - 1];
#pragma unroll
        for (int q = 0; q < P + F - 1; q++) {
          const int x = min(x0 + q, iw - 1);   // clamped: only feeds masked outputs
          v[q] = __ldg(row + (long long)x * k);
        }
#pragma unroll
        for (int dx = 0; dx < F; dx++)
#pragma unroll
          for (int p = 0; p < P; p++) acc[p] = fmaf(v[p + dx], w[j][dy][dx], acc[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
      float v = acc[p];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[p] = v;
    }
    if (lane < P && x0 + lane < ow) {
      float r = 0.f;
#pragma unroll
      for (int p = 0; p < P; p++)
        if (p == lane) r = acc[p];
      r += bias;
      if (relu) r = fmaxf(r, 0.f);
      out[(s * oh + y) * (long long)ow + x0 + lane] = r;
    }
  }
}

// Same arithmetic (same order: bit-identical results), for samples that fit in shared memory
// (training / validation patches): a CTA stages one sample with coalesced 16-byte loads and its
// warps take runs of P output pixels from there -- the overlapping 5-row windows of neighbouring
// runs are then read from shared memory instead of L1 / L2 (C4: 163 -> ~45 us per 2 048 patches).
template <int F, int CPL, int P>
__global__ void __launch_bounds__(NT) n1_forward_smem_kernel(const float* __restrict__ in,
                                                             float* __restrict__ out,
                                                             const float* __restrict__ W,
                                                             const float* __restrict__ B, int k,
                                                             int relu, int iw, int ih, int ow,
                                                             int oh, int S) {
  extern __shared__ __align__(16) float smp[];   // [ih][iw][k]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int runs_per_row = (ow + P - 1) / P;
  const int n4 = ih * iw * k / 4;
  float w[CPL][F][F];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int dy = 0; dy < F; dy++)
#pragma unroll
      for (int dx = 0; dx < F; dx++) w[j][dy][dx] = c < k ? __ldg(W + (dy * F + dx) * k + c) : 0.f;
  }
  const float bias = __ldg(B);
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    __syncthreads();
    // cp.async: every 16-byte piece of the sample is in flight at once (a load -> store loop
    // runs its ~14 iterations per thread as ~14 dependent HBM round trips: that WAS the kernel,
    // 107 us per 3 028 samples of 21x21x32)
    const float4* src = reinterpret_cast<const float4*>(in + (long long)s * ih * iw * k);
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(smp);
    for (int i = threadIdx.x; i < n4; i += NT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + 16u * (uint32_t)i), "l"(src + i)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int run = warp; run < oh * runs_per_row; run += NT / 32) {
      const int y = run / runs_per_row, x0 = (run - y * runs_per_row) * P;
      float acc[P];
#pragma unroll
      for (int p = 0; p < P; p++) acc[p] = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; j++) {
        const int c = lane + 32 * j;
        const int cc = c < k ? c : 0;
#pragma unroll
        for (int dy = 0; dy < F; dy++) {
          const float* row = smp + (y + dy) * iw * k + cc;
          float v[P + F - 1];
#pragma unroll
          for (int q = 0; q < P + F - 1; q++) v[q] = row[min(x0 + q, iw - 1) * k];
#pragma unroll
          for (int dx = 0; dx < F; dx++)
#pragma unroll
            for (int p = 0; p < P; p++) acc[p] = fmaf(v[p + dx], w[j][dy][dx], acc[p]);
        }
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
        float v = acc[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[p] = v;
      }
      if (lane < P && x0 + lane < ow) {
        float r = 0.f;
#pragma unroll
        for (int p = 0; p < P; p++)
          if (p == lane) r = acc[p];
        r += bias;
        if (relu) r = fmaxf(r, 0.f);
        out[((long long)s * oh + y) * ow + x0 + lane] = r;
      }
    }
  }
}

// ------------------------------------------------------------------ delta below an n=1 layer --
// target[s][j][i][c] = [lo[s][j][i][c] > 0] * sum_{dy,dx} W[dy][dx][c] * dn[s][j-dy][i-dx]
// reference: src/kernel/layer_deltas.cl:42-127 with n_next_filter_cnt = 1.
template <int F, int CPL>
__global__ void __launch_bounds__(NT) n1_deltas_kernel(const float* __restrict__ dn,
                                                       const float* __restrict__ lo,
                                                       float* __restrict__ target,
                                                       const float* __restrict__ W, int k, int ow,
                                                       int oh, int S) {
  const int lane = threadIdx.x & 31;
  const int nw = ow - F + 1, nh = oh - F + 1;
  float w[CPL][F][F];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int dy = 0; dy < F; dy++)
#pragma unroll
      for (int dx = 0; dx < F; dx++) w[j][dy][dx] = c < k ? __ldg(W + (dy * F + dx) * k + c) : 0.f;
  }
  const long long total = (long long)S * oh * ow;
  for (long long p = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); p < total;
       p += (long long)gridDim.x * (NT / 32)) {
    const int i = (int)(p % ow);
    const long long t = p / ow;
    const int j = (int)(t % oh);
    const long long s = t / oh;
    float d[F][F];   // warp-uniform window of the next layer's deltas (zero outside)
#pragma unroll
    for (int dy = 0; dy < F; dy++)
#pragma unroll
      for (int dx = 0; dx < F; dx++) {
        const int nj = j - dy, ni = i - dx;
        d[dy][dx] = (nj >= 0 && nj < nh && ni >= 0 && ni < nw)
                        ? __ldg(dn + (s * nh + nj) * (long long)nw + ni)
                        : 0.f;
      }
#pragma unroll
    for (int jc = 0; jc < CPL; jc++) {
      const int c = lane + 32 * jc;
      if (c >= k) continue;
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < F; dy++)
#pragma unroll
        for (int dx = 0; dx < F; dx++) acc = fmaf(d[dy][dx], w[jc][dy][dx], acc);
      const long long idx = p * k + c;
      target[idx] = __ldg(lo + idx) > 0.f ? acc : 0.f;
    }
  }
}

// Same, for samples whose zero-padded next-layer delta map fits in shared memory (training
// patches): one CTA per sample stages the padded map once, so the inner loop has no bounds
// tests and no global-memory latency.
template <int F, int CPL>
__global__ void __launch_bounds__(NT) n1_deltas_smem_kernel(const float* __restrict__ dn,
                                                            const float* __restrict__ lo,
                                                            float* __restrict__ target,
                                                            const float* __restrict__ W, int k,
                                                            int ow, int oh, int S) {
  extern __shared__ float dpad[];   // [(oh + F - 1)][(ow + F - 1)], zero border of F-1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = ow - F + 1, nh = oh - F + 1;
  const int pw = ow + F - 1, ph = oh + F - 1;
  float w[CPL][F][F];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int dy = 0; dy < F; dy++)
#pragma unroll
      for (int dx = 0; dx < F; dx++) w[j][dy][dx] = c < k ? __ldg(W + (dy * F + dx) * k + c) : 0.f;
  }
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < pw * ph; i += NT) {
      const int y = i / pw - (F - 1), x = i % pw - (F - 1);
      dpad[i] = (y >= 0 && y < nh && x >= 0 && x < nw)
                    ? __ldg(dn + ((long long)s * nh + y) * nw + x)
                    : 0.f;
    }
    __syncthreads();
    for (int p = warp; p < ow * oh; p += NT / 32) {
      const int j = p / ow, i = p - j * ow;
      // target pixel (j,i) sees dn[j-dy][i-dx] = dpad[j-dy+F-1][i-dx+F-1]
      const float* base = dpad + (j + F - 1) * pw + i + F - 1;
      float d[F][F];
#pragma unroll
      for (int dy = 0; dy < F; dy++)
#pragma unroll
        for (int dx = 0; dx < F; dx++) d[dy][dx] = base[-dy * pw - dx];
#pragma unroll
      for (int jc = 0; jc < CPL; jc++) {
        const int c = lane + 32 * jc;
        if (c >= k) continue;
        float acc = 0.f;
#pragma unroll
        for (int dy = 0; dy < F; dy++)
#pragma unroll
          for (int dx = 0; dx < F; dx++) acc = fmaf(d[dy][dx], w[jc][dy][dx], acc);
        const long long idx = ((long long)s * oh * ow + p) * k + c;
        target[idx] = __ldg(lo + idx) > 0.f ? acc : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------ gW of an n=1 layer --------
// gW[dy][dx][c] += sum_{s,row,col} d[s][row][col] * in[s][row+dy][col+dx][c];  gB += sum d
// reference: src/kernel/backpropagate.cl:56-114 with n_current_filter_cnt = 1.
// Warp w of a CTA owns the taps {w, w+8, w+16, w+24}; lane = channel.  A CTA walks its samples
// in a fixed order and writes ONE partial vector [f*f*k + 1] to scratch.
template <int F, int CPL>
__global__ void __launch_bounds__(NT) n1_gradw_kernel(const float* __restrict__ d,
                                                      const float* __restrict__ in,
                                                      float* __restrict__ partial, int k, int ow,
                                                      int oh, int S) {
  constexpr int TPW = (F * F + 7) / 8;   // taps per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int iw = ow + F - 1, ih = oh + F - 1;
  float acc[TPW][CPL];
  float gb = 0.f;
  int toff[TPW];
#pragma unroll
  for (int t = 0; t < TPW; t++) {
    const int tap = warp + 8 * t;
    const int tt = tap < F * F ? tap : 0;
    toff[t] = ((tt / F) * iw + (tt % F)) * k;
#pragma unroll
    for (int j = 0; j < CPL; j++) acc[t][j] = 0.f;
  }
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const float* ds = d + (long long)s * ow * oh;
    const float* ins = in + (long long)s * iw * ih * k;
    for (int row = 0; row < oh; row++)
#pragma unroll 4
      for (int col = 0; col < ow; col++) {
        const float dv = __ldg(ds + row * ow + col);
        gb += dv;
        const float* px = ins + ((long long)row * iw + col) * k;
#pragma unroll
        for (int t = 0; t < TPW; t++)
#pragma unroll
          for (int j = 0; j < CPL; j++) {
            const int c = lane + 32 * j;
            const float v = c < k ? __ldg(px + toff[t] + c) : 0.f;
            acc[t][j] = fmaf(v, dv, acc[t][j]);
          }
      }
  }
  float* dst = partial + (long long)blockIdx.x * (F * F * k + 1);
#pragma unroll
  for (int t = 0; t < TPW; t++) {
    const int tap = warp + 8 * t;
    if (tap < F * F)
#pragma unroll
      for (int j = 0; j < CPL; j++) {
        const int c = lane + 32 * j;
        if (c < k) dst[tap * k + c] = acc[t][j];
      }
  }
  if (threadIdx.x == 0) dst[F * F * k] = gb;
}

// ------------------------------------------------------------------ delta below an f=1 layer --
// target[p][n] = [lo[p][n] > 0] * sum_k W[n][k] * dn[p][k]      (p over all S*oh*ow pixels)
// reference: src/kernel/layer_deltas.cl:42-127 with f_next = 1 (layer 2 of 9-1-5): a plain
// [P x K] . [K x N] GEMM with the ReLU mask of the layer's own output as epilogue.
// CTA = 128 pixels x N outputs, 256 threads, thread tile 8 px x (N/8) outputs; the dn tile is
// staged transposed ([k][px]) so every inner-loop access is a conflict-free LDS.128.
template <int N, int K>
__global__ void __launch_bounds__(256) f1_deltas_kernel(const float* __restrict__ dn,
                                                        const float* __restrict__ lo,
                                                        float* __restrict__ target,
                                                        const float* __restrict__ W, long long P) {
  constexpr int PX = 128, TN = N / 8, PITCH = PX + 4;
  __shared__ __align__(16) float sW[K][N];        // sW[k][n] = W[n][k]
  __shared__ __align__(16) float sD[K][PITCH];    // sD[k][px]
  const int tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * PX;
  for (int i = tid; i < N * K; i += 256) {
    const int n = i / K, k = i - n * K;
    sW[k][n] = __ldg(W + i);
  }
  {
    // lane = pixel (consecutive banks), one 16-byte K-quad per step
    const int px = tid & (PX - 1), q0 = tid >> 7;
    const long long p = p0 + px;
#pragma unroll
    for (int q = q0; q < K / 4; q += 2) {
      const float4 v = p < P ? __ldg(reinterpret_cast<const float4*>(dn + p * K) + q)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      sD[4 * q + 0][px] = v.x;
      sD[4 * q + 1][px] = v.y;
      sD[4 * q + 2][px] = v.z;
      sD[4 * q + 3][px] = v.w;
    }
  }
  __syncthreads();
  // 256 threads = 8 output groups (tx) x 32 pixel groups (ty) of 4 pixels
  const int tx = tid & 7, ty = tid >> 3;
  constexpr int TP = PX / 32;              // pixels per thread
  static_assert(TP == 4 && TN % 4 == 0, "float4 accesses");
  float acc[TP][TN];
#pragma unroll
  for (int i = 0; i < TP; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; k++) {
    float a[TP], b[TN];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sD[k][ty * TP]);
#pragma unroll
    for (int j = 0; j < TN; j += 4)
      *reinterpret_cast<float4*>(b + j) = *reinterpret_cast<const float4*>(&sW[k][tx * TN + j]);
#pragma unroll
    for (int i = 0; i < TP; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < TP; i++) {
    const long long p = p0 + ty * TP + i;
    if (p < P) {
      const long long base = p * N + tx * TN;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 o = __ldg(reinterpret_cast<const float4*>(lo + base + j));
        float4 r;
        r.x = o.x > 0.f ? acc[i][j + 0] : 0.f;
        r.y = o.y > 0.f ? acc[i][j + 1] : 0.f;
        r.z = o.z > 0.f ? acc[i][j + 2] : 0.f;
        r.w = o.w > 0.f ? acc[i][j + 3] : 0.f;
        *reinterpret_cast<float4*>(target + base + j) = r;
      }
    }
  }
}

// ------------------------------------------------------------------ fused backward of layer 3 --
// One pass over out2 for: last_layer_delta (src/kernel/last_layer_delta.cl:14-50, quirk Q2 kept),
// the deltas of layer 2 (src/kernel/layer_deltas.cl:42-127 with n_next = 1) and the weight/bias
// gradient of layer 3 (src/kernel/backpropagate.cl:56-114 with n = 1).  All three consume the
// same 5x5 window of d3 around an out2 pixel:
//     d2[j][i][c]       = [out2[j][i][c] > 0] * sum_{dy,dx} W3[dy][dx][c] * d3[j-dy][i-dx]
//     gW3[dy][dx][c]   += out2[j][i][c] * d3[j-dy][i-dx]
// A CTA (5 warps) takes one sample at a time: d3 is computed into a zero-padded shared-memory
// map, warp w walks the out2 rows w, w+5, .. with lane = channel; the window slides through
// registers (5 LDS per pixel for 50 FFMA).  Weight-gradient sums stay in registers across the
// samples of a CTA (fixed order) and leave as ONE partial vector per CTA -> deterministic.
constexpr int B3_WARPS = 5, B3_NT = 32 * B3_WARPS, B3_F = 5;
template <int CPL>
__global__ void __launch_bounds__(B3_NT) bwd3_fused_kernel(
    const float* __restrict__ gt, const float* __restrict__ out3, const float* __restrict__ out2,
    const float* __restrict__ W3, float* __restrict__ d3, float* __restrict__ d2,
    float* __restrict__ partial, int k, int gt_w, int gt_h, int w3, int h3, int S) {
  constexpr int F = B3_F, T = F * F;
  extern __shared__ float b3_smem[];
  const int ow = w3 + F - 1, oh = h3 + F - 1;       // out2 / d2 extent
  const int pw = w3 + 2 * (F - 1), ph = h3 + 2 * (F - 1);
  float* P = b3_smem;                               // [ph][pw] zero-padded d3
  float* red = b3_smem + ((pw * ph + 3) & ~3);      // [B3_WARPS][T * k] at the end
  __shared__ float gb_part[B3_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pad = (gt_w - w3) / 2;
  float wt[CPL][T], acc[CPL][T];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int t = 0; t < T; t++) {
      wt[j][t] = c < k ? __ldg(W3 + t * k + c) : 0.f;
      acc[j][t] = 0.f;
    }
  }
  float gb = 0.f;   // thread 0 only
  for (int i = threadIdx.x; i < pw * ph; i += B3_NT) P[i] = 0.f;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    __syncthreads();   // previous sample's readers are done with P
    // ---- d3 of this sample -> P (interior), global d3, bias-gradient sum
    float part = 0.f;
    for (int i = threadIdx.x; i < w3 * h3; i += B3_NT) {
      const int y = i / w3, x = i - y * w3;
      const float o = __ldg(out3 + (long long)s * w3 * h3 + i);
      const float t = __ldg(gt + ((long long)s * gt_h + y + pad) * gt_w + pad + x);
      const float dv = o > 0.f ? o - t : 0.f;
      P[(y + F - 1) * pw + x + F - 1] = dv;
      d3[(long long)s * w3 * h3 + i] = dv;
      part += dv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) gb_part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < B3_WARPS; w++) t += gb_part[w];
      gb += t;
    }
    // ---- out2 rows of this warp
    const float* o2s = out2 + (long long)s * ow * oh * k;
    float* d2s = d2 + (long long)s * ow * oh * k;
    for (int j = warp; j < oh; j += B3_WARPS) {
      // window rows: d3 row j-dy lives in P row j-dy+F-1; column i-dx+F-1 = i+t, t = F-1-dx
      const float* prow[F];
#pragma unroll
      for (int dy = 0; dy < F; dy++) prow[dy] = P + (j - dy + F - 1) * pw;
      float win[F][F];   // win[dy][(i + t) % F] = column i+t of row dy, t = 0..F-1
#pragma unroll
      for (int dy = 0; dy < F; dy++)
#pragma unroll
        for (int t = 0; t < F - 1; t++) win[dy][t] = prow[dy][t];
      // out2 values of the next group of F pixels are fetched while this group is processed
      // (one warp has ~4 peers per scheduler: an exposed global load costs ~150 issue slots)
      float vn[CPL][F];
#pragma unroll
      for (int jc = 0; jc < CPL; jc++)
#pragma unroll
        for (int u = 0; u < F; u++) {
          const int c = lane + 32 * jc;
          vn[jc][u] = (c < k && u < ow) ? __ldg(o2s + ((long long)j * ow + u) * k + c) : 0.f;
        }
      for (int i0 = 0; i0 < ow; i0 += F) {
        float vc[CPL][F];
#pragma unroll
        for (int jc = 0; jc < CPL; jc++)
#pragma unroll
          for (int u = 0; u < F; u++) {
            vc[jc][u] = vn[jc][u];
            const int c = lane + 32 * jc, i = i0 + F + u;
            vn[jc][u] = (c < k && i < ow) ? __ldg(o2s + ((long long)j * ow + i) * k + c) : 0.f;
          }
#pragma unroll
        for (int u = 0; u < F; u++) {
          const int i = i0 + u;
          if (i < ow) {
#pragma unroll
            for (int dy = 0; dy < F; dy++) win[dy][(u + F - 1) % F] = prow[dy][i + F - 1];
#pragma unroll
            for (int jc = 0; jc < CPL; jc++) {
              const int c = lane + 32 * jc;
              if (c < k) {
                const float v = vc[jc][u];
                float ds[F];   // one chain per filter row, folded in a fixed order
#pragma unroll
                for (int dy = 0; dy < F; dy++) {
                  ds[dy] = 0.f;
#pragma unroll
                  for (int dx = 0; dx < F; dx++) {
                    const float dv = win[dy][(u + F - 1 - dx) % F];
                    ds[dy] = fmaf(wt[jc][dy * F + dx], dv, ds[dy]);
                    acc[jc][dy * F + dx] = fmaf(v, dv, acc[jc][dy * F + dx]);
                  }
                }
                const float dsum = ((ds[0] + ds[1]) + (ds[2] + ds[3])) + ds[4];
                d2s[((long long)j * ow + i) * k + c] = v > 0.f ? dsum : 0.f;
              }
            }
          }
        }
      }
    }
  }
  // ---- one partial per CTA: warps folded in warp order
  __syncthreads();
#pragma unroll
  for (int jc = 0; jc < CPL; jc++) {
    const int c = lane + 32 * jc;
    if (c < k)
#pragma unroll
      for (int t = 0; t < T; t++) red[(warp * T + t) * k + c] = acc[jc][t];
  }
  __syncthreads();
  float* dst = partial + (long long)blockIdx.x * (T * k + 1);
  for (int i = threadIdx.x; i < T * k; i += B3_NT) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < B3_WARPS; w++) t += red[w * T * k + i];
    dst[i] = t;
  }
  if (threadIdx.x == 0) dst[T * k] = gb;
}

// ------------------------------------------------------------------ gW of f=1 / k=1 layers ---
// gW[m][n] += sum_p A[p][m] * d[p][n];  gB[n] += sum_p d[p][n]       (p over all pixels)
//   MODE 0 (f = 1, layer 2): A[p][m] = in[p][m]                       m < K
//   MODE 1 (k = 1, layer 1): A[p][m] = in[s][row+dy][col+dx], m = dy*F+dx   (gathered)
// reference: src/kernel/backpropagate.cl:56-114.
// A CTA streams blocks of PB pixels through a 2-stage cp.async pipeline (all threads copy,
// coalesced 16-byte requests; the k=1 gather uses 4-byte requests), so the memory-level
// parallelism does not depend on occupancy.  Each lane owns an 8x8 register tile of the
// gradient; WPS warps cover the whole (M/8) x (N/8) tile grid and the CTA's SPLITS warp groups
// take disjoint pixels of every block.  One partial [M+1][N] per (CTA, split); the fixed-order
// second stage makes the result deterministic.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N_));
}

template <int MODE, int F, int K, int N>
struct GradCfg {
  static constexpr int M = MODE == 0 ? K : F * F;       // gradient rows
  static constexpr int MP = (M + 7) / 8 * 8;            // padded to whole 8-row tiles
  static constexpr int MG = MP / 8, NG = N / 8, TILES = MG * NG;
  static constexpr int WPS = (TILES + 31) / 32;          // warps per split
  static constexpr int SPLITS = (NT / 32) / WPS;         // pixel splits per CTA
  static constexpr int PB = 64;                          // pixels per pipeline stage
  static constexpr int AP = MP + 4;                      // smem pitch of an A row (bank spread)
  static constexpr int STAGE_FLOATS = PB * AP + PB * N;
  static constexpr size_t SMEM_BYTES = sizeof(float) * 2 * (size_t)STAGE_FLOATS;
  static_assert(SPLITS >= 1 && PB % SPLITS == 0, "tile grid / split shape");
};

template <int MODE, int F, int K, int N>
__global__ void __launch_bounds__(NT) grad_stream_kernel(const float* __restrict__ d,
                                                         const float* __restrict__ in,
                                                         float* __restrict__ partial, int ow,
                                                         int oh, long long P,
                                                         long long blocks_per_cta) {
  using G = GradCfg<MODE, F, K, N>;
  extern __shared__ __align__(16) float gsm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int split = warp / G::WPS;
  const int t = tid - split * G::WPS * 32;               // tile index inside the split
  const bool live = split < G::SPLITS && t < G::TILES;
  const int tt = live ? t : 0;
  const int mi = tt / G::NG, ni = tt % G::NG;
  const int iw = ow + F - 1, ih = oh + F - 1;

  const long long nblocks = (P + G::PB - 1) / G::PB;
  const long long b0 = (long long)blockIdx.x * blocks_per_cta;
  const long long b1 = min(nblocks, b0 + blocks_per_cta);

  auto stage_A = [&](int st) { return gsm + st * G::STAGE_FLOATS; };
  auto stage_B = [&](int st) { return gsm + st * G::STAGE_FLOATS + G::PB * G::AP; };

  // zero the A padding columns once (taps >= F*F); cp.async never touches them
  if (G::MP > G::M) {
    for (int i = tid; i < 2 * G::PB * (G::AP - G::M); i += NT) {
      const int st = i / (G::PB * (G::AP - G::M)), r = i % (G::PB * (G::AP - G::M));
      stage_A(st)[(r / (G::AP - G::M)) * G::AP + G::M + r % (G::AP - G::M)] = 0.f;
    }
  }

  auto load_block = [&](long long blk, int st) {
    const long long p0 = blk * G::PB;
    const int npx = (int)min((long long)G::PB, P - p0);
    float* sA = stage_A(st);
    float* sB = stage_B(st);
    // B: npx * N contiguous floats
    for (int i = tid; i < npx * (N / 4); i += NT)
      cp_async16(sB + i * 4, d + p0 * N + (long long)i * 4);
    if constexpr (MODE == 0) {
      for (int i = tid; i < npx * (K / 4); i += NT) {
        const int px = i / (K / 4), c4 = i % (K / 4);
        cp_async16(sA + px * G::AP + c4 * 4, in + (p0 + px) * K + c4 * 4);
      }
    } else {
      for (int i = tid; i < npx * F; i += NT) {
        const int px = i / F, dy = i % F;
        const long long p = p0 + px;
        long long s;
        int rem;
        if (P <= 0x7fffffffLL) {   // 32-bit divisions: the 64-bit ones cost ~100 instructions
          const unsigned pu = (unsigned)p, per = (unsigned)(ow * oh);
          const unsigned su = pu / per;
          s = su;
          rem = (int)(pu - su * per);
        } else {
          s = p / ((long long)ow * oh);
          rem = (int)(p - s * ow * oh);
        }
        const int row = rem / ow, col = rem - row * ow;
        const float* src = in + (s * ih + row + dy) * (long long)iw + col;
        float* dst = sA + px * G::AP + dy * F;
#pragma unroll
        for (int dx = 0; dx < F; dx++) cp_async4(dst + dx, src + dx);
      }
    }
    // pixels past the end of the data (last block only): zero both operands so they add nothing
    for (int i = tid; i < (G::PB - npx) * N; i += NT) sB[npx * N + i] = 0.f;
    for (int i = tid; i < (G::PB - npx) * G::AP; i += NT) sA[npx * G::AP + i] = 0.f;
  };

  float acc[8][8];
  float gb[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    gb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  }

  if (b0 < b1) load_block(b0, 0);
  cp_async_commit();
  for (long long blk = b0; blk < b1; blk++) {
    const int st = (int)((blk - b0) & 1);
    if (blk + 1 < b1) load_block(blk + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();          // this block's copies have landed (the next may be in flight)
    __syncthreads();
    if (live) {
      const float* sA = stage_A(st) + mi * 8;
      const float* sB = stage_B(st) + ni * 8;
      constexpr int PPS = G::PB / G::SPLITS;             // pixels of the block per split
#pragma unroll 4
      for (int q = 0; q < PPS; q++) {
        const int px = split * PPS + q;
        const float4 a0 = *reinterpret_cast<const float4*>(sA + px * G::AP);
        const float4 a1 = *reinterpret_cast<const float4*>(sA + px * G::AP + 4);
        const float4 c0 = *reinterpret_cast<const float4*>(sB + px * N);
        const float4 c1 = *reinterpret_cast<const float4*>(sB + px * N + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        if (mi == 0) {
#pragma unroll
          for (int j = 0; j < 8; j++) gb[j] += bv[j];
        }
      }
    }
    __syncthreads();             // everyone is done with stage st before it is refilled
  }
  cp_async_wait<0>();
  if (!live) return;
  float* dst = partial + ((long long)blockIdx.x * G::SPLITS + split) * (long long)((G::M + 1) * N);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = mi * 8 + i;
    if (m < G::M) {
      float4* o = reinterpret_cast<float4*>(dst + m * N + ni * 8);
      o[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      o[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  if (mi == 0) {
    float4* o = reinterpret_cast<float4*>(dst + G::M * N + ni * 8);
    o[0] = make_float4(gb[0], gb[1], gb[2], gb[3]);
    o[1] = make_float4(gb[4], gb[5], gb[6], gb[7]);
  }
}

template <int MODE, int F, int K, int N>
inline int launch_grad_stream(srcnn_ctx* ctx, const float* d, const float* in, int ow, int oh,
                              long long P, int* count) {
  using G = GradCfg<MODE, F, K, N>;
  SRCNN_TRY(ensure_func_setup(ctx, grad_stream_kernel<MODE, F, K, N>, G::SMEM_BYTES));
  const long long nblocks = (P + G::PB - 1) / G::PB;
  long long ctas = std::min<long long>(2LL * ctx->sm_count, nblocks);
  const long long bpc = (nblocks + ctas - 1) / ctas;
  ctas = (nblocks + bpc - 1) / bpc;
  *count = (int)(ctas * G::SPLITS);
  SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                           sizeof(float) * (size_t)(*count) * (G::M + 1) * N));
  grad_stream_kernel<MODE, F, K, N><<<(int)ctas, NT, G::SMEM_BYTES, ctx->stream>>>(
      d, in, (float*)ctx->splitk_scratch, ow, oh, P, bpc);
  return SRCNN_OK;
}

// fixed-order sum of the partial vectors, then `+=` into the accumulators
// (layout of one partial: [Mw rows][n] weights followed by [n] bias sums)
// A block owns 32 consecutive outputs; its 8 warps take the partial vectors z = w, w+8, ... (each
// with 4 independent chains), then the 8 per-warp sums are folded in warp order -- the
// association is fixed, so the result does not depend on scheduling.
constexpr int RED_OUT = 32, RED_WARPS = 8;
__global__ void __launch_bounds__(RED_OUT * RED_WARPS)
partial_reduce_kernel(const float* __restrict__ partial, float* grad_w, float* grad_b, int Mw,
                      int n, int count) {
  __shared__ float sh[RED_WARPS][RED_OUT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int id = blockIdx.x * RED_OUT + lane;
  const int total = (Mw + 1) * n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (id < total) {
    int z = warp;
    for (; z + 3 * RED_WARPS < count; z += 4 * RED_WARPS) {
      s0 += __ldg(partial + (long long)z * total + id);
      s1 += __ldg(partial + (long long)(z + RED_WARPS) * total + id);
      s2 += __ldg(partial + (long long)(z + 2 * RED_WARPS) * total + id);
      s3 += __ldg(partial + (long long)(z + 3 * RED_WARPS) * total + id);
    }
    for (; z < count; z += RED_WARPS) s0 += __ldg(partial + (long long)z * total + id);
  }
  sh[warp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (warp == 0 && id < total) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RED_WARPS; w++) s += sh[w][lane];
    if (id < Mw * n)
      grad_w[id] += s;
    else
      grad_b[id - Mw * n] += s;
  }
}

// ================================================================== dispatch ================
inline int grid_for(srcnn_ctx* ctx, long long warps_of_work) {
  long long blocks = (warps_of_work + (NT / 32) - 1) / (NT / 32);
  const long long cap = 8LL * ctx->sm_count;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// forward of an n = 1 layer; returns true when it launched
inline bool n1_forward(srcnn_ctx* ctx, const float* in, float* out, const float* W, const float* B,
                       int k, int n, int f, bool relu, int in_w, int in_h, int S) {
  if (n != 1 || f != 5 || (k != 16 && k != 32 && k != 64)) return false;
  const int ow = in_w - f + 1, oh = in_h - f + 1;
  // patch-sized samples: one sample per CTA staged in shared memory (float4 loads: a sample
  // must start on a 16-byte boundary), runs of 6 or 8 pixels, whichever wastes fewer lanes
  const size_t smp_bytes = sizeof(float) * (size_t)in_w * in_h * k;
  if (smp_bytes <= 72 * 1024 && S >= 4 && (smp_bytes % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smp_bytes));
    const int g = (int)std::min<long long>(S, (long long)per_sm * ctx->sm_count);
    const bool p6 = ((ow + 5) / 6) * 6 < ((ow + 7) / 8) * 8;
#define SRCNN_N1F_LAUNCH(CPL_, P_)                                                              \
    do {                                                                                        \
      auto kern = n1_forward_smem_kernel<5, CPL_, P_>;                                          \
      if (ensure_func_setup(ctx, kern, smp_bytes) != SRCNN_OK) return false;                    \
      kern<<<g, NT, smp_bytes, ctx->stream>>>(in, out, W, B, k, relu ? 1 : 0, in_w, in_h, ow,   \
                                              oh, S);                                           \
    } while (0)
    if (k <= 32) {
      if (p6) SRCNN_N1F_LAUNCH(1, 6); else SRCNN_N1F_LAUNCH(1, 8);
    } else {
      if (p6) SRCNN_N1F_LAUNCH(2, 6); else SRCNN_N1F_LAUNCH(2, 8);
    }
#undef SRCNN_N1F_LAUNCH
    return true;
  }
  const long long runs = (long long)S * oh * ((ow + 7) / 8);
  const int grid = grid_for(ctx, runs);
  if (k <= 32)
    n1_forward_kernel<5, 1><<<grid, NT, 0, ctx->stream>>>(in, out, W, B, k, relu ? 1 : 0, in_w,
                                                          in_h, ow, oh, S);
  else
    n1_forward_kernel<5, 2><<<grid, NT, 0, ctx->stream>>>(in, out, W, B, k, relu ? 1 : 0, in_w,
                                                          in_h, ow, oh, S);
  return true;
}

inline bool n1_deltas(srcnn_ctx* ctx, const float* dn, const float* lo, float* target,
                      const float* W, int n_curr, int f_next, int n_next, int ow, int oh, int S) {
  if (n_next != 1 || f_next != 5 || (n_curr != 16 && n_curr != 32 && n_curr != 64)) return false;
  const size_t pad_bytes = sizeof(float) * (size_t)(ow + f_next - 1) * (oh + f_next - 1);
  if (pad_bytes <= 40 * 1024) {   // patch-sized samples: padded delta map staged in smem
    const int g = (int)std::min<long long>(S, 4LL * ctx->sm_count);
    if (n_curr <= 32)
      n1_deltas_smem_kernel<5, 1><<<g, NT, pad_bytes, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
    else
      n1_deltas_smem_kernel<5, 2><<<g, NT, pad_bytes, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
    return true;
  }
  const int grid = grid_for(ctx, (long long)S * ow * oh);
  if (n_curr <= 32)
    n1_deltas_kernel<5, 1><<<grid, NT, 0, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
  else
    n1_deltas_kernel<5, 2><<<grid, NT, 0, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
  return true;
}

// deltas below an f = 1 layer; returns true when it launched
inline bool f1_deltas(srcnn_ctx* ctx, const float* dn, const float* lo, float* target,
                      const float* W, int n_curr, int f_next, int n_next, int ow, int oh, int S) {
  if (f_next != 1) return false;
  const long long P = (long long)S * ow * oh;
  const unsigned grid = (unsigned)((P + 127) / 128);
  if ((reinterpret_cast<uintptr_t>(dn) | reinterpret_cast<uintptr_t>(lo) |
       reinterpret_cast<uintptr_t>(target)) & 15u)
    return false;
  if (n_curr == 64 && n_next == 32)
    f1_deltas_kernel<64, 32><<<grid, 256, 0, ctx->stream>>>(dn, lo, target, W, P);
  else if (n_curr == 32 && n_next == 16)
    f1_deltas_kernel<32, 16><<<grid, 256, 0, ctx->stream>>>(dn, lo, target, W, P);
  else
    return false;
  return true;
}

// weight/bias gradients; returns 1 when launched, 0 when not handled, < 0 on error
inline int gradw(srcnn_ctx* ctx, const float* d, const float* in, float* grad_w, float* grad_b,
                 int n, int k, int f, int ow, int oh, int S) {
  const long long P = (long long)S * ow * oh;
  int count = 0, Mw = f * f * k;
  if (n == 1 && f == 5 && (k == 16 || k == 32 || k == 64)) {
    count = (int)std::min<long long>(S, 2LL * ctx->sm_count);
    SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                             sizeof(float) * (size_t)count * (Mw + 1)));
    float* part = (float*)ctx->splitk_scratch;
    if (k <= 32)
      n1_gradw_kernel<5, 1><<<count, NT, 0, ctx->stream>>>(d, in, part, k, ow, oh, S);
    else
      n1_gradw_kernel<5, 2><<<count, NT, 0, ctx->stream>>>(d, in, part, k, ow, oh, S);
  } else if (f == 1 && k == 64 && n == 32) {
    SRCNN_TRY((launch_grad_stream<0, 1, 64, 32>(ctx, d, in, ow, oh, P, &count)));
  } else if (f == 1 && k == 128 && n == 64) {
    SRCNN_TRY((launch_grad_stream<0, 1, 128, 64>(ctx, d, in, ow, oh, P, &count)));
  } else if (f == 1 && k == 32 && n == 16) {
    SRCNN_TRY((launch_grad_stream<0, 1, 32, 16>(ctx, d, in, ow, oh, P, &count)));
  } else if (k == 1 && f == 9 && n == 64) {
    SRCNN_TRY((launch_grad_stream<1, 9, 1, 64>(ctx, d, in, ow, oh, P, &count)));
  } else if (k == 1 && f == 9 && n == 128) {
    SRCNN_TRY((launch_grad_stream<1, 9, 1, 128>(ctx, d, in, ow, oh, P, &count)));
  } else if (k == 1 && f == 9 && n == 32) {
    SRCNN_TRY((launch_grad_stream<1, 9, 1, 32>(ctx, d, in, ow, oh, P, &count)));
  } else {
    return 0;
  }
  const int total = (Mw + 1) * n;
  partial_reduce_kernel<<<(total + RED_OUT - 1) / RED_OUT, RED_OUT * RED_WARPS, 0, ctx->stream>>>(
      (const float*)ctx->splitk_scratch, grad_w, grad_b, Mw, n, count);
  return 1;
}

// fused backward of the last layer (f = 5, n = 1); returns 1 when it launched, 0 when not
// handled.  d2 / d3 receive the deltas, grad_w3 / grad_b3 are accumulated into.
inline int bwd3_fused(srcnn_ctx* ctx, const float* gt, const float* out3, const float* out2,
                      const float* W3, float* d3, float* d2, float* grad_w, float* grad_b, int k,
                      int f, int gt_w, int gt_h, int w3, int h3, int S) {
  if (f != B3_F || (k != 16 && k != 32 && k != 64)) return 0;
  const int pw = w3 + 2 * (f - 1), ph = h3 + 2 * (f - 1);
  const size_t smem = sizeof(float) * (((size_t)pw * ph + 3) / 4 * 4 + (size_t)B3_WARPS * f * f * k);
  if (smem > 96 * 1024) return 0;   // image-sized samples: the per-kernel path handles them
  // one wave of resident CTAs, each walking S / count samples: with more CTAs than fit, the
  // leftover ones run as a second wave at a fraction of the occupancy
  // (attribute and occupancy are cached per context: they depend on the device)
  int occ = 0;
  if (k <= 32)
    SRCNN_TRY(ensure_func_setup(ctx, bwd3_fused_kernel<1>, smem, B3_NT, &occ));
  else
    SRCNN_TRY(ensure_func_setup(ctx, bwd3_fused_kernel<2>, smem, B3_NT, &occ));
  if (occ < 1) occ = 1;
  const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int count = (int)std::min<long long>(S, (long long)occ * sms);
  const int Mw = f * f * k;
  SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                           sizeof(float) * (size_t)count * (Mw + 1)));
  float* part = (float*)ctx->splitk_scratch;
  if (k <= 32) {
    bwd3_fused_kernel<1><<<count, B3_NT, smem, ctx->stream>>>(gt, out3, out2, W3, d3, d2, part, k,
                                                             gt_w, gt_h, w3, h3, S);
  } else {
    bwd3_fused_kernel<2><<<count, B3_NT, smem, ctx->stream>>>(gt, out3, out2, W3, d3, d2, part, k,
                                                             gt_w, gt_h, w3, h3, S);
  }
  partial_reduce_kernel<<<(Mw + 1 + RED_OUT - 1) / RED_OUT, RED_OUT * RED_WARPS, 0, ctx->stream>>>(
      part, grad_w, grad_b, Mw, 1, count);
  return 1;
}

}  // namespace train
}  // namespace srcnn
