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

// ------------------------------------------------------------------ gW of an f=1 layer --------
// gW[c][n] += sum_p in[p][c] * d[p][n];  gB[n] += sum_p d[p][n]     (p over all pixels)
// reference: src/kernel/backpropagate.cl:56-114 with f_spatial_size = 1.
// A "slice" is WPS warps covering the (K/8) x (N/8) grid of 8x8 register tiles; slices own
// disjoint, contiguous pixel ranges.  partial[slice][K+1][N].
template <int K, int N>
__global__ void __launch_bounds__(NT) dense_gradw_kernel(const float* __restrict__ d,
                                                         const float* __restrict__ in,
                                                         float* __restrict__ partial,
                                                         long long P, long long pix_per_slice) {
  constexpr int NG = N / 8, TILES = (K / 8) * NG, WPS = TILES / 32;
  static_assert(TILES % 32 == 0 && (NT / 32) % WPS == 0, "tile grid must fill whole warps");
  constexpr int SLICES_PER_CTA = (NT / 32) / WPS;
  const int slice_in_cta = (threadIdx.x >> 5) / WPS;
  const int t = threadIdx.x - slice_in_cta * WPS * 32;   // tile index inside the slice
  const int mi = t / NG, ni = t % NG;
  const long long slice = (long long)blockIdx.x * SLICES_PER_CTA + slice_in_cta;
  const long long p0 = slice * pix_per_slice;
  const long long p1 = min(P, p0 + pix_per_slice);
  float acc[8][8];
  float gb[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    gb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  }
  for (long long p = p0; p < p1; p++) {
    const float4* ap = reinterpret_cast<const float4*>(in + p * K + mi * 8);
    const float4* bp = reinterpret_cast<const float4*>(d + p * N + ni * 8);
    const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    if (mi == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++) gb[j] += bv[j];
    }
  }
  float* dst = partial + slice * (long long)((K + 1) * N);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float4* o = reinterpret_cast<float4*>(dst + (mi * 8 + i) * N + ni * 8);
    o[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    o[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  if (mi == 0) {
    float4* o = reinterpret_cast<float4*>(dst + K * N + ni * 8);
    o[0] = make_float4(gb[0], gb[1], gb[2], gb[3]);
    o[1] = make_float4(gb[4], gb[5], gb[6], gb[7]);
  }
}

// ------------------------------------------------------------------ gW of a k=1 layer ---------
// gW[dy][dx][n] += sum_{s,row,col} in[s][row+dy][col+dx] * d[s][row][col][n];  gB[n] += sum d
// reference: src/kernel/backpropagate.cl:56-114 with n_prev_filter_cnt = 1.
// Register tiles: 8 taps x 8 channels; MG = ceil(f*f/8) tap groups (taps >= f*f read as 0).
template <int F, int N>
__global__ void __launch_bounds__(NT) k1_gradw_kernel(const float* __restrict__ d,
                                                      const float* __restrict__ in,
                                                      float* __restrict__ partial, int ow, int oh,
                                                      int S, long long pix_per_slice) {
  constexpr int FF = F * F, MG = (FF + 7) / 8, NG = N / 8, TILES = MG * NG;
  constexpr int WPS = (TILES + 31) / 32;
  constexpr int SLICES_PER_CTA = (NT / 32) / WPS;
  static_assert(SLICES_PER_CTA >= 1, "tile grid larger than a CTA");
  const int iw = ow + F - 1, ih = oh + F - 1;
  const int slice_in_cta = (threadIdx.x >> 5) / WPS;
  if (slice_in_cta >= SLICES_PER_CTA) return;
  const int t = threadIdx.x - slice_in_cta * WPS * 32;
  const bool live = t < TILES;
  const int tt = live ? t : 0;
  const int mi = tt / NG, ni = tt % NG;
  int toff[8];
  float tmask[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int tap = mi * 8 + i;
    const int tc = tap < FF ? tap : 0;
    toff[i] = (tc / F) * iw + (tc % F);
    tmask[i] = tap < FF ? 1.f : 0.f;
  }
  const long long P = (long long)S * ow * oh;
  const long long slice = (long long)blockIdx.x * SLICES_PER_CTA + slice_in_cta;
  const long long p0 = slice * pix_per_slice;
  const long long p1 = min(P, p0 + pix_per_slice);
  float acc[8][8];
  float gb[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    gb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  }
  // walk (s,row,col) incrementally instead of dividing per pixel
  long long s = p0 / ((long long)ow * oh);
  int rem = (int)(p0 - s * ow * oh);
  int row = rem / ow, col = rem - row * ow;
  for (long long p = p0; p < p1; p++) {
    const float* px = in + (s * ih + row) * (long long)iw + col;
    const float4* bp = reinterpret_cast<const float4*>(d + p * N + ni * 8);
    const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float av[8];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = __ldg(px + toff[i]) * tmask[i];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    if (mi == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++) gb[j] += bv[j];
    }
    if (++col == ow) {
      col = 0;
      if (++row == oh) {
        row = 0;
        ++s;
      }
    }
  }
  if (!live) return;
  float* dst = partial + slice * (long long)((FF + 1) * N);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int tap = mi * 8 + i;
    if (tap < FF) {
      float4* o = reinterpret_cast<float4*>(dst + tap * N + ni * 8);
      o[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      o[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  if (mi == 0) {
    float4* o = reinterpret_cast<float4*>(dst + FF * N + ni * 8);
    o[0] = make_float4(gb[0], gb[1], gb[2], gb[3]);
    o[1] = make_float4(gb[4], gb[5], gb[6], gb[7]);
  }
}

// fixed-order sum of the partial vectors, then `+=` into the accumulators
// (layout of one partial: [Mw rows][n] weights followed by [n] bias sums)
__global__ void partial_reduce_kernel(const float* __restrict__ partial, float* grad_w,
                                      float* grad_b, int Mw, int n, int count) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = (Mw + 1) * n;
  if (id >= total) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // 4 independent chains, fixed association
  int z = 0;
  for (; z + 3 < count; z += 4) {
    s0 += partial[(long long)z * total + id];
    s1 += partial[(long long)(z + 1) * total + id];
    s2 += partial[(long long)(z + 2) * total + id];
    s3 += partial[(long long)(z + 3) * total + id];
  }
  for (; z < count; z++) s0 += partial[(long long)z * total + id];
  const float s = (s0 + s1) + (s2 + s3);
  if (id < Mw * n)
    grad_w[id] += s;
  else
    grad_b[id - Mw * n] += s;
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
  const int grid = grid_for(ctx, (long long)S * ow * oh);
  if (n_curr <= 32)
    n1_deltas_kernel<5, 1><<<grid, NT, 0, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
  else
    n1_deltas_kernel<5, 2><<<grid, NT, 0, ctx->stream>>>(dn, lo, target, W, n_curr, ow, oh, S);
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
  } else if (f == 1 && ((k == 64 && n == 32) || (k == 128 && n == 64))) {
    const int wps = (k / 8) * (n / 8) / 32, spc = (NT / 32) / wps;
    long long slices = std::min<long long>((long long)2 * ctx->sm_count * spc, (P + 63) / 64);
    slices = (slices + spc - 1) / spc * spc;
    const long long pps = (P + slices - 1) / slices;
    count = (int)slices;
    SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                             sizeof(float) * (size_t)count * (Mw + 1) * n));
    float* part = (float*)ctx->splitk_scratch;
    if (k == 64)
      dense_gradw_kernel<64, 32><<<(int)(slices / spc), NT, 0, ctx->stream>>>(d, in, part, P, pps);
    else
      dense_gradw_kernel<128, 64><<<(int)(slices / spc), NT, 0, ctx->stream>>>(d, in, part, P, pps);
  } else if (k == 1 && f == 9 && (n == 64 || n == 128 || n == 32)) {
    const int tiles = 11 * (n / 8), wps = (tiles + 31) / 32, spc = (NT / 32) / wps;
    long long slices = std::min<long long>((long long)2 * ctx->sm_count * spc, (P + 63) / 64);
    slices = (slices + spc - 1) / spc * spc;
    const long long pps = (P + slices - 1) / slices;
    count = (int)slices;
    SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                             sizeof(float) * (size_t)count * (Mw + 1) * n));
    float* part = (float*)ctx->splitk_scratch;
    const int grid = (int)(slices / spc);
    if (n == 64)
      k1_gradw_kernel<9, 64><<<grid, NT, 0, ctx->stream>>>(d, in, part, ow, oh, S, pps);
    else if (n == 128)
      k1_gradw_kernel<9, 128><<<grid, NT, 0, ctx->stream>>>(d, in, part, ow, oh, S, pps);
    else
      k1_gradw_kernel<9, 32><<<grid, NT, 0, ctx->stream>>>(d, in, part, ow, oh, S, pps);
  } else {
    return 0;
  }
  const int total = (Mw + 1) * n;
  partial_reduce_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(
      (const float*)ctx->splitk_scratch, grad_w, grad_b, Mw, n, count);
  return 1;
}

}  // namespace train
}  // namespace srcnn
