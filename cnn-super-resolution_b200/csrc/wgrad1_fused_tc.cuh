// Layer-1 deltas FUSED into the layer-1 weight / bias gradient (9-1-5, n1 = 64, n2 = 32):
//     d1[p][c]        = [out1[p][c] > 0] * sum_k W2[c][k] * d2[p][k]            (never stored)
//     gW1[dy][dx][c] += sum_p in[p + (dy,dx)] * d1[p][c],     gB1[c] += sum_p d1[p][c]
// reference: src/kernel/layer_deltas.cl:42-127 (f_next = 1) followed by
// src/kernel/backpropagate.cl:56-114; ConfigBasedDataPipeline.cpp:258-320 runs them as two
// launches with the layer-1 deltas (n1 floats per pixel, 328 MB per 2 048 patches) written and
// read back in between.
//
// wgrad1_tc_kernel (wgrad_tc.cuh) wants its B operand TRANSPOSED: rows = channels, K = the 64
// pixels of a tile.  The delta GEMM can deliver exactly that: with A = W2 (rows = channels,
// K-major as stored) and B = the d2 tile (rows = pixels, K-major as stored) the accumulator is
//     DT[c][p] = sum_k W2[c][k] * d2[p][k]          TMEM lane = channel, column = pixel
// so the PB warps of the gradient kernel read their operand row out of tensor memory instead of
// global memory, mask it with out1 (loaded with the access pattern they used for d1), split it
// and store it as before.  3xTF32 into ONE accumulator (hi.hi + hi.lo + lo.hi).  W2 is stored
// twice along M (rows 64..127 repeat rows 0..63): lanes 64..127 then hold a copy of the tile,
// and the four PB warps -- one per TMEM lane quarter -- split the pixels of a tile in halves
// without exchanging anything.
//
//   LD (2 warps)  d2 tile -> TF32 hi / lo -> smem                               -> d2_full
//   I2 (1 warp)   DT[i&1] = [W2;W2] x d2^T, 12 MMAs                             -> dt_full
//   PB (4 warps)  DT -> mask(out1) -> bias sums -> split -> B tile              -> full, dt_free
//   PA (8 warps)  im2col gather -> split -> A tiles                             -> full
//   I  (1 warp)   8 K-steps x 2 MMAs into the resident gradient accumulators    -> empty
#pragma once
#include <cuda_runtime.h>

#include "context.cuh"
#include "fused_forward_pl.cuh"
#include "tc_common.cuh"
#include "wgrad_tc.cuh"

namespace srcnn {
namespace wgf {

struct Cfg {
  static constexpr int F = 9, T = F * F, TP = 88, N = 64, K2 = 32;
  static constexpr int PX = 64;
  static constexpr int N_PA = 8, W_PB = N_PA, W_I = W_PB + 4, W_LD = W_I + 1, W_I2 = W_LD + 2;
  static constexpr int NT = (W_I2 + 1) * 32;
  static constexpr int PA_ITEMS = (T * (PX / 4) + N_PA * 32 - 1) / (N_PA * 32);
  static constexpr int SBO = 128 * (PX / 4);     // tiles with K = 64 pixels
  static constexpr int SBO2 = 128 * (K2 / 4);    // operands of the delta GEMM, K = 32
  // shared memory (floats).  An M = 128 MMA reads rows TP..127 of an A tile: they fall into
  // the arrays behind it (A_lo, B), inside the allocation; nobody reads those accumulator rows
  static constexpr int A_FLOATS = TP * PX, B_FLOATS = 2 * N * PX;
  static constexpr int STAGE = 2 * A_FLOATS + B_FLOATS;
  static constexpr int oBase = 2 * STAGE;                // per-pixel input offsets [2][PX]
  static constexpr int oW2 = oBase + 2 * PX;             // [W2hi;W2hi], [W2lo;W2lo]: 2 x [128][K2]
  static constexpr int W2_FLOATS = 128 * K2;
  static constexpr int oD2 = oW2 + 2 * W2_FLOATS;        // d2 stages: 2 x (hi [PX][K2], lo)
  static constexpr int D2_FLOATS = PX * K2;
  static constexpr int TOTAL = oD2 + 2 * 2 * D2_FLOATS;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)TOTAL;
  // tensor memory: gradient accumulators as in wgrad1_tc_kernel, then DT[2] (64 columns each)
  static constexpr uint32_t cDhi = 0, cDlo = 128, cDT = 192, TMEM_COLS = 512;
};
static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(Cfg::NT, 1)
    wgrad1_fused_tc_kernel(const float* __restrict__ d2, const float* __restrict__ out1,
                           const float* __restrict__ W2, const float* __restrict__ in,
                           float* __restrict__ partial, int ow, int oh, long long P,
                           int n_tiles_total) {
  using C = Cfg;
  using namespace tc;
  using fused_pl::elect_one;
  using fused_pl::tmem_ld16_nowait;
  using fused_pl::tmem_ld_wait;
  using fused_ws::mbar_arrive;
  using fused_ws::named_bar_sync;
  extern __shared__ __align__(128) float wg_smem[];
  int* sBase = reinterpret_cast<int*>(wg_smem + C::oBase);
  float* sW2 = wg_smem + C::oW2;
  float* sD2 = wg_smem + C::oD2;
  __shared__ __align__(8) uint64_t full[2], empty[2], d2_full[2], dt_full[2], dt_free[2], done;
  __shared__ float gb_part[2][C::N];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int iw = ow + C::F - 1, ih = oh + C::F - 1;

  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      mbar_init(&full[i], C::N_PA * 32 + 128);
      mbar_init(&empty[i], 1);
      mbar_init(&d2_full[i], 64);
      mbar_init(&dt_full[i], 1);
      mbar_init(&dt_free[i], 128);
    }
    mbar_init(&done, 1);
  }
  // the zero rows of the im2col tiles (taps 81..87) are written once
  for (int i = tid; i < 2 * 2 * (C::TP - C::T) * C::PX; i += C::NT) {
    const int tile = i / ((C::TP - C::T) * C::PX), r = i % ((C::TP - C::T) * C::PX);
    const int t = C::T + r / C::PX, k = r % C::PX;
    wg_smem[(tile >> 1) * C::STAGE + (tile & 1) * C::A_FLOATS + kmajor_offset(t, k, C::PX)] = 0.f;
  }
  // A operand of the delta GEMM: row r = channel r & 63
  for (int i = tid; i < 128 * C::K2; i += C::NT) {
    const int r = i / C::K2, k = i % C::K2;
    float hi, lo;
    split_tf32(__ldg(W2 + (r & (C::N - 1)) * C::K2 + k), hi, lo);
    sW2[kmajor_offset(r, k, C::K2)] = hi;
    sW2[C::W2_FLOATS + kmajor_offset(r, k, C::K2)] = lo;
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = (n_tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const long long per = (long long)ow * oh;

  if (warp < C::W_PB) {
    // ============================ PA: im2col tiles (as wgrad1_tc_kernel) ===================
    for (int i = 0; i < my_tiles; i++) {
      const long long p0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * C::PX;
      if (i >= 2) mbar_wait(&empty[i & 1], (uint32_t)(((i - 2) >> 1) & 1));
      int* base = sBase + (i & 1) * C::PX;
      if (tid < C::PX) {
        const long long p = p0 + tid;
        int b = -1;
        if (p < P) {
          const long long s = p / per;
          const int rem = (int)(p - s * per), row = rem / ow, col = rem - row * ow;
          b = (int)((s * ih + row) * iw + col);
        }
        base[tid] = b;
      }
      named_bar_sync(1, C::N_PA * 32);
      float* sAh = wg_smem + (i & 1) * C::STAGE;
      float* sAl = sAh + C::A_FLOATS;
      float v[C::PA_ITEMS][4];
#pragma unroll
      for (int u = 0; u < C::PA_ITEMS; u++) {
        const int it = tid + C::N_PA * 32 * u;
        const int t = it % C::T, q = it / C::T;
        const int toff = (t / C::F) * iw + (t % C::F);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int b = it < C::T * (C::PX / 4) ? base[4 * q + j] : -1;
          v[u][j] = b >= 0 ? __ldg(in + b + toff) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < C::PA_ITEMS; u++) {
        const int it = tid + C::N_PA * 32 * u;
        if (it < C::T * (C::PX / 4)) {
          const int t = it % C::T, q = it / C::T;
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; j++) split_tf32(v[u][j], hi[j], lo[j]);
          const int off = kmajor_offset(t, 4 * q, C::PX);
          *reinterpret_cast<float4*>(sAh + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(sAl + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[i & 1]);
    }
  } else if (warp < C::W_I) {
    // ============================ PB: DT -> mask -> B tile, bias sums ======================
    // warp pw = TMEM lane quarter pw: channel 32*(pw&1) + lane, pixels 32*(pw>>1) .. +31
    const int pw = warp - C::W_PB;
    const int c = (pw & 1) * 32 + lane;
    const int px0 = (pw >> 1) * 32;
    const uint32_t lane_base = (uint32_t)(pw * 32) << 16;
    float gb = 0.f;
    for (int i = 0; i < my_tiles; i++) {
      const long long p0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * C::PX + px0;
      // the mask operand does not depend on the MMA: all loads in flight while waiting
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; j++) o[j] = p0 + j < P ? __ldg(out1 + (p0 + j) * C::N + c) : 0.f;
      mbar_wait(&dt_full[i & 1], (uint32_t)((i >> 1) & 1));
      tcgen05_fence_after();
      float v[32];
      const uint32_t dt = tmem + lane_base + C::cDT + 64u * (uint32_t)(i & 1) + (uint32_t)px0;
      tmem_ld16_nowait(dt, v);
      tmem_ld16_nowait(dt + 16, v + 16);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(&dt_free[i & 1]);
      if (i >= 2) mbar_wait(&empty[i & 1], (uint32_t)(((i - 2) >> 1) & 1));
      float* sB = wg_smem + (i & 1) * C::STAGE + 2 * C::A_FLOATS;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        float d[4], hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; j++) d[j] = o[4 * u + j] > 0.f ? v[4 * u + j] : 0.f;
        gb += (d[0] + d[1]) + (d[2] + d[3]);
#pragma unroll
        for (int j = 0; j < 4; j++) split_tf32(d[j], hi[j], lo[j]);
        *reinterpret_cast<float4*>(sB + kmajor_offset(c, px0 + 4 * u, C::PX)) =
            make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sB + kmajor_offset(C::N + c, px0 + 4 * u, C::PX)) =
            make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      mbar_arrive(&full[i & 1]);
    }
    gb_part[pw >> 1][c] = gb;
  } else if (warp == C::W_I) {
    // ============================ I: gradient MMA issuer ===================================
    const uint32_t idesc_hi = make_idesc_tf32(128, 2 * C::N);   // A_hi x [B_hi; B_lo]
    const uint32_t idesc_lo = make_idesc_tf32(128, C::N);       // A_lo x B_hi
    for (int i = 0; i < my_tiles; i++) {
      mbar_wait(&full[i & 1], (uint32_t)((i >> 1) & 1));
      tcgen05_fence_after();
      const float* st = wg_smem + (i & 1) * C::STAGE;
      const uint64_t ah = make_desc_kmajor(st, 0, 128, C::SBO);
      const uint64_t al = make_desc_kmajor(st + C::A_FLOATS, 0, 128, C::SBO);
      const uint64_t bd = make_desc_kmajor(st + 2 * C::A_FLOATS, 0, 128, C::SBO);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::PX / 8; ks++) {
          mma_tf32(tmem + C::cDhi, ah + 16 * ks, bd + 16 * ks, idesc_hi, (i | ks) > 0);
          mma_tf32(tmem + C::cDlo, al + 16 * ks, bd + 16 * ks, idesc_lo, (i | ks) > 0);
        }
        mma_commit(&empty[i & 1]);
        if (i == my_tiles - 1) mma_commit(&done);
      }
      __syncwarp();
    }
  } else if (warp < C::W_I2) {
    // ============================ LD: d2 tile -> TF32 hi / lo (rows = pixels) ==============
    const int t = tid - C::W_LD * 32;   // pixel of the tile
    for (int i = 0; i < my_tiles; i++) {
      const long long p = ((long long)blockIdx.x + (long long)i * gridDim.x) * C::PX + t;
      float4 v[C::K2 / 4];
#pragma unroll
      for (int q = 0; q < C::K2 / 4; q++)
        v[q] = p < P ? __ldg(reinterpret_cast<const float4*>(d2 + p * C::K2) + q)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      // the stage is free once the delta GEMM of tile i-2 has completed
      if (i >= 2) mbar_wait(&dt_full[i & 1], (uint32_t)(((i - 2) >> 1) & 1));
      float* sh = sD2 + (i & 1) * 2 * C::D2_FLOATS;
      float* sl = sh + C::D2_FLOATS;
#pragma unroll
      for (int q = 0; q < C::K2 / 4; q++) {
        float hi[4], lo[4];
        split_tf32(v[q].x, hi[0], lo[0]);
        split_tf32(v[q].y, hi[1], lo[1]);
        split_tf32(v[q].z, hi[2], lo[2]);
        split_tf32(v[q].w, hi[3], lo[3]);
        const int off = kmajor_offset(t, 4 * q, C::K2);
        *reinterpret_cast<float4*>(sh + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sl + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      mbar_arrive(&d2_full[i & 1]);
    }
  } else {
    // ============================ I2: delta GEMM issuer ====================================
    const uint32_t idesc = make_idesc_tf32(128, C::PX);
    const uint64_t wh = make_desc_kmajor(sW2, 0, 128, C::SBO2);
    const uint64_t wl = make_desc_kmajor(sW2 + C::W2_FLOATS, 0, 128, C::SBO2);
    for (int i = 0; i < my_tiles; i++) {
      mbar_wait(&d2_full[i & 1], (uint32_t)((i >> 1) & 1));
      // DT[i&1] is free once the PB warps have read tile i-2 out of it
      if (i >= 2) mbar_wait(&dt_free[i & 1], (uint32_t)(((i - 2) >> 1) & 1));
      tcgen05_fence_after();
      const float* st = sD2 + (i & 1) * 2 * C::D2_FLOATS;
      const uint64_t dh = make_desc_kmajor(st, 0, 128, C::SBO2);
      const uint64_t dl = make_desc_kmajor(st + C::D2_FLOATS, 0, 128, C::SBO2);
      const uint32_t dt = tmem + C::cDT + 64u * (uint32_t)(i & 1);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::K2 / 8; ks++) {
          mma_tf32(dt, wh + 16 * ks, dh + 16 * ks, idesc, ks > 0);
          mma_tf32(dt, wh + 16 * ks, dl + 16 * ks, idesc, 1);
          mma_tf32(dt, wl + 16 * ks, dh + 16 * ks, idesc, 1);
        }
        mma_commit(&dt_full[i & 1]);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: one partial [T*N + N] per CTA; thread = filter tap (TMEM lane)
  __syncthreads();   // gb_part complete, all producers done
  float* dst = partial + (long long)blockIdx.x * (C::T * C::N + C::N);
  if (warp < 3) {
    if (my_tiles > 0) {
      mbar_wait(&done, 0);
      tcgen05_fence_after();
    }
    const int t = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int g = 0; g < C::N / 16; g++) {
      float a[16], b[16], l[16];
      if (my_tiles > 0) {
        tmem_ld16_nowait(tmem + lane_base + C::cDhi + g * 16, a);
        tmem_ld16_nowait(tmem + lane_base + C::cDhi + C::N + g * 16, b);
        tmem_ld16_nowait(tmem + lane_base + C::cDlo + g * 16, l);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = b[j] = l[j] = 0.f;
      }
      if (t < C::T) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + t * C::N + g * 16 + j) =
              make_float4((a[j] + b[j]) + l[j], (a[j + 1] + b[j + 1]) + l[j + 1],
                          (a[j + 2] + b[j + 2]) + l[j + 2], (a[j + 3] + b[j + 3]) + l[j + 3]);
      }
    }
  }
  if (tid < C::N) dst[C::T * C::N + tid] = gb_part[0][tid] + gb_part[1][tid];
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

// Layer-1 deltas + layer-1 gradients of an f2 = 1 network in one launch: per-CTA partials go to
// ctx->splitk_scratch (`count` of them) for partial_reduce_kernel.  Returns 1 when it launched,
// 0 when the shape has no instantiation.
inline int wgrad1_fused_tc(srcnn_ctx* ctx, const float* d2, const float* out1, const float* W2,
                           const float* in, int n1, int n2, int f1, int f2, int ow, int oh, int S,
                           int* count) {
  if (f1 != Cfg::F || f2 != 1 || n1 != Cfg::N || n2 != Cfg::K2) return 0;
  if ((reinterpret_cast<uintptr_t>(d2) & 15u) != 0) return 0;
  const long long P = (long long)S * ow * oh;
  const long long in_elems = (long long)S * (ow + f1 - 1) * (oh + f1 - 1);
  if (in_elems > 0x7fffffffLL) return 0;   // 32-bit input offsets
  const long long tiles = (P + Cfg::PX - 1) / Cfg::PX;
  if (tiles > 0x7fffffffLL) return 0;
  static bool configured = false;
  if (!configured) {
    SRCNN_CUDA(cudaFuncSetAttribute(wgrad1_fused_tc_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM_BYTES));
    configured = true;
  }
  const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int grid = (int)(tiles < sms ? tiles : sms);
  *count = grid;
  SRCNN_TRY(ensure_scratch(ctx, &ctx->splitk_scratch, &ctx->splitk_bytes,
                           sizeof(float) * (size_t)grid * (Cfg::T * Cfg::N + Cfg::N)));
  wgrad1_fused_tc_kernel<<<grid, Cfg::NT, Cfg::SMEM_BYTES, ctx->stream>>>(
      d2, out1, W2, in, (float*)ctx->splitk_scratch, ow, oh, P, (int)tiles);
  return 1;
}

}  // namespace wgf
}  // namespace srcnn
