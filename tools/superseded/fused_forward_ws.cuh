// Fused SRCNN inference, warp-specialised tensor-core version (9-1-5, n1=64, n2=32).
// Same mathematics as fused_forward_tc.cuh (all three layers as 3xTF32 tcgen05 GEMMs on
// 128-pixel tiles, accumulators and the A2/A3 operands in TMEM, layer 3 as a tap GEMM + 25-term
// gather) but the stages no longer run in lockstep: four warp roles work on different tiles at
// the same time and hand buffers to each other through mbarriers.
//
//   role            warps   per tile b
//   IM   (im2col)    8..15  wait MMA-1(b-1) done (A1 free) -> build A1(b) hi/lo -> arrive a1_full
//   MMA  (issuer)    16     wait a2_full(b) -> MMA-2(b); wait a1_full(b+1) -> MMA-1(b+1);
//                           wait a3_full(b) -> MMA-3(b)        (tcgen05.commit -> bar1/2/3)
//   E1   (epilogue1) 0..3   wait MMA-1(b) -> D1 -> relu/split -> A2 (TMEM) -> arrive a2_full
//   E2   (epilogue2) 4..7   wait MMA-2(b) -> D2 -> relu/split -> A3 (TMEM) -> arrive a3_full
//   E3   (epi3+gath) 17..20 wait MMA-3(b) -> Q rows -> ring (arrive d3_free); gather out3 rows
//
// Buffer hand-over needs no extra "free" barriers: E1 arrives on a2_full(b) only after it has
// read D1(b), and the issuer waits for a2_full(b) before MMA-1(b+1), so D1 is free by then; the
// same argument covers D2/A3 (a3_full) and D3 (the epilogue-3 of tile b-1 precedes a3_full(b)).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "context.cuh"
#include "fused_forward.cuh"
#include "tc_common.cuh"

namespace srcnn {
namespace fused_ws {

struct Cfg {
  static constexpr int N1 = 64, N2 = 32, F1 = 9, F3 = 5;
  static constexpr int NT = 21 * 32;
  static constexpr int OW2 = 64, RB = 2;
  static constexpr int OW3 = OW2 - (F3 - 1);
  static constexpr int IW = OW2 + F1 - 1, IWP = IW + 4;
  static constexpr int IR = RB + F1 - 1 + RB;   // input ring: one tile of slack
  static constexpr int RING = 8;                // Q ring rows (>= RB + F3 - 1 + RB)
  static constexpr int RPC = 128;
  static constexpr int M = OW2 * RB;            // 128 pixels per MMA tile
  static constexpr int K1 = 88, K2 = N1, NT3 = 32, QP = F3 * F3;
  static constexpr int oA1h = 0;
  static constexpr int oA1l = oA1h + M * K1;
  static constexpr int oW1h = oA1l + M * K1;
  static constexpr int oW1l = oW1h + N1 * K1;
  static constexpr int oW2h = oW1l + N1 * K1;
  static constexpr int oW2l = oW2h + N2 * K2;
  static constexpr int oW3h = oW2l + N2 * K2;
  static constexpr int oW3l = oW3h + NT3 * N2;
  static constexpr int oB1 = oW3l + NT3 * N2;
  static constexpr int oB2 = oB1 + N1;
  static constexpr int oIn = oB2 + N2;
  static constexpr int oQ = oIn + IR * IWP;
  static constexpr int TOTAL = oQ + RING * OW2 * QP;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)TOTAL;
  // tensor memory columns; D1 is double buffered so MMA-1(b+1) can run under epilogue-1(b)
  static constexpr uint32_t cD1 = 0 /* + 64 * (b & 1) */, cD2 = 128, cD3 = 160, cA2h = 192,
                            cA2l = 256, cA3h = 320, cA3l = 352;
  static constexpr uint32_t TMEM_COLS = 512;
  // named barriers (0 is __syncthreads)
  static constexpr int BAR_IM = 1, BAR_E23 = 2;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(Cfg::NT, 1) forward_fused_ws_kernel(fused::Args a) {
  using C = Cfg;
  using namespace tc;
  extern __shared__ __align__(128) float smem[];
  float* sA1h = smem + C::oA1h;
  float* sA1l = smem + C::oA1l;
  float* sW1h = smem + C::oW1h;
  float* sW1l = smem + C::oW1l;
  float* sW2h = smem + C::oW2h;
  float* sW2l = smem + C::oW2l;
  float* sW3h = smem + C::oW3h;
  float* sW3l = smem + C::oW3l;
  float* sB1 = smem + C::oB1;
  float* sB2 = smem + C::oB2;
  float* sIn = smem + C::oIn;
  float* sQ = smem + C::oQ;
  // bar1[i]: MMA-1 into D1 buffer i done (one barrier per buffer: a waiter is never more than
  // one phase behind); bar2/bar3: MMA-2 / MMA-3 done; aN_full: operand of layer N ready
  __shared__ __align__(8) uint64_t bar1[2], bar2, bar3, a1_full, a2_full, a3_full, d3_free;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int X0 = blockIdx.x * C::OW3;
  const int R0 = blockIdx.y * C::RPC;
  const float* img = a.in + (size_t)blockIdx.z * a.w * a.h;
  float* dst = a.out + (size_t)blockIdx.z * a.w3 * a.h3;

  // ---- stage parameters (all threads): B operands split into TF32 hi/lo, [n][k] canonical ----
  for (int i = tid; i < C::N1 * C::K1; i += C::NT) {
    const int n = i / C::K1, k = i % C::K1;
    float hi, lo;
    split_tf32(k < C::F1 * C::F1 ? __ldg(a.pw1 + k * C::N1 + n) : 0.f, hi, lo);
    sW1h[kmajor_offset(n, k, C::K1)] = hi;
    sW1l[kmajor_offset(n, k, C::K1)] = lo;
  }
  for (int i = tid; i < C::N2 * C::K2; i += C::NT) {
    const int n = i / C::K2, k = i % C::K2;
    float hi, lo;
    split_tf32(__ldg(a.pw2 + k * C::N2 + n), hi, lo);
    sW2h[kmajor_offset(n, k, C::K2)] = hi;
    sW2l[kmajor_offset(n, k, C::K2)] = lo;
  }
  for (int i = tid; i < C::NT3 * C::N2; i += C::NT) {
    const int n = i / C::N2, k = i % C::N2;   // n = tap dy*5+dx, k = channel
    float hi, lo;
    split_tf32(n < C::QP ? __ldg(a.pw3 + n * C::N2 + k) : 0.f, hi, lo);
    sW3h[kmajor_offset(n, k, C::N2)] = hi;
    sW3l[kmajor_offset(n, k, C::N2)] = lo;
  }
  for (int i = tid; i < C::N1; i += C::NT) sB1[i] = __ldg(a.pb1 + i);
  for (int i = tid; i < C::N2; i += C::NT) sB2[i] = __ldg(a.pb2 + i);
  const float b3 = __ldg(a.pb3);

  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  if (tid == 0) {
    if (smem_u32(smem) & 127u) __trap();
    mbar_init(&bar1[0], 1);
    mbar_init(&bar1[1], 1);
    mbar_init(&bar2, 1);
    mbar_init(&bar3, 1);
    mbar_init(&a1_full, 256);
    mbar_init(&a2_full, 128);
    mbar_init(&a3_full, 128);
    mbar_init(&d3_free, 128);
  }
  fence_proxy_async();   // the weight operands are read by the tensor core (async proxy)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;

  const int rows_here = min(C::RPC, a.h3 - R0);
  const int n_tiles = (rows_here + (C::F3 - 1) + C::RB - 1) / C::RB;

  if (warp >= 8 && warp < 16) {
    // ============================ IM: im2col producers (256 threads) =======================
    const int t = tid - 256;
    const int im_m = t & (C::M - 1), im_half = t >> 7;
    const int im_r = im_m / C::OW2, im_x = im_m % C::OW2;
    auto load_rows = [&](int first_rel_row, int count) {
      for (int i = t; i < count * C::IW; i += 256) {
        const int rr = first_rel_row + i / C::IW, xx = i % C::IW;
        const int gy = R0 + rr, gx = X0 + xx;
        sIn[(rr % C::IR) * C::IWP + xx] =
            (gy < a.h && gx < a.w) ? __ldg(img + (size_t)gy * a.w + gx) : 0.f;
      }
    };
    auto im2col = [&](int b) {
      const int base_slot = (b * C::RB + im_r) % C::IR;
      const float* rowp[C::F1];
#pragma unroll
      for (int dy = 0; dy < C::F1; dy++) {
        int slot = base_slot + dy;
        slot = slot >= C::IR ? slot - C::IR : slot;
        rowp[dy] = sIn + slot * C::IWP + im_x;
      }
      auto half = [&](auto tag) {
        constexpr int H = decltype(tag)::value;
#pragma unroll
        for (int cc = 0; cc < 11; cc++) {
          const int c = H * 11 + cc;   // 16-byte K chunk
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int k = c * 4 + j;
            float v = 0.f;
            if (k < C::F1 * C::F1) {
              const int dy = k / C::F1, dx = k - dy * C::F1;
              v = rowp[dy][dx];
            }
            split_tf32(v, hi[j], lo[j]);
          }
          const int off = kmajor_offset(im_m, c * 4, C::K1);
          *reinterpret_cast<float4*>(sA1h + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(sA1l + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      };
      if (im_half == 0)
        half(std::integral_constant<int, 0>{});
      else
        half(std::integral_constant<int, 1>{});
    };
    load_rows(0, C::F1 - 1 + C::RB);                         // rows of tile 0
    if (n_tiles > 1) load_rows(C::RB + C::F1 - 1, C::RB);    // rows of tile 1
    named_bar_sync(C::BAR_IM, 256);
    for (int b = 0; b < n_tiles; b++) {
      // rows of tile b+2 are fetched into a register now and stored after the im2col
      float pre = 0.f;
      int pre_idx = -1;
      if (b >= 1 && b + 1 < n_tiles && t < C::RB * C::IW) {
        const int rr = (b + 1) * C::RB + C::F1 - 1 + t / C::IW, xx = t % C::IW;
        const int gy = R0 + rr, gx = X0 + xx;
        pre = (gy < a.h && gx < a.w) ? __ldg(img + (size_t)gy * a.w + gx) : 0.f;
        pre_idx = (rr % C::IR) * C::IWP + xx;
      }
      if (b > 0) mbar_wait(&bar1[(b - 1) & 1], (uint32_t)(((b - 1) >> 1) & 1));   // MMA-1(b-1) done: A1 free
#ifndef EXP_SKIP_IM2COL
      im2col(b);
#endif
      fence_proxy_async();
      mbar_arrive(&a1_full);
      if (pre_idx >= 0) sIn[pre_idx] = pre;                   // rows of tile b+1
      named_bar_sync(C::BAR_IM, 256);                         // ring rows visible / slots reusable
    }
  } else if (warp == 16) {
    // ============================ MMA issuer (one lane) ====================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_tf32(C::M, C::N1);
      const uint32_t idesc2 = make_idesc_tf32(C::M, C::N2);
      const uint32_t idesc3 = make_idesc_tf32(C::M, C::NT3);
      auto issue_mma1 = [&](int b) {
        const uint32_t d1 = tmem + C::cD1 + 64u * (uint32_t)(b & 1);
        const uint32_t sbo = 128 * (C::K1 / 4);
        uint64_t ah = make_desc_kmajor(sA1h, 0, 128, sbo), al = make_desc_kmajor(sA1l, 0, 128, sbo);
        uint64_t bh = make_desc_kmajor(sW1h, 0, 128, sbo), bl = make_desc_kmajor(sW1l, 0, 128, sbo);
#pragma unroll
        for (int ks = 0; ks < C::K1 / 8; ks++) {
#ifndef EXP_MMA_THIRD
          mma_tf32(d1, al, bh, idesc1, ks > 0);
          mma_tf32(d1, ah, bl, idesc1, 1);
          mma_tf32(d1, ah, bh, idesc1, 1);
#else
          mma_tf32(d1, ah, bh, idesc1, ks > 0);
#endif
          ah += 16; al += 16; bh += 16; bl += 16;
        }
        mma_commit(&bar1[b & 1]);
      };
      auto issue_mma2 = [&]() {
        const uint32_t sbo = 128 * (C::K2 / 4);
        uint64_t bh = make_desc_kmajor(sW2h, 0, 128, sbo), bl = make_desc_kmajor(sW2l, 0, 128, sbo);
#pragma unroll
        for (int ks = 0; ks < C::K2 / 8; ks++) {
          mma_tf32_ts(tmem + C::cD2, tmem + C::cA2l + ks * 8, bh, idesc2, ks > 0);
          mma_tf32_ts(tmem + C::cD2, tmem + C::cA2h + ks * 8, bl, idesc2, 1);
          mma_tf32_ts(tmem + C::cD2, tmem + C::cA2h + ks * 8, bh, idesc2, 1);
          bh += 16; bl += 16;
        }
        mma_commit(&bar2);
      };
      auto issue_mma3 = [&]() {
        const uint32_t sbo = 128 * (C::N2 / 4);
        uint64_t bh = make_desc_kmajor(sW3h, 0, 128, sbo), bl = make_desc_kmajor(sW3l, 0, 128, sbo);
#pragma unroll
        for (int ks = 0; ks < C::N2 / 8; ks++) {
          mma_tf32_ts(tmem + C::cD3, tmem + C::cA3l + ks * 8, bh, idesc3, ks > 0);
          mma_tf32_ts(tmem + C::cD3, tmem + C::cA3h + ks * 8, bl, idesc3, 1);
          mma_tf32_ts(tmem + C::cD3, tmem + C::cA3h + ks * 8, bh, idesc3, 1);
          bh += 16; bl += 16;
        }
        mma_commit(&bar3);
      };
      // Issue whatever is ready, in tile order per layer.  Readiness = the operand barrier of
      // that tile has completed AND the accumulator it overwrites has been consumed:
      //   MMA-1(t): a1_full(t);  D1[t&1] was read by E1(t-2)  <=> MMA-2(t-2) already issued
      //   MMA-2(t): a2_full(t);  D2 was read by E23(t-1)      <=> MMA-3(t-1) already issued
      //   MMA-3(t): a3_full(t);  D3 was read by E3(t-1)       <=> d3_free(t-1)
      int n1 = 0, n2 = 0, n3 = 0;
      while (n3 < n_tiles) {
        bool did = false;
        if (n1 < n_tiles && n2 + 2 > n1 && mbar_test(&a1_full, (uint32_t)(n1 & 1))) {
          tcgen05_fence_after();
          issue_mma1(n1);
          n1++;
          did = true;
        }
        if (n2 < n1 && n3 >= n2 && mbar_test(&a2_full, (uint32_t)(n2 & 1))) {
          tcgen05_fence_after();
          issue_mma2();
          n2++;
          did = true;
        }
        if (n3 < n2 && mbar_test(&a3_full, (uint32_t)(n3 & 1)) &&
            (n3 == 0 || mbar_test(&d3_free, (uint32_t)((n3 - 1) & 1)))) {
          tcgen05_fence_after();
          issue_mma3();
          n3++;
          did = true;
        }
        if (!did) __nanosleep(20);
      }
    }
  } else if (warp < 4) {
    // ============================ E1: A2 = split(relu(D1 + b1)) -> TMEM ====================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int b = 0; b < n_tiles; b++) {
      mbar_wait(&bar1[b & 1], (uint32_t)((b >> 1) & 1));       // MMA-1(b) done
      if (b > 0) mbar_wait(&bar2, (uint32_t)((b - 1) & 1));    // MMA-2(b-1) done: A2 free
      tcgen05_fence_after();
      const uint32_t d1 = tmem + lane_base + C::cD1 + 64u * (uint32_t)(b & 1);
#ifdef EXP_SKIP_E1
      if (b < 0)
#endif
#pragma unroll 1
      for (int g = 0; g < 4; g++) {
        float v[16];
        tmem_ld16(d1 + g * 16, v);
#pragma unroll
        for (int h8 = 0; h8 < 2; h8++) {
          float hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; j++)
            split_tf32(fmaxf(v[h8 * 8 + j] + sB1[g * 16 + h8 * 8 + j], 0.f), hi[j], lo[j]);
          tmem_st8(tmem + lane_base + C::cA2h + g * 16 + h8 * 8, hi);
          tmem_st8(tmem + lane_base + C::cA2l + g * 16 + h8 * 8, lo);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(&a2_full);
    }
  } else if (warp < 8) {
    // ============================ E2: A3 = split(relu(D2 + b2)) -> TMEM ====================
    const int q4 = warp - 4;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    for (int b = 0; b < n_tiles; b++) {
      mbar_wait(&bar2, (uint32_t)(b & 1));                     // MMA-2(b) done
      if (b > 0) mbar_wait(&bar3, (uint32_t)((b - 1) & 1));    // MMA-3(b-1) done: A3 free
      tcgen05_fence_after();
      float v[32];
      tmem_ld16(tmem + lane_base + C::cD2, v);
      tmem_ld16(tmem + lane_base + C::cD2 + 16, v + 16);
#pragma unroll
      for (int h8 = 0; h8 < 4; h8++) {
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
          split_tf32(fmaxf(v[h8 * 8 + j] + sB2[h8 * 8 + j], 0.f), hi[j], lo[j]);
        tmem_st8(tmem + lane_base + C::cA3h + h8 * 8, hi);
        tmem_st8(tmem + lane_base + C::cA3l + h8 * 8, lo);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(&a3_full);
    }
  } else if (warp >= 17) {
    // ============================ E3: Q rows -> ring, 25-term gather -> out3 =================
    const int q4 = warp & 3;                  // TMEM lane quarter this warp may access
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int et = tid - 17 * 32;             // 0..127
    const int ep_m = q4 * 32 + lane;          // tile row of this thread's TMEM lane
    const int ep_r = ep_m / C::OW2, ep_x = ep_m % C::OW2;
    const int g_r = et / C::OW3, g_x = et % C::OW3;   // gather: one thread per output pixel
    const bool g_live = et < C::RB * C::OW3;
    for (int b = 0; b < n_tiles; b++) {
      mbar_wait(&bar3, (uint32_t)(b & 1));                     // MMA-3(b) done: Q tile in D3
      tcgen05_fence_after();
      float v[32];
      tmem_ld16(tmem + lane_base + C::cD3, v);
      tmem_ld16(tmem + lane_base + C::cD3 + 16, v + 16);
      tcgen05_fence_before();
      mbar_arrive(&d3_free);                                   // D3 may be overwritten
      const int slot = (b * C::RB + ep_r) % C::RING;
      float* q = sQ + (slot * C::OW2 + ep_x) * C::QP;
#pragma unroll
      for (int j = 0; j < C::QP; j++) q[j] = v[j];
      named_bar_sync(C::BAR_E23, 128);          // Q rows of tile b visible to the gather
      const int j = b * C::RB - (C::F3 - 1) + (g_live ? g_r : 0);
      if (g_live && j >= 0 && j < rows_here) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int dy = 0; dy < C::F3; dy++) {
          const float* qq = sQ + (((j + dy) % C::RING) * C::OW2 + g_x) * C::QP + dy * C::F3;
#pragma unroll
          for (int dx = 0; dx < C::F3; dx++) {
            if ((dy * C::F3 + dx) & 1)
              acc1 += qq[dx * C::QP + dx];
            else
              acc0 += qq[dx * C::QP + dx];
          }
        }
        const int gx = X0 + g_x;
        if (gx < a.w3) dst[(size_t)(R0 + j) * a.w3 + gx] = (acc0 + acc1) + b3;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

inline int configure() {
  SRCNN_CUDA(cudaFuncSetAttribute(forward_fused_ws_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::SMEM_BYTES));
  return SRCNN_OK;
}

inline bool supported(int n1, int n2, int f1, int f2, int f3) {
  return n1 == 64 && n2 == 32 && f1 == 9 && f2 == 1 && f3 == 5;
}

inline int launch(srcnn_ctx* ctx, const fused::Args& a, int S) {
  dim3 grid((a.w3 + Cfg::OW3 - 1) / Cfg::OW3, (a.h3 + Cfg::RPC - 1) / Cfg::RPC, S);
  forward_fused_ws_kernel<<<grid, Cfg::NT, Cfg::SMEM_BYTES, ctx->stream>>>(a);
  return SRCNN_OK;
}

}  // namespace fused_ws
}  // namespace srcnn
