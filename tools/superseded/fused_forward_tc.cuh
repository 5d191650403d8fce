// Fused SRCNN inference, tensor-core version (9-1-5, n1=64, n2=32): all three layers run on
// the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
//   L1  D1[128 px][64] = A1[128 px][88] * W1[64][88]^T      A1 = im2col of the 9x9 window (81 taps
//                                                            + 7 zero columns), built in smem
//   L2  D2[128 px][32] = A2[128 px][64] * W2[32][64]^T      A2 = relu(D1 + b1), TMEM -> regs -> TMEM
//   L3  Q [128 px][32] = A3[128 px][32] * W3[32 taps][32]^T A3 = relu(D2 + b2), TMEM -> regs -> TMEM
//       out3[y][x] = b3 + sum_{dy,dx} Q[(y+dy, x+dx)][dy*5+dx]   (25-term FP32 gather from a ring
//       of Q rows in shared memory).  Layer 3 has one output channel, so instead of an N = 1
//       contraction the 25 taps become the N dimension of a shift-free GEMM: every out2 pixel is
//       multiplied with all 25 tap vectors once, and the spatial shifts move to the gather.
//
// Precision: operands are FP32 values split into TF32 hi + lo; every product is evaluated as
// hi*hi + hi*lo + lo*hi with FP32 accumulation ("3xTF32"), measured at 8e-7 max error against
// fp64 on a K=88 contraction (csrc/probe/tc_probe.cu) -- the same as FP32 FMA.  Plain TF32
// (3e-4) cannot meet the path's 1e-4 tolerance.
//
// Tiling: as in the SIMT kernel a CTA owns 60 output columns x RPC output rows and marches down
// the strip, here 2 rows (= one 128-pixel MMA tile) per step; input rows and out2 rows live in
// circular row buffers.  Operands use the no-swizzle K-major canonical layout (tc_common.cuh).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "context.cuh"
#include "fused_forward.cuh"
#include "tc_common.cuh"

namespace srcnn {
namespace fused_tc {

struct Cfg {
  static constexpr int N1 = 64, N2 = 32, F1 = 9, F3 = 5;
  static constexpr int NW = 512;               // worker threads (im2col, epilogues, gather)
  static constexpr int PARTS = NW / 128;        // workers per tile row / warps per TMEM lane quarter
  static constexpr int NT = NW + 32;           // + one warp that only issues the MMAs
  static constexpr int OW2 = 64, RB = 2;
  static constexpr int OW3 = OW2 - (F3 - 1);
  static constexpr int IW = OW2 + F1 - 1, IWP = IW + 4;
  static constexpr int IR = RB + F1 - 1 + RB;  // one block of slack: rows of tile b+2 land while tile b+1 is read
  static constexpr int RING = RB + F3 - 1;      // ring of Q rows
  static constexpr int NT3 = 32;                // 25 taps padded to an MMA N of 32
  static constexpr int QP = F3 * F3;            // floats per pixel in the Q ring (odd: no conflicts)
  static constexpr int RPC = 128;
  static constexpr int M = OW2 * RB;           // 128 pixels per MMA tile
  static constexpr int K1 = 88;                // 81 taps padded to a multiple of 8
  static constexpr int K2 = N1;
  static constexpr int NPG3 = OW3 / 4;
  // shared memory carve-up (floats)
  static constexpr int oA1h = 0;
  static constexpr int oA1l = oA1h + M * K1;
  static constexpr int oW1h = oA1l + M * K1;
  static constexpr int oW1l = oW1h + N1 * K1;
  static constexpr int oW2h = oW1l + N1 * K1;
  static constexpr int oW2l = oW2h + N2 * K2;
  static constexpr int oW3h = oW2l + N2 * K2;   // B operand of layer 3: [32 taps][32 ch]
  static constexpr int oW3l = oW3h + NT3 * N2;
  static constexpr int oB1 = oW3l + NT3 * N2;
  static constexpr int oB2 = oB1 + N1;
  static constexpr int oIn = oB2 + N2;
  static constexpr int oQ = oIn + IR * IWP;
  static constexpr int TOTAL = oQ + RING * OW2 * QP;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)TOTAL;
  // tensor memory columns: the three accumulators, A2 = relu(D1+b1) and A3 = relu(D2+b2) split
  // into TF32 hi / lo
  static constexpr uint32_t cD1 = 0, cD2 = 64, cD3 = 96, cA2h = 128, cA2l = 192, cA3h = 256,
                            cA3l = 288;
  static constexpr uint32_t TMEM_COLS = 512;
};

__global__ void __launch_bounds__(Cfg::NT, 1) forward_fused_tc_kernel(fused::Args a) {
  using C = Cfg;
  using namespace tc;
  // 128-byte aligned dynamic shared memory (descriptor addresses are in 16-byte units, core
  // matrices are 128 B); indexing the extern array directly keeps the accesses LDS/STS
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
  __shared__ __align__(8) uint64_t bar1, bar2, bar3;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int X0 = blockIdx.x * C::OW3;
  const int R0 = blockIdx.y * C::RPC;
  const float* img = a.in + (size_t)blockIdx.z * a.w * a.h;
  float* dst = a.out + (size_t)blockIdx.z * a.w3 * a.h3;

  // ---- stage parameters: B operands split into TF32 hi/lo, canonical [n][k] layout -------
  for (int i = tid; i < C::N1 * C::K1; i += C::NT) {
    const int n = i / C::K1, k = i % C::K1;
    const float w = k < C::F1 * C::F1 ? __ldg(a.pw1 + k * C::N1 + n) : 0.f;
    float hi, lo;
    split_tf32(w, hi, lo);
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
  for (int i = tid; i < C::N1; i += C::NT) sB1[i] = __ldg(a.pb1 + i);
  for (int i = tid; i < C::N2; i += C::NT) sB2[i] = __ldg(a.pb2 + i);
  for (int i = tid; i < C::NT3 * C::N2; i += C::NT) {
    const int n = i / C::N2, k = i % C::N2;   // n = tap dy*5+dx, k = channel: W3 is [tap][c2]
    float hi, lo;
    split_tf32(n < C::F3 * C::F3 ? __ldg(a.pw3 + n * C::N2 + k) : 0.f, hi, lo);
    sW3h[kmajor_offset(n, k, C::N2)] = hi;
    sW3l[kmajor_offset(n, k, C::N2)] = lo;
  }
  const float b3 = __ldg(a.pb3);

  auto load_rows = [&](int first_rel_row, int count) {
    for (int i = tid; i < count * C::IW; i += C::NW) {
      const int rr = first_rel_row + i / C::IW, xx = i % C::IW;
      const int gy = R0 + rr, gx = X0 + xx;
      const float v = (gy < a.h && gx < a.w) ? __ldg(img + (size_t)gy * a.w + gx) : 0.f;
      sIn[(rr % C::IR) * C::IWP + xx] = v;
    }
  };
  load_rows(0, C::F1 - 1 + C::RB);   // input rows of tile 0

  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  if (tid == 0) {
    if (smem_u32(smem) & 127u) __trap();   // operand layout needs a 128-byte aligned base
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    mbar_init(&bar3, 1);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc1 = make_idesc_tf32(C::M, C::N1);
  const uint32_t idesc2 = make_idesc_tf32(C::M, C::N2);
  const uint32_t idesc3 = make_idesc_tf32(C::M, C::NT3);

  const int rows_here = min(C::RPC, a.h3 - R0);
  const int n_blocks = (rows_here + (C::F3 - 1) + C::RB - 1) / C::RB;

  // im2col role: row m of the tile and half of the K chunks
  const int im_m = tid & (C::M - 1);
  const int im_part = tid >> 7;              // which quarter of the 22 K chunks
  const int im_r = im_m / C::OW2, im_x = im_m % C::OW2;
  // epilogue role: TMEM lane quarter + column half
  const int ep_q = warp & 3, ep_h = (warp >> 2) & 3;   // lane quarter, column part (0..3)
  const int ep_m = ep_q * 32 + lane;
  // gather role: two threads per output pixel of the tile's RB x OW3 outputs
  const int g_px = tid >> 2, g_part = tid & 3;   // four threads per output pixel
  const bool g_live = g_px < C::RB * C::OW3;
  const int g_r = g_live ? g_px / C::OW3 : 0, g_x = g_live ? g_px % C::OW3 : 0;

  // ---- pipeline stages ------------------------------------------------------------------
  // im2col of tile b: A1[m][k] = in[r+dy][x+dx], k = dy*9+dx, split hi/lo
  auto im2col = [&](int b) {
    const int base_slot = (b * C::RB + im_r) % C::IR;
    // ring-row pointers of the nine input rows this pixel reads, computed once
    const float* rowp[C::F1];
#pragma unroll
    for (int dy = 0; dy < C::F1; dy++) {
      int slot = base_slot + dy;
      slot = slot >= C::IR ? slot - C::IR : slot;
      rowp[dy] = sIn + slot * C::IWP + im_x;
    }
    // the K half is a compile-time constant inside `half`, so every (dy, dx) below is too and
    // rowp[] stays in registers
    auto part = [&](auto part_tag) {
      constexpr int PI = decltype(part_tag)::value;
      constexpr int C0 = PI == 0 ? 0 : PI == 1 ? 6 : PI == 2 ? 12 : 17;   // 6 + 6 + 5 + 5 chunks
      constexpr int C1 = PI == 0 ? 6 : PI == 1 ? 12 : PI == 2 ? 17 : 22;
#pragma unroll
      for (int c = C0; c < C1; c++) {   // 16-byte K chunk
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
    if (im_part == 0)
      part(std::integral_constant<int, 0>{});
    else if (im_part == 1)
      part(std::integral_constant<int, 1>{});
    else if (im_part == 2)
      part(std::integral_constant<int, 2>{});
    else
      part(std::integral_constant<int, 3>{});
  };
  // MMA-1: D1 = A1 * W1^T  (33 x M128 N64 K8), one thread.  A k-step advances every operand by
  // two 128-byte core matrices = 16 in the descriptor's 16-byte address units.
  auto issue_mma1 = [&]() {
    tcgen05_fence_after();
    const uint32_t sbo = 128 * (C::K1 / 4);
    uint64_t ah = make_desc_kmajor(sA1h, 0, 128, sbo), al = make_desc_kmajor(sA1l, 0, 128, sbo);
    uint64_t bh = make_desc_kmajor(sW1h, 0, 128, sbo), bl = make_desc_kmajor(sW1l, 0, 128, sbo);
#pragma unroll
    for (int ks = 0; ks < C::K1 / 8; ks++) {
      mma_tf32(tmem + C::cD1, al, bh, idesc1, ks > 0);
      mma_tf32(tmem + C::cD1, ah, bl, idesc1, 1);
      mma_tf32(tmem + C::cD1, ah, bh, idesc1, 1);
      ah += 16; al += 16; bh += 16; bl += 16;
    }
    mma_commit(&bar1);
  };
  // MMA-2: D2 = A2 * W2^T  (24 x M128 N32 K8), A2 hi/lo in tensor memory
  auto issue_mma2 = [&]() {
    tcgen05_fence_after();
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
  // epilogue 1: A2 = split(relu(D1 + b1)) -> tensor memory
  auto epilogue1 = [&]() {
    const uint32_t lane_base = (uint32_t)(ep_q * 32) << 16;
    const int c0 = ep_h * 16;
    float v[16];
    tmem_ld16(tmem + lane_base + C::cD1 + c0, v);
#pragma unroll
    for (int h8 = 0; h8 < 2; h8++) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        split_tf32(fmaxf(v[h8 * 8 + j] + sB1[c0 + h8 * 8 + j], 0.f), hi[j], lo[j]);
      tmem_st8(tmem + lane_base + C::cA2h + c0 + h8 * 8, hi);
      tmem_st8(tmem + lane_base + C::cA2l + c0 + h8 * 8, lo);
    }
    tmem_st_wait();
  };
  // MMA-3: Q = A3 * W3^T  (12 x M128 N32 K8), A3 hi/lo in tensor memory
  auto issue_mma3 = [&]() {
    tcgen05_fence_after();
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
  // epilogue 2: A3 = split(relu(D2 + b2)) -> tensor memory
  auto epilogue2 = [&]() {
    const uint32_t lane_base = (uint32_t)(ep_q * 32) << 16;
    const int c0 = ep_h * 8;
    float v[8], hi[8], lo[8];
    tmem_ld8(tmem + lane_base + C::cD2 + c0, v);
#pragma unroll
    for (int j = 0; j < 8; j++) split_tf32(fmaxf(v[j] + sB2[c0 + j], 0.f), hi[j], lo[j]);
    tmem_st8(tmem + lane_base + C::cA3h + c0, hi);
    tmem_st8(tmem + lane_base + C::cA3l + c0, lo);
    tmem_st_wait();
  };
  // epilogue 3 of tile b: Q rows (25 tap products per out2 pixel) -> ring in shared memory
  auto epilogue3 = [&](int b) {
    const uint32_t taddr = tmem + ((uint32_t)(ep_q * 32) << 16) + C::cD3 + ep_h * 8;
    float v[8];
    tmem_ld8(taddr, v);
    const int r = ep_m / C::OW2, x = ep_m % C::OW2;
    const int slot = (b * C::RB + r) % C::RING;
    float* q = sQ + (slot * C::OW2 + x) * C::QP + ep_h * 8;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (ep_h * 8 + j < C::QP) q[j] = v[j];
  };
  // gather of the output rows completed by tile b:
  // out3[j][x] = b3 + sum_{dy,dx} Q[j+dy][x+dx][dy*5+dx], four threads per pixel
  auto gather = [&](int b) {
    const int j = b * C::RB - (C::F3 - 1) + g_r;     // output row relative to R0
    float acc = 0.f;
    if (g_live && j >= 0) {
#pragma unroll
      for (int dy = 0; dy < C::F3; dy++) {
        const int slot = (j + dy) % C::RING;
        const float* q = sQ + (slot * C::OW2 + g_x) * C::QP + dy * C::F3;
#pragma unroll
        for (int dx = 0; dx < C::F3; dx++) {
          const int tap = dy * C::F3 + dx;
          if (tap % 4 == g_part) acc += q[dx * C::QP + dx];
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (g_live && g_part == 0 && j >= 0 && j < rows_here) {
      const int gx = X0 + g_x;
      if (gx < a.w3) dst[(size_t)(R0 + j) * a.w3 + gx] = acc + b3;
    }
  };

  // ---- software pipeline: MMA-1(b) runs under epilogue-2 + L3 of tile b-1, MMA-2(b) under the
  // im2col of tile b+1.  Warp 8 only issues the MMAs (one lane), so descriptor set-up and the
  // 33 + 24 issue slots are off the workers' critical path; it takes part in the same
  // __syncthreads sequence as the workers.
  const bool is_worker = tid < C::NW;
  // input rows of tile b+2 are fetched into a register before the im2col of tile b+1 and
  // stored after it (one element per thread covers RB rows), hiding the global-load latency
  auto prefetch_row_elem = [&](int first_rel_row, float& v, int& dst_index) {
    dst_index = -1;
    if (tid < C::RB * C::IW) {
      const int rr = first_rel_row + tid / C::IW, xx = tid % C::IW;
      const int gy = R0 + rr, gx = X0 + xx;
      v = (gy < a.h && gx < a.w) ? __ldg(img + (size_t)gy * a.w + gx) : 0.f;
      dst_index = (rr % C::IR) * C::IWP + xx;
    }
  };
  static_assert(C::RB * C::IW <= C::NW, "one prefetched element per worker thread");

  if (is_worker) {
    im2col(0);
    if (n_blocks > 1) load_rows(C::RB + C::F1 - 1, C::RB);   // input rows of tile 1
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();                                      // (S0) A1(0) complete
  if (is_worker) {
    for (int b = 0; b < n_blocks; b++) {
      mbar_wait(&bar1, (uint32_t)(b & 1));              // MMA-1(b) done: D1 ready, A1 free
      tcgen05_fence_after();
      epilogue1();
      tcgen05_fence_before();
      __syncthreads();                                  // (SA) A2 complete -> MMA-2(b)
      if (b + 1 < n_blocks) {
        float pre = 0.f;
        int pre_idx = -1;
        if (b + 2 < n_blocks) prefetch_row_elem((b + 2) * C::RB + C::F1 - 1, pre, pre_idx);
        im2col(b + 1);
        if (pre_idx >= 0) sIn[pre_idx] = pre;
        fence_proxy_async();
      }
      tcgen05_fence_before();
      __syncthreads();                                  // (SB) A1(b+1) complete -> MMA-1(b+1)
      if (b > 0) {
        mbar_wait(&bar3, (uint32_t)((b - 1) & 1));      // MMA-3(b-1) done: Q tile ready, A3 free
        tcgen05_fence_after();
        epilogue3(b - 1);
      }
      mbar_wait(&bar2, (uint32_t)(b & 1));              // MMA-2(b) done: D2 ready
      tcgen05_fence_after();
      epilogue2();
      tcgen05_fence_before();
      __syncthreads();                                  // (SC) A3 complete, D3 read -> MMA-3(b)
      if (b > 0) gather(b - 1);
    }
    mbar_wait(&bar3, (uint32_t)((n_blocks - 1) & 1));
    tcgen05_fence_after();
    epilogue3(n_blocks - 1);
    __syncthreads();                                    // (SD)
    gather(n_blocks - 1);
  } else {
    const bool issuer = lane == 0;
    if (issuer) issue_mma1();                           // tile 0
    __syncwarp();
    for (int b = 0; b < n_blocks; b++) {
      tcgen05_fence_before();
      __syncthreads();                                  // (SA)
      if (issuer) issue_mma2();
      __syncwarp();
      tcgen05_fence_before();
      __syncthreads();                                  // (SB)
      if (issuer && b + 1 < n_blocks) issue_mma1();
      __syncwarp();
      tcgen05_fence_before();
      __syncthreads();                                  // (SC)
      if (issuer) issue_mma3();
      __syncwarp();
    }
    __syncthreads();                                    // (SD)
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

inline int configure() {
  SRCNN_CUDA(cudaFuncSetAttribute(forward_fused_tc_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::SMEM_BYTES));
  return SRCNN_OK;
}

inline bool supported(int n1, int n2, int f1, int f2, int f3) {
  return n1 == 64 && n2 == 32 && f1 == 9 && f2 == 1 && f3 == 5;
}

inline int launch(srcnn_ctx* ctx, const fused::Args& a, int S) {
  dim3 grid((a.w3 + Cfg::OW3 - 1) / Cfg::OW3, (a.h3 + Cfg::RPC - 1) / Cfg::RPC, S);
  forward_fused_tc_kernel<<<grid, Cfg::NT, Cfg::SMEM_BYTES, ctx->stream>>>(a);
  return SRCNN_OK;
}

}  // namespace fused_tc
}  // namespace srcnn
