// Fused SRCNN inference for f2 == 1 networks (9-1-5): layers 1-3 in ONE launch.
//
// reference semantics: three `forward` launches, src/ConfigBasedDataPipeline.cpp:200-241 over
// src/kernel/layer_uber_kernel.cl:36-96, each materialising its output in global memory
// (4.28 GB + 2.14 GB for a 4096x4096 image).  Here the n1- and n2-channel maps never leave
// the SM: HBM traffic is the input luma once (plus an 18-25 % halo re-read that the L2
// absorbs) and the output luma once -- 8 B/pixel instead of 776 B/pixel.
//
// Decomposition (FP32 SIMT, 256 threads, one CTA per SM):
//   * a CTA owns a column strip of OW3 = OW2-4 output columns and RPC output rows and marches
//     down it in blocks of RB rows.  Per block:
//       L1  out1[RB x OW2 px][N1] = relu(b1 + W1 * in-window)   thread tile 8 px x 8 ch, the
//           9 input taps of a row slide through registers; result -> smem, channel-major
//       L2  out2[RB x OW2 px][N2] = relu(b2 + W2 * out1)        thread tile 8 px x 4 ch2;
//           result -> smem ring of RB+4 rows, channel-major
//       L3  out3[RB x OW3 px]     = b3 + W3 * out2-window       thread tile 4 px x RB rows x
//           N2/16 ch2 (taps in registers, each out2 row reused by every output row that
//           sees it); the 16 channel slices are folded with four warp shuffles; -> global
//   * the input rows and the out2 rows live in circular row buffers, so vertical halo
//     recomputation is (RPC+4)/RPC and horizontal OW2/OW3 (1.03 x 1.07 for 64/32).
//   * two __syncthreads per block; weights (W1 20.7 KB, W2 8 KB, W3) stay in smem.
#pragma once
#include <cuda_runtime.h>

#include "context.cuh"

namespace srcnn {
namespace fused {

template <int N1_, int N2_, int OW2_, int RB_>
struct Cfg {
  static constexpr int N1 = N1_, N2 = N2_, OW2 = OW2_, RB = RB_;
  static constexpr int F1 = 9, F3 = 5;
  static constexpr int NT = 256;
  static constexpr int OW3 = OW2 - (F3 - 1);   // output columns per strip
  static constexpr int IW = OW2 + F1 - 1;      // input columns per strip
  static constexpr int IWP = IW + 4;           // padded row pitch of the input ring
  static constexpr int IR = RB + F1 - 1;       // input ring rows
  static constexpr int PX = OW2 * RB;          // pixels per block
  static constexpr int PXP = PX + 4;           // pitch of one out1 channel plane
  static constexpr int RING = RB + F3 - 1;     // out2 ring rows
  static constexpr int OW2P = OW2 + 4;         // pitch of one out2 channel row
  static constexpr int RPC = 128;              // output rows per CTA
  static constexpr int L1_TASKS = (OW2 / 8) * RB * (N1 / 8);
  static constexpr int L2_TASKS = (PX / 8) * (N2 / 4);
  static constexpr int NPG3 = OW3 / 4;
  static_assert(NPG3 * 16 <= NT, "L3 thread tiling");
  static_assert(L1_TASKS == NT, "L1 thread tiling must cover the block exactly");
  static_assert(L2_TASKS == NT, "L2 thread tiling must cover the block exactly");
  static_assert(OW2 % 8 == 0 && OW3 % 4 == 0 && N2 % 16 == 0, "tile shape");
  // shared memory carve-up (floats)
  static constexpr int oW1 = 0;
  static constexpr int oB1 = oW1 + F1 * F1 * N1;
  static constexpr int oW2 = oB1 + N1;
  static constexpr int oB2 = oW2 + N1 * N2;
  static constexpr int oW3 = oB2 + N2;           // repacked, 8 floats per (dy, c2)
  static constexpr int oIn = oW3 + F3 * N2 * 8;
  static constexpr int oO1 = oIn + IR * IWP;
  static constexpr int oO2 = oO1 + N1 * PXP;
  static constexpr int TOTAL = oO2 + RING * N2 * OW2P;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)TOTAL;
};

struct Args {
  const float* in;   // [S][h][w]
  float* out;        // [S][h3][w3]
  const float *pw1, *pb1, *pw2, *pb2, *pw3, *pb3;
  int w, h, w3, h3;
  const int* gate = nullptr;   // device word: run only if *gate != 0 (null: always run)
};

template <class C>
__global__ void __launch_bounds__(256, 1) forward_fused_kernel(Args a) {
  extern __shared__ __align__(16) float smem[];
  float* sW1 = smem + C::oW1;
  float* sB1 = smem + C::oB1;
  float* sW2 = smem + C::oW2;
  float* sB2 = smem + C::oB2;
  float* sW3 = smem + C::oW3;
  float* sIn = smem + C::oIn;
  float* sO1 = smem + C::oO1;
  float* sO2 = smem + C::oO2;
  if (a.gate && *a.gate == 0) return;   // a tensor-core launch ahead of us did the work

  const int tid = threadIdx.x;
  const int X0 = blockIdx.x * C::OW3;   // first output column == first input column
  const int R0 = blockIdx.y * C::RPC;   // first output row    == first input row
  const float* img = a.in + (size_t)blockIdx.z * a.w * a.h;
  float* dst = a.out + (size_t)blockIdx.z * a.w3 * a.h3;

  // ---- stage parameters ------------------------------------------------------------
  for (int i = tid; i < C::F1 * C::F1 * C::N1; i += C::NT) sW1[i] = __ldg(a.pw1 + i);
  for (int i = tid; i < C::N1; i += C::NT) sB1[i] = __ldg(a.pb1 + i);
  for (int i = tid; i < C::N1 * C::N2; i += C::NT) sW2[i] = __ldg(a.pw2 + i);
  for (int i = tid; i < C::N2; i += C::NT) sB2[i] = __ldg(a.pb2 + i);
  // W3 repacked: taps dx 0..3 as [dy][c2][4], tap dx 4 as [dy][c2] behind them, so that the
  // 16 lanes of a pixel group read consecutive 16-byte / 4-byte words (no bank conflicts)
  for (int i = tid; i < C::F3 * C::N2 * 5; i += C::NT) {
    const int dx = i % 5, c2 = (i / 5) % C::N2, dy = i / (5 * C::N2);
    const float v = __ldg(a.pw3 + (dy * C::F3 + dx) * C::N2 + c2);
    if (dx < 4)
      sW3[(dy * C::N2 + c2) * 4 + dx] = v;
    else
      sW3[C::F3 * C::N2 * 4 + dy * C::N2 + c2] = v;
  }
  const float b3 = __ldg(a.pb3);

  // input rows [R0, R0+F1-1) -> ring slots 0..F1-2; rows outside the image read as 0
  auto load_rows = [&](int first_rel_row, int count) {
    for (int i = tid; i < count * C::IW; i += C::NT) {
      const int rr = first_rel_row + i / C::IW, xx = i % C::IW;
      const int gy = R0 + rr, gx = X0 + xx;
      const float v = (gy < a.h && gx < a.w) ? __ldg(img + (size_t)gy * a.w + gx) : 0.f;
      sIn[(rr % C::IR) * C::IWP + xx] = v;
    }
  };
  load_rows(0, C::F1 - 1);
  load_rows(C::F1 - 1, C::RB);   // rows of block 0
  __syncthreads();

  const int rows_here = min(C::RPC, a.h3 - R0);
  const int n_blocks = (rows_here + (C::F3 - 1) + C::RB - 1) / C::RB;

  // thread roles
  // L1: channel group is warp-uniform (weights broadcast), lanes sweep pixel groups / rows
  const int l1_pg = tid % (C::OW2 / 8);
  const int l1_r = (tid / (C::OW2 / 8)) % C::RB;
  const int l1_cg = tid / ((C::OW2 / 8) * C::RB);
  // L2: 8 consecutive lanes sweep 8 pixel groups, channel group varies slower
  const int l2_pg = tid % (C::PX / 8);
  const int l2_cg = tid / (C::PX / 8);

  for (int b = 0; b < n_blocks; b++) {
    const int y0 = b * C::RB;   // first out1/out2 row (relative to R0) of this block

    // ================= L1: 8 px x 8 ch per thread =====================================
    {
      float acc[8][8];
#pragma unroll
      for (int p = 0; p < 8; p++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[p][c] = 0.f;
      int slot = (y0 + l1_r) % C::IR;
#pragma unroll 1
      for (int dy = 0; dy < C::F1; dy++) {
        const float4* ip = reinterpret_cast<const float4*>(sIn + slot * C::IWP + l1_pg * 8);
        const float4 i0 = ip[0], i1 = ip[1], i2 = ip[2], i3 = ip[3];
        const float iv[16] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w,
                              i2.x, i2.y, i2.z, i2.w, i3.x, i3.y, i3.z, i3.w};
        const float4* wp =
            reinterpret_cast<const float4*>(sW1 + (dy * C::F1) * C::N1 + l1_cg * 8);
#pragma unroll
        for (int dx = 0; dx < C::F1; dx++) {
          const float4 wa = wp[dx * (C::N1 / 4)], wb = wp[dx * (C::N1 / 4) + 1];
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int p = 0; p < 8; p++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[p][c] = fmaf(iv[p + dx], wv[c], acc[p][c]);
        }
        slot = slot + 1 == C::IR ? 0 : slot + 1;
      }
      // bias + ReLU -> channel-major plane [c][r*OW2 + x]
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int ch = l1_cg * 8 + c;
        const float bb = sB1[ch];
        float4 v0, v1;
        v0.x = fmaxf(acc[0][c] + bb, 0.f);
        v0.y = fmaxf(acc[1][c] + bb, 0.f);
        v0.z = fmaxf(acc[2][c] + bb, 0.f);
        v0.w = fmaxf(acc[3][c] + bb, 0.f);
        v1.x = fmaxf(acc[4][c] + bb, 0.f);
        v1.y = fmaxf(acc[5][c] + bb, 0.f);
        v1.z = fmaxf(acc[6][c] + bb, 0.f);
        v1.w = fmaxf(acc[7][c] + bb, 0.f);
        float4* op = reinterpret_cast<float4*>(sO1 + ch * C::PXP + l1_r * C::OW2 + l1_pg * 8);
        op[0] = v0;
        op[1] = v1;
      }
    }
    __syncthreads();   // S1: out1 block complete; input ring rows of this block are dead

    // prefetch the next block's input rows (overwrites the RB oldest ring rows)
    if (b + 1 < n_blocks) load_rows((b + 1) * C::RB + C::F1 - 1, C::RB);

    // ================= L2: 8 px x 4 ch2 per thread ====================================
    {
      float acc[8][4];
#pragma unroll
      for (int p = 0; p < 8; p++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[p][c] = 0.f;
      const float* ap = sO1 + l2_pg * 8;
      const float* wp = sW2 + l2_cg * 4;
#pragma unroll 8
      for (int c = 0; c < C::N1; c++) {
        const float4 a0 = *reinterpret_cast<const float4*>(ap + c * C::PXP);
        const float4 a1 = *reinterpret_cast<const float4*>(ap + c * C::PXP + 4);
        const float4 ww = *reinterpret_cast<const float4*>(wp + c * C::N2);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int p = 0; p < 8; p++)
#pragma unroll
          for (int k = 0; k < 4; k++) acc[p][k] = fmaf(av[p], wv[k], acc[p][k]);
      }
      const int pl = l2_pg * 8;             // linear pixel index in the block
      const int r = pl / C::OW2, x = pl % C::OW2;
      const int slot = (y0 + r) % C::RING;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int c2 = l2_cg * 4 + k;
        const float bb = sB2[c2];
        float4 v0, v1;
        v0.x = fmaxf(acc[0][k] + bb, 0.f);
        v0.y = fmaxf(acc[1][k] + bb, 0.f);
        v0.z = fmaxf(acc[2][k] + bb, 0.f);
        v0.w = fmaxf(acc[3][k] + bb, 0.f);
        v1.x = fmaxf(acc[4][k] + bb, 0.f);
        v1.y = fmaxf(acc[5][k] + bb, 0.f);
        v1.z = fmaxf(acc[6][k] + bb, 0.f);
        v1.w = fmaxf(acc[7][k] + bb, 0.f);
        float4* op = reinterpret_cast<float4*>(sO2 + (slot * C::N2 + c2) * C::OW2P + x);
        op[0] = v0;
        op[1] = v1;
      }
    }
    __syncthreads();   // S2: out2 rows of this block complete; next input rows landed

    // ================= L3: 4 px x RB rows x N2/16 ch2 per thread ======================
    // 16 lanes share a pixel group and split the channels (c2 = lane16 + 16*i); each lane
    // keeps the 5x5 taps of its channel in registers and reuses every out2 row it loads for
    // all output rows that see it.  The 16 partial sums are folded with four shuffles.
    {
      const int cgi = tid % 16;
      const int pxg = tid / 16;
      const bool live = pxg < C::NPG3;
      const int pxa = live ? pxg : 0;
      const int j0 = y0 - (C::F3 - 1);        // first output row of this block, rel. to R0
      // ring slot of out2 row j0 (negative in the first block: those slots hold stale data
      // that only ever reaches output rows < 0, which are masked at the store)
      const int slot0 = ((j0 % C::RING) + C::RING) % C::RING;
      float acc[C::RB][4];
#pragma unroll
      for (int r = 0; r < C::RB; r++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[r][p] = 0.f;
#pragma unroll 1
      for (int i = 0; i < C::N2 / 16; i++) {
        const int c2 = cgi + 16 * i;
        float wv[C::F3][C::F3];
#pragma unroll
        for (int dy = 0; dy < C::F3; dy++) {
          const float4 w0 = *reinterpret_cast<const float4*>(sW3 + (dy * C::N2 + c2) * 4);
          wv[dy][0] = w0.x;
          wv[dy][1] = w0.y;
          wv[dy][2] = w0.z;
          wv[dy][3] = w0.w;
          wv[dy][4] = sW3[C::F3 * C::N2 * 4 + dy * C::N2 + c2];
        }
        int slot = slot0;
#pragma unroll
        for (int jj = 0; jj < C::RB + C::F3 - 1; jj++) {
          const float* vp = sO2 + (slot * C::N2 + c2) * C::OW2P + pxa * 4;
          const float4 v0 = *reinterpret_cast<const float4*>(vp);
          const float4 v1 = *reinterpret_cast<const float4*>(vp + 4);
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int r = 0; r < C::RB; r++) {
            const int dy = jj - r;
            if (dy >= 0 && dy < C::F3) {
#pragma unroll
              for (int dx = 0; dx < C::F3; dx++)
#pragma unroll
                for (int p = 0; p < 4; p++) acc[r][p] = fmaf(vv[p + dx], wv[dy][dx], acc[r][p]);
            }
          }
          slot = slot + 1 == C::RING ? 0 : slot + 1;
        }
      }
#pragma unroll
      for (int r = 0; r < C::RB; r++)
#pragma unroll
        for (int p = 0; p < 4; p++) {
          float v = acc[r][p];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          acc[r][p] = v;
        }
      if (live && cgi == 0) {
#pragma unroll
        for (int r = 0; r < C::RB; r++) {
          const int j = j0 + r;
          if (j >= 0 && j < rows_here) {
            const int gy = R0 + j;
#pragma unroll
            for (int p = 0; p < 4; p++) {
              const int gx = X0 + pxg * 4 + p;
              if (gx < a.w3) dst[(size_t)gy * a.w3 + gx] = acc[r][p] + b3;
            }
          }
        }
      }
    }
    // no barrier needed here: the next L1 writes sO1 (last read before S2) and reads the
    // input ring (complete at S2); the next L2 writes ring rows only after the next S1.
  }
}

using Cfg_64_32 = Cfg<64, 32, 64, 4>;
using Cfg_128_64 = Cfg<128, 64, 32, 4>;
using Cfg_32_16 = Cfg<32, 16, 64, 8>;

template <class C>
inline int configure_one() {
  SRCNN_CUDA(cudaFuncSetAttribute(forward_fused_kernel<C>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)C::SMEM_BYTES));
  return SRCNN_OK;
}

inline int configure() {
  SRCNN_TRY(configure_one<Cfg_64_32>());
  SRCNN_TRY(configure_one<Cfg_128_64>());
  SRCNN_TRY(configure_one<Cfg_32_16>());
  return SRCNN_OK;
}

inline bool supported(int n1, int n2, int f1, int f2, int f3) {
  if (f1 != 9 || f2 != 1 || f3 != 5) return false;
  return (n1 == 64 && n2 == 32) || (n1 == 128 && n2 == 64) || (n1 == 32 && n2 == 16);
}

template <class C>
inline int launch_one(srcnn_ctx* ctx, const Args& a, int S) {
  dim3 grid((a.w3 + C::OW3 - 1) / C::OW3, (a.h3 + C::RPC - 1) / C::RPC, S);
  forward_fused_kernel<C><<<grid, C::NT, C::SMEM_BYTES, ctx->stream>>>(a);
  return SRCNN_OK;
}

inline int launch(srcnn_ctx* ctx, int n1, int n2, const Args& a, int S) {
  if (n1 == 64 && n2 == 32) return launch_one<Cfg_64_32>(ctx, a, S);
  if (n1 == 128 && n2 == 64) return launch_one<Cfg_128_64>(ctx, a, S);
  if (n1 == 32 && n2 == 16) return launch_one<Cfg_32_16>(ctx, a, S);
  return fail(SRCNN_EINVAL, "no fused forward instantiation for n1=%d n2=%d", n1, n2);
}

}  // namespace fused
}  // namespace srcnn
