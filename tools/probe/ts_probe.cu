// Stand-alone probe: the tensor-pipe cost of a tcgen05.mma whose A operand lies in TENSOR MEMORY
// (the layer-2 / layer-3 instructions of the fused SRCNN kernels), alone and mixed with
// shared-memory-operand instructions (layer 1).  Question: do the two operand paths (shared
// memory 128 B/clk, tensor memory) overlap when their instructions alternate in the pipe, or is
// a tile's cost the SUM of its instructions?  kind::f16, M = 128, K = 16 per instruction.
//   modes: "ss"  R instructions with A and B in shared memory, N = n_ss
//          "ts"  R instructions with A in tensor memory,        N = n_ts
//          "mix" R of each, alternating in ONE issuing thread
//          "two" R of each from TWO issuing threads (different warps)
//          "blk" R of each from one thread in blocks of 12 ss / 12 ts (the order of a tile)
// Every run checks the accumulator values, so a mis-encoded instruction cannot pass as "fast".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_probe ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

__host__ __device__ inline uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ inline bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__device__ inline void tmem_ld1(uint32_t taddr, float& v) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v = __uint_as_float(r);
}
__device__ inline void tmem_st8u(uint32_t taddr, uint32_t x) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(x)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

enum { SS = 0, TS = 1, MIX = 2, TWO = 3, BLK = 4 };

// TMEM columns: [0,8) A operand (half pairs), D_ss at 64, D_ts at 320
// bg: what warps 4..7 (one per TMEM lane quarter) do while the MMAs run -- 0 nothing,
// 1 tcgen05.ld x16 loops, 2 tcgen05.st x8 loops, 3 ld.shared / st.shared loops: does SIMT traffic
// on tensor memory or shared memory slow the tensor pipe down?
__global__ void __launch_bounds__(256) probe(int mode, int n_ss, int n_ts, int R, int bg, long long* cycles,
                                             float* dval) {
  __shared__ volatile int stop_flag;
  __shared__ float bg_buf[4 * 32 * 33];
  __shared__ __align__(128) __half sA[128 * 16];
  __shared__ __align__(128) __half sB[256 * 16];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  if (tid == 0) stop_flag = 0;
  for (int i = tid; i < 128 * 16; i += 256) sA[i] = __float2half(1.f);
  for (int i = tid; i < 256 * 16; i += 256) sB[i] = __float2half(1.f);
  for (int i = tid; i < 4 * 32 * 33; i += 256) bg_buf[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  {   // A operand in tensor memory: 2.0 in every element (so the two paths give different sums)
    const __half2 two = __floats2half2_rn(2.f, 2.f);
    if (warp < 4) tmem_st8u(tmem + ((uint32_t)(warp * 32) << 16), *reinterpret_cast<const uint32_t*>(&two));
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint64_t ad = make_desc_kmajor(sA, 0, 128, 256);
  const uint64_t bd = make_desc_kmajor(sB, 0, 128, 256);
  const uint32_t i_ss = idesc_f16(128, n_ss), i_ts = idesc_f16(128, n_ts);
  const uint32_t d_ss = tmem + 64, d_ts = tmem + 320;
  const bool two = mode == TWO;
  if (warp == 0 || (two && warp == 1)) {
    const long long t0 = clock64();
    if (elect_one()) {
      if (mode == SS || (two && warp == 0))
        for (int r = 0; r < R; r++) mma_ss(d_ss, ad, bd, i_ss, r > 0);
      else if (mode == TS || (two && warp == 1))
        for (int r = 0; r < R; r++) mma_ts(d_ts, tmem, bd, i_ts, r > 0);
      else if (mode == MIX)
        for (int r = 0; r < R; r++) {
          mma_ss(d_ss, ad, bd, i_ss, r > 0);
          mma_ts(d_ts, tmem, bd, i_ts, r > 0);
        }
      else
        for (int r = 0; r < R; r += 12) {
#pragma unroll
          for (int j = 0; j < 12; j++) mma_ss(d_ss, ad, bd, i_ss, (r + j) > 0);
#pragma unroll
          for (int j = 0; j < 12; j++) mma_ts(d_ts, tmem, bd, i_ts, (r + j) > 0);
        }
      mma_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if ((tid & 31) == 0) cycles[warp] = clock64() - t0;
    if (warp == 0) stop_flag = 1;
  } else if (warp >= 4 && bg != 0) {
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 480;   // unused columns
    float acc = 0.f;
    float* mine = bg_buf + (warp - 4) * 32 * 33 + (tid & 31) * 33;
    while (!stop_flag) {
      if (bg == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
              "%12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15])
              : "r"(ta));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += __uint_as_float(r[0] ^ r[15]);
        }
      } else if (bg == 2) {
#pragma unroll
        for (int i = 0; i < 4; i++) tmem_st8u(ta + 8 * (i & 1), 0u);
      } else {
#pragma unroll
        for (int i = 0; i < 25; i++) mine[i] = acc + (float)i;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 25; i++) acc += mine[(i * 7) % 25];
      }
    }
    if (acc == 123.456f) dval[3] = acc;
  }
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0) {
    float v, w;
    tmem_ld1(d_ss, v);
    tmem_ld1(d_ts, w);
    if (tid == 0) { dval[0] = v; dval[1] = w; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run(int mode, int n_ss, int n_ts, int R, int bg, long long* dc, float* dv) {
  static const char* names[] = {"ss ", "ts ", "mix", "two", "blk"};
  long long best[2] = {1LL << 60, 1LL << 60};
  float v[2] = {0, 0};
  for (int rep = 0; rep < 3; rep++) {
    cudaMemset(dc, 0, 4 * sizeof(long long));
    cudaMemset(dv, 0, 4 * sizeof(float));
    probe<<<1, 256>>>(mode, n_ss, n_ts, R, bg, dc, dv);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", names[mode], cudaGetErrorString(e));
      exit(1);
    }
    long long c[2];
    cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    cudaMemcpy(v, dv, sizeof(v), cudaMemcpyDeviceToHost);
    for (int w = 0; w < 2; w++) best[w] = (c[w] > 0 && c[w] < best[w]) ? c[w] : best[w];
  }
  const bool has_ss = mode != TS, has_ts = mode != SS;
  const float e_ss = has_ss ? 16.f * R : 0.f, e_ts = has_ts ? 32.f * R : 0.f;
  const bool ok = (!has_ss || v[0] == e_ss) && (!has_ts || v[1] == e_ts);
  const long long tot = mode == TWO ? (best[0] > best[1] ? best[0] : best[1]) : best[0];
  const int n_inst = (has_ss ? R : 0) + (has_ts ? R : 0);
  static const char* bgn[] = {"quiet     ", "bg tmem ld", "bg tmem st", "bg smem   "};
  printf("%s ", bgn[bg]);
  printf("%s  N_ss=%3d N_ts=%3d  %3d instructions: %6lld cycles = %5.1f per instruction", names[mode],
         has_ss ? n_ss : 0, has_ts ? n_ts : 0, n_inst, tot, (double)tot / n_inst);
  if (mode >= MIX) printf(" = %5.1f per (ss + ts) pair", (double)tot / R);
  printf("   D = %.0f / %.0f %s\n", v[0], v[1], ok ? "ok" : "WRONG");
}

int main() {
  long long* dc;
  float* dv;
  cudaMalloc(&dc, 4 * sizeof(long long));
  cudaMalloc(&dv, 4 * sizeof(float));
  const int R = 96;
  for (int N : {32, 64, 96, 128, 160, 192, 256}) run(SS, N, 0, R, 0, dc, dv);
  for (int N : {32, 64, 128}) run(TS, 0, N, R, 0, dc, dv);
  for (int mode : {MIX, TWO, BLK}) {
    run(mode, 128, 64, R, 0, dc, dv);
    run(mode, 64, 32, R, 0, dc, dv);
    run(mode, 128, 32, R, 0, dc, dv);
  }
  for (int bg : {1, 2, 3}) {
    run(SS, 128, 0, R, bg, dc, dv);
    run(SS, 64, 0, R, bg, dc, dv);
    run(TS, 0, 64, R, bg, dc, dv);
    run(TS, 0, 32, R, bg, dc, dv);
    run(BLK, 128, 32, R, bg, dc, dv);
  }
  printf("done\n");
  return 0;
}
