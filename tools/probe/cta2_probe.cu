// Stand-alone probe: what does one tcgen05.mma INSTRUCTION cost the tensor pipe, and does a CTA
// pair (cta_group::2, M = 256) pay it once for both SMs?  The fused SRCNN kernels are bound by
// their MMA instruction count (profiles/r2w_fused_kernel_sensitivity.txt), so these numbers say
// what a 2-CTA version of them could gain.  kind::f16, K = 16 per instruction, operands in
// shared memory (K-major, no swizzle), FP32 accumulators in tensor memory.
//   variants: cta_group 1 / 2;  N = 32, 64, 128;  R instructions into ONE accumulator (a dependent
//   chain) or round-robin over 4 accumulators;  one issuing thread or two (different warps).
// Every run checks the accumulator value, so a mis-encoded instruction cannot pass as "fast".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cta2_probe cta2_probe.cu
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
template <int CG>
__device__ inline void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
template <int CG>
__device__ inline void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
  else   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"((unsigned short)3)
        : "memory");
}
template <int CG>
__device__ inline void alloc512(uint32_t* slot) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ inline void dealloc512(uint32_t t) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}
__device__ inline void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ inline uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
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

// out[0] = cycles of the leader's issue loop + completion, out[1..2] = D[0][0] of accumulator 0 in
// CTA 0 / CTA 1
template <int CG>
__global__ void __launch_bounds__(128) probe(int N, int R, int nacc, int issuers, long long* cycles,
                                             float* dval) {
  __shared__ __align__(128) __half sA[128 * 16];
  __shared__ __align__(128) __half sB[256 * 16];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32;
  const uint32_t rank = CG == 2 ? cluster_rank() : 0;
  for (int i = tid; i < 128 * 16; i += 128) sA[i] = __float2half(1.f + (float)rank);
  for (int i = tid; i < 256 * 16; i += 128) sB[i] = __float2half(1.f);
  if (warp == 0) alloc512<CG>(&tmem_slot);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  fence_proxy_async();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint64_t ad = make_desc_kmajor(sA, 0, 128, 256);
  const uint64_t bd = make_desc_kmajor(sB, 0, 128, 256);
  const uint32_t idesc = idesc_f16(128 * CG, N);
  // issuing warps 0 .. issuers-1: the whole warp runs the loop converged and one elected lane
  // issues (a divergent `if (lane == 0)` makes ptxas wrap every UTCHMMA in its own election loop)
  const int who = tid / 32;
  if (rank == 0 && who < issuers) {
    const uint32_t d0 = tmem + (uint32_t)(who * nacc * N);
    const uint32_t step = nacc > 1 ? (uint32_t)N : 0u;
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < R; r += 4) {      // R is a multiple of 4, nacc is 1 or 4
        mma_f16<CG>(d0, ad, bd, idesc, r > 0);
        mma_f16<CG>(d0 + step, ad, bd, idesc, nacc > 1 ? r > 0 : 1);
        mma_f16<CG>(d0 + 2 * step, ad, bd, idesc, nacc > 1 ? r > 0 : 1);
        mma_f16<CG>(d0 + 3 * step, ad, bd, idesc, nacc > 1 ? r > 0 : 1);
      }
      commit<CG>(&bar[who]);
    }
    __syncwarp();
    mbar_wait(&bar[who], 0);
    if ((tid & 31) == 0) cycles[who] = clock64() - t0;
  }
  for (int w = 0; w < issuers; w++) mbar_wait(&bar[w], 0);
  tcgen05_fence_after();
  if (warp == 0) {
    float v;
    tmem_ld1(tmem, v);
    if (tid == 0) dval[rank] = v;
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 0) dealloc512<CG>(tmem);
}

template <int CG>
static void run(int N, int R, int nacc, int issuers, long long* dc, float* dv) {
  cudaMemset(dc, 0, 4 * sizeof(long long));
  cudaMemset(dv, 0, 2 * sizeof(float));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CG, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  long long best[2] = {1LL << 60, 1LL << 60};
  float v[2] = {0, 0};
  for (int rep = 0; rep < 3; rep++) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG>, N, R, nacc, issuers, dc, dv);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("cta_group::%d N=%3d: %s\n", CG, N, cudaGetErrorString(e));
      exit(1);
    }
    long long c[2];
    cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    cudaMemcpy(v, dv, sizeof(v), cudaMemcpyDeviceToHost);
    for (int w = 0; w < issuers; w++) best[w] = c[w] < best[w] ? c[w] : best[w];
  }
  const float expect = (float)(R / nacc) * 16.f;
  const bool ok = v[0] == expect && (CG == 1 || v[1] == 2.f * expect);
  printf("cta_group::%d M=%3d N=%3d  %3d MMAs, %d accumulator(s), %d issuer(s): %6.1f cycles per MMA"
         " per issuer (%lld total)%s  D = %.0f / %.0f (expected %.0f / %.0f) %s\n",
         CG, 128 * CG, N, R, nacc, issuers, (double)best[0] / R, best[0],
         issuers > 1 ? " [both issuers alike]" : "", v[0], v[1], expect, CG == 2 ? 2 * expect : 0.f,
         ok ? "ok" : "WRONG");
}

int main() {
  long long* dc;
  float* dv;
  cudaMalloc(&dc, 4 * sizeof(long long));
  cudaMalloc(&dv, 2 * sizeof(float));
  const int R = 96;
  for (int N : {32, 64, 128}) {
    run<1>(N, R, 1, 1, dc, dv);
    run<1>(N, R, 4, 1, dc, dv);
  }
  run<1>(32, R, 1, 2, dc, dv);     // two issuing threads, each its own accumulator
  run<1>(64, R, 1, 2, dc, dv);
  for (int N : {32, 64, 128, 256}) {
    run<2>(N, R, 1, 1, dc, dv);
    if (N <= 64) run<2>(N, R, 4, 1, dc, dv);
  }
  printf("done\n");
  return 0;
}
