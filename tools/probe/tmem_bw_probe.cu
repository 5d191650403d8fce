// Stand-alone probe: tcgen05.ld / tcgen05.st bandwidth between tensor memory and registers,
// by shape and number of warps, with and without a concurrent stream of tcgen05.mma.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw_probe tmem_bw_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

template <int X>
__device__ __forceinline__ void ld_x(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld_x<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld_x<32>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
// 16x256b.x4: a warp reads 16 lanes x (4 x 256 bits) = 16 registers per thread
__device__ __forceinline__ void ld_16x256_x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void st_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
               "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
               "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// mode 0: ld x16, 1: ld x32, 2: st x8, 3: st x32, 4: ld x32 + st x32 alternating
// n_warps warps (warp w -> lane quarter w&3); mma != 0: warp 15 issues TS N=64 MMAs meanwhile
__global__ void __launch_bounds__(512) tmem_bw_kernel(int mode, int n_warps, int mma, int reps, long long* out,
                                                      uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* s = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < 100 * 1024 / 4; i += 512) s[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); stop = 0; }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t acc = 0;
  if (n_warps == 0 && warp == 0) {
    long long t0 = clock64();
    while (clock64() - t0 < 400000) {}
    if (lane == 0) { out[0] = clock64() - t0; stop = 1; }
  } else if (warp < n_warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[32];
    for (int i = 0; i < 32; i++) r[i] = tid + i;
    long long t0 = clock64();
    for (int it = 0; it < reps; it++) {
      if (mode == 0) { ld_x<16>(base, r); ld_x<16>(base + 16, r + 16); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
      else if (mode == 1) { ld_x<32>(base, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
      else if (mode == 2) { st_x8(base, r); st_x8(base + 8, r + 8); st_x8(base + 16, r + 16); st_x8(base + 24, r + 24);
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
      else if (mode == 3) { st_x32(base, r); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
      else if (mode == 4) { ld_x<32>(base, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
             st_x32(base + 32, r); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
      else { ld_16x256_x4(base, r); ld_16x256_x4(base + ((uint32_t)16 << 16), r + 16);
             asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
      acc += r[0] ^ r[7] ^ r[16] ^ r[31];
    }
    long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (warp == 0 && lane == 0) stop = 1;
  } else if (warp >= 14 && mma) {
    // two issuers of SS N=128 MMAs (pipe-bound at ~65 cycles per MMA when undisturbed)
    if (lane == 0) {
      const uint32_t id128 = make_idesc_tf32(128, 128);
      const uint64_t ad = make_desc_kmajor(s, 0, 16, 128), bd = make_desc_kmajor(s, 16384, 128, 2816);
      const uint32_t d = tmem + 256 + 128 * (warp - 14);
      long long n = 0;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 8; i++) mma_tf32(d, ad + 2 * i, bd + 16 * i, id128, 1);
        n += 8;
      }
      out[warp] = n;
    }
  }
  if (acc == 0x12345) sink[tid] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 16 * 8);
  cudaMalloc(&sink, 512 * 4);
  cudaFuncSetAttribute(tmem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[6] = {"ld 2 x x16", "ld x32", "st 4 x x8", "st x32", "ld x32 + st x32", "ld 2 x 16x256b.x4"};
  const int reps = 2000;
  for (int mma = 1; mma < 2; mma++)
    for (int mode = 0; mode < 6; mode++)
      for (int nw : {0, 4, 8}) {
        cudaMemset(d, 0, 128);
        tmem_bw_kernel<<<1, 512, 100 * 1024>>>(mode, nw, mma, reps, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16];
        cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < (nw ? nw : 1); i++) mx = h[i] > mx ? h[i] : mx;
        const double bytes = (double)reps * nw * 32 * 32 * 4 * (mode == 4 ? 2 : 1);
        printf("%s %-16s %d warps: %7.1f cyc per 32-col access/warp, %6.1f B/cyc/SM%s\n", mma ? "with MMA" : "no MMA  ",
               names[mode], nw, (double)mx / reps, bytes / mx, mma ? "" : "");
        if (mma) printf("      (%lld MMAs issued meanwhile = %.1f cyc/MMA of the pipe)\n", h[14] + h[15], (h[14] + h[15]) ? (double)mx / (h[14] + h[15]) : 0.0);
      }
  return 0;
}
