// Stand-alone probe: tcgen05.shift.down -- which way do the rows move, how many columns does one
// instruction move, does it cross the 32-lane quarters, what does it cost, and is it ordered
// behind MMAs issued by the same thread?  (The layer-3 gather of the fused SRCNN kernel sums
// Q[x + dx][tap] over dx: a row shift in tensor memory would replace its shared-memory staging.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o shift_probe shift_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../cnn-super-resolution_b200/csrc/tc_common.cuh"
using namespace srcnn::tc;

__device__ inline bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__device__ inline void shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}
__device__ inline void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ inline void tmem_ld16u(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// columns [64, 96): value = lane * 100 + column.  `nshift` shifts at column `col0`, then dump.
__global__ void __launch_bounds__(128) probe(int col0, int nshift, int reps, uint32_t* out,
                                             long long* cycles) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t mine = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 64; c0 < 96; c0 += 16) {
    uint32_t v[16];
    for (int j = 0; j < 16; j++) v[j] = (uint32_t)(tid * 100 + c0 + j);
    tmem_st16(mine + c0, v);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; r++)
        for (int s = 0; s < nshift; s++) shift_down(tmem + col0);
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0) cycles[0] = clock64() - t0;
  }
  __syncthreads();
  tcgen05_fence_after();
  for (int c0 = 64; c0 < 96; c0 += 16) {
    uint32_t r[16];
    tmem_ld16u(mine + c0, r);
    for (int j = 0; j < 16; j++) out[tid * 32 + (c0 - 64) + j] = r[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  uint32_t* d;
  long long* dc;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaMalloc(&dc, 8);
  static uint32_t h[128 * 32];
  for (int test = 0; test < 3; test++) {
    const int col0 = test == 2 ? 72 : 64, nshift = test == 1 ? 3 : 1;
    cudaMemset(d, 0, sizeof(h));
    probe<<<1, 128>>>(col0, nshift, 1, d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("--- %d shift(s) at column %d: for each column, which source lane does lane L now hold? (value/100)\n",
           nshift, col0);
    int moved_cols = 0;
    for (int c = 0; c < 32; c++) {
      // classify the column: offset k such that h[L][c] == (L - k) * 100 + 64 + c for most L
      int best_k = 99, best_n = -1;
      for (int k = -4; k <= 4; k++) {
        int n = 0;
        for (int L = 0; L < 128; L++)
          if (L - k >= 0 && L - k < 128 && h[L * 32 + c] == (uint32_t)((L - k) * 100 + 64 + c)) n++;
        if (n > best_n) { best_n = n; best_k = k; }
      }
      printf("col %2d: lane L holds old lane L-%d (%d of 128 lanes match)", 64 + c, best_k, best_n);
      if (best_k != 0) {
        moved_cols++;
        printf("; lanes 0..3 now hold: %u %u %u %u; lanes 31..34: %u %u %u %u", h[0 * 32 + c], h[1 * 32 + c],
               h[2 * 32 + c], h[3 * 32 + c], h[31 * 32 + c], h[32 * 32 + c], h[33 * 32 + c], h[34 * 32 + c]);
      }
      printf("\n");
    }
    printf("columns moved: %d\n", moved_cols);
  }
  // cost: many shifts back to back
  for (int reps : {16, 64}) {
    probe<<<1, 128>>>(64, 1, reps, d, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%d shifts back to back: %lld cycles = %.1f per shift\n", reps, c, (double)c / reps);
  }
  printf("done\n");
  return 0;
}
