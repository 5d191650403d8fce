// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// Minimal OpenCL-C 1.1 -> C++ shim so the reference's own kernel sources
// (/root/reference/src/kernel/*.cl) compile UNMODIFIED with g++ and run on host
// cores.  The .cl files are #include'd from where they lie under /root/reference
// at build time (see oracle/Makefile, oracle/ref_kernels.cpp); nothing is copied
// into this repository.  The driver loops in ref_kernels.cpp play the role of
// clEnqueueNDRangeKernel (reference: src/opencl/Kernel.cpp:72-119).
#pragma once
#include <cstddef>
#include <cstdint>

// address-space / access qualifiers carry no meaning on the host
#define __kernel
#define __global
#define __local
#define __constant const
#define __read_only
#define __write_only
#define __const const

typedef unsigned int uint;
typedef unsigned char uchar;

struct int2 {
  int x, y;
};

namespace clshim {
// work-item ids of the "current" work item; one set per host thread
extern thread_local size_t gid[3];
}  // namespace clshim

static inline size_t get_global_id(uint d) { return clshim::gid[d]; }

static inline float max(float a, float b) { return a > b ? a : b; }

// One work item at a time touches a given address in our NDRange loops (samples
// are serialised, see ref_backpropagate), so a builtin CAS is more than enough.
static inline unsigned int atomic_cmpxchg(volatile unsigned int* p,
                                          unsigned int cmp, unsigned int val) {
  return __sync_val_compare_and_swap(p, cmp, val);
}
