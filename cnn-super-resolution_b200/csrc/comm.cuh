// Multi-GPU plumbing of the device layer: one process (or thread) per GPU, each with its own
// srcnn_ctx, joined by an NCCL communicator that lives in the context and enqueues on the
// context's stream.  The reference has nothing distributed; what the path needs
// is ONE sum all-reduce of the gradient accumulators per update_parameters call
// (src/ConfigBasedDataPipeline.cpp:325-361: the update consumes the gradient of the whole
// training set) and a 1-float all-reduce of the validation squared error
// (src/ConfigBasedDataPipeline.cpp:177-187, src/Main_cl.cpp:174-192).
//
// NCCL is loaded at run time (dlopen of libnccl.so.2) so that the device library has no link-time
// dependency on it: single-GPU users never touch it, and inside a torch process the copy torch
// already loaded is the one that resolves.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include "context.cuh"

namespace srcnn {
namespace comm {

// the slice of the NCCL ABI this layer uses (stable since NCCL 2.0: nccl.h)
using ncclComm_t = void*;
struct UniqueId {
  char internal[128];
};
constexpr int kNcclSuccess = 0, kNcclFloat = 7, kNcclSum = 0;

struct Api {
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

inline Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (a.lib) {
      auto sym = [&](const char* s) { return dlsym(a.lib, s); };
      a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
      a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
      a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
      a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
      a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
      a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
      a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
      a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
      if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.Broadcast ||
          !a.GroupStart || !a.GroupEnd)
        a.lib = nullptr;
    }
  }
  return a.lib ? &a : nullptr;
}

inline int nccl_fail(const char* what, int rc) {
  Api* a = api();
  return fail(SRCNN_ECUDA, "NCCL error in %s: %s (%d)", what,
              (a && a->GetErrorString) ? a->GetErrorString(rc) : "?", rc);
}

#define SRCNN_NCCL(call, what)                                   \
  do {                                                           \
    int _r = (call);                                             \
    if (_r != ::srcnn::comm::kNcclSuccess) return ::srcnn::comm::nccl_fail(what, _r); \
  } while (0)

}  // namespace comm
}  // namespace srcnn
