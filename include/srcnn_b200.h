/*
 * srcnn_b200.h -- C-ABI of the B200 (sm_100a) SRCNN device layer.
 *
 * This is the drop-in boundary: it replaces the reference's OpenCL runtime wrapper
 * (src/opencl/{Context,Kernel,UtilsOpenCL}.{hpp,cpp}) and its ten OpenCL kernels
 * (src/kernel/ *.cl).  The C++ host classes that keep the reference's API
 * (cnn-super-resolution_b200/host: DataPipeline, ConfigBasedDataPipeline, LayerData,
 * Config, `cnn` CLI) call ONLY these functions; so do tests/ (through ctypes) and bench.py.
 * Plain pointers and sizes only -- no C++ or torch types.  Every entry cites the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative SRCNN_E* code on failure; the message
 *     is available from srcnn_last_error() (thread-local).  The reference prints
 *     "[OPENCL ERROR] ..." and throws std::runtime_error (src/opencl/Context.cpp:111-123);
 *     the C++ host shim re-throws from the status code.
 *   - one context = one device + one in-order stream (reference: one in-order command queue,
 *     src/opencl/Context.cpp:70-73).  Not thread-safe, like the reference.
 *   - memory handles are indices into a context-owned table, never reused; the context owns
 *     all device memory until it is destroyed (src/opencl/Context.cpp:164-176, 98-101).
 *   - layouts: activations [S][H][W][C] (C fastest), weights [f][f][C_in][C_out] (C_out
 *     fastest), ground truth [S][H][W]  (src/kernel/layer_uber_kernel.cl:1-34).
 *   - all launches are asynchronous on the context's stream; srcnn_block() waits.
 *   - there is NO CPU fallback: without a CUDA device srcnn_ctx_create fails.
 */
#ifndef SRCNN_B200_H
#define SRCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCNN_OK 0
#define SRCNN_EINVAL (-1)   /* bad argument / failed validation                         */
#define SRCNN_ECUDA (-2)    /* a CUDA runtime call failed                               */
#define SRCNN_ENOMEM (-3)   /* device allocation failed                                 */
#define SRCNN_EHANDLE (-4)  /* unknown / released memory handle                         */
#define SRCNN_ERANGE (-5)   /* read/write/copy outside an allocation                    */

typedef struct srcnn_ctx srcnn_ctx; /* replaces opencl::Context (src/opencl/Context.hpp:72) */
typedef uint64_t srcnn_mem;         /* replaces opencl::MemoryHandle (Context.hpp:48)       */
/* replaces gpu_nullptr (src/DataPipeline.hpp:7) */
#define SRCNN_NULL_MEM ((srcnn_mem)1 << 30)

/* kernel ids, for per-kernel profiling totals (reference: one opencl::Kernel object per
 * .cl entry point, src/DataPipeline.cpp:121-180) */
enum srcnn_kernel_id {
  SRCNN_K_FORWARD = 0,       /* layer_uber_kernel.cl  `forward`            */
  SRCNN_K_SQUARED_ERR,       /* squared_error.cl      `squared_err`        */
  SRCNN_K_LAST_LAYER_DELTA,  /* last_layer_delta.cl   `last_layer_delta`   */
  SRCNN_K_DELTAS,            /* layer_deltas.cl       `deltas`             */
  SRCNN_K_BACKPROPAGATE,     /* backpropagate.cl      `backpropagate`      */
  SRCNN_K_UPDATE_PARAMS,     /* update_parameters.cl  `update_params`      */
  SRCNN_K_SUM,               /* sum.cl                `sum`                */
  SRCNN_K_SUB_FROM_ALL,      /* subtract_from_all.cl  `sub_from_all`       */
  SRCNN_K_EXTRACT_LUMA,      /* extract_luma.cl       `extract_luma`       */
  SRCNN_K_SWAP_LUMA,         /* swap_luma.cl          `swap_luma`          */
  SRCNN_K_FORWARD_FUSED,     /* (new) layers 1-3 in one launch, inference  */
  SRCNN_K_TRAIN_FUSED,       /* (new) fused pieces of the training step    */
  SRCNN_K_COUNT
};

/* ---------------------------------------------------------------- context ---------- */

/* Context::init(bool profile)  (src/opencl/Context.cpp:45-79).  `device` = CUDA ordinal.
 * `profile` != 0 times every launch with a cudaEvent pair and blocks on it, like
 * CL_QUEUE_PROFILING_ENABLE + clWaitForEvents (src/opencl/Kernel.cpp:108-116). */
int srcnn_ctx_create(int device, int profile, srcnn_ctx** out);
/* Same, but enqueue on a caller-owned cudaStream_t (e.g. torch's current stream) so that
 * launches order with the caller's own work (NCCL all-reduce of the gradient). */
int srcnn_ctx_create_on_stream(int device, void* cuda_stream, int profile, srcnn_ctx** out);
/* Context::~Context / _cleanup (src/opencl/Context.cpp:81-109): waits, frees everything. */
int srcnn_ctx_destroy(srcnn_ctx* ctx);
/* message of the last failed call on this thread */
const char* srcnn_last_error(void);
/* Context::block()  (src/opencl/Context.cpp:153-162): flush + finish. */
int srcnn_block(srcnn_ctx* ctx);
/* device name / SM count / global memory, for the "DEVICE:" line (Context.cpp:62) */
int srcnn_device_info(srcnn_ctx* ctx, char* name, size_t name_len, int* sm_count,
                      size_t* global_mem_bytes);
/* accumulated device time (ns) and launch count of one kernel id while profiling
 * (Kernel::get_total_execution_time, src/opencl/Kernel.hpp:64-66) */
int srcnn_profile_get(srcnn_ctx* ctx, int kernel_id, uint64_t* total_ns, uint64_t* launches);
/* number of kernel launches this context has issued (all ids); bench.py's gpu_launches */
int srcnn_launch_count(srcnn_ctx* ctx, uint64_t* launches);
/* the cudaStream_t this context enqueues on */
int srcnn_stream(srcnn_ctx* ctx, void** cuda_stream);

/* ---------------------------------------------------------------- memory ----------- */

/* Context::allocate(flags,size)  (src/opencl/Context.cpp:164-176) */
int srcnn_alloc(srcnn_ctx* ctx, size_t bytes, srcnn_mem* out);
/* register caller-owned device memory (a torch tensor's data_ptr) under a handle; the
 * context never frees it */
int srcnn_wrap(srcnn_ctx* ctx, void* device_ptr, size_t bytes, srcnn_mem* out);
/* RawMemoryHandle::release()  (src/opencl/Context.cpp:26-33); the handle stays invalid */
int srcnn_release(srcnn_ctx* ctx, srcnn_mem mem);
/* RawMemoryHandle::size via Context::raw_memory  (src/opencl/Context.cpp:125-130) */
int srcnn_mem_size(srcnn_ctx* ctx, srcnn_mem mem, size_t* bytes);
int srcnn_mem_ptr(srcnn_ctx* ctx, srcnn_mem mem, void** device_ptr);
/* Context::print_app_memory_usage (src/opencl/Context.cpp:132-149): bytes currently held */
int srcnn_mem_usage(srcnn_ctx* ctx, size_t* buffer_bytes);
/* Context::write_buffer(handle, offset, size, src, block)  (Context.cpp:262-294) */
int srcnn_write(srcnn_ctx* ctx, srcnn_mem mem, size_t offset, size_t bytes, const void* src,
                int block);
/* Context::read_buffer(handle, offset, size, dst, block)  (Context.cpp:235-260) */
int srcnn_read(srcnn_ctx* ctx, srcnn_mem mem, size_t offset, size_t bytes, void* dst,
               int block);
/* Context::copy_buffer(src, dst, dst_offset)  (Context.cpp:312-341): copies ALL of src */
int srcnn_copy(srcnn_ctx* ctx, srcnn_mem src, srcnn_mem dst, size_t dst_offset);
/* partial device-to-device copy (new; used by the batched sample gather) */
int srcnn_copy_region(srcnn_ctx* ctx, srcnn_mem src, size_t src_offset, srcnn_mem dst,
                      size_t dst_offset, size_t bytes);
/* the sample gather of ConfigBasedDataPipeline::execute_batch
 * (src/ConfigBasedDataPipeline.cpp:149-161: one clEnqueueCopyBuffer per sample and tensor) as
 * ONE launch: `n` buffers of `bytes_each` bytes (a multiple of 4) land in consecutive slots of
 * dst */
int srcnn_gather(srcnn_ctx* ctx, const srcnn_mem* src, int n, size_t bytes_each, srcnn_mem dst);
/* Context::fill_float / zeros_float  (Context.cpp:296-310): whole buffer, on the device */
int srcnn_fill_float(srcnn_ctx* ctx, srcnn_mem mem, float value);
/* pinned host staging memory for the e2e path (new) */
int srcnn_host_alloc(size_t bytes, void** host_ptr);
int srcnn_host_free(void* host_ptr);

/* ---------------------------------------------------------------- kernels ---------- */

/* kernel `forward` + DataPipeline::execute_layer's launch
 * (src/kernel/layer_uber_kernel.cl:36-96, src/DataPipeline.cpp:392-409):
 *   out[s][y][x][n] = act(B[n] + sum_{dy,dx,k} W[dy][dx][k][n] * in[s][y+dy][x+dx][k])
 * k = PREVIOUS_FILTER_COUNT, n = CURRENT_FILTER_COUNT, f = F_SPATIAL_SIZE, skip_relu =
 * SKIP_RELU.  Sizes are validated against the allocations. */
int srcnn_forward_layer(srcnn_ctx* ctx, srcnn_mem in, srcnn_mem out, srcnn_mem W, srcnn_mem B,
                        int k, int n, int f, int skip_relu, int in_w, int in_h, int S);

/* kernel `squared_err` (src/kernel/squared_error.cl:36-92, DataPipeline.cpp:416-472):
 * *target = sum_{s,y,x} (algo[s][y][x] - gt[s][y+p][x+p])^2, p = (gt_w - algo_w)/2.
 * `target` is a 1-float buffer; it is zeroed first (DataPipeline.cpp:441-444).  The
 * reduction is deterministic (fixed-order two-stage), unlike the reference's CAS atomics. */
int srcnn_squared_error(srcnn_ctx* ctx, srcnn_mem gt, srcnn_mem algo, srcnn_mem target,
                        int gt_w, int gt_h, int algo_w, int algo_h, int S);

/* kernel `last_layer_delta` (src/kernel/last_layer_delta.cl:14-50):
 * target[s][y][x] = (y - t) * [y > 0]   (quirk Q2 kept). */
int srcnn_last_layer_delta(srcnn_ctx* ctx, srcnn_mem gt, srcnn_mem algo, srcnn_mem target,
                           int gt_w, int gt_h, int algo_w, int algo_h, int S);

/* kernel `deltas` (src/kernel/layer_deltas.cl:42-127, DataPipeline.cpp:522-594):
 * target[s][j][i][n] = [layer_output>0] * sum_{dy,dx,k} W[dy][dx][n][k]*deltas_next[s][j-dy][i-dx][k]
 * n_curr = CURRENT_FILTER_COUNT (layer l-1), f_next/n_next = layer l, out_w/out_h = extent of
 * layer l-1's output. */
int srcnn_deltas(srcnn_ctx* ctx, srcnn_mem deltas_next, srcnn_mem layer_output,
                 srcnn_mem target, srcnn_mem W, int n_curr, int f_next, int n_next, int out_w,
                 int out_h, int S);

/* kernel `backpropagate` (src/kernel/backpropagate.cl:56-114, DataPipeline.cpp:596-663):
 * grad_w[dy][dx][k][n] += sum_{s,row,col} deltas[s][row][col][n]*layer_input[s][row+dy][col+dx][k]
 * grad_b[n]            += sum deltas[..][n].   ACCUMULATES; deterministic order. */
int srcnn_backpropagate(srcnn_ctx* ctx, srcnn_mem deltas, srcnn_mem layer_input,
                        srcnn_mem grad_w, srcnn_mem grad_b, int n, int k, int f, int out_w,
                        int out_h, int S);

/* kernel `update_params` (src/kernel/update_parameters.cl:1-33, DataPipeline.cpp:665-729),
 * quirk Q3 kept: dw = m*prev + lr*g + decay*w; w -= dw/batch; prev = dw (bias: no decay). */
int srcnn_update_params(srcnn_ctx* ctx, srcnn_mem w, srcnn_mem b, srcnn_mem grad_w,
                        srcnn_mem grad_b, srcnn_mem prev_dw, srcnn_mem prev_db, float momentum,
                        float weight_decay, float learning_rate, unsigned batch_size,
                        unsigned weights_size, unsigned bias_size);

/* kernel `sum` (src/kernel/sum.cl:35-68, DataPipeline.cpp:282-313): *target = sum data[i]
 * or sum data[i]^2 over the first `len` floats; target is zeroed first. */
int srcnn_sum(srcnn_ctx* ctx, srcnn_mem data, unsigned len, int squared, srcnn_mem target);

/* kernel `sub_from_all` (src/kernel/subtract_from_all.cl:1-8) */
int srcnn_sub_from_all(srcnn_ctx* ctx, srcnn_mem data, float value, unsigned len);

/* kernel `extract_luma` (src/kernel/extract_luma.cl:7-23, DataPipeline.cpp:186-220):
 * rgba = w*h*4 bytes (RGBA8, the reference's image2d_t), target = w*h floats. */
int srcnn_extract_luma(srcnn_ctx* ctx, srcnn_mem rgba, srcnn_mem target, int w, int h,
                       int normalize);

/* kernel `swap_luma` (src/kernel/swap_luma.cl:18-69, DataPipeline.cpp:222-266):
 * target = gt_w*gt_h*3 bytes RGB8. */
int srcnn_swap_luma(srcnn_ctx* ctx, srcnn_mem rgba, srcnn_mem new_luma, srcnn_mem target,
                    int gt_w, int gt_h, int luma_w, int luma_h);

/* ---------------------------------------------------------------- fused hot path --- */

typedef struct srcnn_net {
  int n1, n2, f1, f2, f3;        /* Config n1,n2,f1,f2,f3 (src/Config.hpp:27-29)          */
  srcnn_mem w[3], b[3];          /* LayerAllocationPool::weights / bias                    */
  srcnn_mem grad_w[3], grad_b[3];/* accumulating_grad_w / _b   (DataPipeline.hpp:11-29)   */
  srcnn_mem prev_dw[3], prev_db[3]; /* previous_batch_delta_w / _b                         */
} srcnn_net;

/* ConfigBasedDataPipeline::forward(w,h,S) for INFERENCE
 * (src/ConfigBasedDataPipeline.cpp:200-241) as ONE launch: layers 1-3 fused, the n1- and
 * n2-channel maps live in shared memory only.  in = [S][h][w] luma, out = [S][h3][w3].
 * Falls back to three srcnn_forward_layer launches through `scratch1/scratch2` for shapes the
 * fused kernel is not instantiated for (pass SRCNN_NULL_MEM to forbid the fallback). */
int srcnn_forward_fused(srcnn_ctx* ctx, const srcnn_net* net, srcnn_mem in, srcnn_mem out,
                        int in_w, int in_h, int S, srcnn_mem scratch1, srcnn_mem scratch2);
/* 1 when srcnn_forward_fused has a fused instantiation for this shape */
int srcnn_forward_fused_supported(const srcnn_net* net);

/* Row-band inference on HOST buffers (the reference-facing e2e call): uploads input rows
 * [row0 - halo .. row1 + halo) of a [h][w] luma image, runs the fused forward, downloads
 * output rows [row0,row1) of the [h3][w3] result into host_out (which points at row row0).
 * halo = f1+f2+f3-3 input rows per band.  Used by bench.py e2e and by the
 * multi-GPU row-band sharding: rank g calls it with its own [row0,row1).
 * Bands of >= 256 output rows are cut into sub-bands whose upload, compute and download
 * overlap on separate streams; that pipeline is captured into a CUDA graph owned by the
 * context and replayed while the call's arguments (both host pointers, shape, band, network
 * buffers) repeat -- the buffers' CONTENTS are read afresh on every call.  Synchronous: the
 * result is in host_out on return.  Pinned host buffers give the PCIe rate; pageable ones work. */
int srcnn_infer_rows_host(srcnn_ctx* ctx, const srcnn_net* net, const float* host_in, int in_w,
                          int in_h, int out_row0, int out_row1, float* host_out);
/* The same call, NON-BLOCKING: the reference's read_buffer(..., block = false) followed by
 * Context::block() (src/opencl/Context.cpp:255-276, 120-126).  Returns once the work is queued;
 * host_out is complete (and host_in may be rewritten) after srcnn_block(ctx) or after the next
 * blocking device-layer call on the parameters.  Two calls can be in flight: consecutive calls
 * alternate between two lanes (stream + staging + graph each), so that for a STREAM of images
 * the download of image i overlaps the upload and the first launches of image i+1 -- the
 * throughput tends to max(H2D, compute, D2H) per image instead of one call's head + tail on top.
 * Calls in flight must not share host_out.  Pinned host buffers are required for the overlap.
 * A device-layer write to the network's parameters, or a call with other parameters, first waits
 * for the calls in flight; in profile mode (and with SRCNN_E2E_GRAPH=0) the call is blocking. */
int srcnn_infer_rows_host_async(srcnn_ctx* ctx, const srcnn_net* net, const float* host_in,
                                int in_w, int in_h, int out_row0, int out_row1, float* host_out);

/* ConfigBasedDataPipeline::forward(sample) for MANY frames (src/ConfigBasedDataPipeline.cpp:
 * 114-126 is called once per image by src/Main_cl.cpp:217-239): `n_frames` luma frames of
 * w x h floats in host memory -> n_frames results of w3 x h3 floats in host_out.  Groups of
 * frames go through a ring of device slots so that the upload of the next group and the download
 * of the previous one overlap the fused forward of the current one.  Synchronous.  Frames are
 * independent: a multi-GPU caller gives each rank its own slice of the frame set (SURVEY 8e,
 * BASELINE config C5). */
int srcnn_infer_frames_host(srcnn_ctx* ctx, const srcnn_net* net, const float* host_in, int w,
                            int h, int n_frames, float* host_out);

/* One training chunk on device-resident samples: forward (keeping out1/out2), last-layer
 * delta, deltas 2<-3 and 1<-2, the three weight/bias gradients accumulated into
 * net->grad_*  (= ConfigBasedDataPipeline::forward + ::backpropagate,
 * src/ConfigBasedDataPipeline.cpp:165-176,243-323).  `work` holds the activations and deltas
 * of the chunk; size it with srcnn_train_workspace_bytes(). */
int srcnn_train_chunk(srcnn_ctx* ctx, const srcnn_net* net, srcnn_mem in, srcnn_mem gt,
                      int w, int h, int S, srcnn_mem work);
size_t srcnn_train_workspace_bytes(const srcnn_net* net, int w, int h, int S);
/* The layer-1 deltas are an intermediate of the layer-1 gradients only
 * (ConfigBasedDataPipeline.cpp:265-270 -> 300-320).  For f2 = 1, n1 = 64, n2 = 32 networks the
 * training-chunk entries compute them inside the layer-1 gradient kernel and leave the d1
 * buffer untouched; returns 1 when the chunk entries write d1 for this network on this
 * context, 0 when they do not (SRCNN_D1_IMPL=separate in the environment restores the
 * separate launch; srcnn_deltas itself always materialises its target). */
int srcnn_train_materializes_d1(srcnn_ctx* ctx, const srcnn_net* net);

/* Same chunk, with the six activation / delta buffers owned by the caller -- the buffers
 * ConfigBasedDataPipeline allocates itself (_out_1.._out_3, _delta_1.._delta_3,
 * src/ConfigBasedDataPipeline.cpp:82-108), so that its forward() + backpropagate() pair for a
 * training chunk is one call.  Every buffer must start 16-byte aligned. */
int srcnn_train_chunk_buffers(srcnn_ctx* ctx, const srcnn_net* net, srcnn_mem in, srcnn_mem gt,
                              int w, int h, int S, srcnn_mem out1, srcnn_mem out2,
                              srcnn_mem out3, srcnn_mem d1, srcnn_mem d2, srcnn_mem d3);

/* ConfigBasedDataPipeline::execute_batch(backpropagate = true, ...)
 * (src/ConfigBasedDataPipeline.cpp:128-195) on HOST-resident samples: `n_samples` inputs and
 * ground truths of w x h floats each are cut into chunks of at most `chunk` samples; the upload
 * of chunk i+1 (copy stream, double-buffered staging owned by the context) overlaps the
 * forward + backward of chunk i (srcnn_train_chunk); the first chunk, whose upload nothing can
 * hide, is half a chunk when there is more than one chunk of samples.  Gradients accumulate in net->grad_*;
 * the caller all-reduces them (multi-GPU) and calls srcnn_update_all.  `work` must hold
 * srcnn_train_workspace_bytes(net, w, h, chunk).  Returns once the host buffers may be reused;
 * the last chunk may still be training on the context stream. */
int srcnn_train_chunks_host(srcnn_ctx* ctx, const srcnn_net* net, const float* host_in,
                            const float* host_gt, int w, int h, int n_samples, int chunk,
                            srcnn_mem work);

/* ConfigBasedDataPipeline::update_parameters (src/ConfigBasedDataPipeline.cpp:325-361) in
 * ONE launch: the three layers with lr[0..2], then the six accumulators are zeroed. */
int srcnn_update_all(srcnn_ctx* ctx, const srcnn_net* net, unsigned batch_size, float momentum,
                     float weight_decay, const float lr[3]);

/* Validation pass of execute_batch (src/ConfigBasedDataPipeline.cpp:177-187): forward +
 * squared error of S device-resident samples; the SSE lands in the 1-float `target`. */
int srcnn_validate_chunk(srcnn_ctx* ctx, const srcnn_net* net, srcnn_mem in, srcnn_mem gt,
                         int w, int h, int S, srcnn_mem work, srcnn_mem target);

/* The fused kernels cache a packed copy of a network's parameters per context while the six
 * parameter buffers are context-owned, were never handed out by srcnn_mem_ptr, and no call of
 * this layer has written them.  A buffer whose raw pointer was obtained with srcnn_mem_ptr is
 * never cached (anything may write through it).  Call this after writing parameters by any
 * other route (a kernel of your own on the context stream, a peer copy ...). */
int srcnn_invalidate_params(srcnn_ctx* ctx);

/* ---------------------------------------------------------------- multi-GPU -------- */
/* New (the reference is single-device: src/opencl/Context.cpp:52-60).  One context per GPU /
 * process; the communicator enqueues on the context stream, so collectives order with the
 * launches around them.  NCCL (libnccl.so.2) is loaded on first use.  Partitioning:
 * inference = row bands, no collective; training = patches split across ranks, replicated
 * parameters, ONE sum all-reduce of the gradient accumulators before each
 * srcnn_update_all(batch_size = GLOBAL sample count); validation = 1-float all-reduce of the
 * squared error (src/ConfigBasedDataPipeline.cpp:177-187, src/Main_cl.cpp:174-192). */
#define SRCNN_COMM_ID_BYTES 128
/* rank 0 creates the id and hands it to the other ranks (file, pipe, launcher ...) */
int srcnn_comm_unique_id(unsigned char id[SRCNN_COMM_ID_BYTES]);
/* collective over all ranks: joins this context to the communicator */
int srcnn_comm_init(srcnn_ctx* ctx, int rank, int world,
                    const unsigned char id[SRCNN_COMM_ID_BYTES]);
int srcnn_comm_destroy(srcnn_ctx* ctx);
/* rank / world of the context (0 / 1 without a communicator) */
int srcnn_comm_info(srcnn_ctx* ctx, int* rank, int* world);
/* in-place sum over ranks of `count` floats at buf[offset_floats..]; no-op on one rank */
int srcnn_allreduce_sum(srcnn_ctx* ctx, srcnn_mem buf, size_t offset_floats, size_t count);
/* in-place sum over ranks of the six gradient accumulators of `net`: one all-reduce when they
 * are consecutive views of one flat buffer, a grouped call otherwise */
int srcnn_allreduce_grads(srcnn_ctx* ctx, const srcnn_net* net);
/* the first `count` floats of `buf` on rank `root` replace everybody's (parameter sync) */
int srcnn_broadcast(srcnn_ctx* ctx, srcnn_mem buf, size_t count, int root);

#ifdef __cplusplus
}
#endif
#endif /* SRCNN_B200_H */
