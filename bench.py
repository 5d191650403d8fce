#!/usr/bin/env python3
"""bench.py -- SRCNN 9-1-5 (n1=64, n2=32) hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Workloads (BASELINE.json):
  * primary  : INFERENCE of one 4096x4096 luma image (config C3), halo row-sharded across the
               N ranks (no collective; strong scaling).  A step = one forward pass of the
               whole image.  metric = input MPix/s = 4096*4096/1e6 / step time.
  * secondary: TRAINING, one epoch over 4096 synthetic 33x33 patches per rank (config C2,
               data-parallel, weak scaling): forward + backward + ONE all-reduce of the flat
               8129-float gradient + momentum/weight-decay update.  Reported under "train".
`value` is measured with inputs resident in HBM; `e2e` goes through the host-buffer entry
points (srcnn_infer_rows_host / write+train+read) with H2D and D2H inside the timed region.
`--impl reference` times the reference's own kernels (oracle/_ref, or the oracle port when
that was not built) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N1, N2, F1, F2, F3 = 64, 32, 9, 1, 5
IMG = 4096
PATCH = 33
PATCHES_PER_RANK = 4096
PAD = F1 + F2 + F3 - 3
MOMENTUM, DECAY = 0.9, 0.001
LR = np.array([1e-4, 1e-4, 1e-5], np.float32)   # example_config.json:8-10

# algorithmic work (SURVEY 8d / BASELINE.md section 2)
FWD_FLOP_C3 = 2 * ((IMG - 8) ** 2 * 81 * 64 + (IMG - 8) ** 2 * 64 * 32 + (IMG - 12) ** 2 * 25 * 32)
FUSED_BYTES_C3 = 4 * (IMG * IMG + (IMG - PAD) ** 2)
TRAIN_FLOP_PER_PATCH = 23.05e6
# dram__bytes_read.sum + dram__bytes_write.sum of one fused launch on C3 (ncu --set full,
# profiles/r1s_fused_hp_ncu_summary.txt): 67.2 MB read (the input, once) + 28.5 MB written back
# during the launch (the rest of the 66.7 MB output is still dirty in the 126 MB L2 at exit)
TRAFFIC_NCU_BYTES = 95.7e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_inputs(seed=1234):
    from helpers import luma_image, make_params, patches
    rng = np.random.default_rng(seed)
    params = make_params(rng, N1, N2, F1, F2, F3)
    return rng, params, luma_image, patches


# ======================================================================= reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle.loader import NetState, Oracle, have
    kind = "reference" if have("reference") else "port"
    orc = Oracle(kind)
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    orc.set_num_threads(len(os.sched_getaffinity(0)))
    cores = orc.num_threads()
    rng, params, luma_image, patches = synthetic_inputs()
    net = NetState(N1, N2, F1, F2, F3, params)
    # bounded sample of C3: a band of `rows` output rows of the 4096-wide image
    rows = args.ref_rows
    band = luma_image(rng, rows + PAD, IMG)
    t = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.net_forward(net, band, IMG, rows + PAD, 1)
        t.append(time.perf_counter() - t0)
    t = t[args.warmup:]
    sec = float(np.mean(t))
    frac = rows / float(IMG - PAD)
    mpix = IMG * IMG * frac / 1e6 / sec
    # training sample: `ref_patches` patches of C2, one epoch incl. the update
    x, gt = patches(rng, args.ref_patches, PATCH, PATCH)
    tt = []
    for i in range(2):
        t0 = time.perf_counter()
        orc.net_train_epoch(net, x, gt, PATCH, PATCH, args.ref_patches, MOMENTUM, DECAY, LR)
        tt.append(time.perf_counter() - t0)
    pps = args.ref_patches / min(tt)
    sample = ("inference: %d of %d output rows of the 4096-wide image per step (%.1f%% of C3), "
              "scaled linearly; train: %d of 4096 patches" % (rows, IMG - PAD, 100 * frac,
                                                              args.ref_patches))
    line = {
        "impl": "reference", "metric": "srcnn_915_inference_mpix_per_s", "value": mpix,
        "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / frac, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": mpix, "unit": "MPix/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": mpix, "unit": "MPix/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "train": {"metric": "srcnn_915_train_patches_per_s", "value": pps, "unit": "patches/s"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "C3: SRCNN 9-1-5 n1=64 n2=32 inference, one 4096x4096 luma image, "
                        "halo row bands over %d GPU(s); secondary C2: training, 4096 patches "
                        "33x33 per GPU, momentum+weight decay" % n,
            "net": "9-1-5 n1=64 n2=32", "image": [IMG, IMG], "patch": [PATCH, PATCH],
            "patches_per_gpu": PATCHES_PER_RANK, "halo_rows": PAD,
            "parallelism": "inference: row bands, no collective; training: dp%d, one all-reduce "
                           "of 8129 floats per update" % n,
            "l2": "inference rotates 4 input/output buffer pairs (>= 4x the band size) so no "
                  "step finds its input in the 126 MB L2; training streams > 1 GB of "
                  "activations per epoch"}


# ======================================================================= our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import _pkg
    pkg = _pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    hbm_peak, bf16_peak, peak_kind = peaks()
    rng, params, luma_image, patches = synthetic_inputs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        ctx = pkg.Context(local_rank, stream=stream.cuda_stream)
        grad = torch.zeros(sum(f * f * k * n + n for (k, n, f) in
                               [(1, N1, F1), (N1, N2, F2), (N2, 1, F3)]),
                           device="cuda", dtype=torch.float32)
        gh = ctx.wrap(grad.data_ptr(), grad.numel() * 4)
        net = pkg.Net(ctx, N1, N2, F1, F2, F3, params, grad_flat=gh)

        # ---------------------------------------------------------------- inference
        h3 = IMG - PAD
        w3 = IMG - PAD
        r0, r1 = pkg.row_bands(h3, world)[rank]
        band_h = r1 - r0 + PAD
        R = 4
        img = pkg.PinnedBuffer((IMG, IMG))
        img.array[:] = luma_image(rng, IMG, IMG)
        out_host = pkg.PinnedBuffer((r1 - r0, w3))
        ins, outs = [], []
        for i in range(R):
            m = ctx.alloc(4 * band_h * IMG)
            ctx.write(m, np.roll(img.array[r0:r0 + band_h], i, axis=1))
            ins.append(m)
            outs.append(ctx.alloc(4 * (r1 - r0) * w3))
        fused = net.fused_supported()
        (w1, h1), (w2, h2), _ = net.out_dims(IMG, band_h)
        s1 = s2 = pkg.NULL_MEM
        if not fused:   # three-launch path needs the n1/n2-channel maps in HBM
            s1, s2 = ctx.alloc(4 * w1 * h1 * N1), ctx.alloc(4 * w2 * h2 * N2)

        def infer_step(i):
            net.forward_fused(ins[i % R], outs[i % R], IMG, band_h, 1, s1, s2)

        def infer_e2e_step(i):
            if fused:
                net.infer_rows_host(img.array, IMG, IMG, r0, r1, out_host.array)
            else:
                ctx.write(ins[i % R], img.array[r0:r0 + band_h], block=False)
                net.forward_fused(ins[i % R], outs[i % R], IMG, band_h, 1, s1, s2)
                ctx.L.srcnn_read(ctx.h, outs[i % R], 0, out_host.nbytes, out_host.ptr, 1)

        def timed(step_fn, steps, warmup):
            for i in range(warmup):
                step_fn(i)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = ctx.launch_count()
            e0.record(stream)
            for i in range(steps):
                step_fn(warmup + i)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            return max_over_ranks(ms) / steps, ctx.launch_count() - n0

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        inf_ms, inf_launches = timed(infer_step, args.steps, args.warmup)
        e2e_ms, _ = timed(infer_e2e_step, max(2, args.steps // 2), 2)

        # ---------------------------------------------------------------- training
        n_loc = PATCHES_PER_RANK
        px = pkg.PinnedBuffer((n_loc, PATCH, PATCH))
        pg = pkg.PinnedBuffer((n_loc, PATCH, PATCH))
        rng_r = np.random.default_rng(99 + rank)
        px.array[:], pg.array[:] = patches(rng_r, n_loc, PATCH, PATCH)
        chunk = min(args.chunk, n_loc)
        d_in, d_gt = ctx.alloc(px.nbytes), ctx.alloc(pg.nbytes)
        ctx.write(d_in, px.array)
        ctx.write(d_gt, pg.array)
        work = ctx.alloc(net.train_workspace_bytes(PATCH, PATCH, chunk))
        bytes_per = 4 * PATCH * PATCH
        chunks = []
        for i in range(0, n_loc, chunk):
            S = min(chunk, n_loc - i)
            base_i, base_g = ctx.mem_ptr(d_in), ctx.mem_ptr(d_gt)
            chunks.append((ctx.wrap(base_i + i * bytes_per, S * bytes_per),
                           ctx.wrap(base_g + i * bytes_per, S * bytes_per), S))
        params_host = pkg.PinnedBuffer((grad.numel(),))

        def train_step(i):
            for (ci, cg, S) in chunks:
                net.train_chunk(ci, cg, PATCH, PATCH, S, work)
            if world > 1:
                dist.all_reduce(grad)       # the ONE exchange step of the path (sum)
            net.update_all(n_loc * world, MOMENTUM, DECAY, LR)

        def train_e2e_step(i):
            # HOST samples in (pinned), upload of chunk i+1 overlapping the training of chunk i
            net.train_chunks_host(px.array, pg.array, PATCH, PATCH, chunk, work)
            if world > 1:
                dist.all_reduce(grad)
            net.update_all(n_loc * world, MOMENTUM, DECAY, LR)
            off = 0
            for l in range(3):     # read the updated parameters back (the step's result)
                for hnd, cnt in ((net.c.w[l], net.sizes[l][0]), (net.c.b[l], net.sizes[l][1])):
                    ctx.L.srcnn_read(ctx.h, hnd, 0, 4 * cnt, params_host.ptr + 4 * off, 0)
                    off += cnt
            ctx.block()

        tr_steps = max(2, args.steps // 2)
        tr_ms, tr_launches = timed(train_step, tr_steps, args.warmup)
        tr_e2e_ms, _ = timed(train_e2e_step, tr_steps, 2)
        clocks = sampler.stop() if rank == 0 else None
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    mpix = IMG * IMG / 1e6 / (inf_ms / 1e3)
    e2e_mpix = IMG * IMG / 1e6 / (e2e_ms / 1e3)
    pps = n_loc * world / (tr_ms / 1e3)
    e2e_pps = n_loc * world / (tr_e2e_ms / 1e3)
    # roofline of the dominant kernel of the primary workload.  One rank's launch processes
    # 1/world of the image; per-launch duration = inf_ms (one fused launch per step) when the
    # fused kernel runs, else the step is three launches and the figure is for the whole step.
    frac_img = (r1 - r0) / float(h3)
    alg_bytes = FUSED_BYTES_C3 * frac_img
    alg_flop = FWD_FLOP_C3 * frac_img
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_clock * 1e6 / 1e12
    achieved_gbs = alg_bytes / (inf_ms / 1e3) / 1e9
    achieved_tf = alg_flop / (inf_ms / 1e3) / 1e12
    impl = os.environ.get("SRCNN_FUSED_IMPL", "hp")
    use_tc = fused and impl != "simt"
    # Dominant kernel of the primary workload = the fused forward launch (its duration is the
    # step time measured above with CUDA events, minus two ~3 us helper launches).  It moves
    # 8 B/pixel, so it is compute-bound: all three layers run on the tensor cores (layer 3 as a
    # tap GEMM + 25-term gather) with error-compensated split operands, because the 1e-4
    # tolerance rules out plain 16/19-bit inputs: every FP32 product costs THREE tensor-core
    # products -- FP16 halves at the bf16 rate (default kernel), or TF32 at half that rate.
    if use_tc:
        split = 6.0 if impl in ("pl", "ws", "tc") else 3.0
        roofline = {
            "kernel": ("forward_fused_hp_kernel (tcgen05 kind::f16, FP16-split operands, all "
                       "three layers)" if split == 3.0 else
                       "forward_fused_%s_kernel (tcgen05 3xTF32, all three layers)" % impl),
            "bound": "tensor", "achieved": achieved_tf, "peak": bf16_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / bf16_peak, "traffic": TRAFFIC_NCU_BYTES * frac_img,
            "peak_kind": peak_kind + " dense bf16 (MEASURED_PEAKS.json)",
            "algorithmic_flop_per_launch": alg_flop,
            "precision_ceiling_tflops": bf16_peak / split,
            "frac_of_precision_ceiling": achieved_tf / (bf16_peak / split),
            "note": "algorithmic FP32 FLOPs / launch time against the measured bf16 peak; the "
                    "split-operand scheme the tolerance requires issues 3 tensor-core products "
                    "per FP32 product (at the fp16/bf16 rate for the FP16-split kernel, at half "
                    "of it for 3xTF32), so peak/%d is the ceiling at this precision" % split,
            "hbm": {"achieved_gbs": achieved_gbs, "peak_gbs": hbm_peak,
                    "frac": achieved_gbs / hbm_peak,
                    "algorithmic_bytes_per_launch": alg_bytes},
        }
    else:
        roofline = {
            "kernel": "forward_fused_kernel (FP32 SIMT)" if fused else "forward x3 (unfused)",
            "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved_gbs / hbm_peak, "traffic": None, "peak_kind": peak_kind,
            "note": "fused inference moves only input+output luma (8 B/px): compute-bound, "
                    "see fp32",
            "fp32": {"achieved_tflops": achieved_tf, "peak_tflops": fp32_peak,
                     "frac": achieved_tf / fp32_peak,
                     "peak_kind": "148 SM x 128 lanes x 2 x median SM clock under load"},
        }
    line = {
        "metric": "srcnn_915_inference_mpix_per_s", "value": mpix, "unit": "MPix/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": inf_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_mpix, "unit": "MPix/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 4 * band_h * IMG, "d2h_bytes_per_step": 4 * (r1 - r0) * w3},
        "gpu_launches": int(inf_launches),
        "roofline": roofline,
        "clocks": clocks,
        "train": {"metric": "srcnn_915_train_patches_per_s", "value": pps, "unit": "patches/s",
                  "ms_per_step": tr_ms, "steps": tr_steps, "scaling": "weak",
                  "patches_per_step": n_loc * world, "chunk": chunk,
                  "gpu_launches": int(tr_launches),
                  "achieved_tflops": TRAIN_FLOP_PER_PATCH * pps / 1e12,
                  "e2e": {"value": e2e_pps, "unit": "patches/s", "ms_per_step": tr_e2e_ms,
                          "h2d_bytes_per_step": 2 * n_loc * bytes_per,
                          "d2h_bytes_per_step": 4 * grad.numel()}},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """The oracle/_ref (reference kernels on host cores) timed on a bounded sample."""
    from oracle.loader import NetState, Oracle, have
    kind = "reference" if have("reference") else "port"
    orc = Oracle(kind)
    orc.set_num_threads(len(os.sched_getaffinity(0)))
    rng, params, luma_image, patches = synthetic_inputs()
    net = NetState(N1, N2, F1, F2, F3, params)
    rows = args.ref_rows
    band = luma_image(rng, rows + PAD, IMG)
    orc.net_forward(net, band, IMG, rows + PAD, 1)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        orc.net_forward(net, band, IMG, rows + PAD, 1)
        t.append(time.perf_counter() - t0)
    frac = rows / float(IMG - PAD)
    mpix = IMG * IMG * frac / 1e6 / min(t)
    x, gt = patches(rng, args.ref_patches, PATCH, PATCH)
    t0 = time.perf_counter()
    orc.net_train_epoch(net, x, gt, PATCH, PATCH, args.ref_patches, MOMENTUM, DECAY, LR)
    pps = args.ref_patches / (time.perf_counter() - t0)
    return {"value": mpix, "unit": "MPix/s", "cores": orc.num_threads(), "kind": kind,
            "sample": "best of 3 on %d of %d output rows of the 4096-wide image (%.1f%% of C3), "
                      "scaled linearly" % (rows, IMG - PAD, 100 * frac),
            "train_patches_per_s": pps, "train_sample": "%d patches, 1 epoch" % args.ref_patches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=2048,
                    help="patches per training chunk (the reference chunks an epoch in two: "
                         "src/Main_cl.cpp:93,128-129)")
    ap.add_argument("--ref-rows", type=int, default=128)
    ap.add_argument("--ref-patches", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
