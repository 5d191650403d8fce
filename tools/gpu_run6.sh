#!/bin/bash
# final multi-GPU session on an 8-GPU box: data-parallel correctness (product communicator, C++ CLI)
# and the bench at N = 2, 4, 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2y}
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_dp.log 2>&1; echo "pytest dp rc $?"
tail -6 gpurun_out/${T}_pytest_dp.log
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench N=$N rc $?"
  tail -2 gpurun_out/${T}_bench_n$N.err
done
python - <<PY
import json
for N in (2, 4, 8):
    for l in open("gpurun_out/${T}_bench_n%d.json" % N):
        try: d = json.loads(l)
        except Exception: continue
        print("N", d["n_gpus"], "C3", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stream", round(d["e2e"]["stream_of_images"]["value"]),
              "| C2", round(d["train"]["value"]), "e2e", round(d["train"]["e2e"]["value"]),
              "| C4", round(d["train_c4"]["value"]), "e2e", round(d["train_c4"]["e2e"]["value"]),
              "| C5", round(d["c5"]["value"]), "e2e", round(d["c5"]["e2e"]["value"]),
              "| pcie", round(d["e2e"]["pcie"]["h2d_gbs_measured"], 1), round(d["e2e"]["pcie"].get("h2d_gbs_both_directions_busy") or 0, 1))
PY
