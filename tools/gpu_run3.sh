#!/bin/bash
# round-2 GPU session: data-parallel correctness + multi-rank bench (N GPUs of one box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2e}; N=${2:-2}
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_dp.log 2>&1; echo "pytest dp rc $?"
tail -15 gpurun_out/${T}_pytest_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench N=$N rc $?"
tail -5 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_n$N.json"):
    try: d = json.loads(l)
    except Exception: continue
    print("N", d["n_gpus"], "C3", round(d["value"]), "MPix/s e2e", round(d["e2e"]["value"]),
          "| C2", round(d["train"]["value"]), "e2e", round(d["train"]["e2e"]["value"]),
          "| C4", round(d["train_c4"]["value"]), "e2e", round(d["train_c4"]["e2e"]["value"]),
          "| C5", round(d["c5"]["value"]), "e2e", round(d["c5"]["e2e"]["value"]))
PY
