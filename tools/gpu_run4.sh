#!/bin/bash
# multi-rank bench of selected workloads (N GPUs of one box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2l}; N=${2:-8}; W=${3:-c3}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 --workloads $W > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench N=$N rc $?"
tail -3 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_n$N.json"):
    try: d = json.loads(l)
    except Exception: continue
    p = d["e2e"]["pcie"]
    print("N", d["n_gpus"], "C3", round(d["value"]), "MPix/s", d["ms_per_step"], "ms | e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "ms | pinned copy GB/s", round(p["h2d_gbs_measured"], 1), round(p["d2h_gbs_measured"], 1), "numa", d["config"].get("numa_node_rank0"))
    if "c5" in d: print("C5", round(d["c5"]["value"]), "e2e", round(d["c5"]["e2e"]["value"]))
PY
lscpu | grep -i "numa\|socket" | head -8
for i in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi -i $i --query-gpu=pci.bus_id --format=csv,noheader | tr A-Z a-z | cut -c5-); echo "gpu $i $b numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)"; done
