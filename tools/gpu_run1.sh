#!/bin/bash
# round-2 GPU session A: parity tests, probes, bench line, ncu evidence (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2a}
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc $?"
timeout 60 tools/probe/mn16_probe > gpurun_out/${T}_mn16.log 2>&1; echo "probe rc $?"
cat gpurun_out/${T}_mn16.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
tail -3 gpurun_out/${T}_bench.err
timeout 300 python tools/profile_targets.py c3 c2 c4 c5 > gpurun_out/${T}_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'forward_fused_hp|bwd3_tc_kernel|wgrad1_fused_tc|wgrad2_tc|conv5_tc_kernel|wgrad5' \
  -o gpurun_out/${T}_prof -f python tools/profile_targets.py c3 c2 c4 c5 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu full rc $?"
timeout 300 python bench.py --steps 2 --warmup 3 --workloads c3,c2 --no-cpu-baseline > gpurun_out/${T}_plain2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --workloads c3,c2 --no-cpu-baseline > gpurun_out/${T}_ncu2.log 2>&1
echo "ncu list rc $?"
