#!/bin/bash
# round-2 GPU session B: the 9-5-5 tensor-core kernels (conv5_tc / wgrad5_tc)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 --timeout 240 \
  -k "conv5 or backpropagate_vs_oracle or baseline_sized or train_chunk_vs or inference_vs or float_alignment" > gpurun_out/${T}_pytest_c4.log 2>&1
echo "pytest(c4) rc $?"; tail -40 gpurun_out/${T}_pytest_c4.log
timeout 600 python bench.py --workloads c2,c4 --steps 20 --no-cpu-baseline > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
echo "bench c4 rc $?"; tail -3 gpurun_out/${T}_bench_c4.err; cat gpurun_out/${T}_bench_c4.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    for k in ('train', 'train_c4'):
        t = d.get(k, {})
        print(k, t.get('value'), 'patches/s', t.get('ms_per_step'), 'ms', t.get('kernels_one_chunk'))
"
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --timeout 300 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/profile_targets.py c2 c4 > gpurun_out/${T}_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'conv5_tc|wgrad5_tc|bwd3|wgrad1|wgrad2|n1_forward|forward_fused_hp|absmax|d3_kernel' \
  -o gpurun_out/${T}_prof -f python tools/profile_targets.py c2 c4 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu full rc $?"
