#!/bin/bash
# memcheck of every tensor-core kernel family on small ragged cases (1 GPU, one sanitizer tool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2n}
timeout 300 python tools/sanitize_targets.py > gpurun_out/${T}_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${T}_memcheck.log \
  python tools/sanitize_targets.py > gpurun_out/${T}_sanitize_run.log 2>&1
echo "memcheck rc $?"
tail -5 gpurun_out/${T}_sanitize_plain.log
tail -15 gpurun_out/${T}_memcheck.log
