#!/bin/bash
# round-2 first GPU contact: smoke under compute-sanitizer, parity tests, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc $?" >> gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/sanitizer.log; tail -30 gpurun_out/pytest_parity.log
