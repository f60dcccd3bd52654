#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic_config or tiny_work or sharded or row_range or randomized" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_sub.log
timeout 600 python tools/sweep_sched.py 1 > gpurun_out/sweep1.log 2>&1; cat gpurun_out/sweep1.log | tail -9
timeout 900 python tools/sweep_sched.py 4 > gpurun_out/sweep4.log 2>&1; cat gpurun_out/sweep4.log | tail -9
