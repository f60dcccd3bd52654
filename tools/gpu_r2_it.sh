#!/bin/bash
# one iteration on the GPU: parity suite, then per-phase traces of configs[1] and [4]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_parity.log
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 1 > gpurun_out/trace1.log 2>&1; echo "rc $?"
grep "prepare\|finalize\|round [0-9]*: stage1\|p1:" gpurun_out/trace1.log | tail -12; tail -3 gpurun_out/trace1.log | cut -c1-400
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 4 > gpurun_out/trace4.log 2>&1; echo "rc $?"
grep "prepare\|p1:" gpurun_out/trace4.log | tail -3; tail -3 gpurun_out/trace4.log | cut -c1-400
