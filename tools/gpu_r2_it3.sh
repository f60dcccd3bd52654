#!/bin/bash
# iteration + launch list of configs[1]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_parity.log
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 1 > gpurun_out/trace1.log 2>&1; echo "rc $?"
grep "prepare\|finalize\|round [0-9]*: stage1\|p1:" gpurun_out/trace1.log | tail -12; tail -3 gpurun_out/trace1.log | cut -c1-400
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 4 > gpurun_out/trace4.log 2>&1; echo "rc $?"
grep "prepare\|p1:" gpurun_out/trace4.log | tail -3; tail -3 gpurun_out/trace4.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c1.csv python tools/profile_join.py 2 1 > gpurun_out/ncu_launch_c1.log 2>&1; echo "ncu rc $?"
