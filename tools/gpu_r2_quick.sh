#!/bin/bash
# quick pass: parity tests + stage breakdown on configs[1], [3], [4]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_parity.log
EJB_TRACE=1 timeout 300 python tools/profile_join.py 4 > gpurun_out/trace1.log 2>&1; echo "rc $?" >> gpurun_out/trace1.log
timeout 300 python tools/profile_join.py 3 3 > gpurun_out/trace3.log 2>&1; echo "rc $?" >> gpurun_out/trace3.log
timeout 600 python tools/profile_join.py 3 4 > gpurun_out/trace4.log 2>&1; echo "rc $?" >> gpurun_out/trace4.log
tail -3 gpurun_out/pytest_parity.log; grep "stage1\|finalize" gpurun_out/trace1.log | tail -8; tail -3 gpurun_out/trace1.log | head -1 | cut -c1-330; tail -3 gpurun_out/trace3.log | head -1 | cut -c1-330;  tail -3 gpurun_out/trace4.log | head -1 | cut -c1-700
