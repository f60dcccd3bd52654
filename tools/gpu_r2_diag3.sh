#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_parity.log
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 1 > gpurun_out/trace1.log 2>&1; echo "rc $?"
grep "prepare\|finalize\|round [0-9]*: stage1\|p1:" gpurun_out/trace1.log | tail -12; tail -3 gpurun_out/trace1.log | cut -c1-400
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 3 > gpurun_out/trace3.log 2>&1; echo "rc $?"
grep "p1:" gpurun_out/trace3.log | tail -1;  tail -3 gpurun_out/trace3.log | cut -c1-400
EJB_TRACE=1 timeout 300 python tools/profile_join.py 3 4 > gpurun_out/trace4.log 2>&1; echo "rc $?"
grep "prepare\|p1:" gpurun_out/trace4.log | tail -3; tail -3 gpurun_out/trace4.log | cut -c1-400
timeout 900 python tools/trace_shards.py 8 4 1 > gpurun_out/shards8.json 2> gpurun_out/shards8.err; echo "shards rc $?"
cut -c1-200 gpurun_out/shards8.json
grep "prepare\|p1:" gpurun_out/shards8.err | head -6
