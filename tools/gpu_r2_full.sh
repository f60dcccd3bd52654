#!/bin/bash
# full pass: all GPU tests (parity + full-size), stage trace, bench N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "parity rc $?" >> gpurun_out/pytest_parity.log
EJB_TRACE=1 timeout 300 python tools/profile_join.py 4 > gpurun_out/trace1.log 2>&1; echo "rc $?" >> gpurun_out/trace1.log
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc $?" >> gpurun_out/pytest_fullsize.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?" >> gpurun_out/bench_n1.err
tail -4 gpurun_out/pytest_parity.log; grep "stage1\|finalize" gpurun_out/trace1.log | tail -7; tail -2 gpurun_out/trace1.log | head -1 | cut -c1-300; tail -6 gpurun_out/pytest_fullsize.log; tail -3 gpurun_out/bench_n1.err
