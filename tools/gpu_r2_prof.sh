#!/bin/bash
# round-2 profile pass: stage breakdown, ncu launch list, ncu full set of the pair kernels, full-size parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
EJB_TRACE=1 timeout 300 python tools/profile_join.py 4 > gpurun_out/trace1.log 2>&1; echo "rc $?" >> gpurun_out/trace1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_join.py 2 > gpurun_out/ncu_launch.log 2>&1
EJB_CAND_CAP=33554432 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stage1|k_stage2a|k_pack_rows|k_forest_round|k_stage2b' -c 60 -o gpurun_out/full_r2 -f python tools/profile_join.py 1 > gpurun_out/ncu_full.log 2>&1
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_fullsize.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_fullsize.log
tail -12 gpurun_out/trace1.log; tail -3 gpurun_out/ncu_launch.log; tail -3 gpurun_out/ncu_full.log; tail -30 gpurun_out/pytest_fullsize.log
