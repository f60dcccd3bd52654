#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_join.py 2 > gpurun_out/ncu_launch.log 2>&1
EJB_CAND_CAP=33554432 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stage1|k_stage2a|k_pack_rows|k_forest_round|k_stage2b' -c 60 -o gpurun_out/full_r2 -f python tools/profile_join.py 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --cells 2000000 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "bench rc $?" >> gpurun_out/bench_dbg.err
tail -3 gpurun_out/ncu_launch.log; tail -3 gpurun_out/ncu_full.log; tail -5 gpurun_out/bench_dbg.err; head -c 1500 gpurun_out/bench_dbg.json
