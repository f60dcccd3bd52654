#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/trace_shards.py 8 4 1 > gpurun_out/shards8.json 2> gpurun_out/shards8.err; echo "shards rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4.csv python tools/profile_join.py 2 4 > gpurun_out/ncu_launch_c4.log 2>&1; echo "ncu rc $?"
cat gpurun_out/shards8.json
grep "prepare:" gpurun_out/shards8.err | head -3
