#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/sweep_shards.py 4 > gpurun_out/sweep_shards4.json 2> gpurun_out/sweep_shards4.err; echo "rc $?"
cat gpurun_out/sweep_shards4.json
timeout 600 python tools/sweep_shards.py 1 > gpurun_out/sweep_shards1.json 2> gpurun_out/sweep_shards1.err; echo "rc $?"
cat gpurun_out/sweep_shards1.json
EJB_TRACE=1 timeout 600 python tools/trace_shards.py 8 4 1 > gpurun_out/shards8.json 2> gpurun_out/shards8.err; echo "shards rc $?"
cut -c1-200 gpurun_out/shards8.json
grep "prepare" gpurun_out/shards8.err | head -2
