#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "ref rc $?"
cat gpurun_out/bench_ref.time; head -c 1500 gpurun_out/bench_ref.json
