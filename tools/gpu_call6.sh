#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest6.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err
timeout 200 python tools/sweep_sched.py > gpurun_out/sweep6.json 2> gpurun_out/sweep6.err
timeout 500 python tools/trace_parts.py 8 > gpurun_out/trace8c.json 2> gpurun_out/trace8c.err
tail -4 gpurun_out/pytest6.log
cut -c1-330 gpurun_out/sweep6.json
grep -v "round [0-9]: stage1" gpurun_out/trace8c.err | tail -24 | cut -c1-300
cut -c1-330 gpurun_out/trace8c.json
