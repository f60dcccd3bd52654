#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 200 python tools/sweep_sched.py > gpurun_out/sweep3.json 2> gpurun_out/sweep3.err
timeout 420 python tools/sim_partition.py --world 8 --weights 40 > gpurun_out/sim8c.json 2> gpurun_out/sim8c.err
tail -5 gpurun_out/pytest3.log
cat gpurun_out/sweep3.json
