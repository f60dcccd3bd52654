#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
timeout 200 python tools/ab_modes.py > gpurun_out/ab_modes.json 2> gpurun_out/ab_modes.err
timeout 200 python tools/sweep_sched.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
timeout 420 python tools/sim_partition.py --world 8 --weights 10,0,30 > gpurun_out/sim8b.json 2> gpurun_out/sim8b.err
tail -3 gpurun_out/pytest2.log
cat gpurun_out/ab_modes.json gpurun_out/ab_modes.err gpurun_out/sweep.json
