#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest9.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err
timeout 400 python tools/ab_lib.py default build/lib_p4_b3.so build/lib_p5_b4.so build/lib_S2A_MINB_2.so build/lib_S2A_MINB_4.so build/lib_S1_MINB2_3.so build/lib_S1_MINB2_5.so build/lib_S1_MINB3_2.so build/lib_S1_MINB3_4.so build/lib_S1_NOVOTE_1.so build/lib_S1_NOVOTE_1__DS1_MINB2_3.so default > gpurun_out/ab9.json 2> gpurun_out/ab9.err
tail -4 gpurun_out/pytest9.log
cut -c1-260 gpurun_out/ab9.json
tail -3 gpurun_out/ab9.err
python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
