#!/bin/bash
# final single-GPU numbers of the round: bench line, ncu launch list and `--set full` capture of the top kernels (each ncu
# run only after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
timeout 120 python tools/profile_join.py 2 > gpurun_out/profile_plain.log 2>&1; echo "plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python tools/profile_join.py 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc=$?"
EJB_CAND_CAP=33554432 timeout 420 ncu --set full --import-source on --clock-control none -k regex:'k_stage1|k_stage2a|k_pack_rows' -c 18 -o gpurun_out/full_final python tools/profile_join.py 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
cat gpurun_out/profile_plain.log | tail -1
python -c "
import json; d=json.load(open('gpurun_out/bench_final_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'], d['roofline']['frac'], d['roofline']['whole_join']['frac'], d['cpu_baseline']['value'])"
