#!/bin/bash
# bench line + ncu captures of configs[1]; the .ncu-rep is exported to CSV on the box and removed (gpurun brings back <= 64 MiB)
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?" >> gpurun_out/bench_n1.err
timeout 300 python tools/profile_join.py 2 1 > gpurun_out/profile_plain.log 2>&1; prc=$?; echo "plain rc $prc" >> gpurun_out/profile_plain.log
if [ $prc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_join.py 2 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc $?"
  EJB_CAND_CAP=33554432 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stage1|k_stage2a|k_pack_rows8|k_forest_round|k_stage2b|k_p1_terms' -c 70 -o /tmp/full_r2 -f python tools/profile_join.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
  ncu -i /tmp/full_r2.ncu-rep --page raw --csv > gpurun_out/full_r2_raw.csv 2> gpurun_out/ncu_export.err
  ncu -i /tmp/full_r2.ncu-rep --page source --csv -k regex:'k_stage1' -c 40 2>> gpurun_out/ncu_export.err | gzip -9 > gpurun_out/full_r2_source_stage1.csv.gz
  ncu -i /tmp/full_r2.ncu-rep --page source --csv -k regex:'k_stage2a' -c 6 2>> gpurun_out/ncu_export.err | gzip -9 > gpurun_out/full_r2_source_stage2a.csv.gz
fi
tail -2 gpurun_out/bench_n1.err; ls -la gpurun_out; du -sh gpurun_out
