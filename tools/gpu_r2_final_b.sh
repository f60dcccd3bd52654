#!/bin/bash
# second half of the final evidence pass (the first one ran every test, bench, reference arm and smoke: profiles/final_run_r2_tests.log;
# its gpurun_out/ was not copied back): bench line, reference arm line, then the ncu captures of configs[1]
cd "$(dirname "$0")/.."
rm -rf gpurun_out; mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?" >> gpurun_out/bench_n1.err
timeout 300 python tools/profile_join.py 2 1 > gpurun_out/profile_plain.log 2>&1; prc=$?; echo "plain rc $prc" >> gpurun_out/profile_plain.log
if [ $prc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_join.py 2 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc $?"
  EJB_CAND_CAP=33554432 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stage1|k_stage2a|k_pack_rows8|k_forest_round|k_stage2b|k_p1_terms' -c 70 -o gpurun_out/full_r2 -f python tools/profile_join.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
fi
tail -2 gpurun_out/bench_n1.err; ls -la gpurun_out
