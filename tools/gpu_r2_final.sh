#!/bin/bash
# final evidence pass of round 2 (one GPU): every GPU test, the bench line, the reference arm, smoke, then the ncu captures.
# Each ncu pass runs only after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "parity rc $?" >> gpurun_out/pytest_parity.log
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc $?" >> gpurun_out/pytest_fullsize.log
timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?" >> gpurun_out/bench_n1.err
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?" >> gpurun_out/bench_ref.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/smoke.log
timeout 300 python tools/profile_join.py 2 1 > gpurun_out/profile_plain.log 2>&1; prc=$?; echo "plain rc $prc" >> gpurun_out/profile_plain.log
if [ $prc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_join.py 2 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc $?"
  EJB_CAND_CAP=33554432 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stage1|k_stage2a|k_pack_rows8|k_forest_round|k_stage2b|k_p1_terms' -c 70 -o gpurun_out/full_r2 -f python tools/profile_join.py 1 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
fi
tail -3 gpurun_out/pytest_parity.log; tail -3 gpurun_out/pytest_fullsize.log; tail -2 gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_ref.err; tail -2 gpurun_out/smoke.log
head -c 1200 gpurun_out/bench_n1.json; echo; head -c 600 gpurun_out/bench_ref.json
