#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in A B C; do
  EJB_LIB=$PWD/enclone_ranger_b200/lib_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_fixture or synthetic_config" > gpurun_out/bisect_$v.log 2>&1
  echo "variant $v rc $?"; tail -2 gpurun_out/bisect_$v.log
  EJB_LIB=$PWD/enclone_ranger_b200/lib_$v.so EJB_TRACE=1 timeout 300 python tools/profile_join.py 4 > gpurun_out/trace_$v.log 2>&1; grep "round 4 host\|round 3 host\|round 4: stage\|round 3: stage" gpurun_out/trace_$v.log | tail -4; tail -3 gpurun_out/trace_$v.log | head -1 | cut -c1-330
done
