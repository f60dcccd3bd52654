#!/bin/bash
# A/B of build/lib_*.so variants (EJB_AB_LIBS): timings on configs[1] and [4], and a parity subset per variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LIBS=${EJB_AB_LIBS:-"build/lib_head.so build/lib_e.so"}
timeout 600 python tools/ab_lib.py $LIBS $LIBS > gpurun_out/ab1.json 2> gpurun_out/ab1.err; echo "ab1 rc $?"
cat gpurun_out/ab1.json
EJB_AB_CONFIG=4 timeout 900 python tools/ab_lib.py $LIBS > gpurun_out/ab4.json 2> gpurun_out/ab4.err; echo "ab4 rc $?"
cat gpurun_out/ab4.json
for lib in $LIBS; do
  case $lib in *head*) continue;; esac
  EJB_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or synthetic_config or randomized or packed_cdr3s or tags_do_not or sharded_contexts" > gpurun_out/pytest_ab.log 2>&1; echo "pytest $lib rc $?"; tail -2 gpurun_out/pytest_ab.log
done
