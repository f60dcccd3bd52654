#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
EJB_TRACE=1 timeout 600 python tools/profile_join.py 2 4 > gpurun_out/trace4.log 2>&1; echo "rc $?" >> gpurun_out/trace4.log
grep "round" gpurun_out/trace4.log | tail -60
