#!/bin/bash
# device time of the remaining BASELINE configs at full size (parity of these sizes: tests/test_gpu_fullsize.py digests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 0 2 3; do timeout 300 python tools/profile_join.py 3 $c > gpurun_out/config$c.log 2>&1; echo "config $c rc $?"; tail -3 gpurun_out/config$c.log | head -1 | cut -c1-500; done
