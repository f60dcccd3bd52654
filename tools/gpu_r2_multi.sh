#!/bin/bash
# multi-GPU bench: tools/gpu_r2_multi.sh <N> [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/smi_n$N.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc $?" >> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err; head -c 3000 gpurun_out/bench_n$N.json
