#!/bin/bash
# Builds every native artefact in-tree: the CUDA join library (sm_100a), the synthetic generator and the CPU oracle.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-Wall -shared \
  -o enclone_ranger_b200/libenclone_join_b200.so enclone_ranger_b200/csrc/ejb_host.cu "$@"
g++ -O2 -std=c++17 -fPIC -shared -pthread -Wall -o enclone_ranger_b200/libejb_synth.so enclone_ranger_b200/csrc/synth.cpp
make -s -C oracle
