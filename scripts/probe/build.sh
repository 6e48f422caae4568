#!/bin/bash
# builds the NVLink read probe (measurement only) next to the repo so that it travels to the GPU box
set -e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_exp_probe
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC \
    scripts/probe/peer_probe.cu -o gpurun_exp_probe/libpeer_probe.so -cudart static
