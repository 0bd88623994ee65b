#!/usr/bin/env bash
# builds the two standalone hardware probes (nvcc cross-compiles for sm_100a without a GPU)
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcuda
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
