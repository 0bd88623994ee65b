#!/usr/bin/env bash
# Round-2 GPU pass Q: the register-tiled dense defocus stencil -- K1 parity tests, then the stencil timing table.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "k1_" > gpurun_out/q_k1_tests.log 2>&1; echo "k1 tests exit $? : $(tail -n 1 gpurun_out/q_k1_tests.log)"
timeout 300 python tools/k1_stencil_bench.py > gpurun_out/q_stencil_bench.txt 2>&1; echo "bench exit $?"; cat gpurun_out/q_stencil_bench.txt
