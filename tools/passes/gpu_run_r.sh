#!/usr/bin/env bash
# Round-2 GPU pass R: epilogue / sweep tests after the if-constexpr tidy of k34_small_kernel, ncu --set full of the stencil
# launches (defocus s1 / s3 / s5 and motion s3 at 8192 x 32x32 and 256 x 224x224).
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "k34 or epilogue or accumul or sweep or end_to_end or c_only or label" > gpurun_out/r_epilogue_tests.log 2>&1; echo "tests exit $? : $(tail -n 1 gpurun_out/r_epilogue_tests.log)"
timeout 600 ncu --set full --clock-control none -k regex:'k1_taps' -o /tmp/ncu/k1_taps -f python tools/k1_stencil_ncu.py > gpurun_out/r_ncu.log 2>&1; echo "ncu exit $?"
python tools/ncu_summary.py full /tmp/ncu/k1_taps.ncu-rep > gpurun_out/r_k1_taps_full.txt 2>&1; echo "summary $(wc -l < gpurun_out/r_k1_taps_full.txt) lines"
