#!/usr/bin/env bash
# Runs the GPU parity groups as separate processes (a sticky CUDA error in one group must not mask the others),
# then a short bench.  Usage on the GPU box:  bash tools/gpu_check.sh [groups...]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest -k expr
  timeout "$2" python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$3" -p no:cacheprovider > "gpurun_out/t_$1.log" 2>&1
  echo "== $1: exit $? : $(tail -n 1 gpurun_out/t_$1.log)"
}
groups=${*:-"k1 k34 conv fwd bench"}
for g in $groups; do
  case $g in
    k1)   run k1 600 "philox or k1" ;;
    k34)  run k34 300 "k3 or k4 or k34 or frame_stats or gate_reproduces or allreduce" ;;
    conv) run conv 300 "conv" ;;
    fwd)  run fwd 600 "forward or cell or sweep_partition or gate_with or full_size or trust_replay or camera_shape or imagenet_shape or smoke_size" ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $? : $(tail -n 1 gpurun_out/smoke.log)" ;;
    bench) timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench: exit $? : $(tail -c 1500 gpurun_out/bench.log)" ;;
  esac
done
