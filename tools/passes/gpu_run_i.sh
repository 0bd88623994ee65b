#!/usr/bin/env bash
# Round-2 GPU pass I (2-GPU box, final code): full GPU suite including the real two-GPU bit-identity test, the driver's 2-rank
# bench command (step mode, weak scaling), a 2-rank strong-scaled C3 sweep with the pipelined path.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/gputest_i.log 2>&1
echo "== pytest exit $? : $(tail -n 1 gpurun_out/gputest_i.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_i.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_c2_2gpu_i.json 2> gpurun_out/bench_c2_2gpu_i.err
echo "== bench C2 N=2 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_2gpu_i.json'));print(d['n_gpus'],round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'])")"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_2gpu_ref_i.json 2> gpurun_out/bench_c2_2gpu_ref_i.err; echo "== bench ref N=2 exit $? : $(head -c 120 gpurun_out/bench_c2_2gpu_ref_i.json)"
for n in 1 2; do
  if [ $n = 1 ]; then CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --config C3 --mode sweep --images 2048 > gpurun_out/sweep_c3_n1_i.json 2> gpurun_out/sweep_c3_n1_i.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 2 --config C3 --mode sweep --images 2048 > gpurun_out/sweep_c3_n2_i.json 2> gpurun_out/sweep_c3_n2_i.err; fi
  echo "== C3 sweep N=$n exit $? : $(python -c "import json;d=json.load(open('gpurun_out/sweep_c3_n${n}_i.json'));print(round(d['value']),d['sweep_wall_s'],d['arena_fnv'])")"
done
