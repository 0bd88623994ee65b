#!/usr/bin/env bash
# Round-2 GPU pass N (4-GPU box, final code): whole-sweep strong scaling of C2 / C3 at 1 / 2 / 4 GPUs with the pipelined sweep.
set -u
mkdir -p gpurun_out/scale4
run() { # gpus(csv) nproc port config images tag
  local devs=$1 n=$2 port=$3 cfg=$4 img=$5 tag=$6
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=$devs timeout 600 python bench.py --gpus 1 --config $cfg --mode sweep --images $img > gpurun_out/scale4/${tag}.json 2> gpurun_out/scale4/${tag}.err
  else
    CUDA_VISIBLE_DEVICES=$devs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --mode sweep --images $img > gpurun_out/scale4/${tag}.json 2> gpurun_out/scale4/${tag}.err
  fi
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/scale4/${tag}.json'));print(d['n_gpus'],round(d['value']),d['sweep_wall_s'],d['arena_fnv'])")"
}
run 0 1 29741 C2 10000 c2_n1 &
run 1 1 29742 C3 4096 c3_n1 &
run 2,3 2 29743 C2 10000 c2_n2 &
wait
run 0,1 2 29744 C3 4096 c3_n2 &
wait
run 0,1,2,3 4 29745 C2 10000 c2_n4
run 0,1,2,3 4 29746 C3 4096 c3_n4
