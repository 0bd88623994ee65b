#!/usr/bin/env bash
# Round-2 GPU pass C (8-GPU box): whole-sweep STRONG scaling of C2 / C3 at 1/2/4/8 GPUs, C4 at 1 and 8 GPUs, the real 2-GPU
# bit-identity test.  Independent runs share the box side by side (disjoint GPUs) to keep the box time short.
set -u
mkdir -p gpurun_out/scale
run() { # gpus(csv) nproc port config images tag
  local devs=$1 n=$2 port=$3 cfg=$4 img=$5 tag=$6
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=$devs timeout 900 python bench.py --gpus 1 --config $cfg --mode sweep --images $img > gpurun_out/scale/${tag}.json 2> gpurun_out/scale/${tag}.err
  else
    CUDA_VISIBLE_DEVICES=$devs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --mode sweep --images $img > gpurun_out/scale/${tag}.json 2> gpurun_out/scale/${tag}.err
  fi
  echo "== $tag exit $? : $(head -c 200 gpurun_out/scale/${tag}.json)"
}
nvidia-smi -L | head -8
# phase 1: five independent runs on 7 GPUs
run 0 1 29701 C2 10000 c2_n1 &
run 1 1 29702 C3 2048 c3_n1 &
run 2 1 29703 C4 512 c4_n1 &
run 3,4 2 29704 C2 10000 c2_n2 &
run 5,6 2 29705 C3 2048 c3_n2 &
wait
# phase 2
run 0,1,2,3 4 29706 C2 10000 c2_n4 &
run 4,5,6,7 4 29707 C3 2048 c3_n4 &
wait
# phase 3: the whole box
run 0,1,2,3,4,5,6,7 8 29708 C2 10000 c2_n8
run 0,1,2,3,4,5,6,7 8 29709 C3 2048 c3_n8
run 0,1,2,3,4,5,6,7 8 29710 C4 512 c4_n8
# weak-scaling step lines of C3 / C4 at 8 GPUs (the driver runs C2 itself)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --config C3 --steps 40 --warmup 3 --soak 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/scale/c3_steps_n8.json 2> gpurun_out/scale/c3_steps_n8.err; echo "== c3 steps n8 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --config C4 --steps 10 --warmup 3 --soak 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/scale/c4_steps_n8.json 2> gpurun_out/scale/c4_steps_n8.err; echo "== c4 steps n8 exit $?"
# the real multi-GPU bit-identity test
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "two_gpu" > gpurun_out/scale/two_gpu_test.log 2>&1; echo "== 2-GPU test exit $? : $(tail -n 1 gpurun_out/scale/two_gpu_test.log)"
python - <<'PY'
import json, glob, os
rows = {}
for f in sorted(glob.glob("gpurun_out/scale/c*_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    tag = os.path.basename(f)[:-5]
    rows[tag] = d
    print(f"{tag:14s} n_gpus={d['n_gpus']} value={d['value']:.0f} {d['unit']} wall={d.get('sweep_wall_s')} fnv={d.get('arena_fnv')} clocks={d.get('clocks', {}).get('sm_mhz') if d.get('clocks') else None} {d.get('clocks', {}).get('reasons') if d.get('clocks') else None}")
for cfg in ("c2", "c3", "c4"):
    f = {rows[k].get("arena_fnv") for k in rows if k.startswith(cfg + "_n")}
    print(cfg, "arena checksums across N:", f, "BIT-IDENTICAL" if len(f) == 1 else "DIFFER")
PY
