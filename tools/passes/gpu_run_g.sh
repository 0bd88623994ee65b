#!/usr/bin/env bash
# Round-2 GPU pass G: the software-pipelined sweep (K1 beside the forward): parity subset, C2 / C3 / C4 bench lines, C2 block-size soak.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -k "sweep or cell_end or full_size or smoke_size or c_only or allreduce" > gpurun_out/gputest_g.log 2>&1
echo "== pytest(sweep) exit $? : $(tail -n 1 gpurun_out/gputest_g.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_g.log | head -20
timeout 900 python bench.py --steps 200 --warmup 5 --no-extras > gpurun_out/bench_c2_g.json 2> gpurun_out/bench_c2_g.err; echo "== bench C2 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_g.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'],d['clocks'])")"
timeout 600 python bench.py --config C3 --steps 75 --warmup 3 --soak 4 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_c3_g.json 2> gpurun_out/bench_c3_g.err; echo "== bench C3 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c3_g.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'])")"
timeout 600 python bench.py --config C4 --steps 20 --warmup 3 --soak 4 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_c4_g.json 2> gpurun_out/bench_c4_g.err; echo "== bench C4 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c4_g.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'])")"
for b in 1024 2048 8192; do
  timeout 600 python bench.py --block $b --steps 100 --warmup 5 --soak 5 --no-cpu-baseline --no-kernel-rooflines --no-extras > gpurun_out/bench_c2_block$b.json 2> gpurun_out/bench_c2_block$b.err
  echo "== C2 block $b exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_block$b.json'));print(round(d['value']),'sust',round(d['sustained']['value']),d['sustained']['clocks']['sm_mhz'],'e2e',round(d['e2e']['value']),d['ms_per_step'])")"
done
