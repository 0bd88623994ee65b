#!/usr/bin/env bash
# Round-2 GPU pass D: full GPU suite on the final K1 kernels, C3 block-size experiment, the final C2 bench line (with the C3 / C4 /
# C5 sub-records), ncu launch list of the bench command, ncu --set full of the K1 / K34 kernels and of the conv launches of a step.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 --durations=5 > gpurun_out/gputest_d.log 2>&1
echo "== pytest exit $? : $(tail -n 2 gpurun_out/gputest_d.log | tr '\n' ' ')"
grep -E "FAILED|ERROR" gpurun_out/gputest_d.log | head -40
for b in 256 512 1024; do
  timeout 600 python bench.py --config C3 --block $b --steps 75 --warmup 3 --soak 4 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_c3_block$b.json 2> gpurun_out/bench_c3_block$b.err
  echo "== C3 block $b exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c3_block$b.json'));print(round(d['value']),'sust',round(d['sustained']['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'])")"
done
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_d.json 2> gpurun_out/bench_c2_d.err; echo "== bench C2 exit $? : $(head -c 200 gpurun_out/bench_c2_d.json)"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref_d.json 2> gpurun_out/bench_c2_ref_d.err; echo "== bench ref exit $? : $(head -c 150 gpurun_out/bench_c2_ref_d.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2_d.csv python bench.py --steps 20 --warmup 3 --soak 0 --no-cpu-baseline --no-kernel-rooflines --no-extras > gpurun_out/ncu_c2_d.log 2>&1; echo "== ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_|k34' -o gpurun_out/k1_full_d -f python tools/k1_ncu.py > gpurun_out/ncu_k1_d.log 2>&1; echo "== ncu k1 full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_' -c 18 -o gpurun_out/conv_full_d -f python tools/conv_ncu.py > gpurun_out/ncu_conv_d.log 2>&1; echo "== ncu conv full exit $?"
