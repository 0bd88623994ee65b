#!/usr/bin/env bash
# Round-2 GPU pass L (final code): full GPU suite + smoke, the final default bench line + reference arm, ncu launch list of the
# bench command, ncu --set full summaries of the K1 / K34 kernels and of the conv launches of a C3 step.
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/gputest_l.log 2>&1
echo "== pytest exit $? : $(tail -n 1 gpurun_out/gputest_l.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_l.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_l.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke_l.log | cut -c1-120)"
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_l.json 2> gpurun_out/bench_c2_l.err; echo "== bench C2 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_l.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'],d['clocks'],d['roofline']['frac'])")"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref_l.json 2> gpurun_out/bench_c2_ref_l.err; echo "== bench ref exit $?"
timeout 600 python bench.py --config C3 --steps 75 --warmup 3 --soak 4 > gpurun_out/bench_c3_l.json 2> gpurun_out/bench_c3_l.err; echo "== bench C3 exit $? : $(head -c 120 gpurun_out/bench_c3_l.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2_l.csv python bench.py --steps 20 --warmup 3 --soak 0 --no-cpu-baseline --no-kernel-rooflines --no-extras > gpurun_out/ncu_c2_l.log 2>&1; echo "== ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:'k1_|k34' -o /tmp/ncu/k1_full -f python tools/k1_ncu.py > gpurun_out/ncu_k1_l.log 2>&1; echo "== ncu k1 full exit $?"
python tools/ncu_summary.py full /tmp/ncu/k1_full.ncu-rep > gpurun_out/k1_k34_full_l.txt 2>&1; echo "== k1 summary $(wc -l < gpurun_out/k1_k34_full_l.txt) lines"
timeout 900 ncu --set full --clock-control none -k regex:'conv' -c 50 -o /tmp/ncu/conv_c3 -f python tools/conv_ncu.py 256 1 resnet50 224 > gpurun_out/ncu_conv_c3_l.log 2>&1; echo "== ncu conv C3 exit $?"
python tools/ncu_summary.py full /tmp/ncu/conv_c3.ncu-rep > gpurun_out/conv_full_c3_l.txt 2>&1; echo "== conv C3 summary: $(tail -n 1 gpurun_out/conv_full_c3_l.txt)"
du -sh gpurun_out
