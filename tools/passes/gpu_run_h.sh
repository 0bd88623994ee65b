#!/usr/bin/env bash
# Round-2 GPU pass H: the final default bench line (C2 at 8192 images / step, pipelined, with the C3 / C4 / C5 sub-records), the
# reference arm, and the ncu launch list of the same bench command.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_h.json 2> gpurun_out/bench_c2_h.err; echo "== bench C2 exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_h.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),'sust',round(d['sustained']['value']),d['ms_per_step'],d['clocks'],d['roofline']['frac'],d['roofline']['achieved'])")"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref_h.json 2> gpurun_out/bench_c2_ref_h.err; echo "== bench ref exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_h20.json 2> gpurun_out/bench_c2_h20.err; echo "== bench C2 (driver-like 20 steps) exit $? : $(python -c "import json;d=json.load(open('gpurun_out/bench_c2_h20.json'));print(round(d['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'],d['clocks'])")"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2_h.csv python bench.py --steps 20 --warmup 3 --soak 0 --no-cpu-baseline --no-kernel-rooflines --no-extras > gpurun_out/ncu_c2_h.log 2>&1; echo "== ncu launches exit $?"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -k "k34 or k3_ or k4_ or end_to_end" > gpurun_out/gputest_h.log 2>&1; echo "== pytest(k34) exit $? : $(tail -n 1 gpurun_out/gputest_h.log)"
