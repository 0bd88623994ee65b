#!/usr/bin/env bash
# Round-2 GPU pass K: chunk-per-thread K1 kernels: K1 parity subset + reset / CLI tests, C2 and C3 bench lines with the K1
# rooflines, ncu --set full summary of the K1 / K34 kernels.
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -k "k1_ or reset or cli or c_only or gate" > gpurun_out/gputest_k.log 2>&1
echo "== pytest(k1) exit $? : $(tail -n 1 gpurun_out/gputest_k.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_k.log | head -20
timeout 900 python bench.py --steps 100 --warmup 5 --soak 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c2_k.json 2> gpurun_out/bench_c2_k.err; echo "== bench C2 exit $? : $(python -c "
import json;d=json.load(open('gpurun_out/bench_c2_k.json'));print(round(d['value']),'e2e',round(d['e2e']['value']));k=d['roofline_k1'];print({c:round(v['frac_mean'],3) for c,v in k['frac_by_corruption'].items()});print({c:round(v,3) for c,v in k['steady_state']['frac'].items()})")"
timeout 600 python bench.py --config C3 --steps 75 --warmup 3 --soak 3 --no-cpu-baseline > gpurun_out/bench_c3_k.json 2> gpurun_out/bench_c3_k.err; echo "== bench C3 exit $? : $(python -c "
import json;d=json.load(open('gpurun_out/bench_c3_k.json'));print(round(d['value']),'e2e',round(d['e2e']['value']));k=d['roofline_k1'];print({c:round(v['frac_mean'],3) for c,v in k['frac_by_corruption'].items()});print({c:round(v,3) for c,v in k['steady_state']['frac'].items()})")"
timeout 900 ncu --set full --clock-control none -k regex:'k1_|k34' -o /tmp/ncu/k1_full -f python tools/k1_ncu.py > gpurun_out/ncu_k1_k.log 2>&1; echo "== ncu k1 full exit $?"
python tools/ncu_summary.py full /tmp/ncu/k1_full.ncu-rep > gpurun_out/k1_k34_full_k.txt 2>&1; echo "== k1 summary $(wc -l < gpurun_out/k1_k34_full_k.txt) lines"
