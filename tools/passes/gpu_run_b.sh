#!/usr/bin/env bash
# Round-2 GPU pass B: full GPU suite with the new K1 / K34 kernels, bench lines (K1 / K34 rooflines), ncu launch list of the
# bench command, ncu --set full of the K1 / K34 kernels at the bench shapes.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 --durations=8 > gpurun_out/gputest_b.log 2>&1
echo "== pytest exit $? : $(tail -n 3 gpurun_out/gputest_b.log | tr '\n' ' ')"
grep -E "FAILED|ERROR" gpurun_out/gputest_b.log | head -40
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_b.json 2> gpurun_out/bench_c2_b.err; echo "== bench C2 exit $? : $(head -c 300 gpurun_out/bench_c2_b.json)"
timeout 600 python bench.py --config C3 --steps 75 --warmup 3 > gpurun_out/bench_c3_b.json 2> gpurun_out/bench_c3_b.err; echo "== bench C3 exit $? : $(head -c 300 gpurun_out/bench_c3_b.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2_b.csv python bench.py --steps 20 --warmup 3 --soak 0 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/ncu_c2_b.log 2>&1; echo "== ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_|k34' -o gpurun_out/k1_full -f python tools/k1_ncu.py > gpurun_out/ncu_k1.log 2>&1; echo "== ncu k1 full exit $? $(ls -la gpurun_out/k1_full.ncu-rep 2>/dev/null | awk '{print $5}')"
