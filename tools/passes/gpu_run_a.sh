#!/usr/bin/env bash
# Round-2 GPU pass A: full GPU parity suite, smoke, bench lines of every config, ncu launch list of the bench command.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 --durations=20 > gpurun_out/gputest.log 2>&1
echo "== pytest exit $? : $(tail -n 3 gpurun_out/gputest.log | tr '\n' ' ')"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log | cut -c1-300)"
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "== bench C2 exit $? : $(head -c 600 gpurun_out/bench_c2.json)"
timeout 600 python bench.py --config C3 --steps 40 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "== bench C3 exit $? : $(head -c 400 gpurun_out/bench_c3.json)"
timeout 600 python bench.py --config C4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "== bench C4 exit $? : $(head -c 400 gpurun_out/bench_c4.json)"
timeout 600 python bench.py --config C5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "== bench C5 exit $? : $(head -c 400 gpurun_out/bench_c5.json)"
timeout 600 python bench.py --config C2 --mode sweep --images 8192 > gpurun_out/sweep_c2_1gpu.json 2> gpurun_out/sweep_c2.err; echo "== sweep C2 exit $? : $(head -c 300 gpurun_out/sweep_c2_1gpu.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --soak 0 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1; echo "== ncu launches exit $?"
tail -n 40 gpurun_out/gputest.log
