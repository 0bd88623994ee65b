#!/usr/bin/env bash
# Round-2 GPU pass F (final): full GPU suite + smoke on the committed code, the conv ncu capture with the flat kernel included,
# stand-alone C4 / C5 lines.
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 --durations=5 > gpurun_out/gputest_f.log 2>&1
echo "== pytest exit $? : $(tail -n 1 gpurun_out/gputest_f.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_f.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke_f.log | cut -c1-200)"
timeout 900 ncu --set full --clock-control none -k regex:'conv' -c 18 -o /tmp/ncu/conv_full -f python tools/conv_ncu.py > gpurun_out/ncu_conv_f.log 2>&1; echo "== ncu conv full exit $?"
python tools/ncu_summary.py full /tmp/ncu/conv_full.ncu-rep > gpurun_out/conv_full_f.txt 2>&1; echo "== conv summary: $(tail -n 1 gpurun_out/conv_full_f.txt)"
timeout 600 python bench.py --config C5 > gpurun_out/bench_c5_f.json 2> gpurun_out/bench_c5_f.err; echo "== bench C5 exit $? : $(head -c 160 gpurun_out/bench_c5_f.json)"
timeout 600 python bench.py --config C4 --steps 20 --warmup 3 --soak 4 > gpurun_out/bench_c4_f.json 2> gpurun_out/bench_c4_f.err; echo "== bench C4 exit $? : $(head -c 160 gpurun_out/bench_c4_f.json)"
timeout 300 python -m fav.sweep --images 2048 --passes 20 --format json --out gpurun_out/sweep_cli_f.json > gpurun_out/sweep_cli_f.log 2>&1; echo "== sweep CLI exit $? : $(head -c 300 gpurun_out/sweep_cli_f.json)"
du -sh gpurun_out
