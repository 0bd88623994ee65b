#!/usr/bin/env bash
# Round-2 GPU pass E: K1 tests on the final kernels, the final C2 bench line (with C3 / C4 / C5 sub-records) + reference arm, ncu launch
# list of the bench command, ncu --set full of the K1 / K34 kernels and of the conv launches of a step (summarised here: the
# .ncu-rep files are too large to bring back).
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -k "k1_ or c_only or smoke_size" > gpurun_out/gputest_e.log 2>&1
echo "== pytest(k1) exit $? : $(tail -n 1 gpurun_out/gputest_e.log)"
grep -E "FAILED|ERROR" gpurun_out/gputest_e.log | head -20
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_e.json 2> gpurun_out/bench_c2_e.err; echo "== bench C2 exit $? : $(head -c 200 gpurun_out/bench_c2_e.json)"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref_e.json 2> gpurun_out/bench_c2_ref_e.err; echo "== bench ref exit $?"
timeout 600 python bench.py --config C3 --steps 75 --warmup 3 --soak 4 > gpurun_out/bench_c3_e.json 2> gpurun_out/bench_c3_e.err; echo "== bench C3 exit $? : $(head -c 120 gpurun_out/bench_c3_e.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c2_e.csv python bench.py --steps 20 --warmup 3 --soak 0 --no-cpu-baseline --no-kernel-rooflines --no-extras > gpurun_out/ncu_c2_e.log 2>&1; echo "== ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:'k1_|k34' -o /tmp/ncu/k1_full -f python tools/k1_ncu.py > gpurun_out/ncu_k1_e.log 2>&1; echo "== ncu k1 full exit $?"
python tools/ncu_summary.py full /tmp/ncu/k1_full.ncu-rep > gpurun_out/k1_k34_full_e.txt 2>&1; echo "== k1 summary $(wc -l < gpurun_out/k1_k34_full_e.txt) lines"
timeout 900 ncu --set full --clock-control none -k regex:'conv_' -c 18 -o /tmp/ncu/conv_full -f python tools/conv_ncu.py > gpurun_out/ncu_conv_e.log 2>&1; echo "== ncu conv full exit $?"
python tools/ncu_summary.py full /tmp/ncu/conv_full.ncu-rep > gpurun_out/conv_full_e.txt 2>&1; echo "== conv summary: $(tail -n 1 gpurun_out/conv_full_e.txt)"
du -sh gpurun_out
