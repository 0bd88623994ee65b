#!/usr/bin/env bash
# Round-2 GPU pass J: compute-sanitizer memcheck and racecheck over every K1 kernel, K3+K4 and a small forward.
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_k1.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "== memcheck exit $? : $(grep -E 'ERROR SUMMARY|done' gpurun_out/sanitize_memcheck.log | tail -n 2 | tr '\n' ' ')"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_k1.py k1only > gpurun_out/sanitize_racecheck.log 2>&1; echo "== racecheck exit $? : $(grep -E 'RACECHECK SUMMARY|done' gpurun_out/sanitize_racecheck.log | tail -n 2 | tr '\n' ' ')"
grep -E "Invalid|hazard|Error|error" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log | head -30
