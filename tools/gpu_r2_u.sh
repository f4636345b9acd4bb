#!/bin/bash
# round 2, call U: compute-sanitizer memcheck over every entry point incl. the staged line kernels, the kernels.cu entry points and the device marker scans
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py 2>&1 | tail -2
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer_i.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_i.log
grep -E "ERROR SUMMARY|Invalid|sanitizer rc|exercised|at 0x|by thread" gpurun_out/sanitizer_i.log | head -40
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 10 python tools/sanitize_target.py jpeg > gpurun_out/racecheck_i.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_i.log
grep -E "RACECHECK SUMMARY|hazard|racecheck rc|exercised" gpurun_out/racecheck_i.log | head -20
