#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 -k "pixfmt" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 200 python tools/pixfmt_sweep.py 2>&1 | grep -E "UYVY . RGBA|v210 . RGB "
timeout 120 python tools/sanitize_target.py 2>&1 | tail -2
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
grep -E "ERROR SUMMARY|Invalid|sanitizer rc|exercised|at 0x|by thread" gpurun_out/sanitizer.log | head -40
