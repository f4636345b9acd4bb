#!/bin/bash
# quick GPU check: bash tools/gpu_quick.sh "<pytest -k expr>" [extra command]
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 150 -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -40 gpurun_out/pytest_quick.log | cut -c1-2000
if [ -n "$2" ]; then bash -c "$2"; fi
