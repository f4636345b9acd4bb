#!/bin/bash
# experiment call 4: JPEG after the compact-kernel rewrite and the clamp removal (timing + byte parity of every route)
mkdir -p gpurun_out
for cap in adaptive 16; do echo "cap $cap"; if [ $cap = adaptive ]; then timeout 60 tools/exp_e2e jpeg 2>&1 | tail -2; else UGB200_JPEG_CAP=$cap timeout 60 tools/exp_e2e jpeg 2>&1 | tail -2; fi; done | tee gpurun_out/exp_jpeg_compact.txt
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/jpeg_launches2.csv tools/exp_e2e jpeg > /dev/null 2>&1; grep -E "jpeg_" gpurun_out/jpeg_launches2.csv | awk -F'","' '{print $5, $NF}' | tail -8
timeout 400 python -m pytest tests/test_jpeg.py tests/test_jpeg_wire.py tests/test_vcompress.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_part4.log 2>&1; tail -3 gpurun_out/pytest_part4.log
