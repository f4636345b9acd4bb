#!/bin/bash
# experiment call 2: persistent / skewed DXT variants, JPEG start skew, JPEG parity after the shared-memory column swizzle
mkdir -p gpurun_out
timeout 150 tools/exp_dxt > gpurun_out/exp_dxt2.txt 2>&1; echo "exp_dxt rc=$?"; cat gpurun_out/exp_dxt2.txt
for sk in 0 12000 40000; do echo "skew $sk"; UGB200_JPEG_SKEW=$sk timeout 60 tools/exp_e2e jpeg 2>&1 | tail -2; done | tee gpurun_out/exp_jpeg_skew.txt
timeout 300 python -m pytest tests/test_jpeg.py tests/test_vcompress.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_part2.log 2>&1; tail -3 gpurun_out/pytest_part2.log
