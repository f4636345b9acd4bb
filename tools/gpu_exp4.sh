#!/bin/bash
# experiment call 3: DXT5-YCoCg after the instruction-count work, JPEG with 7 CTAs per SM (cap 12 / 8) against 6 (cap 16), parity of both
mkdir -p gpurun_out
timeout 100 tools/exp_dxt > gpurun_out/exp_dxt3.txt 2>&1; echo "exp_dxt rc=$?"; cat gpurun_out/exp_dxt3.txt
for cap in 16 12 8 adaptive; do echo "cap $cap"; if [ $cap = adaptive ]; then timeout 60 tools/exp_e2e jpeg 2>&1 | tail -2; else UGB200_JPEG_CAP=$cap timeout 60 tools/exp_e2e jpeg 2>&1 | tail -2; fi; done | tee gpurun_out/exp_jpeg_cap.txt
timeout 420 python -m pytest tests/test_dxt_gpu.py tests/test_jpeg.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_part3.log 2>&1; tail -3 gpurun_out/pytest_part3.log
