#!/bin/bash
# experiment call: DXT launch-shape variants (timing + byte parity), module e2e with host frames, full ncu captures of the shipped kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi2.txt 2>&1
timeout 120 tools/exp_dxt > gpurun_out/exp_dxt.txt 2>&1; echo "exp_dxt rc=$?"; cat gpurun_out/exp_dxt.txt
timeout 180 tools/exp_e2e 48 > gpurun_out/exp_e2e.txt 2>&1; echo "exp_e2e rc=$?"; cat gpurun_out/exp_e2e.txt
timeout 60 tools/exp_e2e jpeg > gpurun_out/exp_jpeg.txt 2>&1; cat gpurun_out/exp_jpeg.txt
for v in d1_b2_t128_m6_true d6_b1_t128_m6_true d1_b1_t128_m7_true; do
  timeout 120 ncu --set full --import-source on --clock-control none -k regex:exp_kernel -s 1 -c 1 -f -o gpurun_out/exp_$v tools/exp_dxt one $v > gpurun_out/ncu_$v.log 2>&1; tail -1 gpurun_out/ncu_$v.log
done
timeout 120 ncu --set full --import-source on --clock-control none -k regex:jpeg_fused -s 2 -c 1 -f -o gpurun_out/exp_jpeg_fused tools/exp_e2e jpeg > gpurun_out/ncu_jpeg2.log 2>&1; tail -1 gpurun_out/ncu_jpeg2.log
timeout 400 python -m pytest tests/test_vcompress.py tests/test_jpeg.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
ls -la gpurun_out | head -30
