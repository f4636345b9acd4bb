#!/bin/bash
# round 2, call N: per-converter staged defaults: converter parity, sweep of the defaults, ncu --set full of v210 -> RG48 and R12L -> RG48 in the direct and the shipped form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixfmt_gpu.py tests/test_cuda_wrapper_kernels.py tests/test_named_line_converters.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n.log
tail -4 gpurun_out/pytest_n.log | cut -c1-600
timeout 900 python tools/pixfmt_sweep.py > gpurun_out/pixfmt_sweep_n.txt 2>&1; tail -70 gpurun_out/pixfmt_sweep_n.txt | cut -c1-200
for pair in "7 27 v210_rg48" "6 27 r12l_rg48"; do
  set -- $pair
  UGB200_LINE_STAGED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_$3_direct -f python tools/profile_target.py conv $1 $2 > gpurun_out/ncu_$3_direct.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_$3_staged -f python tools/profile_target.py conv $1 $2 > gpurun_out/ncu_$3_staged.log 2>&1
done
ls -la gpurun_out/prof_*_staged.ncu-rep gpurun_out/prof_*_direct.ncu-rep
