#!/bin/bash
# round 2, call M: staged line kernel with separate input / output staging (modes 0 direct, 1 both, 2 output only, 3 input only): parity of all forms, sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixfmt_gpu.py tests/test_cuda_wrapper_kernels.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_m.log
tail -4 gpurun_out/pytest_m.log | cut -c1-600
timeout 900 python tools/pixfmt_sweep.py 0,1,2,3 > gpurun_out/pixfmt_sweep_m.txt 2>&1; tail -70 gpurun_out/pixfmt_sweep_m.txt | cut -c1-200
