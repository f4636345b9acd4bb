#!/bin/bash
# round 2, very last call (3 GPU-minutes left): lean line kernel - converter parity, then the 8K sweep with the defaults, with every converter in the direct lean form, and without lean
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_pixfmt_gpu.py tests/test_named_line_converters.py tests/test_cuda_wrapper_kernels.py -m gpu -q -x --timeout 90 > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lean.log
tail -3 gpurun_out/pytest_lean.log | cut -c1-300
timeout 60 python tools/pixfmt_sweep.py > gpurun_out/pixfmt_sweep_lean.txt 2>&1; echo "sweep default done"
UGB200_LINE_STAGED=0 timeout 60 python tools/pixfmt_sweep.py > gpurun_out/pixfmt_sweep_lean_direct.txt 2>&1; echo "sweep direct done"
UGB200_LINE_LEAN=0 timeout 60 python tools/pixfmt_sweep.py > gpurun_out/pixfmt_sweep_nolean.txt 2>&1; echo "sweep nolean done"
