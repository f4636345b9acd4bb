#!/bin/bash
# round 2, call I: balanced entropy coder: byte-exactness of every JPEG route, then A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py tests/test_real_module.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_i.log
tail -12 gpurun_out/pytest_i.log | cut -c1-1500
UGB200_JPEG_BALANCED=0 timeout 600 python -m pytest tests/test_jpeg.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_i0.log 2>&1; echo "pytest block-per-thread rc=$?"; tail -2 gpurun_out/pytest_i0.log
timeout 900 python tools/jpeg_ab.py quick > gpurun_out/jpeg_ab_i.txt 2>&1; cat gpurun_out/jpeg_ab_i.txt | tail -20
