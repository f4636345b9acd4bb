#!/bin/bash
# round 2, call S: JPEG encoder with the tile copy issued at the top of the fused kernel: byte-exactness, A/B timing (compare with profiles/r02_g_jpeg_two_kernels.md)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s.log
tail -5 gpurun_out/pytest_s.log | cut -c1-800
timeout 900 python tools/jpeg_ab.py quick > gpurun_out/jpeg_ab_s.txt 2>&1; cat gpurun_out/jpeg_ab_s.txt | tail -20
