#!/bin/bash
# round 2, call J: entropy threads take the CTA's blocks sorted by non-zero count: byte-exactness of every JPEG route, then A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py tests/test_real_module.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_j.log
tail -12 gpurun_out/pytest_j.log | cut -c1-1500
timeout 900 python tools/jpeg_ab.py quick > gpurun_out/jpeg_ab_j.txt 2>&1; cat gpurun_out/jpeg_ab_j.txt | tail -20
