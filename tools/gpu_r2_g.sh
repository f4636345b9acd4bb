#!/bin/bash
# round 2, call G: two-kernel JPEG form: byte-exactness (JPEG, module, wire tests) in both forms, then the A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py tests/test_real_module.py tests/test_jpeg_wire.py tests/test_jpeg_decode.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -8 gpurun_out/pytest_g.log | cut -c1-800
UGB200_JPEG_ONE_KERNEL=1 timeout 600 python -m pytest tests/test_jpeg.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_g1.log 2>&1; echo "pytest one-kernel rc=$?"; tail -2 gpurun_out/pytest_g1.log
UGB200_JPEG_A8=1 timeout 600 python -m pytest tests/test_jpeg.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_g8.log 2>&1; echo "pytest a8 rc=$?"; tail -2 gpurun_out/pytest_g8.log
timeout 900 python tools/jpeg_ab.py > gpurun_out/jpeg_ab.txt 2>&1; cat gpurun_out/jpeg_ab.txt | tail -20
