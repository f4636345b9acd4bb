#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg_decode.py tests/test_vdecompress.py tests/test_jpeg_wire.py "tests/test_jpeg.py::test_gpu_two_kernel_form_is_byte_identical" -m gpu -q -x --timeout 600 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -4 gpurun_out/pytest_h.log | cut -c1-1200
timeout 600 python tools/jpegdec_ab.py > gpurun_out/jpegdec_ab.txt 2>&1; cat gpurun_out/jpegdec_ab.txt | tail; tail -24 gpurun_out/jpegdec_device.txt
