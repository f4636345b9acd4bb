#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg_decode.py tests/test_vdecompress.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -3 gpurun_out/pytest_h.log | cut -c1-1200
timeout 900 python tools/jpegdec_ab.py > gpurun_out/jpegdec_ab.txt 2>&1; cat gpurun_out/jpegdec_ab.txt | tail
for f in gpurun_out/jpegdec_device_scan_default_for_this_stream_t1.txt gpurun_out/jpegdec_device_scan_default_for_this_stream_t2.txt gpurun_out/jpegdec_host_scan_t2.txt; do echo "--- $f"; sed -n 300,330p $f; done
