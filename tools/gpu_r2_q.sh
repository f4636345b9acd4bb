#!/bin/bash
# round 2, call Q (P + the stream upload on a copy stream, double-buffered device copy): JPEG decoder - word-wise bit reader refill, marker scan of multi-scan (RGB) streams on the device: parity + A/B timing (UYVY and RGB 8K)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_jpeg_decode.py tests/test_vdecompress.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -12 gpurun_out/pytest_q.log | cut -c1-1500
timeout 900 python tools/jpegdec_ab.py short > gpurun_out/jpegdec_ab_q.txt 2>&1; cat gpurun_out/jpegdec_ab_q.txt | tail -12
for f in gpurun_out/jpegdec_RGB_device*_t2.txt gpurun_out/jpegdec_device*_t2.txt; do echo "== $f"; tail -14 "$f"; done
