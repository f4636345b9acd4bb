#!/bin/bash
# round 2, call H: device marker scan of the JPEG decoder (parity with the host scan, A/B timing), two-kernel encoder test, launch list of the decode path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg_decode.py tests/test_vdecompress.py tests/test_jpeg_wire.py "tests/test_jpeg.py::test_gpu_two_kernel_form_is_byte_identical" -m gpu -q -x --timeout 600 > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -8 gpurun_out/pytest_h.log | cut -c1-1200
timeout 600 python tools/jpegdec_ab.py > gpurun_out/jpegdec_ab.txt 2>&1; cat gpurun_out/jpegdec_ab.txt | tail; tail -12 gpurun_out/jpegdec_device.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/jpegdec_launches.csv python tools/profile_target.py jpegdec > gpurun_out/jpegdec_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/jpegdec_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]: print(r[4][:70], r[-1])
PY
