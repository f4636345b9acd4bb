#!/bin/bash
# round 2, last call: the library as finally built - all GPU parity tests + smoke, then the JPEG encoder's lean instantiation against the general kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 300 python tools/jpeg_ab.py lean > gpurun_out/jpeg_ab_lean.txt 2>&1; tail -9 gpurun_out/jpeg_ab_lean.txt
UGB200_JPEG_LEAN=0 timeout 300 python -m pytest tests/test_jpeg.py -m gpu -q -x --timeout 280 -k "equals_oracle_bytes or serial_route or 8k or 7680" > gpurun_out/pytest_general.log 2>&1; tail -2 gpurun_out/pytest_general.log | cut -c1-200
