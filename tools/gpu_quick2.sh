#!/bin/bash
mkdir -p gpurun_out
UGB200_JPEG_SPLIT=1 timeout 40 python -m pytest tests/test_jpeg.py -m gpu -x -q -k "equals_oracle_bytes or largest_ac" > gpurun_out/pytest_compact_split.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_compact_split.log; tail -2 gpurun_out/pytest_compact_split.log
timeout 40 python -m pytest tests/test_vcompress.py tests/test_jpeg_wire.py -m gpu -x -q -k "gpujpeg or wire or jpeg" > gpurun_out/pytest_compact_mod.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_compact_mod.log; tail -2 gpurun_out/pytest_compact_mod.log
