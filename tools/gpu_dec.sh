#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 -k "jpeg_decode or vdecompress" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 200 python tools/profile_target.py jpegdec 2>&1 | tail -6
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_dec.csv python tools/profile_target.py jpegdec > gpurun_out/dec_prof.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/launches_dxtdec.csv python tools/profile_target.py dxtdec > gpurun_out/dxtdec_prof.log 2>&1
python tools/summarize_ncu.py - gpurun_out/launches_dec.csv gpurun_out/dec_sum.md "jpeg decode" > /dev/null 2>&1; cat gpurun_out/dec_sum.md | tail -12
python tools/summarize_ncu.py - gpurun_out/launches_dxtdec.csv gpurun_out/dxtdec_sum.md "dxt decode" > /dev/null 2>&1; cat gpurun_out/dxtdec_sum.md | tail -6
