#!/bin/bash
# end-of-round evidence, most important first: all GPU parity tests, bench line (with extras), full ncu captures of the shipped DXT1 /
# DXT5-YCoCg kernels (first two launches of tools/exp_dxt = the shipped entry points) and of the fused JPEG kernel, smoke, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q -x -n 3 --timeout 280 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 60 ncu --set full --clock-control none --import-source on -k regex:dxt_uyvy_kernel -c 2 -o gpurun_out/prof_dxt_shipped -f tools/exp_dxt one none > gpurun_out/ncu_dxt.log 2>&1; tail -1 gpurun_out/ncu_dxt.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 3 -c 1 -o gpurun_out/prof_jpeg_fused7 -f tools/exp_e2e jpeg > gpurun_out/ncu_jpeg.log 2>&1; tail -1 gpurun_out/ncu_jpeg.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-extra > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
