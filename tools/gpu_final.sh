#!/bin/bash
# end-of-round evidence: all GPU parity tests, smoke, bench line (with extras), ncu launch list of the bench command, full capture of the headline kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxt_uyvy -s 2 -c 1 -o gpurun_out/prof_dxt1_final -f python tools/profile_target.py dxt1 > gpurun_out/ncu_dxt1.log 2>&1; tail -2 gpurun_out/ncu_dxt1.log
