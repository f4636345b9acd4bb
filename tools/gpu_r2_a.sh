#!/bin/bash
# round 2, call A: parity tests, smoke, the new bench line (all workloads), the CPU arm, launch list, in-stream DRAM traffic per kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -25 >> gpurun_out/nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/nproc.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-800
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 8 --passes 1 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
for k in dxt1 dxt6 p010 jpeg; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/traffic_$k.csv python tools/profile_target.py $k > gpurun_out/traffic_$k.log 2>&1
done
tail -4 gpurun_out/traffic_dxt1.csv | cut -c1-400
