#!/bin/bash
# 2-GPU check: module/multi-device tests, then the N=2 bench line and the N=2 reference arm (as the driver launches them)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
timeout 400 python -m pytest tests -m gpu -q -x --timeout 150 -k "jpeg_decode or vdecompress or vcompress" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cat gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
