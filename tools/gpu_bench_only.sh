#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 300 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
timeout 40 tools/exp_e2e jpeg 2>&1 | tail -2
