#!/bin/bash
# round 2, call AA: ncu --set full of the two line converters that stay below 0.70 of the copy bandwidth (v210 -> RGB, UYVY -> RGBA)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_v210_rgb -f python tools/profile_target.py conv 7 12 > gpurun_out/ncu_v210_rgb.log 2>&1; tail -1 gpurun_out/ncu_v210_rgb.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_uyvy_rgba -f python tools/profile_target.py conv 2 1 > gpurun_out/ncu_uyvy_rgba.log 2>&1; tail -1 gpurun_out/ncu_uyvy_rgba.log
ls -la gpurun_out/prof_v210_rgb.ncu-rep gpurun_out/prof_uyvy_rgba.ncu-rep
