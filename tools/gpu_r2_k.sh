#!/bin/bash
# round 2, call K: ncu --set full of the fused JPEG kernel with the sorted hand-out (compare with prof_jpeg_uyvy of state d)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 4 -c 1 -o gpurun_out/prof_jpeg_uyvy_sorted -f python tools/profile_target.py jpeg > gpurun_out/ncu_jpeg_uyvy_sorted.log 2>&1; tail -2 gpurun_out/ncu_jpeg_uyvy_sorted.log
UGB200_JPEG_SORT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 4 -c 1 -o gpurun_out/prof_jpeg_uyvy_unsorted -f python tools/profile_target.py jpeg > gpurun_out/ncu_jpeg_uyvy_unsorted.log 2>&1; tail -2 gpurun_out/ncu_jpeg_uyvy_unsorted.log
ls -la gpurun_out/prof_jpeg_uyvy_*sorted.ncu-rep
