#!/bin/bash
# round 2, call R: ncu --set full of the JPEG decoder's Huffman and IDCT kernels (8K UYVY stream)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_decode_huffman -s 1 -c 1 -o gpurun_out/prof_jpegdec_huffman -f python tools/profile_target.py jpegdec > gpurun_out/ncu_jpegdec_huffman.log 2>&1; tail -3 gpurun_out/ncu_jpegdec_huffman.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_idct_uyvy -s 1 -c 1 -o gpurun_out/prof_jpegdec_idct -f python tools/profile_target.py jpegdec > gpurun_out/ncu_jpegdec_idct.log 2>&1; tail -3 gpurun_out/ncu_jpegdec_idct.log
ls -la gpurun_out/*.ncu-rep
