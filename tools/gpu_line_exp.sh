#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "UGB200_LINE_L1=1" "UGB200_LINE_STAGE=1" "UGB200_LINE_L1=1 UGB200_LINE_STAGE=1"; do
  echo "== $cfg"
  env $cfg timeout 200 python tools/pixfmt_sweep.py 2>&1 | grep -E "^\| (UYVY \| RGB |RGB \| UYVY|v210 \| UYVY|UYVY \| v210|RGB \| RGBA|RGBA \| RGB |v210 \| RGB |UYVY \| RGBA|RG48 \| RGB |UYVY \| RG48|YUYV \| UYVY|BGR \| RGB |R10k \| RGBA)"
done
UGB200_LINE_L1=1 UGB200_LINE_STAGE=1 timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 -k "pixfmt or named" 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -q -x --timeout 150 -k "two_encoders or survives or pitch or named" 2>&1 | tail -2
UGB200_JPEG_TIMING=1 timeout 100 python tools/profile_target.py jpegdec 2>&1 | tail -6
