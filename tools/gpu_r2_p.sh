#!/bin/bash
# round 2, call P: JPEG decoder - word-wise bit reader refill, marker scan of multi-scan (RGB) streams on the device: parity + A/B timing (UYVY and RGB 8K)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_jpeg_decode.py tests/test_vdecompress.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p.log
tail -12 gpurun_out/pytest_p.log | cut -c1-1500
timeout 900 python tools/jpegdec_ab.py short > gpurun_out/jpegdec_ab_p.txt 2>&1; cat gpurun_out/jpegdec_ab_p.txt | tail -12
for f in gpurun_out/jpegdec_RGB_device*_t2.txt gpurun_out/jpegdec_device*_t2.txt; do echo "== $f"; tail -14 "$f"; done
# DXT5-YCoCg decode variants (UGB200_DXT5DEC: bit 0 = no F2I, bit 1 = reciprocal + FMA division), each in its own process
for v in 0 1 2 3; do
UGB200_DXT5DEC=$v python - <<'PY'
import os, torch
from ultragrid_b200 import api
W, H = 7680, 4320
blocks = [torch.randint(0, 256, (W * H,), dtype=torch.uint8, device="cuda") for _ in range(4)]
out = torch.empty(W * H * 3, dtype=torch.uint8, device="cuda")
for i in range(6):
    api.dxt_to_rgb(blocks[i % 4], W, H, 6, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(24):
    api.dxt_to_rgb(blocks[i % 4], W, H, 6, out=out)
e1.record(); torch.cuda.synchronize()
print("dxt5ycocg decode variant", os.environ["UGB200_DXT5DEC"], "%.1f us" % (e0.elapsed_time(e1) / 24 * 1e3), "checksum", int(out.to(torch.int64).sum()))
PY
done
UGB200_DXT5DEC=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dxt5ycocg_decode -s 1 -c 1 -o gpurun_out/prof_dxt5dec_v0 -f python tools/profile_target.py dxtdec > gpurun_out/ncu_dxt5dec_v0.log 2>&1
ls -la gpurun_out/*.ncu-rep
