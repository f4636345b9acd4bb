#!/bin/bash
# JPEG experiment: event timing of the fused and the split path on three 8K contents + one full ncu capture of the fused kernel
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/jpeg_timing.log
import os, sys, subprocess
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import util
from ultragrid_b200 import api
W, H = 7680, 4320
orc = util.oracle()
yy, xx = np.mgrid[0:H, 0:W]
rgb = np.stack([xx * 255 // (W - 1), yy * 255 // (H - 1), (xx + yy) % 256], axis=2).astype(np.uint8)
nat = (rgb.astype(np.int16) + np.random.default_rng(1).integers(-6, 7, rgb.shape, dtype=np.int16)).clip(0, 255).astype(np.uint8)
srcs = {"natural": torch.from_numpy(util.convert_cpu(orc, "orc_convert", 12, 2, nat.reshape(-1), W, H)).cuda(),
        "noise": torch.randint(0, 256, (W * H * 2,), dtype=torch.uint8, device="cuda"),
        "flat": torch.full((W * H * 2,), 128, dtype=torch.uint8, device="cuda")}
enc = api.JpegEncoder()
for name, src in srcs.items():
    for q in (75, 90):
        for i in range(3): enc.encode_device(src, W, H, 2, quality=q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10): enc.encode_device(src, W, H, 2, quality=q)
        e1.record(); torch.cuda.synchronize()
        print(os.environ.get("UGB200_JPEG_SPLIT", "fused"), name, "q", q, "us/frame %.1f" % (e0.elapsed_time(e1) / 10 * 1e3), "bytes", len(enc.result()))
'''
for split in (None, "1"):
    env = dict(os.environ)
    if split: env["UGB200_JPEG_SPLIT"] = "split"
    subprocess.run([sys.executable, "-c", code], env=env)
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:jpeg_fused -s 1 -c 1 -f -o gpurun_out/jpeg_fused python tools/profile_target.py jpeg > gpurun_out/ncu_jpeg.log 2>&1
tail -3 gpurun_out/ncu_jpeg.log
