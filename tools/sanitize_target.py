"""exercises every C-ABI entry point once or twice at small, odd sizes - meant to run under `compute-sanitizer --tool memcheck`"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import util
import planar_cases as pc
from test_oracle_pinning import PAIRS
from ultragrid_b200 import api, _lib, vc_get_linesize
lib = _lib.load()
n = 0
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
for inc, outc in ([] if ONLY == "jpeg" else PAIRS):  # line converters: tight buffers, ragged widths
    for w, h in ((50, 3), (17, 2), (256, 2)):
        ls_i, ls_o = vc_get_linesize(w, inc), vc_get_linesize(w, outc)
        src = torch.randint(0, 256, (ls_i * h + 64,), dtype=torch.uint8, device="cuda")  # MAX_PADDING of over-read slack, video_codec.h:61
        dst = torch.zeros(ls_o * h + 64, dtype=torch.uint8, device="cuda")
        api.pixfmt_convert(inc, outc, src, w, h, dst=dst)
        n += 1
for name, depth in ([] if ONLY == "jpeg" else pc.all_cases()):  # planar converters
    for w, h in ((50, 5), (17, 3), (64, 4)):
        if name == "yuv420_to_i420" and (w % 2 or h % 2):
            continue
        for mode in (0, 2):
            c = pc.Case(name, w, h, seed=3, mode=mode, depth=depth)
            c.run_gpu(lib, torch, 0)
            n += 1
for w, h in ((8, 4), (260, 36)):  # DXT encode / decode
    uy = torch.randint(0, 256, (w * h * 2,), dtype=torch.uint8, device="cuda")
    rgb = torch.randint(0, 256, (w * h * 3,), dtype=torch.uint8, device="cuda")
    for t in (1, 6):
        blocks = api.uyvy_to_dxt(uy, w, h, dxt_type=t)
        api.dxt_to_rgb(blocks, w, h, t)
        api.compat_to_dxt("cuda_rgb_to_dxt1" if t == 1 else "cuda_rgb_to_dxt6", rgb, w, -h)
        n += 3
enc, dec = api.JpegEncoder(), api.JpegDecoder()
for codec, w, h, q, ri in ((2, 100, 52, 90, 0), (2, 98, 50, 100, 1), (12, 77, 33, 85, 8), (12, 64, 64, 100, 5)):  # JPEG encode (fused, serial route, split) / decode
    bpp = 2 if codec == 2 else 3
    src = torch.randint(0, 256, (((w + 1) // 2 * 2) * bpp * h,), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        enc.encode_device(src, w, h, codec, quality=q, restart_interval=ri)
        try:
            s = enc.result()
        except RuntimeError:
            s = None
    if s:
        for out_c in (2, 12, 1, 29):
            dec.decode(s, out_c)
        n += 6
torch.cuda.synchronize()
print("exercised", n, "calls")
