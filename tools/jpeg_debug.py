"""debug: fused vs oracle stream on the failing test image; which blocks differ after decoding both"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, util
import test_jpeg as tj
import test_jpeg_decode as tjd
from ultragrid_b200 import api
orc = util.oracle()
for (codec, w, h, q, ri) in ((tj.UYVY, 64, 32, 75, 2), (tj.UYVY, 100, 52, 90, 0), (tj.RGB, 64, 64, 100, 4)):
    if codec == tj.UYVY:
        src = util.convert_cpu(orc, "orc_convert", 12, 2, tj.natural_rgb(w, h, 5).reshape(-1), w, h)
        src[: w * 2 * min(h, 8)] = util.rng_bytes(w * 2 * min(h, 8), 1)
    else:
        src = tj.natural_rgb(w, h, 7).reshape(-1).copy()
        src[: w * 3 * min(h, 8)] = util.rng_bytes(w * 3 * min(h, 8), 2)
    want = tj.orc_encode(orc, src, w, h, codec, q, ri)
    enc = api.JpegEncoder()
    got = enc.encode(src, w, h, codec, quality=q, restart_interval=ri)
    print("config", codec, w, h, q, ri, "len", len(got), len(want), "equal", got == want)
    if got != want:
        n = min(len(got), len(want))
        first = next(i for i in range(n) if got[i] != want[i])
        print(" first differing byte", first)
        fmt = 0 if codec == tj.UYVY else 1
        _, a = tjd.orc_decode(orc, got, fmt, w, h)
        _, b = tjd.orc_decode(orc, want, fmt, w, h)
        if fmt == 0:
            A, B = a.reshape(h, -1), b.reshape(h, -1)
            for name, sl, bw in (("Y", slice(1, None, 2), 8), ("Cb", slice(0, None, 4), 8), ("Cr", slice(2, None, 4), 8)):
                d = (A[:, sl].astype(int) - B[:, sl].astype(int))
                ys, xs = np.nonzero(d)
                blocks = sorted(set(zip((ys // 8).tolist(), (xs // bw).tolist())))
                print(" ", name, "differing blocks", blocks[:10], "max abs", np.abs(d).max())
                if blocks:
                    by, bx = blocks[0]
                    print(d[by*8:by*8+8, bx*8:bx*8+8])
        else:
            d = a.reshape(h, w, 3).astype(int) - b.reshape(h, w, 3).astype(int)
            ys, xs, cs = np.nonzero(d)
            print("  differing blocks", sorted(set(zip((ys // 8).tolist(), (xs // 8).tolist(), cs.tolist())))[:10], "max", np.abs(d).max())
