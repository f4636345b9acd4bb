"""Small driver for ncu: runs each hot kernel a few times on distinct 8K/4K buffers (no timing here)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultragrid_b200 import api, Codec, vc_get_linesize

which = sys.argv[1] if len(sys.argv) > 1 else "dxt1"
STREAM = len(sys.argv) > 2 and sys.argv[2] == "stream"  # 8 distinct inputs and outputs, 12 launches: steady-state (in-stream) DRAM traffic per launch
NB = 8 if STREAM else 3
NL = 12 if STREAM else 6
dev = torch.device("cuda", 0)
W, H = 7680, 4320
if which == "dxt1":
    src = [torch.randint(0, 256, (W * H * 2,), dtype=torch.uint8, device=dev) for _ in range(NB)]
    out = [torch.empty(W * H // 2, dtype=torch.uint8, device=dev) for _ in range(NB)]
    for i in range(NL):
        api.uyvy_to_dxt(src[i % NB], W, H, out=out[i % NB])
elif which == "dxt6":
    src = [torch.randint(0, 256, (W * H * 2,), dtype=torch.uint8, device=dev) for _ in range(NB)]
    out = [torch.empty(W * H, dtype=torch.uint8, device=dev) for _ in range(NB)]
    for i in range(NL):
        api.uyvy_to_dxt(src[i % NB], W, H, dxt_type=6, out=out[i % NB])
elif which == "p010":
    ls = vc_get_linesize(W, Codec.v210)
    src = [torch.randint(0, 1 << 30, (ls // 4 * H,), dtype=torch.int32, device=dev).view(torch.uint8) for _ in range(NB)]
    oy = [torch.empty(W * 2 * H, dtype=torch.uint8, device=dev) for _ in range(NB)]
    oc = [torch.empty(W * H, dtype=torch.uint8, device=dev) for _ in range(NB)]
    for i in range(NL):
        api.v210_to_p010le(src[i % NB], W, H, out_y=oy[i % NB], out_c=oc[i % NB])
elif which == "uyvy_rgb":
    src = [torch.randint(0, 256, (W * H * 2,), dtype=torch.uint8, device=dev) for _ in range(3)]
    dst = torch.empty(W * H * 3, dtype=torch.uint8, device=dev)
    for i in range(6):
        api.pixfmt_convert(Codec.UYVY, Codec.RGB, src[i % 3], W, H, dst=dst)
elif which in ("jpeg", "jpeg_rgb"):  # ramps + noise ("natural") frames made on the device, as bench.py's JPEG workloads
    xx = torch.arange(W, device=dev, dtype=torch.int32).view(1, W)
    yy = torch.arange(H, device=dev, dtype=torch.int32).view(H, 1)
    base = torch.stack([(xx * 255 // (W - 1)).expand(H, W), (yy * 255 // (H - 1)).expand(H, W), (xx + yy) % 256], dim=2)
    g = torch.Generator(device=dev)
    frames = []
    for k in range(4 if STREAM else 2):
        g.manual_seed(k)
        rgb = (base + torch.randint(-6, 7, base.shape, dtype=torch.int32, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8).reshape(-1)
        frames.append(rgb if which == "jpeg_rgb" else api.pixfmt_convert(12, 2, rgb, W, H))
    codec = 12 if which == "jpeg_rgb" else 2
    enc = api.JpegEncoder()
    for i in range(NL):
        enc.encode_device(frames[i % len(frames)], W, H, codec, quality=90)
    print("jpeg bytes", enc.result_size())
elif which == "conv":
    inc, outc = int(sys.argv[2]), int(sys.argv[3])
    ls_i, ls_o = vc_get_linesize(W, inc), vc_get_linesize(W, outc)
    src = [torch.randint(0, 256, (ls_i * H + 4096,), dtype=torch.uint8, device=dev) for _ in range(3)]
    dst = torch.empty(ls_o * H + 4096, dtype=torch.uint8, device=dev)
    for i in range(4):
        api.pixfmt_convert(inc, outc, src[i % 3], W, H, dst=dst)
elif which in ("jpegdec", "dxtdec"):
    import time
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import util
    if which == "dxtdec":
        for t in (1, 6):
            blocks = torch.randint(0, 256, (W * H // (2 if t == 1 else 1),), dtype=torch.uint8, device=dev)
            out = torch.empty(W * H * 3, dtype=torch.uint8, device=dev)
            for i in range(3):
                api.dxt_to_rgb(blocks, W, H, t, out=out)
    else:
        orc = util.oracle()
        yy, xx = np.mgrid[0:H, 0:W]
        rgb = np.stack([xx * 255 // (W - 1), yy * 255 // (H - 1), (xx + yy) % 256], axis=2).astype(np.uint8)
        rgb = (rgb.astype(np.int16) + np.random.default_rng(1).integers(-6, 7, rgb.shape, dtype=np.int16)).clip(0, 255).astype(np.uint8)
        enc = api.JpegEncoder()
        enc.encode_device(torch.from_numpy(util.convert_cpu(orc, "orc_convert", 12, 2, rgb.reshape(-1), W, H)).cuda(), W, H, 2, quality=90)
        stream = enc.result()
        dec = api.JpegDecoder()
        for i in range(4):
            t0 = time.perf_counter()
            dec.decode(stream, 2, device=True)
            print("decode wall ms", (time.perf_counter() - t0) * 1e3)
        info = api.JpegImageInfo()
        buf = (__import__("ctypes").c_uint8 * len(stream)).from_buffer_copy(stream)
        t0 = time.perf_counter()
        for i in range(10):
            api._L.ugb200_jpeg_get_image_info(buf, len(stream), __import__("ctypes").byref(info))
        print("header probe ms", (time.perf_counter() - t0) * 100)
torch.cuda.synchronize()
