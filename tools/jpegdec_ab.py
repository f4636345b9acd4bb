"""A/B timing of the JPEG decoder's marker scan (host threads vs device kernels), 8K natural UYVY stream: pipelined wall time per frame (no
synchronisation between frames, as bench.py measures it), device time per frame (CUDA events), and the latency of one synchronous decode.
Each variant in a child process (the mode is read when the decoder is created)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import numpy as np, torch
    from ultragrid_b200 import api
    W, H = 7680, 4320
    dev = torch.device("cuda:0")
    xx = torch.arange(W, device=dev, dtype=torch.int32).view(1, W)
    yy = torch.arange(H, device=dev, dtype=torch.int32).view(H, 1)
    base = torch.stack([(xx * 255 // (W - 1)).expand(H, W), (yy * 255 // (H - 1)).expand(H, W), (xx + yy) % 256], dim=2)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    rgb = (base + torch.randint(-6, 7, base.shape, dtype=torch.int32, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8).reshape(-1)
    enc = api.JpegEncoder()
    codec = 12 if os.environ.get("JPEGDEC_AB_RGB") else 2  # RGB: three scans, as GPUJPEG stores RGB (config 3 of BASELINE.json on the receiving side)
    enc.encode_device(rgb if codec == 12 else api.pixfmt_convert(12, 2, rgb, W, H), W, H, codec, quality=90)
    stream = enc.result()
    dec = api.JpegDecoder()
    out = dec.decode(stream, codec, device=True)
    for _ in range(4):
        dec.decode(stream, codec, device=True, out=out, sync=False)
    torch.cuda.synchronize()
    n = 64
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        dec.decode(stream, codec, device=True, out=out, sync=False)
    t_host = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    lat = []
    for _ in range(8):
        t0 = time.perf_counter()
        dec.decode(stream, codec, device=True, out=out, sync=True)
        lat.append(time.perf_counter() - t0)
    print("stream %d B: pipelined wall %.3f ms/frame, host side %.3f ms/frame, device span %.3f ms/frame, synchronous latency %.3f ms (median of 8)"
          % (len(stream), wall / n * 1e3, t_host / n * 1e3, e0.elapsed_time(e1) / n, sorted(lat)[4] * 1e3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        out = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out")
        variants = (("host scan", {"UGB200_JPEG_MARKER_SCAN": "host"}), ("host scan, plain stores", {"UGB200_JPEG_MARKER_SCAN": "host", "UGB200_JPEG_STAGE": "plain"}),
                    ("device scan (default for this stream)", {}), ("device scan, plain stores", {"UGB200_JPEG_STAGE": "plain"}))
        if len(sys.argv) > 1 and sys.argv[1] == "short":  # the two scans only, UYVY and RGB streams
            variants = (("host scan", {"UGB200_JPEG_MARKER_SCAN": "host"}), ("device scan (default for this stream)", {}),
                        ("RGB host scan", {"UGB200_JPEG_MARKER_SCAN": "host", "JPEGDEC_AB_RGB": "1"}), ("RGB device scan (default for this stream)", {"JPEGDEC_AB_RGB": "1"}))
        for name, env in variants:
            print("==", name, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env={**os.environ, **env}, timeout=400)
            tag = name.replace(" ", "_").replace(",", "").replace("(", "").replace(")", "")
            for level in ("1", "2"):  # 1: host laps of the pipelined loop, 2: device stages (serialised)
                subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env={**os.environ, **env, "UGB200_JPEG_TIMING": level}, timeout=400,
                               stdout=subprocess.DEVNULL, stderr=open(os.path.join(out, "jpegdec_%s_t%s.txt" % (tag, level)), "w"))
