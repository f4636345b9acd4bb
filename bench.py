#!/usr/bin/env python3
"""Headline benchmark: 7680x4320 frames/s of the fused UYVY -> DXT1 encode (BASELINE.json metric), with the
HBM roofline of the kernel, the end-to-end number through the C ABI with host buffers, and the CPU baseline.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      CPU arm (rank 0 only)

A "step" = one pass of the hot path over a batch of B distinct 8K frames that are resident in HBM (B x 66 MB is
far larger than the 126 MB L2, so nothing is served from cache).  Frames are independent units: with N GPUs each
rank encodes its own B frames per step (weak scaling, no data-path collective); the only collective is the NCCL
scatter of the int32 frame-index assignment, issued one step ahead on a side stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W8K, H8K = 7680, 4320
ALGO_BYTES_UYVY_DXT1 = W8K * H8K * 2 + W8K * H8K // 2  # 2 + 0.5 B/px, SURVEY.md section 8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=48, help="distinct 8K frames per step per GPU")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary kernels / cpu baseline")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every 10 ms while the timed region runs (nvidia-smi as fallback)"""

    def __init__(self, uuid=None, index=0):
        self.uuid, self.index, self.rows, self.stop_flag, self.thread = uuid, index, [], False, None
        self.max_mhz, self.t0, self.t1 = None, None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                for cand in (f"GPU-{self.uuid}", str(self.uuid)):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), reasons, time.perf_counter()))
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self.rows.append(("error", str(e)))

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def window_begin(self):
        self.t0 = time.perf_counter()

    def window_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        good = [r for r in self.rows if r[0] != "error" and (self.t0 is None or self.t0 <= r[2] <= (self.t1 or 1e30))]
        sm = sorted(r[0] for r in good)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({n for r in good for bit, n in names.items() if r[1] & bit})
        err = [r[1] for r in self.rows if r[0] == "error"]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                **({"error": err[0]} if err else {})}


def cpu_port_fps(seconds=12.0, max_frames=6):
    """oracle port of the path (oracle/dxt_oracle.c, OpenMP over block rows) on the host cores, bounded sample"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    orc = util.oracle()
    orc.orc_set_threads(0)
    cores = orc.orc_get_max_threads()
    src = util.rng_bytes(W8K * H8K * 2, 4)
    out = np.zeros(W8K * H8K // 2, dtype=np.uint8)
    orc.orc_uyvy_to_dxt1(src.ctypes.data, out.ctypes.data, W8K, 256, 0)  # warm the thread pool
    n, t0 = 0, time.perf_counter()
    while n < max_frames and (n == 0 or time.perf_counter() - t0 < seconds):
        orc.orc_uyvy_to_dxt1(src.ctypes.data, out.ctypes.data, W8K, H8K, 0)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, cores, f"{n} noise frames 7680x4320 UYVY->DXT1, {dt:.1f} s wall"


def reference_arm(args, rank):
    """the reference has no CPU implementation of DXT (only CUDA); the CPU arm is therefore the oracle port of the
    same path on all host threads.  Rank 0 alone works."""
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    orc = util.oracle()
    orc.orc_set_threads(0)
    cores = orc.orc_get_max_threads()
    src = util.rng_bytes(W8K * H8K * 2, 4)
    out = np.zeros(W8K * H8K // 2, dtype=np.uint8)
    for _ in range(max(args.warmup, 1)):
        orc.orc_uyvy_to_dxt1(src.ctypes.data, out.ctypes.data, W8K, H8K // 4, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):  # one step = one 8K frame (bounded sample of the GPU arm's batch)
        orc.orc_uyvy_to_dxt1(src.ctypes.data, out.ctypes.data, W8K, H8K, 0)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "7680x4320 frames/sec encode (UYVY->DXT1)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "7680x4320 UYVY->DXT1 encode on host cores, 1 noise frame per step",
                   "note": "UltraGrid has no CPU DXT encoder; this is the CPU port of cuda_dxt's arithmetic (oracle/dxt_oracle.c)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} noise frames 7680x4320, OpenMP over block rows"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from ultragrid_b200 import api, compress, sharding  # raises if libugb200.so is missing: no fallback

    assert torch.cuda.is_available(), "bench.py needs a GPU"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.frames, args.steps, max(args.warmup, 3)
    frame_bytes, out_bytes = W8K * H8K * 2, W8K * H8K // 2

    # B distinct noise frames (worst case for DXT: no flat-block shortcut), resident in HBM
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    frames = torch.randint(0, 256, (B, frame_bytes), dtype=torch.uint8, device=dev, generator=g)
    outs = torch.empty((B, out_bytes), dtype=torch.uint8, device=dev)

    # frame-index assignment: rank 0 owns the queue and scatters int32 indices (the only collective)
    comm = torch.cuda.Stream(device=dev)
    assign = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]

    def scatter_assignment(step):
        with torch.cuda.stream(comm):
            sharding.scatter_assignment(step, B, assign[step % 2])  # NCCL scatter from rank 0 (a copy when world == 1)
            ready[step % 2].record(comm)

    fr = [frames[f] for f in range(B)]
    ou = [outs[f] for f in range(B)]

    def run_step(step):
        scatter_assignment(step + 1)                      # next step's assignment travels while this one encodes
        torch.cuda.current_stream().wait_event(ready[step % 2])
        for f in range(B):                               # global frame assign[step%2][f] lives in local slot f
            api.uyvy_to_dxt(fr[f], W8K, H8K, dxt_type=1, out=ou[f])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    clocks = ClockSampler(uuid, local_rank)
    if rank == 0:
        clocks.start()  # NVML initialises while the warm-up runs; only samples inside the timed window are kept
    scatter_assignment(0)
    for s in range(Wm):
        run_step(s)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.window_begin()
    e0.record()
    for s in range(K):
        run_step(Wm + s)
    e1.record()
    barrier()
    clocks.window_end()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    fps = world * B * K / (ms_max * 1e-3)

    # ---- end to end through the reference-facing plugin: compress_init("cuda_dxt:DXT1") / compress_frame / compress_pop with HOST
    # frames (pinned, like UltraGrid's capture buffers can be); H2D + kernel + D2H are inside the timed region, every frame
    compress.set_cuda_devices([local_rank])
    nhost = 6
    h_in = [torch.empty(frame_bytes, dtype=torch.uint8).pin_memory() for _ in range(nhost)]
    for i, b in enumerate(h_in):
        b.copy_(frames[i % B].cpu())
    h_np = [b.numpy() for b in h_in]
    plugin = compress.Compress("cuda_dxt:DXT1")

    inflight = [0]

    def e2e_step():  # frames stream through the module: up to 3 in flight, results popped in order (zero-copy pinned frames)
        for f in range(B):
            plugin.push(h_np[f % nhost], W8K, H8K, 2)  # codec_t UYVY
            inflight[0] += 1
            if inflight[0] == 3:
                view, _, _ = plugin.pop_ref()
                inflight[0] -= 1
                assert view.size == out_bytes
        while inflight[0]:  # the step's last results are on the host before the step ends
            view, _, _ = plugin.pop_ref()
            inflight[0] -= 1

    Ke = max(3, min(K, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()  # the module synchronises its own stream per frame: host wall clock brackets whole frames
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * Ke / float(e2e_t.item())
    plugin.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        per_launch_ms = ms / (B * K)
        achieved = ALGO_BYTES_UYVY_DXT1 / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dxt_uyvy_kernel_8k_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "7680x4320 frames/sec encode (UYVY->DXT1)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "7680x4320 UYVY->DXT1 fused encode (ugb200_uyvy_to_dxt1_async), uniform-noise frames",
                       "frames_per_step_per_gpu": B, "global_frames_per_step": world * B,
                       "l2": f"inputs larger than L2: {B} distinct 66 MB frames per GPU", "parallelism": f"frame-sharded x{world}",
                       "collective": "NCCL scatter of int32 frame indices, one step ahead on a side stream"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "ugb::dxt_uyvy_kernel<1,2,false>", "algorithmic_bytes_per_launch": ALGO_BYTES_UYVY_DXT1,
                         "us_per_launch": per_launch_ms * 1e3, "peak_source": peak_src},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes, "d2h_bytes_per_step": B * out_bytes,
                    "path": "compress_init('cuda_dxt:DXT1'): pinned host UYVY frame -> compress_frame (H2D, fused kernel, D2H into a pooled "
                            "pinned frame) -> compress_pop; asynchronous module, 3 frames in flight on 3 streams"},
            "gpu_launches": B * K,
            "clocks": clk,
        }
        if world == 1 and not args.no_extra:
            def guarded(fn, *a):  # a failing secondary measurement must not cost the headline line
                try:
                    return fn(*a)
                except Exception as e:  # noqa: BLE001
                    return {"error": f"{type(e).__name__}: {e}"[:300]}
            line["extra"] = guarded(extra_kernels, api, torch, dev)
            line["extra"]["jpeg"] = guarded(extra_jpeg, api, compress, torch, dev)
            line["extra"]["decode"] = guarded(extra_decode, api, torch, dev)
            try:
                v, cores, sample = cpu_port_fps()
                line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": None, "kind": "port", "sample": f"failed: {e}"[:200]}
            line["extra"]["cpu_reference_pixfmt"] = guarded(cpu_reference_pixfmt)
            line["extra"]["reference_gpu_kernels"] = guarded(reference_gpu_kernels, torch, dev)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_kernel(torch, fn, iters, warm=3, graph=False):
    """graph=True: the `iters` launches are captured once into a CUDA graph and replayed, so that a kernel of a few microseconds is not
    timed by the Python/ctypes call overhead of its launcher"""
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    if graph:
        s = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for i in range(iters):
                    fn(i)
            g.replay()
            s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            s.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def extra_kernels(api, torch, dev):
    """secondary kernels of the path (kernel-only, device-resident, distinct buffers cycled so that L2 does not help)"""
    from ultragrid_b200 import Codec, vc_get_linesize
    peak, _ = measured_peaks()
    res = {}

    def rec(name, secs, nbytes, px):
        res[name] = {"us": secs * 1e6, "GBps": nbytes / secs / 1e9, "frac_of_peak": nbytes / secs / 1e9 / peak, "fps": 1.0 / secs,
                     "bytes_per_px": nbytes / px}

    def rnd(n, count):
        return [torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev) for _ in range(count)]

    # config 2: 3840x2160 UYVY -> DXT1
    w, h = 3840, 2160
    src, out = rnd(w * h * 2, 12), torch.empty(w * h // 2, dtype=torch.uint8, device=dev)
    rec("uyvy_dxt1_4k", time_kernel(torch, lambda i: api.uyvy_to_dxt(src[i % 12], w, h, out=out), 60, graph=True), w * h * 2.5, w * h)
    del src
    # config 5: 7680x4320 UYVY -> DXT5-YCoCg (fused)
    w, h = W8K, H8K
    src, out = rnd(w * h * 2, 4), torch.empty(w * h, dtype=torch.uint8, device=dev)
    rec("uyvy_dxt5ycocg_8k", time_kernel(torch, lambda i: api.uyvy_to_dxt(src[i % 4], w, h, dxt_type=6, out=out), 12), w * h * 3.0, w * h)
    del src
    # reference-ABI RGB -> DXT1 at 8K (async variant measured through the compat kernel, includes its stream sync)
    w, h = W8K, H8K
    src, out = rnd(w * h * 3, 4), torch.empty(w * h // 2, dtype=torch.uint8, device=dev)
    rec("rgb_dxt1_8k_sync_abi", time_kernel(torch, lambda i: api.compat_to_dxt("cuda_rgb_to_dxt1", src[i % 4], w, h, out=out), 12),
        w * h * 3.5, w * h)
    del src
    # config 4: v210 -> P010 8K
    ls = vc_get_linesize(w, Codec.v210)
    src = [torch.randint(0, 1 << 30, (ls // 4 * h,), dtype=torch.int32, device=dev).view(torch.uint8) for _ in range(4)]
    oy, oc = torch.empty(w * 2 * h, dtype=torch.uint8, device=dev), torch.empty(w * h, dtype=torch.uint8, device=dev)
    rec("v210_p010_8k", time_kernel(torch, lambda i: api.v210_to_p010le(src[i % 4], w, h, out_y=oy, out_c=oc), 40), ls * h + w * h * 3, w * h)
    del src
    # config 1 on GPU + 8K line conversions
    for name, inc, outc, ww, hh, n in (("uyvy_rgb_1080p", Codec.UYVY, Codec.RGB, 1920, 1080, 64), ("uyvy_rgb_8k", Codec.UYVY, Codec.RGB, w, h, 4),
                                       ("rgb_uyvy_8k", Codec.RGB, Codec.UYVY, w, h, 4), ("v210_uyvy_8k", Codec.v210, Codec.UYVY, w, h, 4)):
        src = rnd(vc_get_linesize(ww, inc) * hh, n)
        dst = torch.empty(vc_get_linesize(ww, outc) * hh, dtype=torch.uint8, device=dev)
        rec(name, time_kernel(torch, lambda i: api.pixfmt_convert(inc, outc, src[i % n], ww, hh, dst=dst), 64 if ww < 3000 else 40, graph=ww < 3000),
            (vc_get_linesize(ww, inc) + vc_get_linesize(ww, outc)) * hh, ww * hh)
        del src
    return res


def extra_jpeg(api, compress, torch, dev):
    """second half of the metric: 7680x4320 UYVY -> JPEG (q=90), kernel-only (device-resident) and through the GPUJPEG module"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    res = {}
    orc = util.oracle()
    yy, xx = np.mgrid[0:H8K, 0:W8K]
    rgb = np.stack([xx * 255 // (W8K - 1), yy * 255 // (H8K - 1), (xx + yy) % 256], axis=2).astype(np.uint8)
    rgb = (rgb.astype(np.int16) + np.random.default_rng(1).integers(-6, 7, rgb.shape, dtype=np.int16)).clip(0, 255).astype(np.uint8)
    natural = util.convert_cpu(orc, "orc_convert", 12, 2, rgb.reshape(-1), W8K, H8K)
    del rgb
    inputs = {"natural": torch.from_numpy(natural).cuda(), "testcard": torch.from_numpy(util.testcard_uyvy(W8K, H8K, orc)).cuda(),
              "noise": torch.randint(0, 256, (W8K * H8K * 2,), dtype=torch.uint8, device=dev)}
    enc = api.JpegEncoder()
    for name, src in inputs.items():
        for _ in range(3):  # the encoder sizes its bit buffers (and with them the kernel instantiation) from the previous frame's statistics
            enc.encode_device(src, W8K, H8K, 2, quality=90)
            nbytes = len(enc.result())
        n = 6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            enc.encode_device(src, W8K, H8K, 2, quality=90)
        e1.record()
        enc.result()
        secs = e0.elapsed_time(e1) / n * 1e-3
        res[name] = {"us": secs * 1e6, "fps": 1 / secs, "stream_bytes": nbytes, "algorithmic_GBps": (W8K * H8K * 2 + nbytes) / secs / 1e9}
    nbytes_natural = res["natural"]["stream_bytes"]
    enc.close()
    # two encoders on two streams, frames alternating: scan + compact of frame n run beside the fused kernel of frame n + 1 (what the GPUJPEG
    # module's lanes do on one device); wall clock between two device synchronisations around 24 frames
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    encs = [api.JpegEncoder(stream=st) for st in streams]
    src = inputs["natural"]
    for e in encs:
        for _ in range(3):
            e.encode_device(src, W8K, H8K, 2, quality=90)
            assert len(e.result()) == nbytes_natural
    torch.cuda.synchronize()
    n = 24
    t0 = time.perf_counter()
    for i in range(n):
        encs[i % 2].encode_device(src, W8K, H8K, 2, quality=90)
    torch.cuda.synchronize()
    secs = (time.perf_counter() - t0) / n
    res["natural_two_streams"] = {"us": secs * 1e6, "fps": 1 / secs}
    for e in encs:
        e.result()
        e.close()
    # through the module with pinned host frames (compress_init("GPUJPEG:q=90") / compress_frame / compress_pop): H2D, kernels and the D2H
    # of the stream into the pooled output frame are inside the timed region.  lanes=1 is the reference's single-device shape
    # (push returns when the frame is done), the default keeps 3 frames in flight on one device.
    hosts = [torch.from_numpy(natural).pin_memory().numpy() for _ in range(3)]
    for cfg, key, depth in (("GPUJPEG:q=90:lanes=1", "natural_e2e_module_sync_fps", 1), ("GPUJPEG:q=90", "natural_e2e_module_fps", 3)):
        c = compress.Compress(cfg)
        def run(n):
            inflight, view = 0, None
            for i in range(n):
                c.push(hosts[i % 3], W8K, H8K, 2)
                inflight += 1
                if inflight == depth:
                    view, _, _ = c.pop_ref()
                    inflight -= 1
            while inflight:
                view, _, _ = c.pop_ref()
                inflight -= 1
            return view

        run(6)  # every lane has its encoder buffers, the pool its pinned frames
        n = 24
        t0 = time.perf_counter()
        view = run(n)
        res[key] = n / (time.perf_counter() - t0)
        assert view.size == nbytes_natural
        c.close()
    return res


def extra_decode(api, torch, dev):
    """decode side and one planar converter, 8K: kernel-only where the input is device-resident; the JPEG decoder takes a
    HOST stream (parse + upload + kernels), so that number is wall clock per frame"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    peak, _ = measured_peaks()
    res = {}
    w, h = W8K, H8K
    for t, name in ((1, "dxt1_rgb_8k"), (6, "dxt5ycocg_rgb_8k")):
        nb = w * h // (2 if t == 1 else 1)
        blocks = [torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev) for _ in range(4)]
        out = torch.empty(w * h * 3, dtype=torch.uint8, device=dev)
        secs = time_kernel(torch, lambda i: api.dxt_to_rgb(blocks[i % 4], w, h, t, out=out), 12)
        res[name] = {"us": secs * 1e6, "fps": 1 / secs, "GBps": (nb + w * h * 3) / secs / 1e9, "frac_of_peak": (nb + w * h * 3) / secs / 1e9 / peak}
        del blocks
    src = [torch.randint(0, 256, (w * h * 2,), dtype=torch.uint8, device=dev) for _ in range(4)]
    y, c = torch.empty(w * h, dtype=torch.uint8, device=dev), torch.empty(w * h // 2, dtype=torch.uint8, device=dev)
    secs = time_kernel(torch, lambda i: api.to_planar("uyvy_to_nv12", src[i % 4], w, h, [y, c], [w, w]), 20)
    res["uyvy_nv12_8k"] = {"us": secs * 1e6, "GBps": w * h * 3.5 / secs / 1e9, "frac_of_peak": w * h * 3.5 / secs / 1e9 / peak}
    del src
    orc = util.oracle()
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([xx * 255 // (w - 1), yy * 255 // (h - 1), (xx + yy) % 256], axis=2).astype(np.uint8)
    rgb = (rgb.astype(np.int16) + np.random.default_rng(1).integers(-6, 7, rgb.shape, dtype=np.int16)).clip(0, 255).astype(np.uint8)
    enc = api.JpegEncoder()
    enc.encode_device(torch.from_numpy(util.convert_cpu(orc, "orc_convert", 12, 2, rgb.reshape(-1), w, h)).cuda(), w, h, 2, quality=90)
    stream = enc.result()
    enc.close()
    dec = api.JpegDecoder()
    out = dec.decode(stream, 2, device=True)
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        dec.decode(stream, 2, device=True, out=out, sync=False)
    torch.cuda.synchronize()
    res["jpeg_decode_natural_8k"] = {"ms_wall_per_frame": (time.perf_counter() - t0) / n * 1e3, "stream_bytes": len(stream),
                                     "output": "UYVY on the device; host stream in: marker scan + staged upload + 4 kernels"}
    dec.close()
    return res


def reference_gpu_kernels(torch, dev):
    """for the DXT configs the reference beside the product is its own CUDA path — the UNMODIFIED cuda_dxt.cu built for
    sm_100a (oracle/_ref/libcuda_dxt_ref.so), i.e. cuda_yuv422_to_yuv444 + cuda_yuv_to_dxt{1,6} as src/video_compress/cuda_dxt.cpp
    :229,257 runs them (each call synchronises its stream, cuda_dxt.cu:759).  Baseline only; never part of `value`."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ref = util.ref_gpu()
    if ref is None:
        return {"unavailable": "oracle/_ref/libcuda_dxt_ref.so not present"}
    w, h = W8K, H8K
    srcs = [torch.randint(0, 256, (w * h * 2,), dtype=torch.uint8, device=dev) for _ in range(3)]
    mid = torch.empty(w * h * 3, dtype=torch.uint8, device=dev)
    out = torch.empty(w * h, dtype=torch.uint8, device=dev)
    res = {}
    for name, fn in (("uyvy_dxt1_8k", ref.cuda_yuv_to_dxt1), ("uyvy_dxt5ycocg_8k", ref.cuda_yuv_to_dxt6)):
        def step(i):
            ref.cuda_yuv422_to_yuv444(ctypes.c_void_p(srcs[i % 3].data_ptr()), ctypes.c_void_p(mid.data_ptr()), w * h, None)
            fn(ctypes.c_void_p(mid.data_ptr()), ctypes.c_void_p(out.data_ptr()), w, h, None)
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 6
        for i in range(n):
            step(i)
        torch.cuda.synchronize()
        secs = (time.perf_counter() - t0) / n
        res[name] = {"us": secs * 1e6, "fps": 1 / secs}
    return res


def cpu_reference_pixfmt():
    """the reference's own CPU pixfmt_conv path (unmodified objects in oracle/_ref) on this host: UYVY->RGB, 1 thread and all
    cores through parallel_pix_conv (src/utils/parallel_conv.c:64-85)"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ref = util.ref_cpu()
    if ref is None:
        return {"unavailable": "oracle/_ref/libugref.so not present"}
    src = util.rng_bytes(W8K * H8K * 2, 9)
    dst = np.zeros(W8K * H8K * 3, dtype=np.uint8)
    out = {"cores": os.cpu_count(), "flags": "-O3 -msse4.1 (tools/Makefile)"}
    for label, fn in (("1_thread", lambda: ref.ref_convert(2, 12, dst.ctypes.data, W8K * 3, src.ctypes.data, W8K * 2, W8K * 3, H8K, 0, 8, 16)),
                      ("all_cores", lambda: ref.ref_convert_parallel(2, 12, dst.ctypes.data, W8K * 3, src.ctypes.data, W8K * 2, H8K, 0))):
        fn()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        out[f"uyvy_rgb_8k_{label}_ms"] = best * 1e3
    return out


if __name__ == "__main__":
    main()
