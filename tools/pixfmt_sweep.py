"""times every line converter (decoders[] row) at 7680x4320 on the device: us, GB/s of algorithmic bytes (in + out), fraction of the measured peak"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ultragrid_b200 import api, Codec, vc_get_linesize
from test_oracle_pinning import PAIRS
W, H = 7680, 4320
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6490.5) if os.path.exists("MEASURED_PEAKS.json") else 6490.5
rows = []
MODES = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [-1]  # ugb200_pixfmt_staged_mode: -1 default, 0 direct, 1 staged
for inc, outc in PAIRS:
    if inc == outc:
        continue
    ls_i, ls_o = vc_get_linesize(W, inc), vc_get_linesize(W, outc)
    src = [torch.randint(0, 256, (ls_i * H + 4096,), dtype=torch.uint8, device="cuda") for _ in range(3)]
    dst = torch.empty(ls_o * H + 4096, dtype=torch.uint8, device="cuda")
    res = []
    for mode in MODES:
        api.pixfmt_staged_mode(mode)
        for i in range(3):
            api.pixfmt_convert(inc, outc, src[i % 3], W, H, dst=dst)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 12
        for i in range(n):
            api.pixfmt_convert(inc, outc, src[i % 3], W, H, dst=dst)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / n * 1e3)
    api.pixfmt_staged_mode(-1)
    us = res[0]
    gb = (ls_i + ls_o) * H / us / 1e3
    rows.append((Codec(inc).name, Codec(outc).name, us, gb, gb / peak, res[1:]))
    del src, dst
rows.sort(key=lambda r: r[4])
extra = "".join(f" us (mode {m}) |" for m in MODES[1:])
print(f"| in | out | us | GB/s | frac |{extra}\n|---|---|---|---|---|" + "---|" * (len(MODES) - 1))
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]:.1f} | {r[3]:.0f} | {r[4]:.2f} |" + "".join(f" {u:.1f} |" for u in r[5]))
