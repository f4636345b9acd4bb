#!/bin/bash
# round 2, call O: DXT5-YCoCg decode without F2I / division routine (parity + timing), ncu --set full of v210 -> RG48 direct vs staged
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dxt_decode.py tests/test_vdecompress.py tests/test_real_module.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_o.log
tail -4 gpurun_out/pytest_o.log | cut -c1-600
python - <<'PY'
import torch
from ultragrid_b200 import api
W, H = 7680, 4320
for t, nbytes in ((1, W * H // 2), (6, W * H)):
    blocks = [torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda") for _ in range(6)]
    outs = [torch.empty(W * H * 3, dtype=torch.uint8, device="cuda") for _ in range(6)]
    for i in range(6):
        api.dxt_to_rgb(blocks[i], W, H, t, out=outs[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 24
    for i in range(n):
        api.dxt_to_rgb(blocks[i % 6], W, H, t, out=outs[i % 6])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"dxt{t} decode 8K noise blocks: {us:.1f} us, {(nbytes + W * H * 3) / us / 1e3:.0f} GB/s, frac {(nbytes + W * H * 3) / us / 1e3 / 6490.5:.2f}")
PY
UGB200_LINE_STAGED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_v210_rg48_direct -f python tools/profile_target.py conv 7 27 > gpurun_out/ncu_v210_rg48_direct.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_conv -s 2 -c 1 -o gpurun_out/prof_v210_rg48_staged -f python tools/profile_target.py conv 7 27 > gpurun_out/ncu_v210_rg48_staged.log 2>&1
ls -la gpurun_out/*.ncu-rep
