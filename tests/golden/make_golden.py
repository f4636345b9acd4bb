"""Generates tests/golden/pixfmt_golden.npz from the UNMODIFIED reference objects (oracle/_ref/libugref.so,
built from /root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixtures are small (a few hundred KB) and are what pins the oracle and the CUDA path on the GPU box, where
/root/reference does not exist."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import util  # noqa: E402
from test_oracle_pinning import PAIRS  # noqa: E402

ref = util.ref_cpu()
assert ref is not None, "build oracle/_ref first: make -C oracle ref"
out = {}
for n, (inc, outc) in enumerate(PAIRS):
    for w, h in ((50, 3), (97, 2)):
        src = util.rng_bytes(ref.ref_vc_get_linesize(w, inc) * h, 7000 + n)
        dst = util.convert_cpu(ref, "ref_convert", inc, outc, src, w, h, linesize=ref.ref_vc_get_linesize)
        k = f"c{inc}_{outc}_{w}x{h}"
        out[k + "_src"], out[k + "_dst"], out[k + "_meta"] = src, dst, np.array([inc, outc, w, h])
w, h = 100, 7
ls = 102 * 2 + 12
src = util.v210_noise(w, h, 4242)
y = np.zeros(ls * h, dtype=np.uint8)
c = np.zeros(ls * ((h + 1) // 2), dtype=np.uint8)
ref.ref_v210_to_p010le(w, h, y.ctypes.data, ls, c.ctypes.data, ls, src.ctypes.data)
out["p010_src"], out["p010_y"], out["p010_c"], out["p010_meta"] = src, y, c, np.array([w, h, ls])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pixfmt_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out) // 3, "cases")
