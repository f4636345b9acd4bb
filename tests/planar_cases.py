"""Shared description of the packed<->planar whole-buffer converters (src/to_planar.h, src/from_planar.h of the reference):
buffer geometry per function, input generation, and runners for the CPU libraries (oracle: orc_<name>, reference: <name>,
both take the struct BY VALUE like the reference) and for the GPU C ABI (ugb200_<name>, pointer to the same struct)."""
import ctypes

import numpy as np

FILL = 0xCD


class ToPlanarData(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("out_data", ctypes.c_void_p * 4),
                ("out_linesize", ctypes.c_uint * 4), ("in_data", ctypes.c_void_p)]


class FromPlanarData(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("out_data", ctypes.c_void_p), ("out_pitch", ctypes.c_uint),
                ("in_data", ctypes.c_void_p * 4), ("in_linesize", ctypes.c_uint * 4), ("in_depth", ctypes.c_int),
                ("log2_chroma_h", ctypes.c_int), ("rgb_shift", ctypes.c_int * 3)]


def _c2(w):
    return (w + 1) // 2


def _ce(w):
    return (w + 1) & ~1


# name -> (input row bytes, [(plane row bytes, plane rows)])
TO_PLANAR = {
    "y216_to_p010le": (lambda w: _c2(w) * 8, lambda w, h: [(2 * w, h), (2 * _ce(w), _c2(h))]),
    "uyvy_to_nv12": (lambda w: 2 * w, lambda w, h: [(w, h), (_ce(w), _c2(h))]),
    "uyvy_to_i420": (lambda w: _c2(w) * 4, lambda w, h: [(w, h), (_c2(w), _c2(h)), (_c2(w), _c2(h))]),
    "rgba_to_bgra": (lambda w: 4 * w, lambda w, h: [(4 * w, h)]),
    "vuya_to_i444": (lambda w: 4 * w, lambda w, h: [(w, h)] * 3),
    "r12l_to_gbrp12le": (lambda w: (w + 7) // 8 * 36, lambda w, h: [((w + 7) // 8 * 16, h)] * 3),
    "r12l_to_gbrp16le": (lambda w: (w + 7) // 8 * 36, lambda w, h: [((w + 7) // 8 * 16, h)] * 3),
    "r12l_to_rgbp12le": (lambda w: (w + 7) // 8 * 36, lambda w, h: [((w + 7) // 8 * 16, h)] * 3),
}

# name -> (fixed depth or None (= XX: uses in_depth) or 8, planes(w, h, bytes-per-sample) -> [(row bytes, rows)], out row bytes(w), needs in_depth list)
def _p444(w, h, b):
    return [(w * b, h)] * 3


def _p422(w, h, b):
    return [(w * b, h), (_c2(w) * b, h), (_c2(w) * b, h)]


def _p420(w, h, b):
    return [(w * b, h), (_c2(w) * b, _c2(h)), (_c2(w) * b, _c2(h))]


FROM_PLANAR = {
    "gbrap_to_rgb": (8, _p444, lambda w: 3 * w), "gbrap_to_rgba": (8, lambda w, h, b: [(w, h)] * 4, lambda w: 4 * w),
    "gbrp10le_to_rgb": (10, _p444, lambda w: 3 * w), "gbrp12le_to_rgb": (12, _p444, lambda w: 3 * w), "gbrp16le_to_rgb": (16, _p444, lambda w: 3 * w),
    "rgbpXX_to_rgb": (None, _p444, lambda w: 3 * w),
    "gbrp10le_to_rgba": (10, _p444, lambda w: 4 * w), "gbrp12le_to_rgba": (12, _p444, lambda w: 4 * w), "gbrp16le_to_rgba": (16, _p444, lambda w: 4 * w),
    "gbrp10le_to_rg48": (10, _p444, lambda w: 6 * w), "gbrp12le_to_rg48": (12, _p444, lambda w: 6 * w), "gbrp16le_to_rg48": (16, _p444, lambda w: 6 * w),
    "rgbpXXle_to_rg48": (None, _p444, lambda w: 6 * w),
    "gbrp10le_to_r10k": (10, _p444, lambda w: 4 * w), "gbrp12le_to_r10k": (12, _p444, lambda w: 4 * w), "gbrp16le_to_r10k": (16, _p444, lambda w: 4 * w),
    "rgbpXXle_to_r10k": (None, _p444, lambda w: 4 * w),
    "gbrp12le_to_r12l": (12, _p444, lambda w: (w + 7) // 8 * 36), "gbrp16le_to_r12l": (16, _p444, lambda w: (w + 7) // 8 * 36),
    "rgbpXXle_to_r12l": (None, _p444, lambda w: (w + 7) // 8 * 36),
    "yuv444p_to_vuya": (8, _p444, lambda w: 4 * w),
    "yuv420p_to_uyvy": (8, _p420, lambda w: 2 * _ce(w)),
    "yuv420_to_i420": (8, _p420, None),
    "yuv422p_to_uyvy": (8, _p422, lambda w: 2 * _ce(w)), "yuv422p_to_yuyv": (8, _p422, lambda w: 2 * _ce(w)),
    "yuv422pXX_to_uyvy": (None, _p422, lambda w: 2 * _ce(w)), "yuv422p10le_to_uyvy": (10, _p422, lambda w: 2 * _ce(w)),
    "yuv422p10le_to_v210": (10, _p422, lambda w: (w + 47) // 48 * 128),
}

XX_DEPTHS = {"rgbpXX_to_rgb": (8, 10, 12, 16), "rgbpXXle_to_rg48": (10, 12, 16), "rgbpXXle_to_r10k": (10, 12, 16), "rgbpXXle_to_r12l": (12, 16),
             "yuv422pXX_to_uyvy": (8, 10, 12, 16)}


def _pad(nbytes, mode):
    """row pitch: mode 0 = 16-aligned tight, 1 = 16-aligned with 32 B of padding, 2 = deliberately only 2-aligned"""
    if mode == 0:
        return (nbytes + 15) // 16 * 16
    if mode == 1:
        return (nbytes + 15) // 16 * 16 + 32
    return (nbytes + 1) // 2 * 2 + 6


class Case:
    """Host-side buffers of one conversion; run with .run_cpu(lib, prefix) or .run_gpu(api)."""

    def __init__(self, name, w, h, seed=1, mode=0, depth=None, valid_bits=True, shifts=(0, 8, 16)):
        self.name, self.w, self.h, self.mode = name, w, h, mode
        self.to = name in TO_PLANAR
        rng = np.random.default_rng(seed)
        slack = 4096
        if self.to:
            in_ls_f, planes_f = TO_PLANAR[name]
            self.in_ls = in_ls_f(w)
            self.src = np.zeros(self.in_ls * h + slack, dtype=np.uint8)
            self.src[:self.in_ls * h] = rng.integers(0, 256, self.in_ls * h, dtype=np.uint8)
            self.planes = [(_pad(rb, mode), rows) for rb, rows in planes_f(w, h)]
            self.depth = None
        else:
            fixed, planes_f, out_f = FROM_PLANAR[name]
            self.depth = fixed if fixed is not None else depth
            b = 1 if self.depth == 8 else 2
            self.shifts = shifts
            self.in_planes = []
            for rb, rows in planes_f(w, h, b):
                ls = _pad(rb, mode)
                if name.startswith("gbrap"):  # the reference indexes every plane with in_linesize[0]
                    ls = _pad(w, mode)
                a = rng.integers(0, 256, ls * rows + slack, dtype=np.uint8)
                if b == 2 and valid_bits:
                    v = a[:(ls * rows) // 2 * 2].view(np.uint16)
                    v &= (1 << self.depth) - 1
                a[ls * rows:] = 0
                self.in_planes.append((a, ls, rows))
            if name == "yuv420_to_i420":
                self.out_pitch, self.out_size = 0, w * h + 2 * (w // 2) * (h // 2)
            else:
                self.out_pitch = _pad(out_f(w), mode)
                self.out_size = self.out_pitch * h

    # ---- buffers -----------------------------------------------------------------------------------------------
    def alloc_out(self):
        if self.to:
            return [np.full(ls * rows + 64, FILL, dtype=np.uint8) for ls, rows in self.planes]
        return [np.full(self.out_size + 64, FILL, dtype=np.uint8)]

    def struct(self, in_ptrs, out_ptrs):
        if self.to:
            d = ToPlanarData()
            d.width, d.height = self.w, self.h
            for i, (ls, _) in enumerate(self.planes):
                d.out_data[i], d.out_linesize[i] = out_ptrs[i], ls
            d.in_data = in_ptrs[0]
            return d
        d = FromPlanarData()
        d.width, d.height = self.w, self.h
        d.out_data, d.out_pitch = out_ptrs[0], self.out_pitch
        for i, (_, ls, _) in enumerate(self.in_planes):
            d.in_data[i], d.in_linesize[i] = in_ptrs[i], ls
        d.in_depth = self.depth or 0
        d.log2_chroma_h = 1 if "420" in self.name else 0
        d.rgb_shift[0], d.rgb_shift[1], d.rgb_shift[2] = self.shifts
        return d

    def inputs(self):
        return [self.src] if self.to else [a for a, _, _ in self.in_planes]

    # ---- runners -----------------------------------------------------------------------------------------------
    def run_cpu(self, lib, prefix):
        fn = getattr(lib, prefix + self.name)
        fn.argtypes = [ToPlanarData if self.to else FromPlanarData]
        fn.restype = None
        ins, outs = self.inputs(), self.alloc_out()
        fn(self.struct([a.ctypes.data for a in ins], [o.ctypes.data for o in outs]))
        return outs

    def run_gpu(self, lib, torch, offset=0):
        """lib = the loaded libugb200 (ctypes); offset misaligns every device pointer by that many bytes"""
        fn = getattr(lib, "ugb200_" + self.name)
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        ins = [torch.from_numpy(np.concatenate([np.zeros(offset, np.uint8), a])).cuda() for a in self.inputs()]
        outs_h = self.alloc_out()
        outs = [torch.from_numpy(np.concatenate([np.full(offset, FILL, np.uint8), o])).cuda() for o in outs_h]
        d = self.struct([t.data_ptr() + offset for t in ins], [t.data_ptr() + offset for t in outs])
        rc = fn(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, (self.name, rc)
        torch.cuda.synchronize()
        return [t.cpu().numpy()[offset:] for t in outs]


def all_cases():
    """(name, depth) pairs of every exported function"""
    out = [(n, None) for n in TO_PLANAR]
    for n, (fixed, _, _) in FROM_PLANAR.items():
        if fixed is None:
            out += [(n, dpt) for dpt in XX_DEPTHS[n]]
        else:
            out.append((n, None))
    return out
