"""The reference's own unit tests for this path (test/codec_conversions_test.cpp), same sizes and patterns, run against the CPU
restatement (always) and the CUDA kernels (-m gpu):
  codec_conversion_test_testcard_uyvy_to_i420 (:28-86)   fixed 'u y v Y' pattern, odd sizes
  codec_conversion_test_y216_to_p010le (:91-170)         16-bit pattern, odd sizes"""
import numpy as np
import pytest

import planar_cases as pc
import util

I420_SIZES = [(1, 2), (2, 1), (16, 1), (16, 16), (127, 255)]
P010_SIZES = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (15, 1), (16, 1), (16, 16), (127, 255), (128, 256), (255, 1), (255, 2)]


def run(name, src, w, h, planes, backend):
    """planes: [(linesize, rows)]; returns the plane arrays"""
    outs = [np.full(ls * rows + 64, 0xCD, np.uint8) for ls, rows in planes]
    srcp = np.concatenate([src, np.zeros(4096, np.uint8)])
    if backend == "oracle":
        d = pc.ToPlanarData()
        d.width, d.height, d.in_data = w, h, srcp.ctypes.data
        for i, ((ls, _), o) in enumerate(zip(planes, outs)):
            d.out_data[i], d.out_linesize[i] = o.ctypes.data, ls
        fn = getattr(util.oracle(), "orc_" + name)
        fn.argtypes, fn.restype = [pc.ToPlanarData], None
        fn(d)
        return outs
    import torch
    from ultragrid_b200 import api
    t = [torch.from_numpy(o).cuda() for o in outs]
    api.to_planar(name, torch.from_numpy(srcp).cuda(), w, h, t, [ls for ls, _ in planes])
    torch.cuda.synchronize()
    return [x.cpu().numpy() for x in t]


def check_uyvy_to_i420(backend):
    for w, h in I420_SIZES:
        row = np.tile(np.frombuffer(b"uyvY", np.uint8), (w + 1) // 2)
        src = np.tile(row, h)
        cw, ch = (w + 1) // 2, (h + 1) // 2
        y, u, v = run("uyvy_to_i420", src, w, h, [(w, h), (cw, ch), (cw, ch)], backend)
        want_y = np.tile(np.array([ord("y"), ord("Y")], np.uint8), (w + 1) // 2)[:w]
        assert np.array_equal(y[:w * h].reshape(h, w), np.tile(want_y, (h, 1))), (w, h)
        assert (u[:cw * ch] == ord("u")).all() and (v[:cw * ch] == ord("v")).all(), (w, h)


def check_y216_to_p010le(backend):
    u, y1, v, y2 = ord("U") << 8 | ord("u"), ord("Y") << 8 | ord("1"), ord("V") << 8 | ord("v"), ord("Y") << 8 | ord("2")
    pat = np.array([y1, u, y2, v], np.uint16)
    for w, h in P010_SIZES:
        we = (w + 1) & ~1
        src = np.tile(np.tile(pat, we // 2), h).view(np.uint8)
        yp, cp = run("y216_to_p010le", src, w, h, [(w * 2, h), (we * 2, (h + 1) // 2)], backend)
        assert np.array_equal(yp[:w * h * 2].view(np.uint16).reshape(h, w), np.tile(np.tile(pat[[0, 2]], we // 2)[:w], (h, 1))), (w, h)
        assert np.array_equal(cp[:we * ((h + 1) // 2) * 2].view(np.uint16).reshape(-1, we), np.tile(np.tile(pat[[1, 3]], we // 2), ((h + 1) // 2, 1))), (w, h)


def test_reference_unit_tests_on_the_restatement():
    check_uyvy_to_i420("oracle")
    check_y216_to_p010le("oracle")


@pytest.mark.gpu
def test_reference_unit_tests_on_the_gpu():
    check_uyvy_to_i420("gpu")
    check_y216_to_p010le("gpu")
