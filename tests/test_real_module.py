"""The compress modules in UltraGrid's REAL ABI (ultragrid_b200/csrc/module/ug_module.cpp, compiled against the reference's own headers) loaded into the
UNMODIFIED compress framework of the reference (src/lib_common.cpp + src/video_compress.cpp + frame pool + module tree, compiled from the reference
tree into oracle/_ref/libugframework.so): dlopen like open_all(), registration through the reference's REGISTER_MODULE / register_library, lookup
through load_library(), frames through compress_init / compress_frame / compress_pop with real struct video_frame's."""
import ctypes
import os

import numpy as np
import pytest

import util

RGBA, UYVY, RGB, DXT1, DXT5, JPEG = 1, 2, 12, 9, 11, 13
FW = os.path.join(util.ROOT, "oracle", "_ref", "libugframework.so")
MODS = [os.path.join(util.ROOT, "ultragrid_b200", "modules", f"ultragrid_vcompress_{m}.so") for m in ("cuda_dxt", "gpujpeg")]
DEC_MODS = [os.path.join(util.ROOT, "ultragrid_b200", "modules", f"ultragrid_vdecompress_{m}.so") for m in ("gpujpeg", "gpujpeg_to_dxt", "dxt_cuda")]
_fw = None


def framework():
    global _fw
    if _fw is not None:
        return _fw
    if not os.path.exists(FW) or not all(os.path.exists(m) for m in MODS + DEC_MODS):
        pytest.skip("reference framework / real-ABI modules not built (reference tree absent)")
    fw = ctypes.CDLL(FW, mode=ctypes.RTLD_GLOBAL)  # the module libraries resolve register_library, vf_*, video_frame_pool, cuda_devices ... here
    fw.fwd_load_module.argtypes = [ctypes.c_char_p]
    fw.fwd_has_module.argtypes = [ctypes.c_char_p]
    fw.fwd_init.argtypes, fw.fwd_init.restype = [ctypes.c_char_p], ctypes.c_void_p
    fw.fwd_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    fw.fwd_pop.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int),
                           ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
    fw.fwd_done.argtypes = [ctypes.c_void_p]
    fw.fwd_has_decompress_module.argtypes = [ctypes.c_char_p]
    fw.fwd_dec_init.argtypes, fw.fwd_dec_init.restype = [ctypes.c_int, ctypes.c_int], ctypes.c_void_p
    fw.fwd_dec_reconfigure.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8
    fw.fwd_dec_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    fw.fwd_dec_accepts_corrupted.argtypes = [ctypes.c_void_p]
    fw.fwd_dec_done.argtypes = [ctypes.c_void_p]
    assert not fw.fwd_has_module(b"cuda_dxt") and not fw.fwd_has_module(b"gpujpeg")  # the framework alone knows neither
    assert not fw.fwd_has_decompress_module(b"gpujpeg") and not fw.fwd_has_decompress_module(b"dxt_cuda")
    for m in MODS + DEC_MODS:
        assert fw.fwd_load_module(m.encode()) == 0, m
    _fw = fw
    return fw


def pop(fw, st, cap):
    out = np.empty(cap, np.uint8)
    n, codec, seq, w, h = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_uint(), ctypes.c_uint(), ctypes.c_uint()
    rc = fw.fwd_pop(st, out.ctypes.data, cap, ctypes.byref(n), ctypes.byref(codec), ctypes.byref(seq), ctypes.byref(w), ctypes.byref(h))
    if rc == 1:
        return None
    assert rc == 0
    return out[:n.value], codec.value, seq.value, w.value, h.value


def test_modules_register_with_the_reference_registry():
    fw = framework()
    assert fw.fwd_has_module(b"cuda_dxt") and fw.fwd_has_module(b"gpujpeg") and fw.fwd_has_module(b"GPUJPEG")
    assert not fw.fwd_has_module(b"libavcodec")


@pytest.mark.parametrize("cfg", ["cuda_dxt", "cuda_dxt:DXT5", "gpujpeg", "gpujpeg:q=90:lanes=1", "gpujpeg:q=80:lanes=4:restart=8"])
def test_lifecycle_through_reference_framework(cfg):
    """compress_init / poison pill / compress_pop / compress_done of the unmodified framework around our module state: host logic only"""
    fw = framework()
    st = fw.fwd_init(cfg.encode())
    assert st
    fw.fwd_frame(st, None, 0, 0, 0, 0, 0.0)
    assert pop(fw, st, 16) is None
    fw.fwd_done(st)


@pytest.mark.parametrize("cfg", ["cuda_dxt:DXT3", "gpujpeg:bogus=1", "gpujpeg:lanes=0", "no_such_module"])
def test_bad_configuration_is_refused_by_compress_init(cfg):
    fw = framework()
    assert not fw.fwd_init(cfg.encode())


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,inc,dxt_type", [("cuda_dxt", UYVY, 1), ("cuda_dxt:DXT5", UYVY, 6), ("cuda_dxt:DXT1", RGB, 1), ("cuda_dxt:DXT1", RGBA, 1)])
def test_cuda_dxt_through_reference_framework(orc, cfg, inc, dxt_type):
    """a real struct video_frame in, a frame of the reference's video_frame_pool out: the bytes of the kernels run by hand"""
    import torch
    from ultragrid_b200 import api, compress
    fw = framework()
    w, h = 1920, 1080
    src = util.rng_bytes(orc.orc_vc_get_linesize(w, inc) * h, 40 + inc)
    d = torch.from_numpy(src).cuda()
    mid_codec = inc if inc in (UYVY, RGB) else compress.get_best_decoder_from(inc, [RGB, UYVY])
    mid = d if mid_codec == inc else api.pixfmt_convert(inc, mid_codec, d, w, h)
    want = (api.uyvy_to_dxt(mid, w, h, dxt_type=dxt_type) if mid_codec == UYVY else
            api.compat_to_dxt("cuda_rgb_to_dxt1" if dxt_type == 1 else "cuda_rgb_to_dxt6", mid, w, h)).cpu().numpy()
    st = fw.fwd_init(cfg.encode())
    assert st
    for rep in range(3):  # pool reuse
        fw.fwd_frame(st, src.ctypes.data, 0, w, h, inc, 30.0)
        got, codec, seq, ow, oh = pop(fw, st, w * h)
        assert codec == (DXT1 if dxt_type == 1 else DXT5) and (ow, oh) == (w, h)
        assert np.array_equal(got, want)
    fw.fwd_frame(st, ctypes.c_void_p(d.data_ptr()), 1, w, h, inc, 30.0)  # mem_location == CUDA_MEM
    got, *_ = pop(fw, st, w * h)
    assert np.array_equal(got, want)
    fw.fwd_frame(st, None, 0, 0, 0, 0, 0.0)
    assert pop(fw, st, 16) is None
    fw.fwd_done(st)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["gpujpeg:q=90:lanes=1", "gpujpeg:q=90"])
def test_gpujpeg_through_reference_framework(orc, cfg):
    from test_jpeg import natural_rgb, orc_encode
    fw = framework()
    w, h, n = 640, 360, 6
    frames = [util.convert_cpu(orc, "orc_convert", RGB, UYVY, natural_rgb(w, h, 200 + i).reshape(-1), w, h) for i in range(n)]
    want = [orc_encode(orc, f, w, h, UYVY, 90) for f in frames]
    st = fw.fwd_init(cfg.encode())
    assert st
    for f in frames:
        fw.fwd_frame(st, f.ctypes.data, 0, w, h, UYVY, 60.0)
    fw.fwd_frame(st, None, 0, 0, 0, 0, 0.0)
    for i in range(n):
        got, codec, seq, ow, oh = pop(fw, st, w * h * 3 + 4096)
        assert codec == JPEG and seq == i and (ow, oh) == (w, h)
        assert got.tobytes() == want[i], i
    assert pop(fw, st, 16) is None
    fw.fwd_done(st)


# ---- decompress side: ultragrid_vdecompress_{gpujpeg,gpujpeg_to_dxt,dxt_cuda}.so in the unmodified src/video_decompress.c -------------------------
def dec_frame(fw, st, data, out_bytes, seq=0):
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.zeros(max(out_bytes, 1), np.uint8)
    props = (ctypes.c_int * 3)()
    rc = fw.fwd_dec_frame(st, dst.ctypes.data, src.ctypes.data, len(src), seq, props)
    return rc, dst, list(props)


def test_decompress_modules_register_with_the_reference_registry():
    fw = framework()
    for name in (b"gpujpeg", b"gpujpeg_to_dxt", b"dxt_cuda"):
        assert fw.fwd_has_decompress_module(name), name
    assert not fw.fwd_has_decompress_module(b"libavcodec")
    assert not fw.fwd_dec_init(UYVY, RGB)  # not a compression: every module answers VDEC_PRIO_NA, decompress_init_multi fails


@pytest.mark.gpu
def test_gpujpeg_decompress_through_reference_framework(orc):
    """decompress_init_multi of the unmodified video_decompress.c picks our module by its priority; probe + decode == the decoder called by hand"""
    from test_jpeg import natural_rgb, orc_encode
    from ultragrid_b200 import api
    fw = framework()
    w, h = 640, 360
    rgb = natural_rgb(w, h, 31)
    uyvy = util.convert_cpu(orc, "orc_convert", RGB, UYVY, rgb.reshape(-1), w, h)
    s_yuv, s_rgb = orc_encode(orc, uyvy, w, h, UYVY, 90), orc_encode(orc, rgb.reshape(-1).copy(), w, h, RGB, 90)
    probe = fw.fwd_dec_init(JPEG, 0)  # out_codec VIDEO_CODEC_NONE: VDEC_PRIO_PROBE_HI
    assert probe and fw.fwd_dec_accepts_corrupted(probe) == 0
    assert fw.fwd_dec_reconfigure(probe, w, h, JPEG, 0, 8, 16, 0, 0)
    rc, _, props = dec_frame(fw, probe, s_yuv, 0)
    assert rc == 2 and props == [8, 4220, 0]  # DECODER_GOT_CODEC
    rc, _, props = dec_frame(fw, probe, s_rgb, 0)
    assert rc == 2 and props == [8, 4440, 1]
    fw.fwd_dec_done(probe)
    dec = api.JpegDecoder()
    for out_c, bpp, shifts in ((UYVY, 2, (0, 8, 16)), (RGB, 3, (0, 8, 16)), (RGBA, 4, (16, 8, 0))):
        st = fw.fwd_dec_init(JPEG, out_c)
        assert st
        pitch = w * bpp
        assert fw.fwd_dec_reconfigure(st, w, h, JPEG, *shifts, pitch, out_c)
        for stream in (s_yuv, s_rgb):
            rc, out, _ = dec_frame(fw, st, stream, pitch * h)
            assert rc == 1  # DECODER_GOT_FRAME
            assert np.array_equal(out, dec.decode(stream, out_c, shifts=shifts))
        assert dec_frame(fw, st, b"not a jpeg at all", pitch * h)[0] == 0  # DECODER_NO_FRAME
        assert not fw.fwd_dec_reconfigure(st, w, h, JPEG, *shifts, pitch, DXT1)  # not an output of this module
        fw.fwd_dec_done(st)


@pytest.mark.gpu
@pytest.mark.parametrize("out_c", [DXT1, DXT5])
def test_gpujpeg_to_dxt_through_reference_framework(orc, out_c):
    from test_jpeg import natural_rgb, orc_encode
    from ultragrid_b200 import api
    fw = framework()
    w, h = 256, 128
    stream = orc_encode(orc, natural_rgb(w, h, 8).reshape(-1).copy(), w, h, RGB, 90)
    st = fw.fwd_dec_init(JPEG, out_c)  # only gpujpeg_to_dxt answers for JPEG -> DXT
    assert st
    pitch = w // (2 if out_c == DXT1 else 1)  # vc_get_linesize(width, DXT1 / DXT5)
    assert fw.fwd_dec_reconfigure(st, w, h, JPEG, 0, 8, 16, pitch, out_c)
    n = w * h // (2 if out_c == DXT1 else 1)
    rc, out, _ = dec_frame(fw, st, stream, n)
    assert rc == 1
    rgb = api.JpegDecoder().decode(stream, RGB, device=True)
    want = api.compat_to_dxt("cuda_rgb_to_dxt1" if out_c == DXT1 else "cuda_rgb_to_dxt6", rgb, w, -h).cpu().numpy()
    assert np.array_equal(out[:want.size], want)
    fw.fwd_dec_done(st)


@pytest.mark.gpu
@pytest.mark.parametrize("comp", [DXT1, DXT5])
def test_dxt_cuda_through_reference_framework(orc, comp):
    import torch
    from ultragrid_b200 import api
    fw = framework()
    w, h = 256, 128
    blocks = util.rng_bytes(w * h // (2 if comp == DXT1 else 1), 4)
    rgb = api.dxt_to_rgb(torch.from_numpy(blocks).cuda(), w, h, 1 if comp == DXT1 else 6).cpu().numpy()
    for out_c, bpp, shifts in ((RGB, 3, (0, 8, 16)), (RGBA, 4, (8, 16, 24)), (UYVY, 2, (0, 8, 16))):
        st = fw.fwd_dec_init(comp, out_c)
        assert st
        assert fw.fwd_dec_reconfigure(st, w, h, comp, *shifts, w * bpp, out_c)
        rc, out, _ = dec_frame(fw, st, blocks.tobytes(), w * h * bpp)
        assert rc == 1
        want = rgb if out_c == RGB else util.convert_cpu(orc, "orc_convert", RGB, out_c, rgb, w, h, shifts=shifts)
        assert np.array_equal(out, want)
        fw.fwd_dec_done(st)
