"""Baseline JPEG decode (the stage src/video_decompress/gpujpeg.c:268-330 delegates to libgpujpeg).
GPUJPEG is absent (parity unpinned against it); pinned instead:
  * CPU: oracle/jpeg_decode_oracle.c against libjpeg (PIL) on streams of our encoder and on libjpeg-made streams: every sample
    within 1 (libjpeg's default IDCT is integer, ours float AAN), and the reference's flat-grey round trip (test/gpujpeg_test.cpp:68-106);
  * GPU: ugb200_jpeg_decode == the oracle, byte for byte; other output codecs == the line converters applied to the native output."""
import ctypes
import io

import numpy as np
import pytest
from PIL import Image

import util
from test_jpeg import RGB, UYVY, natural_rgb, orc_encode, psnr

RGBA, VUYA, I420 = 1, 4, 29


def orc_decode(orc, stream, fmt, w, h):
    orc.orc_jpeg_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
    info = (ctypes.c_int * 6)()
    pitch = (w + 1) // 2 * 4 if fmt == 0 else w * 3
    out = np.zeros(pitch * h + 64, np.uint8)
    b = np.frombuffer(stream, np.uint8)
    rc = orc.orc_jpeg_decode(b.ctypes.data, len(stream), fmt, out.ctypes.data, pitch, info)
    assert rc == 0, rc
    return list(info), out[:pitch * h]


def pil_stream(rgb, quality, subsampling, restart_blocks=0):
    b = io.BytesIO()
    kw = {"restart_marker_blocks": restart_blocks} if restart_blocks else {}
    Image.fromarray(rgb).save(b, "JPEG", quality=quality, subsampling=subsampling, **kw)
    return b.getvalue()


def pil_ycc(stream, w, h):
    im = Image.open(io.BytesIO(stream))
    im.draft("YCbCr", (w, h))
    return np.asarray(im)


STREAMS = [("ours-uyvy", 200, 120, 90, 0), ("ours-uyvy", 1920, 1080, 75, 8), ("ours-uyvy", 98, 50, 90, 1), ("ours-rgb", 200, 120, 90, 0), ("ours-rgb", 130, 37, 75, 32),
           ("pil-444", 200, 120, 85, 0), ("pil-444", 64, 64, 95, 3), ("pil-422", 200, 120, 85, 5), ("pil-420", 200, 120, 85, 4), ("pil-420", 150, 75, 60, 0)]


def make_stream(orc, kind, w, h, q, ri):
    rgb = natural_rgb(w, h, 5)
    if kind == "ours-uyvy":
        return orc_encode(orc, util.convert_cpu(orc, "orc_convert", RGB, UYVY, rgb.reshape(-1), w, h), w, h, UYVY, q, ri), rgb
    if kind == "ours-rgb":
        return orc_encode(orc, rgb.reshape(-1).copy(), w, h, RGB, q, ri), rgb
    return pil_stream(rgb, q, {"pil-444": 0, "pil-422": 1, "pil-420": 2}[kind], ri), rgb


@pytest.mark.parametrize("kind,w,h,q,ri", STREAMS)
def test_oracle_decoder_matches_libjpeg(orc, kind, w, h, q, ri):
    s, rgb = make_stream(orc, kind, w, h, q, ri)
    if kind == "ours-rgb":  # Adobe transform 0: libjpeg hands back the stored RGB
        info, out = orc_decode(orc, s, 1, w, h)
        assert info[:3] == [w, h, 3] and info[5] == 0
        d = np.abs(out.reshape(h, w, 3).astype(int) - np.asarray(Image.open(io.BytesIO(s))).astype(int))
        assert d.max() <= 1 and d.mean() < 0.05
        return
    ycc = pil_ycc(s, w, h)
    if kind in ("ours-uyvy", "pil-422"):
        info, out = orc_decode(orc, s, 0, w, h)
        assert info[3:5] == [2, 1]
        uy = out.reshape(h, -1)
        d = np.abs(uy[:, 1::2][:, :w].astype(int) - ycc[:, :, 0].astype(int))  # luma (chroma: libjpeg upsamples with a triangle filter)
        assert d.max() <= 1 and d.mean() < 0.05
        cb = uy[:, 0::4].astype(int)
        assert np.abs(cb - ycc[:, 0::2, 1][:, :cb.shape[1]].astype(int)).mean() < 2.5
    else:
        info, out = orc_decode(orc, s, 1, w, h)
        o = out.reshape(h, w, 3)
        if kind == "pil-444":
            d = np.abs(o.astype(int) - ycc.astype(int))
        else:
            d = np.abs(o[:, :, 0].astype(int) - ycc[:, :, 0].astype(int))
        assert d.max() <= 1 and d.mean() < 0.05


def test_oracle_flat_grey_roundtrip(orc):
    """test/gpujpeg_test.cpp:68-106 of the reference: encode a flat grey RGB 1080p frame, decode, max |diff| <= 1"""
    w, h = 1920, 1080
    src = np.full(w * h * 3, 127, np.uint8)
    _, out = orc_decode(orc, orc_encode(orc, src, w, h, RGB, 75), 1, w, h)
    assert np.abs(out.astype(int) - 127).max() <= 1


def test_image_info_host_only(orc):
    from ultragrid_b200 import _lib
    lib = _lib.load()

    class Info(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("width", "height", "components", "h_samp", "v_samp", "adobe", "ri", "native")]
    for kind, w, h, q, ri in STREAMS[:1] + STREAMS[3:4] + STREAMS[5:6] + STREAMS[8:9]:
        s, _ = make_stream(orc, kind, w, h, q, ri)
        info = Info()
        assert lib.ugb200_jpeg_get_image_info((ctypes.c_uint8 * len(s)).from_buffer_copy(s), len(s), ctypes.byref(info)) == 0
        assert (info.width, info.height, info.components) == (w, h, 3)
        assert info.native == {"ours-uyvy": UYVY, "ours-rgb": RGB, "pil-444": VUYA, "pil-420": UYVY}[kind]
    assert lib.ugb200_jpeg_get_image_info((ctypes.c_uint8 * 4)(1, 2, 3, 4), 4, ctypes.byref(Info())) != 0


@pytest.mark.parametrize("kind,w,h,q,ri", STREAMS + [("ours-uyvy", 3840, 2160, 90, 0), ("ours-uyvy", 7680, 4320, 90, 0)])  # the last two: 4 / 8 scan threads
def test_stream_parser_finds_restart_segments(orc, kind, w, h, q, ri):
    """host logic of the decoder: the SSE2 marker scan against a plain numpy search of the same stream"""
    from ultragrid_b200 import _lib
    lib = _lib.load()
    s, _ = make_stream(orc, kind, w, h, q, ri)
    a = np.frombuffer(s, np.uint8)
    ff = np.flatnonzero(a[:-1] == 0xFF)
    nxt = a[ff + 1]
    sos = ff[nxt == 0xDA]
    cap = 1 << 17
    begin, end = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    n = lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(s), begin.ctypes.data, end.ctypes.data, cap)
    assert n > 0
    want_b, want_e = [], []
    for k, so in enumerate(sos):
        data0 = so + 2 + (int(a[so + 2]) << 8 | int(a[so + 3]))
        stop = next(p for p, c in zip(ff, nxt) if p >= data0 and c != 0 and not 0xD0 <= c <= 0xD7 and c != 0xFF)
        rst = [p for p, c in zip(ff, nxt) if data0 <= p < stop and 0xD0 <= c <= 0xD7]
        want_b += [data0] + [p + 2 for p in rst]
        want_e += rst + [stop]
    assert n == len(want_b)
    assert begin[:n].tolist() == want_b and end[:n].tolist() == want_e


@pytest.mark.gpu
@pytest.mark.parametrize("kind,w,h,q,ri", STREAMS + [("ours-uyvy", 3840, 2160, 90, 0), ("ours-rgb", 1920, 1080, 90, 0)])
def test_gpu_decoder_equals_oracle(orc, kind, w, h, q, ri):
    from ultragrid_b200 import api
    s, _ = make_stream(orc, kind, w, h, q, ri)
    dec = api.JpegDecoder()
    native = api.jpeg_image_info(s).native_codec
    if native == UYVY:
        _, want = orc_decode(orc, s, 0, w, h)
    elif native == RGB:
        _, want = orc_decode(orc, s, 1, w, h)
    else:  # 4:4:4 YCbCr: the oracle's packed Y Cb Cr against VUYA
        _, ycc = orc_decode(orc, s, 1, w, h)
        ycc = ycc.reshape(h, w, 3)
        want = np.stack([ycc[:, :, 2], ycc[:, :, 1], ycc[:, :, 0], np.full((h, w), 255, np.uint8)], axis=2).reshape(-1)
    got = dec.decode(s, native)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:8]
    assert np.array_equal(dec.decode(s, native, device=True).cpu().numpy(), want)
    if native == UYVY:  # I420 output (GPUJPEG_420_U8_P0P1P2, gpujpeg.c:113-116) = uyvy_to_i420 of the native frame (to_planar.c:343-378)
        uy = want.reshape(h, -1).astype(np.int32)
        rows_b = [min(y + 1, h - 1) for y in range(0, h, 2)]
        cw = (w + 1) // 2
        u = (uy[0::2, 0::4][:, :cw] + uy[rows_b, 0::4][:, :cw] + 1) // 2
        v = (uy[0::2, 2::4][:, :cw] + uy[rows_b, 2::4][:, :cw] + 1) // 2
        i420 = np.concatenate([uy[:, 1::2][:, :w].reshape(-1), u.reshape(-1), v.reshape(-1)]).astype(np.uint8)
        assert np.array_equal(dec.decode(s, I420), i420)
        assert np.array_equal(dec.decode(s, I420, device=True).cpu().numpy(), i420)
    # another output codec = UltraGrid's line converter applied to the native frame
    for out_c in (UYVY, RGB, RGBA):
        if out_c == native or not api.pixfmt_supported(native, out_c):
            continue
        shifts = (16, 8, 0) if out_c == RGBA else (0, 8, 16)
        conv = util.convert_cpu(orc, "orc_convert", native, out_c, want, w, h, shifts=shifts)
        assert np.array_equal(dec.decode(s, out_c, shifts=shifts), conv), (native, out_c)
    dec.close()


@pytest.mark.gpu
def test_gpu_roundtrip_8k(orc):
    """encode -> decode on the device at BASELINE's size: PSNR of the round trip as the quantiser implies, decoder == oracle on a band"""
    import torch
    from ultragrid_b200 import api
    w, h = 7680, 4320
    src = util.convert_cpu(orc, "orc_convert", RGB, UYVY, natural_rgb(w, h, 3).reshape(-1), w, h)
    enc, dec = api.JpegEncoder(), api.JpegDecoder()
    enc.encode_device(torch.from_numpy(src).cuda(), w, h, UYVY, quality=90)
    s = enc.result()
    out = dec.decode(s, UYVY)
    assert psnr(out, src) > 38.0
    _, want = orc_decode(orc, s, 0, w, h)
    assert np.array_equal(out, want)


def test_stream_parser_survives_corruption(orc):
    """host logic: truncated and randomly damaged streams are rejected or parsed, never read out of bounds (segments stay inside the stream)"""
    from ultragrid_b200 import _lib
    lib = _lib.load()
    s, _ = make_stream(orc, "ours-uyvy", 200, 120, 90, 0)
    rng = np.random.default_rng(5)
    cap = 1 << 12
    begin, end = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)

    class Info(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("width", "height", "components", "h_samp", "v_samp", "adobe", "ri", "native")]
    for trial in range(400):
        a = np.frombuffer(s, np.uint8).copy()
        kind = trial % 4
        if kind == 0:
            a = a[:rng.integers(2, len(a))]                       # truncated anywhere
        elif kind == 1:
            pos = rng.integers(0, len(a), rng.integers(1, 8))     # a few damaged bytes, headers included
            a[pos] = rng.integers(0, 256, len(pos))
        elif kind == 2:
            pos = rng.integers(0, min(len(a), 700))               # a damaged header length / marker
            a[pos] = 0xFF
        else:
            a = np.concatenate([a[:rng.integers(2, 600)], rng.integers(0, 256, rng.integers(1, 300)).astype(np.uint8)])
        a = np.ascontiguousarray(a)
        rc = lib.ugb200_jpeg_get_image_info(a.ctypes.data, len(a), ctypes.byref(Info()))
        n = lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(a), begin.ctypes.data, end.ctypes.data, cap)
        if n > 0:
            k = min(n, cap)
            assert (begin[:k] <= end[:k]).all() and (end[:k] <= len(a)).all(), (trial, kind)
        assert rc <= 0


@pytest.mark.gpu
def test_gpu_decoder_survives_corruption(orc):
    """damaged entropy-coded data decodes to *something* of the right size (or is rejected) without faulting the device"""
    import torch
    from ultragrid_b200 import api
    s, _ = make_stream(orc, "ours-uyvy", 200, 120, 90, 0)
    rng = np.random.default_rng(6)
    dec = api.JpegDecoder()
    good = dec.decode(s, UYVY)
    for trial in range(40):
        a = np.frombuffer(s, np.uint8).copy()
        if trial % 2:
            a = a[:rng.integers(700, len(a))]
        else:
            pos = rng.integers(640, len(a), rng.integers(1, 30))
            a[pos] = rng.integers(0, 256, len(pos))
        try:
            out = dec.decode(a.tobytes(), UYVY)
            assert out.size == good.size
        except RuntimeError:
            pass
    torch.cuda.synchronize()
    assert np.array_equal(dec.decode(s, UYVY), good)  # the decoder is still healthy


def test_oversubscribed_huffman_table_is_rejected(orc):
    """untrusted DHT: 255 codes of length 1 pass the 'sum <= 256' check but violate Kraft; the look-up fill would then write far past the
    table (libjpeg: JERR_BAD_HUFF_TABLE).  Host-side parser only."""
    from ultragrid_b200 import _lib
    lib = _lib.load()
    s, _ = make_stream(orc, "ours-uyvy", 64, 32, 90, 0)
    a = np.frombuffer(s, np.uint8).copy()
    ff = np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xC4))
    assert len(ff) >= 1
    begin, end = np.zeros(64, np.uint32), np.zeros(64, np.uint32)
    assert lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(a), begin.ctypes.data, end.ctypes.data, 64) > 0
    for bits in ([255] + [0] * 15, [2, 1] + [0] * 14, [0] * 7 + [250, 7] + [0] * 7):
        b = a.copy()
        p = int(ff[0]) + 5  # marker(2) length(2) Tc/Th(1)
        n_old = int(b[p:p + 16].sum())
        n_new = sum(bits)
        body = np.concatenate([np.array(bits, np.uint8), np.arange(n_new, dtype=np.uint8)])
        seglen = 2 + 1 + 16 + n_new
        b = np.concatenate([b[:int(ff[0]) + 2], np.array([seglen >> 8, seglen & 255, b[int(ff[0]) + 4]], np.uint8), body, b[p + 16 + n_old:]])
        b = np.ascontiguousarray(b)
        assert lib.ugb200_jpeg_debug_segments(b.ctypes.data, len(b), begin.ctypes.data, end.ctypes.data, 64) < 0, bits


def test_absurd_dimensions_are_refused_before_allocation(orc):
    """DRI = 1 with 65535 x 65535 in SOF0 would reserve tens of millions of segment entries: refused by the pixel bound"""
    from ultragrid_b200 import _lib
    lib = _lib.load()
    s, _ = make_stream(orc, "ours-uyvy", 64, 32, 90, 0)
    a = np.frombuffer(s, np.uint8).copy()
    sof = int(np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xC0))[0])
    a[sof + 5:sof + 9] = 0xFF
    begin, end = np.zeros(64, np.uint32), np.zeros(64, np.uint32)
    assert lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(a), begin.ctypes.data, end.ctypes.data, 64) < 0


@pytest.mark.gpu
def test_decoder_refuses_stream_of_another_size(orc):
    """the destination is sized from reconfigure()'s video_desc, the stream's SOF0 says how much is written: a mismatch must be refused
    (module: DECODER_NO_FRAME) instead of overflowing the buffer"""
    from ultragrid_b200 import compress
    small, _ = make_stream(orc, "ours-uyvy", 64, 32, 90, 0)
    big, _ = make_stream(orc, "ours-uyvy", 128, 64, 90, 0)
    d = compress.Decompress(13, UYVY)  # JPEG
    d.reconfigure(64, 32, 13, UYVY)
    st, out, _ = d.frame(small)
    assert st == d.GOT_FRAME and out is not None
    st, out, _ = d.frame(big)
    assert st != d.GOT_FRAME and out is None
    st, out, _ = d.frame(small)
    assert st == d.GOT_FRAME
    d.close()


def _segments(lib, dec, cap=1 << 17):
    begin, end = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    n = lib.ugb200_jpeg_decoder_last_segments(dec._h, begin.ctypes.data, end.ctypes.data, cap)
    return n, begin[:max(n, 0)].tolist(), end[:max(n, 0)].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,w,h,q,ri,damage", [
    ("ours-uyvy", 1920, 1080, 90, 0, None), ("ours-uyvy", 200, 120, 90, 0, None), ("ours-uyvy", 322, 50, 75, 3, None), ("pil-422", 640, 480, 85, 5, None),
    ("pil-444", 333, 211, 90, 2, None), ("pil-420", 640, 480, 85, 0, None),                         # the last one: no DRI, one segment
    ("ours-uyvy", 1920, 1080, 90, 0, "truncate"), ("ours-uyvy", 1920, 1080, 90, 0, "drop-rst"), ("ours-uyvy", 1920, 1080, 90, 0, "extra-rst"),
    ("ours-uyvy", 1920, 1080, 90, 0, "foreign-marker"), ("ours-uyvy", 7680, 4320, 90, 0, None)])
def test_gpu_device_marker_scan_equals_host_scan(orc, monkeypatch, kind, w, h, q, ri, damage):
    """single-scan streams: the restart segments found by the device kernels are the host parser's, intact or damaged; the decoded frame is the same.
    (8K: the automatic choice - device scan for streams of 1 MB and more; the others force it.)"""
    from ultragrid_b200 import _lib, api
    lib = _lib.load()
    s, _ = make_stream(orc, kind, w, h, q, ri)
    a = np.frombuffer(s, np.uint8).copy()
    ff = np.flatnonzero((a[:-1] == 0xFF) & (a[1:] >= 0xD0) & (a[1:] <= 0xD7))
    if damage == "truncate":
        a = a[:len(a) * 6 // 10]
    elif damage == "drop-rst":      # two restart markers turned into entropy-coded bytes: their segments merge, the table's tail is empty
        a[ff[5]:ff[5] + 2] = (0x12, 0x34)
        a[ff[40]:ff[40] + 2] = (0x12, 0x34)
    elif damage == "extra-rst":     # more restart markers than the frame has segments
        a = np.concatenate([a[:-2], np.tile(np.array([0x55, 0xFF, 0xD3], np.uint8), 50), a[-2:]])
    elif damage == "foreign-marker":  # a non-RSTn marker in the middle of the scan ends it (an early EOI; the host parser goes on reading
        a[ff[len(ff) // 2] + 1] = 0xD9  # segments behind any other marker and may reject the stream, the device path never looks behind the scan)
    a = np.ascontiguousarray(a)
    data = a.tobytes()
    monkeypatch.setenv("UGB200_JPEG_MARKER_SCAN", "host")
    host = api.JpegDecoder()
    if w < 7680:
        monkeypatch.setenv("UGB200_JPEG_MARKER_SCAN", "device")
    else:
        monkeypatch.delenv("UGB200_JPEG_MARKER_SCAN")
    dev = api.JpegDecoder()
    monkeypatch.delenv("UGB200_JPEG_MARKER_SCAN", raising=False)
    codec = UYVY if kind in ("ours-uyvy", "pil-422", "pil-420") else VUYA
    for _ in range(2):  # second turn: the decoders' scratch of the first one is reused
        want = host.decode(data, codec, device=True)
        n_h, b_h, e_h = _segments(lib, host)
        got = dev.decode(data, codec, device=True)
        n_d, b_d, e_d = _segments(lib, dev)
        assert n_h == n_d > 0
        assert b_h == b_d and e_h == e_d
        assert bool((want == got).all())
    cap = 1 << 17
    begin, end = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    assert lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(a), begin.ctypes.data, end.ctypes.data, cap) == n_h
    assert begin[:n_h].tolist() == b_h and end[:n_h].tolist() == e_h
    host.close(), dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,q,ri,damage", [
    (640, 360, 90, 0, None), (200, 120, 75, 3, None), (1920, 1080, 90, 0, None), (7680, 4320, 90, 0, None),
    (1920, 1080, 90, 0, "truncate"), (1920, 1080, 90, 0, "truncate-scan1"), (1920, 1080, 90, 0, "drop-rst"), (1920, 1080, 90, 0, "extra-rst"),
    (1920, 1080, 90, 0, "foreign-marker"), (1920, 1080, 90, 0, "swap-tables"), (1920, 1080, 90, 0, "other-component"), (1920, 1080, 90, 0, "dht-between")])
def test_gpu_device_marker_scan_of_multi_scan_streams(orc, monkeypatch, w, h, q, ri, damage):
    """three scans (RGB as GPUJPEG stores it, one per component): the device finds the later SOS headers behind the entropy-coded data, checks that they
    are what the first one promises and builds the segment table of all three scans; a stream that breaks the promise (damaged, tables redefined between the
    scans, another component order) is handed back to the host parser - either way segments and pixels are the host path's."""
    from ultragrid_b200 import _lib, api
    lib = _lib.load()
    s, _ = make_stream(orc, "ours-rgb", w, h, q, ri)
    a = np.frombuffer(s, np.uint8).copy()
    rst = np.flatnonzero((a[:-1] == 0xFF) & (a[1:] >= 0xD0) & (a[1:] <= 0xD7))
    sos = np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xDA))
    assert len(sos) == 3
    if damage == "truncate":            # inside the last scan
        a = a[:(sos[2] + len(a)) // 2]
    elif damage == "truncate-scan1":    # the later scans are missing altogether
        a = a[:(sos[0] + sos[1]) // 2]
    elif damage == "drop-rst":
        a[rst[5]:rst[5] + 2] = (0x12, 0x34)
        k = int(np.searchsorted(rst, sos[1])) + 7
        a[rst[k]:rst[k] + 2] = (0x12, 0x34)
    elif damage == "extra-rst":
        a = np.concatenate([a[:sos[1]], np.tile(np.array([0x55, 0xFF, 0xD3], np.uint8), 50), a[sos[1]:]])
    elif damage == "foreign-marker":    # an EOI in the middle of the second scan
        k = int(np.searchsorted(rst, (sos[1] + sos[2]) // 2))
        a[rst[k] + 1] = 0xD9
    elif damage == "swap-tables":       # legal: the second scan codes with the luminance tables selectors (the data was coded with the others: garbage, equally on both paths)
        a[sos[1] + 6] = 0x00
    elif damage == "other-component":   # legal: the second scan carries the third component and the third scan the second
        a[sos[1] + 5], a[sos[2] + 5] = a[sos[2] + 5], a[sos[1] + 5]
    elif damage == "dht-between":       # legal: a (repeated) DHT segment between the scans
        dht = int(np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xC4))[0])
        L = int(a[dht + 2]) << 8 | int(a[dht + 3])
        a = np.concatenate([a[:sos[1]], a[dht:dht + 2 + L], a[sos[1]:]])
    a = np.ascontiguousarray(a)
    data = a.tobytes()
    monkeypatch.setenv("UGB200_JPEG_MARKER_SCAN", "host")
    host = api.JpegDecoder()
    if w < 7680:
        monkeypatch.setenv("UGB200_JPEG_MARKER_SCAN", "device")
    else:
        monkeypatch.delenv("UGB200_JPEG_MARKER_SCAN")
    dev = api.JpegDecoder()
    monkeypatch.delenv("UGB200_JPEG_MARKER_SCAN", raising=False)
    for _ in range(2):
        try:
            want = host.decode(data, RGB, device=True)
        except RuntimeError:
            with pytest.raises(RuntimeError):  # what the host parser refuses, the device path refuses too (it hands the frame back)
                dev.decode(data, RGB, device=True)
            continue
        n_h, b_h, e_h = _segments(lib, host, 1 << 18)
        got = dev.decode(data, RGB, device=True)
        n_d, b_d, e_d = _segments(lib, dev, 1 << 18)
        assert n_h == n_d > 0
        assert b_h == b_d and e_h == e_d
        assert bool((want == got).all())
    host.close(), dev.close()
