"""Decompress module layer (host mirror of src/video_decompress.h; modules gpujpeg, gpujpeg_to_dxt, dxt_cuda)."""
import numpy as np
import pytest

import util
from test_jpeg import RGB, UYVY, natural_rgb, orc_encode

RGBA, DXT1, DXT5, JPEG, NONE = 1, 9, 11, 13, 0


def test_codec_ids_match_reference_enum():
    from ultragrid_b200 import Codec
    assert (int(Codec.DXT1), int(Codec.DXT5), int(Codec.JPEG), int(Codec.RGBA)) == (DXT1, DXT5, JPEG, RGBA)


def test_no_module_for_unknown_pair_is_reported():
    from ultragrid_b200.compress import Decompress
    with pytest.raises(RuntimeError):
        Decompress(UYVY, RGB)  # not a compression: every module answers -1


@pytest.mark.gpu
def test_gpujpeg_module_probe_and_decode(orc):
    from ultragrid_b200 import api
    from ultragrid_b200.compress import Decompress
    w, h = 320, 200
    rgb = natural_rgb(w, h, 2)
    uyvy = util.convert_cpu(orc, "orc_convert", RGB, UYVY, rgb.reshape(-1), w, h)
    s_yuv, s_rgb = orc_encode(orc, uyvy, w, h, UYVY, 90), orc_encode(orc, rgb.reshape(-1).copy(), w, h, RGB, 90)
    probe = Decompress(JPEG, NONE)
    assert probe.module == "gpujpeg"
    probe.reconfigure(w, h, JPEG, NONE)
    st, _, props = probe.frame(s_yuv)
    assert st == Decompress.GOT_CODEC and props == [8, 4220, 0]
    st, _, props = probe.frame(s_rgb)
    assert st == Decompress.GOT_CODEC and props == [8, 4440, 1]
    dec = api.JpegDecoder()
    for out_c, shifts in ((UYVY, (0, 8, 16)), (RGB, (0, 8, 16)), (RGBA, (16, 8, 0))):
        d = Decompress(JPEG, out_c)
        assert d.module == "gpujpeg"
        for stream in (s_yuv, s_rgb):
            pitch = d.reconfigure(w, h, JPEG, out_c, shifts=shifts)
            st, out, _ = d.frame(stream)
            assert st == Decompress.GOT_FRAME
            assert np.array_equal(out, dec.decode(stream, out_c, shifts=shifts))
        pitch = d.reconfigure(w, h, JPEG, out_c, shifts=shifts, pitch=pitch + 64)  # a padded destination
        st, out, _ = d.frame(s_yuv)
        want = dec.decode(s_yuv, out_c, shifts=shifts).reshape(h, -1)
        assert st == Decompress.GOT_FRAME and np.array_equal(out.reshape(h, pitch)[:, :want.shape[1]], want)
        d.close()
    assert d.frame is not None and Decompress(JPEG, RGB).frame(b"not a jpeg")[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("out_c", [DXT1, DXT5])
def test_gpujpeg_to_dxt_module(orc, out_c):
    """JPEG -> RGB on the device -> cuda_rgb_to_dxt{1,6} with mirrored height (src/video_decompress/gpujpeg_to_dxt.cpp:141-153)"""
    import torch
    from ultragrid_b200 import api
    from ultragrid_b200.compress import Decompress
    w, h = 256, 128
    stream = orc_encode(orc, natural_rgb(w, h, 8).reshape(-1).copy(), w, h, RGB, 90)
    d = Decompress(JPEG, out_c)
    assert d.module == "gpujpeg_to_dxt"
    d.reconfigure(w, h, JPEG, out_c)
    st, out, _ = d.frame(stream)
    assert st == Decompress.GOT_FRAME
    rgb = api.JpegDecoder().decode(stream, RGB, device=True)
    want = api.compat_to_dxt("cuda_rgb_to_dxt1" if out_c == DXT1 else "cuda_rgb_to_dxt6", rgb, w, -h).cpu().numpy()
    assert np.array_equal(out[:want.size], want)


@pytest.mark.gpu
@pytest.mark.parametrize("comp", [DXT1, DXT5])
def test_dxt_cuda_module(orc, comp):
    import torch
    from ultragrid_b200 import api
    from ultragrid_b200.compress import Decompress
    w, h = 256, 128
    blocks = util.rng_bytes(w * h // (2 if comp == DXT1 else 1), 4)
    rgb = api.dxt_to_rgb(torch.from_numpy(blocks).cuda(), w, h, 1 if comp == DXT1 else 6).cpu().numpy()
    for out_c, shifts in ((RGB, (0, 8, 16)), (RGBA, (8, 16, 24)), (UYVY, (0, 8, 16))):
        d = Decompress(comp, out_c)
        assert d.module == "dxt_cuda"
        d.reconfigure(w, h, comp, out_c, shifts=shifts)
        st, out, _ = d.frame(blocks.tobytes())
        assert st == Decompress.GOT_FRAME
        want = rgb if out_c == RGB else util.convert_cpu(orc, "orc_convert", RGB, out_c, rgb, w, h, shifts=shifts)
        assert np.array_equal(out, want)
        d.close()
