"""codec_t surface of UltraGrid (src/types.h:62-112) and the per-codec geometry of
src/video_codec.c:120-206,507-560 — host-side bookkeeping only (buffer sizes, pitches)."""
import enum


class Codec(enum.IntEnum):
    NONE = 0
    RGBA = 1
    UYVY = 2
    YUYV = 3
    VUYA = 4
    R10k = 5
    R12L = 6
    v210 = 7
    DVS10 = 8
    DXT1 = 9
    DXT1_YUV = 10
    DXT5 = 11
    RGB = 12
    JPEG = 13
    BGR = 20
    RG48 = 27
    I420 = 29
    Y216 = 30
    Y416 = 31


# (block_size_bytes, block_size_pixels, h_align) — codec_info[], video_codec.c:120-206
_INFO = {
    Codec.RGBA: (4, 1, 1), Codec.UYVY: (4, 2, 2), Codec.YUYV: (4, 2, 2), Codec.VUYA: (4, 1, 1),
    Codec.R10k: (4, 1, 64), Codec.R12L: (36, 8, 8), Codec.v210: (16, 6, 48), Codec.DVS10: (16, 6, 48),
    Codec.DXT1: (1, 2, 0), Codec.DXT1_YUV: (1, 2, 0), Codec.DXT5: (1, 1, 0), Codec.RGB: (3, 1, 1),
    Codec.JPEG: (1, 1, 0), Codec.BGR: (3, 1, 1), Codec.RG48: (6, 1, 1), Codec.I420: (3, 2, 2),
    Codec.Y216: (8, 2, 2), Codec.Y416: (8, 1, 1),
}


def vc_get_linesize(width: int, codec: Codec) -> int:
    """vc_get_linesize(), video_codec.c:507-521 (line padded to the codec's h_align)."""
    bsz, bpx, h_align = _INFO[Codec(codec)]
    if h_align:
        width = (width + h_align - 1) // h_align * h_align
    return (width + bpx - 1) // bpx * bsz


def vc_get_size(width: int, codec: Codec) -> int:
    """vc_get_size(), video_codec.c:530-538 (no line padding)."""
    bsz, bpx, _ = _INFO[Codec(codec)]
    return (width + bpx - 1) // bpx * bsz


def vc_get_datalen(width: int, height: int, codec: Codec) -> int:
    """vc_get_datalen(), video_codec.c:543-560 for packed codecs; I420 is the only planar codec_t."""
    if Codec(codec) == Codec.I420:
        return width * height + 2 * ((width + 1) // 2) * ((height + 1) // 2)
    return vc_get_linesize(width, codec) * height
