"""Python driver of the C++ video_compress module layer (include/ugb200_vcompress.h) — the call a user of the
reference makes: compress_init("cuda_dxt:DXT1") / compress_frame / compress_pop (src/video_compress.h:95-107)."""
import ctypes

import numpy as np

from . import _lib

_L = _lib.load()


def set_cuda_devices(devices):
    """-D/--cuda-device (src/main.cpp:402-431)"""
    arr = (ctypes.c_int * len(devices))(*devices)
    if _L.ugb200_set_cuda_devices(arr, len(devices)) != 0:
        raise ValueError("bad device list")


def get_best_decoder_from(in_codec, candidates):
    arr = (ctypes.c_int * len(candidates))(*[int(c) for c in candidates])
    return _L.ugb200_get_best_decoder_from(int(in_codec), arr, len(candidates))


class Compress:
    def __init__(self, config):
        self._h = _L.ugb200_compress_init(config.encode())
        if not self._h:
            raise RuntimeError(f"compress_init({config!r}) failed")
        self._keep = []

    def push(self, frame, width, height, codec, fps=60.0):
        """frame: host numpy uint8 array, a CUDA torch tensor (CUDA_MEM), or None (poison pill)"""
        if frame is None:
            rc = _L.ugb200_compress_push(self._h, None, 0, 0, 0, 0, 0.0)
        elif isinstance(frame, np.ndarray):
            self._keep.append(frame)
            rc = _L.ugb200_compress_push(self._h, ctypes.c_void_p(frame.ctypes.data), 0, width, height, int(codec), fps)
        else:
            self._keep.append(frame)
            rc = _L.ugb200_compress_push(self._h, ctypes.c_void_p(frame.data_ptr()), 1, width, height, int(codec), fps)
        if rc != 0:
            raise RuntimeError(f"compress_frame failed ({rc})")

    def pop_into(self, out):
        """pops the next frame into the numpy buffer `out`; returns (nbytes, codec, seq) or None at end of stream"""
        n, codec, seq = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_uint()
        rc = _L.ugb200_compress_pop(self._h, ctypes.c_void_p(out.ctypes.data), out.size, ctypes.byref(n), ctypes.byref(codec), ctypes.byref(seq))
        if rc == 1:
            return None
        if rc != 0:
            raise RuntimeError(f"compress_pop failed ({rc}), frame len {n.value}")
        if self._keep:
            self._keep.pop(0)
        return n.value, codec.value, seq.value

    def pop_ref(self):
        """zero-copy pop: (numpy view into the pooled pinned frame — valid until the next pop, codec, seq) or None at end"""
        ptr, n, codec, seq = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int(), ctypes.c_uint()
        rc = _L.ugb200_compress_pop_ref(self._h, ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(codec), ctypes.byref(seq))
        if rc == 1:
            return None
        if rc != 0:
            raise RuntimeError(f"compress_pop failed ({rc})")
        if self._keep:
            self._keep.pop(0)
        view = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value,))
        return view, codec.value, seq.value

    def pop(self, capacity):
        """returns (bytes array, codec, seq) or None at end of stream"""
        out = np.empty(capacity, dtype=np.uint8)
        r = self.pop_into(out)
        return None if r is None else (out[:r[0]], r[1], r[2])

    def close(self):
        if self._h and _L is not None:
            _L.ugb200_compress_done(self._h)
        self._h = None

    def __del__(self):
        self.close()


class Decompress:
    """decompress_init_multi / decompress_reconfigure / decompress_frame (src/video_decompress.h:173-213) through the C driver.
    Host buffers in and out, like UltraGrid's receiver calls it."""

    GOT_FRAME, GOT_CODEC = 1, 2

    def __init__(self, compression, out_codec):
        self._h = _L.ugb200_decompress_init(int(compression), int(out_codec))
        if not self._h:
            raise RuntimeError(f"no decompress module for {compression} -> {out_codec}")
        self.module = _L.ugb200_decompress_module(self._h).decode()
        self._out_bytes = 0

    def reconfigure(self, width, height, compression, out_codec, shifts=(0, 8, 16), pitch=None):
        from .codec import vc_get_linesize
        pitch = vc_get_linesize(width, out_codec) if pitch is None and int(out_codec) != 0 else (pitch or 0)
        ok = _L.ugb200_decompress_reconfigure(self._h, width, height, int(compression), shifts[0], shifts[1], shifts[2], pitch, int(out_codec))
        if not ok:
            raise RuntimeError("decompress_reconfigure failed")
        self._out_bytes = pitch * height
        return pitch

    def frame(self, data, seq=0):
        """returns (status, output bytes as numpy array or None, [depth, subsampling, rgb])"""
        import numpy as np
        src = np.frombuffer(data, dtype=np.uint8)
        dst = np.zeros(max(self._out_bytes, 1), dtype=np.uint8)
        props = (ctypes.c_int * 3)()
        st = _L.ugb200_decompress_frame(self._h, ctypes.c_void_p(dst.ctypes.data), ctypes.c_void_p(src.ctypes.data), len(data), seq, props)
        return st, (dst if st == self.GOT_FRAME else None), list(props)

    def close(self):
        if self._h and _L is not None:
            _L.ugb200_decompress_done(self._h)
        self._h = None

    def __del__(self):
        self.close()
