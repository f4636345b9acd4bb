"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares (no compute)."""
import ctypes
import glob
import os
import re

import pytest

import util

ROOT = util.ROOT


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        if os.path.basename(h) == "cuda_pix_conv.h":
            continue  # C++ linkage like the reference's header: its (mangled) symbols are checked in test_cuda_pix_conv.py
        text = open(h).read()
        names += re.findall(r"UGB_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)
    return names


def test_headers_declare_reference_abi():
    names = declared_symbols()
    # cuda_dxt/cuda_dxt.h:30-89 and src/cuda_wrapper.h:61-73 of the reference
    for n in ["cuda_rgb_to_dxt1", "cuda_yuv_to_dxt1", "cuda_rgb_to_dxt6", "cuda_yuv_to_dxt6", "cuda_yuv422_to_yuv444",
              "cuda_wrapper_free", "cuda_wrapper_free_host", "cuda_wrapper_host_alloc", "cuda_wrapper_malloc",
              "cuda_wrapper_malloc_host", "cuda_wrapper_memcpy", "cuda_wrapper_last_error_string",
              "cuda_wrapper_set_device", "cuda_wrapper_get_last_error", "cuda_wrapper_get_error_string",
              "cuda_wrapper_print_devices_info", "cuda_wrapper_device_reset"]:
        assert n in names


def test_library_exports_every_declared_symbol():
    from ultragrid_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_headers():
    from ultragrid_b200 import _lib
    assert sorted(_lib.SIGNATURES) == sorted(set(declared_symbols()))
    _lib.load()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ultragrid_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_codec_geometry_matches_oracle(orc):
    from ultragrid_b200 import Codec, vc_get_linesize, vc_get_size
    for c in (Codec.RGBA, Codec.UYVY, Codec.YUYV, Codec.R10k, Codec.R12L, Codec.v210, Codec.RGB, Codec.BGR, Codec.RG48, Codec.Y216,
              Codec.Y416):
        for w in (1, 2, 5, 6, 47, 48, 49, 64, 65, 1920, 3840, 7680):
            assert vc_get_linesize(w, c) == orc.orc_vc_get_linesize(w, int(c)), (c, w)
            assert vc_get_size(w, c) == orc.orc_vc_get_size(w, int(c)), (c, w)
