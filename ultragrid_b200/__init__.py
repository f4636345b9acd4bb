"""B200-native implementation of UltraGrid's per-frame pixel-conversion / block-compression hot path.

The product is the C-ABI shared library ``libugb200.so`` (hand-written sm_100a CUDA behind UltraGrid's
``cuda_dxt.h`` / ``cuda_wrapper.h`` entry points plus fused/asynchronous additions, see ``include/``).
This Python package is only the binding used by the tests and ``bench.py``: torch supplies device
memory and streams, every operation goes through the C ABI.  There is no CPU fallback — importing
:mod:`ultragrid_b200.api` raises if the library is missing.
"""
from .codec import Codec, vc_get_linesize, vc_get_size, vc_get_datalen  # noqa: F401

__all__ = ["Codec", "vc_get_linesize", "vc_get_size", "vc_get_datalen"]
