"""The decision logic of the JPEG decoder's device marker scan for multi-scan streams (ultragrid_b200/csrc/jpeg_marker_bounds.cuh: which candidates end a scan,
what the later SOS headers must look like, where every restart segment lies), compiled for the HOST and compared with the host parser
(ugb200_jpeg_debug_segments = parse_stream) - the same header the kernels include.  Property on intact, damaged and randomly mutated streams:
whatever the device logic ACCEPTS gives exactly the host parser's segment table; what it does not accept goes back to the host parser anyway."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util
from test_jpeg_decode import make_stream

SHIM = r'''
#include <vector>
#include "jpeg_marker_bounds.cuh"
extern "C" int bounds_segments(const uint8_t *s, uint32_t len, uint32_t begin0, int nscans, uint32_t comp_ids, int seg1, int seg2, int nseg, uint32_t *sb, uint32_t *se, int *tabs)
{
        using namespace ugb;
        std::vector<uint32_t> list;
        uint32_t meta[kMetaWords] = { 0 };
        meta[kMetaFirstOther] = 0xffffffffu;
        for (uint32_t p = begin0; p + 1 < len; ++p) {  // the candidates of marker_mask16 (jpeg_decode_kernels.cu), in stream order
                if (s[p] == 0xFF && s[p + 1] != 0x00 && s[p + 1] != 0xFF) {
                        const uint32_t idx = (uint32_t) list.size();
                        list.push_back(p);
                        if (s[p + 1] < 0xD0 || s[p + 1] > 0xD7) {
                                const uint32_t k = meta[kMetaOtherCount]++;
                                if (k < (uint32_t) kMaxOther) {
                                        meta[kMetaOther + (kMaxOther - 1 - k) % kMaxOther] = idx;  // the device's atomics hand the slots out in any order: reversed here
                                }
                        }
                }
        }
        if (meta[kMetaOtherCount] > 0 && meta[kMetaOtherCount] < (uint32_t) kMaxOther) {  // compact the reversed slots to the front, still unordered
                uint32_t tmp[kMaxOther], n = meta[kMetaOtherCount];
                for (uint32_t k = 0; k < n; ++k) {
                        tmp[k] = meta[kMetaOther + kMaxOther - 1 - k];
                }
                for (uint32_t k = 0; k < n; ++k) {
                        meta[kMetaOther + k] = tmp[n - 1 - (k * 5 + 3) % n];  // some permutation
                }
        }
        meta[kMetaTotal] = (uint32_t) list.size();
        list.push_back(0);  // never read
        marker_bounds(s, len, list.data(), meta, begin0, nscans, comp_ids);
        if (meta[kMetaError]) {
                return 1;
        }
        for (int i = 0; i < nseg; ++i) {
                marker_segment_multi(list.data(), meta, nscans, seg1, seg2, nseg, i, sb + i, se + i);
        }
        for (int j = 1; j < nscans; ++j) {
                tabs[2 * j] = (int) meta[kMetaBounds + 6 * j + 4], tabs[2 * j + 1] = (int) meta[kMetaBounds + 6 * j + 5];
        }
        return 0;
}
'''


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("marker_bounds")
    src, so = d / "shim.cpp", d / "libshim.so"
    src.write_text(SHIM)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(util.ROOT, "ultragrid_b200", "csrc"), str(src), "-o", str(so)],
                   check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.bounds_segments.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def layout(a):
    """what the host side of ugb200_jpeg_decode knows before the device runs: data offset of the first scan, component ids, segments per scan"""
    sos = int(np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xDA))[0])
    begin0 = sos + 2 + (int(a[sos + 2]) << 8 | int(a[sos + 3]))
    sof = int(np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xC0))[0])
    h, w = int(a[sof + 5]) << 8 | int(a[sof + 6]), int(a[sof + 7]) << 8 | int(a[sof + 8])
    ids = [int(a[sof + 10 + 3 * i]) for i in range(3)]
    dri = np.flatnonzero((a[:sos] == 0xFF) & (a[1:sos + 1] == 0xDD))
    ri = (int(a[dri[0] + 4]) << 8 | int(a[dri[0] + 5])) if len(dri) else 0
    nmcu = ((w + 7) // 8) * ((h + 7) // 8)
    per = (nmcu + ri - 1) // ri if ri else 1
    return begin0, ids[0] | ids[1] << 8 | ids[2] << 16, per


def both(lib, shim, a):
    from ultragrid_b200 import _lib
    a = np.ascontiguousarray(a)
    begin0, ids, per = layout(a)
    nseg = 3 * per
    sb, se, tabs = np.zeros(nseg, np.uint32), np.zeros(nseg, np.uint32), np.zeros(6, np.int32)
    rc = shim.bounds_segments(a.ctypes.data, len(a), begin0, 3, ids, per, 2 * per, nseg, sb.ctypes.data, se.ctypes.data, tabs.ctypes.data)
    cap = nseg + 16
    hb, he = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    n = lib.ugb200_jpeg_debug_segments(a.ctypes.data, len(a), hb.ctypes.data, he.ctypes.data, cap)
    return rc, sb, se, n, hb[:max(n, 0)], he[:max(n, 0)], nseg


@pytest.mark.parametrize("w,h,q,ri", [(64, 48, 90, 0), (200, 120, 75, 3), (333, 211, 90, 2), (640, 360, 90, 0)])
def test_device_logic_equals_host_parser_on_intact_and_mutated_streams(orc, shim, w, h, q, ri):
    from ultragrid_b200 import _lib
    lib = _lib.load()
    s, _ = make_stream(orc, "ours-rgb", w, h, q, ri)
    a0 = np.frombuffer(s, np.uint8).copy()
    rc, sb, se, n, hb, he, nseg = both(lib, shim, a0)
    assert rc == 0 and n == nseg and np.array_equal(sb, hb) and np.array_equal(se, he)  # the regular stream is accepted and laid out as the host lays it out
    begin0, _, _ = layout(a0)
    rng = np.random.default_rng(1000 + w)
    accepted = refused = 0
    for trial in range(400):
        a = a0.copy()
        kind = trial % 5
        pos = int(rng.integers(begin0, len(a) - 2))
        if kind == 0:      # a random byte
            a[pos] = rng.integers(0, 256)
        elif kind == 1:    # a random marker in the middle of things
            a[pos], a[pos + 1] = 0xFF, rng.choice([0xD0, 0xD3, 0xD7, 0xD9, 0xDA, 0xC4, 0xDB, 0xDD, 0xE0, 0x01])
        elif kind == 2:    # truncation
            a = a[:pos]
        elif kind == 3:    # a restart marker becomes data
            r = np.flatnonzero((a[begin0:-1] == 0xFF) & (a[begin0 + 1:] >= 0xD0) & (a[begin0 + 1:] <= 0xD7)) + begin0
            k = int(rng.integers(0, len(r)))
            a[r[k]:r[k] + 2] = (0x12, 0x34)
        else:              # a byte of a later SOS header
            sos = np.flatnonzero((a[:-1] == 0xFF) & (a[1:] == 0xDA))
            k = int(sos[int(rng.integers(1, len(sos)))])
            a[k + int(rng.integers(2, 10))] = rng.integers(0, 256)
        rc, sb, se, n, hb, he, nseg = both(lib, shim, a)
        if rc == 0:
            accepted += 1
            assert n == nseg, (trial, kind, n, nseg)  # accepted by the device logic => the host parser takes it too, with the same table
            assert np.array_equal(sb, hb) and np.array_equal(se, he), (trial, kind)
        else:
            refused += 1
    assert accepted > 50 and refused > 50  # both outcomes are exercised
