"""CPU-side pins of three pieces of reasoning the kernels rely on (no GPU, no product code involved):
  * the 8-operation FP64 form of the YCoCg transform in dxt6_device.cuh equals the reference's 11-operation form on every possible pixel;
  * the SWAR count -> index map of the DXT5 alpha indices equals the per-pixel formula of cuda_dxt.cu:376-392, including the 48-bit layout;
  * the shared-memory column rotation of the fused JPEG kernel (blk_col) is free of bank conflicts in both access patterns;
  * the per-segment routine of the JPEG compact kernel (csrc/jpeg_compact.cuh: aligned word stores through a funnel shift), compiled for the host,
    equals memcpy for every size and alignment and touches no byte outside its segment."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ycocg_fp64_reassociation_is_exact_on_every_pixel(tmp_path):
    exe = tmp_path / "ycocg_identity"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "oracle", "ycocg_identity.c"), "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "33554432 triples, 0 differ" in r.stdout, r.stdout


def alpha_count_to_index(x):
    """alpha_count_to_index() of dxt6_device.cuh on ten 3-bit fields"""
    lsb, m = 0x09249249, 0xFFFFFFFF
    x1, x2 = x >> 1, x >> 2
    all7 = x & x1 & x2 & lsb
    inc = (x | x1 | x2) & lsb & ~all7 & m
    return (((x & ~(all7 * 7) & m) + inc) | all7) & m


def test_alpha_swar_map_and_layout_equal_the_per_pixel_formula():
    rng = np.random.default_rng(5)
    m = 0xFFFFFFFF
    cases = [[c] * 16 for c in range(8)] + [list(rng.integers(0, 8, 16)) for _ in range(20000)]
    for cnts in cases:
        cnts = [int(c) for c in cnts]
        ix = iy = 0
        for i, cnt in enumerate(cnts):  # cuda_dxt.cu:376-392: index = 1 + count, & 7, ^ (2 > index); pixels 0..5 in word 0 from bit 16, the rest in word 1
            idx = (1 + cnt) & 7
            idx ^= 1 if 2 > idx else 0
            if i < 6:
                ix |= (idx << (3 * i + 16)) & m
            if i == 5:
                iy = idx >> 1
            if i > 5:
                iy |= (idx << (3 * i - 16)) & m
        a = sum(c << (3 * i) for i, c in enumerate(cnts[:10]))
        b = sum(c << (3 * i) for i, c in enumerate(cnts[10:]))
        ia, ib = alpha_count_to_index(a), alpha_count_to_index(b)
        s_lo, s_hi = (ia | (ib << 30)) & m, ib >> 2
        assert ((s_lo << 16) & m, ((s_lo >> 16) | (s_hi << 16)) & m) == (ix, iy), cnts


def test_jpeg_block_columns_are_conflict_free():
    def blk_col(p):
        return (p & ~31) | ((p + (p >> 5)) & 31)
    assert sorted(blk_col(p) for p in range(128)) == list(range(128))  # a permutation of the 128 columns
    for k in range(4):  # phases 1-2: lane l of the warp of component k works on block 4 l + k
        assert len({blk_col(4 * lane + k) % 32 for lane in range(32)}) == 32
    for w in range(4):  # phase 3: thread tid works on block tid
        assert len({blk_col(32 * w + lane) % 32 for lane in range(32)}) == 32


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc (host compilation of a .cu file)")
def test_jpeg_compact_segment_routine_equals_memcpy(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "exp_compact"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), os.path.join(ROOT, "tools", "exp_compact.cu")],
                   check=True, capture_output=True)
    r = subprocess.run([str(exe), "check"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cpu check ok" in r.stdout, r.stdout + r.stderr

