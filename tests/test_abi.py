"""CPU tests: the C-ABI library loads and exports every symbol include/lz4b200.h declares; host-only
helpers (bounds, chunk list, corpus) behave like the reference's; no compute is called without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lzbench_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lz4b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lz4b200_\w+|lzbench_lz4_b200_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L.lib, s), f"{s} declared in include/lz4b200.h but not exported"


def test_compress_bound_matches_reference_macro():
    # lz4/lz4.h:171
    for n in (0, 1, 12, 13, 254, 255, 256, 4096, 65536, 1 << 20, 0x7E000000):
        assert L.compress_bound(n) == n + n // 255 + 16
    assert L.compress_bound(0x7E000001) == 0
    assert L.compress_bound(65536) == 65809


def test_compress_device_bound_covers_every_chunk():
    # device entry point without a capacity argument: the documented size covers n chunks at their bound (+ segment slack)
    for n, mc in ((1, 1), (7, 4096), (16384, 65536), (3, 1 << 20), (2, 4 << 20)):
        assert L.lib.lz4b200_compress_device_bound(n, mc) >= n * L.compress_bound(mc)
    assert L.lib.lz4b200_compress_device_bound(1, 0x7E000001) == 0


def test_chunk_list_matches_lzbench():
    # _lzbench/lzbench.cpp:366-373
    assert L.chunk_list(65536 * 3, 65536).tolist() == [65536] * 3
    assert L.chunk_list(65536 * 3 + 1, 65536).tolist() == [65536] * 3 + [1]
    assert L.chunk_list([10, 70000], 65536).tolist() == [10, 65536, 4464]
    assert L.chunk_list(5, 65536).tolist() == [5]


def test_corpus_is_deterministic_and_block_independent():
    a = L.corpus("mixed", 3 * 65536, seed=2, threads=1)
    b = L.corpus("mixed", 3 * 65536, seed=2, threads=3)
    assert np.array_equal(a, b)
    c = L.corpus("mixed", 65536, seed=2, offset=65536)
    assert np.array_equal(a[65536:131072], c)
    assert not np.array_equal(a, L.corpus("mixed", 3 * 65536, seed=3))
    z = L.corpus("zero", 1000)
    assert not z.any()
    r = L.corpus("repeat", 256)
    assert np.array_equal(r[:64], r[64:128])


def test_no_cpu_fallback_without_gpu():
    if L.lib.lz4b200_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        L.Lz4B200(0)
    assert L.lib.lzbench_lz4_b200_init(65536, 0, 0) is None


def test_product_never_references_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lzbench_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".c", ".h", ".py")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle/" not in txt and "liblz4ref" not in txt, f
