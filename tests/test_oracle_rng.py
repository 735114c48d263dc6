"""Pins the oracle's random number generators (no GPU)."""
import numpy as np
import pytest


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = oracle.philox4x32_10(ctr, key)
        assert [int(v) for v in got] == want


def test_sqb_stream_layout(oracle):
    seed, step = 0x1234567890abcdef, (5 << 32) | 77
    got = oracle.sqb_philox(seed, step, oracle.DOM_BG_SIDE1, 11, 3)
    want = oracle.philox4x32_10([11, 3, 77, (3 << 24) | 5], [seed & 0xffffffff, seed >> 32])
    assert np.array_equal(got, want)


def test_mt19937_known_answers(oracle):
    # mt19937ar reference output for init_genrand(5489): first outputs 3499211612, 581869302, 3890346734 ...
    s = oracle.mt_stream(5489, 5)
    assert [int(v) for v in s[:3]] == [3499211612, 581869302, 3890346734]
    # numpy's legacy seeding is init_genrand on a 32-bit seed as well
    for seed in (1, 13255, 1133557):
        ours = oracle.mt_stream(seed, 2000)
        rs = np.random.RandomState(seed)
        theirs = rs.randint(0, 1 << 32, 2000, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(ours, theirs)


def test_mt_against_reference_object_code(oracle):
    ref = oracle.reflib()
    if ref is None:
        pytest.skip('oracle/_ref not built (reference tree absent)')
    import ctypes as C
    for seed in (0, 7, 1133557, (1 << 40) + 5):
        n = 3000
        want = np.empty(n, np.uint32)
        ref.ref_mt_stream(C.c_ulonglong(seed), n, want.ctypes.data_as(C.c_void_p))
        assert np.array_equal(oracle.mt_stream(seed, n), want)
        wf = np.empty(n, np.float32); wd = np.empty(n, np.float64)
        ref.ref_mt_reals(C.c_ulonglong(seed), n, wf.ctypes.data_as(C.c_void_p), wd.ctypes.data_as(C.c_void_p))
        f, d = oracle.mt_reals(seed, n)
        assert np.array_equal(f, wf) and np.array_equal(d, wd)
        assert f.min() >= 0 and d.min() >= 0 and d.max() < 1


def test_dot_product_against_reference_object_code(oracle):
    """the oracle's AVX2 dot product (oracle.cpp dotv: 'same summation tree') vs the reference's own cpu/Dot_SIMD.cpp, compiled where it
    lies into oracle/_ref: bit-identical on 64-byte aligned, zero-padded rows of every length class (the reference reads whole
    cache lines, Matrix.cpp:7-14 clears the padding)"""
    ref = oracle.reflib()
    if ref is None:
        pytest.skip('oracle/_ref not built (reference tree absent)')
    import ctypes as C
    lib = oracle.lib()
    lib.orc_dot_f32.restype = C.c_float; lib.orc_dot_f64.restype = C.c_double
    ref.ref_dot_f32.restype = C.c_float; ref.ref_dot_f64.restype = C.c_double
    rng = np.random.default_rng(5)

    def aligned(n, dtype):
        per = 64 // np.dtype(dtype).itemsize
        cap = (n + per - 1) // per * per + per
        raw = np.zeros(cap * np.dtype(dtype).itemsize + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        a = raw[off:off + cap * np.dtype(dtype).itemsize].view(dtype)
        a[:n] = rng.standard_normal(n).astype(dtype)
        return a, raw
    for n in (1, 7, 8, 15, 16, 17, 100, 1000, 1024, 4097, 8192):
        for dtype, fo, fr in ((np.float32, lib.orc_dot_f32, ref.ref_dot_f32), (np.float64, lib.orc_dot_f64, ref.ref_dot_f64)):
            a, ka = aligned(n, dtype); b, kb = aligned(n, dtype)
            pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
            got, want = fo(pa, pb, n), fr(pa, pb, n)
            assert got == want, (n, dtype, got, want)
