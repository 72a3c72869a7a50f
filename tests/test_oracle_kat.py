"""Known-answer vectors the reference's own unit tests hold for this path
(/root/reference/tests/gainmapmath_test.cpp), checked against the C restatement (and the reference
build when present)."""
import ctypes as C
import math

import numpy as np
import pytest

import uhdr_testlib as T


@pytest.fixture(scope="module")
def impls(oracle_libs):
    out = [("oracle", oracle_libs.Oracle().lib, "uo_")]
    if oracle_libs.have_ref():
        out.append(("ref", oracle_libs.Ref().lib, "ref_"))
    for _, lib, p in out:
        getattr(lib, p + "compute_gain").restype = C.c_float
        getattr(lib, p + "compute_gain").argtypes = [C.c_float, C.c_float]
        getattr(lib, p + "affine_map_gain").argtypes = [C.c_float] * 4
        getattr(lib, p + "float_to_half").argtypes = [C.c_float]
        getattr(lib, p + "srgb_oetf").restype = C.c_float
        getattr(lib, p + "srgb_oetf").argtypes = [C.c_float]
    return out


def test_float_to_half_vectors(impls):  # gainmapmath_test.cpp:1580-1588
    fmax = float(np.finfo(np.float32).max)
    vec = [(0.1, 0x2E66), (0.0, 0x0), (1.0, 0x3C00), (-1.0, 0xBC00), (fmax, 0x7FFF), (-fmax, 0xFFFF),
           (float(np.float32(2.0) ** -126), 0x0), (0.2, 0x3266), (0.3, 0x34CD)]  # :1576 0x3C0034CD32662E66
    for name, lib, p in impls:
        for f, want in vec:
            assert getattr(lib, p + "float_to_half")(f) == want, (name, f)


def test_affine_map_of_compute_gain_table(impls):  # gainmapmath_test.cpp:1297-1351
    l2 = lambda x: float(np.float32(math.log2(x)))  # noqa: E731
    table = [
        (l2(0.25), l2(4.0), [(0, 1, 255), (1, 0, 0), (0.5, 0, 0), (1, 1, 128), (1, 4, 255), (1, 5, 255), (4, 1, 0),
                             (4, 0.5, 0), (1, 2, 191), (2, 1, 64)]),
        (l2(0.5), l2(2.0), [(1, 2, 255), (2, 1, 0), (1, 1.41421, 191), (1.41421, 1, 64)]),
        (l2(0.125), l2(8.0), [(1, 8, 255), (8, 1, 0), (1, 2.82843, 191), (2.82843, 1, 64)]),
        (l2(1.0), l2(8.0), [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 8, 255), (1, 4, 170), (1, 2, 85)]),
        (l2(0.5), l2(8.0), [(0, 0, 64), (1, 0, 0), (1, 1, 64), (1, 8, 255), (1, 4, 191), (1, 2, 127), (1, 0.7071, 32),
                            (1, 0.5, 0)]),
    ]
    for name, lib, p in impls:
        for mn, mx, rows in table:
            for sdr, hdr, want in rows:
                g = getattr(lib, p + "compute_gain")(sdr, hdr)
                assert getattr(lib, p + "affine_map_gain")(g, mn, mx, 1.0) == want, (name, mn, mx, sdr, hdr)


def test_srgb_oetf_spot_values(impls):  # gainmapmath_test.cpp:1051-1105 (tolerance 1e-4 there)
    for name, lib, p in impls:
        f = getattr(lib, p + "srgb_oetf")
        assert abs(f(0.0) - 0.0) < 1e-6
        assert abs(f(1.0) - 1.0) < 1e-6
        assert abs(f(0.0031308) - 0.04045) < 1e-4
        assert abs(f(0.5) - 0.735357) < 1e-4


def test_lut_nodes_match_functions(oracle_libs):  # gainmapmath_test.cpp:1107-1140
    o = oracle_libs.Oracle()
    srgb = o.lut(0)
    x = np.arange(1024, dtype=np.float32) / np.float32(1023)
    want = np.where(x <= 0.04045, x / np.float32(12.92), ((x.astype(np.float64) + 0.055) / 1.055) ** 2.4)
    assert np.abs(srgb - want).max() < 1e-6
    pq = o.lut(4)
    assert pq[0] == 0.0 and abs(pq[-1] - 1.0) < 1e-6 and (np.diff(pq) >= 0).all()
    hlg = o.lut(3)
    assert hlg[0] == 0.0 and abs(hlg[-1] - 1.0) < 1e-5 and (np.diff(hlg) >= 0).all()
