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


def _lut_at(lut, x):
    n = lut.size
    return float(lut[min(n - 1, int(np.float32(x) * np.float32(n - 1) + 0.5))])


def test_transfer_function_spot_values_through_the_luts(impls, oracle_libs):
    """gainmapmath_test.cpp:1051-1061 (HlgOetf), :1063-1073 (HlgInvOetf), :1083-1093 (PqOetf), :1095-1105
    (PqInvOetf): the reference's expected values, read through the LUTs the hot path uses (node spacing of the
    4096-entry inverse tables is 2.4e-4, hence the wider tolerance there; the reference's is 1e-4)."""
    libs = [oracle_libs.Oracle()] + ([oracle_libs.Ref()] if oracle_libs.have_ref() else [])
    for impl in libs:
        hlg_inv, pq_inv, hlg, pq = impl.lut(1), impl.lut(2), impl.lut(3), impl.lut(4)
        for x, want in ((0.0, 0.0), (0.04167, 0.35357), (0.08333, 0.5), (0.5, 0.87164), (1.0, 1.0)):
            assert abs(_lut_at(hlg, x) - want) < 1e-4, ("hlgOetf", x)
        for x, want in ((0.0, 0.0), (0.01, 0.50808), (0.5, 0.92655), (0.99, 0.99895), (1.0, 1.0)):
            assert abs(_lut_at(pq, x) - want) < 2e-4, ("pqOetf", x)
        for x, want in ((0.0, 0.0), (0.25, 0.02083), (0.5, 0.08333), (0.75, 0.26496), (1.0, 1.0)):
            assert abs(_lut_at(hlg_inv, x) - want) < 5e-4, ("hlgInvOetf", x)
        for x, want in ((0.0, 0.0), (0.01, 2.31017e-7), (0.5, 0.00922), (0.99, 0.90903), (1.0, 1.0)):
            assert abs(_lut_at(pq_inv, x) - want) < 5e-3 * max(want, 0.02), ("pqInvOetf", x)
        # round trips :1075-1081, :1267-1273
        for x in (0.04167, 0.08333, 0.5):
            assert abs(_lut_at(hlg_inv, _lut_at(hlg, x)) - x) < 1e-3
        for x in (0.01, 0.5, 0.99):
            assert abs(_lut_at(pq_inv, _lut_at(pq, x)) - x) < 5e-3


def test_color_to_rgba1010102_and_f16_through_apply(oracle_libs):
    """gainmapmath_test.cpp:1555-1578 (colorToRgba1010102 / colorToRgbaF16 packing) exercised end to end: a
    unit gain map over primary-coloured SDR pixels must come out as the packed constants the reference lists
    (black 0x3<<30 / alpha-only half 0x3C00<<48, white 0xFFFFFFFF / 0x3C003C003C003C00)."""
    from libultrahdr_b200 import ctypes_api as A
    impls = [oracle_libs.Oracle()] + ([oracle_libs.Ref()] if oracle_libs.have_ref() else [])
    w, h = 16, 8
    md = A.GainmapMetadata()
    for i in range(3):
        md.max_content_boost[i] = md.min_content_boost[i] = 1.0   # gain factor exp2(0) = 1 everywhere
        md.gamma[i] = 1.0
        md.offset_sdr[i] = md.offset_hdr[i] = 0.0
    md.hdr_capacity_min, md.hdr_capacity_max, md.use_base_cg = 1.0, 2.0, 1
    gm = np.full((h, w, 1), 128, np.uint8)
    gi = T.gm_image(gm, A.CG_BT709)
    for name, yv in (("black", 0), ("white", 255)):
        buf = np.concatenate([np.full(w * h, yv), np.full(w * h // 2, 128)]).astype(np.uint8)
        sdr, _k = A.yuv420_image(buf, w, h, A.CG_BT709)
        for impl in impls:
            f16 = impl.apply(sdr, gi, md, A.CT_LINEAR).reshape(-1, 4)
            assert (f16[:, 3] == 0x3C00).all()
            assert (f16[:, :3] == (0 if name == "black" else 0x3C00)).all(), name
            pq = impl.apply(sdr, gi, md, A.CT_PQ).reshape(-1)
            assert (pq >> 30 == 3).all()
            if name == "black":
                assert (pq == np.uint32(0x3 << 30)).all()
