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


# ---- stage-level restatements of further gainmapmath_test.cpp vectors --------------------------------------------
_YUV_COEFFS = {  # gainmapmath.cpp:638-674, keyed (src_cg, dst_cg) with BT.709 = 0, P3 (BT.601 matrix) = 1, BT.2100 = 2
    (0, 1): [1.0, 0.101579, 0.196076, 0.0, 0.989854, -0.110653, 0.0, -0.072453, 0.983398],
    (0, 2): [1.0, -0.016969, 0.096312, 0.0, 0.995306, -0.051192, 0.0, 0.011507, 1.002637],
    (1, 0): [1.0, -0.118188, -0.212685, 0.0, 1.018640, 0.114618, 0.0, 0.075049, 1.025327],
    (1, 2): [1.0, -0.128245, -0.115879, 0.0, 1.010016, 0.061592, 0.0, 0.086969, 1.029350],
    (2, 0): [1.0, 0.018149, -0.095132, 0.0, 1.004123, 0.051267, 0.0, -0.011524, 0.996782],
    (2, 1): [1.0, 0.117887, 0.105521, 0.0, 0.995211, -0.059549, 0.0, -0.084085, 0.976518],
}


def _stage_impls(oracle_libs):
    return [oracle_libs.Oracle()] + ([oracle_libs.Ref()] if oracle_libs.have_ref() else [])


def test_transform_yuv420_fixture(oracle_libs):
    """gainmapmath_test.cpp:897-971 on the 4x4 fixture of :154-197: every output sample within 1 of the value the
    test computes from yuvColorGamutConversion of the four covered pixels (chroma: their mean), for all six matrices."""
    y = np.array([0x00, 0x10, 0x20, 0x30, 0x01, 0x11, 0x21, 0x31, 0x02, 0x12, 0x22, 0x32, 0x03, 0x13, 0x23, 0x33], np.uint8)
    u = np.array([0xA0, 0xA1, 0xA2, 0xA3], np.uint8)
    v = np.array([0xB0, 0xB1, 0xB2, 0xB3], np.uint8)
    buf = np.concatenate([y, u, v])
    f32 = np.float32
    for (src, dst), m in _YUV_COEFFS.items():
        m = np.array(m, f32)
        outs = [impl.convert_yuv(buf, 4, 4, src, dst) for impl in _stage_impls(oracle_libs)]
        for o in outs[1:]:
            assert (o == outs[0]).all(), (src, dst)
        o = outs[0]
        oy, ou, ov = o[:16].reshape(4, 4), o[16:20].reshape(2, 2), o[20:].reshape(2, 2)
        for cy in range(2):
            for cx in range(2):
                uu = (f32(int(u[cy * 2 + cx]) - 128)) / f32(255)
                vv = (f32(int(v[cy * 2 + cx]) - 128)) / f32(255)
                su = sv = f32(0)
                for dy in range(2):
                    for dx in range(2):
                        yy = f32(int(y[(2 * cy + dy) * 4 + 2 * cx + dx])) / f32(255)
                        ny = yy * m[0] + uu * m[1] + vv * m[2]
                        su += yy * m[3] + uu * m[4] + vv * m[5]
                        sv += yy * m[6] + uu * m[7] + vv * m[8]
                        want = int(min(max(float(ny) * 255.0 + 0.5, 0), 255))
                        assert abs(int(oy[2 * cy + dy, 2 * cx + dx]) - want) <= 1, (src, dst, cy, cx, dy, dx)
                want_u = int(min(max(float(su) / 4 * 255.0 + 128.0 + 0.5, 0), 255))
                want_v = int(min(max(float(sv) / 4 * 255.0 + 128.0 + 0.5, 0), 255))
                assert abs(int(ou[cy, cx]) - want_u) <= 1 and abs(int(ov[cy, cx]) - want_v) <= 1, (src, dst, cy, cx)


def test_yuv_gamut_conversion_of_primaries(oracle_libs):
    """gainmapmath_test.cpp:738-773: the YUV of a gamut's primaries maps to the YUV of the target gamut's primaries
    (fixture values :84-97), here through the 4:2:0 image transform on uniform 4x4 images, 8-bit quantised (+-2)."""
    prim = {0: [(0.2126, -0.11457, 0.5), (0.7152, -0.38543, -0.45415), (0.0722, 0.5, -0.04585)],
            1: [(0.299, -0.16874, 0.5), (0.587, -0.33126, -0.41869), (0.114, 0.5, -0.08131)],
            2: [(0.2627, -0.13963, 0.5), (0.6780, -0.36037, -0.45979), (0.0593, 0.5, -0.04021)]}
    for cg in prim:
        prim[cg] = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0)] + prim[cg]   # YuvBlack, YuvWhite
    q = lambda c: (int(round(c[0] * 255)), int(min(255, round(c[1] * 255 + 128))), int(min(255, round(c[2] * 255 + 128))))  # noqa: E731
    for impl in _stage_impls(oracle_libs):
        for (src, dst) in _YUV_COEFFS:
            for i in range(5):
                a, want = q(prim[src][i]), q(prim[dst][i])
                buf = np.concatenate([np.full(16, a[0]), np.full(4, a[1]), np.full(4, a[2])]).astype(np.uint8)
                o = impl.convert_yuv(buf, 4, 4, src, dst)
                got = (int(o[0]), int(o[16]), int(o[20]))
                assert all(abs(g - w) <= 2 for g, w in zip(got, want)), (src, dst, i, got, want)


def test_apply_gain_table(oracle_libs):
    """gainmapmath_test.cpp:1353-1429 (applyGain) through the applyGainMap stage: a white SDR pixel under a gain-map
    byte b comes out as min_boost^(1-b/255) * max_boost^(b/255) in every channel (the reference's table rows at
    g = 0, .25, .5, .75, 1), within the GainLUT's 1/1023 index step and half-float precision."""
    from libultrahdr_b200 import ctypes_api as A
    w, h = 16, 8
    buf = np.concatenate([np.full(w * h, 255), np.full(w * h // 2, 128)]).astype(np.uint8)
    sdr, _k = A.yuv420_image(buf, w, h, A.CG_BT709)
    for mn, mx in ((0.25, 4.0), (0.5, 2.0), (0.125, 8.0), (1.0, 8.0), (0.5, 8.0)):
        md = A.GainmapMetadata()
        for i in range(3):
            md.max_content_boost[i], md.min_content_boost[i], md.gamma[i] = mx, mn, 1.0
            md.offset_sdr[i] = md.offset_hdr[i] = 0.0
        md.hdr_capacity_min, md.hdr_capacity_max, md.use_base_cg = mn, mx, 1
        if mn < 1.0:
            md.hdr_capacity_min = 1.0   # the C API's validation wants capacity_min >= 1; the weight stays 1 at full boost
        for b in (0, 64, 128, 191, 255):
            gm = np.full((h, w, 1), b, np.uint8)
            gi = T.gm_image(gm, A.CG_BT709)
            want = math.exp2(math.log2(mn) + (b / 255.0) * (math.log2(mx) - math.log2(mn)))
            for impl in _stage_impls(oracle_libs):
                f16 = impl.apply(sdr, gi, md, A.CT_LINEAR).reshape(-1, 4)[:, :3].copy().view(np.float16).astype(np.float64)
                assert np.allclose(f16, want, rtol=1.5e-2), (mn, mx, b, float(f16[0, 0]), want)
