"""The screened kernels (k_affine_q, the one-pass gain code, toneMap's srgbOetf) decide from an approximate value
whether the exact one could round differently.  Their thresholds are derived by hand in the kernel comments: an error
bound of the hardware approximation (measured over every input on the device, GPU tests) carried through the float
operations behind it.  Here the float32 operation sequences are replayed with numpy on a few million random inputs and
worst-case-signed perturbations: the observed shift of the pre-rounding value must stay within the threshold the
kernels use (constants copied from gainmap_fast.cu / tonemap_fast.cu; keep them in step)."""
import ctypes as C
import ctypes.util

import numpy as np

f32 = np.float32
libm = C.CDLL(ctypes.util.find_library("m"))
libm.fmaf.restype = C.c_float
libm.fmaf.argtypes = [C.c_float] * 3
vfmaf = np.vectorize(libm.fmaf, otypes=[np.float32])

K_LG2_ABS, K_LG2_REL, K_AFFINE_ROUND = f32(8e-7), f32(3.2e-7), f32(2e-4)
K_POW_ABS = f32(3.0e-7)


def _within(g, e, sign):
    """the float32 furthest from g in direction `sign` whose distance to g does not exceed e (what the approximation
    may return: it is itself a float)"""
    g2 = (g.astype(np.float64) + sign * e).astype(np.float32)
    over = np.abs(g2.astype(np.float64) - g.astype(np.float64)) > e
    g2[over] = np.nextafter(g2[over], g[over])
    return g2


def _refined_rcp(d):
    r = (f32(1) / d).astype(np.float32)          # stand-in for rcp.approx + one Newton step: any 1-ulp reciprocal
    return vfmaf(r, vfmaf(-d, r, np.full_like(d, 1, np.float32)), r)


def _affine_t(g, mn, d, rc):
    a = (g - mn).astype(np.float32)
    q = (a * rc).astype(np.float32)
    q = vfmaf(rc, vfmaf(-d, q, a), q)
    return ((q * f32(255)).astype(np.float32) + f32(0.5)).astype(np.float32)


def test_affine_screen_threshold_covers_the_shift():
    rs = np.random.RandomState(7)
    n = 300000
    mn = rs.uniform(-14.3, 2.0, n).astype(np.float32)
    d = rs.uniform(0.1, 29.9, n).astype(np.float32)
    mx = (mn + d).astype(np.float32)
    d = (mx - mn).astype(np.float32)
    rc = _refined_rcp(d)
    g = (mn + d * rs.uniform(-0.002, 1.002, n).astype(np.float32)).astype(np.float32)   # t in about [-0.5, 256]
    gmax = np.maximum(np.abs(mn), np.abs(mx)) + f32(1)
    thr = (K_LG2_ABS + gmax * K_LG2_REL) * (f32(255) * rc * f32(1.0001)) + K_AFFINE_ROUND
    t = _affine_t(g, mn, d, rc).astype(np.float64)
    for sign in (-1.0, 1.0):
        e = (K_LG2_ABS + np.abs(g) * K_LG2_REL).astype(np.float64)
        g2 = _within(g, e, sign)
        t2 = _affine_t(g2, mn, d, rc).astype(np.float64)
        worst = np.max(np.abs(t2 - t) / thr)
        assert worst <= 1.0, worst


def test_onepass_screen_threshold_covers_the_shift():
    rs = np.random.RandomState(8)
    n = 300000
    lmin = np.zeros(n, np.float32)
    lmax = rs.choice(np.array([2.3005, 5.6224, 1.0, 3.0], np.float32), n)
    rng64 = (lmax - lmin).astype(np.float64)
    inv_f = (1.0 / rng64).astype(np.float32)
    g = (rs.uniform(0.0, 1.0, n) * rng64).astype(np.float64)            # exact log2 of the gain (double)
    t_exact = ((g - lmin) / rng64).astype(np.float32) * f32(255)          # float(double quotient) * 255
    gmax = np.maximum(np.abs(lmin), np.abs(lmax))
    thr = f32(255) * (K_LG2_ABS + gmax * K_LG2_REL) * np.abs(inv_f) * f32(1.0001) + f32(1.5e-4)
    for sign in (-1.0, 1.0):
        gf = _within(g.astype(np.float32), (K_LG2_ABS + np.abs(g) * K_LG2_REL).astype(np.float64), sign)   # what lg2.approx may return
        t_fast = (((gf - lmin).astype(np.float32) * inv_f).astype(np.float32) * f32(255)).astype(np.float32)
        worst = np.max(np.abs(t_fast.astype(np.float64) - t_exact.astype(np.float64)) / thr)
        assert worst <= 1.0, worst


def test_tonemap_screen_thresholds_cover_the_shift():
    rs = np.random.RandomState(9)
    n = 200000
    lin = rs.uniform(0.0031308, 1.0, (n, 4, 3))                       # a 2x2 group of linear sRGB triples
    p = (lin ** (1 / 2.4)).astype(np.float32)                          # stand-in for the exact powf
    r_cb = _refined_rcp(np.full(n, 1.772, np.float32))
    r_cr = _refined_rcp(np.full(n, 1.402, np.float32))

    def codes_pre(pw):
        e = ((f32(1.055) * pw).astype(np.float32) - f32(0.055)).astype(np.float32)
        er, eg, eb = e[..., 0], e[..., 1], e[..., 2]
        yy = (((f32(0.299) * er).astype(np.float32) + (f32(0.587) * eg).astype(np.float32)).astype(np.float32)
              + (f32(0.114) * eb).astype(np.float32)).astype(np.float32)

        def div_by(a, rcp, dv):
            q = (a * rcp[:, None]).astype(np.float32)
            return vfmaf(rcp[:, None] + np.zeros_like(q), vfmaf(np.full_like(q, -dv), q, a), q)
        uo = (div_by((eb - yy).astype(np.float32), r_cb, f32(1.772)) + f32(0.5)).astype(np.float32)
        vo = (div_by((er - yy).astype(np.float32), r_cr, f32(1.402)) + f32(0.5)).astype(np.float32)
        su = np.zeros(n, np.float32)
        sv = np.zeros(n, np.float32)
        for i in range(4):
            su = (su + uo[:, i]).astype(np.float32)
            sv = (sv + vo[:, i]).astype(np.float32)
        return ((yy * f32(255)).astype(np.float64), ((su * f32(0.25)) * f32(255)).astype(np.float64),
                ((sv * f32(0.25)) * f32(255)).astype(np.float64))
    k_e = f32(1.055) * K_POW_ABS + f32(1.3e-7)
    k_ey = k_e + f32(2.0e-7)
    thr_y = f32(255) * k_ey + f32(1.6e-5)
    thr_u = f32(255) * ((k_e + k_ey) / f32(1.772) + f32(2.0e-7) + f32(1.3e-7)) + f32(1.6e-5)
    thr_v = f32(255) * ((k_e + k_ey) / f32(1.402) + f32(2.0e-7) + f32(1.3e-7)) + f32(1.6e-5)
    y0, u0, v0 = codes_pre(p)
    for trial in range(6):
        # perturbation patterns that push luma / Cb / Cr the furthest, plus random signs
        if trial == 0:
            sgn = np.ones((n, 4, 3))
        elif trial == 1:
            sgn = -np.ones((n, 4, 3))
        elif trial == 2:
            sgn = np.tile(np.array([-1.0, -1.0, 1.0]), (n, 4, 1))     # b up, r and g down: Cb
        elif trial == 3:
            sgn = np.tile(np.array([1.0, -1.0, -1.0]), (n, 4, 1))     # r up, g and b down: Cr
        else:
            sgn = rs.choice([-1.0, 1.0], (n, 4, 3))
        y1, u1, v1 = codes_pre((p + sgn * K_POW_ABS).astype(np.float32))
        assert np.max(np.abs(y1 - y0)) <= thr_y, (trial, np.max(np.abs(y1 - y0)), thr_y)
        assert np.max(np.abs(u1 - u0)) <= thr_u, (trial, np.max(np.abs(u1 - u0)), thr_u)
        assert np.max(np.abs(v1 - v0)) <= thr_v, (trial, np.max(np.abs(v1 - v0)), thr_v)
