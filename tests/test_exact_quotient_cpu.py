"""encode_gain_norm (gainmap_fast.cu) replaces the reference's fp64 division a / b by three fp64 operations with the
host's correctly rounded reciprocal: q = a*y, r = fma(-q, b, a), q' = fma(r, y, q).  The device performs exactly these
IEEE operations, so the claim -- q' equals the correctly rounded quotient whenever b is a float widened to double --
can be checked on the CPU with libm's fma against the hardware division, here on a few million operand pairs
including divisors with extreme float mantissas and dividends one ulp around multiples of the divisor."""
import ctypes as C
import ctypes.util

import numpy as np

libm = C.CDLL(ctypes.util.find_library("m"))
libm.fma.restype = C.c_double
libm.fma.argtypes = [C.c_double] * 3
vfma = np.vectorize(libm.fma, otypes=[np.float64])


def _check(a, b):
    y = 1.0 / b                     # IEEE division: correctly rounded reciprocal
    q = a * y
    r = vfma(-q, b, a)
    q2 = vfma(r, y, q)
    ref = a / b
    bad = ~((q2 == ref) | (np.isnan(q2) & np.isnan(ref)))
    assert not bad.any(), (a[bad][:4], b[bad][:4], q2[bad][:4], ref[bad][:4])


def test_three_operation_quotient_equals_division():
    rs = np.random.RandomState(12345)
    n = 400000
    # divisors: floats widened to double (what log2_max - log2_min is), incl. all-ones and one-bit mantissas
    b = rs.uniform(0.05, 40.0, n).astype(np.float32)
    special = np.array([np.float32(1.0), np.nextafter(np.float32(2.0), np.float32(0)), np.float32(2.3219280), np.float32(5.6147099),
                        np.nextafter(np.float32(1.0), np.float32(2)), np.float32(0.1), np.float32(3.0), np.float32(1.5)], np.float32)
    b[:special.size] = special
    b = b.astype(np.float64)
    # dividends: doubles of the magnitude log2(gain) - log2_min takes, both signs
    a = rs.uniform(-45.0, 45.0, n)
    _check(a, b)
    # dividends next to exact multiples k/255 of the divisor (the byte boundaries of the one-pass code)
    # (k >= 1: around 0 the neighbours are subnormal, where the sequence loses its guarantee -- on the device the
    # dividend is log2(gain) - log2_min: exactly 0 or at least 2^-52 * |log2_min| resp. 1.7e-7 in magnitude)
    k = rs.randint(1, 256, n).astype(np.float64)
    base = b * (k / 255.0)
    for ulps in (-2, -1, 0, 1, 2):
        aa = base.copy()
        for _ in range(abs(ulps)):
            aa = np.nextafter(aa, np.inf if ulps > 0 else -np.inf)
        _check(aa, b)
    # zero, tiny (normal) and huge dividends
    _check(np.zeros(n), b)
    _check(rs.uniform(-1e-12, 1e-12, n), b)
    _check(rs.uniform(-1e9, 1e9, n), b)
