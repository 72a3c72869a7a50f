"""How often do the screened kernels take their exact paths?  4K bench content and 4K noise.
  python tools/probe_screens.py"""
import ctypes as C
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import uhdr_testlib as T  # noqa: E402
from libultrahdr_b200 import ctypes_api as A  # noqa: E402

gpu = T.Gpu()
lib = gpu.lib
W, H = bench.W4K, bench.H4K


def gstats():
    st = (C.c_ulonglong * 2)()
    lib.uhdr_b200_generate_stats(st)
    return st[0], st[1]


def tstats():
    st = (C.c_ulonglong * 2)()
    lib.uhdr_b200_tonemap_stats(st)
    return st[0], st[1]


first = struct.unpack("<I", struct.pack("<f", 0.0031308))[0]
last = struct.unpack("<I", struct.pack("<f", 1.0))[0]
w = C.c_float(-1)
assert lib.uhdr_b200_probe_pow_fast(C.c_uint(first), C.c_uint(last - first + 1), C.byref(w)) == 0
print("pow(e, 1/2.4) approximation: worst |error| over all %d inputs = %.3e (screen assumes 3e-7)" % (last - first + 1, w.value))
f0 = (127 - 40) << 23
n = ((127 + 40) << 23) - f0
w = C.c_float(-1)
assert lib.uhdr_b200_probe_log2_fast(C.c_uint(f0), C.c_uint(n), C.byref(w)) == 0
print("lg2.approx: worst error / bound over all %d floats of [2^-40, 2^40] = %.3f (must stay <= 0.5)" % (n, w.value))
for name in ("bench", "noise"):
    if name == "bench":
        p010, yuv = bench.make_frame(W, H, 11)
    else:
        p010, yuv = T.make_p010(W, H, "noise"), T.make_yuv420(W, H, "noise")
    sdr, _ks = A.yuv420_image(yuv, W, H, A.CG_BT709)
    hdr, _kh = A.p010_image(p010, W, H, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
    v0, e0 = gstats()
    gpu.generate(sdr, hdr)
    v1, e1 = gstats()
    print("%-5s two-pass generate: %d values, %d through the fp64 log2 (%.4f %%)" % (name, v1 - v0, e1 - e0, 100.0 * (e1 - e0) / max(1, v1 - v0)))
    g0, x0 = tstats()
    gpu.tonemap(hdr)
    g1, x1 = tstats()
    print("%-5s toneMap: %d groups, %d redone with the exact powf (%.4f %%)" % (name, g1 - g0, x1 - x0, 100.0 * (x1 - x0) / max(1, g1 - g0)))
