"""Per-kernel CUDA-event times of one 4K API-0 encode (HDR intent only: toneMap + one-pass generateGainMap + two JPEGs).
  python tools/bench_api0_kernels.py [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import uhdr_testlib as T  # noqa: E402
from libultrahdr_b200 import ctypes_api as A  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
gpu = T.Gpu()
lib = gpu.lib
T.UhdrApi(lib)
p010, _ = bench.make_frame(bench.W4K, bench.H4K, 99)
hdr, _k = A.p010_image(p010, bench.W4K, bench.H4K, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
sl = bench.EncoderSlot(lib)
sl.set_inputs(hdr, None)
lib.uhdr_b200_set_kernel_timing(1)
for _ in range(3):
    sl.encode()
    sl.rearm()
bench.kernel_report(lib)
for _ in range(n):
    sl.rearm()
    sl.encode()
kt = bench.kernel_report(lib)
tot = 0.0
for k, v in sorted(kt.items()):
    tot += v[1] / n
    print("%-18s launches/frame %.1f  avg %.4f ms  per frame %.4f ms" % (k, v[0] / n, v[1] / v[0], v[1] / n))
print("sum per frame %.4f ms" % tot)
st = (__import__("ctypes").c_ulonglong * 2)()
lib.uhdr_b200_tonemap_stats(st)
print("tonemap groups %d, redone exactly %d (%.3f %%)" % (st[0], st[1], 100.0 * st[1] / max(1, st[0])))
