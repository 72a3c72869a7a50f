"""uhdr_decode timing at 4K and 8K with the host and the device entropy decoder:
  python tools/bench_decode.py [4k|8k|both]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import uhdr_testlib as T  # noqa: E402
from libultrahdr_b200 import ctypes_api as A  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
gpu = T.Gpu()
lib = gpu.lib
api = T.UhdrApi(lib)
lib.uhdr_b200_entropy_decoder_stats.restype = None


def stats():
    st = (C.c_ulonglong * 3)()
    lib.uhdr_b200_entropy_decoder_stats(st)
    return list(st)


for name, (w, h) in (("4k", (3840, 2160)), ("8k", (7680, 4320))):
    if what not in (name, "both"):
        continue
    p, y = bench.make_frame(w, h, 7)
    hdr, sdr, keep = bench.frame_descs(p, y, w, h)
    data = api.encode(hdr, sdr)
    buf = np.frombuffer(data, np.uint8).copy()
    ci = A.CompressedImage(buf.ctypes.data, len(data), len(data), -1, -1, -1)
    for mode in (1, 2):
        lib.uhdr_b200_set_entropy_decoder(mode)
        ts = []
        for it in range(5):
            if it == 4:
                lib.uhdr_b200_set_kernel_timing(1)
                bench.kernel_report(lib)
            dec = C.c_void_p(lib.uhdr_create_decoder())
            t0 = time.perf_counter()
            assert lib.uhdr_dec_set_image(dec, C.byref(ci)).error_code == 0
            e = lib.uhdr_decode(dec)
            assert e.error_code == 0, e.detail
            ts.append(time.perf_counter() - t0)
            lib.uhdr_release_decoder(dec)
        kt = bench.kernel_report(lib)
        lib.uhdr_b200_set_kernel_timing(0)
        print(name, "mode", mode, "stream", len(data), "ms", [round(t * 1e3, 2) for t in ts], "stats", stats())
        print("   kernels(ms, last run):", {k: round(v[1], 3) for k, v in kt.items()})
    lib.uhdr_b200_set_entropy_decoder(0)
