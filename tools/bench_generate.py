"""generateGainMap kernel times at 4K (CUDA events, one call in flight): map scale x channels x preset.
  python tools/bench_generate.py"""
import os
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
p010, yuv = bench.make_frame(W, H, 11)
sdr, _ks = A.yuv420_image(yuv, W, H, A.CG_BT709)
hdr, _kh = A.p010_image(p010, W, H, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
lib.uhdr_b200_set_kernel_timing(1)
for scale in (1, 2, 4):
    for multi in (1, 0):
        for preset, pname in ((1, "two-pass"), (0, "realtime")):
            cfg = A.default_gm_config(scale_factor=scale, multichannel=multi, preset=preset)
            gpu.generate(sdr, hdr, cfg)
            bench.kernel_report(lib)
            for _ in range(5):
                gpu.generate(sdr, hdr, cfg)
            kt = bench.kernel_report(lib)
            parts = {k: round(v[1] / v[0], 4) for k, v in kt.items() if k.startswith("gainmap")}
            print("scale %d  %s  %-9s  %s  sum %.4f ms" % (scale, "3ch" if multi else "1ch", pname, parts, sum(parts.values())))
