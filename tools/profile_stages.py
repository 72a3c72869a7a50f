"""Small driver for ncu captures: one 8K applyGainMap (config 3 geometry) and a few 4K API-1
encodes through the C API.  Usage (under gpurun):
  ncu --set full --clock-control none --import-source on -k regex:k_apply -c 2 -o gpurun_out/apply python tools/profile_stages.py apply
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import uhdr_testlib as T  # noqa: E402
from libultrahdr_b200 import ctypes_api as A  # noqa: E402
import bench  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
gpu = T.Gpu()
if what in ("apply", "apply_nat", "all"):
    W, H = 7680, 4320
    if what == "apply_nat":   # the bench's "natural" content: smooth + texture
        _p, sb = bench.make_frame(W, H, 5)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        g0 = 128 + 90 * np.sin(xx / 301.0) * np.cos(yy / 257.0) + np.random.RandomState(3).randn(H, W) * 2
        gm = np.stack([g0, g0 * 0.9 + 10, g0 * 0.8 + 20, np.full_like(g0, 255)], -1).clip(0, 255).astype(np.uint8)
    else:
        sb = T.make_yuv420(W, H, "noise")
        gm = np.random.RandomState(7).randint(0, 256, (H, W, 4)).astype(np.uint8)
    sdr, k2 = A.yuv420_image(sb, W, H, A.CG_BT709)
    md = A.GainmapMetadata()
    for i, (mx, mn) in enumerate(((65.1, 4.9e-5), (845.9, 2.7e-3), (1283.8, 4.9e-5))):
        md.max_content_boost[i], md.min_content_boost[i], md.gamma[i] = mx, mn, 1.0
        md.offset_sdr[i] = md.offset_hdr[i] = 1e-7
    md.hdr_capacity_min, md.hdr_capacity_max, md.use_base_cg = 1.0, 4.926108, 0
    gi = T.gm_image(gm, A.CG_BT2100)
    for _ in range(3):
        gpu.apply(sdr, gi, md, A.CT_LINEAR)
if what in ("encode", "all"):
    api = T.UhdrApi(gpu.lib)
    p, y = bench.make_frame(3840, 2160, 0)
    hdr, sdr, keep = bench.frame_descs(p, y, 3840, 2160)
    for _ in range(3):
        api.encode(hdr, sdr)
    if what == "all":
        api.encode(hdr, None)
        # JpegR's own defaults: map scale 4, one channel (k_gainmap_scaled), two-pass and realtime
        api.encode(hdr, sdr, scale=4, multichannel=0)
        api.encode(hdr, sdr, scale=4, multichannel=0, preset=A.USAGE_REALTIME)
print("done")
