"""Golden vectors produced by the reference's own code on a crop of its own fixtures
(tools/make_golden.py).  CPU: the C restatement reproduces them.  GPU: the CUDA path reproduces
them through the C ABI, including the complete API-1 file byte for byte."""
import os

import numpy as np
import pytest

import uhdr_testlib as T
from libultrahdr_b200 import ctypes_api as A

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "uhdr_golden_320x192.npz"))
W, H = 320, 192
CASES = (("default", {}), ("s4_single", {"scale_factor": 4, "multichannel": 0}), ("onepass", {"preset": A.USAGE_REALTIME}))


def _inputs():
    hb, sb = G["p010"].copy(), G["yuv420"].copy()
    hdr, k1 = A.p010_image(hb, W, H, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
    sdr, k2 = A.yuv420_image(sb, W, H, A.CG_BT709)
    return hdr, sdr, (hb, sb, k1, k2)


def _check_stages(impl, tonemap_exact=True):
    hdr, sdr, keep = _inputs()
    for name, kw in CASES:
        g, m = impl.generate(sdr, hdr, A.default_gm_config(**kw))
        assert (g == G["gm_" + name]).all(), name
        assert bytes(m) == G["md_" + name].tobytes(), name
        gi = T.gm_image(g, A.CG_BT2100)
        assert (impl.apply(sdr, gi, m, A.CT_LINEAR) == G["apply_f16_" + name]).all(), name
        assert (impl.apply(sdr, gi, m, A.CT_PQ) == G["apply_pq_" + name]).all(), name
    tm = impl.tonemap(hdr)[0]
    d = np.abs(tm.astype(int) - G["tonemap"].astype(int))
    assert d.max() == 0  # srgbOetf's powf is glibc's, operation for operation, on the device too
    assert (impl.convert_yuv(G["yuv420"].copy(), W, H, 0, 1) == G["convert_709_601"]).all()


def test_oracle_reproduces_golden(oracle_libs):
    _check_stages(oracle_libs.Oracle())


@pytest.mark.gpu
def test_gpu_reproduces_golden_stages(gpu):
    _check_stages(gpu)


@pytest.mark.gpu
def test_gpu_reproduces_golden_file_and_decode(gpu):
    api = T.UhdrApi(gpu.lib)
    hdr, sdr, keep = _inputs()
    data = api.encode(hdr, sdr)
    assert data == G["file_api1"].tobytes()
    px, gm, md, cg = api.decode(G["file_api1"].tobytes())
    assert (px == G["decoded_f16"]).all()
