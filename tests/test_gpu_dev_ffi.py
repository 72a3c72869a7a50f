"""Device-pointer stage entry points (uhdr_b200_*_dev, include/uhdr_b200.h): planes in HBM, caller's
stream, no PCIe traffic.  Chained generateGainMap -> applyGainMap -> toneMap -> convertYuv on torch device
tensors equal the reference's results for the same inputs (bit-exact)."""
import ctypes as C

import numpy as np
import pytest

import uhdr_testlib as T
from libultrahdr_b200 import ctypes_api as A

pytestmark = pytest.mark.gpu


def _dev_desc(torch, fmt, cg, ct, rng, w, h, tensors, strides):
    img = A.RawImage()
    img.fmt, img.cg, img.ct, img.range, img.w, img.h = fmt, cg, ct, rng, w, h
    for i in range(3):
        img.planes[i] = tensors[i].data_ptr() if i < len(tensors) else None
        img.stride[i] = strides[i] if i < len(strides) else 0
    return img


@pytest.mark.parametrize("preset", [A.USAGE_BEST_QUALITY, A.USAGE_REALTIME])
def test_chained_stages_on_device_memory(gpu, checker, preset):
    import torch
    lib = gpu.lib
    w, h = 640, 368
    hb, sb = T.make_p010(w, h, "smooth"), T.make_yuv420(w, h, "smooth")
    hdr_h, k1 = A.p010_image(hb, w, h, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
    sdr_h, k2 = A.yuv420_image(sb, w, h, A.CG_BT709)
    cfg = A.default_gm_config()
    cfg.preset = preset
    want_gm, want_md = checker.generate(sdr_h, hdr_h, cfg)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        th = torch.from_numpy(hb.view(np.int16).copy()).cuda()          # P010: Y then interleaved UV
        ts = torch.from_numpy(sb.copy()).cuda()
        hdr_d = _dev_desc(torch, A.FMT_P010, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED, w, h, [th, th[w * h:]], [w, w])
        sdr_d = _dev_desc(torch, A.FMT_YUV420, A.CG_BT709, A.CT_SRGB, A.CR_FULL, w, h, [ts, ts[w * h:], ts[w * h * 5 // 4:]], [w, w // 2, w // 2])
        gm_t = torch.zeros(h * w * 3, dtype=torch.uint8, device="cuda")
        gm_d = _dev_desc(torch, A.FMT_RGB888, -1, -1, -1, w, h, [gm_t], [w])
        md = A.GainmapMetadata()
        sp = C.c_void_p(st.cuda_stream)
        rc = lib.uhdr_b200_generate_gainmap_dev(C.byref(sdr_d), C.byref(hdr_d), C.byref(cfg), C.byref(md), C.byref(gm_d), sp)
        assert rc == 0, T.gpu_err(gpu)
        # the map never leaves HBM: applyGainMap reads it where generateGainMap wrote it
        out_t = torch.zeros(h * w * 4, dtype=torch.int16, device="cuda")
        dst_d = _dev_desc(torch, A.FMT_RGBAF16, -1, A.CT_LINEAR, A.CR_FULL, w, h, [out_t], [w])
        rc = lib.uhdr_b200_apply_gainmap_dev(C.byref(sdr_d), C.byref(gm_d), C.byref(md), A.CT_LINEAR, C.c_float(A.FLT_MAX), C.byref(dst_d), sp)
        assert rc == 0, T.gpu_err(gpu)
        tm_t = torch.zeros(w * h * 3 // 2, dtype=torch.uint8, device="cuda")
        tm_d = _dev_desc(torch, A.FMT_YUV420, -1, -1, -1, w, h, [tm_t, tm_t[w * h:], tm_t[w * h * 5 // 4:]], [w, w // 2, w // 2])
        rc = lib.uhdr_b200_tonemap_dev(C.byref(hdr_d), C.byref(tm_d), sp)
        assert rc == 0, T.gpu_err(gpu)
        cv_t = ts.clone()
        cv_d = _dev_desc(torch, A.FMT_YUV420, A.CG_BT709, A.CT_SRGB, A.CR_FULL, w, h, [cv_t, cv_t[w * h:], cv_t[w * h * 5 // 4:]], [w, w // 2, w // 2])
        rc = lib.uhdr_b200_convert_yuv_dev(C.byref(cv_d), A.CG_BT709, A.CG_P3, sp)
        assert rc == 0, T.gpu_err(gpu)
    st.synchronize()
    got_gm = gm_t.cpu().numpy().reshape(h, w, 3)
    assert (got_gm == want_gm).all() and T.md_equal(md, want_md)
    gi = T.gm_image(want_gm, gm_d.cg)
    want_px = checker.apply(sdr_h, gi, want_md, A.CT_LINEAR)
    assert (out_t.cpu().numpy().view(np.uint16).reshape(h, w, 4) == want_px).all()
    want_tm, _ = checker.tonemap(hdr_h)
    assert (tm_t.cpu().numpy() == want_tm).all()
    assert (cv_t.cpu().numpy() == checker.convert_yuv(sb, w, h, A.CG_BT709, A.CG_P3)).all()
