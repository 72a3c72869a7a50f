"""Shared helpers for the test-suite: library loading, seeded synthetic frames and thin ctypes
wrappers.  Three implementations expose the same four stage calls:

  * ``Ref``    -- oracle/_ref/libuhdr_ref.so : the UNMODIFIED reference sources compiled in place
  * ``Oracle`` -- oracle/liboracle.so        : the plain-C restatement (the "port")
  * ``Gpu``    -- libultrahdr_b200/libuhdr_b200.so : the product (CUDA), host-buffer C ABI

so a parity test reads ``assert_same(Gpu().apply(...), Ref().apply(...))``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from libultrahdr_b200.ctypes_api import *  # noqa: F401,F403
from libultrahdr_b200 import ctypes_api as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libuhdr_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
GPU_SO = os.path.join(ROOT, "libultrahdr_b200", "libuhdr_b200.so")
REF_DATA = "/root/reference/tests/data"
SEED = 20240607


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference/lib/src")
                                          and not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"],
                              stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_SO)


# ------------------------------------------------------------------------------------------------
# synthetic frames: noise / smooth / edge, seeded
# ------------------------------------------------------------------------------------------------
def make_p010(w, h, kind="noise", seed=SEED, limited=True):
    rs = np.random.RandomState(seed)
    n = w * h
    if kind == "noise":
        if limited:
            y = rs.randint(64, 941, n)
            uv = rs.randint(64, 961, n // 2)
        else:
            y = rs.randint(0, 1024, n)
            uv = rs.randint(0, 1024, n // 2)
    elif kind == "smooth":
        yy, xx = np.mgrid[0:h, 0:w]
        y = (64 + 876 * (0.5 + 0.5 * np.sin(xx / 97.0) * np.cos(yy / 61.0)) *
             (xx + yy) / (w + h)).astype(np.int64).ravel()
        cy, cx = np.mgrid[0:h // 2, 0:w // 2]
        u = 512 + 200 * np.sin(cx / 53.0)
        v = 512 + 200 * np.cos(cy / 41.0)
        uv = np.stack([u, v], -1).astype(np.int64).ravel()
    elif kind == "black":
        y = np.full(n, 64)
        uv = np.full(n // 2, 512)
    elif kind == "white":
        y = np.full(n, 940)
        uv = np.full(n // 2, 512)
    else:
        raise ValueError(kind)
    buf = np.concatenate([y, uv]).astype(np.uint16) << 6
    return np.ascontiguousarray(buf)


def make_yuv420(w, h, kind="noise", seed=SEED + 1):
    rs = np.random.RandomState(seed)
    n = w * h + 2 * (w // 2) * (h // 2)
    if kind == "noise":
        return rs.randint(0, 256, n).astype(np.uint8)
    if kind == "smooth":
        yy, xx = np.mgrid[0:h, 0:w]
        y = (255 * (0.5 + 0.5 * np.sin(xx / 97.0) * np.cos(yy / 61.0)) * (xx + yy) / (w + h))
        cy, cx = np.mgrid[0:h // 2, 0:w // 2]
        u = 128 + 50 * np.sin(cx / 53.0)
        v = 128 + 50 * np.cos(cy / 41.0)
        return np.concatenate([y.ravel(), u.ravel(), v.ravel()]).astype(np.uint8)
    if kind == "black":
        return np.concatenate([np.zeros(w * h), np.full(n - w * h, 128)]).astype(np.uint8)
    if kind == "white":
        return np.concatenate([np.full(w * h, 255), np.full(n - w * h, 128)]).astype(np.uint8)
    raise ValueError(kind)


def make_rgba1010102(w, h, seed=SEED + 2):
    rs = np.random.RandomState(seed)
    return (rs.randint(0, 1 << 30, w * h).astype(np.uint32) | np.uint32(3 << 30))


def make_rgbaf16(w, h, seed=SEED + 3):
    rs = np.random.RandomState(seed)
    px = np.ones((h * w, 4), np.float16)
    px[:, :3] = (rs.rand(h * w, 3) ** 3 * 20.0).astype(np.float16)
    # sprinkle non-finite / negative values (sanitizePixel, gainmapmath.h:588-593)
    idx = rs.randint(0, h * w, 16)
    px[idx[:4], 0] = np.inf
    px[idx[4:8], 1] = -np.inf
    px[idx[8:12], 2] = np.nan
    px[idx[12:], 0] = -1.0
    return px.view(np.uint16).reshape(-1).copy()


def make_rgba8888(w, h, seed=SEED + 4):
    rs = np.random.RandomState(seed)
    return (rs.randint(0, 1 << 24, w * h).astype(np.uint32) | np.uint32(0xFF000000))


def load_fixture_720p():
    """config 1 inputs; only available in the build container (not on the GPU box)."""
    p = np.fromfile(os.path.join(REF_DATA, "raw_p010_image.p010"), dtype=np.uint16)
    y = np.fromfile(os.path.join(REF_DATA, "raw_yuv420_image.yuv420"), dtype=np.uint8)
    return p, y


# ------------------------------------------------------------------------------------------------
class _Impl:
    """Common calling convention over a library exporting <pfx>generate_gainmap etc."""

    def __init__(self, so, pfx, mode=None):
        self.lib = C.CDLL(so) if mode is None else C.CDLL(so, mode=mode)
        self.pfx = pfx

    def f(self, name):
        return getattr(self.lib, self.pfx + name)

    def generate(self, sdr, hdr, cfg=None):
        """-> (gainmap ndarray (h,w,c) u8, GainmapMetadata)"""
        cfg = cfg or A.default_gm_config()
        s = max(1, cfg.scale_factor)
        mw, mh = sdr.w // s, sdr.h // s
        ch = 3 if cfg.multichannel else 1
        gm = np.zeros((mh, mw, ch), np.uint8)
        gmi = A.raw_image(A.FMT_RGB888 if ch == 3 else A.FMT_Y400, -1, -1, -1, mw, mh, [gm], [mw])
        md = A.GainmapMetadata()
        rc = self.f("generate_gainmap")(C.byref(sdr), C.byref(hdr), C.byref(cfg), C.byref(md),
                                        C.byref(gmi))
        assert rc == 0, f"{self.pfx}generate_gainmap rc={rc}"
        self.last_gm_desc = gmi
        return gm, md

    def apply(self, sdr, gm_img, md, out_ct, max_boost=A.FLT_MAX):
        w, h = sdr.w, sdr.h
        if out_ct == A.CT_LINEAR:
            out = np.zeros((h, w, 4), np.uint16)
            fmt = A.FMT_RGBAF16
        else:
            out = np.zeros((h, w), np.uint32)
            fmt = A.FMT_RGBA1010102
        dst = A.raw_image(fmt, -1, out_ct, A.CR_FULL, w, h, [out], [w])
        fn = self.f("apply_gainmap")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        rc = fn(C.byref(sdr), C.byref(gm_img), C.byref(md), out_ct, fmt, max_boost, C.byref(dst))
        assert rc == 0, f"{self.pfx}apply_gainmap rc={rc}"
        return out

    def tonemap(self, hdr):
        w, h = hdr.w, hdr.h
        if hdr.fmt == A.FMT_P010:
            out = np.zeros(w * h * 3 // 2, np.uint8)
            sdr, _ = A.yuv420_image(out, w, h, -1, -1, -1)
        else:
            out = np.zeros(w * h, np.uint32)
            sdr = A.raw_image(A.FMT_RGBA8888, -1, -1, -1, w, h, [out], [w])
        rc = self.f("tonemap")(C.byref(hdr), C.byref(sdr))
        assert rc == 0, f"{self.pfx}tonemap rc={rc}"
        return out, sdr

    def convert_yuv(self, buf, w, h, src_cg, dst_cg):
        b = buf.copy()
        img, _ = A.yuv420_image(b, w, h, src_cg)
        rc = self.f("convert_yuv")(C.byref(img), src_cg, dst_cg)
        assert rc == 0
        return b

    def lut(self, which):
        n = [1024, 4096, 4096, 65536, 65536][which]
        out = np.zeros(n, np.float32)
        assert self.f("lut")(which, out.ctypes.data_as(C.c_void_p), n) == 0
        return out


class Ref(_Impl):
    def __init__(self):
        super().__init__(REF_SO, "ref_")
        for n in ("srgb_oetf", "compute_gain", "hlg_ootf_1", "hlg_inv_ootf_1"):
            self.f(n).restype = C.c_float


class Oracle(_Impl):
    def __init__(self):
        super().__init__(ORACLE_SO, "uo_")
        for n in ("srgb_oetf", "compute_gain"):
            self.f(n).restype = C.c_float


class Gpu(_Impl):
    def __init__(self):
        if not os.path.exists(GPU_SO):
            raise RuntimeError("libuhdr_b200.so missing: run `python -c 'import __graft_entry__ as g;"
                               " g.build()'`")
        super().__init__(GPU_SO, "uhdr_b200_")


def gm_image(gm, cg=-1, ct=-1, rng=-1):
    """wrap a (h,w,c) u8 gain map as a raw image descriptor (c = 1, 3 or 4)."""
    h, w, c = gm.shape
    fmt = {1: A.FMT_Y400, 3: A.FMT_RGB888, 4: A.FMT_RGBA8888}[c]
    return A.raw_image(fmt, cg, ct, rng, w, h, [gm], [w])


def md_equal(a, b):
    return bytes(a) == bytes(b)
