"""Pins oracle/jpeg_oracle.c (the restated libjpeg-turbo baseline algorithm) against a real
libjpeg-turbo: Pillow's bundled 3.1.x.  Byte-identical streams, FDCT/IDCT per block through the
exported jpeg_fdct_islow / jpeg_idct_islow symbols, decoded planes equal to Pillow's."""
import ctypes as C
import glob
import io
import os

import numpy as np
import pytest

import uhdr_testlib as T
from libultrahdr_b200 import ctypes_api as A

PIL = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def olib(oracle_libs):
    return oracle_libs.Oracle().lib


def _pillow_jpeg():
    import PIL
    c = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libjpeg-*.so*"))
    return C.CDLL(c[0]) if c else None


def test_fdct_matches_libjpeg_turbo(olib):
    pj = _pillow_jpeg()
    if pj is None or not hasattr(pj, "jpeg_fdct_islow"):
        pytest.skip("Pillow's libjpeg does not export jpeg_fdct_islow")
    rs = np.random.RandomState(1)
    for t in range(3000):
        blk = (rs.randint(0, 256, 64) - 128).astype(np.int16) if t % 3 else (rs.randint(0, 2, 64) * 255 - 128).astype(np.int16)
        a, b = blk.copy(), blk.copy()
        olib.jo_fdct_islow(a.ctypes.data_as(C.c_void_p))
        pj.jpeg_fdct_islow(b.ctypes.data_as(C.c_void_p))
        assert (a == b).all()


@pytest.mark.parametrize("w,h", [(64, 64), (320, 240), (318, 237), (8, 8), (17, 9), (960, 540)])
def test_streams_byte_identical_to_pillow(olib, w, h):
    rs = np.random.RandomState(w * 1000 + h)
    for q in (1, 10, 50, 75, 90, 95, 100):
        g = rs.randint(0, 256, (h, w)).astype(np.uint8) if q % 2 == 0 else \
            (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8)
        b = io.BytesIO()
        PIL.fromarray(g).save(b, "JPEG", quality=q)
        img = A.raw_image(A.FMT_Y400, -1, -1, 1, w, h, [g], [w])
        mine = T.oracle_encode(olib, img, q)
        if w % 8 == 0 and h % 8 == 0:  # the raw-data path pads with the helper's rules, not libjpeg's
            assert mine == b.getvalue(), ("gray", w, h, q)
        rgb = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        if q % 2:
            rgb[..., 1] = g
        b = io.BytesIO()
        PIL.fromarray(rgb).save(b, "JPEG", quality=q, subsampling=0)
        img = A.raw_image(A.FMT_RGB888, -1, -1, 1, w, h, [rgb], [w])
        assert T.oracle_encode(olib, img, q) == b.getvalue(), ("rgb", w, h, q)


@pytest.mark.parametrize("w,h", [(64, 64), (318, 237), (17, 9), (480, 270)])
def test_decode_matches_pillow(olib, w, h):
    rs = np.random.RandomState(5)
    FIX = lambda x: int(x * 65536 + 0.5)  # noqa: E731
    for q in (5, 50, 95, 100):
        g = rs.randint(0, 256, (h, w)).astype(np.uint8)
        b = io.BytesIO()
        PIL.fromarray(g).save(b, "JPEG", quality=q)
        ref = np.array(PIL.open(io.BytesIO(b.getvalue())))
        hd, pl = T.oracle_decode(olib, b.getvalue())
        assert (pl[0][:h, :w] == ref).all()
        rgb = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        for ss in (0, 2, 1):
            b = io.BytesIO()
            PIL.fromarray(rgb).save(b, "JPEG", quality=q, subsampling=ss)
            im = PIL.open(io.BytesIO(b.getvalue()))
            hd, pl = T.oracle_decode(olib, b.getvalue())
            if ss == 0:
                ref = np.array(im)
                y, cb, cr = [p[:h, :w].astype(np.int32) for p in pl]
                xr, xb = cr - 128, cb - 128
                R = np.clip(y + ((FIX(1.402) * xr + 32768) >> 16), 0, 255)
                B = np.clip(y + ((FIX(1.772) * xb + 32768) >> 16), 0, 255)
                G = np.clip(y + ((-FIX(0.34414) * xb + 32768 - FIX(0.71414) * xr) >> 16), 0, 255)
                assert (np.stack([R, G, B], -1).astype(np.uint8) == ref).all()
            else:  # luma plane is the raw IDCT output in YCbCr draft mode
                im.draft("YCbCr", (w, h))
                assert (pl[0][:h, :w] == np.array(im)[..., 0]).all()


def test_420_raw_path_roundtrip_through_pillow(olib):
    """the raw_data_in 4:2:0 path cannot be produced by Pillow; check that a real libjpeg-turbo
    decodes our stream to exactly the planes our own decoder reconstructs."""
    w, h = 320, 240
    buf = T.make_yuv420(w, h, "smooth")
    img, keep = A.yuv420_image(buf, w, h, 1)
    data = T.oracle_encode(olib, img, 90)
    hd, pl = T.oracle_decode(olib, data)
    im = PIL.open(io.BytesIO(data))
    im.draft("YCbCr", (w, h))
    assert (np.array(im)[..., 0] == pl[0][:h, :w]).all()
    # q90 of a smooth ramp stays close to the source
    assert np.abs(pl[0][:h, :w].astype(int) - buf[:w * h].reshape(h, w)).max() <= 12


@pytest.mark.parametrize("w,h", [(64, 48), (318, 237), (17, 9), (2, 2), (1, 7), (640, 361), (33, 64)])
@pytest.mark.parametrize("subsampling", [2, 1, 0])  # Pillow: 2 = 4:2:0, 1 = 4:2:2, 0 = 4:4:4
def test_rgb_output_of_subsampled_streams_matches_pillow(olib, w, h, subsampling):
    """jo_planes_to_rgba (fancy upsampling + colour conversion as libjpeg-turbo's default decode
    path does them) against Pillow decoding the same stream to RGB."""
    rs = np.random.RandomState(w * 131 + h * 7 + subsampling)
    for kind in ("noise", "smooth"):
        rgb = rs.randint(0, 256, (h, w, 3)).astype(np.uint8) if kind == "noise" else \
            np.stack([(np.add.outer(np.arange(h) * k, np.arange(w) * (5 - k))) % 256 for k in (1, 2, 3)], -1).astype(np.uint8)
        b = io.BytesIO()
        PIL.fromarray(rgb).save(b, "JPEG", quality=92, subsampling=subsampling)
        data = b.getvalue()
        want = np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))
        hd, planes = T.oracle_decode(olib, data)
        out = np.zeros((h, w, 4), np.uint8)
        pp = (C.c_void_p * 3)(*[p.ctypes.data for p in planes])
        assert olib.jo_planes_to_rgba(C.byref(hd), pp, out.ctypes.data_as(C.c_void_p)) == 0
        assert (out[..., 3] == 255).all()
        assert (out[..., :3] == want).all(), (w, h, subsampling, kind, int((out[..., :3] != want).sum()))
