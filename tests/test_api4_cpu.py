"""Encode API-4 (pre-compressed base image + pre-compressed gain map + metadata -> JPEG/R,
jpegr.cpp:388-434 + appendGainMap :1105-1415) is container work on the host: byte parity with the
reference without a GPU."""
import ctypes as C
import io

import numpy as np
import pytest

import uhdr_testlib as T
from libultrahdr_b200 import ctypes_api as A
from test_probe_cpu import _probe


def _api4(lib, base, gm, md, base_cg=-1, exif=None):
    lib.uhdr_create_encoder.restype = C.c_void_p
    for f in ("uhdr_enc_set_compressed_image", "uhdr_enc_set_gainmap_image", "uhdr_encode", "uhdr_enc_set_exif_data"):
        getattr(lib, f).restype = A.ErrorInfo
    lib.uhdr_get_encoded_stream.restype = C.POINTER(A.CompressedImage)
    enc = C.c_void_p(lib.uhdr_create_encoder())
    try:
        bb, gb = np.frombuffer(base, np.uint8).copy(), np.frombuffer(gm, np.uint8).copy()
        bi = A.CompressedImage(bb.ctypes.data, len(base), len(base), base_cg, -1, -1)
        gi = A.CompressedImage(gb.ctypes.data, len(gm), len(gm), -1, -1, -1)
        e = lib.uhdr_enc_set_compressed_image(enc, C.byref(bi), A.BASE_IMG)
        if e.error_code:
            return ("set_base", e.error_code)
        e = lib.uhdr_enc_set_gainmap_image(enc, C.byref(gi), C.byref(md))
        if e.error_code:
            return ("set_gm", e.error_code)
        if exif is not None:
            xb = np.frombuffer(exif, np.uint8).copy()
            blk = A.MemBlock(xb.ctypes.data, len(exif), len(exif))
            e = lib.uhdr_enc_set_exif_data(enc, C.byref(blk))
            assert e.error_code == 0
        e = lib.uhdr_encode(enc)
        if e.error_code:
            return ("encode", e.error_code)
        o = lib.uhdr_get_encoded_stream(enc).contents
        return C.string_at(o.data, o.data_sz)
    finally:
        lib.uhdr_release_encoder(enc)


@pytest.fixture(scope="module")
def libs(oracle_libs):
    if not oracle_libs.have_ref():
        pytest.skip("reference build not available")
    return C.CDLL(T.GPU_SO), oracle_libs.Ref().lib


@pytest.fixture(scope="module")
def parts(libs):
    _mine, ref = libs
    api = T.UhdrApi(ref)
    w, h = 192, 128
    hb, sb = T.make_p010(w, h, "smooth"), T.make_yuv420(w, h, "smooth")
    hdr, k1 = A.p010_image(hb, w, h, A.CG_BT2100, A.CT_HLG, A.CR_LIMITED)
    sdr, k2 = A.yuv420_image(sb, w, h, A.CG_BT709)
    out = {}
    for name, opts in (("multi", {}), ("single", {"multichannel": 0, "scale": 2})):
        data = api.encode(hdr, sdr, **opts)
        p = _probe(ref, data)
        out[name] = (p["base_image"], p["gainmap_image"], p["md"])
    return out


@pytest.mark.parametrize("which", ["multi", "single"])
def test_api4_matches_reference(libs, parts, which):
    mine, ref = libs
    base, gm, md = parts[which]
    a, b = _api4(mine, base, gm, md), _api4(ref, base, gm, md)
    assert isinstance(b, bytes), b
    assert a == b
    # the result is a JPEG/R both libraries probe identically
    pa, pb = _probe(mine, a), _probe(ref, a)
    assert pa["dims"] == pb["dims"] and T.md_equal(pa["md"], pb["md"])


def test_api4_base_without_icc_and_with_exif(libs, parts):
    """a base image from another encoder (Pillow's libjpeg-turbo): no ICC -> one is written from the
    configured gamut; EXIF inside the base image is carried over; unknown gamut is an error"""
    PIL = pytest.importorskip("PIL.Image")
    mine, ref = libs
    _b, gm, md = parts["single"]
    rgb = (np.add.outer(np.arange(128), np.arange(192)) % 256).astype(np.uint8)
    img = PIL.fromarray(np.stack([rgb, rgb[::-1], rgb.T[:128, :192] if False else rgb], -1))
    plain = io.BytesIO()
    img.save(plain, "JPEG", quality=88)
    exif = b"Exif\x00\x00MM\x00\x2a\x00\x00\x00\x08\x00\x00\x00\x00\x00\x00"
    with_exif = io.BytesIO()
    img.save(with_exif, "JPEG", quality=88, exif=exif)
    for base in (plain.getvalue(), with_exif.getvalue()):
        for cg in (A.CG_BT709, A.CG_P3, A.CG_BT2100):
            a, b = _api4(mine, base, gm, md, cg), _api4(ref, base, gm, md, cg)
            assert isinstance(b, bytes), b
            assert a == b, cg
        a, b = _api4(mine, base, gm, md, -1), _api4(ref, base, gm, md, -1)
        assert a == b and not isinstance(a, bytes)   # same error code from both
    # exif given twice: through the API and inside the base image
    a, b = _api4(mine, with_exif.getvalue(), gm, md, A.CG_BT709, exif=exif), _api4(ref, with_exif.getvalue(), gm, md, A.CG_BT709, exif=exif)
    assert a == b
    a, b = _api4(mine, plain.getvalue(), gm, md, A.CG_BT709, exif=exif), _api4(ref, plain.getvalue(), gm, md, A.CG_BT709, exif=exif)
    assert a == b


def test_api4_rejects_what_the_reference_rejects(libs, parts):
    mine, ref = libs
    base, gm, md = parts["multi"]
    bad = A.GainmapMetadata.from_buffer_copy(bytes(md))
    bad.gamma[1] = -1.0
    assert _api4(mine, base, gm, bad) == _api4(ref, base, gm, bad)
    assert _api4(mine, base[:200], gm, md) == _api4(ref, base[:200], gm, md)
    assert _api4(mine, b"notajpeg" * 10, gm, md) == _api4(ref, b"notajpeg" * 10, gm, md)
    # gain map applied in the alternate image space needs an ICC profile in the gain-map image
    alt = A.GainmapMetadata.from_buffer_copy(bytes(md))
    alt.use_base_cg = 0
    sos = gm.index(b"\xff\xe2")
    seglen = (gm[sos + 2] << 8) | gm[sos + 3]
    if b"ICC_PROFILE" in gm[sos:sos + 20]:
        stripped = gm[:sos] + gm[sos + 2 + seglen:]
        assert _api4(mine, base, stripped, alt) == _api4(ref, base, stripped, alt)
    assert _api4(mine, base, gm, alt) == _api4(ref, base, gm, alt)
