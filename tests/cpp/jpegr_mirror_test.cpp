// Exercises include/uhdr_b200_jpegr.hpp (the C++ mirror of ultrahdr::JpegR) against the plain C API:
// API-1 encode through both must give the same bytes; decodeJPEGR must return the same pixels as the
// handle API.  Exit code 0 = all good; prints the first problem otherwise.  Built and run by
// tests/test_cpp_mirror.py (compile-only without a GPU).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "uhdr_b200_jpegr.hpp"

static int fail(const char* what, const uhdr_error_info_t& st) {
  std::fprintf(stderr, "FAIL %s: code %d %s\n", what, (int)st.error_code, st.has_detail ? st.detail : "");
  return 1;
}

int main(int argc, char** argv) {
  const bool run = argc > 1 && argv[1][0] == 'r';
  const unsigned w = 640, h = 368;
  std::vector<uint16_t> p010((size_t)w * h * 3 / 2);
  std::vector<uint8_t> yuv((size_t)w * h * 3 / 2);
  for (unsigned y = 0; y < h; y++)
    for (unsigned x = 0; x < w; x++) {
      const unsigned v = (x * 3 + y * 5) % 877;
      p010[(size_t)y * w + x] = (uint16_t)((64 + v) << 6);
      yuv[(size_t)y * w + x] = (uint8_t)(16 + (v % 220));
    }
  for (size_t i = (size_t)w * h; i < p010.size(); i++) p010[i] = (uint16_t)((400 + (i * 7) % 300) << 6);
  for (size_t i = (size_t)w * h; i < yuv.size(); i++) yuv[i] = (uint8_t)(100 + (i * 11) % 60);

  uhdr_raw_image_t hdr{}, sdr{};
  hdr.fmt = UHDR_IMG_FMT_24bppYCbCrP010; hdr.cg = UHDR_CG_BT_2100; hdr.ct = UHDR_CT_HLG; hdr.range = UHDR_CR_LIMITED_RANGE;
  hdr.w = w; hdr.h = h;
  hdr.planes[UHDR_PLANE_Y] = p010.data(); hdr.planes[UHDR_PLANE_UV] = p010.data() + (size_t)w * h;
  hdr.stride[UHDR_PLANE_Y] = w; hdr.stride[UHDR_PLANE_UV] = w;
  sdr.fmt = UHDR_IMG_FMT_12bppYCbCr420; sdr.cg = UHDR_CG_BT_709; sdr.ct = UHDR_CT_SRGB; sdr.range = UHDR_CR_FULL_RANGE;
  sdr.w = w; sdr.h = h;
  sdr.planes[UHDR_PLANE_Y] = yuv.data(); sdr.planes[UHDR_PLANE_U] = yuv.data() + (size_t)w * h;
  sdr.planes[UHDR_PLANE_V] = yuv.data() + (size_t)w * h * 5 / 4;
  sdr.stride[UHDR_PLANE_Y] = w; sdr.stride[UHDR_PLANE_U] = w / 2; sdr.stride[UHDR_PLANE_V] = w / 2;

  ultrahdr_b200::JpegR jr(nullptr, 1, 95, true);
  if (!run) {  // link check only: every method must be instantiable
    (void)&ultrahdr_b200::JpegR::decodeJPEGR;
    (void)&ultrahdr_b200::JpegR::getJPEGRInfo;
    (void)&ultrahdr_b200::JpegR::generateGainMap;
    (void)&ultrahdr_b200::JpegR::applyGainMap;
    (void)&ultrahdr_b200::JpegR::toneMap;
    (void)&ultrahdr_b200::JpegR::convertYuv;
    std::puts("linked");
    return 0;
  }

  std::vector<uint8_t> out((size_t)w * h * 6 + 65536);
  uhdr_compressed_image_t dest{};
  dest.data = out.data(); dest.capacity = out.size();
  uhdr_error_info_t st = jr.encodeJPEGR(&hdr, &sdr, &dest, 95, nullptr);
  if (st.error_code != UHDR_CODEC_OK) return fail("JpegR::encodeJPEGR API-1", st);

  // the same through the C handle API with its defaults
  uhdr_codec_private_t* enc = uhdr_create_encoder();
  st = uhdr_enc_set_raw_image(enc, &hdr, UHDR_HDR_IMG);
  if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_raw_image(enc, &sdr, UHDR_SDR_IMG);
  if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_quality(enc, 95, UHDR_BASE_IMG);
  if (st.error_code == UHDR_CODEC_OK) st = uhdr_encode(enc);
  if (st.error_code != UHDR_CODEC_OK) return fail("uhdr_encode", st);
  const uhdr_compressed_image_t* ref = uhdr_get_encoded_stream(enc);
  if (ref->data_sz != dest.data_sz || std::memcmp(ref->data, dest.data, dest.data_sz) != 0) {
    std::fprintf(stderr, "FAIL mirror and C API files differ (%zu vs %zu bytes)\n", dest.data_sz, ref->data_sz);
    return 1;
  }
  uhdr_release_encoder(enc);

  int iw = 0, ih = 0, gw = 0, gh = 0;
  st = jr.getJPEGRInfo(&dest, &iw, &ih, &gw, &gh);
  if (st.error_code != UHDR_CODEC_OK) return fail("getJPEGRInfo", st);
  if (iw != (int)w || ih != (int)h || gw != (int)w || gh != (int)h) { std::fprintf(stderr, "FAIL info %d %d %d %d\n", iw, ih, gw, gh); return 1; }

  std::vector<uint8_t> px((size_t)w * h * 8), gm((size_t)w * h * 4);
  uhdr_raw_image_t pix{}, gmi{};
  pix.planes[0] = px.data(); pix.stride[0] = w;
  gmi.planes[0] = gm.data(); gmi.stride[0] = w;
  uhdr_gainmap_metadata_t md{};
  st = jr.decodeJPEGR(&dest, &pix, FLT_MAX, UHDR_CT_LINEAR, UHDR_IMG_FMT_64bppRGBAHalfFloat, &gmi, &md);
  if (st.error_code != UHDR_CODEC_OK) return fail("decodeJPEGR", st);
  if (pix.w != w || pix.h != h || pix.fmt != UHDR_IMG_FMT_64bppRGBAHalfFloat || gmi.w != w) { std::fprintf(stderr, "FAIL decode geometry\n"); return 1; }

  // stage-level members
  std::vector<uint8_t> map((size_t)w * h * 3);
  uhdr_raw_image_t mapimg{};
  mapimg.planes[0] = map.data();
  uhdr_gainmap_metadata_t md2{};
  st = jr.generateGainMap(&sdr, &hdr, &md2, &mapimg);
  if (st.error_code != UHDR_CODEC_OK) return fail("generateGainMap", st);
  if (std::memcmp(md.max_content_boost, md2.max_content_boost, sizeof md.max_content_boost) != 0) {
    std::fprintf(stderr, "FAIL metadata of generateGainMap differs from the decoded file's\n");
    return 1;
  }
  std::vector<uint8_t> px2((size_t)w * h * 8);
  uhdr_raw_image_t pix2{};
  pix2.fmt = UHDR_IMG_FMT_64bppRGBAHalfFloat; pix2.w = w; pix2.h = h; pix2.planes[0] = px2.data(); pix2.stride[0] = w;
  uhdr_raw_image_t sdr601 = sdr;   // applyGainMap reads the base image as decoded (BT.601 YCbCr); here: plumbing only
  st = jr.applyGainMap(&sdr601, &mapimg, &md2, UHDR_CT_LINEAR, UHDR_IMG_FMT_64bppRGBAHalfFloat, FLT_MAX, &pix2);
  if (st.error_code != UHDR_CODEC_OK) return fail("applyGainMap", st);
  std::puts("ok");
  return 0;
}
