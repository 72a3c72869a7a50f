/*
 * TEST INFRASTRUCTURE ONLY -- part of oracle/_ref.
 *
 * extern "C" wrappers so tests / bench.py (ctypes) can call the UNMODIFIED reference
 * implementation (compiled from /root/reference/lib/src/{jpegr,gainmapmath,...}.cpp) stage by
 * stage.  Everything here forwards to ultrahdr::JpegR; nothing is re-implemented.
 */
#include <cstring>
#include <memory>

#include "ultrahdr_api.h"
#include "ultrahdr/ultrahdrcommon.h"
#include "ultrahdr/jpegr.h"
#include "ultrahdr/gainmapmath.h"
#include "ultrahdr/jpegencoderhelper.h"
#include "ultrahdr/jpegdecoderhelper.h"

using namespace ultrahdr;

#define REF_API extern "C" __attribute__((visibility("default")))

struct ref_gm_config {
  int scale_factor;
  int quality;          /* gain-map jpeg quality (unused by generate) */
  int multichannel;
  float gamma;
  int preset;           /* uhdr_enc_preset_t */
  float min_content_boost, max_content_boost; /* FLT_MIN / FLT_MAX = unset */
  float target_disp_peak_nits;                /* -1 = unset */
  int sdr_is_601;
  int use_luminance;
};

static JpegR make(const ref_gm_config* c) {
  return JpegR(nullptr, c->scale_factor, c->quality, c->multichannel != 0, c->gamma,
               (uhdr_enc_preset_t)c->preset, c->min_content_boost, c->max_content_boost,
               c->target_disp_peak_nits);
}

/* generateGainMap (lib/src/jpegr.cpp:530). gainmap_out->planes[0] must hold map_w*map_h*(3|1)
 * bytes; it is written tightly packed (stride = width). */
REF_API int ref_generate_gainmap(uhdr_raw_image_t* sdr, uhdr_raw_image_t* hdr,
                                 const ref_gm_config* cfg, uhdr_gainmap_metadata_t* md_out,
                                 uhdr_raw_image_t* gainmap_out) {
  JpegR j = make(cfg);
  uhdr_gainmap_metadata_ext_t md(kJpegrVersion);
  std::unique_ptr<uhdr_raw_image_ext_t> gm;
  uhdr_error_info_t st =
      j.generateGainMap(sdr, hdr, &md, gm, cfg->sdr_is_601 != 0, cfg->use_luminance != 0);
  if (st.error_code != UHDR_CODEC_OK) return (int)st.error_code;
  *md_out = md;
  const int bpp = gm->fmt == UHDR_IMG_FMT_24bppRGB888 ? 3 : 1;
  uint8_t* dst = static_cast<uint8_t*>(gainmap_out->planes[0]);
  for (unsigned y = 0; y < gm->h; y++)
    memcpy(dst + (size_t)y * gm->w * bpp,
           static_cast<uint8_t*>(gm->planes[0]) + (size_t)y * gm->stride[0] * bpp,
           (size_t)gm->w * bpp);
  gainmap_out->fmt = gm->fmt;
  gainmap_out->cg = gm->cg;
  gainmap_out->ct = gm->ct;
  gainmap_out->range = gm->range;
  gainmap_out->w = gm->w;
  gainmap_out->h = gm->h;
  gainmap_out->stride[0] = gm->w;
  return 0;
}

/* applyGainMap (lib/src/jpegr.cpp:1533) */
REF_API int ref_apply_gainmap(uhdr_raw_image_t* sdr, uhdr_raw_image_t* gainmap,
                              uhdr_gainmap_metadata_t* md, int output_ct, int output_fmt,
                              float max_display_boost, uhdr_raw_image_t* dest) {
  JpegR j;
  uhdr_gainmap_metadata_ext_t m(*md, kJpegrVersion);
  uhdr_error_info_t st = j.applyGainMap(sdr, gainmap, &m, (uhdr_color_transfer_t)output_ct,
                                        (uhdr_img_fmt_t)output_fmt, max_display_boost, dest);
  return (int)st.error_code;
}

/* toneMap (lib/src/jpegr.cpp:1985) */
REF_API int ref_tonemap(uhdr_raw_image_t* hdr, uhdr_raw_image_t* sdr) {
  JpegR j;
  return (int)j.toneMap(hdr, sdr).error_code;
}

/* convertYuv (lib/src/jpegr.cpp:436) */
REF_API int ref_convert_yuv(uhdr_raw_image_t* img, int src_cg, int dst_cg) {
  JpegR j;
  return (int)j.convertYuv(img, (uhdr_color_gamut_t)src_cg, (uhdr_color_gamut_t)dst_cg).error_code;
}

/* LUTs exactly as the reference builds them (gainmapmath.cpp:126-349): evaluate the LUT
 * accessor at every node. which: 0 srgbInvOetf(1024) 1 hlgInvOetf(4096) 2 pqInvOetf(4096)
 * 3 hlgOetf(65536) 4 pqOetf(65536) */
REF_API int ref_lut(int which, float* out, int n) {
  static const int sizes[5] = {kSrgbInvOETFNumEntries, kHlgInvOETFNumEntries,
                               kPqInvOETFNumEntries, kHlgOETFNumEntries, kPqOETFNumEntries};
  if (which < 0 || which > 4 || n != sizes[which]) return -1;
  for (int i = 0; i < n; i++) {
    float x = static_cast<float>(i) / static_cast<float>(n - 1);
    switch (which) {
      case 0: out[i] = srgbInvOetfLUT(x); break;
      case 1: out[i] = hlgInvOetfLUT(x); break;
      case 2: out[i] = pqInvOetfLUT(x); break;
      case 3: out[i] = hlgOetfLUT(x); break;
      case 4: out[i] = pqOetfLUT(x); break;
    }
  }
  return 0;
}

/* scalar primitives for known-answer tests (gainmapmath.cpp) */
REF_API float ref_srgb_oetf(float x) { return srgbOetf(x); }
REF_API float ref_hlg_ootf_1(float x) { return hlgOotfApprox({{{x, x, x}}}, nullptr).r; }
REF_API float ref_hlg_inv_ootf_1(float x) { return hlgInverseOotfApprox({{{x, x, x}}}).r; }
REF_API float ref_compute_gain(float sdr, float hdr) { return computeGain(sdr, hdr); }
REF_API int ref_affine_map_gain(float g, float mn, float mx, float gamma) {
  return affineMapGain(g, mn, mx, gamma);
}
REF_API int ref_encode_gain(float y_sdr, float y_hdr, uhdr_gainmap_metadata_t* md, float l2min,
                            float l2max, int idx) {
  uhdr_gainmap_metadata_ext_t m(*md, kJpegrVersion);
  return encodeGain(y_sdr, y_hdr, &m, l2min, l2max, idx);
}
REF_API unsigned ref_float_to_half(float f) { return floatToHalf(f); }
REF_API void ref_idw_weights(int scale, int variant, float* out) {
  ShepardsIDW t(scale);
  float* src = variant == 0 ? t.mWeights : variant == 1 ? t.mWeightsNR
               : variant == 2 ? t.mWeightsNB : t.mWeightsC;
  memcpy(out, src, sizeof(float) * scale * scale * 4);
}
REF_API void ref_gain_lut(uhdr_gainmap_metadata_t* md, float weight, float* out /*3*1024*/) {
  uhdr_gainmap_metadata_ext_t m(*md, kJpegrVersion);
  GainLUT lut(&m, weight);
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < kGainFactorNumEntries; i++) {
      /* gamma == 1 assumed by callers of this accessor */
      out[c * kGainFactorNumEntries + i] =
          lut.getGainFactor(static_cast<float>(i) / (kGainFactorNumEntries - 1), c);
    }
}

/* JPEG helpers as the reference calls them (backed by oracle/jpeg_oracle.c in this build) */
REF_API int ref_jpeg_encode(uhdr_raw_image_t* img, int quality, const void* icc, size_t icc_size,
                            uint8_t* out, size_t cap, size_t* out_size) {
  JpegEncoderHelper e;
  uhdr_error_info_t st = e.compressImage(img, quality, icc, icc_size);
  if (st.error_code != UHDR_CODEC_OK) return (int)st.error_code;
  *out_size = e.getCompressedImageSize();
  if (*out_size > cap) return -1;
  memcpy(out, e.getCompressedImagePtr(), *out_size);
  return 0;
}

/* ICC profile as the reference writes it (lib/src/icc.cpp:404); data includes the
 * "ICC_PROFILE\0" + chunk bytes prefix.  Used by tools/make_icc_blobs.py */
#include "ultrahdr/icc.h"
REF_API int ref_icc_profile(int ct, int cg, uint8_t* out, size_t cap) {
  std::shared_ptr<DataStruct> icc = IccHelper::writeIccProfile((uhdr_color_transfer_t)ct, (uhdr_color_gamut_t)cg);
  if (!icc) return -1;
  if ((size_t)icc->getLength() > cap) return -(int)icc->getLength();
  memcpy(out, icc->getData(), icc->getLength());
  return icc->getLength();
}
REF_API int ref_icc_gamut(void* data, size_t n) { return IccHelper::readIccColorGamut(data, n); }
