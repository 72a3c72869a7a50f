/*
 * uhdr_b200_jpegr.hpp -- header-only C++ mirror of the reference's ultrahdr::JpegR surface for the hot
 * path (lib/include/ultrahdr/jpegr.h:52-222, inherited UltraHdr members
 * lib/include/ultrahdr/ultrahdrcommon.h:471-546), implemented over the C ABI of libuhdr_b200.so.
 *
 * Same method names, argument meaning and error convention (uhdr_error_info_t by value) as the
 * reference, so code written against ultrahdr::JpegR ports by changing the namespace:
 *
 *   ultrahdr_b200::JpegR jr(nullptr, 1, 95, true);                   // jpegr.h:54-60
 *   jr.encodeJPEGR(&hdr, &sdr, &dest, 95, nullptr);                  // API-1, jpegr.h:101-102
 *   jr.decodeJPEGR(&file, &pixels, FLT_MAX, UHDR_CT_LINEAR, UHDR_IMG_FMT_64bppRGBAHalfFloat);
 *
 * Differences, all forced by the header being a thin shim: generateGainMap fills a caller-provided
 * descriptor instead of allocating a unique_ptr<uhdr_raw_image_ext_t> (jpegr.cpp:714); encode APIs
 * 2-4 (re-muxing of already compressed inputs) are not part of the B200 hot path.
 */
#ifndef UHDR_B200_JPEGR_HPP
#define UHDR_B200_JPEGR_HPP

#include <cfloat>
#include <cstdio>
#include <cstring>

#include "uhdr_b200.h"
#include "ultrahdr_api.h"

namespace ultrahdr_b200 {

class JpegR {
 public:
  // jpegr.h:54-60 (same defaults as the reference's library build: scale 1 / quality 95 /
  // multichannel / gamma 1 / BEST_QUALITY are the ultrahdr_api.h defaults, ultrahdrcommon.h:424-447)
  explicit JpegR(void* /*uhdrGLESCtxt*/ = nullptr, int mapDimensionScaleFactor = 1, int mapCompressQuality = 95,
                 bool useMultiChannelGainMap = true, float gamma = 1.0f, uhdr_enc_preset_t preset = UHDR_USAGE_BEST_QUALITY,
                 float minContentBoost = FLT_MIN, float maxContentBoost = FLT_MAX, float targetDispPeakBrightness = -1.0f) {
    cfg_.scale_factor = mapDimensionScaleFactor;
    cfg_.quality = mapCompressQuality;
    cfg_.multichannel = useMultiChannelGainMap ? 1 : 0;
    cfg_.gamma = gamma;
    cfg_.preset = preset;
    cfg_.min_content_boost = minContentBoost;
    cfg_.max_content_boost = maxContentBoost;
    cfg_.target_disp_peak_nits = targetDispPeakBrightness;
    cfg_.sdr_is_601 = 0;
    cfg_.use_luminance = 1;
  }

  /* Encode API-0, jpegr.cpp:179-244 */
  uhdr_error_info_t encodeJPEGR(uhdr_raw_image_t* hdr_intent, uhdr_compressed_image_t* dest, int quality, uhdr_mem_block_t* exif) {
    return encode(hdr_intent, nullptr, dest, quality, exif);
  }
  /* Encode API-1, jpegr.cpp:247-291 */
  uhdr_error_info_t encodeJPEGR(uhdr_raw_image_t* hdr_intent, uhdr_raw_image_t* sdr_intent, uhdr_compressed_image_t* dest,
                                int quality, uhdr_mem_block_t* exif) {
    if (!sdr_intent) return error(UHDR_CODEC_INVALID_PARAM, "received nullptr for sdr intent");
    return encode(hdr_intent, sdr_intent, dest, quality, exif);
  }

  /* jpegr.cpp:1469-1531.  dest->planes[0] (and gainmap_img->planes[0] when given) are caller memory, like
   * in the reference. */
  uhdr_error_info_t decodeJPEGR(uhdr_compressed_image_t* uhdr_compressed_img, uhdr_raw_image_t* dest,
                                float max_display_boost = FLT_MAX, uhdr_color_transfer_t output_ct = UHDR_CT_LINEAR,
                                uhdr_img_fmt_t output_format = UHDR_IMG_FMT_64bppRGBAHalfFloat,
                                uhdr_raw_image_t* gainmap_img = nullptr, uhdr_gainmap_metadata_t* gainmap_metadata = nullptr) {
    if (!uhdr_compressed_img || !uhdr_compressed_img->data || !dest || !dest->planes[0])
      return error(UHDR_CODEC_INVALID_PARAM, "received nullptr for compressed image or destination");
    uhdr_codec_private_t* dec = uhdr_create_decoder();
    if (!dec) return error(UHDR_CODEC_MEM_ERROR, "unable to allocate a decoder instance");
    uhdr_error_info_t st = uhdr_dec_set_image(dec, uhdr_compressed_img);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_dec_set_out_img_format(dec, output_format);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_dec_set_out_color_transfer(dec, output_ct);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_dec_set_out_max_display_boost(dec, max_display_boost);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_decode(dec);
    if (st.error_code == UHDR_CODEC_OK) {
      st = copy_out(uhdr_get_decoded_image(dec), dest);
      if (st.error_code == UHDR_CODEC_OK && gainmap_img) st = copy_out(uhdr_get_decoded_gainmap_image(dec), gainmap_img);
      if (st.error_code == UHDR_CODEC_OK && gainmap_metadata) {
        const uhdr_gainmap_metadata_t* md = uhdr_dec_get_gainmap_metadata(dec);
        if (md) *gainmap_metadata = *md;
      }
    }
    uhdr_release_decoder(dec);
    return st;
  }

  /* jpegr.cpp:1417-1430: sizes only (the full jr_info_ptr with exif / icc blocks needs a live handle:
   * use uhdr_dec_probe + the uhdr_dec_get_* getters for those) */
  uhdr_error_info_t getJPEGRInfo(uhdr_compressed_image_t* uhdr_compressed_img, int* width, int* height, int* gm_width = nullptr,
                                 int* gm_height = nullptr) {
    uhdr_codec_private_t* dec = uhdr_create_decoder();
    if (!dec) return error(UHDR_CODEC_MEM_ERROR, "unable to allocate a decoder instance");
    uhdr_error_info_t st = uhdr_dec_set_image(dec, uhdr_compressed_img);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_dec_probe(dec);
    if (st.error_code == UHDR_CODEC_OK) {
      if (width) *width = uhdr_dec_get_image_width(dec);
      if (height) *height = uhdr_dec_get_image_height(dec);
      if (gm_width) *gm_width = uhdr_dec_get_gainmap_width(dec);
      if (gm_height) *gm_height = uhdr_dec_get_gainmap_height(dec);
    }
    uhdr_release_decoder(dec);
    return st;
  }

  /* UltraHdr::generateGainMap, jpegr.cpp:530-1058.  gainmap_img->planes[0]: caller memory of
   * ceil(w/scale)*ceil(h/scale)*(3 or 1) bytes; fmt / w / h / stride are filled in. */
  uhdr_error_info_t generateGainMap(uhdr_raw_image_t* sdr_intent, uhdr_raw_image_t* hdr_intent, uhdr_gainmap_metadata_t* gainmap_metadata,
                                    uhdr_raw_image_t* gainmap_img, bool sdr_is_601 = false, bool use_luminance = true) {
    uhdr_b200_gm_config_t c = cfg_;
    c.sdr_is_601 = sdr_is_601 ? 1 : 0;
    c.use_luminance = use_luminance ? 1 : 0;
    return from_rc(uhdr_b200_generate_gainmap(sdr_intent, hdr_intent, &c, gainmap_metadata, gainmap_img));
  }
  /* UltraHdr::applyGainMap, jpegr.cpp:1533-1831 */
  uhdr_error_info_t applyGainMap(uhdr_raw_image_t* sdr_intent, uhdr_raw_image_t* gainmap_img, uhdr_gainmap_metadata_t* gainmap_metadata,
                                 uhdr_color_transfer_t output_ct, uhdr_img_fmt_t output_format, float max_display_boost,
                                 uhdr_raw_image_t* dest) {
    return from_rc(uhdr_b200_apply_gainmap(sdr_intent, gainmap_img, gainmap_metadata, output_ct, output_format, max_display_boost, dest));
  }
  /* UltraHdr::toneMap, jpegr.cpp:1985-2222 */
  uhdr_error_info_t toneMap(uhdr_raw_image_t* hdr_intent, uhdr_raw_image_t* sdr_intent) { return from_rc(uhdr_b200_tonemap(hdr_intent, sdr_intent)); }
  /* UltraHdr::convertYuv, jpegr.cpp:436-518 */
  uhdr_error_info_t convertYuv(uhdr_raw_image_t* image, uhdr_color_gamut_t src_encoding, uhdr_color_gamut_t dst_encoding) {
    return from_rc(uhdr_b200_convert_yuv(image, src_encoding, dst_encoding));
  }

 private:
  uhdr_b200_gm_config_t cfg_{};

  static uhdr_error_info_t error(uhdr_codec_err_t code, const char* msg) {
    uhdr_error_info_t st;
    std::memset(&st, 0, sizeof st);
    st.error_code = code;
    st.has_detail = 1;
    std::snprintf(st.detail, sizeof st.detail, "%s", msg);
    return st;
  }
  static uhdr_error_info_t from_rc(int rc) {
    if (rc == 0) {
      uhdr_error_info_t ok;
      std::memset(&ok, 0, sizeof ok);
      return ok;
    }
    return error(static_cast<uhdr_codec_err_t>(rc), uhdr_b200_last_error());
  }
  static uhdr_error_info_t copy_out(const uhdr_raw_image_t* src, uhdr_raw_image_t* dst) {
    if (!src || !src->planes[0]) return error(UHDR_CODEC_ERROR, "decoder returned no image");
    if (!dst->planes[0]) return error(UHDR_CODEC_INVALID_PARAM, "destination image has no memory");
    const size_t bpp = src->fmt == UHDR_IMG_FMT_64bppRGBAHalfFloat ? 8 : (src->fmt == UHDR_IMG_FMT_8bppYCbCr400 ? 1 : 4);
    const unsigned dstride = dst->stride[0] ? dst->stride[0] : src->w;
    dst->fmt = src->fmt; dst->cg = src->cg; dst->ct = src->ct; dst->range = src->range;
    dst->w = src->w; dst->h = src->h; dst->stride[0] = dstride;
    for (unsigned y = 0; y < src->h; y++)
      std::memcpy(static_cast<char*>(dst->planes[0]) + (size_t)y * dstride * bpp,
                  static_cast<const char*>(src->planes[0]) + (size_t)y * src->stride[0] * bpp, (size_t)src->w * bpp);
    return from_rc(0);
  }
  uhdr_error_info_t encode(uhdr_raw_image_t* hdr, uhdr_raw_image_t* sdr, uhdr_compressed_image_t* dest, int quality, uhdr_mem_block_t* exif) {
    if (!hdr || !dest || !dest->data) return error(UHDR_CODEC_INVALID_PARAM, "received nullptr for hdr intent or destination");
    uhdr_codec_private_t* enc = uhdr_create_encoder();
    if (!enc) return error(UHDR_CODEC_MEM_ERROR, "unable to allocate an encoder instance");
    uhdr_error_info_t st = uhdr_enc_set_raw_image(enc, hdr, UHDR_HDR_IMG);
    if (st.error_code == UHDR_CODEC_OK && sdr) st = uhdr_enc_set_raw_image(enc, sdr, UHDR_SDR_IMG);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_quality(enc, quality, UHDR_BASE_IMG);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_quality(enc, cfg_.quality, UHDR_GAIN_MAP_IMG);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_gainmap_scale_factor(enc, cfg_.scale_factor);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_using_multi_channel_gainmap(enc, cfg_.multichannel);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_gainmap_gamma(enc, cfg_.gamma);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_enc_set_preset(enc, static_cast<uhdr_enc_preset_t>(cfg_.preset));
    if (st.error_code == UHDR_CODEC_OK && (cfg_.min_content_boost != FLT_MIN || cfg_.max_content_boost != FLT_MAX))
      st = uhdr_enc_set_min_max_content_boost(enc, cfg_.min_content_boost, cfg_.max_content_boost);
    if (st.error_code == UHDR_CODEC_OK && cfg_.target_disp_peak_nits != -1.0f)
      st = uhdr_enc_set_target_display_peak_brightness(enc, cfg_.target_disp_peak_nits);
    if (st.error_code == UHDR_CODEC_OK && exif) st = uhdr_enc_set_exif_data(enc, exif);
    if (st.error_code == UHDR_CODEC_OK) st = uhdr_encode(enc);
    if (st.error_code == UHDR_CODEC_OK) {
      const uhdr_compressed_image_t* out = uhdr_get_encoded_stream(enc);
      if (!out || out->data_sz > dest->capacity) {
        st = error(UHDR_CODEC_MEM_ERROR, "destination buffer too small for the compressed image");
      } else {
        std::memcpy(dest->data, out->data, out->data_sz);
        dest->data_sz = out->data_sz;
        dest->cg = out->cg; dest->ct = out->ct; dest->range = out->range;
      }
    }
    uhdr_release_encoder(enc);
    return st;
  }
};

}  // namespace ultrahdr_b200

#endif
