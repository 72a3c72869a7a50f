/*
 * TEST INFRASTRUCTURE ONLY -- part of oracle/_ref.
 *
 * Bodies for the reference's JpegEncoderHelper / JpegDecoderHelper classes (declared in
 * /root/reference/lib/include/ultrahdr/jpeg{en,de}coderhelper.h) implemented on top of
 * oracle/jpeg_oracle.c, because libjpeg-turbo headers are not available here.  Behaviour follows
 * lib/src/jpegencoderhelper.cpp:101-244 and lib/src/jpegdecoderhelper.cpp:169-555: same marker
 * order, same quality/table/sampling choices, same plane/stride bookkeeping, same validation.
 */
#include <cmath>
#include <cstring>

#include "ultrahdr/ultrahdrcommon.h"
#include "ultrahdr/jpegencoderhelper.h"
#include "ultrahdr/jpegdecoderhelper.h"
#include "../jpeg_oracle.h"

namespace ultrahdr {

const int kMinWidth = 8;
const int kMinHeight = 8;
const int kMaxWidth = 8192;
const int kMaxHeight = 8192;

static uhdr_error_info_t err(uhdr_codec_err_t code, const char* msg) {
  uhdr_error_info_t s;
  s.error_code = code;
  s.has_detail = 1;
  snprintf(s.detail, sizeof s.detail, "%s", msg);
  return s;
}

uhdr_error_info_t JpegEncoderHelper::compressImage(const uhdr_raw_image_t* img, const int qfactor,
                                                   const void* iccBuffer, const size_t iccSize) {
  const uint8_t* planes[3]{reinterpret_cast<uint8_t*>(img->planes[UHDR_PLANE_Y]),
                           reinterpret_cast<uint8_t*>(img->planes[UHDR_PLANE_U]),
                           reinterpret_cast<uint8_t*>(img->planes[UHDR_PLANE_V])};
  const unsigned int strides[3]{img->stride[UHDR_PLANE_Y], img->stride[UHDR_PLANE_U],
                                img->stride[UHDR_PLANE_V]};
  return compressImage(planes, strides, img->w, img->h, img->fmt, qfactor, iccBuffer, iccSize);
}

uhdr_error_info_t JpegEncoderHelper::compressImage(const uint8_t* planes[3],
                                                   const unsigned int strides[3], const int width,
                                                   const int height, const uhdr_img_fmt_t format,
                                                   const int qfactor, const void* iccBuffer,
                                                   const size_t iccSize) {
  if (format != UHDR_IMG_FMT_24bppRGB888 && format != UHDR_IMG_FMT_8bppYCbCr400 &&
      format != UHDR_IMG_FMT_12bppYCbCr420 && format != UHDR_IMG_FMT_24bppYCbCr444 &&
      format != UHDR_IMG_FMT_16bppYCbCr422)
    return err(UHDR_CODEC_INVALID_PARAM, "oracle shim: unsupported input format");
  const bool isGainMapImg =
      format == UHDR_IMG_FMT_24bppRGB888 || format == UHDR_IMG_FMT_8bppYCbCr400;
  char comment[255];
  snprintf(comment, sizeof comment,
           "Source: google libuhdr v%s, Coder: libjpeg v%d, Attrib: GainMap Image",
           UHDR_LIB_VERSION_STR, JPEG_LIB_VERSION);
  uint8_t* out = nullptr;
  size_t n = 0;
  int rc = jo_encode(planes, strides, width, height, (int)format, qfactor,
                     static_cast<const uint8_t*>(iccBuffer), iccSize,
                     isGainMapImg ? comment : nullptr, &out, &n);
  if (rc != 0) return err(UHDR_CODEC_ERROR, "oracle jpeg encode failed");
  mDestMgr.mResultBuffer.assign(out, out + n);
  free(out);
  return g_no_error;
}

uhdr_compressed_image_t JpegEncoderHelper::getCompressedImage() {
  uhdr_compressed_image_t img;
  img.data = mDestMgr.mResultBuffer.data();
  img.capacity = img.data_sz = mDestMgr.mResultBuffer.size();
  img.cg = UHDR_CG_UNSPECIFIED;
  img.ct = UHDR_CT_UNSPECIFIED;
  img.range = UHDR_CR_UNSPECIFIED;
  return img;
}

/* ------------------------------------------------------------------------------------------- */

static const uint8_t kICCSig[] = {'I', 'C', 'C', '_', 'P', 'R', 'O', 'F', 'I', 'L', 'E', '\0'};
static const uint8_t kXmpNameSpace[] = "http://ns.adobe.com/xap/1.0/";
static const uint8_t kExifIdCode[] = {'E', 'x', 'i', 'f', '\0', '\0'};
static const uint8_t kIsoNameSpace[] = "urn:iso:std:iso:ts:21496:-1";

/* jpegdecoderhelper.cpp:119-139 */
static void extract_marker(const uint8_t* data, const jo_header_t& h, uint8_t code,
                           const uint8_t* fourcc, size_t fourcc_len, std::vector<JOCTET>& dst,
                           long& payload_offset) {
  unsigned int pos = 2;
  payload_offset = -1;
  for (int i = 0; i < h.nmarkers; i++) {
    pos += 4;
    const jo_marker_t& m = h.markers[i];
    if (m.id == code && m.length > fourcc_len && !memcmp(data + m.offset, fourcc, fourcc_len)) {
      dst.assign(data + m.offset, data + m.offset + m.length);
      payload_offset = pos;
      return;
    }
    pos += m.length;
  }
}

static uhdr_img_fmt_t sampling_format(const jo_frame_t& f) {
  if (f.ncomp == 1) return UHDR_IMG_FMT_8bppYCbCr400;
  float r[6];
  for (int i = 0; i < 3; i++) {
    r[i * 2] = ((float)f.comp[i].h_samp) / f.max_h;
    r[i * 2 + 1] = ((float)f.comp[i].v_samp) / f.max_v;
  }
  if (r[0] == 1 && r[1] == 1 && r[2] == r[4] && r[3] == r[5]) {
    if (r[2] == 1 && r[3] == 1) return UHDR_IMG_FMT_24bppYCbCr444;
    if (r[2] == 1 && r[3] == 0.5) return UHDR_IMG_FMT_16bppYCbCr440;
    if (r[2] == 0.5 && r[3] == 1) return UHDR_IMG_FMT_16bppYCbCr422;
    if (r[2] == 0.5 && r[3] == 0.5) return UHDR_IMG_FMT_12bppYCbCr420;
    if (r[2] == 0.25 && r[3] == 1) return UHDR_IMG_FMT_12bppYCbCr411;
    if (r[2] == 0.25 && r[3] == 0.5) return UHDR_IMG_FMT_10bppYCbCr410;
  }
  return UHDR_IMG_FMT_UNSPECIFIED;
}

uhdr_error_info_t JpegDecoderHelper::decompressImage(const void* image, size_t length,
                                                     decode_mode_t mode) {
  if (image == nullptr) return err(UHDR_CODEC_INVALID_PARAM, "received nullptr for compressed image data");
  if (length <= 0) return err(UHDR_CODEC_INVALID_PARAM, "received bad compressed image size");
  mResultBuffer.clear();
  mXMPBuffer.clear();
  mEXIFBuffer.clear();
  mICCBuffer.clear();
  mIsoMetadataBuffer.clear();
  mOutFormat = UHDR_IMG_FMT_UNSPECIFIED;
  mNumComponents = 1;
  for (int i = 0; i < kMaxNumComponents; i++) {
    mPlanesMCURow[i].reset();
    mPlaneWidth[i] = mPlaneHeight[i] = mPlaneHStride[i] = mPlaneVStride[i] = 0;
  }
  mExifPayLoadOffset = -1;

  const uint8_t* data = static_cast<const uint8_t*>(image);
  jo_header_t h;
  if (jo_read_header(data, length, &h) != 0)
    return err(UHDR_CODEC_ERROR, "oracle shim: unable to parse jpeg header");
  long off = -1;
  extract_marker(data, h, 0xE1, kXmpNameSpace, sizeof kXmpNameSpace, mXMPBuffer, off);
  extract_marker(data, h, 0xE1, kExifIdCode, sizeof kExifIdCode, mEXIFBuffer, mExifPayLoadOffset);
  extract_marker(data, h, 0xE2, kICCSig, sizeof kICCSig, mICCBuffer, off);
  extract_marker(data, h, 0xE2, kIsoNameSpace, sizeof kIsoNameSpace, mIsoMetadataBuffer, off);

  const jo_frame_t& f = h.frame;
  if (f.width < 1 || f.height < 1) return err(UHDR_CODEC_ERROR, "received bad image width or height");
  if (f.width > kMaxWidth || f.height > kMaxHeight) return err(UHDR_CODEC_ERROR, "image too large");
  mNumComponents = f.ncomp;
  for (int i = 0; i < f.ncomp; i++) {
    mPlaneWidth[i] = std::ceil(((float)f.width * f.comp[i].h_samp) / f.max_h);
    mPlaneHStride[i] = mPlaneWidth[i];
    mPlaneHeight[i] = std::ceil(((float)f.height * f.comp[i].v_samp) / f.max_v);
    mPlaneVStride[i] = mPlaneHeight[i];
  }
  if (f.ncomp == 3) {
    if (mPlaneWidth[1] > mPlaneWidth[0] || mPlaneHeight[2] > mPlaneHeight[0])
      return err(UHDR_CODEC_ERROR, "cb, cr planes are upsampled wrt luma plane");
    if (mPlaneWidth[1] != mPlaneWidth[2] || mPlaneHeight[1] != mPlaneHeight[2])
      return err(UHDR_CODEC_ERROR, "cb, cr planes are not sampled identically");
  }
  if (PARSE_STREAM == mode) return g_no_error;
  if (DECODE_STREAM == mode) mode = f.ncomp == 1 ? DECODE_TO_YCBCR_CS : DECODE_TO_RGB_CS;

  std::vector<int16_t> cbuf[3];
  std::vector<uint8_t> pbuf[3];
  int16_t* coefs[3] = {nullptr, nullptr, nullptr};
  uint8_t* planes[3] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < f.ncomp; c++) {
    cbuf[c].resize((size_t)f.comp[c].wblocks * f.comp[c].hblocks * 64);
    pbuf[c].resize((size_t)f.comp[c].wblocks * f.comp[c].hblocks * 64);
    coefs[c] = cbuf[c].data();
    planes[c] = pbuf[c].data();
  }
  if (jo_decode_coefs(data, length, &h, coefs) != 0)
    return err(UHDR_CODEC_ERROR, "oracle shim: entropy decode failed");
  jo_inverse(&h, coefs, planes);

  if (DECODE_TO_RGB_CS == mode) {
    mPlaneHStride[0] = f.width;
    mPlaneVStride[0] = f.height;
    for (int i = 1; i < kMaxNumComponents; i++) mPlaneHStride[i] = mPlaneVStride[i] = 0;
    mResultBuffer.resize((size_t)f.width * f.height * 4);
    /* jpeg_read_scanlines with JCS_EXT_RGBA: libjpeg's default fancy upsampling + jdcolor */
    if (f.ncomp != 3 || jo_planes_to_rgba(&h, planes, mResultBuffer.data()) != 0)
      return err(UHDR_CODEC_UNSUPPORTED_FEATURE,
                 "oracle shim: RGB output only for 4:4:4 / 4:2:2 / 4:2:0 three-component streams");
    mOutFormat = UHDR_IMG_FMT_32bppRGBA8888;
  } else {
    size_t size = 0;
    for (int i = 0; i < f.ncomp; i++) {
      mPlaneHStride[i] = ((mPlaneWidth[i] + f.max_h - 1) / f.max_h) * f.max_h;
      mPlaneVStride[i] = ((mPlaneHeight[i] + f.max_v - 1) / f.max_v) * f.max_v;
      size += (size_t)mPlaneHStride[i] * mPlaneVStride[i];
    }
    mResultBuffer.assign(size, 0);
    mOutFormat = sampling_format(f);
    uint8_t* dst = mResultBuffer.data();
    for (int i = 0; i < f.ncomp; i++) {
      const int pw = f.comp[i].wblocks * 8;
      /* jpegdecoderhelper.cpp:468-535: rows < VStride are written; full aligned rows when the
       * stride is already a multiple of 8, else only mPlaneWidth bytes are copied back */
      const bool aligned = (mPlaneHStride[i] % 8) == 0;
      for (unsigned y = 0; y < mPlaneVStride[i] && y < (unsigned)f.comp[i].hblocks * 8; y++)
        memcpy(dst + (size_t)y * mPlaneHStride[i], planes[i] + (size_t)y * pw,
               aligned ? mPlaneHStride[i] : mPlaneWidth[i]);
      dst += (size_t)mPlaneHStride[i] * mPlaneVStride[i];
    }
  }
  return g_no_error;
}

uhdr_raw_image_t JpegDecoderHelper::getDecompressedImage() {
  uhdr_raw_image_t img;
  img.fmt = mOutFormat;
  img.cg = UHDR_CG_UNSPECIFIED;
  img.ct = UHDR_CT_UNSPECIFIED;
  img.range = UHDR_CR_FULL_RANGE;
  img.w = mPlaneWidth[0];
  img.h = mPlaneHeight[0];
  uint8_t* data = mResultBuffer.data();
  for (int i = 0; i < 3; i++) {
    img.planes[i] = data;
    img.stride[i] = mPlaneHStride[i];
    data += (size_t)mPlaneHStride[i] * mPlaneVStride[i];
  }
  return img;
}

}  // namespace ultrahdr
