/*
 * jpeg_oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the baseline-JPEG arithmetic that the reference reaches through
 * libjpeg-turbo (pinned 3.1.0 at /root/reference/CMakeLists.txt:507-508; the library itself is
 * NOT under /root/reference).  Call sites being restated:
 *   encode: lib/src/jpegencoderhelper.cpp:131-244 (jpeg_set_defaults, jpeg_set_quality(q,TRUE),
 *           raw_data_in, JDCT_ISLOW, APP2/COM markers), :246-309 (iMCU row feeder + padding)
 *   decode: lib/src/jpegdecoderhelper.cpp:212-411 (marker capture, raw_data_out / JCS_EXT_RGBA,
 *           JDCT_ISLOW), :446-535
 * The published algorithm restated here is libjpeg-turbo's jcparam.c (quality scaling, Annex-K
 * tables), jccolor.c (RGB->YCbCr fixed point), jfdctint.c (LL&M islow FDCT), jcdctmgr.c
 * (quantiser), jccoefct.c (dummy blocks), jchuff.c (baseline Huffman), jcmarker.c (headers),
 * jdhuff.c / jidctint.c / jdcolor.c (inverse path).
 *
 * Pinned (tests/test_oracle_jpeg.py) against Pillow's bundled libjpeg-turbo 3.1.4.1:
 * byte-identical streams for gray / RGB 4:4:4 at many qualities, `jpeg_fdct_islow` /
 * `jpeg_idct_islow` via ctypes, and decoded planes equal to Pillow's decode.
 */
#ifndef UHDR_JPEG_ORACLE_H
#define UHDR_JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* image layouts understood by the codec (values = uhdr_img_fmt_t, ultrahdr_api.h) */
enum {
  JO_FMT_YUV420 = 1,  /* UHDR_IMG_FMT_12bppYCbCr420 */
  JO_FMT_Y400 = 2,    /* UHDR_IMG_FMT_8bppYCbCr400  */
  JO_FMT_RGBA8888 = 3,/* decode output only */
  JO_FMT_YUV444 = 6,  /* UHDR_IMG_FMT_24bppYCbCr444 */
  JO_FMT_YUV422 = 7,  /* UHDR_IMG_FMT_16bppYCbCr422 */
  JO_FMT_RGB888 = 11  /* UHDR_IMG_FMT_24bppRGB888   */
};

/* quality -> two 8-bit quant tables in natural (row-major) order. jcparam.c jpeg_set_quality
 * with force_baseline = TRUE */
void jo_quant_tables(int quality, uint16_t lum[64], uint16_t chr[64]);

/* in-place islow forward DCT of one block of level-shifted samples (output scaled by 8) */
void jo_fdct_islow(int16_t blk[64]);
/* quantise FDCT output with divisor 8*q, round half away from zero */
void jo_quantize(const int16_t in[64], const uint16_t q[64], int16_t out[64]);
/* dequantise + islow inverse DCT + 128 + clamp, writes 8 rows of 8 samples */
void jo_idct_islow(const int16_t coef[64], const uint16_t q[64], uint8_t* out, int out_stride);

/* RGB -> YCbCr (jccolor.c) and YCbCr -> RGB (jdcolor.c), one pixel */
void jo_rgb_to_ycc(int r, int g, int b, uint8_t* y, uint8_t* cb, uint8_t* cr);
void jo_ycc_to_rgb(int y, int cb, int cr, uint8_t* r, uint8_t* g, uint8_t* b);

/* Geometry of one colour component after libjpeg's block padding */
typedef struct {
  int h_samp, v_samp;     /* sampling factors */
  int width, height;      /* real plane size (ceil) */
  int wblocks, hblocks;   /* width_in_blocks, height_in_blocks */
  int tq;                 /* quant table selector */
} jo_comp_t;

typedef struct {
  int ncomp, width, height, max_h, max_v;
  int mcus_per_row, mcu_rows;
  jo_comp_t comp[3];
  uint16_t qt[2][64];     /* natural order */
} jo_frame_t;

/* Fill frame geometry for (fmt, width, height, quality). returns 0 on success */
int jo_frame_init(jo_frame_t* f, int fmt, int width, int height, int quality);

/* Stage 1 of encode: padded sample planes -> quantised coefficients.
 * coefs[c] receives wblocks*hblocks blocks (raster order), 64 int16 each, natural order.
 * planes/strides as the reference passes them to JpegEncoderHelper::compressImage; the padding
 * rules of compressYCbCr (jpegencoderhelper.cpp:254-296) and of libjpeg's edge expansion for the
 * scanline (RGB888) path are applied here. */
int jo_forward(const jo_frame_t* f, int fmt, const uint8_t* const planes[3],
               const unsigned strides[3], int16_t* coefs[3]);

/* Stage 2 of encode: coefficient blocks -> complete JFIF stream (SOI..EOI), laid out exactly as
 * libjpeg writes it: SOI, JFIF APP0, [APP2 icc], [COM], DQT.., SOF0, DHT.., SOS, scan, EOI.
 * *out is malloc'ed. */
int jo_write_stream(const jo_frame_t* f, int16_t* const coefs[3], const uint8_t* icc,
                    size_t icc_size, const char* comment, uint8_t** out, size_t* out_size);

/* One call: both stages. */
int jo_encode(const uint8_t* const planes[3], const unsigned strides[3], int width, int height,
              int fmt, int quality, const uint8_t* icc, size_t icc_size, const char* comment,
              uint8_t** out, size_t* out_size);

/* Decoder ------------------------------------------------------------------------------------ */
typedef struct {
  uint8_t id;        /* 0xE0 + n */
  size_t offset;     /* offset of payload (after the 2 length bytes) in the stream */
  size_t length;     /* payload length */
} jo_marker_t;

typedef struct {
  jo_frame_t frame;
  int comp_id[3];
  int restart_interval;
  /* entropy-coded segment */
  size_t scan_offset, scan_end;
  /* APP0..APP2 markers in stream order (what jpeg_save_markers keeps) */
  jo_marker_t markers[64];
  int nmarkers;
  /* Huffman tables */
  uint8_t bits[2][2][17];    /* [class dc/ac][id][1..16] */
  uint8_t vals[2][2][256];
  int have_tbl[2][2];
  int dc_sel[3], ac_sel[3];
} jo_header_t;

/* Parse headers up to and including SOS (jpeg_read_header + start of scan). 0 on success. */
int jo_read_header(const uint8_t* data, size_t size, jo_header_t* h);

/* Entropy-decode the single baseline scan into quantised coefficients (raster block order per
 * component, natural order inside a block).  coefs[c] must hold wblocks*hblocks*64 int16. */
int jo_decode_coefs(const uint8_t* data, size_t size, const jo_header_t* h, int16_t* coefs[3]);

/* Inverse stage: coefficients -> planes of wblocks*8 x hblocks*8 samples (stride wblocks*8). */
void jo_inverse(const jo_header_t* h, int16_t* const coefs[3], uint8_t* planes[3]);
/* fancy chroma upsampling (jdsample.c) + YCbCr->RGB (jdcolor.c) of the planes jo_inverse produced */
int jo_planes_to_rgba(const jo_header_t* h, uint8_t* const planes[3], uint8_t* rgba);

#ifdef __cplusplus
}
#endif
#endif
