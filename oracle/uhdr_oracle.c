/*
 * uhdr_oracle.c -- TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the reference's gain-map
 * math; see uhdr_oracle.h.  Every function cites the reference lines it follows
 * (paths relative to /root/reference).  Compile WITHOUT fp contraction / fast-math.
 */
#include "uhdr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { FMT_P010 = 0, FMT_YUV420 = 1, FMT_Y400 = 2, FMT_RGBA8888 = 3, FMT_RGBAF16 = 4,
       FMT_RGBA1010102 = 5, FMT_YUV444 = 6, FMT_YUV422 = 7, FMT_RGB888 = 11, FMT_YUV444_10 = 12 };
enum { CG_709 = 0, CG_P3 = 1, CG_2100 = 2 };
enum { CT_LINEAR = 0, CT_HLG = 1, CT_PQ = 2, CT_SRGB = 3 };
enum { CR_LIMITED = 0, CR_FULL = 1 };

typedef struct { float r, g, b; } color_t; /* also y,u,v */

/* ------------------------------------------------------------------------------------------- */
/* constants: lib/src/gainmapmath.cpp:86,94,104-105,156,163-164,174-175,187,194,226-227,236,285,
 * 309-311; lib/include/ultrahdr/gainmapmath.h:44-48,549-550,570 */
static const float kSdrWhiteNits = 203.0f, kHlgMaxNits = 1000.0f, kPqMaxNits = 10000.0f;
static const float kSrgbR = 0.212639f, kSrgbG = 0.715169f, kSrgbB = 0.072192f;
#define kSrgbCb (2 * (1 - kSrgbB))
#define kSrgbCr (2 * (1 - kSrgbR))
static const float kP3R = 0.2289746f, kP3G = 0.6917385f, kP3B = 0.0792869f;
static const float kP3YR = 0.299f, kP3YG = 0.587f, kP3YB = 0.114f;
static const float kP3Cb = 1.772f, kP3Cr = 1.402f;
static const float kBt2100R = 0.2627f, kBt2100G = 0.677998f, kBt2100B = 0.059302f;
#define kBt2100Cb (2 * (1 - kBt2100B))
#define kBt2100Cr (2 * (1 - kBt2100R))
static const float kHlgA = 0.17883277f, kHlgB = 0.28466892f, kHlgC = 0.55991073f;
static const float kOotfGamma = 1.2f;
static const float kHdrOffset = 1e-7f, kSdrOffset = 1e-7f;
static const float kMaxPixelFloatHdrLinear = 10000.0f / 203.0f;

/* file-scope statics of the reference are computed once in float; keep them as variables so the
 * compiler folds them in float exactly like g++ does for `static const float` */
static float cSrgbCb, cSrgbCr, cSrgbGCb, cSrgbGCr, cP3GCb, cP3GCr, cBtCb, cBtCr, cBtGCb, cBtGCr;
static float cPqM1, cPqM2, cPqC1, cPqC2, cPqC3;
static int g_init;
static void init_consts(void) {
  if (g_init) return;
  cSrgbCb = kSrgbCb;
  cSrgbCr = kSrgbCr;
  cSrgbGCb = kSrgbB * cSrgbCb / kSrgbG;
  cSrgbGCr = kSrgbR * cSrgbCr / kSrgbG;
  cP3GCb = kP3YB * kP3Cb / kP3YG;
  cP3GCr = kP3YR * kP3Cr / kP3YG;
  cBtCb = kBt2100Cb;
  cBtCr = kBt2100Cr;
  cBtGCb = kBt2100B * cBtCb / kBt2100G;
  cBtGCr = kBt2100R * cBtCr / kBt2100G;
  cPqM1 = 2610.0f / 16384.0f;
  cPqM2 = 2523.0f / 4096.0f * 128.0f;
  cPqC1 = 3424.0f / 4096.0f;
  cPqC2 = 2413.0f / 4096.0f * 32.0f;
  cPqC3 = 2392.0f / 4096.0f * 32.0f;
  g_init = 1;
}

static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static inline float clip_neg(float v) { return v < 0.0f ? 0.0f : v; }

/* ------------------------------------------------------------------------------------------- */
/* transfer functions, gainmapmath.cpp:114-148, 238-265, 313-333.  Unqualified pow/log/exp/sqrt
 * there are the C double functions. */
static float srgb_inv_oetf(float e) {
  if (e <= 0.04045f) return e / 12.92f;
  return (float)pow((double)((e + 0.055f) / 1.055f), (double)2.4f);
}
float uo_srgb_oetf(float e) { /* :139-148, float std::pow */
  const float kPowerExponent = 1.0f / 2.4f;
  if (e <= 0.0031308f) return 12.92f * e;
  return (1.0f + 0.055f) * powf(e, kPowerExponent) - 0.055f;
}
static float hlg_oetf(float e) {
  if (e <= 1.0f / 12.0f) return (float)sqrt((double)(3.0f * e));
  return (float)((double)kHlgA * log((double)(12.0f * e - kHlgB)) + (double)kHlgC);
}
static float hlg_inv_oetf(float e) {
  if (e <= 0.5f) return (float)(pow((double)e, (double)2.0f) / (double)3.0f);
  return (float)((exp((double)((e - kHlgC) / kHlgA)) + (double)kHlgB) / (double)12.0f);
}
static float pq_oetf(float e) {
  if (e <= 0.0f) return 0.0f;
  double p = pow((double)e, (double)cPqM1);
  return (float)pow(((double)cPqC1 + (double)cPqC2 * p) / (1 + (double)cPqC3 * p), (double)cPqM2);
}
static float pq_inv_oetf(float e) {
  float val = (float)pow((double)e, (double)(1 / cPqM2));
  float num = val - cPqC1;
  if (num < 0.0f) num = 0.0f; /* (std::max)(val - kPqC1, 0.0f) */
  return (float)pow((double)(num / (cPqC2 - cPqC3 * val)), (double)(1 / cPqM1));
}

static float *g_lut[5];
static const int g_lut_n[5] = {1024, 4096, 4096, 65536, 65536};
static void build_luts(void) { /* LookUpTable, gainmapmath.h:345-357 */
  init_consts();
  if (g_lut[0]) return;
  for (int w = 0; w < 5; w++) {
    int n = g_lut_n[w];
    float* t = (float*)malloc(sizeof(float) * n);
    for (int i = 0; i < n; i++) {
      float v = (float)i / (float)(n - 1);
      t[i] = w == 0 ? srgb_inv_oetf(v) : w == 1 ? hlg_inv_oetf(v) : w == 2 ? pq_inv_oetf(v)
             : w == 3 ? hlg_oetf(v) : pq_oetf(v);
    }
    g_lut[w] = t;
  }
}
int uo_lut(int which, float* out, int n) {
  build_luts();
  if (which < 0 || which > 4 || n != g_lut_n[which]) return -1;
  memcpy(out, g_lut[which], sizeof(float) * n);
  return 0;
}
/* LUT lookup: index = int32(float(x*(N-1)) + 0.5 (a double literal)), clamped. :126-132 etc. */
static inline float lut_at(int which, float x) {
  int n = g_lut_n[which];
  int32_t v = (int32_t)((double)(x * (float)(n - 1)) + 0.5);
  v = v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
  return g_lut[which][v];
}
static inline color_t lut3(int which, color_t c) {
  color_t o = {lut_at(which, c.r), lut_at(which, c.g), lut_at(which, c.b)};
  return o;
}

/* ------------------------------------------------------------------------------------------- */
/* yuv<->rgb, luminance  gainmapmath.cpp:88-233 */
static color_t yuv_to_rgb(int cg, color_t e) {
  float cr = cg == CG_709 ? cSrgbCr : cg == CG_P3 ? kP3Cr : cBtCr;
  float cb = cg == CG_709 ? cSrgbCb : cg == CG_P3 ? kP3Cb : cBtCb;
  float gcb = cg == CG_709 ? cSrgbGCb : cg == CG_P3 ? cP3GCb : cBtGCb;
  float gcr = cg == CG_709 ? cSrgbGCr : cg == CG_P3 ? cP3GCr : cBtGCr;
  color_t o = {clamp01(e.r + cr * e.b), clamp01(e.r - gcb * e.g - gcr * e.b),
               clamp01(e.r + cb * e.g)};
  return o;
}
static color_t p3_rgb_to_yuv(color_t e) { /* :166-169 */
  float y = kP3YR * e.r + kP3YG * e.g + kP3YB * e.b;
  color_t o = {y, (e.b - y) / kP3Cb, (e.r - y) / kP3Cr};
  return o;
}
static float luminance(int cg, color_t e) {
  if (cg == CG_709) return kSrgbR * e.r + kSrgbG * e.g + kSrgbB * e.b;
  if (cg == CG_P3) return kP3R * e.r + kP3G * e.g + kP3B * e.b;
  return kBt2100R * e.r + kBt2100G * e.g + kBt2100B * e.b;
}

/* gamut matrices, gainmapmath.cpp:603-615 */
static const float kGamut[3][3][9] = {
    /* dst 709 */ {{1, 0, 0, 0, 1, 0, 0, 0, 1},
                   {1.22494f, -0.22494f, 0.0f, -0.042057f, 1.042057f, 0.0f, -0.019638f, -0.078636f,
                    1.098274f},
                   {1.660491f, -0.587641f, -0.07285f, -0.124551f, 1.1329f, -0.008349f, -0.018151f,
                    -0.100579f, 1.11873f}},
    /* dst P3 */ {{0.822462f, 0.177537f, 0.000001f, 0.033194f, 0.966807f, -0.000001f, 0.017083f,
                   0.072398f, 0.91052f},
                  {1, 0, 0, 0, 1, 0, 0, 0, 1},
                  {1.343578f, -0.282179f, -0.061399f, -0.065298f, 1.075788f, -0.01049f, 0.002822f,
                   -0.019598f, 1.016777f}},
    /* dst 2100 */ {{0.627404f, 0.329282f, 0.043314f, 0.069097f, 0.919541f, 0.011362f, 0.016392f,
                     0.088013f, 0.895595f},
                    {0.753833f, 0.198597f, 0.04757f, 0.045744f, 0.941777f, 0.012479f, -0.00121f,
                     0.017601f, 0.983608f},
                    {1, 0, 0, 0, 1, 0, 0, 0, 1}}};
/* getGamutConversionFn(dst, src), :1087-1129; identity returns the input untouched */
static color_t gamut(int dst, int src, color_t e) {
  if (dst == src) return e;
  const float* c = kGamut[dst][src];
  color_t o = {c[0] * e.r + c[1] * e.g + c[2] * e.b, c[3] * e.r + c[4] * e.g + c[5] * e.b,
               c[6] * e.r + c[7] * e.g + c[8] * e.b};
  return o;
}

/* yuv encoding conversion matrices :638-674, index [src][dst] */
static const float kYuvConv[3][3][9] = {
    {{0}, {1.0f, 0.101579f, 0.196076f, 0.0f, 0.989854f, -0.110653f, 0.0f, -0.072453f, 0.983398f},
     {1.0f, -0.016969f, 0.096312f, 0.0f, 0.995306f, -0.051192f, 0.0f, 0.011507f, 1.002637f}},
    {{1.0f, -0.118188f, -0.212685f, 0.0f, 1.018640f, 0.114618f, 0.0f, 0.075049f, 1.025327f}, {0},
     {1.0f, -0.128245f, -0.115879, 0.0f, 1.010016f, 0.061592f, 0.0f, 0.086969f, 1.029350f}},
    {{1.0f, 0.018149f, -0.095132f, 0.0f, 1.004123f, 0.051267f, 0.0f, -0.011524f, 0.996782f},
     {1.0f, 0.117887f, 0.105521f, 0.0f, 0.995211f, -0.059549f, 0.0f, -0.084085f, 0.976518f}, {0}}};

/* ------------------------------------------------------------------------------------------- */
/* pixel access  gainmapmath.cpp:354-492 */
static float half_to_float(uint16_t h) { /* gainmapmath.h:193-216 */
  union { uint32_t u; float f; } o, magic;
  magic.u = 126u << 23;
  unsigned e = (h >> 10) & 0x1f, m = h & 0x3ff;
  if (e == 0) {
    o.u = magic.u + m;
    o.f -= magic.f;
  } else {
    o.u = m << 13;
    o.u |= e == 0x1f ? (255u << 23) : ((127 - 15 + e) << 23);
  }
  o.u |= (uint32_t)(h >> 15) << 31;
  return o.f;
}
static float sanitize1(float v) { /* gainmapmath.h:572-593 */
  if (isfinite(v)) return v < 0.0f ? 0.0f : (v > kMaxPixelFloatHdrLinear ? kMaxPixelFloatHdrLinear : v);
  if (isinf(v)) return v > 0 ? kMaxPixelFloatHdrLinear : 0.0f;
  return 0.0f;
}
static color_t get_pixel(const uo_image_t* im, size_t x, size_t y) {
  color_t c = {0, 0, 0};
  switch (im->fmt) {
    case FMT_YUV444:
    case FMT_YUV422:
    case FMT_YUV420: {
      int hf = im->fmt == FMT_YUV444 ? 1 : 2, vf = im->fmt == FMT_YUV420 ? 2 : 1;
      const uint8_t* Y = (const uint8_t*)im->planes[0];
      const uint8_t* U = (const uint8_t*)im->planes[1];
      const uint8_t* V = (const uint8_t*)im->planes[2];
      uint8_t yy = Y[x + y * im->stride[0]];
      uint8_t u = U[x / hf + (y / vf) * im->stride[1]];
      uint8_t v = V[x / hf + (y / vf) * im->stride[2]];
      c.r = (float)yy * (1 / 255.0f);
      c.g = (float)(u - 128) * (1 / 255.0f);
      c.b = (float)(v - 128) * (1 / 255.0f);
      return c;
    }
    case FMT_Y400: {
      const uint8_t* Y = (const uint8_t*)im->planes[0];
      c.r = (float)Y[x + y * im->stride[0]] * (1 / 255.0f);
      return c;
    }
    case FMT_P010:
    case FMT_YUV444_10: {
      uint16_t yy, u, v;
      if (im->fmt == FMT_P010) {
        const uint16_t* Y = (const uint16_t*)im->planes[0];
        const uint16_t* UV = (const uint16_t*)im->planes[1];
        size_t ui = (y >> 1) * im->stride[1] + (x & ~(size_t)1);
        yy = Y[y * im->stride[0] + x] >> 6;
        u = UV[ui] >> 6;
        v = UV[ui + 1] >> 6;
      } else {
        yy = ((const uint16_t*)im->planes[0])[y * im->stride[0] + x];
        u = ((const uint16_t*)im->planes[1])[y * im->stride[1] + x];
        v = ((const uint16_t*)im->planes[2])[y * im->stride[2] + x];
      }
      if (im->range == CR_FULL) {
        c.r = (float)yy / 1023.0f;
        c.g = (float)u / 1023.0f - 0.5f;
        c.b = (float)v / 1023.0f - 0.5f;
      } else {
        c.r = (float)(yy - 64) * (1 / 876.0f);
        c.g = (float)(u - 64) * (1 / 896.0f) - 0.5f;
        c.b = (float)(v - 64) * (1 / 896.0f) - 0.5f;
      }
      return c;
    }
    case FMT_RGB888: {
      const uint8_t* p = (const uint8_t*)im->planes[0] + x * 3 + y * im->stride[0] * 3;
      c.r = (float)p[0] / 255.0f;
      c.g = (float)p[1] / 255.0f;
      c.b = (float)p[2] / 255.0f;
      return c;
    }
    case FMT_RGBA8888: {
      uint32_t p = ((const uint32_t*)im->planes[0])[x + y * im->stride[0]];
      c.r = (float)(p & 0xff) / 255.0f;
      c.g = (float)((p >> 8) & 0xff) / 255.0f;
      c.b = (float)((p >> 16) & 0xff) / 255.0f;
      return c;
    }
    case FMT_RGBA1010102: {
      uint32_t p = ((const uint32_t*)im->planes[0])[x + y * im->stride[0]];
      c.r = (float)(p & 0x3ff) / 1023.0f;
      c.g = (float)((p >> 10) & 0x3ff) / 1023.0f;
      c.b = (float)((p >> 20) & 0x3ff) / 1023.0f;
      return c;
    }
    case FMT_RGBAF16: {
      uint64_t p = ((const uint64_t*)im->planes[0])[x + y * im->stride[0]];
      c.r = sanitize1(half_to_float((uint16_t)(p & 0xffff)));
      c.g = sanitize1(half_to_float((uint16_t)((p >> 16) & 0xffff)));
      c.b = sanitize1(half_to_float((uint16_t)((p >> 32) & 0xffff)));
      return c;
    }
  }
  return c;
}
/* samplePixels :494-504: float accumulation in row-major order, then / float(s*s) */
static color_t sample_pixels(const uo_image_t* im, size_t s, size_t x, size_t y) {
  color_t e = {0, 0, 0};
  for (size_t dy = 0; dy < s; dy++)
    for (size_t dx = 0; dx < s; dx++) {
      color_t p = get_pixel(im, x * s + dx, y * s + dy);
      e.r += p.r;
      e.g += p.g;
      e.b += p.b;
    }
  float d = (float)(s * s);
  e.r /= d;
  e.g /= d;
  e.b /= d;
  return e;
}
static int is_rgb_fmt(int f) { return f == FMT_RGBAF16 || f == FMT_RGBA8888 || f == FMT_RGBA1010102; }

/* ------------------------------------------------------------------------------------------- */
/* gain encode helpers :758-789 */
int uo_encode_gain(float y_sdr, float y_hdr, const uo_metadata_t* md, float l2min, float l2max,
                   int idx) {
  float gain = 1.0f;
  if (y_sdr > 0.0f) gain = y_hdr / y_sdr;
  if (gain < md->min_content_boost[idx]) gain = md->min_content_boost[idx];
  if (gain > md->max_content_boost[idx]) gain = md->max_content_boost[idx];
  float gn = (float)((log2((double)gain) - (double)l2min) / (double)(l2max - l2min));
  float gg = powf(gn, md->gamma[idx]);
  return (uint8_t)(gg * 255.0f);
}
float uo_compute_gain(float sdr, float hdr) {
  float gain = (float)log2((double)((hdr + kHdrOffset) / (sdr + kSdrOffset)));
  if (sdr < 2.f / 255.0f) gain = gain < 2.3f ? gain : 2.3f; /* (std::min)(gain, 2.3f) */
  return gain;
}
int uo_affine_map_gain(float g, float mn, float mx, float gamma) {
  float m = (g - mn) / (mx - mn);
  if (gamma != 1.0f) m = (float)pow((double)m, (double)gamma);
  m *= 255;
  float t = m + 0.5f;
  t = t < 0 ? 0 : (t > 255 ? 255 : t);
  return (uint8_t)t;
}
unsigned uo_float_to_half(float f) { /* gainmapmath.h:160-173 */
  union { uint32_t u; float f; } x;
  x.f = f;
  const uint32_t b = x.u + 0x00001000;
  const int32_t e = (b & 0x7F800000) >> 23;
  const uint32_t m = b & 0x007FFFFF;
  uint32_t r = (b & 0x80000000) >> 16;
  if (e > 112) r |= (((uint32_t)(e - 112) << 10) & 0x7C00) | (m >> 13);
  if (e < 113 && e > 101) r |= (((0x007FF000 + m) >> (125 - e)) + 1) >> 1;
  if (e > 143) r |= 0x7FFF;
  return r & 0xFFFF;
}

/* ShepardsIDW :39-80 */
static void fill_idw(float* w, int s, int incR, int incB) {
  for (int y = 0; y < s; y++)
    for (int x = 0; x < s; x++) {
      float pos_x = ((float)x) / s, pos_y = ((float)y) / s;
      int curr_x = (int)floor((double)pos_x), curr_y = (int)floor((double)pos_y);
      int next_x = curr_x + incR, next_y = curr_y + incB;
#define DIST(x1, x2, y1, y2) \
  ((float)sqrt((double)((((y2) - (y1)) * ((y2) - (y1))) + ((x2) - (x1)) * ((x2) - (x1)))))
      float cx = (float)curr_x, cy = (float)curr_y, nx = (float)next_x, ny = (float)next_y;
      float e1d = DIST(pos_x, cx, pos_y, cy);
      float* o = w + (y * s + x) * 4;
      if (e1d == 0) {
        o[0] = 1.f;
        o[1] = o[2] = o[3] = 0.f;
      } else {
        float e1 = 1.f / e1d;
        float e2 = 1.f / DIST(pos_x, cx, pos_y, ny);
        float e3 = 1.f / DIST(pos_x, nx, pos_y, cy);
        float e4 = 1.f / DIST(pos_x, nx, pos_y, ny);
        float tot = e1 + e2 + e3 + e4;
        o[0] = e1 / tot;
        o[1] = e2 / tot;
        o[2] = e3 / tot;
        o[3] = e4 / tot;
      }
    }
}
void uo_idw_weights(int scale, int variant, float* out) {
  static const int inc[4][2] = {{1, 1}, {0, 1}, {1, 0}, {0, 0}};
  fill_idw(out, scale, inc[variant][0], inc[variant][1]);
}

/* GainLUT gainmapmath.h:452-489 */
static int md_single_channel(const uo_metadata_t* m) {
#define SAME(a) (m->a[0] == m->a[1] && m->a[0] == m->a[2])
  return SAME(max_content_boost) && SAME(min_content_boost) && SAME(gamma) && SAME(offset_sdr) &&
         SAME(offset_hdr);
}
void uo_gain_lut(const uo_metadata_t* md, float weight, float* out) {
  int single = md_single_channel(md);
  for (int c = 0; c < (single ? 1 : 3); c++)
    for (int i = 0; i < 1024; i++) {
      float value = (float)i / (float)1023;
      float logBoost = (float)(log2((double)md->min_content_boost[c]) * (double)(1.0f - value) +
                               log2((double)md->max_content_boost[c]) * (double)value);
      out[c * 1024 + i] = (float)exp2((double)(logBoost * weight));
    }
  if (single) {
    memcpy(out + 1024, out, 4096);
    memcpy(out + 2048, out, 4096);
  }
}
static inline float gain_factor(const float* lut, float gain, float gamma_inv) {
  if (gamma_inv != 1.0f) gain = (float)pow((double)gain, (double)gamma_inv);
  int32_t idx = (int32_t)((double)(gain * (float)1023) + 0.5);
  idx = idx < 0 ? 0 : (idx > 1023 ? 1023 : idx);
  return lut[idx];
}

/* ------------------------------------------------------------------------------------------- */
/* generateGainMap  jpegr.cpp:530-1058 */
static float ref_nits(int ct) { /* gainmapmath.cpp:20-34 */
  return ct == CT_LINEAR ? kPqMaxNits : ct == CT_HLG ? kHlgMaxNits : ct == CT_PQ ? kPqMaxNits
         : ct == CT_SRGB ? kSdrWhiteNits : -1.0f;
}
static color_t hdr_to_linear(int ct, color_t g) {
  color_t l;
  if (ct == CT_HLG) {
    l = lut3(1, g);
    /* hlgOotfApprox :293-295, float std::pow */
    l.r = powf(l.r, kOotfGamma);
    l.g = powf(l.g, kOotfGamma);
    l.b = powf(l.b, kOotfGamma);
    return l;
  }
  if (ct == CT_PQ) return lut3(2, g);
  if (ct == CT_SRGB) return lut3(0, g);
  return g;
}

int uo_generate_gainmap(const uo_image_t* sdr, const uo_image_t* hdr, const uo_gm_config_t* cfg,
                        uo_metadata_t* md, uo_image_t* out) {
  build_luts();
  if (sdr->fmt != FMT_YUV444 && sdr->fmt != FMT_YUV422 && sdr->fmt != FMT_YUV420 &&
      sdr->fmt != FMT_RGBA8888)
    return 6;
  if (hdr->fmt != FMT_P010 && hdr->fmt != FMT_YUV444_10 && hdr->fmt != FMT_RGBA1010102 &&
      hdr->fmt != FMT_RGBAF16)
    return 6;
  float hdr_white_nits = ref_nits(hdr->ct);
  if (hdr_white_nits == -1.0f) return 6;
  /* use_sdr_cg rule :607-638 with kWriteXmpMetadata = false (UHDR_WRITE_XMP off) */
  int use_sdr_cg = 1;
  if (sdr->cg != hdr->cg)
    use_sdr_cg = !(hdr->cg == CG_2100 || (hdr->cg == CG_P3 && sdr->cg != CG_2100));
  md->use_base_cg = use_sdr_cg;
  const int sdr_yuv_cg = cfg->sdr_is_601 ? CG_P3 : sdr->cg;
  int s = cfg->scale_factor;
  unsigned mw = sdr->w / s, mh = sdr->h / s;
  if (mw == 0 || mh == 0) {
    int sf = (int)(sdr->w < sdr->h ? sdr->w : sdr->h);
    s = sf >= 8 ? sf / 8 : 1;
    mw = sdr->w / s;
    mh = sdr->h / s;
  }
  const int multi = cfg->multichannel != 0;
  const int nch = multi ? 3 : 1;
  out->fmt = multi ? FMT_RGB888 : FMT_Y400;
  out->cg = hdr->cg;
  out->ct = hdr->ct;
  out->range = hdr->range;
  out->w = mw;
  out->h = mh;
  out->stride[0] = mw;
  uint8_t* dst = (uint8_t*)out->planes[0];
  const float hdr_nits_factor = hdr->ct == CT_LINEAR ? kSdrWhiteNits : hdr_white_nits;
  const int onepass = cfg->preset == 0; /* UHDR_USAGE_REALTIME */
  float* gains = NULL;
  float gmin[3] = {127.0f, 127.0f, 127.0f}, gmax[3] = {-128.0f, -128.0f, -128.0f};
  float l2min = 0, l2max = 0;
  if (onepass) { /* :724-737 */
    for (int i = 0; i < 3; i++) {
      md->max_content_boost[i] = hdr_white_nits / kSdrWhiteNits;
      md->min_content_boost[i] = 1.0f;
      md->gamma[i] = cfg->gamma;
      md->offset_sdr[i] = 0.0f;
      md->offset_hdr[i] = 0.0f;
    }
    md->hdr_capacity_min = 1.0f;
    md->hdr_capacity_max = cfg->target_disp_peak_nits != -1.0f
                               ? cfg->target_disp_peak_nits / kSdrWhiteNits
                               : md->max_content_boost[0];
    l2min = log2f(md->min_content_boost[0]); /* jpegr.cpp has `using namespace std` -> float */
    l2max = log2f(md->max_content_boost[0]);
  } else {
    gains = (float*)malloc(sizeof(float) * (size_t)mw * mh * nch);
    if (!gains) return 4;
  }
  for (size_t y = 0; y < mh; y++)
    for (size_t x = 0; x < mw; x++) {
      color_t sg = sample_pixels(sdr, s, x, y);
      if (!is_rgb_fmt(sdr->fmt)) sg = yuv_to_rgb(sdr_yuv_cg, sg);
      color_t sl = lut3(0, sg);
      if (!use_sdr_cg) sl = gamut(hdr->cg, sdr->cg, sl);
      sl.r = clip_neg(sl.r);
      sl.g = clip_neg(sl.g);
      sl.b = clip_neg(sl.b);
      color_t hg = sample_pixels(hdr, s, x, y);
      if (!is_rgb_fmt(hdr->fmt)) hg = yuv_to_rgb(hdr->cg, hg);
      color_t hl = hdr_to_linear(hdr->ct, hg);
      if (use_sdr_cg) hl = gamut(sdr->cg, hdr->cg, hl);
      hl.r = clip_neg(hl.r);
      hl.g = clip_neg(hl.g);
      hl.b = clip_neg(hl.b);
      float sv[3], hv[3];
      if (multi) {
        sv[0] = sl.r * kSdrWhiteNits; sv[1] = sl.g * kSdrWhiteNits; sv[2] = sl.b * kSdrWhiteNits;
        hv[0] = hl.r * hdr_nits_factor; hv[1] = hl.g * hdr_nits_factor; hv[2] = hl.b * hdr_nits_factor;
      } else if (cfg->use_luminance) {
        sv[0] = luminance(sdr->cg, sl) * kSdrWhiteNits;
        hv[0] = luminance(sdr->cg, hl) * hdr_nits_factor;
      } else {
        sv[0] = fmaxf(sl.r, fmaxf(sl.g, sl.b)) * kSdrWhiteNits;
        hv[0] = fmaxf(hl.r, fmaxf(hl.g, hl.b)) * hdr_nits_factor;
      }
      size_t idx = (x + y * mw) * nch;
      for (int c = 0; c < nch; c++) {
        if (onepass) {
          dst[idx + c] = (uint8_t)uo_encode_gain(sv[c], hv[c], md, l2min, l2max, c);
        } else {
          float g = uo_compute_gain(sv[c], hv[c]);
          gains[idx + c] = g;
          if (g < gmin[c]) gmin[c] = g;
          if (g > gmax[c]) gmax[c] = g;
        }
      }
    }
  if (!onepass) { /* :969-1048 */
    for (int c = 0; c < nch; c++) {
      gmin[c] = gmin[c] < -14.3f ? -14.3f : (gmin[c] > 15.6f ? 15.6f : gmin[c]);
      gmax[c] = gmax[c] < -14.3f ? -14.3f : (gmax[c] > 15.6f ? 15.6f : gmax[c]);
      if (cfg->max_content_boost != FLT_MAX) {
        float sug = log2f(cfg->max_content_boost);
        if (sug < gmax[c]) gmax[c] = sug;
      }
      if (cfg->min_content_boost != FLT_MIN) {
        float sug = log2f(cfg->min_content_boost);
        if (sug > gmin[c]) gmin[c] = sug;
      }
      if (fabsf(gmax[c] - gmin[c]) < FLT_EPSILON) gmax[c] += 0.1f;
    }
    size_t n = (size_t)mw * mh * nch;
    for (size_t i = 0; i < n; i++)
      dst[i] = (uint8_t)uo_affine_map_gain(gains[i], gmin[i % nch], gmax[i % nch], cfg->gamma);
    free(gains);
    for (int i = 0; i < 3; i++) {
      int c = multi ? i : 0;
      md->max_content_boost[i] = exp2f(gmax[c]);
      md->min_content_boost[i] = exp2f(gmin[c]);
      md->gamma[i] = cfg->gamma;
      md->offset_sdr[i] = kSdrOffset;
      md->offset_hdr[i] = kHdrOffset;
    }
    md->hdr_capacity_min = 1.0f;
    md->hdr_capacity_max = cfg->target_disp_peak_nits != -1.0f
                               ? cfg->target_disp_peak_nits / kSdrWhiteNits
                               : hdr_white_nits / kSdrWhiteNits;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* applyGainMap  jpegr.cpp:1533-1831; samplers gainmapmath.cpp:871-1080 */
static inline size_t minz(size_t a, size_t b) { return a < b ? a : b; }
static inline float map_u8(uint8_t v) { return (float)v / 255.0f; }
static float pyth(float xd, float yd) {
  return (float)sqrt(pow((double)xd, (double)2.0f) + pow((double)yd, (double)2.0f));
}

static void sample_map(const uo_image_t* map, int nch_stride, int nch, float scale, int integer,
                       const float* idw[4], size_t x, size_t y, float out[3]) {
  const uint8_t* data = (const uint8_t*)map->planes[0];
  size_t stride = map->stride[0];
  size_t xl, xu, yl, yu;
  float w[4];
  int early = -1;
  if (integer) {
    size_t s = (size_t)scale;
    xl = x / s; xu = xl + 1; yl = y / s; yu = yl + 1;
    xl = minz(xl, map->w - 1); xu = minz(xu, map->w - 1);
    yl = minz(yl, map->h - 1); yu = minz(yu, map->h - 1);
    size_t ox = x % s, oy = y % s;
    const float* t = idw[0];
    if (xl == xu && yl == yu) t = idw[3];
    else if (xl == xu) t = idw[1];
    else if (yl == yu) t = idw[2];
    t += oy * s * 4 + ox * 4;
    w[0] = t[0]; w[1] = t[1]; w[2] = t[2]; w[3] = t[3];
  } else {
    float xm = (float)x / scale, ym = (float)y / scale;
    xl = (size_t)floor((double)xm); xu = xl + 1;
    yl = (size_t)floor((double)ym); yu = yl + 1;
    xl = minz(xl, map->w - 1); xu = minz(xu, map->w - 1);
    yl = minz(yl, map->h - 1); yu = minz(yu, map->h - 1);
    float d1 = pyth(xm - (float)xl, ym - (float)yl);
    float d2 = pyth(xm - (float)xl, ym - (float)yu);
    float d3 = pyth(xm - (float)xu, ym - (float)yl);
    float d4 = pyth(xm - (float)xu, ym - (float)yu);
    if (d1 == 0.0f) early = 0;
    else if (d2 == 0.0f) early = 1;
    else if (d3 == 0.0f) early = 2;
    else if (d4 == 0.0f) early = nch == 1 ? 1 : 3; /* :908 returns e2 in the 1-channel code */
    else {
      float w1 = 1.0f / d1, w2 = 1.0f / d2, w3 = 1.0f / d3, w4 = 1.0f / d4;
      float tot = w1 + w2 + w3 + w4;
      w[0] = w1 / tot; w[1] = w2 / tot; w[2] = w3 / tot; w[3] = w4 / tot;
    }
  }
  size_t i1 = (xl + yl * stride) * nch_stride, i2 = (xl + yu * stride) * nch_stride,
         i3 = (xu + yl * stride) * nch_stride, i4 = (xu + yu * stride) * nch_stride;
  for (int c = 0; c < nch; c++) {
    float e1 = map_u8(data[i1 + c]), e2 = map_u8(data[i2 + c]), e3 = map_u8(data[i3 + c]),
          e4 = map_u8(data[i4 + c]);
    if (early >= 0) out[c] = early == 0 ? e1 : early == 1 ? e2 : early == 2 ? e3 : e4;
    else out[c] = e1 * w[0] + e2 * w[1] + e3 * w[2] + e4 * w[3];
  }
}

int uo_apply_gainmap(const uo_image_t* sdr, const uo_image_t* gm, const uo_metadata_t* md,
                     int output_ct, int output_fmt, float max_display_boost, uo_image_t* dest) {
  (void)output_fmt;
  build_luts();
  if (!dest || !dest->planes[0] || dest->stride[0] < dest->w) return 3;
  if (output_ct != CT_LINEAR && output_ct != CT_HLG && output_ct != CT_PQ) return 3;
  if ((output_ct == CT_LINEAR && dest->fmt != FMT_RGBAF16) ||
      (output_ct != CT_LINEAR && dest->fmt != FMT_RGBA1010102))
    return 3;
  int sdr_cg = sdr->cg == -1 ? CG_709 : sdr->cg;
  int hdr_cg = gm->cg == -1 ? sdr_cg : gm->cg;
  dest->cg = hdr_cg;
  {
    float pa = (float)sdr->w / sdr->h, ga = (float)gm->w / gm->h;
    if (fabsf(pa - ga) / pa > 0.01f) return 6; /* resize path (editorhelper) not restated */
  }
  float scale = (float)sdr->w / gm->w;
  int srnd = (int)roundf(scale);
  if (srnd < 1) srnd = 1;
  float* idwbuf = (float*)malloc(sizeof(float) * 4 * 4 * srnd * srnd);
  const float* idw[4];
  for (int v = 0; v < 4; v++) {
    uo_idw_weights(srnd, v, idwbuf + (size_t)v * 4 * srnd * srnd);
    idw[v] = idwbuf + (size_t)v * 4 * srnd * srnd;
  }
  float display_boost = max_display_boost < md->hdr_capacity_max ? max_display_boost
                                                                  : md->hdr_capacity_max;
  float weight;
  if (display_boost != md->hdr_capacity_max) { /* float log2: jpegr.cpp `using namespace std` */
    weight = (log2f(display_boost) - log2f(md->hdr_capacity_min)) /
             (log2f(md->hdr_capacity_max) - log2f(md->hdr_capacity_min));
    weight = weight < 0.0f ? 0.0f : (weight > 1.0f ? 1.0f : weight);
  } else {
    weight = 1.0f;
  }
  float* glut = (float*)malloc(sizeof(float) * 3 * 1024);
  uo_gain_lut(md, weight, glut);
  int single_md = md_single_channel(md);
  float ginv[3];
  for (int c = 0; c < 3; c++) ginv[c] = 1.0f / md->gamma[single_md ? 0 : c];
  const int integer = scale == floorf(scale);
  const int map1 = gm->fmt == FMT_Y400;
  const int nstride = gm->fmt == FMT_RGBA8888 ? 4 : (map1 ? 1 : 3);
  for (size_t y = 0; y < sdr->h; y++)
    for (size_t x = 0; x < sdr->w; x++) {
      color_t g = get_pixel(sdr, x, y);
      /* always BT.601 (:1723); isPixelFormatRgb() is false for RGB888 so the reference runs
       * the yuv->rgb step on it too -- restated as is */
      if (!is_rgb_fmt(sdr->fmt)) g = yuv_to_rgb(CG_P3, g);
      color_t l = lut3(0, g);
      if (!md->use_base_cg) l = gamut(hdr_cg, sdr_cg, l);
      float gain[3];
      sample_map(gm, nstride, map1 ? 1 : 3, scale, integer, idw, x, y, gain);
      color_t h;
      if (map1) { /* applyGainLUT(Color, float) :807-810: table 0, offsets [0] */
        float f = gain_factor(glut, gain[0], ginv[0]);
        h.r = ((l.r + md->offset_sdr[0]) * f) - md->offset_hdr[0];
        h.g = ((l.g + md->offset_sdr[0]) * f) - md->offset_hdr[0];
        h.b = ((l.b + md->offset_sdr[0]) * f) - md->offset_hdr[0];
      } else {
        float fr = gain_factor(glut, gain[0], ginv[0]);
        float fg = gain_factor(glut + 1024, gain[1], ginv[1]);
        float fb = gain_factor(glut + 2048, gain[2], ginv[2]);
        h.r = ((l.r + md->offset_sdr[0]) * fr) - md->offset_hdr[0];
        h.g = ((l.g + md->offset_sdr[1]) * fg) - md->offset_hdr[1];
        h.b = ((l.b + md->offset_sdr[2]) * fb) - md->offset_hdr[2];
      }
      size_t pi = x + y * dest->stride[0];
      if (output_ct == CT_LINEAR) {
        if (md->use_base_cg) h = gamut(hdr_cg, sdr_cg, h);
        float v[3] = {h.r, h.g, h.b};
        uint64_t px = (uint64_t)uo_float_to_half(1.0f) << 48;
        for (int c = 0; c < 3; c++) {
          float t = v[c] < 0.0f ? 0.0f : (v[c] > kMaxPixelFloatHdrLinear ? kMaxPixelFloatHdrLinear : v[c]);
          px |= (uint64_t)uo_float_to_half(t) << (16 * c);
        }
        ((uint64_t*)dest->planes[0])[pi] = px;
      } else {
        float maxn = output_ct == CT_HLG ? kHlgMaxNits : kPqMaxNits;
        h.r = h.r * kSdrWhiteNits / maxn;
        h.g = h.g * kSdrWhiteNits / maxn;
        h.b = h.b * kSdrWhiteNits / maxn;
        if (md->use_base_cg) h = gamut(hdr_cg, sdr_cg, h);
        h.r = clamp01(h.r);
        h.g = clamp01(h.g);
        h.b = clamp01(h.b);
        if (output_ct == CT_HLG) { /* hlgInverseOotfApprox :303-306, float std::pow */
          const float p = 1.0f / kOotfGamma;
          h.r = powf(h.r, p);
          h.g = powf(h.g, p);
          h.b = powf(h.b, p);
        }
        color_t e = lut3(output_ct == CT_HLG ? 3 : 4, h);
        /* colorToRgba1010102 :1279-1284 */
        float v[3] = {e.r, e.g, e.b};
        uint32_t px = 0x3u << 30;
        for (int c = 0; c < 3; c++) {
          float t = v[c] * 1023 + 0.5f;
          t = t < 0.0f ? 0.0f : (t > 1023.0f ? 1023.0f : t);
          px |= (uint32_t)t << (10 * c);
        }
        ((uint32_t*)dest->planes[0])[pi] = px;
      }
    }
  free(glut);
  free(idwbuf);
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* toneMap  jpegr.cpp:1945-2222 */
static uint8_t scale_to_8bit(float v) { /* :1979-1983 std::round */
  int i = (int)roundf(v * 255.0f);
  return (uint8_t)(i < 0 ? 0 : (i > 255 ? 255 : i));
}
static color_t global_tonemap(color_t in, float headroom, int normalized) { /* :1951-1977 */
  color_t h = in;
  if (normalized) { h.r = in.r * headroom; h.g = in.g * headroom; h.b = in.b * headroom; }
  float max_hdr = h.r;
  if (h.g > max_hdr) max_hdr = h.g; /* std::max_element: first largest */
  if (h.b > max_hdr) max_hdr = h.b;
  float o = 1.0f + max_hdr / (headroom * headroom); /* ReinhardMap :1945-1949 */
  o /= 1.0f + max_hdr;
  float max_sdr = o * max_hdr;
  color_t s;
  s.r = h.r > 0.0f ? h.r * max_sdr / max_hdr : 0.0f;
  s.g = h.g > 0.0f ? h.g * max_sdr / max_hdr : 0.0f;
  s.b = h.b > 0.0f ? h.b * max_sdr / max_hdr : 0.0f;
  return s;
}
int uo_tonemap(const uo_image_t* hdr, uo_image_t* sdr) {
  build_luts();
  if (hdr->fmt != FMT_P010 && hdr->fmt != FMT_YUV444_10 && hdr->fmt != FMT_RGBA1010102 &&
      hdr->fmt != FMT_RGBAF16)
    return 6;
  if (hdr->fmt == FMT_P010 && sdr->fmt != FMT_YUV420) return 6;
  if (hdr->fmt == FMT_YUV444_10 && sdr->fmt != FMT_YUV444) return 6;
  if ((hdr->fmt == FMT_RGBA1010102 || hdr->fmt == FMT_RGBAF16) && sdr->fmt != FMT_RGBA8888) return 6;
  float nits = ref_nits(hdr->ct);
  if (nits == -1.0f) return 6;
  sdr->cg = CG_P3;
  sdr->ct = CT_SRGB;
  sdr->range = CR_FULL;
  const int f = hdr->fmt == FMT_P010 ? 2 : 1;
  const int normalized = hdr->ct != CT_LINEAR;
  const float headroom = nits / kSdrWhiteNits;
  uint8_t* Y = (uint8_t*)sdr->planes[0];
  uint8_t* U = (uint8_t*)sdr->planes[1];
  uint8_t* V = (uint8_t*)sdr->planes[2];
  for (size_t y = 0; y < hdr->h; y += f)
    for (size_t x = 0; x < hdr->w; x += f) {
      float su = 0.0f, sv = 0.0f;
      for (int i = 0; i < f; i++)
        for (int j = 0; j < f; j++) {
          color_t g = get_pixel(hdr, x + j, y + i);
          if (!is_rgb_fmt(hdr->fmt)) g = yuv_to_rgb(hdr->cg, g);
          color_t l = hdr_to_linear(hdr->ct, g);
          color_t t = global_tonemap(l, headroom, normalized);
          t = gamut(CG_P3, hdr->cg, t);
          t.r = clamp01(t.r);
          t.g = clamp01(t.g);
          t.b = clamp01(t.b);
          color_t e = {uo_srgb_oetf(t.r), uo_srgb_oetf(t.g), uo_srgb_oetf(t.b)};
          if (sdr->fmt == FMT_RGBA8888) { /* putRgba8888Pixel :538-552 */
            float v[3] = {e.r * 255.0f + 0.5f, e.g * 255.0f + 0.5f, e.b * 255.0f + 0.5f};
            uint32_t px = 255u << 24;
            for (int c = 0; c < 3; c++) {
              float q = v[c] < 0.0f ? 0.0f : (v[c] > 255.0f ? 255.0f : v[c]);
              px |= (uint32_t)(int32_t)q << (8 * c);
            }
            ((uint32_t*)sdr->planes[0])[(x + j) + (y + i) * sdr->stride[0]] = px;
          } else {
            color_t yuv = p3_rgb_to_yuv(e);
            yuv.g += 0.5f;
            yuv.b += 0.5f;
            if (sdr->fmt == FMT_YUV444) { /* putYuv444Pixel :579-596 */
              float v[3] = {yuv.r * 255.0f + 0.5f, yuv.g * 255.0f + 0.5f, yuv.b * 255.0f + 0.5f};
              for (int c = 0; c < 3; c++) v[c] = v[c] < 0.0f ? 0.0f : (v[c] > 255.0f ? 255.0f : v[c]);
              Y[(x + j) + (y + i) * sdr->stride[0]] = (uint8_t)v[0];
              U[(x + j) + (y + i) * sdr->stride[1]] = (uint8_t)v[1];
              V[(x + j) + (y + i) * sdr->stride[2]] = (uint8_t)v[2];
            } else {
              Y[(y + i) * sdr->stride[0] + x + j] = scale_to_8bit(yuv.r);
              su += yuv.g;
              sv += yuv.b;
            }
          }
        }
      if (sdr->fmt == FMT_YUV420) {
        su /= (float)(f * f);
        sv /= (float)(f * f);
        U[x / f + (y / f) * sdr->stride[1]] = scale_to_8bit(su);
        V[x / f + (y / f) * sdr->stride[2]] = scale_to_8bit(sv);
      }
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* convertYuv jpegr.cpp:436-518, transformYuv420/444 gainmapmath.cpp:676-748 */
static inline color_t yuv_conv(color_t e, const float* c) {
  color_t o = {e.r * c[0] + e.g * c[1] + e.b * c[2], e.r * c[3] + e.g * c[4] + e.b * c[5],
               e.r * c[6] + e.g * c[7] + e.b * c[8]};
  return o;
}
static inline uint8_t clip255(float v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
int uo_convert_yuv(uo_image_t* im, int src, int dst) {
  if (src < 0 || src > 2 || dst < 0 || dst > 2) return 3;
  if (src == dst) return 0;
  const float* c = kYuvConv[src][dst];
  uint8_t* Y = (uint8_t*)im->planes[0];
  uint8_t* U = (uint8_t*)im->planes[1];
  uint8_t* V = (uint8_t*)im->planes[2];
  if (im->fmt == FMT_YUV420) {
    for (size_t y = 0; y < im->h / 2; y++)
      for (size_t x = 0; x < im->w / 2; x++) {
        color_t p[4];
        for (int k = 0; k < 4; k++)
          p[k] = yuv_conv(get_pixel(im, x * 2 + (k & 1), y * 2 + (k >> 1)), c);
        float nu = (((p[0].g + p[1].g) + p[2].g) + p[3].g) / 4.0f;
        float nv = (((p[0].b + p[1].b) + p[2].b) + p[3].b) / 4.0f;
        for (int k = 0; k < 4; k++)
          Y[(x * 2 + (k & 1)) + (y * 2 + (k >> 1)) * im->stride[0]] = clip255(p[k].r * 255.0f + 0.5f);
        U[x + y * im->stride[1]] = clip255(nu * 255.0f + 128.0f + 0.5f);
        V[x + y * im->stride[2]] = clip255(nv * 255.0f + 128.0f + 0.5f);
      }
    return 0;
  }
  if (im->fmt == FMT_YUV444) {
    for (size_t y = 0; y < im->h; y++)
      for (size_t x = 0; x < im->w; x++) {
        color_t p = yuv_conv(get_pixel(im, x, y), c);
        Y[x + y * im->stride[0]] = clip255(p.r * 255.0f + 0.5f);
        U[x + y * im->stride[1]] = clip255(p.g * 255.0f + 128.0f + 0.5f);
        V[x + y * im->stride[2]] = clip255(p.b * 255.0f + 128.0f + 0.5f);
      }
    return 0;
  }
  return 6;
}

/* reference expression of computeGain's logarithm (gainmapmath.cpp:774): double log2 of a float,
 * narrowed to float -- vector form for the device log2 probe test */
void uo_log2_of_float(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = (float)log2((double)in[i]);
}

/* float std::pow as the reference calls it (gainmapmath.cpp:147,294,304): vector form for the
 * device powf probe test */
void uo_powf_vec(const float* in, float y, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = powf(in[i], y);
}
