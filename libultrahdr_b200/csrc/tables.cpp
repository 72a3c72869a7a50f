// See tables.h.  Host code; compiled with -ffp-contract=off so no FMA is formed.
#include "tables.h"

#include <cmath>
#include <cstring>

namespace uhdr_b200 {

namespace {
// BT.2100 / sRGB / P3 constants (gainmapmath.cpp:86,156,163-164,187,236,309-311)
constexpr float kSrgbR = 0.212639f, kSrgbG = 0.715169f, kSrgbB = 0.072192f;
constexpr float kP3R = 0.2289746f, kP3G = 0.6917385f, kP3B = 0.0792869f;
constexpr float kP3YR = 0.299f, kP3YG = 0.587f, kP3YB = 0.114f, kP3Cb = 1.772f, kP3Cr = 1.402f;
constexpr float kBtR = 0.2627f, kBtG = 0.677998f, kBtB = 0.059302f;
constexpr float kHlgA = 0.17883277f, kHlgB = 0.28466892f, kHlgC = 0.55991073f;

// volatile-free float evaluation helpers: everything is plain float arithmetic
inline float f2(float a) { return 2 * (1 - a); }

struct PqConst {
  float m1 = 2610.0f / 16384.0f;
  float m2 = 2523.0f / 4096.0f * 128.0f;
  float c1 = 3424.0f / 4096.0f;
  float c2 = 2413.0f / 4096.0f * 32.0f;
  float c3 = 2392.0f / 4096.0f * 32.0f;
};

// transfer functions.  The reference calls the *double* libm functions on float arguments at
// these sites (gainmapmath.cpp:114-120, 238-244, 259-265, 313-333) and narrows the result.
float srgb_eotf(float v) {
  if (v <= 0.04045f) return v / 12.92f;
  double base = (double)((v + 0.055f) / 1.055f);
  return (float)std::pow(base, (double)2.4f);
}
float hlg_oetf(float e) {
  if (e <= 1.0f / 12.0f) return (float)std::sqrt((double)(3.0f * e));
  return (float)((double)kHlgA * std::log((double)(12.0f * e - kHlgB)) + (double)kHlgC);
}
float hlg_eotf(float v) {
  if (v <= 0.5f) return (float)(std::pow((double)v, 2.0) / 3.0);
  return (float)((std::exp((double)((v - kHlgC) / kHlgA)) + (double)kHlgB) / 12.0);
}
float pq_oetf(float e) {
  static const PqConst k;
  if (e <= 0.0f) return 0.0f;
  double p = std::pow((double)e, (double)k.m1);
  double ratio = ((double)k.c1 + (double)k.c2 * p) / (1 + (double)k.c3 * p);
  return (float)std::pow(ratio, (double)k.m2);
}
float pq_eotf(float v) {
  static const PqConst k;
  float val = (float)std::pow((double)v, (double)(1 / k.m2));
  float num = val - k.c1;
  if (num < 0.0f) num = 0.0f;
  return (float)std::pow((double)(num / (k.c2 - k.c3 * val)), (double)(1 / k.m1));
}

template <class F>
void fill(float* dst, int n, F fn) {
  for (int i = 0; i < n; i++) dst[i] = fn(static_cast<float>(i) / static_cast<float>(n - 1));
}
}  // namespace

void build_lut_blob(float* out) {
  fill(out + kLutSrgbInv, 1024, srgb_eotf);
  fill(out + kLutHlgInv, 4096, hlg_eotf);
  // hlgOotfApprox (gainmapmath.cpp:293-295) is a per-channel float std::pow(x, 1.2f) applied to
  // the output of the 4096-entry inverse-OETF LUT: fold it into a second table (host powf, i.e.
  // the very libm the reference CPU path uses on this machine).
  for (int i = 0; i < 4096; i++) out[kLutHlgInvOotf + i] = std::pow(out[kLutHlgInv + i], 1.2f);
  fill(out + kLutPqInv, 4096, pq_eotf);
  fill(out + kLutHlgOetf, 65536, hlg_oetf);
  fill(out + kLutPqOetf, 65536, pq_oetf);
  for (int i = 0; i < 256; i++) out[kLutU8Div255 + i] = static_cast<float>(i) / 255.0f;
}

// gainmapmath.cpp:39-80
static float idw_dist(float x1, float x2, float y1, float y2) {
  return (float)std::sqrt((double)(((y2 - y1) * (y2 - y1)) + (x2 - x1) * (x2 - x1)));
}
void build_idw_tables(int scale, float* out) {
  static const int inc[4][2] = {{1, 1}, {0, 1}, {1, 0}, {0, 0}};
  const size_t per = (size_t)scale * scale * 4;
  for (size_t i = 0; i < per * 4; i++) out[i] = 0.0f;
  for (int v = 0; v < 4; v++) {
    float* w = out + per * v;
    for (int y = 0; y < scale; y++)
      for (int x = 0; x < scale; x++) {
        float px = ((float)x) / scale, py = ((float)y) / scale;
        float cx = (float)(int)std::floor((double)px), cy = (float)(int)std::floor((double)py);
        float nx = cx + (float)inc[v][0], ny = cy + (float)inc[v][1];
        // next_x = curr_x + incR is an int add in the reference; values are 0/1 so exact
        float* o = w + ((size_t)y * scale + x) * 4;
        float d1 = idw_dist(px, cx, py, cy);
        if (d1 == 0) {
          o[0] = 1.f;
          continue;
        }
        float w1 = 1.f / d1;
        float w2 = 1.f / idw_dist(px, cx, py, ny);
        float w3 = 1.f / idw_dist(px, nx, py, cy);
        float w4 = 1.f / idw_dist(px, nx, py, ny);
        float tot = w1 + w2 + w3 + w4;
        o[0] = w1 / tot;
        o[1] = w2 / tot;
        o[2] = w3 / tot;
        o[3] = w4 / tot;
      }
  }
}

bool metadata_single_channel(const GainmapMetadata& m) {
  auto same = [](const float* a) { return a[0] == a[1] && a[0] == a[2]; };
  return same(m.max_content_boost) && same(m.min_content_boost) && same(m.gamma) &&
         same(m.offset_sdr) && same(m.offset_hdr);
}

// gainmapmath.h:452-470 (double log2/exp2: the header is parsed before any `using namespace std`)
void build_gain_lut(const GainmapMetadata& md, float weight, float* out) {
  const bool single = metadata_single_channel(md);
  for (int c = 0; c < (single ? 1 : 3); c++) {
    const double lmin = std::log2((double)md.min_content_boost[c]);
    const double lmax = std::log2((double)md.max_content_boost[c]);
    for (int i = 0; i < 1024; i++) {
      float value = static_cast<float>(i) / static_cast<float>(1023);
      float log_boost = (float)(lmin * (double)(1.0f - value) + lmax * (double)value);
      out[c * 1024 + i] = (float)std::exp2((double)(log_boost * weight));
    }
  }
  if (single) {
    std::memcpy(out + 1024, out, 1024 * sizeof(float));
    std::memcpy(out + 2048, out, 1024 * sizeof(float));
  }
}

// gainmapmath.cpp:603-615
static const float kGamutTbl[3][3][9] = {
    {{1, 0, 0, 0, 1, 0, 0, 0, 1},
     {1.22494f, -0.22494f, 0.0f, -0.042057f, 1.042057f, 0.0f, -0.019638f, -0.078636f, 1.098274f},
     {1.660491f, -0.587641f, -0.07285f, -0.124551f, 1.1329f, -0.008349f, -0.018151f, -0.100579f,
      1.11873f}},
    {{0.822462f, 0.177537f, 0.000001f, 0.033194f, 0.966807f, -0.000001f, 0.017083f, 0.072398f,
      0.91052f},
     {1, 0, 0, 0, 1, 0, 0, 0, 1},
     {1.343578f, -0.282179f, -0.061399f, -0.065298f, 1.075788f, -0.01049f, 0.002822f, -0.019598f,
      1.016777f}},
    {{0.627404f, 0.329282f, 0.043314f, 0.069097f, 0.919541f, 0.011362f, 0.016392f, 0.088013f,
      0.895595f},
     {0.753833f, 0.198597f, 0.04757f, 0.045744f, 0.941777f, 0.012479f, -0.00121f, 0.017601f,
      0.983608f},
     {1, 0, 0, 0, 1, 0, 0, 0, 1}}};

bool gamut_matrix(int dst, int src, float out[9], bool* identity) {
  if (dst < 0 || dst > 2 || src < 0 || src > 2) return false;
  std::memcpy(out, kGamutTbl[dst][src], sizeof(float) * 9);
  *identity = dst == src;
  return true;
}

// gainmapmath.cpp:638-674, [src][dst]
static const float kYuvTbl[3][3][9] = {
    {{1, 0, 0, 0, 1, 0, 0, 0, 1},
     {1.0f, 0.101579f, 0.196076f, 0.0f, 0.989854f, -0.110653f, 0.0f, -0.072453f, 0.983398f},
     {1.0f, -0.016969f, 0.096312f, 0.0f, 0.995306f, -0.051192f, 0.0f, 0.011507f, 1.002637f}},
    {{1.0f, -0.118188f, -0.212685f, 0.0f, 1.018640f, 0.114618f, 0.0f, 0.075049f, 1.025327f},
     {1, 0, 0, 0, 1, 0, 0, 0, 1},
     // third coefficient is a *double* literal narrowed to float in the reference (:660)
     {1.0f, -0.128245f, (float)-0.115879, 0.0f, 1.010016f, 0.061592f, 0.0f, 0.086969f, 1.029350f}},
    {{1.0f, 0.018149f, -0.095132f, 0.0f, 1.004123f, 0.051267f, 0.0f, -0.011524f, 0.996782f},
     {1.0f, 0.117887f, 0.105521f, 0.0f, 0.995211f, -0.059549f, 0.0f, -0.084085f, 0.976518f},
     {1, 0, 0, 0, 1, 0, 0, 0, 1}}};

bool yuv_matrix(int src, int dst, float out[9]) {
  if (dst < 0 || dst > 2 || src < 0 || src > 2) return false;
  std::memcpy(out, kYuvTbl[src][dst], sizeof(float) * 9);
  return true;
}

// gainmapmath.cpp:94,104-111,164,174-181,194,226-233: {cr, cb, gcb, gcr}
bool rgb2yuv_coeffs(int cg, float out[5]) {
  if (cg == 0) {
    out[0] = kSrgbR; out[1] = kSrgbG; out[2] = kSrgbB; out[3] = f2(kSrgbB); out[4] = f2(kSrgbR);
  } else if (cg == 1) {
    out[0] = kP3YR; out[1] = kP3YG; out[2] = kP3YB; out[3] = kP3Cb; out[4] = kP3Cr;
  } else if (cg == 2) {
    out[0] = kBtR; out[1] = kBtG; out[2] = kBtB; out[3] = f2(kBtB); out[4] = f2(kBtR);
  } else {
    return false;
  }
  return true;
}
bool yuv2rgb_coeffs(int cg, float out[4]) {
  if (cg == 0) {
    float cb = f2(kSrgbB), cr = f2(kSrgbR);
    out[0] = cr;
    out[1] = cb;
    out[2] = kSrgbB * cb / kSrgbG;
    out[3] = kSrgbR * cr / kSrgbG;
  } else if (cg == 1) {
    out[0] = kP3Cr;
    out[1] = kP3Cb;
    out[2] = kP3YB * kP3Cb / kP3YG;
    out[3] = kP3YR * kP3Cr / kP3YG;
  } else if (cg == 2) {
    float cb = f2(kBtB), cr = f2(kBtR);
    out[0] = cr;
    out[1] = cb;
    out[2] = kBtB * cb / kBtG;
    out[3] = kBtR * cr / kBtG;
  } else {
    return false;
  }
  return true;
}

bool luminance_coeffs(int cg, float out[3]) {
  if (cg == 0) { out[0] = kSrgbR; out[1] = kSrgbG; out[2] = kSrgbB; }
  else if (cg == 1) { out[0] = kP3R; out[1] = kP3G; out[2] = kP3B; }
  else if (cg == 2) { out[0] = kBtR; out[1] = kBtG; out[2] = kBtB; }
  else return false;
  return true;
}

// gainmapmath.cpp:20-34
float reference_display_peak_nits(int ct) {
  switch (ct) {
    case 0: return 10000.0f;  // LINEAR
    case 1: return 1000.0f;   // HLG
    case 2: return 10000.0f;  // PQ
    case 3: return 203.0f;    // SRGB
  }
  return -1.0f;
}

}  // namespace uhdr_b200
