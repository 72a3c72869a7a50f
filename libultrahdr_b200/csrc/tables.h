// Host-side constant tables of the gain-map path: transfer-function LUTs, Shepard IDW weights,
// gain LUT, gamut / YUV matrices.  These are built on the HOST with the same libm calls and the
// same float/double promotion as the reference (lib/src/gainmapmath.cpp:114-349,
// lib/include/ultrahdr/gainmapmath.h:222-251,345-357,452-489) and uploaded once per device, so
// that every transcendental on the device hot path is a table fetch and the remaining per-pixel
// arithmetic is IEEE binary32 + - * / (compiled with -fmad=false).
#pragma once
#include <cstdint>
#include <vector>

namespace uhdr_b200 {

// offsets (in floats) of each table inside the LUT blob that lives in device memory and that is
// broadcast rank0 -> all ranks with one NCCL broadcast in multi-GPU runs
enum : int {
  kLutSrgbInv = 0,                       // 1024  srgbInvOetf
  kLutHlgInv = kLutSrgbInv + 1024,       // 4096  hlgInvOetf
  kLutHlgInvOotf = kLutHlgInv + 4096,    // 4096  powf(hlgInvOetf, 1.2f)   (hlgOotfApprox folded in)
  kLutPqInv = kLutHlgInvOotf + 4096,     // 4096  pqInvOetf
  kLutHlgOetf = kLutPqInv + 4096,        // 65536 hlgOetf
  kLutPqOetf = kLutHlgOetf + 65536,      // 65536 pqOetf
  kLutU8Div255 = kLutPqOetf + 65536,     // 256   float(i) / 255.0f   (mapUintToFloat)
  kLutTotalFloats = kLutU8Div255 + 256
};

// Fills kLutTotalFloats floats.
void build_lut_blob(float* out);

// ShepardsIDW: 4 variants (default, no-right, no-bottom, corner) x scale*scale*4 floats.
void build_idw_tables(int scale, float* out);   // 16 * scale * scale floats, caller storage

struct GainmapMetadata {  // == uhdr_gainmap_metadata_t
  float max_content_boost[3], min_content_boost[3], gamma[3], offset_sdr[3], offset_hdr[3];
  float hdr_capacity_min, hdr_capacity_max;
  int use_base_cg;
};
bool metadata_single_channel(const GainmapMetadata& m);

// GainLUT: 3 x 1024 floats (channels aliased when the metadata is single channel)
void build_gain_lut(const GainmapMetadata& md, float weight, float* out);

// gamut conversion dst <- src (identity when equal); returns false for unknown gamuts
bool gamut_matrix(int dst_cg, int src_cg, float out[9], bool* identity);
// yuv encoding change src -> dst
bool yuv_matrix(int src_cg, int dst_cg, float out[9]);
// yuv->rgb coefficients {cr, cb, gcb, gcr} of a gamut
bool yuv2rgb_coeffs(int cg, float out[4]);
// {yr, yg, yb, cb, cr} of srgb/p3/bt2100RgbToYuv (gainmapmath.cpp:96-99,166-169,196-199)
bool rgb2yuv_coeffs(int cg, float out[5]);
// luminance coefficients
bool luminance_coeffs(int cg, float out[3]);
float reference_display_peak_nits(int ct);

}  // namespace uhdr_b200
