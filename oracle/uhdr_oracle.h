/*
 * uhdr_oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement (plain C, scalar) of the reference's per-pixel gain-map path:
 *   lib/src/gainmapmath.cpp, lib/include/ultrahdr/gainmapmath.h (primitives, LUTs, samplers)
 *   lib/src/jpegr.cpp:436-518 (convertYuv), :530-1058 (generateGainMap), :1533-1831
 *   (applyGainMap), :1945-2222 (toneMap).
 * Build with the flags in oracle/Makefile (x86-64 baseline, no FMA contraction): the float /
 * double promotion at every site follows the "precision map" of SURVEY.md section 8.
 * Pinned against oracle/_ref (the reference's own sources compiled in place) in
 * tests/test_oracle_vs_ref.py and against the reference's known-answer vectors
 * (tests/gainmapmath_test.cpp) in tests/test_oracle_kat.py.
 */
#ifndef UHDR_ORACLE_H
#define UHDR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout-compatible with uhdr_raw_image_t (ultrahdr_api.h:227-246) */
typedef struct {
  int fmt, cg, ct, range;
  unsigned w, h;
  void* planes[3];
  unsigned stride[3];
} uo_image_t;

/* layout-compatible with uhdr_gainmap_metadata_t (ultrahdr_api.h:262-283) */
typedef struct {
  float max_content_boost[3], min_content_boost[3], gamma[3], offset_sdr[3], offset_hdr[3];
  float hdr_capacity_min, hdr_capacity_max;
  int use_base_cg;
} uo_metadata_t;

/* JpegR ctor arguments + generateGainMap flags */
typedef struct {
  int scale_factor, quality, multichannel;
  float gamma;
  int preset;
  float min_content_boost, max_content_boost, target_disp_peak_nits;
  int sdr_is_601, use_luminance;
} uo_gm_config_t;

/* LUTs: which = 0 srgbInvOetf(1024) 1 hlgInvOetf(4096) 2 pqInvOetf(4096) 3 hlgOetf(65536)
 * 4 pqOetf(65536) */
int uo_lut(int which, float* out, int n);
void uo_idw_weights(int scale, int variant, float* out);
void uo_gain_lut(const uo_metadata_t* md, float weight, float* out /* 3*1024 */);

float uo_srgb_oetf(float x);
float uo_compute_gain(float sdr, float hdr);
int uo_affine_map_gain(float g, float mn, float mx, float gamma);
int uo_encode_gain(float y_sdr, float y_hdr, const uo_metadata_t* md, float l2min, float l2max,
                   int idx);
unsigned uo_float_to_half(float f);
void uo_log2_of_float(const float* in, float* out, size_t n);
void uo_powf_vec(const float* in, float y, float* out, size_t n);

int uo_generate_gainmap(const uo_image_t* sdr, const uo_image_t* hdr, const uo_gm_config_t* cfg,
                        uo_metadata_t* md_out, uo_image_t* gainmap_out /* tight, caller mem */);
int uo_apply_gainmap(const uo_image_t* sdr, const uo_image_t* gainmap, const uo_metadata_t* md,
                     int output_ct, int output_fmt, float max_display_boost, uo_image_t* dest);
int uo_tonemap(const uo_image_t* hdr, uo_image_t* sdr);
int uo_convert_yuv(uo_image_t* img, int src_cg, int dst_cg);

#ifdef __cplusplus
}
#endif
#endif
