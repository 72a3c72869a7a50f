// generateGainMap fast path (jpegr.cpp:753-817 one-pass, :866-931 pass 1 of two-pass) for P010 HDR intent
// (HLG or PQ) + YUV 4:2:0 SDR intent: map scale 1 (the configuration the API-0/API-1 benchmarks exercise,
// k_gainmap_fast) and map scales 2 / 4 (JpegR's own default is 4, k_gainmap_scaled further down).  Arithmetic, operand order and tables are those of the generic kernels in
// kernels.cu; what changes is the instruction count:
//   * persistent CTAs, 256x8-pixel tiles handed out through an atomic ticket; one thread = a 4x2
//     pixel tile (chroma terms of both images computed once per 2x2)
//   * every fp32 multiply / add on packed pairs (two horizontally adjacent pixels, packed_f32.cuh)
//   * inverse-OETF tables in shared memory in "doubled" form so that the reference's LUT index
//     int32(double(x*(N-1)) + 0.5) becomes one multiply, one add toward zero (the index is read out
//     of the mantissa) and one mask
//   * computeGain's double-precision log2 of a float quotient through a 128-entry table + degree-8
//     polynomial in fp64 (error < 2^-50 before narrowing to float, like glibc's / CUDA's log2,
//     ~6x fewer instructions than the library routine), coefficients as constant-bank operands,
//     widenings done with integer ops; the IEEE division without its range-check slow path
// Also here: the affine pass (min/max finalisation folded in) and the 4:2:0 convertYuv kernel.
#include <cmath>
#include <mutex>

#include "kernels.cuh"
#include "packed_f32.cuh"
#include "powf_glibc.cuh"
#include "runtime.h"
#include "tables.h"

namespace uhdr_b200 {

namespace {

// ---- log2 ---------------------------------------------------------------------------------------
// z = 2^k * m, m in [OFF, 2*OFF), OFF = 0x3f330000 (0.69921875); m's top 7 mantissa bits select
// (invc, logc) with c near the centre of the sub-interval; the two sub-intervals touching 1.0 use
// c = 1 exactly so that results near zero keep full relative accuracy.
struct Log2Tab { double invc[128], logc[128]; };
constexpr unsigned kLogOff = 0x3f330000u;

// polynomial of log2(1+r)/r, highest degree first; in the constant bank so that every DFMA takes its
// coefficient as a c[][] operand instead of materialising it with two moves
__constant__ double kLog2Poly[8] = {
    -0.18033688011112042,  // -1/(8 ln2)
    0.20609929155556619,   //  1/(7 ln2)
    -0.24044917348149390,  // -1/(6 ln2)
    0.28853900817779268,   //  1/(5 ln2)
    -0.36067376022224085,  // -1/(4 ln2)
    0.48089834696298783,   //  1/(3 ln2)
    -0.72134752044448170,  // -1/(2 ln2)
    1.4426950408889634};   //  1/ln2

__device__ __forceinline__ double log2_core(float q, const double2* __restrict__ tab /* smem: {invc, logc}[128] */) {
  const unsigned ix = __float_as_uint(q);
  const unsigned tmp = ix - kLogOff;
  const int i = (tmp >> 16) & 127;
  const int k = (int)tmp >> 23;
  const unsigned im = ix - (tmp & 0xff800000u);  // bits of m (a normal float in [OFF, 2*OFF))
  // exact widenings done with integer ops (the conversion unit is the busiest pipe of this kernel):
  // m: re-bias the exponent, shift the mantissa;  k: 2^52 + 2^31 + k minus the same constant
  const double md = __hiloint2double((int)((im >> 3) + 0x38000000u), (int)(im << 29));
  const double kd = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)k)) - 4503601774854144.0;
  const double2 t = tab[i];
  const double r = fma(md, t.x, -1.0);
  // log2(1+r) = r * P(r), P = sum_{j>=0} (-1)^j r^j / ((j+1) ln2)
  double p = kLog2Poly[0];
#pragma unroll
  for (int j = 1; j < 8; j++) p = fma(p, r, kLog2Poly[j]);
  return fma(p, r, kd + t.y);
}


// encodeGain's normalisation (gainmapmath.cpp:766): float((log2(gain) - log2_min) / double(log2_max - log2_min)), an fp64
// division per channel and pixel in the reference.  The divisor b is the same for every pixel and is a float widened
// to double, so the quotient comes out of three fp64 operations with y = RN(1/b) from the host (IEEE division):
//   q = RN(a*y);  r = a - b*q (exact in the fma);  q' = RN(q + r*y)
// q is within one ulp of a/b, so q + r*y differs from a/b by less than 2^-51 ulp before rounding; with a 24-bit
// divisor a/b stays at least 2^-25 ulp away from every rounding boundary of a double (|A*2^k - B*(2n+1)| >= 1 for
// integers A < 2^53, B < 2^24), hence q' = RN(a/b): the same double the division returns, bit for bit.  (No underflow:
// a is exactly 0 or at least 2^-52 * |log2_min| resp. log2(1 + 2^-23) in magnitude.  tests/test_exact_quotient_cpu.py
// runs the three operations against the hardware division on a few million operand pairs.)
__device__ __forceinline__ float encode_gain_norm(double log2_gain, const GainmapGenParams& p) {
  const double a = log2_gain - (double)p.log2_min;
  const double b = (double)(p.log2_max - p.log2_min);
  if (p.inv_log2_range == 0.0) return (float)(a / b);   // degenerate range: the reference's own division (inf / NaN)
  const double q = a * p.inv_log2_range;
  const double r = fma(-q, b, a);
  return (float)fma(r, p.inv_log2_range, q);
}

// q mode of the two-pass kernels: key slots and start values (a quotient is a positive finite float)
constexpr float kQMinInit = 3.0e38f;
constexpr int kQDarkKeys = 16;   // minmax[16..18] min q of dark pixels, [19..21] max q of dark pixels ([0..5]: the others)

constexpr float kLg2Abs = 8e-7f, kLg2Rel = 3.2e-7f, kAffineRound = 2e-4f;   // measured worst case: 0.35 of this bound

__device__ __forceinline__ float lg2_fast(float x) {
  float r;
  asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_fast_bound(float g) { return kLg2Abs + fabsf(g) * kLg2Rel; }

// one-pass code of a gain: trunc(encode_gain_norm(log2(gain)) * 255).  Only the byte leaves the kernel: it is taken from
// lg2.approx and float arithmetic unless t lies within `thr` of an integer, thr bounding everything the short cut can
// differ by (lg2_fast_bound scaled by 255 / range, plus 1.5e-4 for the float roundings on either side); then the fp64
// path decides.  thr <= 0: always exact (degenerate range).
// Gains clamped to min_boost / max_boost (every pixel that is not brighter in the HDR rendition, for one) sit exactly
// on t = 0 / 255: their codes come from two exact evaluations per thread (code_lo, code_hi), not from the screen.
__device__ __forceinline__ unsigned encode_gain_code(float gain, const GainmapGenParams& p, float inv_range_f, float thr,
                                                     unsigned code_lo, unsigned code_hi, const double2* __restrict__ tab) {
  if (gain <= p.min_boost) return code_lo;
  if (gain >= p.max_boost) return code_hi;
  float t = 0.0f;
  bool exact = !(thr > 0.0f);
  if (!exact) {
    t = ((lg2_fast(gain) - p.log2_min) * inv_range_f) * 255.0f;
    exact = fabsf(t - rintf(t)) < thr;
  }
  if (exact) t = encode_gain_norm(log2_core(gain, tab), p) * 255.0f;
  return (unsigned)__float2int_rz(t) & 0xff;
}
__device__ __forceinline__ float onepass_threshold(const GainmapGenParams& p, float inv_range_f) {
  if (p.inv_log2_range == 0.0) return -1.0f;
  const float gmax = fmaxf(fabsf(p.log2_min), fabsf(p.log2_max));
  return 255.0f * (kLg2Abs + gmax * kLg2Rel) * fabsf(inv_range_f) * 1.0001f + 1.5e-4f;   // six float roundings of 6e-8 on t <= 255: 9.2e-5
}

// ---- shared memory ------------------------------------------------------------------------------
struct GmSmem {
  double2 log2tab[128];  // {invc, logc}
  float srgb2[2048];   // srgb2[j] = srgbInvLUT[(j+1)>>1]
  float hdr2[8192];    // hdr2[j]  = hdrInvLUT[(j+1)>>1], 4096-entry source table
};
__device__ __forceinline__ float fetch2(const float* t, float x, float scale8) {  // x in [0,1]
  // trunc(x*scale8) without the conversion unit: adding 2^23 toward zero leaves it in the mantissa
  const int off = __float_as_int(__fadd_rz(x * scale8, 8388608.0f)) & 0x7ffffc;
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(t) + off);
}

// a / b per lane, same steps as div_pos (na = -b is formed by the caller's packed multiply by -1)
__device__ __forceinline__ V2 div_pos2(V2 a, V2 b, unsigned long long nz) {
  float b0, b1, r0, r1;
  un(b, b0, b1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(b1));
  const V2 nb = vmul(b, bc(-1.0f), nz);
  V2 r = v2(r0, r1);
  const V2 e = vfma(nb, r, bc(1.0f));
  r = vfma(r, e, r);
  const V2 q = vmul(a, r, nz);
  const V2 rem = vfma(nb, q, a);
  return vfma(r, rem, q);
}

__device__ __forceinline__ float fetch_off(const float* t, unsigned mant) {  // mant: mantissa bits holding trunc(x*8(N-1))
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(t) + (mant & 0x7ffffc));
}

// order-independent min / max reduction of a CTA into the pass-1 keys (same as k_gainmap_pass1)
template <int NCH>
__device__ __forceinline__ void reduce_minmax(const float mn[3], const float mx[3], unsigned* __restrict__ minmax, int tid, int nt) {
  __shared__ unsigned s_mn[3][8], s_mx[3][8];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    unsigned a = __float_as_uint(mn[c]), b = __float_as_uint(mx[c]);
    a = (a & 0x80000000u) ? ~a : (a | 0x80000000u);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    for (int o = 16; o; o >>= 1) {
      a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) { s_mn[c][warp] = a; s_mx[c][warp] = b; }
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = nt >> 5;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      unsigned a = lane < nw ? s_mn[c][lane] : 0xffffffffu, b = lane < nw ? s_mx[c][lane] : 0u;
      for (int o = 4; o; o >>= 1) {
        a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      if (lane == 0) { atomicMin(minmax + c, a); atomicMax(minmax + 3 + c, b); }
    }
  }
}

template <bool ONEPASS, int NCH, int GAMUT /*0 none, 1 on sdr, 2 on hdr*/, bool LIMITED, bool QMODE /*two-pass: quotient plane*/>
__global__ void __launch_bounds__(256, 4) k_gainmap_fast(const GainmapGenParams p, const double* __restrict__ log2tab_g, const int tiles_x,
                                                         const int ntiles, unsigned* __restrict__ sched, const unsigned long long nz) {
  extern __shared__ double2 smem_d[];
  GmSmem& sm = *reinterpret_cast<GmSmem*>(smem_d);
  __shared__ int s_tile[2];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  for (int i = tid; i < 128; i += nt) sm.log2tab[i] = make_double2(log2tab_g[i], log2tab_g[128 + i]);
  for (int i = tid; i < 2048; i += nt) sm.srgb2[i] = __ldg(p.luts + kLutSrgbInv + min((i + 1) >> 1, 1023));
  // hdr inverse OETF table: 4096 entries for HLG (OOTF folded in) / PQ, the 1024-entry sRGB one for an sRGB
  // "hdr" intent (reachable through JpegR::generateGainMap, getInverseOetfFn gainmapmath.cpp:1175-1180)
  const int hN = p.hdr_ct == CT_SRGB ? 1024 : 4096;
  const float* hsrc = p.luts + (p.hdr_ct == CT_HLG ? kLutHlgInvOotf : (p.hdr_ct == CT_PQ ? kLutPqInv : kLutSrgbInv));
  for (int i = tid; i < 2 * hN; i += nt) sm.hdr2[i] = __ldg(hsrc + min((i + 1) >> 1, hN - 1));
  const float hscale8 = (float)(8 * (hN - 1));
  // persistent CTAs; 256x8-pixel tiles handed out through a ticket counter (zeroed by the caller)
  if (tid == 0) s_tile[0] = (int)atomicAdd(sched, 1u);
  __syncthreads();

  // gains mode: extremes of the gains; q mode (store_q): extremes of the quotient, non-dark and dark pixels apart
  const float mn0 = QMODE ? kQMinInit : 127.0f, mx0 = QMODE ? 0.0f : -128.0f;
  float mn[3] = {mn0, mn0, mn0}, mx[3] = {mx0, mx0, mx0};
  float dmn[3] = {kQMinInit, kQMinInit, kQMinInit}, dmx[3] = {0.0f, 0.0f, 0.0f};
  const V2 k8184 = bc(8184.0f), k32760 = bc(hscale8), keps = bc(1e-7f);
  const V2 snits = bc(p.sdr_nits), hnits = bc(p.hdr_nits);
  const float inv_range_f = (float)p.inv_log2_range, thr1 = onepass_threshold(p, inv_range_f);   // one-pass screen
  unsigned code_lo = 0, code_hi = 0;
  if (ONEPASS) {   // exact codes of the two clamp values (tables are staged: see the barrier above)
    code_lo = (unsigned)__float2int_rz(encode_gain_norm(log2_core(p.min_boost, sm.log2tab), p) * 255.0f) & 0xff;
    code_hi = (unsigned)__float2int_rz(encode_gain_norm(log2_core(p.max_boost, sm.log2tab), p) * 255.0f) & 0xff;
  }
#pragma unroll 1
  for (int it = 0;; it++) {
    const int t = s_tile[it & 1];
    if (t >= ntiles) break;
    if (tid == 0) s_tile[(it + 1) & 1] = (int)atomicAdd(sched, 1u);
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x = (tx * 64 + threadIdx.x) * 4;
    const int y = (ty * 4 + threadIdx.y) * 2;
    if (x < p.map_w && y < p.map_h) {
      // ---- loads: 4x2 luma of both images, 2 chroma pairs each
      const uint16_t* HY = (const uint16_t*)p.hdr.p[0];
      const uint2 hy0 = __ldg((const uint2*)(HY + (size_t)y * p.hdr.stride[0] + x));
      const uint2 hy1 = __ldg((const uint2*)(HY + (size_t)(y + 1) * p.hdr.stride[0] + x));
      const uint2 huv = __ldg((const uint2*)((const uint16_t*)p.hdr.p[1] + (size_t)(y >> 1) * p.hdr.stride[1] + x));
      const uint8_t* SY = (const uint8_t*)p.sdr.p[0];
      const unsigned sy0 = __ldg((const unsigned*)(SY + (size_t)y * p.sdr.stride[0] + x));
      const unsigned sy1 = __ldg((const unsigned*)(SY + (size_t)(y + 1) * p.sdr.stride[0] + x));
      const unsigned su = __ldg((const uint16_t*)((const uint8_t*)p.sdr.p[1] + (size_t)(y >> 1) * p.sdr.stride[1] + (x >> 1)));
      const unsigned sv = __ldg((const uint16_t*)((const uint8_t*)p.sdr.p[2] + (size_t)(y >> 1) * p.sdr.stride[2] + (x >> 1)));
      // ---- chroma terms (yuv->rgb: r = y + cr*v, g = y - gcb*u - gcr*v, b = y + cb*u)
      float h_crv[2], h_cbu[2], h_gcbu[2], h_gcrv[2], s_crv[2], s_cbu[2], s_gcbu[2], s_gcrv[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const unsigned uvw = k ? huv.y : huv.x;
        const int u10 = (int)((uvw & 0xffff) >> 6), v10 = (int)(uvw >> 22);
        float hu, hv;
        if (LIMITED) {
          hu = (float)(u10 - 64) * (1 / 896.0f) - 0.5f;
          hv = (float)(v10 - 64) * (1 / 896.0f) - 0.5f;
        } else {
          hu = (float)u10 / 1023.0f - 0.5f;
          hv = (float)v10 / 1023.0f - 0.5f;
        }
        h_crv[k] = p.hdr_y2r[0] * hv; h_cbu[k] = p.hdr_y2r[1] * hu;
        h_gcbu[k] = p.hdr_y2r[2] * hu; h_gcrv[k] = p.hdr_y2r[3] * hv;
        const float u = (float)((int)((su >> (8 * k)) & 0xff) - 128) * (1 / 255.0f);
        const float v = (float)((int)((sv >> (8 * k)) & 0xff) - 128) * (1 / 255.0f);
        s_crv[k] = p.sdr_y2r[0] * v; s_cbu[k] = p.sdr_y2r[1] * u;
        s_gcbu[k] = p.sdr_y2r[2] * u; s_gcrv[k] = p.sdr_y2r[3] * v;
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const uint2 hyw = r ? hy1 : hy0;
        const unsigned syw = r ? sy1 : sy0;
        float gout[12];
        unsigned bout[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 2; k++) {  // pixel pair (2k, 2k+1): same chroma sample, packed arithmetic
          // sdr: getYuv420Pixel -> yuvToRgb -> srgbInvOetfLUT [-> gamut] -> clipNegatives
          const V2 syf = vmul(v2((float)((syw >> (16 * k)) & 0xff), (float)((syw >> (16 * k + 8)) & 0xff)), bc(1 / 255.0f), nz);
          float a0, a1, t0, t1;
          un(syf, a0, a1);
          un(vsub(syf, bc(s_gcbu[k])), t0, t1);
          unsigned i0, i1, j0, j1, l0, l1;
          un(vtrunc_bits(vmul(v2(__saturatef(a0 + s_crv[k]), __saturatef(a1 + s_crv[k])), k8184, nz)), i0, i1);
          un(vtrunc_bits(vmul(v2(__saturatef(t0 - s_gcrv[k]), __saturatef(t1 - s_gcrv[k])), k8184, nz)), j0, j1);
          un(vtrunc_bits(vmul(v2(__saturatef(a0 + s_cbu[k]), __saturatef(a1 + s_cbu[k])), k8184, nz)), l0, l1);
          V2 sr = v2(fetch_off(sm.srgb2, i0), fetch_off(sm.srgb2, i1));
          V2 sg = v2(fetch_off(sm.srgb2, j0), fetch_off(sm.srgb2, j1));
          V2 sb = v2(fetch_off(sm.srgb2, l0), fetch_off(sm.srgb2, l1));
          // hdr: getP010Pixel -> yuvToRgb -> invOETF(+OOTF) LUT [-> gamut] -> clipNegatives
          const unsigned hw = k ? hyw.y : hyw.x;
          const int ya = (int)((hw & 0xffff) >> 6), yb = (int)(hw >> 22);
          V2 hyf;
          if (LIMITED) hyf = vmul(v2((float)(ya - 64), (float)(yb - 64)), bc(1 / 876.0f), nz);
          else hyf = v2((float)ya / 1023.0f, (float)yb / 1023.0f);
          un(hyf, a0, a1);
          un(vsub(hyf, bc(h_gcbu[k])), t0, t1);
          un(vtrunc_bits(vmul(v2(__saturatef(a0 + h_crv[k]), __saturatef(a1 + h_crv[k])), k32760, nz)), i0, i1);
          un(vtrunc_bits(vmul(v2(__saturatef(t0 - h_gcrv[k]), __saturatef(t1 - h_gcrv[k])), k32760, nz)), j0, j1);
          un(vtrunc_bits(vmul(v2(__saturatef(a0 + h_cbu[k]), __saturatef(a1 + h_cbu[k])), k32760, nz)), l0, l1);
          V2 hr = v2(fetch_off(sm.hdr2, i0), fetch_off(sm.hdr2, i1));
          V2 hg = v2(fetch_off(sm.hdr2, j0), fetch_off(sm.hdr2, j1));
          V2 hb = v2(fetch_off(sm.hdr2, l0), fetch_off(sm.hdr2, l1));
          if (GAMUT != 0) {
            V2& xr = GAMUT == 1 ? sr : hr;
            V2& xg = GAMUT == 1 ? sg : hg;
            V2& xb = GAMUT == 1 ? sb : hb;
            const V2 a = vadd(vadd(vmul(bc(p.gamut[0]), xr, nz), vmul(bc(p.gamut[1]), xg, nz)), vmul(bc(p.gamut[2]), xb, nz));
            const V2 b = vadd(vadd(vmul(bc(p.gamut[3]), xr, nz), vmul(bc(p.gamut[4]), xg, nz)), vmul(bc(p.gamut[5]), xb, nz));
            const V2 c = vadd(vadd(vmul(bc(p.gamut[6]), xr, nz), vmul(bc(p.gamut[7]), xg, nz)), vmul(bc(p.gamut[8]), xb, nz));
            float f0, f1;
            un(a, f0, f1); xr = v2(fmaxf(f0, 0.0f), fmaxf(f1, 0.0f));
            un(b, f0, f1); xg = v2(fmaxf(f0, 0.0f), fmaxf(f1, 0.0f));
            un(c, f0, f1); xb = v2(fmaxf(f0, 0.0f), fmaxf(f1, 0.0f));
          }
          V2 sv3[3], hv3[3];
          if (NCH == 3) {
            sv3[0] = vmul(sr, snits, nz); sv3[1] = vmul(sg, snits, nz); sv3[2] = vmul(sb, snits, nz);
            hv3[0] = vmul(hr, hnits, nz); hv3[1] = vmul(hg, hnits, nz); hv3[2] = vmul(hb, hnits, nz);
          } else if (p.use_luminance) {
            sv3[0] = vmul(vadd(vadd(vmul(bc(p.lum[0]), sr, nz), vmul(bc(p.lum[1]), sg, nz)), vmul(bc(p.lum[2]), sb, nz)), snits, nz);
            hv3[0] = vmul(vadd(vadd(vmul(bc(p.lum[0]), hr, nz), vmul(bc(p.lum[1]), hg, nz)), vmul(bc(p.lum[2]), hb, nz)), hnits, nz);
          } else {
            float r0, r1, g0, g1, b0, b1;
            un(sr, r0, r1); un(sg, g0, g1); un(sb, b0, b1);
            sv3[0] = vmul(v2(fmaxf(r0, fmaxf(g0, b0)), fmaxf(r1, fmaxf(g1, b1))), snits, nz);
            un(hr, r0, r1); un(hg, g0, g1); un(hb, b0, b1);
            hv3[0] = vmul(v2(fmaxf(r0, fmaxf(g0, b0)), fmaxf(r1, fmaxf(g1, b1))), hnits, nz);
          }
#pragma unroll
          for (int c = 0; c < NCH; c++) {
            float s0, s1, q0, q1;
            un(sv3[c], s0, s1);
            if (ONEPASS) {  // encodeGain gainmapmath.cpp:758-771 (gamma 1: powf(x, 1) == x)
              float h0, h1;
              un(hv3[c], h0, h1);
              const float hv2[2] = {h0, h1}, sv2[2] = {s0, s1};
#pragma unroll
              for (int e = 0; e < 2; e++) {
                float gain = 1.0f;
                if (sv2[e] > 0.0f) gain = div_pos(hv2[e], sv2[e]);
                if (gain < p.min_boost) gain = p.min_boost;
                if (gain > p.max_boost) gain = p.max_boost;
                const unsigned code = encode_gain_code(gain, p, inv_range_f, thr1, code_lo, code_hi, sm.log2tab);
                const int bi = (2 * k + e) * NCH + c;
                bout[bi >> 2] |= code << (8 * (bi & 3));
              }
            } else {        // computeGain :773-782
              un(div_pos2(vadd(hv3[c], keps), vadd(sv3[c], keps), nz), q0, q1);
              if (QMODE) {
                // the log2 waits for pass 2 (k_affine_q): log2, its narrowing to float and the dark-pixel cap are
                // monotone, so the extremes of the gains are the images of the extremes of q per class
                const bool d0 = s0 < 2.f / 255.0f, d1 = s1 < 2.f / 255.0f;
                gout[(2 * k) * NCH + c] = d0 ? -q0 : q0;
                gout[(2 * k + 1) * NCH + c] = d1 ? -q1 : q1;
                if (d0) { dmn[c] = fminf(dmn[c], q0); dmx[c] = fmaxf(dmx[c], q0); } else { mn[c] = fminf(mn[c], q0); mx[c] = fmaxf(mx[c], q0); }
                if (d1) { dmn[c] = fminf(dmn[c], q1); dmx[c] = fmaxf(dmx[c], q1); } else { mn[c] = fminf(mn[c], q1); mx[c] = fmaxf(mx[c], q1); }
              } else {
                float g0 = (float)log2_core(q0, sm.log2tab), g1 = (float)log2_core(q1, sm.log2tab);
                if (s0 < 2.f / 255.0f) g0 = fminf(g0, 2.3f);
                if (s1 < 2.f / 255.0f) g1 = fminf(g1, 2.3f);
                gout[(2 * k) * NCH + c] = g0;
                gout[(2 * k + 1) * NCH + c] = g1;
                mn[c] = fminf(mn[c], fminf(g0, g1));
                mx[c] = fmaxf(mx[c], fmaxf(g0, g1));
              }
            }
          }
        }
        const int yy = y + r;
        if (ONEPASS) {
          uint8_t* d = p.dst + ((size_t)yy * p.dst_stride + x) * NCH;
          if (NCH == 3) { ((unsigned*)d)[0] = bout[0]; ((unsigned*)d)[1] = bout[1]; ((unsigned*)d)[2] = bout[2]; }
          else *(unsigned*)d = bout[0];
        } else {
          float4* d = (float4*)(p.gains + ((size_t)yy * p.map_w + x) * NCH);
          if (NCH == 3) {
            d[0] = make_float4(gout[0], gout[1], gout[2], gout[3]);
            d[1] = make_float4(gout[4], gout[5], gout[6], gout[7]);
            d[2] = make_float4(gout[8], gout[9], gout[10], gout[11]);
          } else {
            d[0] = make_float4(gout[0], gout[1], gout[2], gout[3]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!ONEPASS) {
    reduce_minmax<NCH>(mn, mx, p.minmax, tid, nt);
    if (QMODE) {
      __syncthreads();
      reduce_minmax<NCH>(dmn, dmx, p.minmax + kQDarkKeys, tid, nt);
    }
  }
}

// ---- map scale 2 / 4 (the reference's JpegR default is scale 4, one channel, ultrahdrcommon.h:450-457) ----
// samplePixels (gainmapmath.cpp:494-504): the S x S source pixels of a map pixel are fetched as normalised
// YUV floats, summed in raster order (three sequential chains, starting from 0) and divided by S*S; the rest
// of the pixel is the scale-1 arithmetic once per map pixel.  The sampling is the work here (16 source
// pixels of each image per map pixel at S = 4): one thread = one map pixel, each source row arrives with
// one load per plane (8 / 8 / 4 / 2 / 2 bytes at S = 4), the adds stay in the reference's order.
// read-only loads that stay where they are written (ptxas keeps volatile instructions in program order): the scaled
// kernel wants every source row of a thread in flight before the first dependent instruction
__device__ __forceinline__ unsigned long long ldv_u64(const void* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ldv_u32(const void* p) {
  unsigned v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ldv_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ldv_u8(const void* p) {
  unsigned v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <bool ONEPASS, int NCH, int GAMUT, bool LIMITED, int S, bool QMODE>
__global__ void __launch_bounds__(256, 4) k_gainmap_scaled(const GainmapGenParams p, const double* __restrict__ log2tab_g) {
  extern __shared__ double2 smem_d[];
  GmSmem& sm = *reinterpret_cast<GmSmem*>(smem_d);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  for (int i = tid; i < 128; i += nt) sm.log2tab[i] = make_double2(log2tab_g[i], log2tab_g[128 + i]);
  for (int i = tid; i < 2048; i += nt) sm.srgb2[i] = __ldg(p.luts + kLutSrgbInv + min((i + 1) >> 1, 1023));
  // hdr inverse OETF table: 4096 entries for HLG (OOTF folded in) / PQ, the 1024-entry sRGB one for an sRGB
  // "hdr" intent (reachable through JpegR::generateGainMap, getInverseOetfFn gainmapmath.cpp:1175-1180)
  const int hN = p.hdr_ct == CT_SRGB ? 1024 : 4096;
  const float* hsrc = p.luts + (p.hdr_ct == CT_HLG ? kLutHlgInvOotf : (p.hdr_ct == CT_PQ ? kLutPqInv : kLutSrgbInv));
  for (int i = tid; i < 2 * hN; i += nt) sm.hdr2[i] = __ldg(hsrc + min((i + 1) >> 1, hN - 1));
  const float hscale8 = (float)(8 * (hN - 1));
  __syncthreads();
  // gains mode: extremes of the gains; q mode (store_q): extremes of the quotient, non-dark and dark pixels apart
  const float mn0 = QMODE ? kQMinInit : 127.0f, mx0 = QMODE ? 0.0f : -128.0f;
  float mn[3] = {mn0, mn0, mn0}, mx[3] = {mx0, mx0, mx0};
  float dmn[3] = {kQMinInit, kQMinInit, kQMinInit}, dmx[3] = {0.0f, 0.0f, 0.0f};
  const int tiles_x = (p.map_w + 63) / 64, ntiles = tiles_x * ((p.map_h + 3) / 4);
  const float inv_range_f = (float)p.inv_log2_range, thr1 = onepass_threshold(p, inv_range_f);   // one-pass screen
  unsigned code_lo = 0, code_hi = 0;
  if (ONEPASS) {   // exact codes of the two clamp values (tables are staged: see the barrier above)
    code_lo = (unsigned)__float2int_rz(encode_gain_norm(log2_core(p.min_boost, sm.log2tab), p) * 255.0f) & 0xff;
    code_hi = (unsigned)__float2int_rz(encode_gain_norm(log2_core(p.max_boost, sm.log2tab), p) * 255.0f) & 0xff;
  }
#pragma unroll 1
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x = tx * 64 + threadIdx.x, y = ty * 4 + threadIdx.y;
    if (x >= p.map_w || y >= p.map_h) continue;
    // ---- sampling.  All S source rows are requested before the first sum: a thread's rows are S independent
    // DRAM round trips, and with 42 KB of tables per CTA occupancy alone does not hide them.
    unsigned long long hyw[S], huvw[S];   // per row: S luma words of 16 bit, S/2 chroma pairs
    unsigned syw[S], suw[S], svw[S];
#pragma unroll
    for (int dy = 0; dy < S; dy++) {
      const int yy = y * S + dy;
      const uint16_t* hyp = (const uint16_t*)p.hdr.p[0] + (size_t)yy * p.hdr.stride[0] + x * S;
      const uint16_t* hcp = (const uint16_t*)p.hdr.p[1] + (size_t)(yy >> 1) * p.hdr.stride[1] + x * S;
      const uint8_t* syp = (const uint8_t*)p.sdr.p[0] + (size_t)yy * p.sdr.stride[0] + x * S;
      const uint8_t* sup = (const uint8_t*)p.sdr.p[1] + (size_t)(yy >> 1) * p.sdr.stride[1] + x * (S / 2);
      const uint8_t* svp = (const uint8_t*)p.sdr.p[2] + (size_t)(yy >> 1) * p.sdr.stride[2] + x * (S / 2);
      if (S == 4) {
        hyw[dy] = ldv_u64(hyp);
        syw[dy] = ldv_u32(syp);
        if (!(dy & 1)) {   // rows 2k and 2k+1 share their chroma row
          huvw[dy] = ldv_u64(hcp);
          suw[dy] = ldv_u16(sup);
          svw[dy] = ldv_u16(svp);
        } else {
          huvw[dy] = huvw[dy - 1]; suw[dy] = suw[dy - 1]; svw[dy] = svw[dy - 1];
        }
      } else {
        hyw[dy] = ldv_u32(hyp);
        syw[dy] = ldv_u16(syp);
        if (!(dy & 1)) {
          huvw[dy] = ldv_u32(hcp);
          suw[dy] = ldv_u8(sup);
          svw[dy] = ldv_u8(svp);
        } else {
          huvw[dy] = huvw[dy - 1]; suw[dy] = suw[dy - 1]; svw[dy] = svw[dy - 1];
        }
      }
    }
    float sy = 0.f, su = 0.f, sv = 0.f, hy = 0.f, hu = 0.f, hv = 0.f;
#pragma unroll
    for (int dy = 0; dy < S; dy++) {
#pragma unroll
      for (int dx = 0; dx < S; dx++) {
        // getYuv420Pixel (gainmapmath.cpp:354-372)
        sy += (float)((syw[dy] >> (8 * dx)) & 0xff) * (1 / 255.0f);
        su += (float)((int)((suw[dy] >> (8 * (dx >> 1))) & 0xff) - 128) * (1 / 255.0f);
        sv += (float)((int)((svw[dy] >> (8 * (dx >> 1))) & 0xff) - 128) * (1 / 255.0f);
        // getP010Pixel (:412-445)
        const int y10 = (int)((hyw[dy] >> (16 * dx + 6)) & 0x3ff);
        const int u10 = (int)((huvw[dy] >> (32 * (dx >> 1) + 6)) & 0x3ff), v10 = (int)((huvw[dy] >> (32 * (dx >> 1) + 22)) & 0x3ff);
        if (LIMITED) {
          hy += (float)(y10 - 64) * (1 / 876.0f);
          hu += (float)(u10 - 64) * (1 / 896.0f) - 0.5f;
          hv += (float)(v10 - 64) * (1 / 896.0f) - 0.5f;
        } else {
          hy += (float)y10 / 1023.0f;
          hu += (float)u10 / 1023.0f - 0.5f;
          hv += (float)v10 / 1023.0f - 0.5f;
        }
      }
    }
    const float inv = 1.0f / (float)(S * S);   // a power of two: the product equals the reference's quotient
    sy *= inv; su *= inv; sv *= inv; hy *= inv; hu *= inv; hv *= inv;
    // ---- yuvToRgb -> inverse OETF tables [-> gamut] -> clipNegatives: the scale-1 expressions on scalars
    float sr = fetch2(sm.srgb2, __saturatef(sy + p.sdr_y2r[0] * sv), 8184.0f);
    float sg = fetch2(sm.srgb2, __saturatef((sy - p.sdr_y2r[2] * su) - p.sdr_y2r[3] * sv), 8184.0f);
    float sb = fetch2(sm.srgb2, __saturatef(sy + p.sdr_y2r[1] * su), 8184.0f);
    float hr = fetch2(sm.hdr2, __saturatef(hy + p.hdr_y2r[0] * hv), hscale8);
    float hg = fetch2(sm.hdr2, __saturatef((hy - p.hdr_y2r[2] * hu) - p.hdr_y2r[3] * hv), hscale8);
    float hb = fetch2(sm.hdr2, __saturatef(hy + p.hdr_y2r[1] * hu), hscale8);
    if (GAMUT != 0) {
      float& xr = GAMUT == 1 ? sr : hr;
      float& xg = GAMUT == 1 ? sg : hg;
      float& xb = GAMUT == 1 ? sb : hb;
      const float a = (p.gamut[0] * xr + p.gamut[1] * xg) + p.gamut[2] * xb;
      const float b = (p.gamut[3] * xr + p.gamut[4] * xg) + p.gamut[5] * xb;
      const float c = (p.gamut[6] * xr + p.gamut[7] * xg) + p.gamut[8] * xb;
      xr = fmaxf(a, 0.0f); xg = fmaxf(b, 0.0f); xb = fmaxf(c, 0.0f);
    }
    float s3[3], h3[3];
    if (NCH == 3) {
      s3[0] = sr * p.sdr_nits; s3[1] = sg * p.sdr_nits; s3[2] = sb * p.sdr_nits;
      h3[0] = hr * p.hdr_nits; h3[1] = hg * p.hdr_nits; h3[2] = hb * p.hdr_nits;
    } else if (p.use_luminance) {
      s3[0] = ((p.lum[0] * sr + p.lum[1] * sg) + p.lum[2] * sb) * p.sdr_nits;
      h3[0] = ((p.lum[0] * hr + p.lum[1] * hg) + p.lum[2] * hb) * p.hdr_nits;
    } else {
      s3[0] = fmaxf(sr, fmaxf(sg, sb)) * p.sdr_nits;
      h3[0] = fmaxf(hr, fmaxf(hg, hb)) * p.hdr_nits;
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      if (ONEPASS) {   // encodeGain gainmapmath.cpp:758-771, gamma 1
        float gain = 1.0f;
        if (s3[c] > 0.0f) gain = div_pos(h3[c], s3[c]);
        if (gain < p.min_boost) gain = p.min_boost;
        if (gain > p.max_boost) gain = p.max_boost;
        p.dst[((size_t)y * p.dst_stride + x) * NCH + c] = (uint8_t)encode_gain_code(gain, p, inv_range_f, thr1, code_lo, code_hi, sm.log2tab);
      } else {         // computeGain :773-782
        const float q = div_pos(h3[c] + 1e-7f, s3[c] + 1e-7f);
        const bool dark = s3[c] < 2.f / 255.0f;
        if (QMODE) {
          p.gains[((size_t)y * p.map_w + x) * NCH + c] = dark ? -q : q;
          if (dark) { dmn[c] = fminf(dmn[c], q); dmx[c] = fmaxf(dmx[c], q); } else { mn[c] = fminf(mn[c], q); mx[c] = fmaxf(mx[c], q); }
        } else {
          float g = (float)log2_core(q, sm.log2tab);
          if (dark) g = fminf(g, 2.3f);
          p.gains[((size_t)y * p.map_w + x) * NCH + c] = g;
          mn[c] = fminf(mn[c], g);
          mx[c] = fmaxf(mx[c], g);
        }
      }
    }
  }
  if (!ONEPASS) {
    reduce_minmax<NCH>(mn, mx, p.minmax, tid, nt);
    if (QMODE) {
      __syncthreads();
      reduce_minmax<NCH>(dmn, dmx, p.minmax + kQDarkKeys, tid, nt);
    }
  }
}

// ---- pass 2 (jpegr.cpp:988-1013, affineMapGain gainmapmath.cpp:784-789), gamma 1, tight rows ------
// The gains and the map are walked as flat arrays, one float4 -> one packed u32 per step.  The
// grid has a multiple of 3 threads, so with three channels every thread keeps the same channel
// phase for its whole walk and holds (min, range, refined 1/range) per element in registers.
// The clamp / hint step between the passes (jpegr.cpp:969-986, k_gainmap_finalize in kernels.cu) is
// folded in: every thread derives the final min / max of its channels from the pass-1 keys (a dozen
// instructions), thread 0 also leaves them in minmax_f for the metadata.
__device__ __forceinline__ void finalized_minmax(const GainmapFinalizeParams& f, int c, float& mn, float& mx) {
  const unsigned kmn = f.minmax[c < f.nch ? c : 0], kmx = f.minmax[3 + (c < f.nch ? c : 0)];
  mn = __uint_as_float((kmn & 0x80000000u) ? (kmn & 0x7fffffffu) : ~kmn);
  mx = __uint_as_float((kmx & 0x80000000u) ? (kmx & 0x7fffffffu) : ~kmx);
  mn = mn < -14.3f ? -14.3f : (mn > 15.6f ? 15.6f : mn);
  mx = mx < -14.3f ? -14.3f : (mx > 15.6f ? 15.6f : mx);
  if (f.has_user_max) mx = fminf(mx, f.log2_user_max);
  if (f.has_user_min) mn = fmaxf(mn, f.log2_user_min);
  if (fabsf(mx - mn) < 1.1920928955078125e-07f) mx += 0.1f;
}

template <int NCH>
__global__ void __launch_bounds__(192) k_affine_fast(const AffineParams p, const GainmapFinalizeParams fin, const long long n4) {
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  if (t0 < 3) {
    float a, b;
    finalized_minmax(fin, (int)t0, a, b);
    fin.minmax_f[t0] = a;
    fin.minmax_f[3 + t0] = b;
  }
  float mn[4], d[4], rc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = NCH == 3 ? (int)((t0 * 4 + j) % 3) : 0;
    float mxj;
    finalized_minmax(fin, c, mn[j], mxj);
    d[j] = mxj - mn[j];
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d[j]));
    rc[j] = __fmaf_rn(r, __fmaf_rn(-d[j], r, 1.0f), r);
  }
  const float4* g4 = reinterpret_cast<const float4*>(p.gains);
  unsigned* o = reinterpret_cast<unsigned*>(p.dst);
#pragma unroll 4
  for (long long f = t0; f < n4; f += nthr) {
    const float4 g = __ldcs(g4 + f);
    const float gv[4] = {g.x, g.y, g.z, g.w};
    unsigned w = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float a = gv[j] - mn[j];
      float q = __fmul_rn(a, rc[j]);           // a / d[j], same steps as div_pos
      q = __fmaf_rn(rc[j], __fmaf_rn(-d[j], q, a), q);
      float t = q * 255.0f + 0.5f;
      t = fminf(fmaxf(t, 0.0f), 255.0f);
      w |= (__float_as_uint(__fadd_rz(t, 8388608.0f)) & 0xffu) << (8 * j);
    }
    o[f] = w;
  }
}

// ---- pass 2 on the quotient plane (store_q) ---------------------------------------------------------
// gain = float(log2(double(q))) capped at 2.3 for dark pixels, then affineMapGain -> byte.  Only the byte leaves the
// kernel, so the fp64 log2 is needed only where the byte could depend on it: every value first goes through the
// hardware's fp32 lg2.approx, whose distance to the exact path is bounded by kLg2Abs + |g| * kLg2Rel (checked over every
// float of the quotient's range on the device, tests/test_gpu_stages.py::test_fast_log2_error_bound, with a factor 2 to
// spare); the affine map turns that into a bound on t = 255 * (g - min) / (max - min) + 0.5, and the byte is
// trunc(clamp(t)).  When t is further than the bound (plus the float roundings of the map itself) from the nearest
// integer, the exact t truncates to the same byte; otherwise the value takes the exact path.  On natural content a few
// values in ten thousand do.
// log2 with the table in global memory (a dozen evaluations per thread at kernel start)
__device__ __forceinline__ float log2f_exact_g(float q, const double* __restrict__ tab_g) {
  const unsigned ix = __float_as_uint(q);
  const unsigned tmp = ix - kLogOff;
  const int i = (tmp >> 16) & 127;
  const int k = (int)tmp >> 23;
  const unsigned im = ix - (tmp & 0xff800000u);
  const double md = __hiloint2double((int)((im >> 3) + 0x38000000u), (int)(im << 29));
  const double kd = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)k)) - 4503601774854144.0;
  const double invc = tab_g[i], logc = tab_g[128 + i];
  const double r = fma(md, invc, -1.0);
  double pp = kLog2Poly[0];
#pragma unroll
  for (int j = 1; j < 8; j++) pp = fma(pp, r, kLog2Poly[j]);
  return (float)fma(pp, r, kd + logc);
}

// min / max of the gains of channel c from the images g[] of the extreme quotients (g[4*c + {0: min q, 1: max q,
// 2: min q dark, 3: max q dark}], a class without pixels flagged by a non-positive max q): fold them like the reference
// folds every pixel (start values 127 / -128, jpegr.cpp:866-868), then the clamps / hints of finalized_minmax
__device__ __forceinline__ void fold_minmax_q(const GainmapFinalizeParams& f, const float* g, const bool* has, float& mn, float& mx) {
  mn = 127.0f;
  mx = -128.0f;
  if (has[0]) {   // non-dark pixels
    mn = fminf(mn, g[0]);
    mx = fmaxf(mx, g[1]);
  }
  if (has[1]) {   // dark pixels: gain = min(log2, 2.3)
    mn = fminf(mn, fminf(g[2], 2.3f));
    mx = fmaxf(mx, fminf(g[3], 2.3f));
  }
  mn = mn < -14.3f ? -14.3f : (mn > 15.6f ? 15.6f : mn);
  mx = mx < -14.3f ? -14.3f : (mx > 15.6f ? 15.6f : mx);
  if (f.has_user_max) mx = fminf(mx, f.log2_user_max);
  if (f.has_user_min) mn = fmaxf(mn, f.log2_user_min);
  if (fabsf(mx - mn) < 1.1920928955078125e-07f) mx += 0.1f;
}

template <int NCH>
__global__ void __launch_bounds__(192) k_affine_q(const AffineParams p, const GainmapFinalizeParams fin, const long long n4,
                                                  const double* __restrict__ tab_g, unsigned* __restrict__ exact_count,
                                                  const unsigned long long nz) {
  __shared__ double2 tab[128];
  __shared__ float s_mm[6];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(tab_g[i], tab_g[128 + i]);
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  __shared__ float s_g[12];
  __shared__ bool s_has[6];
  if (threadIdx.x < 12) {   // once per CTA: the fp64 log2 of the twelve extreme quotients, one per thread
    const int c = threadIdx.x >> 2, k = threadIdx.x & 3, cc = c < fin.nch ? c : 0;
    const unsigned key = fin.minmax[(k >> 1) * kQDarkKeys + (k & 1) * 3 + cc];
    const float q = __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
    if (k & 1) s_has[c * 2 + (k >> 1)] = q > 0.0f;
    s_g[threadIdx.x] = q > 0.0f && q < 3.0e38f ? log2f_exact_g(q, tab_g) : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a, b;
    fold_minmax_q(fin, s_g + 4 * threadIdx.x, s_has + 2 * threadIdx.x, a, b);
    s_mm[threadIdx.x] = a;
    s_mm[3 + threadIdx.x] = b;
    if (blockIdx.x == 0) {
      fin.minmax_f[threadIdx.x] = a;
      fin.minmax_f[3 + threadIdx.x] = b;
    }
  }
  __syncthreads();
  // per element of this thread's float4 (fixed channel phase, see k_affine_fast): min, -range, refined 1/range and the
  // screen's threshold.  A value whose t lands in [-1, 257] has |g| <= max(|min|, |max|) + 1: the threshold is constant.
  float mn[4], nd[4], rc[4], thr[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = NCH == 3 ? (int)((t0 * 4 + j) % 3) : 0;
    mn[j] = s_mm[c];
    const float dj = s_mm[3 + c] - mn[j];
    nd[j] = -dj;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dj));
    rc[j] = __fmaf_rn(r, __fmaf_rn(-dj, r, 1.0f), r);
    const float gmax = fmaxf(fabsf(mn[j]), fabsf(s_mm[3 + c])) + 1.0f;
    thr[j] = (kLg2Abs + gmax * kLg2Rel) * (255.0f * fabsf(rc[j]) * 1.0001f) + kAffineRound;   // |rc|: hints can turn the range around
  }
  const V2 mnv[2] = {v2(mn[0], mn[1]), v2(mn[2], mn[3])}, ndv[2] = {v2(nd[0], nd[1]), v2(nd[2], nd[3])};
  const V2 rcv[2] = {v2(rc[0], rc[1]), v2(rc[2], rc[3])};
  const V2 k255 = bc(255.0f), khalf = bc(0.5f), kmagic = bc(12582912.0f);   // 1.5 * 2^23: adding it rounds to an integer
  const float4* g4 = reinterpret_cast<const float4*>(p.gains);
  unsigned* o = reinterpret_cast<unsigned*>(p.dst);
  unsigned n_exact = 0;
  // two loads ahead of the value being processed (four: 68 registers, slower): the byte stores may alias the plane as far as the compiler knows, so
  // it would not move the next load above them by itself
  float4 q1 = t0 < n4 ? __ldcs(g4 + t0) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 q2 = t0 + nthr < n4 ? __ldcs(g4 + t0 + nthr) : q1;
#pragma unroll 1
  for (long long f = t0; f < n4; f += nthr) {
    const float4 qv = q1;
    q1 = q2;
    if (f + 2 * nthr < n4) q2 = __ldcs(g4 + f + 2 * nthr);
    const float qs[4] = {qv.x, qv.y, qv.z, qv.w};
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      g[j] = lg2_fast(fabsf(qs[j]));
      if (qs[j] < 0.0f) g[j] = fminf(g[j], 2.3f);   // dark pixel
    }
    unsigned w = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      // the steps of k_affine_fast on a pair: a = g - min; a / range (div_pos steps); * 255; + 0.5
      const V2 a = vsub(v2(g[2 * h], g[2 * h + 1]), mnv[h]);
      V2 qq = vmul(a, rcv[h], nz);
      qq = vfma(rcv[h], vfma(ndv[h], qq, a), qq);
      const V2 t = vadd(vmul(qq, k255, nz), khalf);
      float tt[2], dl[2];
      un(t, tt[0], tt[1]);
      un(vsub(t, vsub(vadd(t, kmagic), kmagic)), dl[0], dl[1]);   // t - nearest integer
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int j = 2 * h + e;
        if (fabsf(dl[e]) < thr[j]) {   // the byte could depend on the last bits of the log2: exact path
          float ge = (float)log2_core(fabsf(qs[j]), tab);
          if (qs[j] < 0.0f) ge = fminf(ge, 2.3f);
          const float ae = ge - mn[j];
          float qe = __fmul_rn(ae, rc[j]);
          qe = __fmaf_rn(rc[j], __fmaf_rn(nd[j], qe, ae), qe);
          tt[e] = qe * 255.0f + 0.5f;
          n_exact++;
        }
        const float tc = fminf(fmaxf(tt[e], 0.0f), 255.0f);
        w |= (__float_as_uint(__fadd_rz(tc, 8388608.0f)) & 0xffu) << (8 * j);
      }
    }
    o[f] = w;
  }
  if (exact_count && n_exact) atomicAdd(exact_count, n_exact);
}

// diagnostic: out[i] = |lg2.approx(in[i]) - float(log2(double(in[i])))| / lg2_fast_bound(lg2.approx(in[i]))
__global__ void k_log2_fast_probe(unsigned first_bits, unsigned count, float* __restrict__ worst, const double* __restrict__ tab_g) {
  __shared__ double2 tab[128];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(tab_g[i], tab_g[128 + i]);
  __syncthreads();
  float w = 0.0f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float q = __uint_as_float(first_bits + i);
    const float a = lg2_fast(q), e = (float)log2_core(q, tab);
    w = fmaxf(w, fabsf(a - e) / lg2_fast_bound(a));
  }
  for (int o = 16; o; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned*>(worst), __float_as_uint(w));   // w >= 0: bit order == value order
}

// ---- convertYuv 4:2:0 (transformYuv420, gainmapmath.cpp:686-726), in place -----------------------
// thread = 4x2 luma samples + their two chroma pairs (32-bit / 16-bit accesses instead of single
// bytes), the per-pixel products and sums on packed pairs.  Same operand order as k_yuv_convert.
__global__ void __launch_bounds__(256) k_yuv420_fast(const YuvConvParams p, const unsigned long long nz) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = (blockIdx.y * blockDim.y + threadIdx.y) * 2;
  if (x >= p.w || y >= p.h) return;
  const unsigned yw[2] = {*reinterpret_cast<const unsigned*>(p.p[0] + (size_t)y * p.stride[0] + x),
                          *reinterpret_cast<const unsigned*>(p.p[0] + (size_t)(y + 1) * p.stride[0] + x)};
  const unsigned uu = *reinterpret_cast<const uint16_t*>(p.p[1] + (size_t)(y >> 1) * p.stride[1] + (x >> 1));
  const unsigned vv = *reinterpret_cast<const uint16_t*>(p.p[2] + (size_t)(y >> 1) * p.stride[2] + (x >> 1));
  const V2 k255i = bc(1 / 255.0f), k255 = bc(255.0f), khalf = bc(0.5f);
  const V2 m0 = bc(p.m[0]), m3 = bc(p.m[3]), m6 = bc(p.m[6]);
  unsigned oy[2] = {0, 0}, ou = 0, ov = 0;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const float u = (float)((int)((uu >> (8 * k)) & 0xff) - 128) * (1 / 255.0f);
    const float v = (float)((int)((vv >> (8 * k)) & 0xff) - 128) * (1 / 255.0f);
    const V2 um1 = bc(u * p.m[1]), vm2 = bc(v * p.m[2]), um4 = bc(u * p.m[4]), vm5 = bc(v * p.m[5]);
    const V2 um7 = bc(u * p.m[7]), vm8 = bc(v * p.m[8]);
    float cu[4], cv[4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const V2 yy = vmul(v2((float)((yw[r] >> (16 * k)) & 0xff), (float)((yw[r] >> (16 * k + 8)) & 0xff)), k255i, nz);
      const V2 ny = vadd(vadd(vmul(yy, m0, nz), um1), vm2);
      un(vadd(vadd(vmul(yy, m3, nz), um4), vm5), cu[2 * r], cu[2 * r + 1]);
      un(vadd(vadd(vmul(yy, m6, nz), um7), vm8), cv[2 * r], cv[2 * r + 1]);
      float t0, t1;
      un(vadd(vmul(ny, k255, nz), khalf), t0, t1);
      unsigned b0, b1;
      un(vtrunc_bits(v2(fminf(fmaxf(t0, 0.0f), 255.0f), fminf(fmaxf(t1, 0.0f), 255.0f))), b0, b1);
      oy[r] |= ((b0 & 0xff) << (16 * k)) | ((b1 & 0xff) << (16 * k + 8));
    }
    const float su = (((cu[0] + cu[1]) + cu[2]) + cu[3]) * 0.25f;   // / 4.0f: exact either way
    const float sv = (((cv[0] + cv[1]) + cv[2]) + cv[3]) * 0.25f;
    const float tu = fminf(fmaxf(su * 255.0f + 128.0f + 0.5f, 0.0f), 255.0f);
    const float tv = fminf(fmaxf(sv * 255.0f + 128.0f + 0.5f, 0.0f), 255.0f);
    ou |= (__float_as_uint(__fadd_rz(tu, 8388608.0f)) & 0xff) << (8 * k);
    ov |= (__float_as_uint(__fadd_rz(tv, 8388608.0f)) & 0xff) << (8 * k);
  }
  *reinterpret_cast<unsigned*>(p.d[0] + (size_t)y * p.dstride[0] + x) = oy[0];
  *reinterpret_cast<unsigned*>(p.d[0] + (size_t)(y + 1) * p.dstride[0] + x) = oy[1];
  *reinterpret_cast<uint16_t*>(p.d[1] + (size_t)(y >> 1) * p.dstride[1] + (x >> 1)) = (uint16_t)ou;
  *reinterpret_cast<uint16_t*>(p.d[2] + (size_t)(y >> 1) * p.dstride[2] + (x >> 1)) = (uint16_t)ov;
}

// host: table of the log2 kernel, uploaded once per device
int log2_table_dev(const double** out) {
  static std::mutex mu;
  static thread_local int cached_dev = -1;
  static thread_local const double* cached = nullptr;
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (cached && cached_dev == dev) { *out = cached; return E_OK; }
  std::lock_guard<std::mutex> lk(mu);
  static double* per_dev[64] = {nullptr};
  if (dev < 64 && per_dev[dev]) { cached = per_dev[dev]; cached_dev = dev; *out = cached; return E_OK; }
  double host[256];
  for (int i = 0; i < 128; i++) {
    const unsigned lo = kLogOff + ((unsigned)i << 16), hi = lo + (1u << 16);
    float flo, fhi;
    memcpy(&flo, &lo, 4);
    memcpy(&fhi, &hi, 4);
    long double c = ((long double)flo + (long double)fhi) / 2;
    if (flo <= 1.0f && fhi >= 1.0f) c = 1.0L;       // the two sub-intervals that touch 1.0
    const double invc = (double)(1.0L / c);
    host[i] = invc;
    host[128 + i] = (invc == 1.0) ? 0.0 : (double)(-log2l((long double)invc));  // log2 of the c actually used
  }
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof host));
  CUDA_TRY(cudaMemcpy(d, host, sizeof host, cudaMemcpyHostToDevice));
  if (dev < 64) per_dev[dev] = d;
  cached = d;
  cached_dev = dev;
  *out = d;
  return E_OK;
}

struct FastLaunch {
  const double* tab;
  int tiles_x, ntiles;
  unsigned* sched;
  size_t smem;
  cudaStream_t s;
};
template <bool ONEPASS, int NCH, int GAMUT, bool LIMITED, bool QMODE>
cudaError_t launch_kernel_q(const GainmapGenParams& p, const FastLaunch& L) {
  // persistent grid = the CTAs that are co-resident (asked once per instantiation and device)
  static int resident[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto fn = k_gainmap_fast<ONEPASS, NCH, GAMUT, LIMITED, QMODE>;
  if (dev < 0 || dev >= 64) dev = 0;
  if (!resident[dev]) {
    int per_sm = 0, sms = 0;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, L.smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    resident[dev] = per_sm * (sms > 0 ? sms : 148);
  }
  const int ctas = resident[dev] < L.ntiles ? resident[dev] : L.ntiles;
  count_launches(1);
  fn<<<ctas, dim3(64, 4), L.smem, L.s>>>(p, L.tab, L.tiles_x, L.ntiles, L.sched, kNegZero2);
  return cudaGetLastError();
}
template <bool ONEPASS, int NCH, int GAMUT, bool LIMITED, int S, bool QMODE>
cudaError_t launch_scaled_q(const GainmapGenParams& p, const FastLaunch& L) {
  static int resident[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  auto fn = k_gainmap_scaled<ONEPASS, NCH, GAMUT, LIMITED, S, QMODE>;
  if (dev < 0 || dev >= 64) dev = 0;
  if (!resident[dev]) {
    int per_sm = 0, sms = 0;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, L.smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    resident[dev] = per_sm * (sms > 0 ? sms : 148);
  }
  const int ntiles = ((p.map_w + 63) / 64) * ((p.map_h + 3) / 4);
  const int ctas = resident[dev] < ntiles ? resident[dev] : ntiles;
  count_launches(1);
  fn<<<ctas, dim3(64, 4), L.smem, L.s>>>(p, L.tab);
  return cudaGetLastError();
}
template <bool ONEPASS, int NCH, int GAMUT, bool LIMITED>
cudaError_t launch_kernel(const GainmapGenParams& p, const FastLaunch& L) {
  if (!ONEPASS && p.store_q) return launch_kernel_q<ONEPASS, NCH, GAMUT, LIMITED, !ONEPASS>(p, L);
  return launch_kernel_q<ONEPASS, NCH, GAMUT, LIMITED, false>(p, L);
}
template <bool ONEPASS, int NCH, int GAMUT, bool LIMITED, int S>
cudaError_t launch_scaled(const GainmapGenParams& p, const FastLaunch& L) {
  if (!ONEPASS && p.store_q) return launch_scaled_q<ONEPASS, NCH, GAMUT, LIMITED, S, !ONEPASS>(p, L);
  return launch_scaled_q<ONEPASS, NCH, GAMUT, LIMITED, S, false>(p, L);
}
template <bool ONEPASS, int NCH, int GAMUT>
cudaError_t launch_range(const GainmapGenParams& p, const FastLaunch& L) {
  if (p.scale == 4)
    return p.hdr.full_range ? launch_scaled<ONEPASS, NCH, GAMUT, false, 4>(p, L) : launch_scaled<ONEPASS, NCH, GAMUT, true, 4>(p, L);
  if (p.scale == 2)
    return p.hdr.full_range ? launch_scaled<ONEPASS, NCH, GAMUT, false, 2>(p, L) : launch_scaled<ONEPASS, NCH, GAMUT, true, 2>(p, L);
  return p.hdr.full_range ? launch_kernel<ONEPASS, NCH, GAMUT, false>(p, L) : launch_kernel<ONEPASS, NCH, GAMUT, true>(p, L);
}
template <bool ONEPASS, int NCH>
cudaError_t launch_g(const GainmapGenParams& p, const FastLaunch& L) {
  const int gm = p.gamut_identity ? 0 : (p.gamut_on_hdr ? 2 : 1);
  if (gm == 0) return launch_range<ONEPASS, NCH, 0>(p, L);
  if (gm == 1) return launch_range<ONEPASS, NCH, 1>(p, L);
  return launch_range<ONEPASS, NCH, 2>(p, L);
}

}  // namespace

bool affine_fast_eligible(const AffineParams& p) {
  if (p.gamma != 1.0f || (p.nch != 1 && p.nch != 3) || p.dst_stride != p.map_w) return false;
  if (((size_t)p.map_w * p.map_h * p.nch) & 3) return false;
  return !(((size_t)p.gains & 15) || ((size_t)p.dst & 3));
}
// q mode: pass 2 on the quotient plane.  exact_count (device word, may be null): how many values took the fp64 path.
cudaError_t launch_affine_q(const AffineParams& p, const GainmapFinalizeParams& fin, unsigned* exact_count, cudaStream_t s) {
  const double* tab = nullptr;
  if (log2_table_dev(&tab) != E_OK) return cudaErrorUnknown;
  const long long n4 = (long long)p.map_w * p.map_h * p.nch / 4;
  long long ctas = (n4 + 192 * 4 - 1) / (192 * 4);
  // grid-stride walk: exactly the co-resident CTAs (a partial second wave would cost a whole one)
  static int resident[2] = {0, 0};
  int& res = resident[p.nch == 3 ? 1 : 0];
  if (!res) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const cudaError_t oe = p.nch == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_affine_q<3>, 192, 0)
                                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_affine_q<1>, 192, 0);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    res = per_sm * (sms > 0 ? sms : 148);
  }
  if (ctas > res) ctas = res;
  if (ctas < 1) ctas = 1;
  if (p.nch == 3) k_affine_q<3><<<(unsigned)ctas, 192, 0, s>>>(p, fin, n4, tab, exact_count, kNegZero2);
  else k_affine_q<1><<<(unsigned)ctas, 192, 0, s>>>(p, fin, n4, tab, exact_count, kNegZero2);
  return cudaGetLastError();
}
// start values of the q keys (and the tile-ticket word 8)
__global__ void k_init_q_keys(unsigned* mm) {
  const int t = threadIdx.x;
  auto key = [](float f) { const unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); };
  if (t < 3 || (t >= kQDarkKeys && t < kQDarkKeys + 3)) mm[t] = key(kQMinInit);
  else if ((t >= 3 && t < 6) || (t >= kQDarkKeys + 3 && t < kQDarkKeys + 6)) mm[t] = key(0.0f);
  else if (t == 8 || t == 9) mm[t] = 0;   // 8: tile tickets of k_gainmap_fast, 9: exact-path counter of k_affine_q
}
cudaError_t launch_init_q_keys(unsigned* minmax, cudaStream_t s) {
  count_launches(1);
  k_init_q_keys<<<1, 32, 0, s>>>(minmax);
  return cudaGetLastError();
}
// worst[0] (device float, zeroed by the caller) = max over the `count` floats starting at bit pattern first_bits of
// |lg2.approx - exact| / bound
cudaError_t launch_log2_fast_probe(unsigned first_bits, unsigned count, float* d_worst, cudaStream_t s) {
  const double* tab = nullptr;
  if (log2_table_dev(&tab) != E_OK) return cudaErrorUnknown;
  k_log2_fast_probe<<<148 * 8, 256, 0, s>>>(first_bits, count, d_worst, tab);
  return cudaGetLastError();
}

cudaError_t launch_affine_fast(const AffineParams& p, const GainmapFinalizeParams& fin, cudaStream_t s) {
  const long long n4 = (long long)p.map_w * p.map_h * p.nch / 4;
  long long ctas = (n4 + 192 * 4 - 1) / (192 * 4);
  if (ctas > 148 * 10) ctas = 148 * 10;
  if (ctas < 1) ctas = 1;
  if (p.nch == 3) k_affine_fast<3><<<(unsigned)ctas, 192, 0, s>>>(p, fin, n4);
  else k_affine_fast<1><<<(unsigned)ctas, 192, 0, s>>>(p, fin, n4);
  return cudaGetLastError();
}

bool yuv420_fast_eligible(const YuvConvParams& p) {
  if (p.fmt != F_YUV420 || (p.w & 3) || (p.h & 1)) return false;
  if ((p.stride[0] & 3) || (p.stride[1] & 1) || (p.stride[2] & 1)) return false;
  if ((p.dstride[0] & 3) || (p.dstride[1] & 1) || (p.dstride[2] & 1)) return false;
  if (((size_t)p.d[0] & 3) || ((size_t)p.d[1] & 1) || ((size_t)p.d[2] & 1)) return false;
  return !(((size_t)p.p[0] & 3) || ((size_t)p.p[1] & 1) || ((size_t)p.p[2] & 1));
}
cudaError_t launch_yuv420_fast(const YuvConvParams& p, cudaStream_t s) {
  dim3 b(64, 4), g((p.w / 4 + 63) / 64, (p.h / 2 + 3) / 4);
  k_yuv420_fast<<<g, b, 0, s>>>(p, kNegZero2);
  return cudaGetLastError();
}

bool gainmap_fast_eligible(const GainmapGenParams& p, bool onepass) {
  if (p.hdr.fmt != F_P010 || p.sdr.fmt != F_YUV420) return false;
  if (p.hdr_ct != CT_HLG && p.hdr_ct != CT_PQ && p.hdr_ct != CT_SRGB) return false;
  if (p.scale == 2 || p.scale == 4) {
    // k_gainmap_scaled: one load per source row and plane -> S luma samples / S/2 chroma pairs must be aligned
    const int S = p.scale;
    if (p.map_w != p.hdr.w / S || p.map_h != p.hdr.h / S || p.map_w < 1 || p.map_h < 1) return false;
    if (p.sdr.w < p.map_w * S || p.sdr.h < p.map_h * S) return false;
    const size_t am = (size_t)S * 2 - 1;   // bytes of a P010 row fragment - 1
    if ((p.hdr.stride[0] * 2 & am) || (p.hdr.stride[1] * 2 & am) || ((size_t)p.hdr.p[0] & am) || ((size_t)p.hdr.p[1] & am)) return false;
    if ((p.sdr.stride[0] & (S - 1)) || ((size_t)p.sdr.p[0] & (S - 1))) return false;
    if ((p.sdr.stride[1] & (S / 2 - 1)) || (p.sdr.stride[2] & (S / 2 - 1)) || ((size_t)p.sdr.p[1] & (S / 2 - 1)) || ((size_t)p.sdr.p[2] & (S / 2 - 1))) return false;
    if (onepass) return p.gamma == 1.0f;
    return true;
  }
  if (p.scale != 1) return false;
  if ((p.map_w & 3) || (p.map_h & 1) || p.map_w != p.hdr.w || p.map_h != p.hdr.h) return false;
  if ((p.hdr.stride[0] & 3) || (p.hdr.stride[1] & 3) || (p.sdr.stride[0] & 3) || (p.sdr.stride[1] & 1) || (p.sdr.stride[2] & 1)) return false;
  if (((size_t)p.hdr.p[0] & 7) || ((size_t)p.hdr.p[1] & 7) || ((size_t)p.sdr.p[0] & 3) || ((size_t)p.sdr.p[1] & 1) || ((size_t)p.sdr.p[2] & 1)) return false;
  if (onepass) {
    if (p.gamma != 1.0f) return false;
    if (((size_t)p.dst & 3) || ((p.dst_stride * p.nch) & 3)) return false;
  } else if ((size_t)p.gains & 15) {
    return false;
  }
  return true;
}

namespace {
__global__ void k_powf_probe(const float* __restrict__ in, float y, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = powf_glibc(in[i], y);
}
__global__ void k_log2_probe(const float* __restrict__ in, float* __restrict__ out, int n, const double* __restrict__ tab_g) {
  __shared__ double2 tab[128];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = make_double2(tab_g[i], tab_g[128 + i]);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)log2_core(in[i], tab);
}
}  // namespace

// diagnostic: out[i] = float(log2(double(in[i]))) as the gain-map kernels evaluate it (device buffers)
cudaError_t launch_log2_probe(const float* d_in, float* d_out, int n, cudaStream_t s) {
  const double* tab = nullptr;
  if (log2_table_dev(&tab) != E_OK) return cudaErrorUnknown;
  k_log2_probe<<<(n + 255) / 256, 256, 0, s>>>(d_in, d_out, n, tab);
  return cudaGetLastError();
}

cudaError_t launch_powf_probe(const float* d_in, float y, float* d_out, int n, cudaStream_t s) {
  k_powf_probe<<<(n + 255) / 256, 256, 0, s>>>(d_in, y, d_out, n);
  return cudaGetLastError();
}

// sched: one zeroed device word (tile tickets of the persistent grid)
cudaError_t launch_gainmap_fast(const GainmapGenParams& p, bool onepass, unsigned* sched, cudaStream_t s) {
  FastLaunch L;
  if (log2_table_dev(&L.tab) != E_OK) return cudaErrorUnknown;
  L.tiles_x = (p.map_w / 4 + 63) / 64;
  L.ntiles = L.tiles_x * ((p.map_h + 7) / 8);
  L.sched = sched;
  L.smem = sizeof(GmSmem);
  L.s = s;
  if (onepass) return p.nch == 3 ? launch_g<true, 3>(p, L) : launch_g<true, 1>(p, L);
  return p.nch == 3 ? launch_g<false, 3>(p, L) : launch_g<false, 1>(p, L);
}

}  // namespace uhdr_b200
