/* Exhaustive proof of the half-float conversion k_apply_lin1 / k_apply_fast use (apply_fast.cu
 * pack_half4): for EVERY binary32 value in [0, 10000/203] -- the range clampPixelFloatLinear leaves --
 * the reference's floatToHalf (lib/include/ultrahdr/gainmapmath.h:160-173: add 0x1000 to the bits and
 * truncate, i.e. round half UP, also in its denormal branch) equals the IEEE round-to-nearest-even
 * conversion of the float whose last mantissa bit was forced to 1 (what cvt.rn.f16x2.f32 computes on the
 * device from `bits | 1`).  The forced bit only ever moves an exact tie upwards.
 *
 *   gcc -O2 -mf16c -fopenmp tools/check_pack_half4.c -o /tmp/check_pack_half4 && /tmp/check_pack_half4 [stride]
 * stride 1 = all 1,111,821,148 values (a few seconds on 8 cores); tests/test_pack_half4_cpu.py runs a
 * strided subset plus every tie / boundary pattern in the CPU suite and the full sweep when asked. */
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint16_t ref_float_to_half(uint32_t bits) { /* restated from the reference, integer form */
  const uint32_t b = bits + 0x00001000u;
  const int32_t e = (int32_t)((b & 0x7F800000u) >> 23);
  const uint32_t m = b & 0x007FFFFFu;
  uint32_t r = (b & 0x80000000u) >> 16;
  if (e > 112) r |= ((((uint32_t)(e - 112) << 10) & 0x7C00u) | (m >> 13));
  if (e < 113 && e > 101) r |= ((((0x007FF000u + m) >> (125 - e)) + 1u) >> 1);
  if (e > 143) r |= 0x7FFFu;
  return (uint16_t)r;
}
static uint16_t hw_rn_of_bits_or_1(uint32_t bits) {
  const uint32_t forced = bits | 1u;
  float f;
  memcpy(&f, &forced, 4);
  return (uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
}

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;
  const float top = 10000.0f / 203.0f;
  uint32_t top_bits;
  memcpy(&top_bits, &top, 4);
  unsigned long long bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
  for (long long i = 0; i <= (long long)top_bits; i += stride) {
    const uint32_t b = (uint32_t)i;
    n++;
    if (ref_float_to_half(b) != hw_rn_of_bits_or_1(b)) bad++;
  }
  /* every half-way pattern (low 13 bits == 0x1000) and its neighbours, whatever the stride */
  for (uint32_t hi = 0; hi <= (top_bits >> 13); hi++)
    for (int d = -2; d <= 2; d++) {
      const uint32_t b = (hi << 13) + 0x1000u + (uint32_t)d;
      if (b > top_bits) continue;
      n++;
      if (ref_float_to_half(b) != hw_rn_of_bits_or_1(b)) bad++;
    }
  printf("checked %llu values in [0, %.9g] (stride %u): %llu mismatches\n", n, top, stride, bad);
  return bad ? 1 : 0;
}
