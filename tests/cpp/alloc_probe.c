/* Heap-call probe for libuhdr_b200.so ("allocation-free" host container / marker layer).
 *
 * This executable defines malloc/calloc/realloc/free itself (glibc lets a program replace them; the shared
 * libraries it loads then use these too) and forwards to glibc's __libc_* entry points.  While `armed`, every
 * allocation whose call stack passes through libuhdr_b200.so before it reaches any other library is counted and
 * its stack printed.  Allocations that libcuda makes for itself (stack reaches libcuda first) are reported
 * separately: they are the driver's business, not ours.
 *
 *   alloc_probe api4 base.jpg gainmap.jpg     CPU only: uhdr_encode API-4 (container assembly) + probe of the result
 *   alloc_probe gpu  W H                      on a GPU: re-armed API-1 uhdr_encode, reset+set_raw_image+encode,
 *                                             reset+set_image+uhdr_decode, on warmed handles
 * prints "ours=<n> cuda=<n> other=<n>" per phase; exit status 0 iff ours == 0 everywhere.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <execinfo.h>
#include <link.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ultrahdr_api.h"
#include "uhdr_b200.h"

extern void* __libc_malloc(size_t);
extern void* __libc_calloc(size_t, size_t);
extern void* __libc_realloc(void*, size_t);
extern void* __libc_memalign(size_t, size_t);
extern void __libc_free(void*);

static volatile int armed;
static int count_warmup;   /* ALLOC_PROBE_COUNT_WARMUP=1: also count the warm-up iterations (self check: must see > 0) */
static __thread int inside;
static long n_ours, n_cuda, n_other;
static uintptr_t ours_lo, ours_hi, cuda_lo, cuda_hi, self_lo, self_hi, c_lo[4], c_hi[4];
static int n_c;

static int phdr_cb(struct dl_phdr_info* info, size_t size, void* data) {
  (void)size; (void)data;
  uintptr_t lo = UINTPTR_MAX, hi = 0;
  for (int i = 0; i < info->dlpi_phnum; i++)
    if (info->dlpi_phdr[i].p_type == PT_LOAD) {
      const uintptr_t a = info->dlpi_addr + info->dlpi_phdr[i].p_vaddr;
      if (a < lo) lo = a;
      if (a + info->dlpi_phdr[i].p_memsz > hi) hi = a + info->dlpi_phdr[i].p_memsz;
    }
  const char* nm = info->dlpi_name ? info->dlpi_name : "";
  if (strstr(nm, "libuhdr_b200")) { ours_lo = lo; ours_hi = hi; }
  else if (strstr(nm, "libcuda.so")) { cuda_lo = lo; cuda_hi = hi; }
  else if (!nm[0] && !self_hi) { self_lo = lo; self_hi = hi; }
  else if ((strstr(nm, "libc.so") || strstr(nm, "libstdc++") || strstr(nm, "libgcc_s") || strstr(nm, "libm.so")) && n_c < 4) {
    c_lo[n_c] = lo; c_hi[n_c] = hi; n_c++;   /* pass-through frames: operator new, realloc internals ... */
  }
  return 0;
}

static void note(size_t bytes) {
  if (!armed || inside) return;
  inside = 1;
  void* bt[24];
  const int n = backtrace(bt, 24);
  int who = 2;   /* other */
  for (int i = 1; i < n; i++) {
    const uintptr_t a = (uintptr_t)bt[i];
    int skip = a >= self_lo && a < self_hi;   /* our own malloc wrappers */
    for (int k = 0; k < n_c && !skip; k++) skip = a >= c_lo[k] && a < c_hi[k];
    if (skip) continue;
    if (a >= ours_lo && a < ours_hi) who = 0;
    else if (a >= cuda_lo && a < cuda_hi) who = 1;
    break;
  }
  if (who == 0) {
    n_ours++;
    if (n_ours <= 8) {
      fprintf(stderr, "--- heap call of %zu bytes from libuhdr_b200:\n", bytes);
      backtrace_symbols_fd(bt, n, 2);
    }
  } else if (who == 1) {
    n_cuda++;
  } else {
    n_other++;
  }
  inside = 0;
}

void* malloc(size_t n) { note(n); return __libc_malloc(n); }
void* calloc(size_t a, size_t b) { note(a * b); return __libc_calloc(a, b); }
void* realloc(void* p, size_t n) { note(n); return __libc_realloc(p, n); }
void* memalign(size_t a, size_t n) { note(n); return __libc_memalign(a, n); }
void* aligned_alloc(size_t a, size_t n) { note(n); return __libc_memalign(a, n); }
int posix_memalign(void** out, size_t a, size_t n) {
  note(n);
  void* p = __libc_memalign(a, n);
  if (!p) return 12;
  *out = p;
  return 0;
}
void free(void* p) { __libc_free(p); }

static int report(const char* phase) {
  printf("%-44s ours=%ld cuda=%ld other=%ld\n", phase, n_ours, n_cuda, n_other);
  const int bad = n_ours != 0;
  n_ours = n_cuda = n_other = 0;
  return bad;
}
#define CHECK(e) do { uhdr_error_info_t s_ = (e); if (s_.error_code != UHDR_CODEC_OK) { fprintf(stderr, "%s failed: %s\n", #e, s_.has_detail ? s_.detail : ""); exit(2); } } while (0)

static unsigned char* slurp(const char* path, size_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  *n = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char* b = (unsigned char*)__libc_malloc(*n);
  if (fread(b, 1, *n, f) != *n) exit(2);
  fclose(f);
  return b;
}

static int run_api4(const char* base_path, const char* gm_path) {
  size_t bn, gn;
  unsigned char* base = slurp(base_path, &bn);
  unsigned char* gm = slurp(gm_path, &gn);
  uhdr_gainmap_metadata_t md;
  memset(&md, 0, sizeof md);
  for (int i = 0; i < 3; i++) {
    md.max_content_boost[i] = 4.0f + i; md.min_content_boost[i] = 1.0f; md.gamma[i] = 1.0f;
    md.offset_sdr[i] = md.offset_hdr[i] = 1.0f / 64;
  }
  md.hdr_capacity_min = 1.0f; md.hdr_capacity_max = 4.9f; md.use_base_cg = 1;
  uhdr_compressed_image_t b = {base, bn, bn, UHDR_CG_BT_709, UHDR_CT_SRGB, UHDR_CR_FULL_RANGE};
  uhdr_compressed_image_t g = {gm, gn, gn, UHDR_CG_UNSPECIFIED, UHDR_CT_UNSPECIFIED, UHDR_CR_UNSPECIFIED};
  uhdr_codec_private_t* enc = uhdr_create_encoder();
  uhdr_codec_private_t* dec = uhdr_create_decoder();
  int bad = 0;
  for (int it = 0; it < 4; it++) {   /* iteration 0 warms the handles (buffers grow once), then counted */
    armed = it > 0 || count_warmup;
    uhdr_reset_encoder(enc);
    CHECK(uhdr_enc_set_compressed_image(enc, &b, UHDR_BASE_IMG));
    CHECK(uhdr_enc_set_gainmap_image(enc, &g, &md));
    CHECK(uhdr_encode(enc));
    uhdr_compressed_image_t* out = uhdr_get_encoded_stream(enc);
    uhdr_reset_decoder(dec);
    CHECK(uhdr_dec_set_image(dec, out));
    CHECK(uhdr_dec_probe(dec));
    if (uhdr_dec_get_image_width(dec) <= 0 || !uhdr_dec_get_gainmap_metadata(dec)) exit(2);
    armed = 0;
  }
  bad |= report("api-4 encode + probe (host only)");
  uhdr_release_encoder(enc);
  uhdr_release_decoder(dec);
  return bad;
}

static int run_gpu(int w, int h) {
  const size_t npx = (size_t)w * h;
  uint16_t* p010 = (uint16_t*)__libc_malloc(npx * 3);   /* Y + interleaved UV */
  uint8_t* yuv = (uint8_t*)__libc_malloc(npx * 3 / 2);
  uint32_t s = 12345;
  for (size_t i = 0; i < npx * 3 / 2; i++) {
    s = s * 1664525u + 1013904223u;
    const unsigned ramp = (unsigned)((i % (size_t)w) * 255 / (size_t)w);
    p010[i] = (uint16_t)(((64 + ((ramp * 3 + (s >> 28)) % 876)) & 0x3ff) << 6);
    yuv[i] = (uint8_t)((ramp + (s >> 29)) & 0xff);
  }
  uhdr_raw_image_t hdr, sdr;
  memset(&hdr, 0, sizeof hdr);
  memset(&sdr, 0, sizeof sdr);
  hdr.fmt = UHDR_IMG_FMT_24bppYCbCrP010; hdr.cg = UHDR_CG_BT_2100; hdr.ct = UHDR_CT_HLG; hdr.range = UHDR_CR_LIMITED_RANGE;
  hdr.w = w; hdr.h = h; hdr.planes[0] = p010; hdr.planes[1] = p010 + npx; hdr.stride[0] = w; hdr.stride[1] = w;
  sdr.fmt = UHDR_IMG_FMT_12bppYCbCr420; sdr.cg = UHDR_CG_BT_709; sdr.ct = UHDR_CT_SRGB; sdr.range = UHDR_CR_FULL_RANGE;
  sdr.w = w; sdr.h = h; sdr.planes[0] = yuv; sdr.planes[1] = yuv + npx; sdr.planes[2] = yuv + npx + npx / 4;
  sdr.stride[0] = w; sdr.stride[1] = sdr.stride[2] = w / 2;
  int bad = 0;
  uhdr_codec_private_t* enc = uhdr_create_encoder();
  CHECK(uhdr_enc_set_raw_image(enc, &hdr, UHDR_HDR_IMG));
  CHECK(uhdr_enc_set_raw_image(enc, &sdr, UHDR_SDR_IMG));
  for (int it = 0; it < 6; it++) {   /* resident inputs, re-armed */
    armed = it >= 3 || count_warmup;
    if (it) uhdr_b200_enc_rearm(enc);
    CHECK(uhdr_encode(enc));
    if (!uhdr_get_encoded_stream(enc)) exit(2);
    armed = 0;
  }
  bad |= report("api-1 encode, resident inputs (re-armed)");
  for (int it = 0; it < 6; it++) {   /* the C API sequence of a streaming caller */
    armed = it >= 3 || count_warmup;
    uhdr_reset_encoder(enc);
    CHECK(uhdr_enc_set_raw_image(enc, &hdr, UHDR_HDR_IMG));
    CHECK(uhdr_enc_set_raw_image(enc, &sdr, UHDR_SDR_IMG));
    CHECK(uhdr_encode(enc));
    armed = 0;
  }
  bad |= report("api-1 reset + set_raw_image x2 + encode");
  uhdr_compressed_image_t* out = uhdr_get_encoded_stream(enc);
  uhdr_codec_private_t* dec = uhdr_create_decoder();
  for (int it = 0; it < 6; it++) {
    armed = it >= 3 || count_warmup;
    uhdr_reset_decoder(dec);
    CHECK(uhdr_dec_set_image(dec, out));
    CHECK(uhdr_decode(dec));
    if (!uhdr_get_decoded_image(dec)) exit(2);
    armed = 0;
  }
  bad |= report("reset + set_image + uhdr_decode (half float)");
  uhdr_release_decoder(dec);
  uhdr_release_encoder(enc);
  return bad;
}

int main(int argc, char** argv) {
  void* warm[4];
  count_warmup = getenv("ALLOC_PROBE_COUNT_WARMUP") != NULL;
  backtrace(warm, 4);   /* loads libgcc's unwinder now, not inside the first counted call */
  if (argc < 2) { fprintf(stderr, "usage: alloc_probe api4 base.jpg gainmap.jpg | gpu W H\n"); return 2; }
  /* touch the library so that it is mapped, then find the address ranges */
  uhdr_codec_private_t* tmp = uhdr_create_decoder();
  uhdr_release_decoder(tmp);
  int bad;
  if (!strcmp(argv[1], "api4") && argc == 4) {
    dl_iterate_phdr(phdr_cb, NULL);
    bad = run_api4(argv[2], argv[3]);
  } else if (!strcmp(argv[1], "gpu") && argc == 4) {
    /* first CUDA use maps libcuda: do one throw-away encode before reading the ranges */
    dl_iterate_phdr(phdr_cb, NULL);
    bad = run_gpu(atoi(argv[2]), atoi(argv[3]));
  } else {
    return 2;
  }
  if (!ours_hi) { fprintf(stderr, "libuhdr_b200.so not found among the loaded objects\n"); return 2; }
  return bad ? 1 : 0;
}
