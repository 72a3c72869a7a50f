// ASAN/UBSAN harness for the host-side code of the decoder: container split, JPEG header walk, ISO 21496-1
// metadata, ICC gamut read-out and the host entropy decoder, fed with mutated copies of a valid JPEG/R
// file (exact-size heap copies, so any over-read trips the sanitizer).  Built and run by
// tests/test_probe_cpu.py::test_host_parsers_under_sanitizers:
//   g++ -fsanitize=address,undefined -I libultrahdr_b200/csrc -I include -I $CUDA/include \
//       tests/cpp/host_parsers_fuzz.cpp libultrahdr_b200/csrc/container.cpp libultrahdr_b200/csrc/jpeg_host.cpp
//   ./a.out seed.jpg <rng seed> <iterations>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "container.h"
#include "jpeg.h"

namespace uhdr_b200 {
int fail(int code, const char*, ...) { return code; }
void set_last_error(const std::string&) {}
const char* last_error() { return ""; }
}
using namespace uhdr_b200;

static void probe(const uint8_t* d, size_t n) {
  size_t po, pl, go, gl;
  if (split_jpegr(d, n, &po, &pl, &go, &gl)) return;
  JpegHeader ph, gh;
  if (jpeg_read_header(d + po, pl, &ph)) return;
  if (jpeg_read_header(d + go, gl, &gh)) return;
  for (int which = 0; which < 2; which++) {   // host entropy decoder on both scans
    const JpegHeader& hh = which ? gh : ph;
    const uint8_t* jd = d + (which ? go : po);
    const size_t jn = which ? gl : pl;
    const JpegFrame& fr = hh.frame;
    if (fr.ncomp < 1 || fr.ncomp > 3 || fr.total_blocks() > (1u << 20)) continue;
    std::vector<std::vector<int16_t>> store(3);
    int16_t* coefs[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < fr.ncomp; c++) { store[c].resize(fr.blocks(c) * 64 + 64); coefs[c] = store[c].data(); }
    jpeg_host_decode_coefs(jd, jn, hh, coefs);
  }
  for (auto& m : gh.markers) {
    if (m.id == 0xE2 && m.length > 28 && m.offset + m.length <= gl && !memcmp(d + go + m.offset, "urn:iso:std:iso:ts:21496:-1", 28)) {
      uhdr_gainmap_metadata_t md;
      iso_decode_metadata(d + go + m.offset + 28, m.length - 28, &md);
    }
    if (m.id == 0xE2 && m.length > 14 && m.offset + m.length <= gl && !memcmp(d + go + m.offset, "ICC_PROFILE", 12))
      icc_read_gamut(d + go + m.offset, m.length);
  }
}

// one plain JPEG, unmutated: header walk + host entropy decoder (regression inputs, e.g. a stream whose
// DHT carries a malformed table that no scan component selects)
static int single(const uint8_t* d, size_t n) {
  JpegHeader h;
  if (jpeg_read_header(d, n, &h)) return 1;
  const JpegFrame& fr = h.frame;
  if (fr.ncomp < 1 || fr.ncomp > 3 || fr.total_blocks() > (1u << 20)) return 2;
  std::vector<std::vector<int16_t>> store(3);
  int16_t* coefs[3] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < fr.ncomp; c++) { store[c].resize(fr.blocks(c) * 64 + 64); coefs[c] = store[c].data(); }
  return jpeg_host_decode_coefs(d, n, h, coefs) ? 3 : 0;
}

// the heap-free writers with exact-size heap buffers: ISO 21496-1 block and JPEG/R assembly at every capacity from 0 up
// to what they need (too small: an error, never a write past the end), and a header with more APPn markers than the
// fixed marker list keeps
static int writers(const uint8_t* d, size_t n) {
  size_t po, pl, go, gl;
  if (split_jpegr(d, n, &po, &pl, &go, &gl)) return 1;
  uhdr_gainmap_metadata_t md;
  memset(&md, 0, sizeof md);
  for (int i = 0; i < 3; i++) {
    md.max_content_boost[i] = 3.5f + i; md.min_content_boost[i] = 0.9f; md.gamma[i] = 1.0f + 0.1f * i;
    md.offset_sdr[i] = md.offset_hdr[i] = 1.0f / 64;
  }
  md.hdr_capacity_min = 1.0f; md.hdr_capacity_max = 5.5f; md.use_base_cg = 1;
  size_t need = 0;
  {
    uint8_t big[kIsoMetadataMaxBytes];
    if (iso_encode_metadata(md, big, sizeof big, &need)) return 2;
  }
  int refused = 0;
  for (size_t cap = 0; cap <= need; cap++) {
    uint8_t* b = (uint8_t*)malloc(cap ? cap : 1);
    size_t got = 0;
    const int rc = iso_encode_metadata(md, b, cap, &got);
    if (rc) refused++;
    else if (cap != need || got != need) return 3;
    free(b);
  }
  if (refused != (int)need) return 4;
  JpegPieces pb, pg;
  pb.head = d + po; pb.head_len = pl; pb.scan = nullptr; pb.scan_len = 0; pb.whole = true;
  pg.head = d + go; pg.head_len = gl; pg.scan = nullptr; pg.scan_len = 0; pg.whole = true;
  std::vector<uint8_t> full(n + 4096);
  size_t out_n = 0;
  if (assemble_jpegr(pb, pg, nullptr, 0, md, full.data(), full.size(), &out_n)) return 5;
  for (size_t cap : {(size_t)0, (size_t)1, (size_t)100, out_n / 2, out_n - 1, out_n}) {
    uint8_t* b = (uint8_t*)malloc(cap ? cap : 1);
    size_t got = 0;
    const int rc = assemble_jpegr(pb, pg, nullptr, 0, md, b, cap, &got);
    if ((cap < out_n) != (rc != 0)) return 6;
    if (!rc && memcmp(b, full.data(), out_n)) return 7;
    free(b);
  }
  // 100 APP11 markers in front of a header: the list keeps its first kMax, the walk still finds SOF / SOS
  std::vector<uint8_t> many(d + po, d + po + 2);
  for (int i = 0; i < 100; i++) { const uint8_t m[8] = {0xFF, 0xEB, 0x00, 0x06, 'a', 'b', 'c', (uint8_t)i}; many.insert(many.end(), m, m + 8); }
  many.insert(many.end(), d + po + 2, d + po + pl);
  uint8_t* mp = (uint8_t*)malloc(many.size());
  memcpy(mp, many.data(), many.size());
  JpegHeader h;
  const int hrc = jpeg_read_header(mp, many.size(), &h);
  const bool ok = !hrc && h.markers.size() <= (size_t)JpegMarkerList::kMax && h.frame.width > 0;
  free(mp);
  return ok ? 0 : 8;
}

int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  std::vector<uint8_t> good;
  int c;
  while ((c = fgetc(f)) != EOF) good.push_back((uint8_t)c);
  fclose(f);
  if (argc > 2 && !strcmp(argv[2], "single")) {
    uint8_t* p = (uint8_t*)malloc(good.size());
    memcpy(p, good.data(), good.size());
    const int rc = single(p, good.size());
    free(p);
    printf("single rc=%d\nharness done\n", rc);
    return 0;
  }
  if (argc > 2 && !strcmp(argv[2], "writers")) {
    uint8_t* p = (uint8_t*)malloc(good.size());
    memcpy(p, good.data(), good.size());
    const int rc = writers(p, good.size());
    free(p);
    printf("writers rc=%d\nharness done\n", rc);
    return 0;
  }
  std::mt19937 rs(atoi(argv[2]));
  const size_t n = good.size();
  for (int it = 0; it < atoi(argv[3]); it++) {
    std::vector<uint8_t> bad = good;
    switch (it % 4) {
      case 0: for (int k = 0; k < 1 + (int)(rs() % 5); k++) bad[rs() % n] = (uint8_t)rs(); break;
      case 1: bad.resize(1 + rs() % n); break;
      case 2: { size_t a = rs() % n, b = std::min(n, a + 1 + rs() % 64); bad.erase(bad.begin() + a, bad.begin() + b); break; }
      default: { size_t a = rs() % std::min<size_t>(n, 1200); for (int k = 0; k < 1 + (int)(rs() % 4); k++) bad[std::min(bad.size() - 1, a + rs() % 32)] = (uint8_t)rs(); }
    }
    // exact-size heap copy so that any over-read trips ASAN
    uint8_t* p = (uint8_t*)malloc(bad.size());
    memcpy(p, bad.data(), bad.size());
    probe(p, bad.size());
    free(p);
  }
  puts("harness done");
  return 0;
}
