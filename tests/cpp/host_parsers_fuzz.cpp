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
