#include "container.h"

#include <cctype>
#include <cmath>
#include <cstring>
#include <sstream>
#include <string>

#include "runtime.h"

namespace uhdr_b200 {

#include "icc_blobs.inc"

// ---- rational conversion (gainmapmath.cpp:1616-1680), all intermediate math in double ---------
static bool to_unsigned_fraction(float v, uint32_t max_num, uint32_t* num, uint32_t* den) {
  if (std::isnan(v) || v < 0 || v > (float)max_num) return false;
  const uint64_t max_d = (v <= 1) ? UINT32_MAX : (uint64_t)std::floor((double)((float)max_num / v));
  *den = 1;
  uint32_t prev_d = 0;
  double cur = (double)v - std::floor((double)v);
  for (int it = 0; it < 39; it++) {
    const double nd = (double)(*den) * (double)v;
    if (nd > (double)max_num) return false;
    *num = (uint32_t)std::round(nd);
    if (std::fabs(nd - (double)(*num)) == 0.0) return true;
    cur = 1.0 / cur;
    const double new_d = (double)prev_d + std::floor(cur) * (double)(*den);
    if (new_d > (double)max_d) return true;
    prev_d = *den;
    if (new_d > (double)UINT32_MAX) return false;
    *den = (uint32_t)new_d;
    cur -= std::floor(cur);
  }
  *num = (uint32_t)std::round((double)(*den) * (double)v);
  return true;
}
static bool to_signed_fraction(float v, int32_t* num, uint32_t* den) {
  uint32_t pos;
  if (!to_unsigned_fraction(std::fabs(v), INT32_MAX, &pos, den)) return false;
  *num = (int32_t)pos;
  if (v < 0) *num *= -1;
  return true;
}

namespace {
struct Frac {
  int32_t min_n[3], max_n[3], base_off_n[3], alt_off_n[3];
  uint32_t min_d[3], max_d[3], gamma_n[3], gamma_d[3], base_off_d[3], alt_off_d[3];
  uint32_t base_headroom_n, base_headroom_d, alt_headroom_n, alt_headroom_d;
  bool backward, use_base;
};
void be16(ByteSink& o, unsigned v) { o.u16(v); }
void be32(ByteSink& o, uint32_t v) { o.u32(v); }
bool md_single(const uhdr_gainmap_metadata_t& m) {
  auto same = [](const float* a) { return a[0] == a[1] && a[0] == a[2]; };
  return same(m.max_content_boost) && same(m.min_content_boost) && same(m.gamma) && same(m.offset_sdr) && same(m.offset_hdr);
}
}  // namespace

int validate_metadata(const uhdr_gainmap_metadata_t& m) {  // ultrahdr_api.cpp:431-503
  int rc = E_OK;
  for (int i = 0; i < 3; i++) {
    if (!std::isfinite(m.min_content_boost[i]) || !std::isfinite(m.max_content_boost[i]) ||
        !std::isfinite(m.offset_sdr[i]) || !std::isfinite(m.offset_hdr[i]) || !std::isfinite(m.hdr_capacity_min) ||
        !std::isfinite(m.hdr_capacity_max) || !std::isfinite(m.gamma[i]))
      rc = fail(E_INVALID_PARAM, "Field(s) of gainmap metadata descriptor are either NaN or infinite.");
    else if (m.max_content_boost[i] < m.min_content_boost[i])
      rc = fail(E_INVALID_PARAM, "received bad value for content boost max %f, expects to be >= content boost min %f",
                m.max_content_boost[i], m.min_content_boost[i]);
    else if (m.min_content_boost[i] <= 0.0f)
      return fail(E_INVALID_PARAM, "received bad value for min boost %f, expects > 0.0f", m.min_content_boost[i]);
    else if (m.gamma[i] <= 0.0f)
      rc = fail(E_INVALID_PARAM, "received bad value for gamma %f, expects > 0.0f", m.gamma[i]);
    else if (m.offset_sdr[i] < 0.0f)
      rc = fail(E_INVALID_PARAM, "received bad value for offset sdr %f, expects to be >= 0.0f", m.offset_sdr[i]);
    else if (m.offset_hdr[i] < 0.0f)
      rc = fail(E_INVALID_PARAM, "received bad value for offset hdr %f, expects to be >= 0.0f", m.offset_hdr[i]);
    else if (m.hdr_capacity_max <= m.hdr_capacity_min)
      rc = fail(E_INVALID_PARAM, "received bad value for hdr capacity max %f, expects to be > hdr capacity min %f",
                m.hdr_capacity_max, m.hdr_capacity_min);
    else if (m.hdr_capacity_min < 1.0f)
      rc = fail(E_INVALID_PARAM, "received bad value for hdr capacity min %f, expects to be >= 1.0f", m.hdr_capacity_min);
  }
  return rc;
}

int iso_encode_metadata(const uhdr_gainmap_metadata_t& md, uint8_t* out, size_t cap, size_t* out_size) {
  Frac f;
  memset(&f, 0, sizeof f);
  f.backward = false;
  f.use_base = md.use_base_cg != 0;
  const bool single = md_single(md);
#define SFRAC(v, n, d) if (!to_signed_fraction((float)(v), n, d)) return fail(E_INVALID_PARAM, "encountered error while representing float %f as a rational number (p/q form) ", (double)(v))
#define UFRAC(v, n, d) if (!to_unsigned_fraction((float)(v), UINT32_MAX, n, d)) return fail(E_INVALID_PARAM, "encountered error while representing float %f as a rational number (p/q form) ", (double)(v))
  for (int i = 0; i < (single ? 1 : 3); i++) {  // double log2 narrowed to float (gainmapmetadata.cpp:388-399)
    SFRAC(std::log2((double)md.max_content_boost[i]), &f.max_n[i], &f.max_d[i]);
    SFRAC(std::log2((double)md.min_content_boost[i]), &f.min_n[i], &f.min_d[i]);
    UFRAC(md.gamma[i], &f.gamma_n[i], &f.gamma_d[i]);
    SFRAC(md.offset_sdr[i], &f.base_off_n[i], &f.base_off_d[i]);
    SFRAC(md.offset_hdr[i], &f.alt_off_n[i], &f.alt_off_d[i]);
  }
  if (single)
    for (int i = 1; i < 3; i++) {
      f.max_n[i] = f.max_n[0]; f.max_d[i] = f.max_d[0]; f.min_n[i] = f.min_n[0]; f.min_d[i] = f.min_d[0];
      f.gamma_n[i] = f.gamma_n[0]; f.gamma_d[i] = f.gamma_d[0];
      f.base_off_n[i] = f.base_off_n[0]; f.base_off_d[i] = f.base_off_d[0];
      f.alt_off_n[i] = f.alt_off_n[0]; f.alt_off_d[i] = f.alt_off_d[0];
    }
  UFRAC(std::log2((double)md.hdr_capacity_min), &f.base_headroom_n, &f.base_headroom_d);
  UFRAC(std::log2((double)md.hdr_capacity_max), &f.alt_headroom_n, &f.alt_headroom_d);
#undef SFRAC
#undef UFRAC
  // serialisation gainmapmetadata.cpp:113-193
  auto ident = [](const auto* a) { return a[0] == a[1] && a[0] == a[2]; };
  const bool one = ident(f.min_n) && ident(f.min_d) && ident(f.max_n) && ident(f.max_d) && ident(f.gamma_n) &&
                   ident(f.gamma_d) && ident(f.base_off_n) && ident(f.base_off_d) && ident(f.alt_off_n) && ident(f.alt_off_d);
  const int channels = one ? 1 : 3;
  ByteSink o(out, cap);
  be16(o, 0);
  be16(o, 0);
  uint8_t flags = 0;
  if (channels == 3) flags |= 0x80;
  if (f.use_base) flags |= 0x40;
  if (f.backward) flags |= 4;
  const uint32_t denom = f.base_headroom_d;
  bool common = f.alt_headroom_d == denom;
  for (int c = 0; c < channels; c++)
    if (f.min_d[c] != denom || f.max_d[c] != denom || f.gamma_d[c] != denom || f.base_off_d[c] != denom || f.alt_off_d[c] != denom)
      common = false;
  if (common) flags |= 8;
  o.u8(flags);
  if (common) {
    be32(o, denom);
    be32(o, f.base_headroom_n);
    be32(o, f.alt_headroom_n);
    for (int c = 0; c < channels; c++) {
      be32(o, (uint32_t)f.min_n[c]); be32(o, (uint32_t)f.max_n[c]); be32(o, f.gamma_n[c]);
      be32(o, (uint32_t)f.base_off_n[c]); be32(o, (uint32_t)f.alt_off_n[c]);
    }
  } else {
    be32(o, f.base_headroom_n); be32(o, f.base_headroom_d); be32(o, f.alt_headroom_n); be32(o, f.alt_headroom_d);
    for (int c = 0; c < channels; c++) {
      be32(o, (uint32_t)f.min_n[c]); be32(o, f.min_d[c]); be32(o, (uint32_t)f.max_n[c]); be32(o, f.max_d[c]);
      be32(o, f.gamma_n[c]); be32(o, f.gamma_d[c]); be32(o, (uint32_t)f.base_off_n[c]); be32(o, f.base_off_d[c]);
      be32(o, (uint32_t)f.alt_off_n[c]); be32(o, f.alt_off_d[c]);
    }
  }
  if (!o.ok()) return fail(E_MEM, "gain map metadata block of %zu bytes does not fit the %zu-byte buffer", o.size(), cap);
  *out_size = o.size();
  return E_OK;
}

int iso_decode_metadata(const uint8_t* d, size_t n, uhdr_gainmap_metadata_t* md) {
  size_t p = 0;
  auto need = [&](size_t k) { return p <= n && n - p >= k; };
  auto r16 = [&](unsigned& v) { if (!need(2)) return false; v = (d[p] << 8) | d[p + 1]; p += 2; return true; };
  auto r32 = [&](uint32_t& v) { if (!need(4)) return false; v = ((uint32_t)d[p] << 24) | (d[p + 1] << 16) | (d[p + 2] << 8) | d[p + 3]; p += 4; return true; };
#define RD(x) if (!(x)) return fail(E_MEM, "attempting to read past the end of the gain map metadata (size %d)", (int)n)
  unsigned minv = 0xffff, wv = 0xffff;
  RD(r16(minv));
  if (minv != 0) return fail(E_UNSUPPORTED, "received unexpected minimum version %d, expected 0", minv);
  RD(r16(wv));
  RD(need(1));
  const uint8_t flags = d[p++];
  const int channels = (flags & 0x80) ? 3 : 1;
  Frac f;
  memset(&f, 0, sizeof f);
  f.use_base = (flags & 0x40) != 0;
  f.backward = (flags & 4) != 0;
  uint32_t u;
  if (flags & 8) {
    uint32_t den = 1;
    RD(r32(den)); RD(r32(f.base_headroom_n)); f.base_headroom_d = den; RD(r32(f.alt_headroom_n)); f.alt_headroom_d = den;
    for (int c = 0; c < channels; c++) {
      RD(r32(u)); f.min_n[c] = (int32_t)u; f.min_d[c] = den;
      RD(r32(u)); f.max_n[c] = (int32_t)u; f.max_d[c] = den;
      RD(r32(f.gamma_n[c])); f.gamma_d[c] = den;
      RD(r32(u)); f.base_off_n[c] = (int32_t)u; f.base_off_d[c] = den;
      RD(r32(u)); f.alt_off_n[c] = (int32_t)u; f.alt_off_d[c] = den;
    }
  } else {
    RD(r32(f.base_headroom_n)); RD(r32(f.base_headroom_d)); RD(r32(f.alt_headroom_n)); RD(r32(f.alt_headroom_d));
    for (int c = 0; c < channels; c++) {
      RD(r32(u)); f.min_n[c] = (int32_t)u; RD(r32(f.min_d[c]));
      RD(r32(u)); f.max_n[c] = (int32_t)u; RD(r32(f.max_d[c]));
      RD(r32(f.gamma_n[c])); RD(r32(f.gamma_d[c]));
      RD(r32(u)); f.base_off_n[c] = (int32_t)u; RD(r32(f.base_off_d[c]));
      RD(r32(u)); f.alt_off_n[c] = (int32_t)u; RD(r32(f.alt_off_d[c]));
    }
  }
#undef RD
  for (int c = channels; c < 3; c++) {
    f.min_n[c] = f.min_n[0]; f.min_d[c] = f.min_d[0]; f.max_n[c] = f.max_n[0]; f.max_d[c] = f.max_d[0];
    f.gamma_n[c] = f.gamma_n[0]; f.gamma_d[c] = f.gamma_d[0]; f.base_off_n[c] = f.base_off_n[0];
    f.base_off_d[c] = f.base_off_d[0]; f.alt_off_n[c] = f.alt_off_n[0]; f.alt_off_d[c] = f.alt_off_d[0];
  }
  // gainmapMetadataFractionToFloat :301-347 (double exp2 of a float quotient, narrowed)
  if (!f.base_headroom_d || !f.alt_headroom_d) return fail(E_INVALID_PARAM, "received 0 (bad value) for field HdrHeadroom denominator");
  for (int i = 0; i < 3; i++)
    if (!f.max_d[i] || !f.gamma_d[i] || !f.min_d[i] || !f.base_off_d[i] || !f.alt_off_d[i])
      return fail(E_INVALID_PARAM, "received 0 (bad value) for a gain map metadata denominator");
  if (f.backward) return fail(E_UNSUPPORTED, "hdr intent as base rendition is not supported");
  for (int i = 0; i < 3; i++) {
    md->max_content_boost[i] = (float)std::exp2((double)((float)f.max_n[i] / f.max_d[i]));
    md->min_content_boost[i] = (float)std::exp2((double)((float)f.min_n[i] / f.min_d[i]));
    md->gamma[i] = (float)f.gamma_n[i] / f.gamma_d[i];
    md->offset_sdr[i] = (float)f.base_off_n[i] / f.base_off_d[i];
    md->offset_hdr[i] = (float)f.alt_off_n[i] / f.alt_off_d[i];
  }
  md->hdr_capacity_max = (float)std::exp2((double)((float)f.alt_headroom_n / f.alt_headroom_d));
  md->hdr_capacity_min = (float)std::exp2((double)((float)f.base_headroom_n / f.base_headroom_d));
  md->use_base_cg = f.use_base;
  return validate_metadata(*md);
}

const uint8_t* icc_profile(int ct, int cg, size_t* size) {
  if (ct < 0 || ct > 3 || cg < 0 || cg > 2) return nullptr;
  *size = kIccSizes[ct * 3 + cg];
  return kIccBlobs[ct * 3 + cg];
}

// ---- hdrgm XMP (Ultra HDR v1 / Apple) -----------------------------------------------------------
namespace {
// The reference feeds the packet to a SAX parser whose handler only looks at the rdf:Description
// element: its attributes (hdrgm:*) and, for Apple files, its child elements HDRGainMapVersion /
// HDRGainMapHeadroom.  The same information is collected here with a flat scan of the start tag.
struct XmpFields {
  std::string version, gmax, gmin, gamma, off_sdr, off_hdr, cap_min, cap_max, base_is_hdr;
  bool has_version = false, has_gmax = false, has_gmin = false, has_gamma = false, has_off_sdr = false,
       has_off_hdr = false, has_cap_min = false, has_cap_max = false, has_base_is_hdr = false;
  bool apple = false, found_description = false;
};
bool is_name_char(char c) { return isalnum((unsigned char)c) || c == ':' || c == '_' || c == '-' || c == '.'; }

void scan_xmp(const std::string& x, XmpFields* f) {
  size_t p = x.find("<rdf:Description");
  if (p == std::string::npos) return;
  f->found_description = true;
  p += 16;
  // attributes of the start tag
  bool self_closed = false;
  while (p < x.size()) {
    while (p < x.size() && isspace((unsigned char)x[p])) p++;
    if (p >= x.size()) return;
    if (x[p] == '>') { p++; break; }
    if (x[p] == '/' && p + 1 < x.size() && x[p + 1] == '>') { self_closed = true; p += 2; break; }
    size_t a = p;
    while (p < x.size() && is_name_char(x[p])) p++;
    if (p == a) { p++; continue; }
    const std::string name = x.substr(a, p - a);
    while (p < x.size() && isspace((unsigned char)x[p])) p++;
    if (p >= x.size() || x[p] != '=') continue;
    p++;
    while (p < x.size() && isspace((unsigned char)x[p])) p++;
    if (p >= x.size() || (x[p] != '"' && x[p] != '\'')) continue;
    const char quote = x[p++];
    a = p;
    while (p < x.size() && x[p] != quote) p++;
    const std::string val = x.substr(a, p - a);
    if (p < x.size()) p++;
    auto take = [&](const char* n, std::string* dst, bool* flag) {
      if (name == n) { *dst = val; *flag = true; }
    };
    take("hdrgm:Version", &f->version, &f->has_version);
    take("hdrgm:GainMapMax", &f->gmax, &f->has_gmax);
    take("hdrgm:GainMapMin", &f->gmin, &f->has_gmin);
    take("hdrgm:Gamma", &f->gamma, &f->has_gamma);
    take("hdrgm:OffsetSDR", &f->off_sdr, &f->has_off_sdr);
    take("hdrgm:OffsetHDR", &f->off_hdr, &f->has_off_hdr);
    take("hdrgm:HDRCapacityMin", &f->cap_min, &f->has_cap_min);
    take("hdrgm:HDRCapacityMax", &f->cap_max, &f->has_cap_max);
    take("hdrgm:BaseRenditionIsHDR", &f->base_is_hdr, &f->has_base_is_hdr);
  }
  if (self_closed) return;
  // child elements up to </rdf:Description>: Apple's <...:HDRGainMapVersion>n</...> and <...:HDRGainMapHeadroom>x</...>
  const size_t end = x.find("</rdf:Description", p);
  const std::string body = x.substr(p, end == std::string::npos ? std::string::npos : end - p);
  auto child_text = [&](const char* key, std::string* dst) {
    size_t q = 0;
    while ((q = body.find(key, q)) != std::string::npos) {
      // must be inside a start tag: the nearest '<' before it is not "</"
      const size_t lt = body.rfind('<', q);
      if (lt != std::string::npos && lt + 1 < body.size() && body[lt + 1] != '/') {
        const size_t gt = body.find('>', q);
        if (gt == std::string::npos) return false;
        const size_t lt2 = body.find('<', gt);
        *dst = body.substr(gt + 1, lt2 == std::string::npos ? std::string::npos : lt2 - gt - 1);
        return true;
      }
      q += strlen(key);
    }
    return false;
  };
  std::string v;
  if (child_text("HDRGainMapVersion", &v)) {
    f->version = v;
    f->has_version = true;
    f->apple = true;
  }
  if (child_text("HDRGainMapHeadroom", &v)) {
    f->gmax = v;
    f->has_gmax = true;
  }
}
// `stringstream >> float` of the reference
bool parse_float(const std::string& str, float* out) {
  std::stringstream ss(str);
  float v;
  if (ss >> v) { *out = v; return true; }
  return false;
}

bool rd16(const uint8_t* d, size_t n, uint16_t* v, size_t* off, bool be) {
  if (*off > n || n - *off < 2) return false;
  *v = be ? (uint16_t)((d[*off] << 8) | d[*off + 1]) : (uint16_t)(d[*off] | (d[*off + 1] << 8));
  *off += 2;
  return true;
}
bool rd32(const uint8_t* d, size_t n, uint32_t* v, size_t* off, bool be) {
  if (*off > n || n - *off < 4) return false;
  *v = be ? ((uint32_t)d[*off] << 24) | ((uint32_t)d[*off + 1] << 16) | ((uint32_t)d[*off + 2] << 8) | d[*off + 3]
          : (uint32_t)d[*off] | ((uint32_t)d[*off + 1] << 8) | ((uint32_t)d[*off + 2] << 16) | ((uint32_t)d[*off + 3] << 24);
  *off += 4;
  return true;
}
// getExifAppleHeadroom, jpegrutils.cpp:506-644: Apple maker-note tags 33 and 48 -> headroom
bool exif_apple_headroom(const uint8_t* exif, size_t size, float* headroom) {
  *headroom = 0.0f;
  size_t offset = 0;
  if (size < 6 || memcmp(exif, "Exif\0\0", 6) != 0) {
    bool found = false;
    for (size_t i = 0; i + 4 <= size; i++)
      if ((exif[i] == 'I' && exif[i + 1] == 'I' && exif[i + 2] == 0x2A && exif[i + 3] == 0) ||
          (exif[i] == 'M' && exif[i + 1] == 'M' && exif[i + 2] == 0 && exif[i + 3] == 0x2A)) {
        offset = i;
        found = true;
        break;
      }
    if (!found) return false;
  } else {
    offset = 6;
  }
  if (offset + 4 > size) return false;
  bool be = exif[offset] == 'M';
  offset += 4;
  uint32_t ifd;
  if (!rd32(exif, size, &ifd, &offset, be)) return false;
  static const uint8_t kAppleHdr[] = {'A', 'p', 'p', 'l', 'e', ' ', 'i', 'O', 'S', 0, 0, 1, 'M', 'M'};
  bool in_apple = false, has = false;
  double m33 = 0.0, m48 = 0.0;
  int nifd = 0;
  const size_t tiff = offset - 8;
  while (ifd != 0 && nifd++ < 3) {
    offset = tiff + ifd;
    bool next_set = false;
    uint16_t count;
    if (!rd16(exif, size, &count, &offset, be)) return false;
    for (uint16_t fidx = 0; fidx < count; ++fidx) {
      uint16_t tag, fmt;
      uint32_t ncomp, data;
      if (!rd16(exif, size, &tag, &offset, be) || !rd16(exif, size, &fmt, &offset, be) ||
          !rd32(exif, size, &ncomp, &offset, be) || !rd32(exif, size, &data, &offset, be))
        return false;
      if (tag == 0x8769) {
        ifd = data;
        next_set = true;
        break;
      } else if (tag == 0x927c) {
        if (tiff + data + sizeof kAppleHdr <= size && !memcmp(exif + tiff + data, kAppleHdr, sizeof kAppleHdr)) {
          ifd = data + (uint32_t)sizeof kAppleHdr;
          in_apple = true;
          next_set = true;
          be = true;
          break;
        }
      } else if (in_apple && (tag == 33 || tag == 48) && fmt == 10) {
        if (tiff + ifd < sizeof kAppleHdr) return false;
        size_t t = tiff + ifd - sizeof kAppleHdr;
        if (t > SIZE_MAX - (size_t)data) return false;
        t += data;
        uint32_t num, den;
        if (!rd32(exif, size, &num, &t, be) || !rd32(exif, size, &den, &t, be)) return false;
        if (den != 0) {
          const double v = (double)(int32_t)num / den;
          (tag == 33 ? m33 : m48) = v;
          has = true;
        }
      }
    }
    if (!next_set && !rd32(exif, size, &ifd, &offset, be)) return false;
  }
  if (!has) return false;
  double stops;
  if (m33 < 1.0) stops = m48 <= 0.01 ? -20.0 * m48 + 1.8 : -0.101 * m48 + 1.601;
  else stops = m48 <= 0.01 ? -70.0 * m48 + 3.0 : -0.303 * m48 + 2.303;
  *headroom = (float)pow(2.0, stops);
  return true;
}
}  // namespace

int xmp_decode_metadata(const uint8_t* xmp, size_t n, const uint8_t* exif, size_t exif_n, uhdr_gainmap_metadata_t* md) {
  static const char kNs[] = "http://ns.adobe.com/xap/1.0/";  // 28 chars + NUL in the marker
  const size_t ns = sizeof kNs - 1;
  if (n < ns + 2) return fail(E_ERROR, "size of xmp block is expected to be atleast %zd bytes, received only %zd bytes", ns + 2, n);
  if (strncmp((const char*)xmp, kNs, ns)) return fail(E_ERROR, "mismatch in namespace of xmp block. Expected %s", kNs);
  const std::string x((const char*)xmp + ns + 1, n - ns - 1);
  XmpFields f;
  scan_xmp(x, &f);
  float v;
  if (f.apple) {  // jpegrutils.cpp:723-760
    for (int c = 0; c < 3; c++) {
      md->gamma[c] = 1.0f;
      md->min_content_boost[c] = 1.0f;
      md->offset_sdr[c] = md->offset_hdr[c] = 0.0f;
    }
    md->hdr_capacity_min = 1.0f;
    float boost;
    if (f.has_gmax && parse_float(f.gmax, &v)) boost = exp2(v);
    else if (!(exif && exif_n > 0 && exif_apple_headroom(exif, exif_n, &boost)))
      return fail(E_ERROR, "xml parse error, could not find attribute HDRGainMapHeadroom and Exif Headroom missing");
    for (int c = 0; c < 3; c++) md->max_content_boost[c] = boost;
    md->hdr_capacity_max = boost;
    // the reference leaves use_base_cg of its (uninitialised) descriptor untouched on this branch; any
    // non-zero garbage reads as true, which is also what the non-Apple XMP branch sets
    md->use_base_cg = 1;
    return E_OK;
  }
  // required: Version, GainMapMax, HDRCapacityMax; the others default (:762-873).  exp2 is the double
  // function narrowed to float, like the reference's (jpegrutils.cpp has `using namespace std`: float
  // overload, same value for float arguments)
  if (!f.has_version) return fail(E_ERROR, "xml parse error, could not find attribute %s", "hdrgm:Version");
  if (!(f.has_gmax && parse_float(f.gmax, &v))) return fail(E_ERROR, "xml parse error, could not find attribute %s", "hdrgm:GainMapMax");
  md->max_content_boost[0] = std::exp2(v);
  if (!(f.has_cap_max && parse_float(f.cap_max, &v))) return fail(E_ERROR, "xml parse error, could not find attribute %s", "hdrgm:HDRCapacityMax");
  md->hdr_capacity_max = std::exp2(v);
  auto opt = [&](bool has, const std::string& str, const char* name, float dflt, bool exp, float* dst) -> int {
    float t;
    if (parse_float(str, &t)) { *dst = exp ? std::exp2(t) : t; return E_OK; }
    if (has) return fail(E_ERROR, "xml parse error, unable to parse attribute %s", name);
    *dst = dflt;
    return E_OK;
  };
  int rc;
  if ((rc = opt(f.has_gmin, f.gmin, "hdrgm:GainMapMin", 1.0f, true, &md->min_content_boost[0]))) return rc;
  if ((rc = opt(f.has_gamma, f.gamma, "hdrgm:Gamma", 1.0f, false, &md->gamma[0]))) return rc;
  if ((rc = opt(f.has_off_sdr, f.off_sdr, "hdrgm:OffsetSDR", 1.0f / 64.0f, false, &md->offset_sdr[0]))) return rc;
  if ((rc = opt(f.has_off_hdr, f.off_hdr, "hdrgm:OffsetHDR", 1.0f / 64.0f, false, &md->offset_hdr[0]))) return rc;
  if ((rc = opt(f.has_cap_min, f.cap_min, "hdrgm:HDRCapacityMin", 1.0f, true, &md->hdr_capacity_min))) return rc;
  if (f.has_base_is_hdr) {
    if (f.base_is_hdr == "True") return fail(E_ERROR, "hdr intent as base rendition is not supported");
    if (f.base_is_hdr != "False") return fail(E_ERROR, "xml parse error, unable to parse attribute %s", "hdrgm:BaseRenditionIsHDR");
  }
  md->use_base_cg = 1;
  for (int c = 1; c < 3; c++) {
    md->min_content_boost[c] = md->min_content_boost[0];
    md->max_content_boost[c] = md->max_content_boost[0];
    md->gamma[c] = md->gamma[0];
    md->offset_hdr[c] = md->offset_hdr[0];
    md->offset_sdr[c] = md->offset_sdr[0];
  }
  return validate_metadata(*md);
}

int parse_gainmap_metadata(const uint8_t* iso, size_t iso_n, const uint8_t* xmp, size_t xmp_n, const uint8_t* exif,
                           size_t exif_n, uhdr_gainmap_metadata_t* md) {
  static const size_t kIsoNs = 28;  // "urn:iso:std:iso:ts:21496:-1" + NUL
  if (iso_n > 0) {
    if (iso_n < kIsoNs) return fail(E_ERROR, "iso block size needs to be atleast %zd but got %zd", kIsoNs, iso_n);
    return iso_decode_metadata(iso + kIsoNs, iso_n - kIsoNs, md);
  }
  if (xmp_n > 0) return xmp_decode_metadata(xmp, xmp_n, exif, exif_n, md);
  return fail(E_INVALID_PARAM, "received no valid buffer to parse gainmap metadata");
}

int icc_read_gamut(const uint8_t* d, size_t n) {
  const size_t kPrefix = 14, kHeader = 132;
  if (!d || n < kHeader + kPrefix || memcmp(d, "ICC_PROFILE", 12) != 0) return UHDR_CG_UNSPECIFIED;
  const uint8_t* icc = d + kPrefix;
  const size_t psize = n - kPrefix;
  auto be = [](const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; };
  const size_t tags = be(icc + 128);
  if (tags > (psize - kHeader) / 12) return UHDR_CG_UNSPECIFIED;
  size_t off[4] = {0, 0, 0, 0}, len[4] = {0, 0, 0, 0};  // rXYZ gXYZ bXYZ cicp
  static const char* sig[4] = {"rXYZ", "gXYZ", "bXYZ", "cicp"};
  for (size_t t = 0; t < tags; t++) {
    const uint8_t* e = icc + kHeader + t * 12;
    for (int k = 0; k < 4; k++)
      if (off[k] == 0 && !memcmp(e, sig[k], 4)) { off[k] = be(e + 4); len[k] = be(e + 8); break; }
  }
  if (off[3] && len[3] == 12 && off[3] <= psize && len[3] <= psize - off[3]) {
    const uint8_t prim = icc[off[3] + 8];
    if (prim == 1) return UHDR_CG_BT_709;
    if (prim == 12) return UHDR_CG_DISPLAY_P3;
    if (prim == 9) return UHDR_CG_BT_2100;
  }
  for (int k = 0; k < 3; k++)
    if (off[k] == 0 || len[k] != 20 || off[k] > psize || len[k] > psize - off[k]) return UHDR_CG_UNSPECIFIED;
  // colorant matrices icc.h:125-145
  const float F = 1.52587890625e-5f;
  const float mats[3][3][3] = {
      {{0x6FA2 * F, 0x6299 * F, 0x24A0 * F}, {0x38F5 * F, 0xB785 * F, 0x0F84 * F}, {0x0390 * F, 0x18DA * F, 0xB6CF * F}},
      {{0.515102f, 0.291965f, 0.157153f}, {0.241182f, 0.692236f, 0.0665819f}, {-0.00104941f, 0.0418818f, 0.784378f}},
      {{0.673459f, 0.165661f, 0.125100f}, {0.279033f, 0.675338f, 0.0456288f}, {-0.00193139f, 0.0299794f, 0.797162f}}};
  for (int g = 0; g < 3; g++) {
    bool ok = true;
    for (int col = 0; col < 3 && ok; col++)
      for (int row = 0; row < 3; row++) {
        const float v = (float)(int32_t)be(icc + off[col] + 8 + 4 * row) * F;
        if (std::fabs(v - mats[g][row][col]) > 0.001f) { ok = false; break; }
      }
    if (ok) return g;
  }
  return UHDR_CG_UNSPECIFIED;
}

// ---- MPF (multipictureformat.cpp) ---------------------------------------------------------------
constexpr size_t kMpfBytes = 86;   // MPF signature + TIFF header + index IFD (3 tags) + two 16-byte MP entries
static void make_mpf(size_t primary_size, size_t secondary_size, size_t secondary_offset, uint8_t out[kMpfBytes]) {
  ByteSink o(out, kMpfBytes);
  static const uint8_t head[8] = {'M', 'P', 'F', 0, 0x4D, 0x4D, 0x00, 0x2A};
  o.raw(head, 8);
  be32(o, 8);            // index IFD offset
  be16(o, 3);            // tag count
  be16(o, 0xB000); be16(o, 7); be32(o, 4); o.raw("0100", 4);
  be16(o, 0xB001); be16(o, 4); be32(o, 1); be32(o, 2);
  be16(o, 0xB002); be16(o, 7); be32(o, 32);
  be32(o, (uint32_t)(o.size() - 4 + 4 + 4));  // MP entry offset
  be32(o, 0);                                  // attribute IFD offset
  be32(o, 0x030000); be32(o, (uint32_t)primary_size); be32(o, 0); be16(o, 0); be16(o, 0);
  be32(o, 0); be32(o, (uint32_t)secondary_size); be32(o, (uint32_t)secondary_offset); be16(o, 0); be16(o, 0);
}

int assemble_jpegr(const JpegPieces& base, const JpegPieces& gm, const uint8_t* exif, size_t exif_size,
                   const uhdr_gainmap_metadata_t& md, uint8_t* out, size_t cap, size_t* out_size,
                   const uint8_t* icc_arg, size_t icc_arg_size) {
  static const char kIsoNs[] = "urn:iso:std:iso:ts:21496:-1";  // 27 chars + NUL
  const size_t ns_len = sizeof kIsoNs;
  uint8_t iso[kIsoMetadataMaxBytes];   // no heap on this path: fixed-size blocks, everything else goes straight to `out`
  size_t iso_n = 0;
  int rc = iso_encode_metadata(md, iso, sizeof iso, &iso_n);
  if (rc) return rc;
  const size_t iso_secondary_len = 2 + ns_len + iso_n;
  const size_t secondary_size = gm.total() + 2 + iso_secondary_len;
  size_t pos = 0;
  auto put = [&](const void* p, size_t n) {
    if (pos + n > cap) return false;
    memcpy(out + pos, p, n);
    pos += n;
    return true;
  };
  auto put_marker = [&](uint8_t m, size_t payload_len) {
    const uint8_t h[4] = {0xFF, m, (uint8_t)(((payload_len + 2) >> 8) & 0xff), (uint8_t)((payload_len + 2) & 0xff)};
    return put(h, 4);
  };
#define W(x) if (!(x)) return fail(E_MEM, "output buffer of %zu bytes is too small for the encoded stream", cap)
  const uint8_t soi[2] = {0xFF, 0xD8}, eoi[2] = {0xFF, 0xD9};
  W(put(soi, 2));
  const uint8_t* b = base.head;
  const size_t bn = base.head_len;
  size_t bp = 2;
  if (bn >= 6 && b[2] == 0xFF && b[3] == 0xE0) {  // JFIF first (:1225-1237)
    const size_t l = (b[4] << 8) | b[5];
    if (2 + 2 + l <= bn) W(put(b + 2, 2 + l));
    bp += 2 + l;
  }
  // EXIF and ICC carried by the base stream are re-emitted here (jpegr.cpp:1173-1217,1239-1275)
  const uint8_t* icc = nullptr;
  const uint8_t* base_exif = nullptr;
  size_t icc_len = 0, base_exif_len = 0;
  for (size_t q = 2; q + 4 <= bn && b[q] == 0xFF && b[q + 1] != 0xDA;) {
    const size_t l = (b[q + 2] << 8) | b[q + 3];
    if (l < 2 || q + 2 + l > bn) break;
    if (b[q + 1] == 0xE2 && l > 2 + 12 && !memcmp(b + q + 4, "ICC_PROFILE", 12) && !icc) { icc = b + q + 4; icc_len = l - 2; }
    if (b[q + 1] == 0xE1 && l > 2 + 6 && !memcmp(b + q + 4, "Exif\0\0", 6) && !base_exif) { base_exif = b + q + 4; base_exif_len = l - 2; }
    q += 2 + l;
  }
  if (base_exif) {
    if (exif && exif_size)
      return fail(E_INVALID_PARAM, "received exif from uhdr_enc_set_exif_data() while the base image intent already contains "
                  "exif, unsure which one to use");
    exif = base_exif;
    exif_size = base_exif_len;
  }
  if (exif && exif_size) { W(put_marker(0xE1, exif_size)); W(put(exif, exif_size)); }
  if (icc_arg && icc_arg_size) { icc = icc_arg; icc_len = icc_arg_size; }
  if (icc) { W(put_marker(0xE2, icc_len)); W(put(icc, icc_len)); }
  {  // ISO version-only block (:1277-1292)
    const uint8_t zeros[4] = {0, 0, 0, 0};
    W(put_marker(0xE2, ns_len + 4)); W(put(kIsoNs, ns_len)); W(put(zeros, 4));
  }
  size_t sos = 0;
  while (bp < bn) {  // DQT / SOF / DHT up to SOS, APPn dropped (:1313-1336)
    if (b[bp] != 0xFF || bp + 1 >= bn) break;
    const uint8_t m = b[bp + 1];
    if (m == 0xDA) { sos = bp; break; }
    if (m == 0x00 || m == 0xFF || (m >= 0xD0 && m <= 0xD7)) { bp += 2; continue; }
    if (m == 0xD9 || bp + 4 > bn) break;
    const size_t l = (b[bp + 2] << 8) | b[bp + 3];
    if (bp + 2 + l > bn) break;
    if (!(m >= 0xE0 && m <= 0xEF)) W(put(b + bp, 2 + l));
    bp += 2 + l;
  }
  if (!sos) return fail(E_INVALID_PARAM, "SOS marker not found while reordering base jpeg segments, unable to append gainmap");
  {
    const size_t mpf_len = 2 + kMpfBytes;
    const size_t tail = (bn - sos) + base.scan_len + (base.whole ? 0 : 2);  // SOS header + entropy-coded data + EOI
    const size_t primary_size = pos + 2 + mpf_len + tail;
    const size_t secondary_offset = primary_size - pos - 8;
    uint8_t mpf[kMpfBytes];
    make_mpf(primary_size, secondary_size, secondary_offset, mpf);
    W(put_marker(0xE2, sizeof mpf)); W(put(mpf, sizeof mpf));
  }
  W(put(b + sos, bn - sos));
  if (base.scan_len) W(put(base.scan, base.scan_len));
  if (!base.whole) W(put(eoi, 2));
  W(put(soi, 2));
  W(put_marker(0xE2, ns_len + iso_n)); W(put(kIsoNs, ns_len)); W(put(iso, iso_n));
  W(put(gm.head + 2, gm.head_len - 2));
  if (gm.scan_len) W(put(gm.scan, gm.scan_len));
  if (!gm.whole) W(put(eoi, 2));
#undef W
  *out_size = pos;
  return E_OK;
}

// ---- splitter: what image_io's JpegScanner + JpegInfoBuilder(limit 2) report -----------------------
static bool scan_one(const uint8_t* d, size_t n, size_t start, size_t* end) {
  size_t p = start + 2;
  while (p + 2 <= n) {
    if (d[p] != 0xFF) return false;
    const uint8_t m = d[p + 1];
    if (m == 0xFF) { p++; continue; }
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
    if (m == 0xD9) { *end = p + 2; return true; }
    if (p + 4 > n) return false;
    const size_t l = (d[p + 2] << 8) | d[p + 3];
    if (l < 2 || p + 2 + l > n) return false;
    p += 2 + l;
    if (m == 0xDA) {  // entropy-coded data up to the next real marker (hop from 0xFF to 0xFF)
      while (p + 1 < n) {
        const uint8_t* ff = static_cast<const uint8_t*>(memchr(d + p, 0xFF, n - 1 - p));
        if (!ff) { p = n - 1; break; }
        p = (size_t)(ff - d);
        if (d[p + 1] != 0x00 && !(d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7) && d[p + 1] != 0xFF) break;
        p++;
      }
    }
  }
  return false;
}
int count_jpeg_images(const uint8_t* d, size_t n, size_t* first_off, size_t* first_len) {
  int found = 0;
  size_t p = 0;
  while (p + 2 <= n) {
    if (d[p] == 0xFF && d[p + 1] == 0xD8) {
      size_t e;
      if (!scan_one(d, n, p, &e)) return found ? found : -1;
      if (!found) {
        if (first_off) *first_off = p;
        if (first_len) *first_len = e - p;
      }
      found++;
      p = e;
    } else {
      p++;
    }
  }
  return found;
}

int split_jpegr(const uint8_t* d, size_t n, size_t* po, size_t* pl, size_t* go, size_t* gl) {
  size_t found = 0, p = 0, off[2], len[2];
  while (found < 2 && p + 2 <= n) {
    if (d[p] == 0xFF && d[p + 1] == 0xD8) {
      size_t e;
      if (!scan_one(d, n, p, &e)) break;
      off[found] = p;
      len[found] = e - p;
      found++;
      p = e;
    } else {
      p++;
    }
  }
  if (found == 0) return fail(E_INVALID_PARAM, "input uhdr image does not contain any valid images");
  *po = off[0];
  *pl = len[0];
  if (found == 1) return fail(E_INVALID_PARAM, "input uhdr image does not contain gainmap image");
  *go = off[1];
  *gl = len[1];
  return E_OK;
}

}  // namespace uhdr_b200
