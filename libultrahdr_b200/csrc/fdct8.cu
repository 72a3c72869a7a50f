// Forward block stage: 8 lanes per 8x8 block.
//   pass 1: lane (b, r) loads row r of block b (8 bytes; RGB: 24 bytes + jccolor.c conversion),
//           runs the 1-D islow DCT on it in registers and parks the row in a warp-private,
//           bank-padded shared-memory tile
//   pass 2: lane (b, c) reads column c, runs the 1-D DCT, quantises with the exact reciprocal
//           (floor(a/d) == umulhi(a, ceil(2^32/d)) while a*d < 2^32), scatters the 8 coefficients
//           to their (zigzag) positions in the tile
//   store : lane (b, j) writes 16 bytes; a warp writes 4 blocks = 512 contiguous bytes
// A warp owns 4 horizontally adjacent blocks, a CTA 32; all planes of an image go in ONE launch
// (grid.z), the three components of an RGB888 gain map are produced from one read of the pixels.
// Arithmetic is libjpeg-turbo's jccolor.c / jfdctint.c / jcdctmgr.c integer arithmetic: bit-exact.
#include <cstring>
#include <mutex>

#include "kernels.cuh"
#include "runtime.h"

namespace uhdr_b200 {

namespace {

#define C_BITS 13
#define P1_BITS 2
#define DESC(x, n) (((x) + (1 << ((n)-1))) >> (n))

template <int PASS>
__device__ __forceinline__ void dct1d(int d[8]) {
  const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
  const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  constexpr int sh = PASS == 0 ? C_BITS - P1_BITS : C_BITS + P1_BITS;
  if (PASS == 0) {
    d[0] = (tmp10 + tmp11) << P1_BITS;
    d[4] = (tmp10 - tmp11) << P1_BITS;
  } else {
    d[0] = DESC(tmp10 + tmp11, P1_BITS);
    d[4] = DESC(tmp10 - tmp11, P1_BITS);
  }
  int z1 = (tmp12 + tmp13) * 4433;
  d[2] = DESC(z1 + tmp13 * 6270, sh);
  d[6] = DESC(z1 + tmp12 * (-15137), sh);
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * 9633;
  const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 = z3 * (-16069) + z5;
  z4 = z4 * (-3196) + z5;
  d[7] = DESC(t4 + z1 + z3, sh);
  d[5] = DESC(t5 + z2 + z4, sh);
  d[3] = DESC(t6 + z2 + z3, sh);
  d[1] = DESC(t7 + z1 + z4, sh);
}

// natural index -> zigzag position
__device__ const uint8_t kUnzigTab[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                                          41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                                          46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
__device__ __forceinline__ int unzig_rt(int n) { return kUnzigTab[n]; }

constexpr int kTileStride = 72;  // ints per block tile: 64 + 8 padding (4 blocks of a warp on distinct banks)

// Entropy-coder front end (jchuff.c encode_one_block), run while the quantised block is still in the
// warp's shared-memory tile (zigzag order).  Per block it leaves
//   * "meta", one uint4: x = number of code bits of its AC part (run/size Huffman codes, magnitude
//     bits, a ZRL per 16 zeros, EOB unless coefficient 63 is non-zero) << 16 | the DC value; y, z, w =
//     the first 96 bits of the AC part's finished bit string (EOB included), MSB first;
//   * bit strings longer than 96 bits: all their words in the block's slot.
// huffman.cu then only prepends the DC code (which needs the neighbouring block) and concatenates bit
// strings; coefficients are never stored for the device path.
// Step 1, lane j of the block's 8 lanes takes zigzag positions j, j+8, ... (the few non-zeros of a
// typical block sit at the lowest positions: they spread over the lanes): code word of each non-zero
// coefficient -> ent[rank].  Step 2, lane j takes entries j, j+8, ...: prefix sum of the lengths over
// the 8 lanes -> bit offset -> OR into the block's bit-string image.
__device__ __forceinline__ void bs_place(uint32_t* bs, unsigned off, unsigned long long pat, unsigned plen) {
  const unsigned long long v = pat << (64 - plen);  // left aligned
  const unsigned A = (unsigned)(v >> 32), B = (unsigned)v;
  const unsigned w = off >> 5, sh = off & 31;
  const unsigned x0 = A >> sh;
  const unsigned x1 = sh ? (A << (32 - sh)) | (B >> sh) : B;
  const unsigned x2 = sh ? B << (32 - sh) : 0u;
  if (x0) atomicOr(bs + w, x0);
  if (x1) atomicOr(bs + w + 1, x1);
  if (x2) atomicOr(bs + w + 2, x2);
}

constexpr int kBsWords = 56;  // bit-string image per block: 63 x 26 bits = 52 words at most, padded to whole uint4

__device__ __forceinline__ void block_code(const int16_t* t16, const uint4 q, int lane_r, const uint32_t* acb, uint32_t* ent,
                                           uint32_t* bs, uint32_t* gout_words, uint4* meta_out) {
  auto nz2 = [](unsigned w) { return ((w & 0xffffu) ? 1u : 0u) | ((w >> 16) ? 2u : 0u); };
  const unsigned m8 = nz2(q.x) | (nz2(q.y) << 2) | (nz2(q.z) << 4) | (nz2(q.w) << 6);  // positions 8r .. 8r+7
  unsigned lo = lane_r < 4 ? m8 << (8 * lane_r) : 0u, hi = lane_r >= 4 ? m8 << (8 * (lane_r - 4)) : 0u;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {  // OR over the 8 lanes of the block (aligned group: xor stays inside)
    lo |= __shfl_xor_sync(0xffffffffu, lo, o);
    hi |= __shfl_xor_sync(0xffffffffu, hi, o);
  }
  const unsigned long long mask = ((unsigned long long)hi << 32) | lo;
  const unsigned long long anchored = mask | 1ull;  // runs are counted from the DC position
  const unsigned long long mask_ac = mask & ~1ull;
  unsigned bits = 0;
  const unsigned zrl = acb[0xF0] & 0xff, zcode = acb[0xF0] >> 8;
  // step 1: code words.  entry = [24:0] Huffman code followed by the magnitude bits (bit 25 of a 26-bit
  // word is always 1: only the 16-bit codes, which all start with a one, can reach 26 bits), [29:25] its
  // length, [31:30] the number of ZRL codes in front of it
#pragma unroll
  for (int half = 0; half < 2; half++) {
    unsigned mj = ((half ? hi : lo) >> lane_r) & 0x01010101u;
    if (half == 0 && lane_r == 0) mj &= ~1u;  // the DC coefficient is not part of the AC code
    while (mj) {
      const int k = lane_r + (__ffs(mj) - 1) + 32 * half;
      mj &= mj - 1;
      const unsigned long long below = (1ull << k) - 1ull;
      const int prev = 63 - __clzll((long long)(anchored & below));
      const int run = k - prev - 1;
      const int v = t16[k];
      const int nb = 32 - __clz(abs(v));
      const uint32_t e = acb[((run & 15) << 4) | nb];
      const unsigned low = (unsigned)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
      const unsigned len = (e & 0xff) + nb;
      bits += len + (run >> 4) * zrl;
      ent[__popcll(mask_ac & below)] = ((((e >> 8) << nb) | low) & 0x1ffffffu) | (len << 25) | ((unsigned)(run >> 4) << 30);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
  const bool eob = !(hi >> 31);
  if (eob) bits += acb[0] & 0xff;
  // step 2: bit string.  Items = the entries plus, if needed, EOB.  Strings of up to 96 bits (nearly all
  // blocks of natural images) are assembled in three registers per lane and ORed over the 8 lanes; they
  // travel inside the meta word.  Longer strings are built in the shared-memory image and go to the slot.
  const int n = __popcll(mask_ac);
  const int items = n + (eob ? 1 : 0);
  int imax = items;  // rounds are warp-uniform (shuffles): the longest of the warp's 4 blocks decides
  imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, 8));
  imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, 16));
  const unsigned nw = (bits + 31) >> 5;
  const bool longb = bits > 96;
  if (longb)
    for (unsigned w = lane_r; w < nw; w += 8) bs[w] = 0;
  __syncwarp();
  const unsigned eob_ent = ((acb[0] >> 8) & 0x1ffffffu) | ((acb[0] & 0xff) << 25);
  unsigned base = 0, c0 = 0, c1 = 0, c2 = 0;
  for (int t = 0; t * 8 < imax; t++) {
    const int idx = t * 8 + lane_r;
    const unsigned e = idx < n ? ent[idx] : (idx == n && eob ? eob_ent : 0u);
    const unsigned zr = e >> 30, len = (e >> 25) & 31u;
    const unsigned tl = len + zr * zrl;
    unsigned incl = tl;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, incl, o, 8);
      if (lane_r >= o) incl += y;
    }
    const unsigned off = base + incl - tl;
    base += __shfl_sync(0xffffffffu, incl, 7, 8);
    if (tl) {
      unsigned long long pat = (e & 0x1ffffffu) | (len == 26 ? 1u << 25 : 0u);
      for (unsigned z = 0; z < zr; z++) pat |= (unsigned long long)zcode << (len + z * zrl);
      if (longb) {
        bs_place(bs, off, pat, tl);
      } else {
        const unsigned long long v = pat << (64 - tl);  // left aligned
        const unsigned A = (unsigned)(v >> 32), B = (unsigned)v;
        const unsigned w = off >> 5, sh = off & 31;
        const unsigned x0 = A >> sh;
        const unsigned x1 = sh ? (A << (32 - sh)) | (B >> sh) : B;
        const unsigned x2 = sh ? B << (32 - sh) : 0u;
        if (w == 0) { c0 |= x0; c1 |= x1; c2 |= x2; }
        else if (w == 1) { c1 |= x0; c2 |= x1; }
        else c2 |= x0;
      }
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    c0 |= __shfl_xor_sync(0xffffffffu, c0, o);
    c1 |= __shfl_xor_sync(0xffffffffu, c1, o);
    c2 |= __shfl_xor_sync(0xffffffffu, c2, o);
  }
  __syncwarp();
  if (longb) {
    c0 = bs[0]; c1 = bs[1]; c2 = bs[2];
    if (gout_words)  // 16 bytes per lane and round; the tail of the last vector is don't-care
      for (unsigned w4 = 4 * lane_r; w4 < nw; w4 += 32) *(uint4*)(gout_words + w4) = *(const uint4*)(bs + w4);
  }
  if (lane_r == 0 && meta_out) *meta_out = make_uint4((bits << 16) | ((unsigned)(int)t16[0] & 0xffffu), c0, c1, c2);
}

template <bool ZIGZAG>
__device__ __forceinline__ void block_stage(int d[8], int* tile, int lane_b, int lane_r, const unsigned* sdiv,
                                            const unsigned* smag, const uint8_t* sunzig, int16_t* gout_block_base,
                                            const uint32_t* acb, uint32_t* ent, uint32_t* bs, uint4* meta_out) {
  // pass 1 on this lane's row, park it
  dct1d<0>(d);
  int* t = tile + lane_b * kTileStride;
  *(int4*)(t + lane_r * 8) = make_int4(d[0], d[1], d[2], d[3]);
  *(int4*)(t + lane_r * 8 + 4) = make_int4(d[4], d[5], d[6], d[7]);
  __syncwarp();
  // pass 2 on column c = lane_r
  const int c = lane_r;
#pragma unroll
  for (int k = 0; k < 8; k++) d[k] = t[k * 8 + c];
  dct1d<1>(d);
  __syncwarp();
  // quantise; element k of this column is natural index k*8 + c
  int16_t* t16 = reinterpret_cast<int16_t*>(t);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int n = k * 8 + c;
    const unsigned dv = sdiv[n];
    const unsigned a = (unsigned)abs(d[k]) + (dv >> 1);
    int q = (int)__umulhi(a, smag[n]);
    q = d[k] < 0 ? -q : q;
    const int pos = ZIGZAG ? (int)sunzig[n] : n;
    t16[pos] = (int16_t)q;
  }
  __syncwarp();
  // 16 bytes per lane: coefficients [8*lane_r, 8*lane_r + 8) of block lane_b
  const uint4 q = *(const uint4*)(t16 + lane_r * 8);
  if (ZIGZAG) {
    // device entropy coder follows: the AC bit string instead of coefficients (all 32 lanes take part in
    // the shuffles; dead lanes store nothing).  The block's slot holds 64 words.
    block_code(t16, q, lane_r, acb, ent + lane_b * 64, bs + lane_b * kBsWords,
               gout_block_base ? reinterpret_cast<uint32_t*>(gout_block_base) : nullptr, meta_out);
  } else if (gout_block_base) {
    *(uint4*)(gout_block_base + lane_r * 8) = q;
  }
  __syncwarp();
}

template <bool ZIGZAG>
__global__ void __launch_bounds__(256, 4) k_fdct8(const Fdct8Params P) {
  __shared__ unsigned sdiv[2][64], smag[2][64];
  __shared__ int tiles[8][4 * kTileStride];
  __shared__ uint8_t sunzig[64];
  __shared__ uint32_t sacb[ZIGZAG ? 2 : 1][ZIGZAG ? 256 : 1];       // AC code books (code << 8 | length)
  __shared__ uint32_t sent[ZIGZAG ? 8 : 1][ZIGZAG ? 4 * 64 : 1];      // per warp: code-word entries of its 4 blocks
  __shared__ __align__(16) uint32_t sbs[ZIGZAG ? 8 : 1][ZIGZAG ? 4 * kBsWords : 4];  // per warp: their bit-string images
  if (ZIGZAG) {
    sacb[0][threadIdx.x] = __ldg(P.acbooks + threadIdx.x);
    sacb[1][threadIdx.x] = __ldg(P.acbooks + 256 + threadIdx.x);
  }
  if (threadIdx.x >= 128 && threadIdx.x < 192) sunzig[threadIdx.x - 128] = (uint8_t)unzig_rt(threadIdx.x - 128);
  if (threadIdx.x < 128) {
    const int t = threadIdx.x >> 6, i = threadIdx.x & 63;
    sdiv[t][i] = (unsigned)P.q[t][i] << 3;
    smag[t][i] = P.mag[t][i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lane_b = lane >> 3, lane_r = lane & 7;
  const int total_tiles = P.tile_end[P.nplanes - 1];
  // persistent CTAs: tiles of 32 horizontally adjacent blocks, all planes in one flat index space
#pragma unroll 1
  for (int tidx = blockIdx.x; tidx < total_tiles; tidx += gridDim.x) {
  const int pi = tidx < P.tile_end[0] ? 0 : (tidx < P.tile_end[1] ? 1 : 2);
  const Fdct8Plane& pl = P.plane[pi];
  const int local = tidx - (pi ? P.tile_end[pi - 1] : 0);
  const int tiles_x = (pl.wblocks + 31) >> 5;
  const int by = local / tiles_x, tx = local - by * tiles_x;
  const int bx = tx * 32 + warp * 4 + lane_b;
  const bool live = bx < pl.wblocks;
  const int bxc = live ? bx : pl.wblocks - 1;  // dead lanes compute on a valid block, store nothing
  int* tile = tiles[warp];
  int d[8];
  if (!pl.rgb) {
    int y = by * 8 + lane_r;
    if (y >= pl.h) {  // rows past the plane: the encoder helper's pad row (jpegencoderhelper.cpp:254-296)
#pragma unroll
      for (int k = 0; k < 8; k++) d[k] = pl.fill - 128;
    } else {
      const uint2 v = __ldg((const uint2*)(pl.src + (size_t)y * pl.stride + bxc * 8));
#pragma unroll
      for (int k = 0; k < 4; k++) {
        d[k] = (int)((v.x >> (8 * k)) & 0xff) - 128;
        d[4 + k] = (int)((v.y >> (8 * k)) & 0xff) - 128;
      }
    }
    const size_t bidx = (size_t)by * pl.wblocks + bx;
    int16_t* out = live ? pl.coefs[0] + bidx * (ZIGZAG ? 128 : 64) : nullptr;
    block_stage<ZIGZAG>(d, tile, lane_b, lane_r, sdiv[pl.tq[0]], smag[pl.tq[0]], sunzig, out, sacb[ZIGZAG ? pl.hsel[0] : 0],
                        sent[ZIGZAG ? warp : 0], sbs[ZIGZAG ? warp : 0], live && pl.meta[0] ? pl.meta[0] + bidx : nullptr);
  } else {
    // RGB888: libjpeg's scanline path replicates the last column / row (jcsample.c, jcprepct.c)
    const int y = min(by * 8 + lane_r, pl.h - 1);
    const uint8_t* row = pl.src + (size_t)y * pl.stride * 3;
    int r[8], g[8], b[8];
    if (bxc * 8 + 8 <= pl.w) {
      const uint2* p = (const uint2*)(row + (size_t)bxc * 24);
      const uint2 a = __ldg(p), bb = __ldg(p + 1), cc = __ldg(p + 2);
      const unsigned w[6] = {a.x, a.y, bb.x, bb.y, cc.x, cc.y};
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int o = 3 * k;
        r[k] = (w[o >> 2] >> (8 * (o & 3))) & 0xff;
        g[k] = (w[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 0xff;
        b[k] = (w[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 0xff;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint8_t* px = row + (size_t)min(bxc * 8 + k, pl.w - 1) * 3;
        r[k] = __ldg(px); g[k] = __ldg(px + 1); b[k] = __ldg(px + 2);
      }
    }
#pragma unroll
    for (int comp = 0; comp < 3; comp++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {  // jccolor.c rgb_ycc_convert, SCALEBITS 16
        int v;
        if (comp == 0) v = (19595 * r[k] + 38470 * g[k] + 7471 * b[k] + 32768) >> 16;
        else if (comp == 1) v = (-11059 * r[k] - 21709 * g[k] + 32768 * b[k] + (128 << 16) + 32767) >> 16;
        else v = (32768 * r[k] - 27439 * g[k] - 5329 * b[k] + (128 << 16) + 32767) >> 16;
        d[k] = v - 128;
      }
      const size_t bidx = (size_t)by * pl.wblocks + bx;
      int16_t* out = live ? pl.coefs[comp] + bidx * (ZIGZAG ? 128 : 64) : nullptr;
      block_stage<ZIGZAG>(d, tile, lane_b, lane_r, sdiv[pl.tq[comp]], smag[pl.tq[comp]], sunzig, out, sacb[ZIGZAG ? pl.hsel[comp] : 0],
                          sent[ZIGZAG ? warp : 0], sbs[ZIGZAG ? warp : 0], live && pl.meta[comp] ? pl.meta[comp] + bidx : nullptr);
    }
  }
  }
}

}  // namespace

void jpeg_std_codebook(int which, uint32_t out[256]);  // jpeg_host.cpp: (code << 8 | length) per symbol

// the two AC code books (luminance, chrominance), once per device
static int fdct_device_books(const uint32_t** out) {
  static std::mutex mu;
  static uint32_t* per_dev[64] = {nullptr};
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(E_ERROR, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  if (!per_dev[dev]) {
    uint32_t host[512];
    jpeg_std_codebook(1, host);
    jpeg_std_codebook(3, host + 256);
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof host));
    CUDA_TRY(cudaMemcpy(d, host, sizeof host, cudaMemcpyHostToDevice));
    per_dev[dev] = d;
  }
  *out = per_dev[dev];
  return E_OK;
}

cudaError_t launch_fdct8(const Fdct8Params& Pin, cudaStream_t s) {
  count_launches(1);
  Fdct8Params P = Pin;
  if (P.zigzag) {
    int rc = fdct_device_books(&P.acbooks);
    if (rc) return cudaErrorUnknown;
  }
  for (int t = 0; t < 2; t++)
    for (int i = 0; i < 64; i++) {
      const unsigned d = (unsigned)P.q[t][i] << 3;
      P.mag[t][i] = d ? (unsigned)((0x100000000ull + d - 1) / d) : 0u;
    }
  int total = 0;
  for (int i = 0; i < 3; i++) {
    if (i < P.nplanes) total += ((P.plane[i].wblocks + 31) / 32) * P.plane[i].hblocks;
    P.tile_end[i] = total;
  }
  if (total == 0) return cudaSuccess;
  static int resident_tab[2] = {0, 0};  // CTAs of one wave, per instantiation
  int& resident = resident_tab[P.zigzag ? 1 : 0];
  if (!resident) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const cudaError_t oe = P.zigzag ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fdct8<true>, 256, 0)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fdct8<false>, 256, 0);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    resident = per_sm * (sms > 0 ? sms : 148);
  }
  const int ctas = total < resident ? total : resident;
  if (P.zigzag) k_fdct8<true><<<ctas, 256, 0, s>>>(P);
  else k_fdct8<false><<<ctas, 256, 0, s>>>(P);
  return cudaGetLastError();
}

}  // namespace uhdr_b200
