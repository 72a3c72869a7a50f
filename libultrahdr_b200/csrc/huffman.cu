// Baseline-JPEG entropy coding on the device, bit-identical to libjpeg-turbo's jchuff.c for the
// reference's settings (default Annex-K tables, one interleaved scan, no restart markers).
//
//   E1  k_huff_blocks  : one thread per 8x8 block (in scan order).  DC prediction reads the
//                        previous block of the same component; the block's code bits go to a
//                        word-transposed scratch (word w of block s at scratch[w*nblocks+s], so a
//                        warp's stores are contiguous) and its bit count to bits[s].
//   E2  scan           : exclusive prefix sum of bits[] -> bit offset of every block.
//   E3  k_huff_concat  : blocks OR their words into the MSB-first bitstream at their offsets.
//   E4  k_ff_count / scan / k_stuff : pad the last byte with ones, insert 0x00 after every 0xFF.
// Only the final stuffed segment (a few MB at 4K) crosses PCIe.
#include <cstring>

#include "jpeg.h"

namespace uhdr_b200 {

void jpeg_std_codebook(int which, uint32_t out[256]);

namespace {

constexpr int kMaxWordsPerBlock = 52;  // (20 + 63*26 bits) / 32 rounded up
constexpr int kPersistentCtas = 148 * 2;  // one wave of 1024-thread CTAs on the 148 SMs

struct HuffFrame {
  const int16_t* coefs[3];
  int wblocks[3], mw[3], mh[3], koff[3];
  int ncomp, mcus_per_row, blocks_per_mcu;
  unsigned nblocks;
};

__device__ __forceinline__ void locate(const HuffFrame& f, unsigned s, int& c, unsigned& blk, long long& prev) {
  const unsigned m = s / f.blocks_per_mcu, k = s - m * f.blocks_per_mcu;
  c = (f.ncomp > 2 && k >= (unsigned)f.koff[2]) ? 2 : ((f.ncomp > 1 && k >= (unsigned)f.koff[1]) ? 1 : 0);
  const unsigned kk = k - f.koff[c];
  const unsigned mx = m % f.mcus_per_row, my = m / f.mcus_per_row;
  const unsigned mw = f.mw[c], mh = f.mh[c];
  blk = (my * mh + kk / mw) * f.wblocks[c] + mx * mw + kk % mw;
  if (kk > 0) {
    const unsigned pk = kk - 1;
    prev = (long long)(my * mh + pk / mw) * f.wblocks[c] + mx * mw + pk % mw;
  } else if (m > 0) {
    const unsigned pm = m - 1, pk = mw * mh - 1;
    const unsigned pmx = pm % f.mcus_per_row, pmy = pm / f.mcus_per_row;
    prev = (long long)(pmy * mh + pk / mw) * f.wblocks[c] + pmx * mw + pk % mw;
  } else {
    prev = -1;
  }
}

struct BitSink {
  unsigned* scratch;
  unsigned nblocks, s;
  unsigned long long acc;
  int fill, words;
  unsigned total;
  __device__ __forceinline__ void put(unsigned bits, int n) {
    acc = (acc << n) | bits;
    fill += n;
    total += n;
    if (fill >= 32) {
      scratch[(size_t)words * nblocks + s] = (unsigned)(acc >> (fill - 32));
      words++;
      fill -= 32;
    }
  }
  __device__ __forceinline__ void finish() {
    if (fill > 0) scratch[(size_t)words * nblocks + s] = (unsigned)(acc << (32 - fill));
  }
};

// coefficients are stored in zigzag order by k_fdct_quant (zigzag_out = 1)
__global__ void __launch_bounds__(128) k_huff_blocks(const HuffFrame f, const uint32_t* __restrict__ books,
                                                     unsigned* __restrict__ scratch, unsigned* __restrict__ bits) {
  __shared__ uint32_t sb[4 * 256];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sb[i] = books[i];
  __syncthreads();
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= f.nblocks) return;
  int c;
  unsigned blk;
  long long prev;
  locate(f, s, c, blk, prev);
  const int16_t* __restrict__ base = f.coefs[c];
  const uint4* src = (const uint4*)(base + (size_t)blk * 64);
  unsigned w[32];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint4 q = __ldg(src + i);
    w[i * 4] = q.x; w[i * 4 + 1] = q.y; w[i * 4 + 2] = q.z; w[i * 4 + 3] = q.w;
  }
  const int pred = prev >= 0 ? (int)__ldg(base + (size_t)prev * 64) : 0;
  const uint32_t* dcb = sb + (c == 0 ? 0 : 512);
  const uint32_t* acb = sb + (c == 0 ? 256 : 768);
  BitSink o{scratch, f.nblocks, s, 0ull, 0, 0, 0u};
  {
    const int dc = (int)(short)(w[0] & 0xffff);
    const int diff = dc - pred;
    const int mag = abs(diff);
    const int nb = mag ? 32 - __clz(mag) : 0;
    const uint32_t e = dcb[nb];
    const unsigned low = (unsigned)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u);
    o.put(((e >> 8) << nb) | low, (int)(e & 0xff) + nb);
  }
  int run = 0;
#pragma unroll
  for (int k = 1; k < 64; k++) {
    const int v = (int)(short)((w[k >> 1] >> ((k & 1) * 16)) & 0xffff);
    if (v == 0) {
      run++;
    } else {
      while (run > 15) {
        const uint32_t z = acb[0xF0];
        o.put(z >> 8, (int)(z & 0xff));
        run -= 16;
      }
      const int mag = abs(v);
      const int nb = 32 - __clz(mag);
      const uint32_t e = acb[(run << 4) | nb];
      const unsigned low = (unsigned)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
      o.put(((e >> 8) << nb) | low, (int)(e & 0xff) + nb);
      run = 0;
    }
  }
  if (run > 0) {
    const uint32_t e = acb[0];
    o.put(e >> 8, (int)(e & 0xff));
  }
  o.finish();
  bits[s] = o.total;
}

// ---- exclusive scan of u32 (three-phase, tiles of 1024) -------------------------------------------
constexpr int kScanTile = 1024;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {
  __shared__ unsigned warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    unsigned ws = lane < nw ? warp_sums[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += y;
    }
    warp_sums[lane] = ws;  // inclusive
  }
  __syncthreads();
  const unsigned base = wid ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
  const unsigned r = base + x - v;
  __syncthreads();
  return r;
}

// n may come from device memory (n_ptr) so dependent stages need no host round trip
__global__ void __launch_bounds__(kScanTile) k_scan_reduce(const unsigned* __restrict__ in, unsigned n,
                                                           const unsigned* n_ptr, unsigned* __restrict__ partial) {
  if (n_ptr) n = *n_ptr;
  const unsigned i = blockIdx.x * kScanTile + threadIdx.x;
  if (blockIdx.x * kScanTile >= n && blockIdx.x > 0) return;
  unsigned t;
  block_exclusive_scan(i < n ? in[i] : 0u, &t);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void __launch_bounds__(kScanTile) k_scan_partials(unsigned* __restrict__ partial, unsigned ntiles_max,
                                                             unsigned n, const unsigned* n_ptr, unsigned* total_out) {
  if (n_ptr) n = *n_ptr;
  const unsigned ntiles = min(ntiles_max, (n + kScanTile - 1) / kScanTile);
  unsigned carry = 0;
  for (unsigned base = 0; base < ntiles; base += kScanTile) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < ntiles ? partial[i] : 0u;
    unsigned t;
    const unsigned e = block_exclusive_scan(v, &t);
    if (i < ntiles) partial[i] = carry + e;
    carry += t;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(kScanTile) k_scan_apply(const unsigned* __restrict__ in, unsigned n, const unsigned* n_ptr,
                                                          const unsigned* __restrict__ partial, unsigned* __restrict__ out) {
  if (n_ptr) n = *n_ptr;
  if (blockIdx.x * kScanTile >= n) return;
  const unsigned i = blockIdx.x * kScanTile + threadIdx.x;
  const unsigned v = i < n ? in[i] : 0u;
  const unsigned e = block_exclusive_scan(v, nullptr);
  if (i < n) out[i] = partial[blockIdx.x] + e;
}

// ---- E3: concatenate -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_huff_concat(const unsigned* __restrict__ scratch, const unsigned* __restrict__ bits,
                                                     const unsigned* __restrict__ offs, unsigned nblocks,
                                                     unsigned* __restrict__ stream, unsigned cap_words, unsigned* overflow) {
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nblocks) return;
  const unsigned tb = bits[s];
  const unsigned off = offs[s];
  const unsigned nw = (tb + 31) >> 5;
  if (((unsigned long long)off + tb + 31) / 32 + 1 > cap_words) {
    *overflow = 1;
    return;
  }
  const unsigned sh = off & 31;
  unsigned idx = off >> 5;
  for (unsigned w = 0; w < nw; w++, idx++) {
    const unsigned v = scratch[(size_t)w * nblocks + s];  // unused low bits of the last word are zero
    if (sh == 0) {
      atomicOr(stream + idx, v);
    } else {
      atomicOr(stream + idx, v >> sh);
      const unsigned lo = v << (32 - sh);
      if (lo) atomicOr(stream + idx + 1, lo);
    }
  }
}

// ---- E4: byte stuffing ---------------------------------------------------------------------------
// logical word j of `stream` holds stream bytes 4j..4j+3, first byte in the most significant bits
__device__ __forceinline__ unsigned load_padded_word(const unsigned* stream, unsigned j, unsigned total_bits) {
  unsigned v = stream[j];
  const unsigned total_bytes = (total_bits + 7) >> 3;
  const unsigned pad = total_bytes * 8 - total_bits;  // jchuff.c flush_bits: fill with ones
  if (pad && j == (total_bytes - 1) >> 2) {
    const unsigned byte_in_word = (total_bytes - 1) & 3;
    v |= ((1u << pad) - 1u) << (24 - 8 * byte_in_word);
  }
  return v;
}
__device__ __forceinline__ unsigned ff_count(unsigned v, unsigned j, unsigned total_bytes) {
  unsigned c = 0;
#pragma unroll
  for (int b = 0; b < 4; b++)
    if (4 * j + b < total_bytes && ((v >> (24 - 8 * b)) & 0xff) == 0xff) c++;
  return c;
}
// zero exactly the words the concatenation will OR into (total known on the device only)
__global__ void __launch_bounds__(256) k_zero_stream(unsigned* __restrict__ stream, const unsigned* total_bits_ptr,
                                                     unsigned cap_words) {
  unsigned n = (*total_bits_ptr + 31) / 32 + 2;
  if (n > cap_words) n = cap_words;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) stream[i] = 0u;
}
// persistent grid: CTAs stride over the tiles actually in use
__global__ void __launch_bounds__(kScanTile) k_ff_count(const unsigned* __restrict__ stream, const unsigned* total_bits_ptr,
                                                        unsigned* __restrict__ tile_ff, unsigned* __restrict__ nwords_out) {
  const unsigned total_bits = *total_bits_ptr;
  const unsigned total_bytes = (total_bits + 7) >> 3;
  const unsigned nwords = (total_bytes + 3) >> 2;
  if (blockIdx.x == 0 && threadIdx.x == 0) *nwords_out = nwords;
  for (unsigned tile = blockIdx.x; tile * kScanTile < nwords; tile += gridDim.x) {
    const unsigned j = tile * kScanTile + threadIdx.x;
    unsigned c = 0;
    if (j < nwords) c = ff_count(load_padded_word(stream, j, total_bits), j, total_bytes);
    unsigned t;
    block_exclusive_scan(c, &t);
    if (threadIdx.x == 0) tile_ff[tile] = t;
  }
}
__global__ void __launch_bounds__(kScanTile) k_stuff(const unsigned* __restrict__ stream, const unsigned* total_bits_ptr,
                                                     const unsigned* __restrict__ tile_off, const unsigned* total_ff,
                                                     uint8_t* __restrict__ out, unsigned out_cap, unsigned* out_bytes,
                                                     unsigned* overflow) {
  const unsigned total_bits = *total_bits_ptr;
  const unsigned total_bytes = (total_bits + 7) >> 3;
  const unsigned nwords = (total_bytes + 3) >> 2;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *out_bytes = total_bytes + *total_ff;
    if (total_bytes + *total_ff > out_cap) *overflow = 1;
  }
  if (total_bytes + *total_ff > out_cap) return;
  for (unsigned tile = blockIdx.x; tile * kScanTile < nwords; tile += gridDim.x) {
    const unsigned j = tile * kScanTile + threadIdx.x;
    unsigned v = 0, c = 0;
    if (j < nwords) {
      v = load_padded_word(stream, j, total_bits);
      c = ff_count(v, j, total_bytes);
    }
    const unsigned e = block_exclusive_scan(c, nullptr);
    if (j < nwords) {
      unsigned pos = 4 * j + tile_off[tile] + e;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        if (4 * j + b >= total_bytes) break;
        const uint8_t byte = (uint8_t)((v >> (24 - 8 * b)) & 0xff);
        out[pos++] = byte;
        if (byte == 0xff) out[pos++] = 0;
      }
    }
  }
}

struct DeviceBooks {
  uint32_t* d = nullptr;
};
static int device_books(const uint32_t** out) {
  static thread_local int cached_dev = -1;
  static thread_local uint32_t* cached = nullptr;
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  if (cached && cached_dev == dev) { *out = cached; return E_OK; }
  uint32_t host[1024];
  for (int t = 0; t < 4; t++) jpeg_std_codebook(t, host + 256 * t);
  uint32_t* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof host));
  CUDA_TRY(cudaMemcpy(d, host, sizeof host, cudaMemcpyHostToDevice));
  cached = d;
  cached_dev = dev;
  *out = d;
  return E_OK;
}

}  // namespace

bool gpu_entropy_available() { return true; }

int jpeg_entropy_dev(Workspace& ws, JpegEncodeJob* job) {
  const JpegFrame& fr = job->frame;
  if (fr.has_dummy_blocks()) return fail(E_UNSUPPORTED, "device entropy coder needs whole MCUs");
  const uint32_t* books = nullptr;
  int rc = device_books(&books);
  if (rc) return rc;
  HuffFrame f;
  memset(&f, 0, sizeof f);
  f.ncomp = fr.ncomp;
  f.mcus_per_row = fr.mcus_per_row;
  int k = 0;
  for (int c = 0; c < fr.ncomp; c++) {
    f.coefs[c] = job->d_coefs[c];
    f.wblocks[c] = fr.comp[c].wblocks;
    f.mw[c] = fr.ncomp == 1 ? 1 : fr.comp[c].h_samp;
    f.mh[c] = fr.ncomp == 1 ? 1 : fr.comp[c].v_samp;
    f.koff[c] = k;
    k += f.mw[c] * f.mh[c];
  }
  f.blocks_per_mcu = k;
  const size_t nblocks = fr.total_blocks();
  f.nblocks = (unsigned)nblocks;
  // capacity of the entropy-coded segment: the reference's whole output buffer is w*h*6 bytes
  // (ultrahdr_api.cpp:1294); a single scan can never need more than that in a valid encode
  const size_t cap = ((size_t)fr.width * fr.height * 6 + 4096 + 3) / 4 * 4;
  const unsigned cap_words = (unsigned)(cap / 4);
  const unsigned ntiles_blocks = (unsigned)((nblocks + kScanTile - 1) / kScanTile);
  const unsigned ntiles_words = (cap_words + kScanTile - 1) / kScanTile;
  unsigned* scratch = (unsigned*)ws.dalloc(nblocks * kMaxWordsPerBlock * sizeof(unsigned));
  unsigned* bits = (unsigned*)ws.dalloc(nblocks * 4);
  unsigned* offs = (unsigned*)ws.dalloc(nblocks * 4);
  unsigned* part = (unsigned*)ws.dalloc((size_t)(ntiles_blocks > ntiles_words ? ntiles_blocks : ntiles_words) * 4 + 64);
  unsigned* stream = (unsigned*)ws.dalloc(cap + 64);
  unsigned* tile_ff = (unsigned*)ws.dalloc((size_t)ntiles_words * 4 + 64);
  unsigned* ctl = (unsigned*)ws.dalloc(64);  // [0] total_bits [1] tiles_in_use [2] total_ff [3] out_bytes [4] overflow
  job->d_scan = (uint8_t*)ws.dalloc(cap + 64);
  job->h_scan_bytes = (unsigned*)ws.halloc(64);
  if (!scratch || !bits || !offs || !part || !stream || !tile_ff || !ctl || !job->d_scan || !job->h_scan_bytes) return E_MEM;
  job->d_scan_bytes = ctl + 3;
  job->scan_capacity = cap;
  cudaStream_t st = ws.stream();
  CUDA_TRY(cudaMemsetAsync(ctl, 0, 64, st));

  ws.t_begin("huff_blocks");
  k_huff_blocks<<<(unsigned)((nblocks + 127) / 128), 128, 0, st>>>(f, books, scratch, bits);
  ws.t_end();
  ws.t_begin("huff_scan");
  k_scan_reduce<<<ntiles_blocks, kScanTile, 0, st>>>(bits, (unsigned)nblocks, nullptr, part);
  k_scan_partials<<<1, kScanTile, 0, st>>>(part, ntiles_blocks, (unsigned)nblocks, nullptr, ctl + 0);
  k_scan_apply<<<ntiles_blocks, kScanTile, 0, st>>>(bits, (unsigned)nblocks, nullptr, part, offs);
  ws.t_end();
  ws.t_begin("huff_concat");
  k_zero_stream<<<kPersistentCtas, 256, 0, st>>>(stream, ctl + 0, cap_words);
  k_huff_concat<<<(unsigned)((nblocks + 255) / 256), 256, 0, st>>>(scratch, bits, offs, (unsigned)nblocks, stream, cap_words, ctl + 4);
  ws.t_end();
  ws.t_begin("huff_stuff");
  k_ff_count<<<kPersistentCtas, kScanTile, 0, st>>>(stream, ctl + 0, tile_ff, ctl + 1);
  k_scan_partials<<<1, kScanTile, 0, st>>>(tile_ff, ntiles_words, 0, ctl + 1, ctl + 2);
  k_stuff<<<kPersistentCtas, kScanTile, 0, st>>>(stream, ctl + 0, tile_ff, ctl + 2, job->d_scan, (unsigned)cap, ctl + 3, ctl + 4);
  ws.t_end();
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(job->h_scan_bytes, ctl, 32, cudaMemcpyDeviceToHost, st));
  return E_OK;
}

// second phase after the sizes are on the host: fetch exactly the bytes produced
int jpeg_entropy_fetch(Workspace& ws, JpegEncodeJob* job) {
  const unsigned* ctl = job->h_scan_bytes;
  if (ctl[4]) return fail(E_MEM, "entropy-coded segment exceeds the %zu byte device buffer", job->scan_capacity);
  const unsigned n = ctl[3];
  job->h_scan = (uint8_t*)ws.halloc(n + 64);
  if (!job->h_scan) return E_MEM;
  CUDA_TRY(cudaMemcpyAsync(job->h_scan, job->d_scan, n, cudaMemcpyDeviceToHost, ws.stream()));
  return E_OK;
}

}  // namespace uhdr_b200
