// Baseline-JPEG entropy coding on the device, bit-identical to libjpeg-turbo's jchuff.c for the
// reference's settings (default Annex-K tables, one interleaved scan, no restart markers).
//
//   E1  k_huff_encode  : CTAs of 128 consecutive blocks (scan order): stage the coefficients in
//                        shared memory, count each block's code bits, chain the bit offsets across
//                        CTAs with a decoupled look-back, encode into a word-aligned shared-memory
//                        image of the CTA's segment and store it; partial boundary words travel
//                        from CTA to CTA (details at the kernel).
//   E2  k_stuff_lb     : pad the last byte with ones, insert 0x00 after every 0xFF (count, look-back
//                        over 1024-word tiles and write in one launch).
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

// ---- E1: encode, chain the bit offsets across CTAs, write the stream ------------------------------
// One CTA = kEncBlocks consecutive blocks of the scan, taken in ticket order so that a CTA's
// predecessors are always already running (the cross-CTA steps below spin on them).
//   a. the coefficient blocks are staged in shared memory with coalesced loads (stride-129 word
//      tile: thread j then reads word r of its block at tile[r*129 + j] without bank conflicts)
//   b. pass A: each thread builds the 64-bit non-zero mask of its block and counts its code bits
//      (only the non-zero coefficients are visited)
//   c. CTA-wide exclusive scan -> bit offset of each block inside the CTA's segment and the total
//   d. decoupled look-back over per-CTA status words (aggregate / inclusive prefix) -> the
//      segment's absolute bit offset
//   e. pass B: the blocks are encoded again, now ORing their bits into a shared-memory image of the
//      segment that is aligned to the 32-bit words of the output stream
//   f. interior words are stored coalesced; the trailing partial word is handed to the successor
//      CTA, which merges it with its own leading bits and stores the word.  No pre-zeroed stream,
//      no scratch, no global atomics on the stream.
constexpr int kEncBlocks = 128;
constexpr int kTileStride = kEncBlocks + 1;
// shared-memory image of the CTA's segment: 1024 words = 256 bits per block on average.  Heavier
// segments (noise at high quality; the worst case is 52 words per block) are produced in several
// windows of this size, pass B running once per window.
constexpr unsigned kSegWords = 1024;
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagMask = 3ull << 62;

// Decoupled look-back (one warp): sum of the values of all predecessors of `idx`.  status[i] carries a
// 2-bit flag and a 32-bit value: kFlagAgg = the value of i alone, kFlagPrefix = the inclusive prefix up
// to i.  Predecessors are running or finished (ticket order), so the spin loops terminate.
__device__ __forceinline__ unsigned lookback_sum(const unsigned long long* status, unsigned idx, int lane) {
  unsigned base = 0;
  long long hi = (long long)idx - 1;  // highest predecessor not yet accounted for
  for (;;) {
    const long long k = hi - lane;
    unsigned long long st = kFlagPrefix;  // before element 0: nothing
    if (k >= 0) {
      do { st = *reinterpret_cast<const volatile unsigned long long*>(status + k); } while ((st & kFlagMask) == 0);
    }
    const bool is_prefix = (st & kFlagMask) == kFlagPrefix;
    const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
    const int first_prefix = pm ? __ffs(pm) - 1 : 32;  // nearest predecessor carrying an inclusive prefix
    unsigned v = (lane <= first_prefix && k >= 0) ? (unsigned)(st & 0xffffffffu) : 0u;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    base += v;
    if (pm) return base;
    hi -= 32;
  }
}

struct EncSmem {
  uint32_t tile[32 * kTileStride];
  uint32_t seg[kSegWords];
};

template <bool EMIT>
__device__ __forceinline__ unsigned encode_block(const uint32_t* __restrict__ tile, int j, unsigned long long mask, int dc_diff,
                                                 const uint32_t* __restrict__ dcb, const uint32_t* __restrict__ acb, uint32_t* seg,
                                                 unsigned bitpos, unsigned win) {
  unsigned total = 0;
  unsigned long long acc = 0;
  int fill = (int)(bitpos & 31);
  unsigned widx = bitpos >> 5;
  auto put = [&](unsigned bits, int n) {
    total += n;
    if (EMIT) {
      acc = (acc << n) | bits;
      fill += n;
      if (fill >= 32) {
        if (widx - win < kSegWords) atomicOr(seg + (widx - win), (unsigned)(acc >> (fill - 32)));
        widx++;
        fill -= 32;
      }
    }
  };
  {
    const int mag = abs(dc_diff);
    const int nb = mag ? 32 - __clz(mag) : 0;
    const uint32_t e = __ldg(dcb + nb);
    const unsigned low = (unsigned)(dc_diff < 0 ? dc_diff - 1 : dc_diff) & ((1u << nb) - 1u);
    put(((e >> 8) << nb) | low, (int)(e & 0xff) + nb);
  }
  int last = 0;
  const uint32_t zrl = __ldg(acb + 0xF0);
  while (mask) {
    const int k = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    int run = k - last - 1;
    last = k;
    while (run > 15) {
      put(zrl >> 8, (int)(zrl & 0xff));
      run -= 16;
    }
    const uint32_t w = tile[(k >> 1) * kTileStride + j];
    const int v = (int)(short)((w >> ((k & 1) * 16)) & 0xffff);
    const int mag = abs(v);
    const int nb = 32 - __clz(mag);
    const uint32_t e = __ldg(acb + ((run << 4) | nb));
    const unsigned low = (unsigned)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
    put(((e >> 8) << nb) | low, (int)(e & 0xff) + nb);
  }
  if (last != 63) {
    const uint32_t e = __ldg(acb);
    put(e >> 8, (int)(e & 0xff));
  }
  if (EMIT && (fill & 31) && widx - win < kSegWords) atomicOr(seg + (widx - win), (unsigned)(acc << (32 - fill)));
  return total;
}

// status[i]: look-back word of CTA i (flag | bits);  tails[i]: (1 << 32 | trailing partial word) once known
// ctl[0] <- total bits, ctl[5] = ticket counter (zeroed by the caller together with status / tails)
__global__ void __launch_bounds__(kEncBlocks) k_huff_encode(const HuffFrame f, const uint32_t* __restrict__ books,
                                                            unsigned long long* status, unsigned long long* tails,
                                                            unsigned* __restrict__ stream, unsigned cap_words, unsigned* ctl) {
  extern __shared__ uint32_t enc_smem_raw[];
  EncSmem& sm = *reinterpret_cast<EncSmem*>(enc_smem_raw);
  __shared__ unsigned s_cta, s_base, s_total;
  __shared__ const int16_t* s_src[kEncBlocks];
  const int j = threadIdx.x;
  if (j == 0) s_cta = atomicAdd(ctl + 5, 1u);
  __syncthreads();
  const unsigned cta = s_cta;
  const unsigned s = cta * kEncBlocks + j;
  const bool live = s < f.nblocks;
  int c = 0;
  int pred = 0;
  if (live) {
    unsigned blk;
    long long prev;
    locate(f, s, c, blk, prev);
    s_src[j] = f.coefs[c] + (size_t)blk * 64;
    pred = prev >= 0 ? (int)__ldg(f.coefs[c] + (size_t)prev * 64) : 0;
  } else {
    s_src[j] = nullptr;
  }
  __syncthreads();
  // a. 8 lanes per block, 16 bytes each: 4 blocks = 512 contiguous bytes per warp instruction when the
  //    blocks are neighbours in their plane
  for (int i = j; i < kEncBlocks * 8; i += kEncBlocks) {
    const int b = i >> 3, ch = i & 7;
    const int16_t* src = s_src[b];
    if (src) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + ch);
      uint32_t* t = sm.tile + (ch * 4) * kTileStride + b;
      t[0] = q.x; t[kTileStride] = q.y; t[2 * kTileStride] = q.z; t[3 * kTileStride] = q.w;
    }
  }
  __syncthreads();
  // b. non-zero mask and bit count
  unsigned long long mask = 0;
  int dc_diff = 0;
  unsigned nbits = 0;
  // code books (4 KB, hot in L1) are read through the read-only path: keeping them out of shared
  // memory buys two more resident CTAs per SM
  const uint32_t* dcb = books + (c == 0 ? 0 : 512);
  const uint32_t* acb = books + (c == 0 ? 256 : 768);
  if (live) {
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const uint32_t w = sm.tile[r * kTileStride + j];
      if (w & 0xffffu) mask |= 1ull << (2 * r);
      if (w >> 16) mask |= 2ull << (2 * r);
    }
    dc_diff = (int)(short)(sm.tile[j] & 0xffff) - pred;
    mask &= ~1ull;
    nbits = encode_block<false>(sm.tile, j, mask, dc_diff, dcb, acb, nullptr, 0, 0);
  }
  // c. offsets inside the CTA
  unsigned total;
  const unsigned off = block_exclusive_scan(nbits, &total);
  // d. absolute offset: publish the aggregate, then look back (warp 0, 32 predecessors at a time)
  if (j == 0) {
    *reinterpret_cast<volatile unsigned long long*>(status + cta) = (cta == 0 ? kFlagPrefix : kFlagAgg) | total;
    s_total = total;
  }
  if (j < 32 && cta > 0) {
    const unsigned base = lookback_sum(status, cta, j);
    if (j == 0) {
      s_base = base;
      *reinterpret_cast<volatile unsigned long long*>(status + cta) = kFlagPrefix | (unsigned long long)(base + total);
    }
  } else if (j == 0 && cta == 0) {
    s_base = 0;
  }
  __syncthreads();
  const unsigned base = s_base;
  const unsigned sh = base & 31, first = base >> 5;
  const unsigned endbit = sh + total;              // relative to word `first`
  const unsigned nwords = (endbit + 31) >> 5;      // words of the segment image in use
  const bool overflow = (unsigned long long)first + nwords + 1 > cap_words;
  // e./f. segment image, one window of kSegWords words at a time; full words are stored coalesced,
  //       the head word (index 0) and the trailing partial word (index lastw) are kept for step g
  const unsigned tailbits = endbit & 31;
  const unsigned lastw = endbit >> 5;              // index (relative) of the word holding the trailing partial bits
  const bool is_last = (cta + 1) * (unsigned)kEncBlocks >= f.nblocks;
  __shared__ unsigned s_head, s_tail;
  for (unsigned win = 0; win < nwords; win += kSegWords) {
    const unsigned wn = min(kSegWords, nwords - win);
    for (unsigned i = j; i < wn; i += kEncBlocks) sm.seg[i] = 0;
    __syncthreads();
    if (live) encode_block<true>(sm.tile, j, mask, dc_diff, dcb, acb, sm.seg, sh + off, win);
    __syncthreads();
    if (!overflow)
      for (unsigned i = j; i < wn; i += kEncBlocks) {
        const unsigned w = win + i;
        if (w >= 1 && w < lastw) stream[first + w] = sm.seg[i];
      }
    if (j == 0) {
      if (win == 0) s_head = sm.seg[0];
      if (lastw >= win && lastw < win + wn) s_tail = sm.seg[lastw - win];
    }
    __syncthreads();
  }
  // g. boundary words
  if (j == 0) {
    // a non-degenerate segment hands its trailing partial word on *before* waiting for the
    // predecessor's: otherwise every CTA would wait for the whole chain in front of it
    const unsigned tail = (lastw > 0 && tailbits) ? s_tail : 0u;
    if (lastw > 0) {
      *reinterpret_cast<volatile unsigned long long*>(tails + cta) = (1ull << 32) | tail;
      if (is_last && tailbits && !overflow) stream[first + lastw] = tail;
    }
    unsigned head = s_head;
    if (sh > 0) {  // leading bits of word `first` belong to the predecessor
      unsigned long long t;
      do { t = *reinterpret_cast<volatile unsigned long long*>(tails + cta - 1); } while ((t >> 32) == 0);
      head |= (unsigned)t;
    }
    if (lastw > 0) {
      if (!overflow) stream[first] = head;
    } else {
      // the whole segment lies inside word `first` (only a short last CTA can be this small): the
      // merged word is also our trailing partial word
      if (is_last && !overflow) stream[first] = head;
      *reinterpret_cast<volatile unsigned long long*>(tails + cta) = (1ull << 32) | head;
    }
    if (overflow) ctl[4] = 1;
    if (is_last) ctl[0] = base + total;
  }
}

// ---- E2: byte stuffing ---------------------------------------------------------------------------
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
// count + look-back + write in one launch: persistent CTAs take 1024-word tiles in ticket order
// ctl[0] total bits (in), ctl[3] <- stuffed bytes, ctl[4] <- overflow, ctl[6] = tile tickets (zeroed by the caller)
__global__ void __launch_bounds__(kScanTile) k_stuff_lb(const unsigned* __restrict__ stream, unsigned* ctl, unsigned long long* status,
                                                        uint8_t* __restrict__ out, unsigned out_cap) {
  __shared__ unsigned s_tile, s_base;
  const unsigned total_bits = ctl[0];
  const unsigned total_bytes = (total_bits + 7) >> 3;
  const unsigned nwords = (total_bytes + 3) >> 2;
  const unsigned ntiles = (nwords + kScanTile - 1) / kScanTile;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(ctl + 6, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= ntiles) break;
    const unsigned j = tile * kScanTile + threadIdx.x;
    unsigned v = 0, c = 0;
    if (j < nwords) {
      v = load_padded_word(stream, j, total_bits);
      c = ff_count(v, j, total_bytes);
    }
    unsigned t;
    const unsigned e = block_exclusive_scan(c, &t);
    if (threadIdx.x == 0) {
      *reinterpret_cast<volatile unsigned long long*>(status + tile) = (tile == 0 ? kFlagPrefix : kFlagAgg) | t;
      if (tile == 0) s_base = 0;
    }
    if (threadIdx.x < 32 && tile > 0) {
      const unsigned b = lookback_sum(status, tile, (int)threadIdx.x);
      if (threadIdx.x == 0) {
        s_base = b;
        *reinterpret_cast<volatile unsigned long long*>(status + tile) = kFlagPrefix | (unsigned long long)(b + t);
      }
    }
    __syncthreads();
    const unsigned base = s_base;
    if (tile == ntiles - 1 && threadIdx.x == 0) {
      ctl[3] = total_bytes + base + t;
      if (total_bytes + base + t > out_cap) ctl[4] = 1;
    }
    if (j < nwords) {
      unsigned pos = 4 * j + base + e;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        if (4 * j + b >= total_bytes) break;
        const uint8_t byte = (uint8_t)((v >> (24 - 8 * b)) & 0xff);
        if (pos < out_cap) out[pos] = byte;
        pos++;
        if (byte == 0xff) {
          if (pos < out_cap) out[pos] = 0;
          pos++;
        }
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
  const unsigned ntiles_words = (cap_words + kScanTile - 1) / kScanTile;
  const unsigned ncta = (unsigned)((nblocks + kEncBlocks - 1) / kEncBlocks);
  // [ctl 64 B][status ncta x 8][tails ncta x 8], zeroed together
  const size_t ctl_bytes = 64 + (size_t)ncta * 16 + (size_t)ntiles_words * 8;
  unsigned* ctl = (unsigned*)ws.dalloc(ctl_bytes);  // [0] total_bits [3] out_bytes [4] overflow [5] encoder CTA tickets [6] stuffing tile tickets
  unsigned* stream = (unsigned*)ws.dalloc(cap + 64);
  job->d_scan = (uint8_t*)ws.dalloc(cap + 64);
  job->h_scan_bytes = (unsigned*)ws.halloc(64);
  if (!stream || !ctl || !job->d_scan || !job->h_scan_bytes) return E_MEM;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(ctl + 16);
  unsigned long long* tails = status + ncta;
  unsigned long long* stuff_status = tails + ncta;
  job->d_scan_bytes = ctl + 3;
  job->scan_capacity = cap;
  cudaStream_t st = ws.stream();
  CUDA_TRY(cudaMemsetAsync(ctl, 0, ctl_bytes, st));
  static bool smem_set[64] = {false};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !smem_set[dev]) {
      CUDA_TRY(cudaFuncSetAttribute(k_huff_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem)));
      smem_set[dev] = true;
    }
  }

  count_launches(2);
  ws.t_begin("huff_encode");
  k_huff_encode<<<ncta, kEncBlocks, sizeof(EncSmem), st>>>(f, books, status, tails, stream, cap_words, ctl);
  ws.t_end();
  ws.t_begin("huff_stuff");
  k_stuff_lb<<<kPersistentCtas, kScanTile, 0, st>>>(stream, ctl, stuff_status, job->d_scan, (unsigned)cap);
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
