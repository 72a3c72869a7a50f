// Baseline-JPEG entropy *decoding* on the device: the serial half of JpegDecoderHelper
// (jpegdecoderhelper.cpp:397-411 -> libjpeg-turbo jdhuff.c decode_mcu) restated as a
// data-parallel fixed-point iteration.
//
// A Huffman-coded scan has no block index: a decoder must know where the previous symbol ended.
// But a decoder started at a wrong place falls into step with the true symbol sequence after a
// few blocks (codes self-synchronise; an end-of-block resets the coefficient index; a wrong
// position inside the MCU meets the wrong table at the next luma/chroma change and is knocked
// out of step again until it lands on the right one).  So:
//   1. the host removes the FF00 stuffing while staging the scan in pinned memory
//   2. the scan is cut into subsequences of kSeqBits bits, one thread each.  out[i] is the
//      decoder state (bit position, coefficient index, block-in-MCU) at which subsequence i+1
//      starts.  Every round each thread re-decodes its subsequence from out[i-1] and replaces
//      out[i]; subsequence 0 always starts from the true state, so the true states spread from
//      the left at least one subsequence per round -- and in practice in a few rounds, because
//      most exit states are already right.  A round that changes nothing is a fixed point, and
//      the only fixed point is the sequential decoder's state sequence (induction over i).
//   3. the per-subsequence block counts are prefix-summed, and a last pass decodes once more,
//      now writing coefficients ([block][64], natural order) and DC differences
//   4. DC prediction (a running sum per component over the scan order, dummy edge blocks
//      included like jdhuff.c) is a segmented prefix sum + scatter.
// Streams with restart markers, or that do not reach a fixed point in kMaxRounds rounds, return
// kHuffDecFallback and the caller uses the host decoder of jpeg_host.cpp.
#include <atomic>
#include <mutex>
#include <climits>
#include <cstring>

#include "jpeg.h"
#include "runtime.h"

namespace uhdr_b200 {

__constant__ uint8_t kZigzagDev[64];  // zigzag position -> natural index (copy of kZigzag)

namespace {

constexpr int kSeqBits = 1024;
constexpr int kLutBits = 9;
constexpr int kRoundsPerBatch = 16;  // settled rounds cost ~a launch each (CTAs without work exit at once): one host check usually suffices
constexpr int kMaxRounds = 512;

struct HdTables {  // 0 DC table 0, 1 DC table 1, 2 AC table 0, 3 AC table 1
  uint16_t lut[4][1 << kLutBits];  // (len << 8) | symbol; 0 = code longer than kLutBits bits
  int maxcode[4][18];              // canonical decode of the long codes; [17] = INT_MAX
  int valoff[4][17];               // index of a code's symbol = valoff[len] + code
  uint8_t vals[4][256];
};
struct HdFrame {
  int bpm;                       // blocks per MCU
  int comp_of[10], bi[10], bj[10], kin[10];  // per MCU position: component, block offset, index within component
  int dc_tab[3], ac_tab[3];      // indices into HdTables
  int h[3], v[3], hv[3], wblocks[3], hblocks[3];
  int mcus_per_row;
  unsigned total_bits, nseq, total_blocks;
};
struct HdShared {
  HdTables t;
  HdFrame f;
};

__device__ __forceinline__ unsigned peek32(const uint32_t* __restrict__ bits, unsigned p) {
  const unsigned k = p >> 5;
  const uint32_t w0 = __byte_perm(__ldg(bits + k), 0, 0x0123), w1 = __byte_perm(__ldg(bits + k + 1), 0, 0x0123);
  return __funnelshift_l(w1, w0, p & 31);
}
__device__ __forceinline__ int extend(unsigned v, unsigned s) {  // jdhuff.c HUFF_EXTEND
  return v < (1u << (s - 1)) ? (int)v - (int)(1u << s) + 1 : (int)v;
}

struct HdOut {  // WRITE pass destinations
  int16_t* coefs[3];
  int* dcd[3];
  unsigned* err;
};

// Decodes the symbols that start in [p, end_bit).  State in/out: p, z (0 = DC symbol next), c.
template <bool WRITE>
__device__ __forceinline__ unsigned decode_seq(const HdShared& S, const uint32_t* __restrict__ bits, unsigned& p, unsigned& z,
                                               unsigned& c, const unsigned end_bit, unsigned b, const HdOut& o) {
  const HdFrame& f = S.f;
  unsigned nblk = 0;
  int16_t* blk = nullptr;
  unsigned mcu = 0;
  auto locate = [&]() {  // block b at MCU position c -> coefficient block (or nullptr for dummy / surplus blocks)
    const int comp = f.comp_of[c];
    const unsigned mx = mcu % (unsigned)f.mcus_per_row, my = mcu / (unsigned)f.mcus_per_row;
    const int bx = (int)mx * f.h[comp] + f.bi[c], by = (int)my * f.v[comp] + f.bj[c];
    blk = (b < f.total_blocks && bx < f.wblocks[comp] && by < f.hblocks[comp]) ? o.coefs[comp] + ((size_t)by * f.wblocks[comp] + bx) * 64 : nullptr;
  };
  if (WRITE) {
    mcu = b / (unsigned)f.bpm;
    locate();
  }
  while (p < end_bit) {
    const unsigned w = peek32(bits, p);
    const int comp = f.comp_of[c];
    const int t = z == 0 ? f.dc_tab[comp] : f.ac_tab[comp];
    const unsigned e = S.t.lut[t][w >> (32 - kLutBits)];
    unsigned len = e >> 8, sym = e & 0xff;
    if (len == 0) {  // long code: canonical search, like jdhuff.c's slow path
      len = kLutBits + 1;
      int code = (int)(w >> (32 - len));
      while (code > S.t.maxcode[t][len]) {
        len++;
        code = (int)(w >> (32 - len));
      }
      if (len > 16) {  // not a code of this table (possible only while out of step, or corrupt data)
        len = 16;
        sym = 0;
        if (WRITE) *o.err = 1;
      } else {
        sym = S.t.vals[t][(S.t.valoff[t][len] + code) & 255];
      }
    }
    const unsigned s = sym & 15;
    const unsigned v = s ? (w << len) >> (32 - s) : 0;
    p += len + s;
    if (z == 0) {
      if (WRITE && b < f.total_blocks) o.dcd[comp][(size_t)mcu * f.hv[comp] + f.kin[c]] = s ? extend(v, s) : 0;
      z = 1;
    } else {
      const unsigned r = sym >> 4;
      if (s) {
        z += r;
        if (WRITE) {
          if (z > 63) *o.err = 1;
          else if (blk) blk[kZigzagDev[z]] = (int16_t)extend(v, s);
        }
        z++;
      } else {
        z = r == 15 ? z + 16 : 64;
      }
    }
    if (z >= 64) {
      z = 0;
      nblk++;
      c = c + 1 == (unsigned)f.bpm ? 0 : c + 1;
      if (WRITE) {
        b++;
        if (c == 0) mcu++;
        locate();
      }
    }
  }
  return nblk;
}

__device__ __forceinline__ void stage_shared(HdShared& S, const HdShared* __restrict__ g) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&S);
  for (unsigned i = threadIdx.x; i < sizeof(HdShared) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
}

// one relaxation round, in place (64-bit states are read and written atomically)
__global__ void __launch_bounds__(128) k_hd_sync(const uint32_t* __restrict__ bits, unsigned long long* out, unsigned long long* used, unsigned* cnt,
                                                 unsigned* changed, const HdShared* __restrict__ gs) {
  __shared__ HdShared S;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned nseq = gs->f.nseq;
  // state word: bit position | (z | c << 8) << 32
  const unsigned long long entry = (i == 0 || i >= nseq) ? 0ull : *reinterpret_cast<volatile unsigned long long*>(out + i - 1);
  const bool todo = i < nseq && entry != used[i];  // same start as last time: same result
  // after the second round nearly every subsequence is settled: a CTA without work leaves before it
  // stages the 6 KB of tables, so the later rounds cost little more than their launch
  if (!__syncthreads_or(todo ? 1 : 0)) return;
  stage_shared(S, gs);
  if (!todo) return;
  unsigned p = (unsigned)entry, z = (unsigned)(entry >> 32) & 0xff, c = (unsigned)(entry >> 40);
  const unsigned end_bit = min((i + 1) * (unsigned)kSeqBits, S.f.total_bits);
  const HdOut none = {};
  const unsigned n = decode_seq<false>(S, bits, p, z, c, end_bit, 0, none);
  const unsigned long long now = (unsigned long long)p | ((unsigned long long)(z | (c << 8)) << 32);
  used[i] = entry;
  cnt[i] = n;
  if (now != out[i]) {
    *reinterpret_cast<volatile unsigned long long*>(out + i) = now;
    *changed = 1;
  }
}

__global__ void k_hd_init(unsigned long long* out, unsigned long long* used, unsigned* cnt, unsigned nseq) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq) return;
  out[i] = (unsigned long long)((i + 1) * (unsigned)kSeqBits);
  used[i] = ~0ull;
  cnt[i] = 0;
}

// exclusive prefix sum of cnt (one CTA; nseq is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) k_hd_scan(const unsigned* __restrict__ cnt, unsigned* __restrict__ base, unsigned n, unsigned* total) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned start = 0; start < n; start += 1024) {
    const unsigned i = start + threadIdx.x;
    const unsigned v = i < n ? cnt[i] : 0;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned s = warp_sums[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      warp_sums[threadIdx.x] = s;
    }
    __syncthreads();
    const unsigned wbase = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) base[i] = carry + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += wbase + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(128) k_hd_write(const uint32_t* __restrict__ bits, const unsigned long long* __restrict__ out, const unsigned* __restrict__ base,
                                                  const HdShared* __restrict__ gs, HdOut o) {
  __shared__ HdShared S;
  stage_shared(S, gs);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.f.nseq) return;
  const unsigned long long entry = i == 0 ? 0ull : out[i - 1];
  unsigned p = (unsigned)entry, z = (unsigned)(entry >> 32) & 0xff, c = (unsigned)(entry >> 40);
  const unsigned b = base[i];
  if (b % (unsigned)S.f.bpm != c) *o.err = 2;  // the states and the block counts must agree
  const unsigned end_bit = min((i + 1) * (unsigned)kSeqBits, S.f.total_bits);
  decode_seq<true>(S, bits, p, z, c, end_bit, b, o);
}

// ---- DC prediction: inclusive scan of the differences per component, scatter into the blocks ----
struct DcPlan {
  int* dcd;
  int16_t* coefs;
  unsigned n;  // blocks of this component in scan order (dummy blocks included)
  int h, v, hv, wblocks, hblocks, mcus_per_row;
  int* sums;   // per-CTA totals
};
constexpr int kDcCta = 1024;
__device__ __forceinline__ int cta_inclusive_scan(int v, int* warp_sums /* [32] */) {
  int x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int s = warp_sums[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (threadIdx.x >= o) s += y;
    }
    warp_sums[threadIdx.x] = s;
  }
  __syncthreads();
  return x + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0);
}
__global__ void __launch_bounds__(kDcCta) k_dc_local(DcPlan d) {
  __shared__ int ws[32];
  const unsigned i = blockIdx.x * kDcCta + threadIdx.x;
  const int x = cta_inclusive_scan(i < d.n ? d.dcd[i] : 0, ws);
  if (i < d.n) d.dcd[i] = x;
  if (threadIdx.x == kDcCta - 1) d.sums[blockIdx.x] = x;
}
__global__ void __launch_bounds__(kDcCta) k_dc_sums(int* sums, unsigned n) {  // in place, exclusive, one CTA
  __shared__ int ws[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned start = 0; start < n; start += kDcCta) {
    const unsigned i = start + threadIdx.x;
    const int v = i < n ? sums[i] : 0;
    const int x = cta_inclusive_scan(v, ws);
    if (i < n) sums[i] = carry + x - v;
    __syncthreads();
    if (threadIdx.x == kDcCta - 1) carry += x;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_dc_apply(DcPlan d) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  const int dc = d.dcd[i] + d.sums[i / kDcCta];
  const unsigned mcu = i / (unsigned)d.hv, k = i % (unsigned)d.hv;
  const int bx = (int)(mcu % (unsigned)d.mcus_per_row) * d.h + (int)(k % (unsigned)d.h);
  const int by = (int)(mcu / (unsigned)d.mcus_per_row) * d.v + (int)(k / (unsigned)d.h);
  if (bx < d.wblocks && by < d.hblocks) d.coefs[((size_t)by * d.wblocks + bx) * 64] = (int16_t)dc;
}

void build_tables(const JpegHeader& h, HdTables* t) {
  memset(t, 0, sizeof *t);
  for (int cls = 0; cls < 2; cls++)
    for (int id = 0; id < 2; id++) {
      const int ti = cls * 2 + id;
      for (int l = 0; l < 18; l++) t->maxcode[ti][l] = -1;
      t->maxcode[ti][17] = INT_MAX;
      if (!h.have_tbl[cls][id]) {  // absent table: every lookup is a 16-bit "invalid" (never selected by a valid header)
        for (int l = 0; l < 17; l++) t->maxcode[ti][l] = -1;
        continue;
      }
      const uint8_t* bits = h.bits[cls][id];
      memcpy(t->vals[ti], h.vals[cls][id], 256);
      int code = 0, k = 0;
      for (int len = 1; len <= 16; len++) {
        t->valoff[ti][len] = k - code;
        for (int i = 0; i < bits[len]; i++, k++, code++)
          if (len <= kLutBits && code < (1 << len))  // guard: never index past `lut` whatever the lengths say
            for (int r = 0; r < (1 << (kLutBits - len)); r++)
              t->lut[ti][(code << (kLutBits - len)) | r] = (uint16_t)((len << 8) | h.vals[cls][id][k & 255]);
        t->maxcode[ti][len] = bits[len] ? code - 1 : -1;
        code <<= 1;
      }
    }
}

}  // namespace

namespace {
std::atomic<unsigned long long> g_hd_done{0}, g_hd_declined{0}, g_hd_rounds{0};
int declined() {
  g_hd_declined.fetch_add(1);
  return kHuffDecFallback;
}
}  // namespace
void jpeg_entropy_decoder_stats(unsigned long long out[3]) {
  out[0] = g_hd_done.load();
  out[1] = g_hd_declined.load();
  out[2] = g_hd_rounds.load();
}

// Removes byte stuffing from the entropy-coded segment that starts at data[from].  Returns the
// clean length, or -1 when a restart marker is met.  *consumed = offset of the terminating marker.
static long unstuff_scan(const uint8_t* data, size_t size, size_t from, uint8_t* dst) {
  size_t p = from;
  uint8_t* o = dst;
  while (p < size) {
    const uint8_t* ff = (const uint8_t*)memchr(data + p, 0xFF, size - p);
    const size_t run = ff ? (size_t)(ff - (data + p)) : size - p;
    memcpy(o, data + p, run);
    o += run;
    p += run;
    if (!ff) break;
    if (p + 1 >= size) break;  // FF at the very end: treat as end of data
    const uint8_t nx = data[p + 1];
    if (nx == 0x00) {
      *o++ = 0xFF;
      p += 2;
    } else if (nx == 0xFF) {
      p += 1;  // fill byte before a marker
    } else if (nx >= 0xD0 && nx <= 0xD7) {
      return -1;
    } else {
      break;  // EOI or any other marker ends the segment
    }
  }
  return (long)(o - dst);
}

int jpeg_entropy_decode_dev(Workspace& ws, const uint8_t* data, size_t size, const JpegHeader& h, int16_t* d_coefs[3]) {
  const JpegFrame& f = h.frame;
  if (h.restart_interval) return declined();
  HdShared hs;
  memset(&hs, 0, sizeof hs);
  HdFrame& hf = hs.f;
  int bpm = 0;
  for (int c = 0; c < f.ncomp; c++) {
    const JpegComp& k = f.comp[c];
    const int mw = f.ncomp == 1 ? 1 : k.h_samp, mh = f.ncomp == 1 ? 1 : k.v_samp;
    hf.h[c] = mw;
    hf.v[c] = mh;
    hf.hv[c] = mw * mh;
    hf.wblocks[c] = k.wblocks;
    hf.hblocks[c] = k.hblocks;
    if (h.dc_sel[c] < 0 || h.dc_sel[c] > 1 || h.ac_sel[c] < 0 || h.ac_sel[c] > 1) return declined();
    if (!h.have_tbl[0][h.dc_sel[c]] || !h.have_tbl[1][h.ac_sel[c]]) return declined();
    hf.dc_tab[c] = h.dc_sel[c];
    hf.ac_tab[c] = 2 + h.ac_sel[c];
    for (int j = 0; j < mh; j++)
      for (int i = 0; i < mw; i++) {
        if (bpm >= 10) return declined();
        hf.comp_of[bpm] = c;
        hf.bi[bpm] = i;
        hf.bj[bpm] = j;
        hf.kin[bpm] = j * mw + i;
        bpm++;
      }
  }
  hf.bpm = bpm;
  hf.mcus_per_row = f.mcus_per_row;
  const size_t mcus = (size_t)f.mcus_per_row * f.mcu_rows;
  if (mcus * bpm > 0xfffffff0u) return declined();
  hf.total_blocks = (unsigned)(mcus * bpm);
  build_tables(h, &hs.t);

  // 1. clean bit stream in pinned memory, then on the device (8 zero bytes of slack for the reader)
  if (size <= h.scan_offset) return fail(E_ERROR, "Corrupt JPEG data: no entropy-coded segment");
  const size_t cap = size - h.scan_offset + 16;
  uint8_t* h_bits = (uint8_t*)ws.halloc(cap);
  if (!h_bits) return E_MEM;
  PhaseTrace tr;
  const long clean = unstuff_scan(data, size, h.scan_offset, h_bits);
  tr.mark("  unstuff");
  if (clean < 0) return declined();
  if ((size_t)clean * 8 > 0xfffffff0u - kSeqBits) return declined();
  memset(h_bits + clean, 0, 16);
  const size_t padded = ((size_t)clean + 16 + 3) & ~(size_t)3;
  hf.total_bits = (unsigned)(clean * 8);
  hf.nseq = (hf.total_bits + kSeqBits - 1) / kSeqBits;
  if (hf.nseq == 0) return fail(E_ERROR, "Corrupt JPEG data: empty entropy-coded segment");
  const unsigned nseq = hf.nseq;

  uint32_t* d_bits = (uint32_t*)ws.dalloc(padded);
  HdShared* d_hs = (HdShared*)ws.dalloc(sizeof(HdShared));
  HdShared* h_hs = (HdShared*)ws.halloc(sizeof(HdShared));
  unsigned long long* d_out = (unsigned long long*)ws.dalloc(sizeof(unsigned long long) * nseq);
  unsigned long long* d_used = (unsigned long long*)ws.dalloc(sizeof(unsigned long long) * nseq);
  unsigned* d_cnt = (unsigned*)ws.dalloc(sizeof(unsigned) * nseq);
  unsigned* d_base = (unsigned*)ws.dalloc(sizeof(unsigned) * nseq);
  unsigned* d_flags = (unsigned*)ws.dalloc(sizeof(unsigned) * (kMaxRounds + 8));
  unsigned* h_flags = (unsigned*)ws.halloc(sizeof(unsigned) * (kMaxRounds + 8));
  if (!d_bits || !d_hs || !h_hs || !d_out || !d_used || !d_cnt || !d_base || !d_flags || !h_flags) return E_MEM;
  memcpy(h_hs, &hs, sizeof hs);
  cudaStream_t s = ws.stream();
  CUDA_TRY(cudaMemcpyAsync(d_bits, h_bits, padded, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d_hs, h_hs, sizeof hs, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(unsigned) * (kMaxRounds + 8), s));
  {  // once per device, synchronously and under a lock: decodes run concurrently on several streams / threads
    static std::mutex zig_mu;
    static bool zig_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(zig_mu);
    if (dev < 0 || dev >= 64 || !zig_done[dev]) {
      CUDA_TRY(cudaMemcpyToSymbol(kZigzagDev, kZigzag, 64, 0, cudaMemcpyHostToDevice));
      if (dev >= 0 && dev < 64) zig_done[dev] = true;
    }
  }
  // coefficient blocks start as zeros; only non-zero coefficients are written
  int* d_dcd[3] = {nullptr, nullptr, nullptr};
  for (int c = 0; c < f.ncomp; c++) {
    d_coefs[c] = (int16_t*)ws.dalloc(f.blocks(c) * 128);
    d_dcd[c] = (int*)ws.dalloc(sizeof(int) * mcus * hf.hv[c]);
    if (!d_coefs[c] || !d_dcd[c]) return E_MEM;
    CUDA_TRY(cudaMemsetAsync(d_coefs[c], 0, f.blocks(c) * 128, s));
  }

  // 2. relaxation rounds until one of them changes nothing
  const unsigned grid = (nseq + 127) / 128;
  ws.t_begin("huffdec_sync");
  k_hd_init<<<(nseq + 255) / 256, 256, 0, s>>>(d_out, d_used, d_cnt, nseq);
  count_launches(1);
  int rounds = 0, first_quiet = 0;
  bool converged = false;
  while (!converged && rounds < kMaxRounds) {
    count_launches(kRoundsPerBatch);
    for (int r = 0; r < kRoundsPerBatch; r++) k_hd_sync<<<grid, 128, 0, s>>>(d_bits, d_out, d_used, d_cnt, d_flags + rounds + r, d_hs);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_flags + rounds, d_flags + rounds, sizeof(unsigned) * kRoundsPerBatch, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int r = 0; r < kRoundsPerBatch && !converged; r++)
      if (!h_flags[rounds + r]) {
        converged = true;
        first_quiet = rounds + r + 1;
      }
    rounds += kRoundsPerBatch;
  }
  ws.t_end();
  tr.mark("  relaxation rounds");
  if (!converged) return declined();

  // 3. block offsets, then the writing pass
  unsigned* d_total = d_flags + kMaxRounds;      // [0] total blocks, [1] error flag
  HdOut o;
  memset(&o, 0, sizeof o);
  for (int c = 0; c < f.ncomp; c++) { o.coefs[c] = d_coefs[c]; o.dcd[c] = d_dcd[c]; }
  o.err = d_total + 1;
  ws.t_begin("huffdec_write");
  count_launches(2 + 3 * f.ncomp);
  k_hd_scan<<<1, 1024, 0, s>>>(d_cnt, d_base, nseq, d_total);
  k_hd_write<<<grid, 128, 0, s>>>(d_bits, d_out, d_base, d_hs, o);
  ws.t_end();
  CUDA_TRY(cudaGetLastError());
  // 4. DC prediction
  ws.t_begin("huffdec_dc");
  for (int c = 0; c < f.ncomp; c++) {
    DcPlan d;
    d.dcd = d_dcd[c];
    d.coefs = d_coefs[c];
    d.n = (unsigned)(mcus * hf.hv[c]);
    d.h = hf.h[c]; d.v = hf.v[c]; d.hv = hf.hv[c];
    d.wblocks = hf.wblocks[c]; d.hblocks = hf.hblocks[c];
    d.mcus_per_row = hf.mcus_per_row;
    const unsigned nct = (d.n + kDcCta - 1) / kDcCta;
    d.sums = (int*)ws.dalloc(sizeof(int) * nct);
    if (!d.sums) return E_MEM;
    k_dc_local<<<nct, kDcCta, 0, s>>>(d);
    k_dc_sums<<<1, kDcCta, 0, s>>>(d.sums, nct);
    k_dc_apply<<<(d.n + 255) / 256, 256, 0, s>>>(d);
  }
  ws.t_end();
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h_flags, d_total, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  tr.mark("  write + dc");
  if (h_flags[1] || h_flags[0] < hf.total_blocks) return declined();  // let the host decoder produce the diagnosis
  g_hd_done.fetch_add(1);
  g_hd_rounds.store((unsigned long long)first_quiet);
  return E_OK;
}

}  // namespace uhdr_b200
