#include "runtime.h"

#include <cstdarg>
#include <cstdlib>
#include <atomic>
#include <cstdio>
#include <new>
#include <map>
#include <mutex>

#include "tables.h"

namespace uhdr_b200 {

static thread_local std::string g_err;
void set_last_error(const std::string& s) { g_err = s; }
const char* last_error() { return g_err.c_str(); }
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ---- LUT residency ------------------------------------------------------------------------------
static std::mutex g_lut_mu;
static std::map<int, float*> g_luts;  // device ordinal -> device blob

static float* lut_slot(int dev) {
  auto it = g_luts.find(dev);
  if (it != g_luts.end()) return it->second;
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float) * kLutTotalFloats) != cudaSuccess) return nullptr;
  g_luts[dev] = d;
  return d;
}

const float* device_luts() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    fail(E_ERROR, "no usable CUDA device: %s (libuhdr_b200 has no CPU fallback)", cudaGetErrorString(e));
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_lut_mu);
  auto it = g_luts.find(dev);
  if (it != g_luts.end()) return it->second;
  float* d = lut_slot(dev);
  if (!d) {
    fail(E_MEM, "cudaMalloc of LUT blob failed");
    return nullptr;
  }
  std::vector<float> host(kLutTotalFloats);
  build_lut_blob(host.data());
  e = cudaMemcpy(d, host.data(), sizeof(float) * kLutTotalFloats, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    fail(E_ERROR, "LUT upload failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  return d;
}

int install_luts_from_device(const void* dptr) {
  int dev = -1;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_lut_mu);
  float* d = lut_slot(dev);
  if (!d) return fail(E_MEM, "cudaMalloc of LUT blob failed");
  CUDA_TRY(cudaMemcpy(d, dptr, sizeof(float) * kLutTotalFloats, cudaMemcpyDeviceToDevice));
  return E_OK;
}

int read_back_luts(float* host_out) {
  const float* d = device_luts();
  if (!d) return E_ERROR;
  CUDA_TRY(cudaMemcpy(host_out, d, sizeof(float) * kLutTotalFloats, cudaMemcpyDeviceToHost));
  return E_OK;
}

// ---- arenas -------------------------------------------------------------------------------------
// Blocks of released arenas are parked in a process-wide cache instead of going back to the driver:
// the reference's usage pattern is create handle / run / release per image, and pinning a few
// hundred MB of host memory (or cudaMalloc of as much HBM) costs far more than the decode itself.
namespace {
struct ParkedBlock { char* base; size_t size; int device; };
std::mutex g_park_mu;
std::vector<ParkedBlock> g_parked[2];              // [0] device memory, [1] pinned host memory
size_t g_parked_bytes[2] = {0, 0};
constexpr size_t kParkCap[2] = {(size_t)8 << 30, (size_t)4 << 30};

char* take_parked(bool pinned, int device, size_t bytes, size_t* got) {
  std::lock_guard<std::mutex> lk(g_park_mu);
  auto& v = g_parked[pinned ? 1 : 0];
  int best = -1;
  for (int i = 0; i < (int)v.size(); i++) {
    if (v[i].size < bytes || (!pinned && v[i].device != device)) continue;
    if (best < 0 || v[i].size < v[best].size) best = i;
  }
  // a handle of the same kind asks for the same block sizes in the same order, so near-exact matches
  // are the rule; handing a much larger block to a small request would only force the next large
  // request back to the driver
  if (best < 0 || v[best].size > bytes + bytes / 4 + ((size_t)1 << 20)) return nullptr;
  char* base = v[best].base;
  *got = v[best].size;
  g_parked_bytes[pinned ? 1 : 0] -= v[best].size;
  v.erase(v.begin() + best);
  return base;
}
bool park(bool pinned, int device, char* base, size_t size) {
  std::lock_guard<std::mutex> lk(g_park_mu);
  const int k = pinned ? 1 : 0;
  if (g_parked_bytes[k] + size > kParkCap[k]) return false;
  g_parked[k].push_back({base, size, device});
  g_parked_bytes[k] += size;
  return true;
}
}  // namespace

size_t trim_parked_blocks() {
  std::vector<ParkedBlock> take[2];
  {
    std::lock_guard<std::mutex> lk(g_park_mu);
    for (int k = 0; k < 2; k++) {
      take[k].swap(g_parked[k]);
      g_parked_bytes[k] = 0;
    }
  }
  size_t freed = 0;
  int cur = -1;
  cudaGetDevice(&cur);
  for (auto& b : take[0]) {
    if (b.device >= 0 && b.device != cur) cudaSetDevice(b.device);
    cudaFree(b.base);
    freed += b.size;
    if (b.device >= 0 && b.device != cur && cur >= 0) cudaSetDevice(cur);
  }
  for (auto& b : take[1]) {
    cudaFreeHost(b.base);
    freed += b.size;
  }
  return freed;
}

// Callers (Workspace) drain their stream before the arenas go away: a parked block may be handed to
// another handle on another stream at once, and unlike cudaFree parking does not synchronise.
Arena::~Arena() {
  for (auto& b : blocks_) {
    if (park(pinned_, device_, b.base, b.size)) continue;
    if (pinned_) cudaFreeHost(b.base);
    else cudaFree(b.base);
  }
}
void* Arena::alloc(size_t bytes, size_t align) {
  if (bytes == 0) bytes = 1;
  for (auto& b : blocks_) {
    size_t off = (b.used + align - 1) / align * align;
    if (off + bytes <= b.size) {
      b.used = off + bytes;
      return b.base + off;
    }
  }
  const size_t min_block = pinned_ ? (size_t)32 << 20 : (size_t)64 << 20;
  size_t sz = bytes > min_block ? bytes : min_block;
  sz = (sz + 4095) / 4096 * 4096;
  if (device_ < 0) cudaGetDevice(&device_);
  char* base = take_parked(pinned_, device_, sz, &sz);
  if (!base) {
    cudaError_t e = pinned_ ? cudaHostAlloc((void**)&base, sz, cudaHostAllocPortable)
                            : cudaMalloc((void**)&base, sz);
    if (e != cudaSuccess) {
      fail(E_MEM, "%s of %zu bytes failed: %s", pinned_ ? "cudaHostAlloc" : "cudaMalloc", sz,
           cudaGetErrorString(e));
      return nullptr;
    }
  }
  blocks_.push_back({base, sz, bytes, 0});
  return base;
}
void Arena::rewind() {
  for (auto& b : blocks_) b.used = b.floor;
}
void Arena::set_floor() {
  for (auto& b : blocks_) b.floor = b.used;
}
void Arena::clear_floor() {
  for (auto& b : blocks_) b.floor = b.used = 0;
}
size_t Arena::reserved() const {
  size_t s = 0;
  for (auto& b : blocks_) s += b.size;
  return s;
}

Workspace::Workspace() {}
Workspace::~Workspace() {
  if (stream_) {
    // error returns can leave kernels / async copies in flight on this stream; they must have
    // finished before the arena blocks (destroyed after this body) are parked for other handles
    cudaStreamSynchronize(stream_);
    cudaStreamDestroy(stream_);
  }
  for (auto& s : spans_) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
  if (sync_ev_) cudaEventDestroy(sync_ev_);
}
int Workspace::init() {
  if (stream_) return E_OK;
  luts_ = device_luts();
  if (!luts_) return E_ERROR;
  CUDA_TRY(cudaGetDevice(&device_));
  CUDA_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  return E_OK;
}
// A host thread waiting for its stream normally spins (lowest latency).  With many handles per GPU and several GPUs
// per host that is one busy core per waiting thread; UHDR_B200_BLOCKING_SYNC=1 makes the wait a sleep on an event
// created with cudaEventBlockingSync (a few tens of microseconds more per wait, no core burnt).
static bool blocking_sync_wanted() {
  static const bool on = [] { const char* e = getenv("UHDR_B200_BLOCKING_SYNC"); return e && *e && *e != '0'; }();
  return on;
}
int Workspace::sync() {
  if (blocking_sync_wanted()) {
    if (!sync_ev_) CUDA_TRY(cudaEventCreateWithFlags(&sync_ev_, cudaEventBlockingSync | cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(sync_ev_, stream()));
    CUDA_TRY(cudaEventSynchronize(sync_ev_));
  } else {
    CUDA_TRY(cudaStreamSynchronize(stream()));
  }
  if (!spans_.empty()) t_collect();
  return E_OK;
}
// ---- kernel timing ------------------------------------------------------------------------------
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mu;
struct TimingAcc { unsigned long long n = 0; double total = 0, mn = 1e30, mx = 0; };
static std::map<std::string, TimingAcc> g_timing_acc;
void set_kernel_timing(bool on) { g_timing = on; }
bool kernel_timing_enabled() { return g_timing; }
std::string kernel_timing_report(bool reset) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  std::string out;
  char line[256];
  for (auto& kv : g_timing_acc) {
    snprintf(line, sizeof line, "%s %llu %.6f %.6f %.6f\n", kv.first.c_str(), kv.second.n, kv.second.total, kv.second.mn, kv.second.mx);
    out += line;
  }
  if (reset) g_timing_acc.clear();
  return out;
}
cudaEvent_t Workspace::get_event() {
  if (!ev_pool_.empty()) {
    cudaEvent_t e = ev_pool_.back();
    ev_pool_.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void Workspace::t_begin(const char* name) {
  if (!g_timing) return;
  Span s{name, get_event(), get_event()};
  cudaEventRecord(s.a, stream());
  spans_.push_back(s);
}
void Workspace::t_end() {
  if (!g_timing || spans_.empty()) return;
  cudaEventRecord(spans_.back().b, stream());
}
void Workspace::t_collect() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& s : spans_) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      auto& acc = g_timing_acc[s.name];
      acc.n++;
      acc.total += ms;
      if (ms < acc.mn) acc.mn = ms;
      if (ms > acc.mx) acc.mx = ms;
    }
    ev_pool_.push_back(s.a);
    ev_pool_.push_back(s.b);
  }
  spans_.clear();
}

}  // namespace uhdr_b200
