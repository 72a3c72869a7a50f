#include "runtime.h"

#include <cstdarg>
#include <cstdio>
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
Arena::~Arena() {
  for (auto& b : blocks_) {
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
  char* base = nullptr;
  cudaError_t e = pinned_ ? cudaHostAlloc((void**)&base, sz, cudaHostAllocDefault)
                          : cudaMalloc((void**)&base, sz);
  if (e != cudaSuccess) {
    fail(E_MEM, "%s of %zu bytes failed: %s", pinned_ ? "cudaHostAlloc" : "cudaMalloc", sz,
         cudaGetErrorString(e));
    return nullptr;
  }
  blocks_.push_back({base, sz, bytes});
  return base;
}
void Arena::rewind() {
  for (auto& b : blocks_) b.used = 0;
}
size_t Arena::reserved() const {
  size_t s = 0;
  for (auto& b : blocks_) s += b.size;
  return s;
}

Workspace::Workspace() {}
Workspace::~Workspace() {
  if (stream_) cudaStreamDestroy(stream_);
}
int Workspace::init() {
  if (stream_) return E_OK;
  luts_ = device_luts();
  if (!luts_) return E_ERROR;
  CUDA_TRY(cudaGetDevice(&device_));
  CUDA_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  return E_OK;
}
int Workspace::sync() {
  CUDA_TRY(cudaStreamSynchronize(stream_));
  return E_OK;
}

}  // namespace uhdr_b200
