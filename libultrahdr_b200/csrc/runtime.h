// Device runtime: per-device LUT residency, a grow-only arena per worker (device + pinned host)
// and a stream.  180 GB of HBM per GPU means we never free inside a codec call: arenas are
// rewound, not released.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include <cstddef>
#include <string>
#include <vector>

namespace uhdr_b200 {

// uhdr_codec_err_t values
enum : int { E_OK = 0, E_ERROR = 1, E_UNKNOWN = 2, E_INVALID_PARAM = 3, E_MEM = 4,
             E_INVALID_OP = 5, E_UNSUPPORTED = 6 };

void set_last_error(const std::string& s);
const char* last_error();
int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::uhdr_b200::fail(_e == cudaErrorMemoryAllocation ? E_MEM : E_ERROR,          \
                               "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e),      \
                               __FILE__, __LINE__, #expr);                                 \
  } while (0)

// Device-resident LUT blob for the current device (built on first use, or installed from a
// broadcast).  Returns nullptr + sets last error when no CUDA device is usable.
const float* device_luts();
int install_luts_from_device(const void* dptr);
int read_back_luts(float* host_out);

void set_kernel_timing(bool on);
bool kernel_timing_enabled();
// "name count total_ms\n" lines; resets the accumulators when `reset`
std::string kernel_timing_report(bool reset);

#define TIMED(ws, name, call)        \
  do {                               \
    (ws).t_begin(name);              \
    cudaError_t _te = (call);        \
    (ws).t_end();                    \
    CUDA_TRY(_te);                   \
  } while (0)

struct PhaseTrace {  // UHDR_B200_TRACE=1: wall-clock phases of one call on stderr
  bool on = getenv("UHDR_B200_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[uhdr_b200] %-26s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// release every parked arena block back to the driver; returns the bytes freed
size_t trim_parked_blocks();

class Arena {
 public:
  explicit Arena(bool pinned_host) : pinned_(pinned_host) {}
  ~Arena();
  void* alloc(size_t bytes, size_t align = 256);  // nullptr on failure (last error set)
  void rewind();
  // keep everything allocated so far, release what comes later (inputs stay resident while the
  // per-encode scratch is recycled)
  void set_floor();
  void clear_floor();
  size_t reserved() const;

 private:
  struct Block { char* base; size_t size, used, floor; };
  std::vector<Block> blocks_;
  bool pinned_;
  int device_ = -1;  // device the blocks belong to (set at the first allocation)
};

class Workspace {
 public:
  Workspace();
  ~Workspace();
  int init();  // binds to the current device, creates the stream
  cudaStream_t stream() const { return ext_stream_set_ ? ext_stream_ : stream_; }
  // run the following calls on a caller-owned stream (the *_dev stage entry points); clear_external_stream()
  // returns to the workspace's own
  void use_external_stream(cudaStream_t s) { ext_stream_ = s; ext_stream_set_ = true; }
  void clear_external_stream() { ext_stream_set_ = false; }
  void* dalloc(size_t bytes) { return dev_.alloc(bytes); }
  void* halloc(size_t bytes) { return host_.alloc(bytes); }
  void rewind() { dev_.rewind(); host_.rewind(); }
  void set_floor() { dev_.set_floor(); host_.set_floor(); }
  void clear_floor() { dev_.clear_floor(); host_.clear_floor(); }
  // per-kernel CUDA-event timing (enabled globally with set_kernel_timing); begin/end bracket one
  // launch on this workspace's stream, collect() must run after the stream was synchronised
  void t_begin(const char* name);
  void t_end();
  void t_collect();
  const float* luts() const { return luts_; }
  int sync();

 private:
  Arena dev_{false}, host_{true};
  cudaStream_t stream_ = nullptr;
  cudaStream_t ext_stream_ = nullptr;
  cudaEvent_t sync_ev_ = nullptr;   // UHDR_B200_BLOCKING_SYNC=1: sync() sleeps on this event instead of spinning
  bool ext_stream_set_ = false;
  const float* luts_ = nullptr;
  int device_ = -1;
  struct Span { const char* name; cudaEvent_t a, b; };
  std::vector<Span> spans_;
  std::vector<cudaEvent_t> ev_pool_;
  cudaEvent_t get_event();
};

}  // namespace uhdr_b200
