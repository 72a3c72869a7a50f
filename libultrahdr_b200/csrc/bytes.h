// Host-side byte plumbing without the heap: the marker / container writers fill storage the caller owns (a
// stack array or a block of the workspace's pinned arena), look-ups hand out views into the input stream.
// (the ISO 21496-1 / MPF / ICC / splitter layer, "allocation-free".)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace uhdr_b200 {

// Fixed-capacity big-endian byte sink.  Writing past the capacity is counted, not performed: the caller
// checks ok() once at the end.
struct ByteSink {
  uint8_t* p;
  size_t cap, n = 0;
  ByteSink(uint8_t* buf, size_t capacity) : p(buf), cap(capacity) {}
  bool ok() const { return n <= cap; }
  size_t size() const { return n; }
  const uint8_t* data() const { return p; }
  void u8(unsigned b) {
    if (n < cap) p[n] = (uint8_t)b;
    n++;
  }
  void u16(unsigned w) { u8(w >> 8); u8(w & 0xff); }
  void u32(uint32_t w) { u16(w >> 16); u16(w & 0xffff); }
  void raw(const void* src, size_t len) {
    if (n + len <= cap) memcpy(p + n, src, len);
    n += len;
  }
};

// a view into somebody else's bytes
struct ByteView {
  const uint8_t* data = nullptr;
  size_t size = 0;
  bool empty() const { return size == 0; }
};

// std::map<int, V> for the keys 0..N-1 without node allocations (the find / end / [] / clear subset the C API uses)
template <class V, int N>
struct SlotMap {
  struct Entry {
    int first = 0;
    V second{};
    bool used = false;
  };
  Entry e[N];
  using iterator = Entry*;
  iterator end() { return e + N; }
  iterator find(int k) { return (k >= 0 && k < N && e[k].used) ? e + k : end(); }
  V& operator[](int k) {
    e[k].used = true;
    e[k].first = k;
    return e[k].second;
  }
  template <class F>
  void clear(F reset) {
    for (Entry& x : e) {
      x.used = false;
      reset(x.second);
    }
  }
  void clear() {
    clear([](V& v) { v = V(); });
  }
};

}  // namespace uhdr_b200
