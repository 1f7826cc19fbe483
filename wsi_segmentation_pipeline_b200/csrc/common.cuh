// common.cuh — shared helpers for libwsi_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wsi_b200.h"

namespace wsi {

using bf16 = __nv_bfloat16;

// ---- error plumbing ---------------------------------------------------------------------------
struct Error {
  int status;
  std::string msg;
};
void set_global_error(const std::string& m);

#define WSI_THROW(status, ...)                                   \
  do {                                                           \
    char _b[512];                                                \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                       \
    throw ::wsi::Error{(status), std::string(_b)};               \
  } while (0)

#define CUDA_CHECK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      WSI_THROW(WSI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define WSI_REQUIRE(cond, status, ...) \
  do {                                 \
    if (!(cond)) WSI_THROW((status), __VA_ARGS__); \
  } while (0)

// Timing-only experiment switches inside the conv kernels (WSI_STREAM_DBG / WSI_UP_DBG: skip MMAs, skip loads ... —
// results are garbage) exist only in builds with -DWSI_DEBUG_SWITCHES; release builds compile them out.
#ifdef WSI_DEBUG_SWITCHES
#define WSI_DBG(p) ((p).dbg)
#else
#define WSI_DBG(p) 0
#endif

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// host bf16 round-to-nearest-even (same as __float2bfloat16_rn for finite values)
inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}
inline float bf16_bits_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// ---- device buffer (RAII) ---------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool owned = true;        // false: a view into another buffer (never freed here)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), owned(o.owned) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; owned = o.owned; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    bytes = 0;
    owned = true;
  }
  void alloc(size_t n) {
    if (n <= bytes && p) return;
    release();
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      p = nullptr;
      WSI_THROW(WSI_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    }
    bytes = n;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <class T>
inline void upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s = 0) {
  b.alloc(v.size() * sizeof(T));
  if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

// ---- launch counter (bench.py "gpu_launches") --------------------------------------------------
struct LaunchCounter {
  int64_t n = 0;
};

}  // namespace wsi
