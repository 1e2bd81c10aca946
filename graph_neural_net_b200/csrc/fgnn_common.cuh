// Shared host/device helpers for libfgnn_b200 (internal; the public ABI is include/fgnn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/fgnn_b200.h"

namespace fgnn {

extern thread_local char g_last_error[512];
extern thread_local int64_t g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

#define FGNN_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return ::fgnn::fail(FGNN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define FGNN_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ::fgnn::fail(FGNN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                               \
  } while (0)

// every kernel launch in the library goes through this so gpu_launches is an honest count
#define FGNN_LAUNCHED()                                                                      \
  do {                                                                                       \
    ++::fgnn::g_launches;                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess)                                                                  \
      return ::fgnn::fail(FGNN_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                 \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                      \
  } while (0)

// Optional per-kernel-class device timing (CUDA events on the launching stream); off by default.
namespace prof {
enum Kind { kMlp = 0, kMatmul = 1, kStats = 2, kGlue = 3, kNumKinds = 4 };
void begin(Kind k, cudaStream_t st);
void end(Kind k, cudaStream_t st);
}  // namespace prof

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t size;
  size_t off = 0;
  bool dry;  // dry run: only measure
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t count, size_t align = 256) {
    off = align_up(off, align);
    T* r = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return dry || off <= size; }
};

}  // namespace fgnn
