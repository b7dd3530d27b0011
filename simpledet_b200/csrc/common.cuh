// Shared host-side plumbing for the C-ABI library: thread-local error string, launch counter,
// argument-check macros.  No framework types; only the CUDA runtime.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "simpledet_b200.h"

namespace sdet {

extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace sdet

#define SDET_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return sdet::fail(SDET_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define SDET_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return sdet::fail(SDET_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)

// Launch check: catches configuration errors synchronously without synchronising the stream.
#define SDET_LAUNCH_CHECK(name)                                                               \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return sdet::fail(SDET_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    sdet::count_launch();                                                                     \
  } while (0)
