// Shared host-side plumbing for libmimamo_b200.so: error channel, launch counter, checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/mimamo_b200.h"

namespace mimamo {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define MM_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::mimamo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MIMAMO_E_CUDA;                                                              \
    }                                                                                    \
  } while (0)

#define MM_REQUIRE(cond, code, ...)                                                      \
  do {                                                                                   \
    if (!(cond)) { ::mimamo::set_error(__VA_ARGS__); return (code); }                    \
  } while (0)

// after a kernel launch: surface launch-configuration errors without synchronising
#define MM_LAUNCH_OK()                                                                   \
  do {                                                                                   \
    ::mimamo::count_launch();                                                            \
    MM_CUDA(cudaPeekAtLastError());                                                      \
  } while (0)

template <typename T>
inline int upload(T** dev, const T* host, size_t count) {
  MM_CUDA(cudaMalloc((void**)dev, count * sizeof(T)));
  MM_CUDA(cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
  return MIMAMO_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }

// One-time setup that CUDA keeps PER DEVICE (cudaFuncSetAttribute, __constant__ tables, occupancy queries): a bit per
// device ordinal instead of a process-wide flag, so a second GPU used by the same process is set up as well.
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  bool need() const { return !(done.load(std::memory_order_acquire) & (1ull << (current_device() & 63))); }
  void mark() { done.fetch_or(1ull << (current_device() & 63), std::memory_order_release); }
};

// Handles own device memory allocated on the device that was current at create time; calls must come with that
// device current (the Python layer creates handles per device).
#define MM_CHECK_DEVICE(owner)                                                                        \
  do {                                                                                                \
    const int _cur = ::mimamo::current_device();                                                      \
    if (_cur != (owner)) {                                                                            \
      ::mimamo::set_error("this handle lives on CUDA device %d but device %d is current", (owner), _cur); \
      return MIMAMO_E_VALUE;                                                                          \
    }                                                                                                 \
  } while (0)

}  // namespace mimamo
