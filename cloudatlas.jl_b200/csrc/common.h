// Shared helpers for libcloudatlas_b200: status codes, thread-local error text, CUDA checks.
#pragma once
#include <functional>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/cloudatlas_b200.h"

namespace ca {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

void set_last_error(const std::string& msg);
[[noreturn]] void fail(int32_t code, const char* fmt, ...);

extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CA_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      ::ca::fail(CA_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                 cudaGetErrorString(e__));                                                \
  } while (0)

// Wraps the body of an extern "C" entry point: exceptions become status codes.
template <typename F>
int32_t guarded(F&& body) {
  try {
    body();
    return CA_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return CA_ERR_NOMEM;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return CA_ERR_INVALID;
  }
}

// Host worker threads for the build stages: CA_NUM_THREADS, else min(16, hardware threads).
unsigned worker_threads();
// fn(begin, end, chunk) over [0, n) cut into `chunks` contiguous pieces, on the worker threads (dynamic assignment)
void parallel_chunks(size_t n, size_t chunks, const std::function<void(size_t, size_t, size_t)>& fn);

// Wall-clock phase timer, printed to stderr when CA_TIMING is set (host-side build diagnostics).
struct PhaseTimer {
  const char* what;
  double t0;
  bool on;
  explicit PhaseTimer(const char* w);
  void lap(const char* phase);
};

// Requires a CUDA device of compute capability 10.x; the library has no other code path.
void require_device();

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold raised, so
// that the multi-gigabyte temporaries of assembly and model build are recycled inside the process: returning them to
// the driver with cudaFree and asking again costs up to a second per build at m ~ 3000 (fresh pages are mapped and
// scrubbed).  Frees keep cudaFree's semantics -- the device is synchronised first, because handles use non-blocking
// streams that the pool's stream order (legacy stream) does not see.  CA_NO_POOL=1 restores cudaMalloc / cudaFree.
void* device_alloc(size_t bytes);
void device_free(void* ptr);

template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) device_free(ptr);
    ptr = nullptr;
    count = 0;
  }
  void alloc(size_t n) {
    release();
    if (n == 0) return;
    ptr = static_cast<T*>(device_alloc(n * sizeof(T)));
    count = n;
  }
  void upload(const T* host, size_t n, cudaStream_t s = nullptr) {
    if (n > count) alloc(n);
    if (n) CA_CUDA(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

}  // namespace ca
