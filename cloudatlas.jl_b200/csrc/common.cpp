#include "common.h"
#include <vector>
#include <atomic>

#include <algorithm>
#include <cstring>
#include <thread>
#include <ctime>

namespace ca {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};

void set_last_error(const std::string& msg) { t_last_error = msg; }

void fail(int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
PhaseTimer::PhaseTimer(const char* w) : what(w), t0(now_s()), on(getenv("CA_TIMING") != nullptr) {}
void PhaseTimer::lap(const char* phase) {
  if (!on) return;
  const double t = now_s();
  fprintf(stderr, "[ca timing] %s: %s %.3f s\n", what, phase, t - t0);
  t0 = t;
}

unsigned worker_threads() {
  if (const char* e = getenv("CA_NUM_THREADS")) {
    const int v = atoi(e);
    if (v >= 1) return (unsigned)std::min(v, 64);
  }
  return std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
}

void parallel_chunks(size_t n, size_t chunks, const std::function<void(size_t, size_t, size_t)>& fn) {
  if (n == 0) return;
  chunks = std::max<size_t>(1, std::min(chunks, n));
  const unsigned nthreads = (unsigned)std::min<size_t>(worker_threads(), chunks);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (size_t c = next.fetch_add(1); c < chunks; c = next.fetch_add(1)) fn(n * c / chunks, n * (c + 1) / chunks, c);
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

namespace {
bool use_pool() {
  static const bool on = getenv("CA_NO_POOL") == nullptr;
  return on;
}
// raise the release threshold of the current device's default pool once per device
void prepare_pool(int dev) {
  static std::atomic<unsigned long long> done{0};
  if (dev < 64 && (done.load() >> dev) & 1ull) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;  // never hand freed blocks back to the driver on synchronisation
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (dev < 64) done.fetch_or(1ull << dev);
}
}  // namespace

void* device_alloc(size_t bytes) {
  void* p = nullptr;
  if (!use_pool()) {
    CA_CUDA(cudaMalloc(&p, bytes));
    return p;
  }
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  prepare_pool(dev);
  cudaError_t e = cudaMallocAsync(&p, bytes, nullptr);
  if (e == cudaErrorMemoryAllocation) {  // the pool holds blocks that do not fit: give them back and retry once
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      cudaDeviceSynchronize();
      cudaMemPoolTrimTo(pool, 0);
    }
    e = cudaMallocAsync(&p, bytes, nullptr);
  }
  if (e != cudaSuccess) fail(e == cudaErrorMemoryAllocation ? CA_ERR_NOMEM : CA_ERR_CUDA, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  CA_CUDA(cudaStreamSynchronize(nullptr));  // the block is usable from every stream from here on
  return p;
}

void device_free(void* ptr) {
  if (!ptr) return;
  if (!use_pool()) {
    cudaFree(ptr);
    return;
  }
  cudaDeviceSynchronize();  // cudaFree's guarantee: nothing on any stream still touches the block
  if (cudaFreeAsync(ptr, nullptr) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(ptr);
  }
}

void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    fail(CA_ERR_NODEVICE, "no CUDA device available (%s); libcloudatlas_b200 has no CPU fallback",
         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  int major = 0;
  CA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    fail(CA_ERR_NODEVICE, "device %d has compute capability %d.x; this library is built for sm_100a only",
         dev, major);
}

}  // namespace ca

extern "C" {

const char* ca_last_error(void) { return ca::t_last_error.c_str(); }

int32_t ca_version(void) { return 100; }

int64_t ca_launch_count(void) { return ca::g_launches.load(); }

int32_t ca_device_info(int32_t* ndevices, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                       char* name, int32_t name_len) {
  return ca::guarded([&] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) n = 0;
    if (ndevices) *ndevices = n;
    if (n == 0) ca::fail(CA_ERR_NODEVICE, "no CUDA device available");
    int dev = 0;
    CA_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CA_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (name && name_len > 0) {
      strncpy(name, p.name, name_len - 1);
      name[name_len - 1] = 0;
    }
  });
}

int32_t ca_set_device(int32_t device) {
  return ca::guarded([&] {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) ca::fail(CA_ERR_NODEVICE, "no CUDA device available");
    if (device < 0 || device >= n) ca::fail(CA_ERR_INVALID, "device %d out of range 0:%d", device, n - 1);
    CA_CUDA(cudaSetDevice(device));
  });
}

int32_t ca_get_device(int32_t* device) {
  return ca::guarded([&] {
    if (!device) ca::fail(CA_ERR_INVALID, "null argument");
    int d = 0;
    CA_CUDA(cudaGetDevice(&d));
    *device = d;
  });
}

int32_t ca_pin(void* host_ptr, int64_t bytes) {
  return ca::guarded([&] {
    if (!host_ptr || bytes <= 0) ca::fail(CA_ERR_INVALID, "ca_pin: null pointer or non-positive size");
    CA_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterDefault));
  });
}

int32_t ca_unpin(void* host_ptr) {
  return ca::guarded([&] {
    if (!host_ptr) ca::fail(CA_ERR_INVALID, "ca_unpin: null pointer");
    CA_CUDA(cudaHostUnregister(host_ptr));
  });
}

}  // extern "C"
