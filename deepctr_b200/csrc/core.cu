// core.cu — ABI version, thread-local error string, launch counter.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace b2ctr {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long* g_oob[64] = {nullptr};
unsigned long long* oob_counter() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_oob[dev]) {
    // cudaMalloc is illegal while the calling thread captures a graph: the counter is created by the first
    // (eager) launch of a model, long before its step is captured
    unsigned long long* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, sizeof(unsigned long long));
    g_oob[dev] = p;
  }
  return g_oob[dev];
}
}  // namespace b2ctr

extern "C" {
int32_t b2ctr_abi_version(void) { return 1; }
b2ctr_status_t b2ctr_embed_oob_count(int64_t* count, int32_t reset, void* stream) {
  B2_REQUIRE(count, "embed_oob_count: NULL count");
  unsigned long long* d = b2ctr::oob_counter();
  *count = 0;
  if (!d) return B2CTR_OK;
  unsigned long long h = 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && reset && true) e = cudaMemsetAsync(d, 0, sizeof(h), st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    b2ctr::set_error("embed_oob_count: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B2CTR_ERR_CUDA;
  }
  *count = (int64_t)h;
  return B2CTR_OK;
}
const char* b2ctr_last_error(void) { return b2ctr::g_err; }
int64_t b2ctr_launch_count(void) { return (int64_t)b2ctr::g_launches.load(); }
void b2ctr_reset_launch_count(void) { b2ctr::g_launches.store(0); }
b2ctr_status_t b2ctr_set_l2_fetch_granularity(int32_t bytes) {
  // cudaLimitMaxL2FetchGranularity is a hint (32 / 64 / 128): with random 4-byte and 128-byte row reads the
  // default 64 B granule doubles the DRAM traffic of every dim-1 (linear-term) lookup
  cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
  if (e != cudaSuccess) {
    b2ctr::set_error("set_l2_fetch_granularity(%d): %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return B2CTR_ERR_CUDA;
  }
  return B2CTR_OK;
}
b2ctr_status_t b2ctr_l2_persist_reserve(int64_t bytes, int64_t* granted, int64_t* max_window) {
  int dev = 0, max_persist = 0, max_win = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  size_t want = bytes < 0 ? 0 : (size_t)bytes;
  if (want > (size_t)max_persist) want = (size_t)max_persist;
  if (e == cudaSuccess) e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  size_t got = 0;
  if (e == cudaSuccess) e = cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
  if (e != cudaSuccess) {
    b2ctr::set_error("l2_persist_reserve(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return B2CTR_ERR_CUDA;
  }
  if (granted) *granted = (int64_t)got;
  if (max_window) *max_window = (int64_t)max_win;
  return B2CTR_OK;
}
b2ctr_status_t b2ctr_enable_peer_access(int32_t peer_device) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return B2CTR_OK;
  }
  if (e != cudaSuccess) {
    b2ctr::set_error("enable_peer_access(%d): %s", peer_device, cudaGetErrorString(e));
    cudaGetLastError();
    return B2CTR_ERR_CUDA;
  }
  return B2CTR_OK;
}
}
