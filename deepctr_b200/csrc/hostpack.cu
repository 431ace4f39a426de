// hostpack.cu — host-side input staging: copy the per-feature arrays of one batch into the flat pinned
// buffer that is uploaded with a single cudaMemcpyAsync (deepctr_b200/inputs.py Feeder).
// The reference hands its numpy inputs to Keras' data adapter (deepctr examples: model.fit(model_input, y));
// here the equivalent step must keep up with a ~1.3 ms training step, i.e. move ~10 MB per batch at well
// above the ~6 GB/s one core manages, without holding the Python GIL: a small persistent thread pool.
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <string.h>
#include "common.cuh"

namespace b2ctr {
namespace {

struct Job {
  const void* const* src;
  const int64_t* nbytes;
  const int64_t* dst_off;
  unsigned char* dst;
  int n;
};

class CopyPool {
 public:
  explicit CopyPool(int threads) : stop_(false), epoch_(0), pending_(0) {
    for (int i = 0; i < threads; ++i) workers_.emplace_back([this, i, threads] { loop(i, threads); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  void run(const Job& job) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = job;
    pending_ = (int)workers_.size();
    ++epoch_;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  // worker i copies a contiguous share of the BYTES (blocks are split where needed), so that a batch made
  // of a few wide inputs is spread as evenly as one made of many narrow ones
  static void copy_share(const Job& j, int i, int n_workers) {
    int64_t total = 0;
    for (int b = 0; b < j.n; ++b) total += j.nbytes[b];
    const int64_t per = ((total + n_workers - 1) / n_workers + 63) & ~(int64_t)63;
    const int64_t lo = per * i, hi = lo + per < total ? lo + per : total;
    int64_t pos = 0;
    for (int b = 0; b < j.n && pos < hi; ++b) {
      const int64_t beg = pos, end = pos + j.nbytes[b];
      pos = end;
      const int64_t s = beg > lo ? beg : lo, e = end < hi ? end : hi;
      if (s < e)
        memcpy(j.dst + j.dst_off[b] + (s - beg), (const unsigned char*)j.src[b] + (s - beg), (size_t)(e - s));
    }
  }
  void loop(int i, int n_workers) {
    uint64_t seen = 0;
    for (;;) {
      Job job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        job = job_;
      }
      copy_share(job, i, n_workers);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  Job job_;
  bool stop_;
  uint64_t epoch_;
  int pending_;
};

std::mutex g_pool_mu;
CopyPool* g_pool = nullptr;   // intentionally leaked at exit (threads are parked on a condition variable)

}  // namespace
}  // namespace b2ctr

using namespace b2ctr;

extern "C" b2ctr_status_t b2ctr_host_pack(const void* const* src, const int64_t* nbytes, const int64_t* dst_off,
                                          int32_t n, void* dst, int32_t threads) {
  B2_REQUIRE(n >= 0 && (n == 0 || (src && nbytes && dst_off && dst)), "host_pack: NULL argument");
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    B2_REQUIRE(nbytes[i] >= 0 && dst_off[i] >= 0 && (nbytes[i] == 0 || src[i]), "host_pack: bad block %d", i);
    total += nbytes[i];
  }
  if (total == 0) return B2CTR_OK;
  if (threads == 1 || total < (1 << 20)) {          // small batches: not worth waking the pool
    for (int i = 0; i < n; ++i) memcpy((unsigned char*)dst + dst_off[i], src[i], (size_t)nbytes[i]);
    return B2CTR_OK;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);          // one pack at a time (the staging thread is single)
  if (!g_pool) {
    unsigned hw = std::thread::hardware_concurrency();
    int want = threads > 0 ? threads : (int)(hw >= 32 ? 8 : hw >= 8 ? 4 : 2);
    g_pool = new CopyPool(want);
  }
  Job job{src, nbytes, dst_off, (unsigned char*)dst, n};
  g_pool->run(job);
  return B2CTR_OK;
}
