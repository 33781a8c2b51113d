// Host side of record emission: the device appends match records with atomics, so their order varies from run to run; the
// C ABI promises them sorted by (offset, item_type, len).  Matches are spread over the log, so the sort is a one-level
// partition by offset range (monotone bucket function: concatenated sorted buckets are the sorted whole, equal offsets share a
// bucket) followed by small independent sorts — both steps on a pool of worker threads that lives as long as the context.
// A 10 GB step of config 2 ends with 53 K records: std::sort on one thread took 4-6 ms of a 12 ms step (bench.py value_wall).
// Plain C++ (no CUDA), so that the CPU test suite can compile and check it on its own.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/matchy_b200.h"

namespace mgpu {

// n workers, each runs job(worker index) once per parallel() call; parallel() returns when all are done.
class WorkerPool {
 public:
  explicit WorkerPool(unsigned n) {
    for (unsigned k = 0; k < n; k++) th_.emplace_back([this, k] { run(k); });
  }
  ~WorkerPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  unsigned size() const { return (unsigned)th_.size(); }
  void parallel(const std::function<void(unsigned)>& f) {
    std::unique_lock<std::mutex> l(m_);
    job_ = &f; pending_ = (unsigned)th_.size(); gen_++;
    cv_.notify_all();
    done_.wait(l, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void run(unsigned k) {
    unsigned seen = 0;
    std::unique_lock<std::mutex> l(m_);
    for (;;) {
      cv_.wait(l, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      const std::function<void(unsigned)>* j = job_;
      l.unlock();
      (*j)(k);
      l.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(unsigned)>* job_ = nullptr;
  unsigned gen_ = 0, pending_ = 0;
  bool stop_ = false;
};

inline bool record_less(const mgpu_match& x, const mgpu_match& y) {
  if (x.offset != y.offset) return x.offset < y.offset;
  if (x.item_type != y.item_type) return x.item_type < y.item_type;
  return x.len < y.len;
}

// r[0..n) sorted by (offset, item_type, len).  tmp: scratch the caller keeps between calls; pool may be null (one thread).
inline void sort_records(mgpu_match* r, size_t n, std::vector<mgpu_match>& tmp, WorkerPool* pool) {
  if (n < 4096 || !pool || pool->size() < 2) { std::sort(r, r + n, record_less); return; }
  const unsigned T = pool->size();
  // offset range (one slice per worker)
  std::vector<uint64_t> lo_s(T, ~0ull), hi_s(T, 0);
  auto slice = [&](unsigned k) { return n * k / T; };
  pool->parallel([&](unsigned k) {
    uint64_t lo = ~0ull, hi = 0;
    for (size_t i = slice(k); i < slice(k + 1); i++) { lo = std::min(lo, r[i].offset); hi = std::max(hi, r[i].offset); }
    lo_s[k] = lo; hi_s[k] = hi;
  });
  const uint64_t lo = *std::min_element(lo_s.begin(), lo_s.end()), hi = *std::max_element(hi_s.begin(), hi_s.end());
  // bucket = (offset - lo) >> shift, about 512 records per bucket when matches are spread evenly
  const uint64_t span = hi - lo;
  size_t want = std::min<size_t>(std::max<size_t>(n / 512, 16), (size_t)1 << 16);
  unsigned shift = 0;
  while ((span >> shift) >= want) shift++;
  const size_t B = (size_t)(span >> shift) + 1;
  std::vector<uint32_t> hist((size_t)T * B, 0);  // (n < 2^32: a batch holds at most cap_rec < 2^31 records per piece... sums below are 64-bit)
  pool->parallel([&](unsigned k) {
    uint32_t* h = hist.data() + (size_t)k * B;
    for (size_t i = slice(k); i < slice(k + 1); i++) h[(r[i].offset - lo) >> shift]++;
  });
  std::vector<size_t> start((size_t)T * B), bucket_at(B + 1);
  size_t run = 0;
  for (size_t b = 0; b < B; b++) {
    bucket_at[b] = run;
    for (unsigned k = 0; k < T; k++) { start[(size_t)k * B + b] = run; run += hist[(size_t)k * B + b]; }
  }
  bucket_at[B] = run;
  if (tmp.size() < n) tmp.resize(n);
  mgpu_match* t = tmp.data();
  pool->parallel([&](unsigned k) {
    size_t* s = start.data() + (size_t)k * B;
    for (size_t i = slice(k); i < slice(k + 1); i++) t[s[(r[i].offset - lo) >> shift]++] = r[i];
  });
  // sort the buckets (handed out in groups: neighbouring buckets are neighbours in memory) and copy them back in place
  std::atomic<size_t> next{0};
  const size_t group = std::max<size_t>(1, B / (T * 8));
  pool->parallel([&](unsigned) {
    for (size_t g = next.fetch_add(group); g < B; g = next.fetch_add(group)) {
      const size_t g1 = std::min(B, g + group);
      for (size_t b = g; b < g1; b++) std::sort(t + bucket_at[b], t + bucket_at[b + 1], record_less);
      memcpy(r + bucket_at[g], t + bucket_at[g], (bucket_at[g1] - bucket_at[g]) * sizeof(mgpu_match));
    }
  });
}

}  // namespace mgpu
