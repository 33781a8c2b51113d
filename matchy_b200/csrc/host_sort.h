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

// n workers — n - 1 threads and the caller itself, so that a pool as wide as the machine does not oversubscribe it —, each runs
// job(worker index) once per parallel() call; parallel() returns when all are done.  The phases of one
// sort follow each other within microseconds, so a worker spins on the generation counter for a short while after a job
// before it goes to sleep on the condition variable (a sleeping pool costs a futex wake per worker and phase).
class WorkerPool {
 public:
  explicit WorkerPool(unsigned n) {
    for (unsigned k = 0; k + 1 < n; k++) th_.emplace_back([this, k] { run(k); });
  }
  ~WorkerPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  unsigned size() const { return (unsigned)th_.size() + 1; }
  void parallel(const std::function<void(unsigned)>& f) {
    job_ = &f;
    pending_.store((unsigned)th_.size(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> l(m_);  // (pairs with the sleepers' predicate check: no lost wake-up)
      gen_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire)) cv_.notify_all();
    f((unsigned)th_.size());  // the caller is the last worker
    for (unsigned spin = 0; pending_.load(std::memory_order_acquire) != 0; spin++)
      if (spin > 2000) std::this_thread::yield();
  }

 private:
  void run(unsigned k) {
    unsigned seen = 0;
    for (;;) {
      bool have = false;
      for (unsigned spin = 0; spin < 20000 && !have; spin++) have = gen_.load(std::memory_order_acquire) != seen;  // ~20-50 us
      if (!have) {
        std::unique_lock<std::mutex> l(m_);
        sleepers_.fetch_add(1, std::memory_order_release);
        cv_.wait(l, [&] { return stop_ || gen_.load(std::memory_order_acquire) != seen; });
        sleepers_.fetch_sub(1, std::memory_order_release);
        if (stop_) return;
      }
      seen = gen_.load(std::memory_order_acquire);
      (*job_)(k);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  const std::function<void(unsigned)>* job_ = nullptr;
  std::atomic<unsigned> gen_{0}, pending_{0}, sleepers_{0};
  bool stop_ = false;
};

inline bool record_less(const mgpu_match& x, const mgpu_match& y) {
  if (x.offset != y.offset) return x.offset < y.offset;
  if (x.item_type != y.item_type) return x.item_type < y.item_type;
  return x.len < y.len;
}

// r[0..n) sorted by (offset, item_type, len).  Every offset lies in [lo, hi] (the scanned byte range).  tmp: scratch the caller
// keeps between calls; pool may be null (one thread).
inline void sort_records(mgpu_match* r, size_t n, uint64_t lo, uint64_t hi, std::vector<mgpu_match>& tmp, WorkerPool* pool) {
  if (n < 4096 || !pool || pool->size() < 2) { std::sort(r, r + n, record_less); return; }
  const unsigned T = pool->size();
  auto slice = [&](unsigned k) { return n * k / T; };  // (every worker takes part even for small n: the workers that find nothing
                                                        //  to do are awake for the next phase either way)
  // bucket = (offset - lo) >> shift, about 512 records per bucket when matches are spread evenly
  const uint64_t span = hi - lo;
  size_t want = std::min<size_t>(std::max<size_t>(n / 512, 16), (size_t)1 << 16);
  unsigned shift = 0;
  while ((span >> shift) >= want) shift++;
  const size_t B = (size_t)(span >> shift) + 1;
  std::vector<uint32_t> hist((size_t)T * B, 0);  // (n < 2^32: a batch holds at most cap_rec < 2^31 records per piece... sums below are 64-bit)
  std::atomic<bool> outside{false};
  pool->parallel([&](unsigned k) {
    uint32_t* h = hist.data() + (size_t)k * B;
    for (size_t i = slice(k); i < slice(k + 1); i++) {
      if (r[i].offset < lo || r[i].offset > hi) { outside.store(true, std::memory_order_relaxed); return; }
      h[(r[i].offset - lo) >> shift]++;
    }
  });
  if (outside.load()) { std::sort(r, r + n, record_less); return; }  // (a caller that got the range wrong still gets sorted records)
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

// Are r[0..n) in (offset, item_type, len) order already?  (The device sorts every piece's records before they leave it; this is
// the host's check, and what decides whether sort_records still has to run.)
inline bool records_sorted(const mgpu_match* r, size_t n, WorkerPool* pool) {
  if (n < 2) return true;
  if (n < 65536 || !pool || pool->size() < 2) { for (size_t i = 1; i < n; i++) if (record_less(r[i], r[i - 1])) return false; return true; }
  const unsigned T = pool->size();
  std::atomic<bool> bad{false};
  pool->parallel([&](unsigned k) {
    const size_t a = std::max<size_t>(1, n * k / T), b = n * (k + 1) / T;  // (slice k also compares its first element with the one before it)
    for (size_t i = a; i < b; i++) if (record_less(r[i], r[i - 1])) { bad.store(true, std::memory_order_relaxed); return; }
  });
  return !bad.load();
}

// The id pairs of pattern records, gathered in record order into `packed` (ids_index rewritten to match); the device appends
// them in whatever order its warps finish.  Returns the number of pairs.  Random reads of `ids`: latency-bound, so in parallel.
inline size_t repack_ids(mgpu_match* r, size_t n, const mgpu_id_pair* ids, std::vector<mgpu_id_pair>& packed, WorkerPool* pool) {
  const unsigned T = (pool && n >= 4096) ? pool->size() : 1;
  auto slice = [&](unsigned k) { return n * k / T; };
  std::vector<size_t> first(T + 1, 0);
  // Pattern records are a few per cent of a match-heavy result (config 5: 1 M of 41 M per 100 GB): the first pass notes where
  // they are, the second only visits those (one read of the records instead of two).
  std::vector<std::vector<size_t>> where(T);
  auto count = [&](unsigned k) {
    size_t c = 0;
    std::vector<size_t>& w = where[k];
    for (size_t i = slice(k); i < slice(k + 1); i++) if (r[i].kind == MGPU_KIND_PATTERN) { c += r[i].n_ids; w.push_back(i); }
    first[k + 1] = c;
  };
  auto gather = [&](unsigned k) {
    size_t at = first[k];
    for (size_t i : where[k]) {
      memcpy(packed.data() + at, ids + r[i].ids_index, (size_t)r[i].n_ids * sizeof(mgpu_id_pair));
      r[i].ids_index = (uint32_t)at;
      at += r[i].n_ids;
    }
  };
  if (T == 1) count(0); else pool->parallel(count);
  for (unsigned k = 0; k < T; k++) first[k + 1] += first[k];
  if (packed.size() < first[T]) packed.resize(first[T]);
  if (T == 1) gather(0); else pool->parallel(gather);
  return first[T];
}

}  // namespace mgpu
