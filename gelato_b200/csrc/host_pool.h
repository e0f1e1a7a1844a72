// Host side of update mode: the packed x-dependent Jacobian slots of a batch, scattered into the caller's
// Jacobian buffers by a process-wide pool of worker threads that lives across calls (spawning threads per
// call cost more than the scatter itself: profiles/r01v_e2e.txt).  Plain C++, no CUDA; also compiled into the
// host emulator so the CPU test tier covers it.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

namespace gelato_host {

// A job is n_tasks calls fn(ctx, task); run() returns when all of them are done.  The calling thread works too.
class HostPool {
 public:
  typedef void (*TaskFn)(void* ctx, int task);

  // Never destroyed: the workers are detached and end with the process (no join at exit, nothing to
  // tear down in a forked child, where the caller simply does all the work itself).
  static HostPool& instance() {
    static HostPool* pool = new HostPool;
    return *pool;
  }

  void run(int n_tasks, int threads, TaskFn fn, void* ctx) {
    if (n_tasks <= 0) return;
    if (threads <= 1 || n_tasks == 1) {
      for (int t = 0; t < n_tasks; t++) fn(ctx, t);
      return;
    }
    std::lock_guard<std::mutex> one_job(run_mu_);
    grow(threads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = fn;
      ctx_ = ctx;
      n_tasks_ = n_tasks;
      next_.store(0, std::memory_order_relaxed);
      done_ = 0;
      wanted_ = threads - 1;  // workers beyond this stay asleep-equivalent: they find the job full
      joined_ = 0;
      gen_++;
    }
    cv_work_.notify_all();
    int mine = 0;
    for (int t; (t = next_.fetch_add(1, std::memory_order_relaxed)) < n_tasks;) {
      fn(ctx, t);
      mine++;
    }
    std::unique_lock<std::mutex> lk(mu_);
    done_ += mine;
    cv_done_.wait(lk, [&] { return done_ == n_tasks_ && inside_ == 0; });
    fn_ = nullptr;
  }

  int workers() const { return n_workers_; }

 private:
  HostPool() {}
  HostPool(const HostPool&);
  void operator=(const HostPool&);

  void grow(int n) {
    for (; n_workers_ < n; n_workers_++) std::thread(&HostPool::worker, this).detach();
  }

  void worker() {
    unsigned long long seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_work_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (!fn_ || joined_ >= wanted_) continue;  // the job is over already, or has the threads it asked for
      joined_++;
      inside_++;
      const TaskFn fn = fn_;
      void* const ctx = ctx_;
      const int n = n_tasks_;
      lk.unlock();
      int mine = 0;
      for (int t; (t = next_.fetch_add(1, std::memory_order_relaxed)) < n;) {
        fn(ctx, t);
        mine++;
      }
      lk.lock();
      done_ += mine;
      inside_--;
      if (done_ == n_tasks_ && inside_ == 0) cv_done_.notify_one();
    }
  }

  std::mutex run_mu_;  // one job at a time
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  int n_workers_ = 0;
  TaskFn fn_ = nullptr;
  void* ctx_ = nullptr;
  int n_tasks_ = 0, done_ = 0, inside_ = 0, wanted_ = 0, joined_ = 0;
  std::atomic<int> next_{0};
  unsigned long long gen_ = 0;
};

// vals[s][idx[i]] = packed[s][i] for scenarios [s0, s1)
inline void scatter_scenarios(const int64_t* idx, long long n_idx, const double* packed, double* vals, long long n_vals,
                              int s0, int s1) {
  for (int s = s0; s < s1; s++) {
    const double* src = packed + (size_t)s * n_idx;
    double* dst = vals + (size_t)s * n_vals;
    // isolated slots (the node-diagonals of the dense D blocks) cost one cache-line fill each: ask for
    // the lines a few dozen writes ahead so the fills overlap (non-temporal stores measured 5x slower)
    const long long ahead = 48;
    long long i = 0;
    for (; i + ahead < n_idx; i++) {
      __builtin_prefetch(dst + idx[i + ahead], 1, 0);
      dst[idx[i]] = src[i];
    }
    for (; i < n_idx; i++) dst[idx[i]] = src[i];
  }
}

struct ScatterJob {
  const int64_t* idx;
  long long n_idx;
  const double* packed;
  double* vals;
  long long n_vals;
  int s0, s1, parts;
};

inline void scatter_task(void* ctx, int task) {
  const ScatterJob& j = *static_cast<const ScatterJob*>(ctx);
  const long long n = j.s1 - j.s0;
  const int a = j.s0 + (int)(n * task / j.parts), b = j.s0 + (int)(n * (task + 1) / j.parts);
  if (a < b) scatter_scenarios(j.idx, j.n_idx, j.packed, j.vals, j.n_vals, a, b);
}

// packed -> vals for scenarios [s0, s1) on up to `threads` threads of the pool (the caller's included)
inline void scatter_parallel(const int64_t* idx, long long n_idx, const double* packed, double* vals, long long n_vals,
                             int s0, int s1, int threads) {
  if (s1 <= s0) return;
  if (threads > s1 - s0) threads = s1 - s0;
  ScatterJob job = {idx, n_idx, packed, vals, n_vals, s0, s1, threads < 1 ? 1 : threads};
  HostPool::instance().run(job.parts, threads, scatter_task, &job);
}

}  // namespace gelato_host
