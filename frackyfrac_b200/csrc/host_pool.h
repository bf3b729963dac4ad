// Host worker pool of a context (plain C++, no CUDA): used by frc_create (table staging, tree tasks) and by
// frc_next (widening of fp32 bands).  Kept in a header of its own so that tests/test_abi.py can build it into a
// ThreadSanitizer harness (tests/pool_tsan.cpp) without a GPU.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace frc_host {

// Small persistent worker pool (host-side validation / staging copies): waking
// parked threads costs microseconds, spawning them per job cost ~0.5 ms.
class Pool {
 public:
  explicit Pool(int n) {
    for (int t = 0; t < n; ++t) workers_.emplace_back([this, t] { loop(t + 1); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; epoch_.fetch_add(1, std::memory_order_release); }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return static_cast<int>(workers_.size()) + 1; }
  // Runs fn(t) for t in [0, T) on T <= size() threads (t = 0 on the caller).
  void run(int T, const std::function<void(int)>& fn) {
    if (T <= 1) { fn(0); return; }
    post(&fn, /*shift=*/0, /*active=*/T, /*pending=*/T - 1);
    fn(0);
    wait();
  }
  // Starts fn(t) for t in [0, T) on the worker threads alone and returns; wait() joins them.
  // T <= size() - 1; a pool without workers runs the shares inline.
  void start(int T, const std::function<void(int)>& fn) {
    if (workers_.empty()) { for (int t = 0; t < T; ++t) fn(t); return; }
    post(&fn, /*shift=*/1, /*active=*/T + 1, /*pending=*/T);
  }
  void wait() {
    // the shares are tens of microseconds long: poll before paying a futex round trip
    spin_while([this] { return pending_.load(std::memory_order_acquire) != 0; }, 300);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_.load(std::memory_order_acquire) == 0; });
    fn_ = nullptr;
  }

 private:
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  template <class Pred>
  static void spin_while(Pred busy, int max_us) {
    const auto t0 = std::chrono::steady_clock::now();
    while (busy()) {
      for (int k = 0; k < 32 && busy(); ++k) cpu_relax();
      if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(max_us)) return;
    }
  }
  void post(const std::function<void(int)>* fn, int shift, int active, int pending) {
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = fn; shift_ = shift; active_ = active;
      pending_.store(pending, std::memory_order_release);
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      int arg = 0;
      // a job's calls arrive ~0.1 ms apart (one per output band): stay hot that long before parking
      spin_while([&] { return epoch_.load(std::memory_order_acquire) == seen; }, 200);
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
        seen = epoch_.load(std::memory_order_acquire);
        if (stop_) return;
        if (id < active_) { fn = fn_; arg = id - shift_; }
      }
      if (fn) {
        (*fn)(arg);
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
          std::lock_guard<std::mutex> g(m_);
          done_.notify_one();
        }
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int active_ = 0, shift_ = 0;
  std::atomic<int> pending_{0};
  std::atomic<uint64_t> epoch_{0};
  bool stop_ = false;
};

}  // namespace frc_host
