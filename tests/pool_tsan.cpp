// ThreadSanitizer harness for the context's host worker pool (csrc/host_pool.h); built and run by
// tests/test_abi.py::test_host_pool_under_tsan.  Exercises the patterns the engine uses: run() with the
// caller taking share 0, start() + caller-side work + wait(), dynamic hand-out through a counter, back-to-back
// jobs of different widths (workers still spinning or already parked), and destruction while idle.
#include <atomic>
#include <cstdio>
#include <numeric>
#include <vector>

#include "host_pool.h"

using frc_host::Pool;

int main() {
  long long checks = 0;
  for (int workers : {0, 1, 3, 7}) {
    Pool pool(workers);
    std::vector<long long> acc(pool.size(), 0);
    for (int round = 0; round < 300; ++round) {
      const int T = 1 + round % pool.size();
      // run(): static shares
      std::vector<int> hit(T, 0);
      pool.run(T, [&](int t) { hit[t]++; acc[t] += t; });
      for (int t = 0; t < T; ++t) if (hit[t] != 1) { printf("run: share %d ran %d times\n", t, hit[t]); return 1; }
      // start() + wait(): one task list handed out through a counter, the caller joins in
      const int n_tasks = 5 + round % 40;
      std::vector<int> done(n_tasks, 0);
      std::atomic<int> next{0};
      const std::function<void(int)> work = [&](int t) {
        for (int k = next.fetch_add(1); k < n_tasks; k = next.fetch_add(1)) { done[k]++; acc[t] += k; }
      };
      const int W = pool.size() - 1;
      if (W > 0) pool.start(W, work);
      if (round % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(round % 7 * 100));  // parked vs spinning workers
      work(W);
      if (W > 0) pool.wait();
      for (int k = 0; k < n_tasks; ++k) if (done[k] != 1) { printf("start: task %d ran %d times\n", k, done[k]); return 1; }
      checks += T + n_tasks;
    }
  }
  printf("ok %lld\n", checks);
  return 0;
}
