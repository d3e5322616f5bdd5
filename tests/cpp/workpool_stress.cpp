// Stress test of csrc/workpool.h (no GPU): several driver threads publish parallel-for jobs at once while others poll help_one() the way
// Batch::wait does; every item must run exactly once, a throwing item must resurface on the job's owner and leave the pool usable.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../xivo_b200/csrc/workpool.h"

int main() {
  setenv("XIVO_THREADS", "6", 1);
  setenv("XIVO_PIN", "0", 1);
  setenv("XIVO_SPIN_MS", "1", 1);
  xb::WorkPool& pool = xb::WorkPool::get();
  if (pool.workers() != 5) { printf("FAIL workers %d\n", pool.workers()); return 1; }
  std::atomic<bool> stop{false};
  std::atomic<long> helped{0}, errors{0}, items{0};
  std::vector<std::thread> th;
  for (int w = 0; w < 2; ++w)
    th.emplace_back([&] {
      while (!stop.load()) {
        if (pool.help_one()) helped.fetch_add(1);
        else xb::cpu_relax();
      }
    });
  for (int d = 0; d < 4; ++d)
    th.emplace_back([&, d] {
      std::mt19937 rng(1234 + d);
      for (int it = 0; it < 3000; ++it) {
        const int n = 2 + (int)(rng() % 96);
        std::vector<std::atomic<int>> hit(n);
        for (auto& h : hit) h.store(0);
        const bool thrower = (it % 97) == 5;
        const int bad = thrower ? (int)(rng() % n) : -1;
        bool caught = false;
        try {
          pool.pfor(n, [&](int i) {
            hit[i].fetch_add(1);
            volatile double x = 0;
            for (int k = 0; k < 50 + (i & 7) * 40; ++k) x += k * 0.5;
            if (i == bad) throw std::out_of_range("item");
          });
        } catch (const std::out_of_range&) {
          caught = true;
        }
        if (caught != thrower) errors.fetch_add(1);
        for (int i = 0; i < n; ++i)
          if (hit[i].load() != 1) errors.fetch_add(1);
        items.fetch_add(n);
      }
    });
  for (size_t i = 2; i < th.size(); ++i) th[i].join();
  stop.store(true);
  th[0].join();
  th[1].join();
  // the availability counter is back at zero: an idle poll does not find phantom work
  if (pool.help_one()) errors.fetch_add(1);
  printf("%s items %ld helped_by_pollers %ld errors %ld\n", errors.load() ? "FAIL" : "OK", items.load(), helped.load(), errors.load());
  return errors.load() ? 1 : 0;
}
