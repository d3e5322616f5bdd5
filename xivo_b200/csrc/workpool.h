// Process-wide host worker pool for the per-sequence estimator logic.
//
// Every Batch runs its per-sequence host phases (track bookkeeping, graph / slot management, gating
// decisions) as parallel-for jobs on this one pool, so several batches driven from several caller
// threads share a fixed set of workers instead of each spawning a team: the host phases of one batch
// fill the GPU waits of the others without oversubscribing a CPU quota.  CUDA calls stay on the
// caller ("driver") thread of each batch; the driver also executes items of its own job.
//
// Sizing: XIVO_THREADS workers if set, else (cgroup CPU quota or hardware threads, divided by
// LOCAL_WORLD_SIZE when launched by torchrun) minus one for the driver, capped at 32.
// Placement (XIVO_PIN, default 2): workers and drivers are pinned to distinct physical cores (all SMT
// siblings of a core) taken from this rank's contiguous share of the allowed cores (LOCAL_RANK of
// LOCAL_WORLD_SIZE), idlest first as measured from /proc/stat -- the GPU boxes are shared hosts and a
// core already loaded by another tenant halves the speed of whatever phase lands on it.
#pragma once
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <mutex>
#include <algorithm>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace xb {

inline void cpu_relax() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

class WorkPool {
 public:
  static WorkPool& get() {
    static WorkPool p;
    return p;
  }
  int workers() const { return (int)threads_.size(); }

  // run fn(i) for i in [0, n) on the pool and the calling thread; returns when all items are done
  template <typename Fn>
  void pfor(int n, Fn&& fn) {
    if (n <= 0) return;
    if (threads_.empty() || n == 1) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    Job job;
    job.n = n;
    job.ctx = &fn;
    job.call = [](void* c, int i) { (*static_cast<typename std::remove_reference<Fn>::type*>(c))(i); };
    int slot = -1;
    for (;;) {  // publish
      lock();
      for (int s = 0; s < kSlots; ++s)
        if (!slots_[s]) { slots_[s] = &job; slot = s; break; }
      unlock();
      if (slot >= 0) break;
      cpu_relax();
    }
    avail_.fetch_add(1, std::memory_order_release);
    epoch_.fetch_add(1);
    if (sleepers_.load() > 0) {
      std::lock_guard<std::mutex> lk(mu_);
      cv_.notify_all();
    }
    run(job);
    while (job.done.load(std::memory_order_acquire) < n) cpu_relax();
    lock();
    slots_[slot] = nullptr;
    unlock();
    while (job.refs.load(std::memory_order_acquire) > 0) cpu_relax();
    // an item that threw (FlatMap::at, bad_alloc ...) was counted as done on whatever thread ran it; the first exception resurfaces here,
    // on the thread that owns the job, once no worker refers to the job any more
    if (job.failed.load(std::memory_order_acquire)) std::rethrow_exception(job.err);
  }

 private:
  static constexpr int kSlots = 16;
  struct Job {
    std::atomic<int> next{0}, done{0}, refs{0};
    int n = 0;
    void (*call)(void*, int) = nullptr;
    void* ctx = nullptr;
    std::atomic<bool> failed{false};
    std::exception_ptr err;
  };
  // published jobs that still have unclaimed items: drivers polling for the GPU call help_one() in a tight loop, and a look at this
  // counter keeps them off the slot lock (and its cache line) while there is nothing to take
  std::atomic<int> avail_{0};
  Job* slots_[kSlots] = {nullptr};
  std::atomic_flag lk_ = ATOMIC_FLAG_INIT;
  std::atomic<unsigned long long> epoch_{0};
  std::atomic<int> sleepers_{0};
  std::atomic<bool> stop_{false};
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<std::thread> threads_;

  void lock() {
    while (lk_.test_and_set(std::memory_order_acquire)) cpu_relax();
  }
  void unlock() { lk_.clear(std::memory_order_release); }

  void run_item(Job& j, int i) {
    if (i == j.n - 1) avail_.fetch_sub(1, std::memory_order_relaxed);  // whoever claims the last item retires the job from the counter
    try {
      j.call(j.ctx, i);
    } catch (...) {
      if (!j.failed.exchange(true, std::memory_order_acq_rel)) j.err = std::current_exception();
    }
    j.done.fetch_add(1, std::memory_order_release);
  }
  void run(Job& j) {
    for (;;) {
      const int i = j.next.fetch_add(1, std::memory_order_relaxed);
      if (i >= j.n) break;
      run_item(j, i);
    }
  }
  Job* grab() {  // a published job that still has unclaimed items, with a reference held
    if (avail_.load(std::memory_order_acquire) <= 0) return nullptr;
    Job* r = nullptr;
    lock();
    for (int s = 0; s < kSlots; ++s) {
      Job* j = slots_[s];
      if (j && j->next.load(std::memory_order_relaxed) < j->n) {
        j->refs.fetch_add(1, std::memory_order_relaxed);
        r = j;
        break;
      }
    }
    unlock();
    return r;
  }
  void worker(int k, std::vector<int> cpus) {
    if (!cpus.empty()) {
      cpu_set_t set;
      CPU_ZERO(&set);
      for (int c : cpus) CPU_SET(c, &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    (void)k;
    const auto spin_for = std::chrono::milliseconds(spin_ms_);
    auto idle_since = std::chrono::steady_clock::now();
    int polls = 0;
    while (!stop_.load(std::memory_order_relaxed)) {
      // read the publish counter before looking for work, so a job published in between is not missed
      const unsigned long long seen = epoch_.load(std::memory_order_acquire);
      if (Job* j = grab()) {
        run(*j);
        j->refs.fetch_sub(1, std::memory_order_release);
        idle_since = std::chrono::steady_clock::now();
        polls = 0;
        continue;
      }
      // nothing to do: spin on the (read-shared) publish counter, then sleep
      while (epoch_.load(std::memory_order_acquire) == seen && !stop_.load(std::memory_order_relaxed)) {
        cpu_relax();
        if ((++polls & 1023) == 0 && std::chrono::steady_clock::now() - idle_since > spin_for) {
          std::unique_lock<std::mutex> lk(mu_);
          sleepers_.fetch_add(1);
          cv_.wait(lk, [&] { return stop_.load() || epoch_.load() != seen; });
          sleepers_.fetch_sub(1);
          idle_since = std::chrono::steady_clock::now();
        }
      }
    }
  }

  int spin_ms_ = 20;

  static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
  }
  // CPUs this process may use: cgroup quota (v2 cpu.max, v1 cfs_quota_us) and the affinity mask
  static int cpu_budget() {
    int hw = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = CPU_COUNT(&set);
    if (hw <= 0) hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 1;
    double quota = 0;
    {
      std::ifstream f("/sys/fs/cgroup/cpu.max");
      std::string q;
      double period = 0;
      if (f && (f >> q >> period) && q != "max" && period > 0) quota = atof(q.c_str()) / period;
    }
    if (quota <= 0) {
      std::ifstream fq("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"), fp("/sys/fs/cgroup/cpu/cpu.cfs_period_us");
      double q = 0, p = 0;
      if (fq && fp && (fq >> q) && (fp >> p) && q > 0 && p > 0) quota = q / p;
    }
    int n = hw;
    if (quota > 0 && quota < n) n = (int)quota;
    return n < 1 ? 1 : n;
  }
  // physical cores (lists of SMT siblings) among the CPUs of the affinity mask, in CPU order
  static std::vector<std::vector<int>> cores() {
    std::vector<std::vector<int>> out;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return out;
    std::set<int> seen;
    for (int c = 0; c < CPU_SETSIZE; ++c) {
      if (!CPU_ISSET(c, &set) || seen.count(c)) continue;
      std::vector<int> sib;
      std::ifstream f("/sys/devices/system/cpu/cpu" + std::to_string(c) + "/topology/thread_siblings_list");
      std::string line;
      if (f && std::getline(f, line)) {
        std::stringstream ss(line);
        std::string tok;
        while (std::getline(ss, tok, ',')) {
          const size_t dash = tok.find('-');
          const int a = atoi(tok.c_str()), b = dash == std::string::npos ? a : atoi(tok.c_str() + dash + 1);
          for (int x = a; x <= b; ++x)
            if (x >= 0 && x < CPU_SETSIZE && CPU_ISSET(x, &set)) sib.push_back(x);
        }
      }
      if (sib.empty()) sib.push_back(c);
      for (int x : sib) seen.insert(x);
      out.push_back(sib);
    }
    return out;
  }

  // busy fraction of every logical CPU over a short window (/proc/stat shows the host's CPUs, including
  // the load of other tenants of a shared box); empty when unreadable
  static std::vector<double> cpu_load(int window_ms) {
    auto snap = [](std::vector<std::pair<double, double>>& v) {
      std::ifstream f("/proc/stat");
      std::string line;
      while (f && std::getline(f, line)) {
        if (line.compare(0, 3, "cpu") != 0 || line.size() < 4 || line[3] < '0' || line[3] > '9') continue;
        std::stringstream ss(line.substr(3));
        int id;
        double x, tot = 0, idle = 0;
        ss >> id;
        for (int k = 0; k < 8 && (ss >> x); ++k) {
          tot += x;
          if (k == 3 || k == 4) idle += x;
        }
        if (id >= 0 && id < CPU_SETSIZE) {
          if ((int)v.size() <= id) v.resize(id + 1, {0.0, 0.0});
          v[id] = {tot, idle};
        }
      }
    };
    std::vector<std::pair<double, double>> a, b;
    snap(a);
    std::this_thread::sleep_for(std::chrono::milliseconds(window_ms));
    snap(b);
    std::vector<double> load;
    if (a.empty() || a.size() != b.size()) return load;
    load.resize(a.size(), 0.0);
    for (size_t i = 0; i < a.size(); ++i) {
      const double dt = b[i].first - a[i].first, di = b[i].second - a[i].second;
      load[i] = dt > 0 ? 1.0 - di / dt : 0.0;
    }
    return load;
  }

  std::vector<std::vector<int>> driver_cpus_;
  std::atomic<int> next_driver_{0};

 public:
  static int budget() { return cpu_budget(); }  // CPUs this process may use (no pool is created by asking)
  // Run at most one item of any published job on the calling thread (a driver that is waiting for the GPU
  // lends its CPU to the other batches).  Returns false when there was nothing to do.
  bool help_one() {
    if (threads_.empty()) return false;
    Job* j = grab();
    if (!j) return false;
    const int i = j->next.fetch_add(1, std::memory_order_relaxed);
    if (i < j->n) run_item(*j, i);
    j->refs.fetch_sub(1, std::memory_order_release);
    return true;
  }

  // Pin the calling (driver) thread to one of the cores reserved for drivers; once per thread.  Changing the affinity of a thread
  // that belongs to the host application is a lasting side effect, so it is opt-in: XIVO_PIN_DRIVERS=1 (bench.py sets it).
  void pin_driver() {
    static thread_local bool done = false;
    static const bool enabled = env_int("XIVO_PIN_DRIVERS", 0) != 0;
    if (!enabled || done || driver_cpus_.empty()) return;
    done = true;
    const std::vector<int>& cpus = driver_cpus_[next_driver_.fetch_add(1) % driver_cpus_.size()];
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }

 private:
  WorkPool() {
    const int local_world = std::max(1, env_int("LOCAL_WORLD_SIZE", 1));
    const int local_rank = std::max(0, env_int("LOCAL_RANK", 0)) % local_world;
    const int share = std::max(1, cpu_budget() / local_world);
    int n = env_int("XIVO_THREADS", 0);
    if (n <= 0) n = std::min(32, share) - 1;  // one CPU is the driver's
    else n -= 1;                              // XIVO_THREADS counts the driver
    if (n < 0) n = 0;
    spin_ms_ = env_int("XIVO_SPIN_MS", 20);
    const int n_drivers = std::max(1, env_int("XIVO_DRIVERS", 1));
    // XIVO_PIN: 0 = leave placement to the OS, 1 = physical cores in CPU order, 2 (default) = the idlest
    // physical cores of this rank's share of the machine, measured over 100 ms
    const int pin = env_int("XIVO_PIN", 2);
    std::vector<std::vector<int>> cs;
    if (pin) cs = cores();
    if (!cs.empty()) {
      // this rank's contiguous share of the cores (ranks of one node must not overlap).  XIVO_CPU_SLICE="i/n" overrides the torchrun
      // numbering: an application that has already narrowed the affinity mask to the NUMA node of its GPU says which of the n ranks of
      // THAT node it is (bench.py does)
      int si = local_rank, sn = local_world;
      if (const char* sl = getenv("XIVO_CPU_SLICE")) {
        int a = 0, b = 0;
        if (sscanf(sl, "%d/%d", &a, &b) == 2 && b > 0 && a >= 0 && a < b) { si = a; sn = b; }
      }
      const size_t lo = cs.size() * (size_t)si / sn, hi = cs.size() * (size_t)(si + 1) / sn;
      std::vector<std::vector<int>> mine(cs.begin() + lo, cs.begin() + std::max(hi, lo + 1));
      if (pin >= 2) {
        const std::vector<double> load = cpu_load(100);
        if (!load.empty()) {
          auto core_load = [&](const std::vector<int>& c) {
            double l = 0;
            for (int x : c) l += x < (int)load.size() ? load[x] : 0.0;
            return l;
          };
          std::stable_sort(mine.begin(), mine.end(), [&](const std::vector<int>& a, const std::vector<int>& b) { return core_load(a) < core_load(b); });
        }
      }
      cs.swap(mine);
      for (int d = 0; d < n_drivers; ++d) driver_cpus_.push_back(cs[d % cs.size()]);
    }
    for (int k = 0; k < n; ++k) {
      std::vector<int> cpus;
      if (!cs.empty()) cpus = cs[(size_t)(n_drivers + k) % cs.size()];
      threads_.emplace_back([this, k, cpus] { worker(k, cpus); });
    }
  }
  ~WorkPool() {
    stop_ = true;
    {
      std::lock_guard<std::mutex> lk(mu_);
      cv_.notify_all();
    }
    for (auto& t : threads_) t.join();
  }
};

inline int host_cpu_budget() { return WorkPool::budget(); }

}  // namespace xb
