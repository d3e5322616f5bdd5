// Optional per-kernel timing with CUDA events on the launching stream, host phase timers and
// H2D/D2H byte counters.  bench.py turns it on for the timed region: roofline.achieved must come from
// event durations measured live there, not from a profiler run.  Safe with several batches driving
// different streams from different host threads.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace xb {
struct ProfEntry {
  unsigned long long calls = 0;
  double ms = 0;
  double work = 0;  // algorithmic bytes or flops attributed by the host at the launch site
};
struct ProfRec {
  std::string name;
  cudaEvent_t a = nullptr, b = nullptr;
  bool armed = false;
};
// Lock-free accumulators for the per-sequence host scopes (many threads, microsecond regions).
struct FineTable {
  struct Slot {
    std::atomic<const char*> name{nullptr};
    std::atomic<unsigned long long> ns{0}, calls{0};
    char pad[40];
  } slot[64];
  void add(const char* n, unsigned long long ns) {
    for (auto& e : slot) {
      const char* cur = e.name.load(std::memory_order_relaxed);
      if (!cur) {
        const char* expect = nullptr;
        if (e.name.compare_exchange_strong(expect, n)) cur = n;
        else cur = expect;
      }
      if (cur == n) {
        e.ns.fetch_add(ns, std::memory_order_relaxed);
        e.calls.fetch_add(1, std::memory_order_relaxed);
        return;
      }
    }
  }
  void reset() {
    for (auto& e : slot) { e.ns = 0; e.calls = 0; }
  }
};
class Prof {
 public:
  FineTable fine_tab;
  static Prof& get() {
    static Prof p;
    return p;
  }
  std::atomic<bool> enabled{false};
  std::atomic<bool> fine{false};  // per-sequence host scopes ("x_*"): contended, only for host-side diagnosis
  std::atomic<bool> kernels{true};  // false: host scopes only (no event pair per launch: eight runtime calls each, which distorts the host picture)
  std::atomic<unsigned long long> h2d{0}, d2h{0};
  ProfRec* start(const char* name, cudaStream_t st) {
    if (!enabled.load(std::memory_order_relaxed) || !kernels.load(std::memory_order_relaxed)) return nullptr;
    std::unique_ptr<ProfRec> r(new ProfRec());
    r->name = name;
    cudaEventCreate(&r->a);
    cudaEventCreate(&r->b);
    cudaEventRecord(r->a, st);
    ProfRec* p = r.get();
    std::lock_guard<std::mutex> lk(mu_);
    pending_.push_back(std::move(r));
    return p;
  }
  void stop(ProfRec* r, cudaStream_t st) {
    if (!r) return;
    cudaEventRecord(r->b, st);
    std::lock_guard<std::mutex> lk(mu_);
    r->armed = true;
  }
  // harvest every finished record (records of other streams still in flight stay pending)
  void collect(bool wait = false) {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<std::unique_ptr<ProfRec>> keep;
    for (auto& r : pending_) {
      if (!r->armed) {
        keep.push_back(std::move(r));
        continue;
      }
      if (wait) cudaEventSynchronize(r->b);
      if (cudaEventQuery(r->b) != cudaSuccess) {
        keep.push_back(std::move(r));
        continue;
      }
      float ms = 0;
      if (cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) {
        auto& e = acc_[r->name];
        e.calls++;
        e.ms += ms;
      }
      cudaEventDestroy(r->a);
      cudaEventDestroy(r->b);
    }
    pending_.swap(keep);
  }
  void add_host(const char* name, double ms) {
    std::lock_guard<std::mutex> lk(mu_);
    auto& e = acc_[std::string("host:") + name];
    e.calls++;
    e.ms += ms;
  }
  void add_work(const char* name, double w) {
    if (!enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(mu_);
    acc_[name].work += w;
  }
  void reset() {
    collect(true);
    std::lock_guard<std::mutex> lk(mu_);
    acc_.clear();
    fine_tab.reset();
    h2d = 0;
    d2h = 0;
  }
  std::string json() {
    collect(true);
    std::lock_guard<std::mutex> lk(mu_);
    std::string s = "{";
    bool first = true;
    for (auto& kv : acc_) {
      char buf[256];
      snprintf(buf, sizeof(buf), "%s\"%s\": {\"calls\": %llu, \"ms\": %.6f, \"work\": %.6e}", first ? "" : ", ", kv.first.c_str(), kv.second.calls,
               kv.second.ms, kv.second.work);
      s += buf;
      first = false;
    }
    for (auto& e : fine_tab.slot) {
      const char* n = e.name.load();
      if (!n || !e.calls.load()) continue;
      char buf[256];
      snprintf(buf, sizeof(buf), "%s\"host:%s\": {\"calls\": %llu, \"ms\": %.6f, \"work\": 0}", first ? "" : ", ", n, e.calls.load(), e.ns.load() * 1e-6);
      s += buf;
      first = false;
    }
    char tail[128];
    snprintf(tail, sizeof(tail), "%s\"_h2d_bytes\": %llu, \"_d2h_bytes\": %llu}", first ? "" : ", ", h2d.load(), d2h.load());
    return s + tail;
  }

 private:
  std::mutex mu_;
  std::vector<std::unique_ptr<ProfRec>> pending_;
  std::map<std::string, ProfEntry> acc_;
};

struct ProfScope {
  ProfRec* rec;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t s) : rec(Prof::get().start(name, s)), st(s) {}
  ~ProfScope() { Prof::get().stop(rec, st); }
};
}  // namespace xb
