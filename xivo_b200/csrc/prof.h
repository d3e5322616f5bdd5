// Optional per-kernel timing with CUDA events on the launching stream, and H2D/D2H byte counters.
// bench.py turns it on for the timed region: roofline.achieved must come from event durations
// measured live there, not from a profiler run.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace xb {
struct ProfEntry {
  unsigned long long calls = 0;
  double ms = 0;
  double work = 0;  // algorithmic bytes or flops attributed by the host at the launch site
};
class Prof {
 public:
  static Prof& get() {
    static Prof p;
    return p;
  }
  std::atomic<bool> enabled{false};
  std::atomic<unsigned long long> h2d{0}, d2h{0};
  // record start; returns an index to pass to stop()
  int start(const char* name, cudaStream_t st) {
    if (!enabled.load(std::memory_order_relaxed)) return -1;
    std::lock_guard<std::mutex> lk(mu_);
    Rec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    pending_.push_back(r);
    return (int)pending_.size() - 1;
  }
  void stop(int idx, cudaStream_t st) {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(mu_);
    cudaEventRecord(pending_[idx].b, st);
  }
  // call after the stream has been synchronised
  void collect() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& r : pending_) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        auto& e = acc_[r.name];
        e.calls++;
        e.ms += ms;
      }
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    pending_.clear();
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
    collect();
    std::lock_guard<std::mutex> lk(mu_);
    acc_.clear();
    h2d = 0;
    d2h = 0;
  }
  std::string json() {
    collect();
    std::lock_guard<std::mutex> lk(mu_);
    std::string s = "{";
    bool first = true;
    for (auto& kv : acc_) {
      char buf[256];
      snprintf(buf, sizeof(buf), "%s\"%s\": {\"calls\": %llu, \"ms\": %.6f, \"work\": %.6e}", first ? "" : ", ", kv.first.c_str(), kv.second.calls, kv.second.ms,
               kv.second.work);
      s += buf;
      first = false;
    }
    char tail[128];
    snprintf(tail, sizeof(tail), "%s\"_h2d_bytes\": %llu, \"_d2h_bytes\": %llu}", first ? "" : ", ", h2d.load(), d2h.load());
    return s + tail;
  }

 private:
  struct Rec {
    std::string name;
    cudaEvent_t a, b;
  };
  std::mutex mu_;
  std::vector<Rec> pending_;
  std::map<std::string, ProfEntry> acc_;
};

struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t s) : idx(Prof::get().start(name, s)), st(s) {}
  ~ProfScope() { Prof::get().stop(idx, st); }
};
}  // namespace xb
