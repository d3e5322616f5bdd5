// CPU micro-benchmark of the host state machine (xivo_b200/csrc/estimator_host.cpp.inc): one estimator on a synthetic point-cloud
// stream, every device phase replaced by a trivial stand-in (sub-filter = identity, all Mahalanobis distances inside the gate, zero
// correction), so that only the host bookkeeping is timed.  Authoring aid (scripts/host_perf.sh); not a parity check — the decisions
// themselves are pinned by tests/test_host_twin.py.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>

#include "../../xivo_b200/csrc/estimator.h"
#include "../../xivo_b200/csrc/homography.h"

namespace xb {
static std::map<std::string, std::pair<double, long>> g_t;
struct HostScope {
  const char* n;
  std::chrono::steady_clock::time_point t0;
  explicit HostScope(const char* name) : n(name), t0(std::chrono::steady_clock::now()) {}
  ~HostScope() {
    auto& e = g_t[n];
    e.first += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    e.second++;
  }
};
}  // namespace xb
#include "../../xivo_b200/csrc/estimator_host.cpp.inc"

using namespace xb;
struct Seq {
  std::unique_ptr<Estimator> e;
  std::vector<double> pack, x0, y0;
  std::vector<char> tracked;
  uint64_t ts = 0;
  int frames = 0;
};

int main(int argc, char** argv) {
  std::ifstream f(argv[1]);
  std::stringstream ss;
  ss << f.rdbuf();
  const int G = argc > 2 ? atoi(argv[2]) : 4, F = argc > 3 ? atoi(argv[3]) : 14, NF = argc > 4 ? atoi(argv[4]) : 400;
  const int NE = argc > 5 ? atoi(argv[5]) : 1;  // estimators stepped round-robin: > ~64 makes every phase start cache-cold, like a bench batch
  const Json cfg = Json::parse(ss.str());
  const int NP = 300, W = 640, H = 480;
  std::vector<Seq> seqs(NE);
  for (int q = 0; q < NE; ++q) {
    Seq& s = seqs[q];
    s.e.reset(new Estimator(cfg, EkfLayout{G, F}, false));
    s.e->sim_initialize_depths = true;
    const int N = s.e->lay.N();
    s.pack.assign(2 * N + 529, 0.0);
    for (int i = 0; i < N; ++i) s.pack[N + 529 + i] = 1e-3 * (1 + i % 7);
    for (int i = 0; i < 23; ++i) s.pack[N + i * 23 + i] = 1e-4;
    s.x0.resize(NP); s.y0.resize(NP);
    unsigned rng = 12345 + 77 * q;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return (rng >> 8) / double(1 << 24); };
    for (int i = 0; i < NP; ++i) { s.x0[i] = rnd() * (W + 300); s.y0[i] = 20 + rnd() * (H - 40); }
    s.tracked.assign(NP * 64, 0);
  }
  double gyro[3] = {0, 0, 0}, accel[3] = {0, 0, 9.8};
  auto t_begin = std::chrono::steady_clock::now();
  long frames = 0;
  auto run_msg = [&](Seq& s, Msg& m) {
    Estimator& e = *s.e;
    const int N = e.lay.N();
    if (m.type == 0) {
      HostScope hs("1 inertial (8 per frame)");
      e.inertial_internal(m.ts, m.gyro, m.accel);
      if (e.needs_state_now()) e.stages.clear();
      return;
    }
    { HostScope hs("2 visual_begin (propagate)"); if (!e.visual_begin(m.ts, m.type)) return; }
    e.stages.clear();
    { HostScope hs("3 predict"); e.predict_features(); }
    for (size_t k = 0; k < m.ids.size(); ++k) e.ids_to_depths.insert({m.ids[k], m.xp_depth[3 * k + 2]});
    { HostScope hs("4 tracker_update_pointcloud"); e.tracker_update_pointcloud(m.ids, m.xp_depth); }
    { HostScope hs("5 update_step_pre"); e.update_step_pre(); }
    std::vector<SubfilterOut> so(e.subfilter_list.size());
    for (size_t i = 0; i < so.size(); ++i) {
      Feature* ft = e.subfilter_list[i];
      memcpy(so[i].x, ft->x, 24); memcpy(so[i].P, ft->P, 72);
      for (int k = 0; k < 9; ++k) so[i].P[k] *= 0.8;
      so[i].outlier_counter = 0;
    }
    { HostScope hs("6 after_subfilter (select/add)"); e.update_step_after_subfilter(so.data()); }
    e.edits.clear();
    std::vector<double> mh(e.lay.F, 1.0);
    { HostScope hs("7 after_gate"); e.update_step_after_gate(mh.data()); }
    e.edits.clear();
    { HostScope hs("8 after_update (absorb, manage)"); e.update_step_after_update(s.pack.data(), s.pack.data() + N, s.pack.data() + N + 529, !e.in_update.empty()); }
    e.edits.clear();
    ++frames;
    ++s.frames;
    if (&s == &seqs[0] && s.frames % 100 == 0) printf("  frame %d: tracks %zu instate %zu groups %zu\n", s.frames, e.tracks.size(), e.instate_features.size(), e.graph.groups.size());
    if (e.error) { printf("error %d %s\n", e.error, e.error_msg.c_str()); exit(1); }
  };
  for (int fr = 0; fr < NF; ++fr) {
    // like a lock-step batch: every sequence runs its IMU samples, then every sequence its frame
    for (Seq& s : seqs)
      for (int k = 0; k < 8; ++k) {
        Msg m; m.ts = s.ts; m.type = 0; memcpy(m.gyro, gyro, 24); memcpy(m.accel, accel, 24);
        s.ts += 5000000;
        s.e->push(std::move(m));
        Msg o;
        if (s.e->pop_ready(&o)) run_msg(s, o);
      }
    for (Seq& s : seqs) {
      Msg v; v.ts = (uint64_t)fr * 40000000ull; v.type = 3;
      // emulate the image tracker's policy: tracks leave one by one, new ones arrive in a burst when fewer than 120 remain
      auto vis = [&](int i, int* id, double* x) {
        const double xx = s.x0[i] + 1.5 * fr;
        *id = i + 1000 * (int)(xx / (W + 300.0));
        *x = fmod(xx, W + 300.0) - 150.0;
        return *x > 10 && *x < W - 10;
      };
      int ntr = 0;
      for (int i = 0; i < NP; ++i) { int id; double x; if (vis(i, &id, &x) && s.tracked[id % (NP * 64)] == 1) ++ntr; }
      const bool burst = ntr < 120;
      for (int i = 0; i < NP; ++i) {
        int id; double x;
        if (!vis(i, &id, &x)) continue;
        char& t = s.tracked[id % (NP * 64)];
        if (t != 1 && burst && ntr < 150) { t = 1; ++ntr; }
        if (t == 1) { v.ids.push_back(id); v.xp_depth.push_back(x); v.xp_depth.push_back(s.y0[i]); v.xp_depth.push_back(2.0); }
      }
      s.e->push(std::move(v));
      Msg o;
      if (s.e->pop_ready(&o)) run_msg(s, o);
    }
  }
  const double total = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
  Estimator& e = *seqs[0].e;
  printf("estimators %d  frames %ld  tracks %zu  instate %zu  groups %zu  features in graph %zu   total %.1f us/frame\n", NE, frames, e.tracks.size(), e.instate_features.size(),
         e.graph.groups.size(), e.graph.features.size(), total / frames);
  for (auto& kv : g_t) printf("  %-34s %8.2f us/frame  (%ld calls)\n", kv.first.c_str(), kv.second.first / frames, kv.second.second);
  return 0;
}
