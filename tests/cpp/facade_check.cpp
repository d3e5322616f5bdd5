// Compile/link check of include/xivo_b200.hpp: instantiates every member of the facade so that a signature
// drift between the header and the C ABI fails at build time.  Run without a GPU it must report the
// "no device" error (there is no CPU fallback); with a GPU it runs a two-frame point-cloud smoke.
#include <cstdio>
#include <cstring>

#include "xivo_b200.hpp"

int main(int argc, char** argv) {
  const std::string cfg = argc > 1 ? argv[1] : "";
  try {
    auto est = xivo::Estimator::CreateFromFile(cfg, 4, 14);
    est->InitWithSimDepths();
    for (int k = 0; k < 30; ++k) {
      est->InertialMeas(xivo::timestamp_t(k * 5000000LL), {0, 0, 0}, {0, 0, 9.8});
      if (k % 8 == 0) est->VisualMeasPointCloud(xivo::timestamp_t(k * 5000000LL), {1, 2, 3}, {100, 100, 2, 200, 120, 2.5, 300, 200, 3});
    }
    auto g = est->gsb();
    auto P = est->P();
    std::printf("ok N=%d tracked=%zu T=(%g %g %g) P00=%g inst=%d groups=%d ts=%lld\n", est->state_dim(), est->tracked_features_no_descriptor().size(),
                g[3], g[7], g[11], P[0], est->num_instate_features(), est->num_instate_groups(), (long long)est->ts().count());
    (void)est->gbc(); (void)est->gsc(); (void)est->Pstate(); (void)est->Vsb(); (void)est->bg(); (void)est->ba(); (void)est->Rsg();
    (void)est->MeasurementUpdateInitialized(); (void)est->VisionInitialized(); (void)est->gauge_group(); (void)est->num_mh_rejected();
    (void)est->num_tracker_failed_to_track(); (void)est->num_tracker_new_detections();
    (void)est->InstateFeatureIDs(); (void)est->InstateFeatureSinds(); (void)est->InstateFeatureRefGroups(); (void)est->InstateFeaturePositions();
    (void)est->InstateGroupIDs(); (void)est->InstateGroupSinds(); (void)est->InstateGroupPoses();
    // the rest of the read-back surface (estimator_accessors.cpp, pybind11/pyxivo.cpp:332-398), both overloads: always compiled (a
    // signature drift fails the build), executed when asked for (tests/test_gpu_widen_1_readback.py)
    if (argc > 2 && !std::strcmp(argv[2], "--readback")) {
      const size_t count = est->InstateFeatureIDs(0).size();  // (int n_output) overloads have max(count, n_output) rows
      if (est->InstateFeatureIDs(count + 3).size() != count + 3 || est->InstateFeaturePositions(2).size() != 3 * std::max<size_t>(count, 2)) return 3;
      (void)est->InstateFeatureSinds(4); (void)est->InstateFeatureRefGroups(4); (void)est->InstateFeatureXc(); (void)est->InstateFeatureXc(4);
      (void)est->InstateFeaturexc(); (void)est->InstateFeaturexc(4); (void)est->InstateFeaturePreds(); (void)est->InstateFeaturePreds(4);
      (void)est->InstateFeatureMeas(); (void)est->InstateFeatureMeas(4); (void)est->InstateFeatureCovs(); (void)est->InstateFeatureCovs(4);
      if (est->InstateGroupPoses().size() != 7 * est->InstateGroupIDs().size() || est->InstateGroupCovs().size() != 21 * est->InstateGroupIDs().size()) return 4;
      (void)est->InstateGroupCovBlocks(); (void)est->JustDroppedFeatureIDs(); (void)est->td(); (void)est->Ca(); (void)est->Cg();
      const auto intr = est->CameraIntrinsics();
      if (intr[0] != 275.0 || intr[2] != 320.0 || est->CameraDistortionType() != 0) return 5;  // cfg/pcw_sim.json camera
      (void)est->num_tracker_outlier_rejected(); (void)est->num_oneptransac_rejected(); (void)est->UsingLoopClosure(); est->CloseLoop();
      const auto v0 = est->Vsb();
      est->ScaleInitVelocity(2.0);
      if (est->Vsb()[0] != v0[0] / 2.0) return 6;
      // stateful Tracker surface (src/tracker.h:25-54) on a tracker-only session
      auto st = xivo::Tracker::Create(std::string("{\"simulation\": false, \"camera_cfg\": {\"model\": \"pinhole\", \"rows\": 64, \"cols\": 64}, "
                                                  "\"tracker_cfg\": {\"num_features_min\": 5, \"num_features_max\": 10, \"margin\": 4, \"KLT\": {\"max_level\": 2}, "
                                                  "\"FAST\": {\"threshold\": 20}}}"));
      // random 4x4 blocks + per-pixel dither: plenty of FAST-9 corners (cv2 finds 151 on this very texture, 95 inside the margin), no score ties
      std::vector<uint8_t> tex(64 * 64);
      unsigned seed = 12345;
      auto lcg = [&seed]() { seed = (seed * 1103515245u + 12345u) & 0x7fffffffu; return (seed >> 16); };
      for (int by = 0; by < 16; ++by)
        for (int bx = 0; bx < 16; ++bx) {
          const int v = (int)(lcg() & 0xff);
          for (int y = 0; y < 4; ++y)
            for (int x = 0; x < 4; ++x) tex[64 * (4 * by + y) + 4 * bx + x] = (uint8_t)v;
        }
      for (int i = 0; i < 64 * 64; ++i) {
        const int v = (int)tex[i] + (int)(lcg() % 7) - 3;
        tex[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
      xivo::ImageView tv{tex.data(), 64, 64, 1};
      st->Update(tv);
      st->Update(tv);
      if (st->features().empty() || st->num_rejected_outliers() != 0) return 7;
      (void)st->num_new_detections(); (void)st->num_failed_to_track();
      try { xivo::Tracker kernel_only; kernel_only.Update(tv); return 8; } catch (const xivo::Error& e) { if (e.code != XIVO_ERR_STATE) return 9; }
      std::printf("readback ok\n");
    }
    std::vector<uint8_t> img(64 * 64, 0);
    xivo::ImageView v{img.data(), 64, 64, 1};
    xivo::Tracker trk;
    auto kp = trk.Detect(v, 20);
    auto fl = trk.TrackLK(v, v, {32.f, 32.f});
    std::printf("tracker ok kp=%d status=%d\n", kp.total, (int)fl.status[0]);
    try {
      est->VisualMeas(xivo::timestamp_t(1000000000LL), v);  // simulation mode: must throw like the reference (estimator.cpp:1112-1115)
      est->VisualMeasTrackerOnly(xivo::timestamp_t(1000000000LL), v);
      est->VisualMeasPointCloudTrackerOnly(xivo::timestamp_t(1000000000LL), {1}, {1, 2, 3});
    } catch (const xivo::Error& e) {
      std::printf("expected error: %d\n", e.code);
    }
    return 0;
  } catch (const xivo::Error& e) {
    std::printf("xivo::Error %d: %s\n", e.code, e.what());
    return e.code == XIVO_ERR_CUDA ? 42 : 1;
  }
}
