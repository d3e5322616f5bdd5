// C entry points over the REFERENCE's own estimator (src/estimator.cpp, update.cpp, manager.cpp, feature.cpp, ... compiled
// unmodified from /root/reference by oracle/build_ref.py with header shims for OpenCV / glog): the point-cloud path
// Estimator::InertialMeas + VisualMeasPointCloud (the reference's simulation entry, scripts/pyxivo_pcw.py) and the read-back
// accessors.  Used (a) to generate golden trajectories that pin oracle/estimator_oracle.py and the CUDA pipeline on the
// real reference and (b) as the "reference" kind of CPU baseline for the EKF half.  One estimator per process (the
// reference's singletons).  Test infrastructure, built into oracle/_ref/ only.
#include <chrono>
#include <cstring>
#include <sstream>
#include <string>

#include "estimator.h"
#include "json/json.h"

namespace xivo {
EstimatorPtr CreateSystem(const Json::Value& cfg);
}

static xivo::EstimatorPtr g_est = nullptr;
static std::string g_msg;

extern "C" {
const char* ref_error() { return g_msg.c_str(); }
int ref_state_dim() { return xivo::kFullSize; }
int ref_max_groups() { return xivo::kMaxGroup; }
int ref_max_features() { return xivo::kMaxFeature; }

int ref_create(const char* cfg_text) {
  try {
    Json::Value cfg;
    Json::CharReaderBuilder b;
    b["collectComments"] = false;
    std::istringstream is(cfg_text);
    std::string errs;
    if (!Json::parseFromStream(b, is, &cfg, &errs)) { g_msg = errs; return -1; }
    g_est = xivo::CreateSystem(cfg);
    return g_est ? 0 : -1;
  } catch (const std::exception& e) {
    g_msg = e.what();
    return -1;
  }
}
void ref_init_with_sim_depths() { g_est->InitWithSimDepths(); }

int ref_inertial(unsigned long long ts, const double* gyro, const double* accel) {
  try {
    g_est->InertialMeas(xivo::timestamp_t(ts), xivo::Vec3(gyro[0], gyro[1], gyro[2]), xivo::Vec3(accel[0], accel[1], accel[2]));
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}
// xp_depth: n x 3 row-major (x, y, depth)
int ref_visual_pointcloud(unsigned long long ts, int n, const int* ids, const double* xp_depth) {
  try {
    xivo::VecXi vi(n);
    xivo::MatX3 m(n, 3);
    for (int i = 0; i < n; ++i) {
      vi(i) = ids[i];
      for (int k = 0; k < 3; ++k) m(i, k) = xp_depth[3 * i + k];
    }
    g_est->VisualMeasPointCloud(xivo::timestamp_t(ts), vi, m);
    return 0;
  } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}
void ref_gsb(double* out12) {  // 3x4 row-major [R | T]
  const auto g = g_est->gsb();
  const auto M = g.matrix3x4();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) out12[4 * i + j] = M(i, j);
}
void ref_gbc(double* out12) {
  const auto M = g_est->gbc().matrix3x4();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) out12[4 * i + j] = M(i, j);
}
void ref_motion(double* Vsb, double* bg, double* ba) {
  const auto v = g_est->Vsb(), g = g_est->bg(), a = g_est->ba();
  for (int i = 0; i < 3; ++i) { Vsb[i] = v(i); bg[i] = g(i); ba[i] = a(i); }
}
void ref_P(double* out) {  // N x N row-major
  const xivo::MatX P = g_est->P();
  const int N = (int)P.rows();
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) out[(size_t)i * N + j] = P(i, j);
}
unsigned long long ref_time_ns() { return (unsigned long long)g_est->ts().count(); }
// {num_instate_features, num_instate_groups, gauge_group, num_mh_rejected, MeasurementUpdateInitialized, VisionInitialized}
void ref_counters(int* out) {
  out[0] = g_est->num_instate_features(); out[1] = g_est->num_instate_groups(); out[2] = g_est->gauge_group(); out[3] = g_est->num_mh_rejected();
  out[4] = g_est->MeasurementUpdateInitialized() ? 1 : 0; out[5] = g_est->VisionInitialized() ? 1 : 0;
}
int ref_instate_feature_ids(int* ids, int* sinds, int max_n) {
  const xivo::VecXi a = g_est->InstateFeatureIDs(), s = g_est->InstateFeatureSinds();
  const int n = (int)a.size();
  for (int i = 0; i < n && i < max_n; ++i) { ids[i] = a(i); sinds[i] = s(i); }
  return n;
}
// every live track of Tracker::features_: id, FeatureStatus, TrackStatus, lifetime, local state x (3), diag of its 3x3 covariance
int ref_tracks(int* ids, int* status, int* tstatus, int* lifetime, double* x3, int max_n) {
  int n = 0;
  for (auto f : xivo::Tracker::instance()->features_) {
    if (n < max_n) {
      ids[n] = f->id(); status[n] = (int)f->status(); tstatus[n] = (int)f->track_status(); lifetime[n] = f->lifetime();
      const auto x = f->x();
      for (int k = 0; k < 3; ++k) x3[3 * n + k] = x(k);
    }
    ++n;
  }
  return n;
}
int ref_instate_group_ids(int* ids, int* sinds, int max_n) {
  const xivo::VecXi a = g_est->InstateGroupIDs(), s = g_est->InstateGroupSinds();
  const int n = (int)a.size();
  for (int i = 0; i < n && i < max_n; ++i) { ids[i] = a(i); sinds[i] = s(i); }
  return n;
}
}
