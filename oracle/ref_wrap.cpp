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

// ---- the read-back surface of pybind11/pyxivo.cpp:332-398, dumped for boundary-parity tests -------------------------------
// Per-feature table through the reference's accessors (estimator_accessors.cpp): n_output < 0 -> the no-argument overloads
// (instate_features_ order), otherwise the (int n_output) overloads (sorted by covariance norm, max(size, n_output) rows of
// which the first min(size, n_output) are written).  Returns the number of rows; columns: ids, sinds, ref group ids,
// Xs(3), Xc(3), xc(3), pred(2), meas(2), cov(6).
extern "C" int ref_feature_table(int n_output, int* ids, int* sinds, int* refs, double* Xs, double* Xc, double* xc, double* pred, double* meas,
                                 double* cov6, int max_rows) {
  const bool all = n_output < 0;
  const xivo::VecXi a = all ? g_est->InstateFeatureIDs() : g_est->InstateFeatureIDs(n_output);
  const xivo::VecXi s = all ? g_est->InstateFeatureSinds() : g_est->InstateFeatureSinds(n_output);
  const xivo::VecXi r = all ? g_est->InstateFeatureRefGroups() : g_est->InstateFeatureRefGroups(n_output);
  const xivo::MatX3 p = all ? g_est->InstateFeaturePositions() : g_est->InstateFeaturePositions(n_output);
  const xivo::MatX3 c = all ? g_est->InstateFeatureXc() : g_est->InstateFeatureXc(n_output);
  const xivo::MatX3 x = all ? g_est->InstateFeaturexc() : g_est->InstateFeaturexc(n_output);
  const xivo::MatX2 pr = all ? g_est->InstateFeaturePreds() : g_est->InstateFeaturePreds(n_output);
  const xivo::MatX2 me = all ? g_est->InstateFeatureMeas() : g_est->InstateFeatureMeas(n_output);
  const xivo::MatX6 cv = all ? g_est->InstateFeatureCovs() : g_est->InstateFeatureCovs(n_output);
  const int rows = (int)a.size();
  // rows = max(#in-state features of the graph, n_output); rows past min(#, n_output) are uninitialised in the reference (the caller
  // learns # from a call with n_output = 0 and ignores them)
  for (int i = 0; i < rows && i < max_rows; ++i) {
    ids[i] = a(i); sinds[i] = s(i); refs[i] = r(i);
    for (int k = 0; k < 3; ++k) { Xs[3 * i + k] = p(i, k); Xc[3 * i + k] = c(i, k); xc[3 * i + k] = x(i, k); }
    for (int k = 0; k < 2; ++k) { pred[2 * i + k] = pr(i, k); meas[2 * i + k] = me(i, k); }
    for (int k = 0; k < 6; ++k) cov6[6 * i + k] = cv(i, k);
  }
  return rows;
}
// InstateGroupIDs / Sinds / Poses (x y z w T) / Covs (as the reference fills them: estimator_accessors.cpp InstateGroupCovs)
extern "C" int ref_group_table(int* ids, int* sinds, double* pose7, double* cov21, int max_rows) {
  const xivo::VecXi a = g_est->InstateGroupIDs(), s = g_est->InstateGroupSinds();
  const xivo::MatX7 p = g_est->InstateGroupPoses();
  const xivo::MatX c = g_est->InstateGroupCovs();
  const int n = (int)a.size();
  for (int i = 0; i < n && i < max_rows; ++i) {
    ids[i] = a(i); sinds[i] = s(i);
    for (int k = 0; k < 7; ++k) pose7[7 * i + k] = p(i, k);
    for (int k = 0; k < 21; ++k) cov21[21 * i + k] = c(i, k);
  }
  return n;
}
// Ca(9) Cg(9) row-major, td, Rsg(9) row-major, camera intrinsics(9), distortion type, Pstate(81) row-major
extern "C" void ref_calibration(double* Ca, double* Cg, double* td, double* Rsg, double* intr9, int* dist_type, double* Pstate) {
  const xivo::Mat3 a = g_est->Ca(), g = g_est->Cg(), r = g_est->Rsg().matrix();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { Ca[3 * i + j] = a(i, j); Cg[3 * i + j] = g(i, j); Rsg[3 * i + j] = r(i, j); }
  *td = g_est->td();
  const auto in = xivo::CameraManager::instance()->GetIntrinsics();
  for (int i = 0; i < 9; ++i) intr9[i] = in(i);
  *dist_type = (int)xivo::CameraManager::instance()->GetDistortionType();
  const xivo::MatX ps = g_est->Pstate();
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) Pstate[9 * i + j] = ps(i, j);
}
extern "C" int ref_just_dropped(int* ids, int max_n) {
  const xivo::VecXi a = g_est->JustDroppedFeatureIDs();
  for (int i = 0; i < (int)a.size() && i < max_n; ++i) ids[i] = a(i);
  return (int)a.size();
}
// {num_tracker_outlier_rejected, num_tracker_failed_to_track, num_tracker_new_detections, num_oneptransac_rejected}
extern "C" void ref_tracker_counters(int* out) {
  out[0] = g_est->num_tracker_outlier_rejected(); out[1] = g_est->num_tracker_failed_to_track();
  out[2] = g_est->num_tracker_new_detections(); out[3] = g_est->num_oneptransac_rejected();
}
extern "C" void ref_scale_init_velocity(double s) { g_est->ScaleInitVelocity(s); }
// tracked_features_no_descriptor(): ids and last pixels of Tracker::features_
extern "C" int ref_tracked_features(int* ids, double* xy, int max_n) {
  const auto v = g_est->tracked_features_no_descriptor();
  int n = 0;
  for (const auto& t : v) {
    if (n < max_n) { ids[n] = std::get<0>(t); xy[2 * n] = std::get<1>(t)(0); xy[2 * n + 1] = std::get<1>(t)(1); }
    ++n;
  }
  return n;
}
