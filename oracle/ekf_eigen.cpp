// oracle/ekf_eigen.cpp — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU baseline for the EKF side of the hot path: the reference's own Eigen expression sequence
// for the measurement update (Estimator::UpdateJosephForm, /root/reference/src/estimator.cpp:1257-1288)
// and the Mahalanobis gate (Estimator::MHGating, /root/reference/src/update.cpp:60-69), compiled
// against the reference's VENDORED Eigen 3.3.9 (/root/reference/thirdparty/eigen, read-only include
// path; nothing is copied).  estimator.cpp itself cannot be compiled here (it needs OpenCV C++
// headers), so this is a port ("kind": "port") of those ~30 lines on the reference's
// linear-algebra library, single-threaded like the reference (no OpenMP/BLAS for Eigen).
// Built by oracle/build_ref.py into oracle/_ref/libekf_eigen.so (git-ignored, travels to the GPU box).
#include <chrono>

#include "Eigen/Dense"

using MatX = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic>;  // col-major, like common/alias.h
using VecX = Eigen::Matrix<double, Eigen::Dynamic, 1>;
using RowMat = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;

extern "C" {

// One update; inputs/outputs row-major.  Returns seconds spent in the update expressions.
double ref_update_joseph(int N, int M, const double* H_rm, double* P_rm, const double* inn_, const double* diagR_, double* err_out) {
  MatX H_ = Eigen::Map<const RowMat>(H_rm, M, N);
  MatX P_ = Eigen::Map<const RowMat>(P_rm, N, N);
  VecX inn = Eigen::Map<const VecX>(inn_, M), diagR = Eigen::Map<const VecX>(diagR_, M), err(N);
  MatX S_, K_, I_KH_;
  auto t0 = std::chrono::high_resolution_clock::now();
  S_ = H_ * P_ * H_.transpose();
  for (int i = 0; i < diagR.size(); ++i) S_(i, i) += diagR(i);
  K_.setZero(N, H_.rows());
  K_.transpose() = S_.ldlt().solve(H_ * P_);
  err = K_ * inn;
  I_KH_ = K_ * H_;
  for (int i = 0; i < N; ++i) I_KH_(i, i) -= 1;
  P_ = I_KH_ * P_ * I_KH_.transpose();
  for (int i = 0; i < K_.cols(); ++i) K_.block(0, i, K_.rows(), 1) *= std::sqrt(diagR(i));
  P_.noalias() += K_ * K_.transpose();
  auto t1 = std::chrono::high_resolution_clock::now();
  Eigen::Map<RowMat>(P_rm, N, N) = P_;
  Eigen::Map<VecX>(err_out, N) = err;
  return std::chrono::duration<double>(t1 - t0).count();
}

// n gates with dense 2 x N Jacobians (row-major n x 2 x N), as the reference computes them.
double ref_mh_gating(int N, int n, const double* J_rm, const double* P_rm, const double* inn2, double R, double* dist) {
  MatX P_ = Eigen::Map<const RowMat>(P_rm, N, N);
  auto t0 = std::chrono::high_resolution_clock::now();
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix<double, 2, Eigen::Dynamic> J = Eigen::Map<const Eigen::Matrix<double, 2, Eigen::Dynamic, Eigen::RowMajor>>(J_rm + (size_t)i * 2 * N, 2, N);
    Eigen::Vector2d res(inn2[2 * i], inn2[2 * i + 1]);
    Eigen::Matrix2d S = J * P_ * J.transpose();
    S(0, 0) += R;
    S(1, 1) += R;
    dist[i] = res.dot(S.llt().solve(res));
  }
  auto t1 = std::chrono::high_resolution_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
}
