// Tracker-level outlier rejection: the inlier mask of cv::findHomography(pts0, pts1, LMEDS | RANSAC, ...) as
// Tracker::OutlierRejection calls it (/root/reference/src/tracker.cpp:705-753, options `outlier_rejection`
// tracker.cpp:118-131).  Host side, float points, double model: ~55 four-point hypotheses per frame for LMedS.
// The arithmetic is OpenCV's (un-vendored): calib3d fundam.cpp (HomographyEstimatorCallback, findHomography),
// ptsetreg.cpp (RANSAC / LMedS registrators, RANSACUpdateNumIters), levmarq.cpp (LMSolver), core's cv::RNG —
// restated in oracle/homography_oracle.py and pinned there on cv2 4.13 (identical masks).  The mask is the one
// OpenCV >= 4.x returns: inliers of the LM-refined model at the reprojection threshold (3.4 returned the
// estimator's own mask).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "triangulate.h"  // tri_detail::jacobi_right_vectors

namespace xb {
namespace homography {

constexpr int kLMEDS = 4, kRANSAC = 8;  // cv::LMEDS, cv::RANSAC

struct CvRNG {  // cv::RNG: multiply-with-carry
  uint64_t state = 0xffffffffffffffffULL;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

inline int update_num_iters(double p, double ep, int model_points, int max_iters) {  // RANSACUpdateNumIters
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::lrint(num / denom);
}

inline bool have_collinear_points(const float* m, int count) {  // only the last selected point is tested (fundam.cpp)
  const int i = count - 1;
  for (int j = 0; j < i; ++j) {
    const double dx1 = (double)m[2 * j] - m[2 * i], dy1 = (double)m[2 * j + 1] - m[2 * i + 1];
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)m[2 * k] - m[2 * i], dy2 = (double)m[2 * k + 1] - m[2 * i + 1];
      if (std::fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2))) return true;
    }
  }
  return false;
}

inline double det3(const double a[3][3]) {
  return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
         a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

inline bool check_subset(const float* ms1, const float* ms2, int count) {  // HomographyEstimatorCallback::checkSubset
  if (have_collinear_points(ms1, count) || have_collinear_points(ms2, count)) return false;
  if (count == 4) {
    static const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int negative = 0;
    for (const auto& t : tt) {
      double A[3][3], B[3][3];
      for (int r = 0; r < 3; ++r) {
        A[r][0] = ms1[2 * t[r]]; A[r][1] = ms1[2 * t[r] + 1]; A[r][2] = 1.;
        B[r][0] = ms2[2 * t[r]]; B[r][1] = ms2[2 * t[r] + 1]; B[r][2] = 1.;
      }
      negative += det3(A) * det3(B) < 0;
    }
    if (negative != 0 && negative != 4) return false;
  }
  return true;
}

// HomographyEstimatorCallback::runKernel: normalised DLT; H row-major, H[8] = 1.  false = degenerate.
inline bool run_kernel(const float* M, const float* m, int count, double H[9]) {
  double cM[2] = {0, 0}, cm[2] = {0, 0}, sM[2] = {0, 0}, sm[2] = {0, 0};
  for (int i = 0; i < count; ++i) { cm[0] += m[2 * i]; cm[1] += m[2 * i + 1]; cM[0] += M[2 * i]; cM[1] += M[2 * i + 1]; }
  for (int k = 0; k < 2; ++k) { cm[k] /= count; cM[k] /= count; }
  for (int i = 0; i < count; ++i)
    for (int k = 0; k < 2; ++k) { sm[k] += std::fabs(m[2 * i + k] - cm[k]); sM[k] += std::fabs(M[2 * i + k] - cM[k]); }
  if (std::fabs(sm[0]) < DBL_EPSILON || std::fabs(sm[1]) < DBL_EPSILON || std::fabs(sM[0]) < DBL_EPSILON || std::fabs(sM[1]) < DBL_EPSILON) return false;
  for (int k = 0; k < 2; ++k) { sm[k] = count / sm[k]; sM[k] = count / sM[k]; }
  const double inv_hnorm[9] = {1. / sm[0], 0, cm[0], 0, 1. / sm[1], cm[1], 0, 0, 1};
  const double hnorm2[9] = {sM[0], 0, -cM[0] * sM[0], 0, sM[1], -cM[1] * sM[1], 0, 0, 1};
  double H0[9], T[9];
  bool have_h0 = false;
  if (count == 4) {
    // Minimal sample: L (8 x 9) has an exact null vector, which is what the smallest eigenvector of L^T L is.  Solve L h = 0 with
    // h[8] = 1 by elimination (500 flop instead of a 9 x 9 Jacobi sweep set; ~55 of these per frame for LMedS); the general path
    // below takes over when h[8] = 0 makes the system singular.
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
      const double x = (m[2 * i] - cm[0]) * sm[0], y = (m[2 * i + 1] - cm[1]) * sm[1];
      const double X = (M[2 * i] - cM[0]) * sM[0], Y = (M[2 * i + 1] - cM[1]) * sM[1];
      const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
      for (int k = 0; k < 8; ++k) { a[2 * i][k] = Lx[k]; a[2 * i + 1][k] = Ly[k]; }
      a[2 * i][8] = -Lx[8];
      a[2 * i + 1][8] = -Ly[8];
    }
    double amax = 0;
    for (auto& row : a)
      for (int k = 0; k < 8; ++k) amax = std::max(amax, std::fabs(row[k]));
    bool ok = amax > 0;
    for (int k = 0; k < 8 && ok; ++k) {
      int p = k;
      for (int r = k + 1; r < 8; ++r)
        if (std::fabs(a[r][k]) > std::fabs(a[p][k])) p = r;
      if (!(std::fabs(a[p][k]) > 1e-10 * amax)) { ok = false; break; }
      if (p != k)
        for (int c = 0; c < 9; ++c) std::swap(a[p][c], a[k][c]);
      for (int r = k + 1; r < 8; ++r) {
        const double f = a[r][k] / a[k][k];
        for (int c = k; c < 9; ++c) a[r][c] -= f * a[k][c];
      }
    }
    if (ok) {
      for (int r = 7; r >= 0; --r) {
        double acc = a[r][8];
        for (int c = r + 1; c < 8; ++c) acc -= a[r][c] * H0[c];
        H0[r] = acc / a[r][r];
      }
      H0[8] = 1;
      have_h0 = true;
    }
  }
  if (!have_h0) {
  double LtL[81] = {0};
  for (int i = 0; i < count; ++i) {
    const double x = (m[2 * i] - cm[0]) * sm[0], y = (m[2 * i + 1] - cm[1]) * sm[1];
    const double X = (M[2 * i] - cM[0]) * sM[0], Y = (M[2 * i + 1] - cM[1]) * sM[1];
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; ++j)
      for (int k = 0; k < 9; ++k) LtL[9 * j + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  // eigenvector of the smallest eigenvalue of the symmetric PSD L^T L = its right singular vector of the smallest singular value
  double V[81], sg[9];
  tri_detail::jacobi_right_vectors<9, 9>(LtL, V, sg);
  int k0 = 0;
  for (int j = 1; j < 9; ++j)
    if (sg[j] < sg[k0]) k0 = j;
  for (int r = 0; r < 9; ++r) H0[r] = V[9 * r + k0];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = inv_hnorm[3 * r] * H0[c] + inv_hnorm[3 * r + 1] * H0[3 + c] + inv_hnorm[3 * r + 2] * H0[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) H[3 * r + c] = T[3 * r] * hnorm2[c] + T[3 * r + 1] * hnorm2[3 + c] + T[3 * r + 2] * hnorm2[6 + c];
  const double s = 1. / H[8];
  for (int i = 0; i < 9; ++i) H[i] *= s;
  return true;
}

inline void compute_error(const float* M, const float* m, int count, const double H[9], float* err) {  // computeError: model and arithmetic in float
  float Hf[9];
  for (int i = 0; i < 9; ++i) Hf[i] = (float)H[i];
  for (int i = 0; i < count; ++i) {
    const float ww = 1.f / (Hf[6] * M[2 * i] + Hf[7] * M[2 * i + 1] + 1.f);
    const float dx = (Hf[0] * M[2 * i] + Hf[1] * M[2 * i + 1] + Hf[2]) * ww - m[2 * i];
    const float dy = (Hf[3] * M[2 * i] + Hf[4] * M[2 * i + 1] + Hf[5]) * ww - m[2 * i + 1];
    const float e = dx * dx + dy * dy;
    err[i] = e == e ? e : INFINITY;  // a degenerate model gives NaN: same answers as NaN in every comparison below (never an inlier, never the
                                     // best median), but a valid ordering for std::nth_element
  }
}

inline bool get_subset(const float* m1, const float* m2, int count, CvRNG& rng, float ms1[8], float ms2[8], int max_attempts) {
  for (int iters = 0; iters < max_attempts; ++iters) {
    int idx[4];
    for (int i = 0; i < 4; ++i) {
      int k = rng.uniform(0, count);
      while (std::find(idx, idx + i, k) != idx + i) k = rng.uniform(0, count);
      idx[i] = k;
      ms1[2 * i] = m1[2 * k]; ms1[2 * i + 1] = m1[2 * k + 1];
      ms2[2 * i] = m2[2 * k]; ms2[2 * i + 1] = m2[2 * k + 1];
    }
    if (check_subset(ms1, ms2, 4)) return true;
  }
  return false;
}

// The estimator's own mask (what OpenCV 3.4 returned): LMedS with sigma = 2.5*1.4826*(1+5/(n-4))*sqrt(median), or RANSAC's best sample.
inline bool estimator_mask(const float* m1, const float* m2, int n, int method, double thresh, int max_iters, double confidence, std::vector<uint8_t>& mask) {
  mask.assign(n, 0);
  if (n < 4) return false;
  double H[9], best[9];
  if (n == 4) {
    const bool ok = run_kernel(m1, m2, 4, H);
    if (ok) mask.assign(n, 1);
    return ok;
  }
  CvRNG rng;
  std::vector<float> err(n), tmp(n);
  float ms1[8], ms2[8];
  if (method == kLMEDS) {
    const int niters = std::max(update_num_iters(confidence, 0.45, 4, max_iters), 3);
    double min_median = DBL_MAX;
    for (int it = 0; it < niters; ++it) {
      if (!get_subset(m1, m2, n, rng, ms1, ms2, 1000)) {
        if (it == 0) return false;
        break;
      }
      if (!run_kernel(ms1, ms2, 4, H)) continue;
      compute_error(m1, m2, n, H, err.data());
      tmp = err;
      std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
      const double median = tmp[n / 2];
      if (median < min_median) { min_median = median; std::copy(H, H + 9, best); }
    }
    if (!(min_median < DBL_MAX)) return false;
    const double sigma = std::max(2.5 * 1.4826 * (1 + 5. / (n - 4)) * std::sqrt(min_median), 0.001);
    const float t = (float)(sigma * sigma);
    compute_error(m1, m2, n, best, err.data());
    int good = 0;
    for (int i = 0; i < n; ++i) good += (mask[i] = err[i] <= t);
    if (good < 4) { mask.assign(n, 0); return false; }
    return true;
  }
  // RANSAC
  int niters = max_iters, max_good = 0;
  const float t = (float)(thresh * thresh);
  std::vector<uint8_t> cur(n);
  for (int it = 0; it < niters; ++it) {
    if (!get_subset(m1, m2, n, rng, ms1, ms2, 10000)) {
      if (it == 0) return false;
      break;
    }
    if (!run_kernel(ms1, ms2, 4, H)) continue;
    compute_error(m1, m2, n, H, err.data());
    int good = 0;
    for (int i = 0; i < n; ++i) good += (cur[i] = err[i] <= t);
    if (good > std::max(max_good, 3)) {
      mask = cur;
      max_good = good;
      niters = update_num_iters(confidence, (double)(n - good) / n, 4, niters);
    }
  }
  if (max_good <= 0) { mask.assign(n, 0); return false; }
  return true;
}

// Symmetric 8x8 solve by Gaussian elimination with partial pivoting (cv::solve(..., DECOMP_EIG) on a positive definite A + lambda D).
inline bool solve8(const double A[64], const double b[8], double x[8]) {
  double a[8][9];
  for (int r = 0; r < 8; ++r) { for (int c = 0; c < 8; ++c) a[r][c] = A[8 * r + c]; a[r][8] = b[r]; }
  for (int k = 0; k < 8; ++k) {
    int p = k;
    for (int r = k + 1; r < 8; ++r)
      if (std::fabs(a[r][k]) > std::fabs(a[p][k])) p = r;
    if (!(std::fabs(a[p][k]) > 0)) return false;
    if (p != k)
      for (int c = 0; c < 9; ++c) std::swap(a[p][c], a[k][c]);
    for (int r = k + 1; r < 8; ++r) {
      const double f = a[r][k] / a[k][k];
      for (int c = k; c < 9; ++c) a[r][c] -= f * a[k][c];
    }
  }
  for (int r = 7; r >= 0; --r) {
    double s = a[r][8];
    for (int c = r + 1; c < 8; ++c) s -= a[r][c] * x[c];
    x[r] = s / a[r][r];
  }
  return true;
}

// HomographyRefineCallback::compute: residuals r (2n), and on request A = J^T J (8x8), v = J^T r, all in double.
inline void refine_eval(const double h[8], const float* M, const float* m, int n, std::vector<double>& r, double* A, double* v) {
  r.resize(2 * (size_t)n);
  if (A) { std::fill(A, A + 64, 0.0); std::fill(v, v + 8, 0.0); }
  for (int i = 0; i < n; ++i) {
    const double Mx = M[2 * i], My = M[2 * i + 1];
    double ww = h[6] * Mx + h[7] * My + 1.;
    ww = std::fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
    const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
    r[2 * i] = xi - m[2 * i];
    r[2 * i + 1] = yi - m[2 * i + 1];
    if (A) {
      const double Jx[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
      const double Jy[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
      for (int a = 0; a < 8; ++a) {
        for (int b = 0; b < 8; ++b) A[8 * a + b] += Jx[a] * Jx[b] + Jy[a] * Jy[b];
        v[a] += Jx[a] * r[2 * i] + Jy[a] * r[2 * i + 1];
      }
    }
  }
}

// cv::LMSolver (levmarq.cpp LMSolverImpl::run) as findHomography uses it: 10 iterations, epsx = epsf = FLT_EPSILON.
inline void lm_refine(double H[9], const float* M, const float* m, int n, int max_iters = 10) {
  double x[8], xd[8], A[64], v[8], D[8], Ap[64], d[8];
  for (int i = 0; i < 8; ++i) x[i] = H[i];
  std::vector<double> r, rd;
  refine_eval(x, M, m, n, r, A, v);
  auto sq = [](const std::vector<double>& a) { double s = 0; for (double e : a) s += e * e; return s; };
  double S = sq(r);
  for (int i = 0; i < 8; ++i) D[i] = A[9 * i];
  double lambda = 1, lc = 0.75;
  for (int iter = 0;;) {
    for (int i = 0; i < 64; ++i) Ap[i] = A[i];
    for (int i = 0; i < 8; ++i) Ap[9 * i] += lambda * D[i];
    if (!solve8(Ap, v, d)) break;
    for (int i = 0; i < 8; ++i) xd[i] = x[i] - d[i];
    refine_eval(xd, M, m, n, rd, nullptr, nullptr);
    const double Sd = sq(rd);
    double dS = 0, dv = 0;
    for (int a = 0; a < 8; ++a) {
      double Ad = 0;
      for (int b = 0; b < 8; ++b) Ad += A[8 * a + b] * d[b];
      dS += d[a] * (2 * v[a] - Ad);
      dv += d[a] * v[a];
    }
    const double R = (S - Sd) / (std::fabs(dS) > DBL_EPSILON ? dS : 1);
    if (R > 0.75) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < 0.25) {
      double nu = (Sd - S) / (std::fabs(dv) > DBL_EPSILON ? dv : 1) + 2;
      nu = std::min(std::max(nu, 2.), 10.);
      if (lambda == 0) {  // lambda = lc = 1 / max |diag(A^-1)|
        double maxval = DBL_EPSILON;
        for (int k = 0; k < 8; ++k) {
          double e[8] = {0}, col[8];
          e[k] = 1;
          if (solve8(A, e, col)) maxval = std::max(maxval, std::fabs(col[k]));
        }
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (Sd < S) {
      S = Sd;
      for (int i = 0; i < 8; ++i) x[i] = xd[i];
      refine_eval(x, M, m, n, r, A, v);
    }
    ++iter;
    double dinf = 0, rinf = 0;
    for (double e : d) dinf = std::max(dinf, std::fabs(e));
    for (double e : r) rinf = std::max(rinf, std::fabs(e));
    if (!(iter < max_iters && dinf >= FLT_EPSILON && rinf >= FLT_EPSILON)) break;
  }
  for (int i = 0; i < 8; ++i) H[i] = x[i];
  H[8] = 1;
}

// cv::findHomography's mask (OpenCV 4.x): estimator -> re-estimate on its inliers -> LM refine -> inliers of the refined model.
inline bool find_homography_mask(const float* pts0, const float* pts1, int n, int method, double thresh, int max_iters, double confidence,
                                 std::vector<uint8_t>& mask) {
  if (thresh <= 0) thresh = 3;
  if (!estimator_mask(pts0, pts1, n, method, thresh, max_iters, confidence, mask)) { mask.assign(n, 0); return false; }
  if (n == 4) return true;
  std::vector<float> a, b;
  for (int i = 0; i < n; ++i)
    if (mask[i]) { a.push_back(pts0[2 * i]); a.push_back(pts0[2 * i + 1]); b.push_back(pts1[2 * i]); b.push_back(pts1[2 * i + 1]); }
  const int k = (int)a.size() / 2;
  double H[9];
  if (k == 0 || !run_kernel(a.data(), b.data(), k, H)) { mask.assign(n, 0); return false; }
  lm_refine(H, a.data(), b.data(), k);
  std::vector<float> err(n);
  compute_error(pts0, pts1, n, H, err.data());
  const float t = (float)(thresh * thresh);
  for (int i = 0; i < n; ++i) mask[i] = err[i] <= t;
  return true;
}

// Tracker::OutlierRejection (tracker.cpp:705-753): clears the status of the outliers among the points whose status is set.
// Returns false (and leaves *num_rejected alone, like the reference's stale member) when fewer than 4 points are valid.
inline bool tracker_outlier_rejection(const float* pts0, const float* pts1, int n, std::vector<uint8_t>& status, int method, double thresh,
                                      int max_iters, double confidence, int* num_rejected) {
  std::vector<float> a, b;
  std::vector<int> where;
  for (int i = 0; i < n; ++i)
    if (status[i]) { a.push_back(pts0[2 * i]); a.push_back(pts0[2 * i + 1]); b.push_back(pts1[2 * i]); b.push_back(pts1[2 * i + 1]); where.push_back(i); }
  if ((int)where.size() < 4) return false;
  std::vector<uint8_t> mask;
  find_homography_mask(a.data(), b.data(), (int)where.size(), method, thresh, max_iters, confidence, mask);
  int rej = 0;
  for (size_t k = 0; k < where.size(); ++k)
    if (!mask[k]) { status[where[k]] = 0; ++rej; }
  *num_rejected = rej;
  return true;
}

}  // namespace homography
}  // namespace xb
