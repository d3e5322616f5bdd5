// Two-view depth triangulation before the depth sub-filter (host side, fp64).
//   Feature::Triangulate            /root/reference/src/feature.cpp:686-751
//   DirectLinearTransformSVD/Avg    /root/reference/src/helpers.cpp:103-154
//   L1Angular / L2Angular / LinfAngular (Lee & Civera, arXiv:1903.09115)   helpers.cpp:157-327
//   check_cheirality / check_angular_reprojection / check_parallax        helpers.cpp:330-372
// It runs once in a feature's life (on its second observation, manager.cpp:229-231) and costs a few
// hundred flop, so it stays on the host next to Feature::Predict; the sub-filter kernel then reads the
// triangulated state like any other.  The reference keeps several intermediates in `float`
// (a0/a1, the lambdas, the angles; the thresholds are float parameters): same narrowings here.
// One documented deviation (oracle/ekf_oracle.py:_acos_f32): the cosine handed to acos
// is clamped to [-1, 1]; the reference's value for the ray that L1Angular leaves unchanged is
// acos(1 +- ulp), i.e. NaN or 0 depending on the compiler's FMA contraction.
#pragma once
#include <cmath>
#include <string>

#include "hostmath.h"

namespace xb {

enum class TriMethod { DLT_SVD, DLT_AVG, L1, L2, LINF };

inline bool tri_method_from_string(const std::string& s, TriMethod* m) {
  if (s == "direct_linear_transform_svd") *m = TriMethod::DLT_SVD;
  else if (s == "direct_linear_transform_avg") *m = TriMethod::DLT_AVG;
  else if (s == "l1_angular") *m = TriMethod::L1;
  else if (s == "l2_angular") *m = TriMethod::L2;
  else if (s == "linf_angular") *m = TriMethod::LINF;
  else return false;
  return true;
}

struct TriOptions {  // TriangulateOptions (options.h:35-41) as parsed at estimator.cpp:159-164
  TriMethod method = TriMethod::L1;
  double zmin = 0.05, zmax = 5.0;
  double max_theta_thresh = 0.1 * M_PI / 180, beta_thresh = 0.25 * M_PI / 180;  // radians
};

namespace tri_detail {
inline double dot(const V3& a, const V3& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
inline V3 unit(const V3& a) { return v3_scale(a, 1.0 / v3_norm(a)); }
inline V3 normalized_bearing(const double xc[2]) {  // Vec3 f{x, y, 1}; f.normalize()
  const V3 f{{xc[0], xc[1], 1.0}};
  const double n = v3_norm(f);
  return V3{{f.v[0] / n, f.v[1] / n, f.v[2] / n}};
}
inline float acos_clamped(double c) {
  if (c != c) return std::nanf("");
  return (float)std::acos(c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c));
}
inline float std_max(float a, float b) { return a < b ? b : a; }

// One-sided (Hestenes) Jacobi SVD of an R x C matrix stored row-major in a[R*C]: rotates column pairs until
// they are orthogonal, accumulating V (C x C, row-major).  Afterwards column j of `a` has norm sigma_j and
// column j of V is the matching right singular vector.  R, C <= 4.
template <int R, int C>
inline void jacobi_right_vectors(double* a, double* V, double* sigma) {
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) V[i * C + j] = i == j;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p + 1 < C; ++p)
      for (int q = p + 1; q < C; ++q) {
        double app = 0, aqq = 0, apq = 0;
        for (int r = 0; r < R; ++r) {
          app += a[r * C + p] * a[r * C + p];
          aqq += a[r * C + q] * a[r * C + q];
          apq += a[r * C + p] * a[r * C + q];
        }
        if (apq == 0.0 || std::fabs(apq) <= 1e-17 * std::sqrt(app * aqq)) continue;
        rotated = true;
        const double zeta = (aqq - app) / (2.0 * apq);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < R; ++r) {
          const double x = a[r * C + p], y = a[r * C + q];
          a[r * C + p] = cs * x - sn * y;
          a[r * C + q] = sn * x + cs * y;
        }
        for (int r = 0; r < C; ++r) {
          const double x = V[r * C + p], y = V[r * C + q];
          V[r * C + p] = cs * x - sn * y;
          V[r * C + q] = sn * x + cs * y;
        }
      }
    if (!rotated) break;
  }
  for (int j = 0; j < C; ++j) {
    double s = 0;
    for (int r = 0; r < R; ++r) s += a[r * C + j] * a[r * C + j];
    sigma[j] = std::sqrt(s);
  }
}

inline bool checks(const V3& z, const V3& t10, const V3& m0, const V3& Rf0p, const V3& m1, const V3& f1p, float max_theta, float beta_thresh) {
  const double zn2 = std::pow(v3_norm(z), 2);
  const float lambda0 = (float)(dot(z, v3_cross(t10, f1p)) / zn2);
  const float lambda1 = (float)(dot(z, v3_cross(t10, Rf0p)) / zn2);
  if (lambda0 <= 0 || lambda1 <= 0) return false;  // check_cheirality
  const float theta0 = acos_clamped(dot(m0, Rf0p) / (v3_norm(m0) * v3_norm(Rf0p)));
  const float theta1 = acos_clamped(dot(m1, f1p) / (v3_norm(m1) * v3_norm(f1p)));
  if (std_max(theta0, theta1) > max_theta) return false;  // check_angular_reprojection
  const float beta = acos_clamped(dot(f1p, Rf0p) / (v3_norm(f1p) * v3_norm(Rf0p)));
  if (beta < beta_thresh) return false;  // check_parallax
  return true;
}

// common tail of the three angular methods: depth along the corrected ray of view 1, moved to view 0
inline bool angular_finish(const SE3h& g01, const V3& t10, const V3& m0, const V3& m1, const V3& m0p, const V3& m1p, float max_theta,
                           float beta_thresh, V3* X) {
  const V3 z = v3_cross(m1p, m0p);
  const double s = dot(z, v3_cross(t10, m0p)) / std::pow(v3_norm(z), 2);
  *X = se3_apply(g01, v3_scale(m1p, s));
  return checks(z, t10, m0, m0p, m1, m1p, max_theta, beta_thresh);
}
inline V3 reject(const V3& m, const V3& n) { return v3_sub(m, v3_scale(n, dot(m, n))); }  // m - (m.n) n
}  // namespace tri_detail

// g01: pose of view 1 (the newest observation) in view 0 (the reference group's camera); xc0 / xc1 the
// normalised image coordinates of the first / newest observation.  X: the point in view 0.
inline bool triangulate_two_view(const TriOptions& o, const SE3h& g01, const double xc0[2], const double xc1[2], V3* X) {
  using namespace tri_detail;
  const float max_theta = (float)o.max_theta_thresh, beta_thresh = (float)o.beta_thresh;
  if (o.method == TriMethod::DLT_SVD) {  // helpers.cpp:103-129
    const M3 Rt = m3_t(g01.R);
    const V3 tt = m3_mulv(Rt, g01.T);
    const double P1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
    const double P2[3][4] = {{Rt.m[0], Rt.m[1], Rt.m[2], -tt.v[0]}, {Rt.m[3], Rt.m[4], Rt.m[5], -tt.v[1]}, {Rt.m[6], Rt.m[7], Rt.m[8], -tt.v[2]}};
    const V3 f1 = normalized_bearing(xc0), f2 = normalized_bearing(xc1);
    double A[16], V[16], sg[4];
    for (int j = 0; j < 4; ++j) {
      A[0 * 4 + j] = f1.v[0] * P1[2][j] - f1.v[2] * P1[0][j];
      A[1 * 4 + j] = f1.v[1] * P1[2][j] - f1.v[2] * P1[1][j];
      A[2 * 4 + j] = f2.v[0] * P2[2][j] - f2.v[2] * P2[0][j];
      A[3 * 4 + j] = f2.v[1] * P2[2][j] - f2.v[2] * P2[1][j];
    }
    jacobi_right_vectors<4, 4>(A, V, sg);
    int k = 0;
    for (int j = 1; j < 4; ++j)
      if (sg[j] < sg[k]) k = j;  // V(:, 3) of a decreasing-order SVD = the smallest singular value's vector
    const double w = V[3 * 4 + k];
    *X = V3{{V[0 * 4 + k] / w, V[1 * 4 + k] / w, V[2 * 4 + k] / w}};
    return true;
  }
  if (o.method == TriMethod::DLT_AVG) {  // helpers.cpp:131-154 (mid-point of the closest points of the two rays)
    const V3 f1 = normalized_bearing(xc0), f2 = normalized_bearing(xc1);
    const V3 f2u = m3_mulv(g01.R, f2);
    const double b0 = dot(g01.T, f1), b1 = dot(g01.T, f2u);
    const double a00 = dot(f1, f1), a10 = dot(f1, f2u), a01 = -a10, a11 = -dot(f2u, f2u);
    const double invdet = 1.0 / (a00 * a11 - a01 * a10);  // Eigen's 2x2 inverse: adjugate * (1 / det)
    const double l0 = (a11 * invdet) * b0 + (-a01 * invdet) * b1;
    const double l1 = (-a10 * invdet) * b0 + (a00 * invdet) * b1;
    const V3 xm = v3_scale(f1, l0), xn = v3_add(g01.T, v3_scale(f2u, l1));
    *X = v3_scale(v3_add(xm, xn), 0.5);
    return true;
  }
  const M3 R10 = m3_t(g01.R);
  const V3 t10 = v3_scale(m3_mulv(R10, g01.T), -1.0);
  const V3 m0 = m3_mulv(R10, normalized_bearing(xc0)), m1 = normalized_bearing(xc1);
  V3 m0p, m1p;
  if (o.method == TriMethod::L1) {  // helpers.cpp:157-215
    const float a0 = (float)v3_norm(v3_cross(unit(m0), t10));
    const float a1 = (float)v3_norm(v3_cross(unit(m1), t10));
    if (a0 <= a1) {
      m0p = reject(m0, unit(v3_cross(m1, t10)));
      m1p = m1;
    } else {
      m0p = m0;
      m1p = reject(m1, unit(v3_cross(m0, t10)));
    }
  } else if (o.method == TriMethod::L2) {  // helpers.cpp:218-275
    const V3 m0h = unit(m0), m1h = unit(m1), th = unit(t10);
    // B = [m0^ m1^]^T (I - t^ t^T) is 2 x 3; its right singular vectors are the orthogonalised columns of B^T (3 x 2):
    // V.col(1) of the reference's JacobiSVD = the one with the smaller singular value (sign cancels below)
    double Bt[6], V2[4], sg[2];
    const V3 r0 = reject(m0h, th), r1 = reject(m1h, th);  // rows of B: m^T (I - t t^T) = (m - (m.t) t)^T
    for (int i = 0; i < 3; ++i) { Bt[i * 2 + 0] = r0.v[i]; Bt[i * 2 + 1] = r1.v[i]; }
    jacobi_right_vectors<3, 2>(Bt, V2, sg);
    const int k = sg[0] <= sg[1] ? 0 : 1;
    const V3 n{{Bt[0 * 2 + k] / sg[k], Bt[1 * 2 + k] / sg[k], Bt[2 * 2 + k] / sg[k]}};
    m0p = reject(m0, n);
    m1p = reject(m1, n);
  } else {  // LINF, helpers.cpp:277-327; n' is used unnormalised there (`n_prime_hat` is n_a or n_b as computed): kept
    const V3 m0h = unit(m0), m1h = unit(m1);
    const V3 na = v3_cross(v3_add(m0h, m1h), t10), nb = v3_cross(v3_sub(m0h, m1h), t10);
    const V3 n = v3_norm(na) >= v3_norm(nb) ? na : nb;
    m0p = reject(m0, n);
    m1p = reject(m1, n);
  }
  return angular_finish(g01, t10, m0, m1, m0p, m1p, max_theta, beta_thresh, X);
}

// Feature::Triangulate's acceptance (feature.cpp:733-748): returns true and the new feature state
// [x/z, y/z, log z] when the triangulated depth lies in [zmin, zmax].
inline bool triangulate_feature_state(const TriOptions& o, const SE3h& g01, const double xc0[2], const double xc1[2], double x_out[3]) {
  V3 X;
  if (!triangulate_two_view(o, g01, xc0, xc1, &X)) return false;
  const double z = X.v[2];
  if (z < o.zmin || z > o.zmax) return false;
  x_out[0] = X.v[0] / z;
  x_out[1] = X.v[1] / z;
  x_out[2] = std::log(z);
  return true;
}

}  // namespace xb
