// Host-side fp64 geometry helpers for the estimator state machine (SO(3), SE(3), cameras).
// The nominal state lives on the host (it is a few hundred doubles); only the covariance and the
// per-feature hot loops live on the device.
#pragma once
#include <array>
#include <cmath>

#include "common.cuh"

namespace xb {

__host__ __device__ inline M3 so3_exp(const V3& w) {
  // Sophus SO3::exp goes through a unit quaternion; same map to rounding.
  const double th2 = w.v[0] * w.v[0] + w.v[1] * w.v[1] + w.v[2] * w.v[2];
  const double th = sqrt(th2);
  const M3 W = m3_hat(w);
  const M3 W2 = m3_mul(W, W);
  double a, b;
  if (th < 1e-10) {
    a = 1.0;
    b = 0.5;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / th2;
  }
  M3 R = m3_eye();
  for (int i = 0; i < 9; ++i) R.m[i] += a * W.m[i] + b * W2.m[i];
  return R;
}

__host__ __device__ inline V3 so3_log(const M3& R) {
  double c = 0.5 * (R.m[0] + R.m[4] + R.m[8] - 1.0);
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = acos(c);
  V3 v{{R.m[7] - R.m[5], R.m[2] - R.m[6], R.m[3] - R.m[1]}};
  const double k = th < 1e-10 ? 0.5 : th / (2.0 * sin(th));
  return V3{{k * v.v[0], k * v.v[1], k * v.v[2]}};
}

// SO3::normalize(): Sophus renormalises its unit quaternion.  Matrix -> quaternion -> normalise ->
// matrix is the same projection for a nearly orthonormal R.
__host__ __device__ inline M3 so3_normalize(const M3& R) {
  double q[4];  // w x y z
  const double tr = R.m[0] + R.m[4] + R.m[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s;
    q[1] = (R.m[7] - R.m[5]) / s;
    q[2] = (R.m[2] - R.m[6]) / s;
    q[3] = (R.m[3] - R.m[1]) / s;
  } else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) {
    double s = sqrt(1.0 + R.m[0] - R.m[4] - R.m[8]) * 2;
    q[0] = (R.m[7] - R.m[5]) / s;
    q[1] = 0.25 * s;
    q[2] = (R.m[1] + R.m[3]) / s;
    q[3] = (R.m[2] + R.m[6]) / s;
  } else if (R.m[4] > R.m[8]) {
    double s = sqrt(1.0 + R.m[4] - R.m[0] - R.m[8]) * 2;
    q[0] = (R.m[2] - R.m[6]) / s;
    q[1] = (R.m[1] + R.m[3]) / s;
    q[2] = 0.25 * s;
    q[3] = (R.m[5] + R.m[7]) / s;
  } else {
    double s = sqrt(1.0 + R.m[8] - R.m[0] - R.m[4]) * 2;
    q[0] = (R.m[3] - R.m[1]) / s;
    q[1] = (R.m[2] + R.m[6]) / s;
    q[2] = (R.m[5] + R.m[7]) / s;
    q[3] = 0.25 * s;
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  return M3{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
             2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
}

struct SE3h {
  M3 R = m3_eye();
  V3 T{{0, 0, 0}};
};
inline SE3h se3_mul(const SE3h& a, const SE3h& b) { return SE3h{m3_mul(a.R, b.R), v3_add(m3_mulv(a.R, b.T), a.T)}; }
inline SE3h se3_inv(const SE3h& a) {
  const M3 Rt = m3_t(a.R);
  const V3 t = m3_mulv(Rt, a.T);
  return SE3h{Rt, V3{{-t.v[0], -t.v[1], -t.v[2]}}};
}
inline V3 se3_apply(const SE3h& a, const V3& x) { return v3_add(m3_mulv(a.R, x), a.T); }
__host__ __device__ inline V3 v3_scale(const V3& a, double s) { return V3{{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; }
__host__ __device__ inline double v3_norm(const V3& a) { return sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]); }
__host__ __device__ inline V3 v3_cross(const V3& a, const V3& b) {
  return V3{{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}

// Camera unprojection (pixels -> normalised camera coordinates).
// pinhole:  /root/reference/common/camera_pinhole.h:39-52
// equidist: /root/reference/common/camera_equidist.h:97-160 (15 Newton iterations on theta)
inline void camera_unproject(const CameraParams& c, double u, double v, double* x, double* y) {
  if (c.model == 0) {
    *x = (u - c.cx) / c.fx;
    *y = (v - c.cy) / c.fy;
    return;
  }
  const double xn = u - c.cx, yn = v - c.cy;
  const double b = c.fx * yn, a = c.fy * xn;
  const double phi = atan2(b, a);
  const double cp = cos(phi), sp = sin(phi);
  const double rth = xn / (c.fx * cp);
  double th = rth;
  for (int i = 0; i < 15; ++i) {
    const double th2 = th * th, th3 = th2 * th, th4 = th2 * th2, th6 = th4 * th2;
    const double x0 = c.k0 * th3 + c.k1 * th4 * th + c.k2 * th6 * th + c.k3 * th6 * th3 - rth + th;
    const double x1 = 3 * c.k0 * th2 + 5 * c.k1 * th4 + 7 * c.k2 * th6 + 9 * c.k3 * th6 * th2 + 1;
    const double d = 2 * x0 * x1;
    const double d2 = 4 * th * x0 * (3 * c.k0 + 10 * c.k1 * th2 + 21 * c.k2 * th4 + 36 * c.k3 * th6) + 2 * x1 * x1;
    th -= d / d2;
  }
  const double t = tan(th);
  *x = t * cp;
  *y = t * sp;
}

}  // namespace xb
