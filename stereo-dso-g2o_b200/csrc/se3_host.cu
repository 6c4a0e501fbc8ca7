// Host-side SE3 arithmetic in double on row-major [R|t] pairs. Semantics of thirdparty/Sophus/sophus/se3.hpp
// (tangent = [upsilon(3); omega(3)], exp via the V matrix :407-428, log :560-586, Adj :131-139) expressed
// with rotation matrices (Rodrigues) instead of Sophus' unit quaternions. Used for the O(n^2) per-window
// bookkeeping that the reference also does on the host in double: FrameFramePrecalc::set, setAdjointsF,
// the numeric nullspaces of FrameHessian::setStateZero.
#include "ctx.h"
#include <cmath>

namespace sdso {

static inline void hat(const double w[3], double W[9]) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1];
  W[3] = w[2]; W[4] = 0; W[5] = -w[0];
  W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
static inline void mul33(const double A[9], const double B[9], double C[9]) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

void se3_exp(const double a[6], double T[12]) {
  const double* w = a + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  double A, B, C;  // R = I + A W + B W^2 ; V = I + B W + C W^2
  if (th < 1e-5) {
    A = 1 - th2 / 6 + th2 * th2 / 120;
    B = 0.5 - th2 / 24 + th2 * th2 / 720;
    C = 1.0 / 6 - th2 / 120 + th2 * th2 / 5040;
  } else {
    A = std::sin(th) / th;
    B = (1 - std::cos(th)) / th2;
    C = (th - std::sin(th)) / (th2 * th);
  }
  double W[9], W2[9];
  hat(w, W);
  mul33(W, W, W2);
  double R[9], V[9];
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + A * W[i] + B * W2[i];
    V[i] = I + B * W[i] + C * W2[i];
  }
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = V[r * 3] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  }
}

void se3_mul(const double A[12], const double B[12], double C[12]) {
  double out[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
    out[r * 4 + 3] = A[r * 4] * B[3] + A[r * 4 + 1] * B[7] + A[r * 4 + 2] * B[11] + A[r * 4 + 3];
  }
  for (int i = 0; i < 12; i++) C[i] = out[i];
}

void se3_inv(const double A[12], double B[12]) {
  double out[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[r * 4 + c] = A[c * 4 + r];
    out[r * 4 + 3] = -(A[0 * 4 + r] * A[3] + A[1 * 4 + r] * A[7] + A[2 * 4 + r] * A[11]);
  }
  for (int i = 0; i < 12; i++) B[i] = out[i];
}

void se3_log(const double T[12], double a[6]) {
  const double tr = T[0] + T[5] + T[10];
  // omega from the skew part; theta from atan2(|skew|, (tr-1)/2): accurate for the small and moderate
  // angles of inter-keyframe motion (the hot path never takes the log of a rotation near pi)
  const double s[3] = {0.5 * (T[9] - T[6]), 0.5 * (T[2] - T[8]), 0.5 * (T[4] - T[1])};
  const double sn = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  const double cs = 0.5 * (tr - 1);
  const double th = std::atan2(sn, cs);
  double w[3];
  const double k = (sn < 1e-9) ? (1.0 + th * th / 6.0) : th / sn;
  for (int i = 0; i < 3; i++) w[i] = k * s[i];
  const double th2 = th * th;
  double D;  // V^-1 = I - 1/2 W + D W^2
  if (th < 1e-5) D = 1.0 / 12 + th2 / 720;
  else D = (1 - 0.5 * th * std::cos(0.5 * th) / std::sin(0.5 * th)) / th2;
  double W[9], W2[9];
  hat(w, W);
  mul33(W, W, W2);
  const double t[3] = {T[3], T[7], T[11]};
  for (int r = 0; r < 3; r++) {
    double v = 0;
    for (int c = 0; c < 3; c++) v += ((r == c ? 1.0 : 0.0) - 0.5 * W[r * 3 + c] + D * W2[r * 3 + c]) * t[c];
    a[r] = v;
    a[3 + r] = w[r];
  }
}

// The reference's SE3 is a unit quaternion + translation (thirdparty/Sophus/sophus/so3.hpp:95-110, 215-232): a pose that crosses
// into it as a matrix is converted to a quaternion, and products renormalise it, so a rotation never leaves the manifold. Poses
// cross the C ABI as double[12]; this puts the 3x3 block back onto SO(3) the same way (matrix -> quaternion (Shepperd, as
// Eigen::Quaternion(Matrix3)) -> normalise -> matrix). Without it a caller that composes poses with R^T as the inverse (any
// constant-velocity model does) feeds the round-off non-orthonormality of one frame into the next and it grows exponentially.
void so3_normalize(double T[12]) {
  const double m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
  double w, x, y, z;
  const double tr = m00 + m11 + m22;
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0);
    w = 0.5 * t; t = 0.5 / t;
    x = (m21 - m12) * t; y = (m02 - m20) * t; z = (m10 - m01) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    auto M = [&](int r, int c) { return T[r * 4 + c]; };
    double q[3];
    double t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    w = (M(k, j) - M(j, k)) * t;
    q[j] = (M(j, i) + M(i, j)) * t;
    q[k] = (M(k, i) + M(i, k)) * t;
    x = q[0]; y = q[1]; z = q[2];
  }
  const double nrm = std::sqrt(w * w + x * x + y * y + z * z);
  if (!(nrm > 0) || !std::isfinite(nrm)) return;   // not a rotation at all: leave it to the caller's checks
  w /= nrm; x /= nrm; y /= nrm; z /= nrm;
  // Eigen::Quaternion::toRotationMatrix
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T[0] = 1 - (tyy + tzz); T[1] = txy - twz; T[2] = txz + twy;
  T[4] = txy + twz; T[5] = 1 - (txx + tzz); T[6] = tyz - twx;
  T[8] = txz - twy; T[9] = tyz + twx; T[10] = 1 - (txx + tyy);
}

// 6x6 row-major: [R, hat(t) R; 0, R]
void se3_adj(const double T[12], double Ad[36]) {
  double R[9], tx[9], tR[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r * 3 + c] = T[r * 4 + c];
  const double t[3] = {T[3], T[7], T[11]};
  hat(t, tx);
  mul33(tx, R, tR);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
    Ad[r * 6 + c] = R[r * 3 + c];
    Ad[r * 6 + 3 + c] = tR[r * 3 + c];
    Ad[(3 + r) * 6 + c] = 0;
    Ad[(3 + r) * 6 + 3 + c] = R[r * 3 + c];
  }
}

}  // namespace sdso
