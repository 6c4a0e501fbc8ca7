// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// SE3 restated from thirdparty/Sophus/sophus/{so3.hpp,se3.hpp}: unit quaternion + translation,
// tangent order [upsilon(3); omega(3)], small-angle epsilon 1e-10 (sophus.hpp:45-46).
#include "oracle_common.hpp"

namespace orc {

static const double kEps = 1e-10;  // SophusConstants<double>::epsilon()

void SE3::rotationMatrix(double R[9]) const {
  // Eigen::QuaternionBase::toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void SE3::toMat34(double M[12]) const {
  double R[9]; rotationMatrix(R);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) M[r * 4 + c] = R[r * 3 + c]; M[r * 4 + 3] = t[r]; }
}

SE3 SE3::fromMat34(const double M[12]) {
  // Eigen quaternion-from-matrix (internal::quaternionbase_assign_impl<..,3,3>)
  SE3 s;
  const double m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[4], m11 = M[5], m12 = M[6], m20 = M[8], m21 = M[9], m22 = M[10];
  double tr = m00 + m11 + m22;
  if (tr > 0) {
    double tt = std::sqrt(tr + 1.0);
    s.q.w = 0.5 * tt; tt = 0.5 / tt;
    s.q.x = (m21 - m12) * tt; s.q.y = (m02 - m20) * tt; s.q.z = (m10 - m01) * tt;
  } else {
    const double m[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
    int i = 0; if (m11 > m00) i = 1; if (m22 > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double tt = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * tt; tt = 0.5 / tt;
    s.q.w = (m[k][j] - m[j][k]) * tt;
    v[j] = (m[j][i] + m[i][j]) * tt;
    v[k] = (m[k][i] + m[i][k]) * tt;
    s.q.x = v[0]; s.q.y = v[1]; s.q.z = v[2];
  }
  double n = std::sqrt(s.q.w * s.q.w + s.q.x * s.q.x + s.q.y * s.q.y + s.q.z * s.q.z);
  s.q.w /= n; s.q.x /= n; s.q.y /= n; s.q.z /= n;
  s.t[0] = M[3]; s.t[1] = M[7]; s.t[2] = M[11];
  return s;
}

static Quat so3_exp(const double omega[3], double* theta) {
  // so3.hpp:343-369
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  *theta = std::sqrt(theta_sq);
  const double half_theta = 0.5 * (*theta);
  double imag_factor, real_factor;
  if ((*theta) < kEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real_factor = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double sin_half_theta = std::sin(half_theta);
    imag_factor = sin_half_theta / (*theta);
    real_factor = std::cos(half_theta);
  }
  Quat q; q.w = real_factor; q.x = imag_factor * omega[0]; q.y = imag_factor * omega[1]; q.z = imag_factor * omega[2];
  return q;
}

static void hat(const double w[3], double O[9]) {
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
static void mm3(const double A[9], const double B[9], double C[9]) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
    double s = 0; for (int k = 0; k < 3; k++) s += A[r * 3 + k] * B[k * 3 + c];
    C[r * 3 + c] = s;
  }
}

SE3 SE3::exp(const double a[6]) {
  // se3.hpp:407-428
  const double* omega = a + 3;
  double theta;
  SE3 out; out.q = so3_exp(omega, &theta);
  double Omega[9], Omega_sq[9], V[9];
  hat(omega, Omega); mm3(Omega, Omega, Omega_sq);
  if (theta < kEps) {
    out.rotationMatrix(V);
  } else {
    const double theta_sq = theta * theta;
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Omega[i] + c2 * Omega_sq[i];
  }
  for (int r = 0; r < 3; r++) out.t[r] = V[r * 3 + 0] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  return out;
}

void SE3::log(double out[6]) const {
  // so3.hpp:491-531
  const double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const double n = std::sqrt(squared_n);
  const double w = q.w;
  double two_atan_nbyw_by_n;
  if (n < kEps) {
    const double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - 2.0 * squared_n / (w * squared_w);
  } else {
    if (std::fabs(w) < kEps) two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
    else two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
  }
  const double theta = two_atan_nbyw_by_n * n;
  double om[3] = {two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z};
  // se3.hpp:560-586
  double Omega[9], Omega_sq[9], Vinv[9];
  hat(om, Omega); mm3(Omega, Omega, Omega_sq);
  double c;
  if (std::fabs(theta) < kEps) c = 1.0 / 12.0;
  else c = (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
  for (int i = 0; i < 9; i++) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * Omega[i] + c * Omega_sq[i];
  for (int r = 0; r < 3; r++) out[r] = Vinv[r * 3 + 0] * t[0] + Vinv[r * 3 + 1] * t[1] + Vinv[r * 3 + 2] * t[2];
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

static void qrot(const Quat& q, const double p[3], double out[3]) {
  // Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 * q.vec x v
  double uv[3] = {q.y * p[2] - q.z * p[1], q.z * p[0] - q.x * p[2], q.x * p[1] - q.y * p[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = p[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = p[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = p[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

void SE3::act(const double p[3], double out[3]) const {
  double r[3]; qrot(q, p, r);
  out[0] = r[0] + t[0]; out[1] = r[1] + t[1]; out[2] = r[2] + t[2];
}

SE3 SE3::inverse() const {
  SE3 o; o.q.w = q.w; o.q.x = -q.x; o.q.y = -q.y; o.q.z = -q.z;
  double mt[3] = {-t[0], -t[1], -t[2]};
  qrot(o.q, mt, o.t);
  return o;
}

SE3 SE3::operator*(const SE3& b) const {
  SE3 o;
  double rt[3]; qrot(q, b.t, rt);
  o.t[0] = t[0] + rt[0]; o.t[1] = t[1] + rt[1]; o.t[2] = t[2] + rt[2];
  o.q.w = q.w * b.q.w - q.x * b.q.x - q.y * b.q.y - q.z * b.q.z;
  o.q.x = q.w * b.q.x + q.x * b.q.w + q.y * b.q.z - q.z * b.q.y;
  o.q.y = q.w * b.q.y + q.y * b.q.w + q.z * b.q.x - q.x * b.q.z;
  o.q.z = q.w * b.q.z + q.z * b.q.w + q.x * b.q.y - q.y * b.q.x;
  // so3.hpp normalize(): divide by the norm
  double n = std::sqrt(o.q.w * o.q.w + o.q.x * o.q.x + o.q.y * o.q.y + o.q.z * o.q.z);
  o.q.w /= n; o.q.x /= n; o.q.y /= n; o.q.z /= n;
  return o;
}

void SE3::Adj(double A[36]) const {
  // se3.hpp:131-139 : [R, hat(t) R; 0, R]
  double R[9], T[9], TR[9];
  rotationMatrix(R); hat(t, T); mm3(T, R, TR);
  for (int i = 0; i < 36; i++) A[i] = 0;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
    A[r * 6 + c] = R[r * 3 + c];
    A[(r + 3) * 6 + (c + 3)] = R[r * 3 + c];
    A[r * 6 + (c + 3)] = TR[r * 3 + c];
  }
}

}  // namespace orc
