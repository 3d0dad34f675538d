// oracle/lie.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).  PARITY UNPINNED.
//
// Minimal SO3 / SE3 / Sim3 algebra used by the CPU restatement of the
// LSD-SLAM hot path.  The reference application links these from Sophus
// (old SE3Group/Sim3Group API) through the un-vendored `lsd-slam` fips import
// (/root/reference/fips.yml:1-4); neither Sophus nor Eigen exist in this
// image, so the published formulas are restated here.
//
// Conventions (Sophus):  tangent order = (translation[3], rotation[3] [, log-scale]),
// quaternion stored (x,y,z,w), increments applied on the LEFT:  T' = exp(inc) * T.
#pragma once
#include <cmath>
#include <cstring>

namespace lsdo {

template <typename S> struct Eps;
template <> struct Eps<float>  { static constexpr float  v = 1e-5f; };
template <> struct Eps<double> { static constexpr double v = 1e-10; };

template <typename S> struct Vec3 {
  S x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(S a, S b, S c) : x(a), y(b), z(c) {}
  S operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  S dot(const Vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
  S norm() const { return std::sqrt(dot(*this)); }
  Vec3 operator+(const Vec3 &o) const { return {x + o.x, y + o.y, z + o.z}; }
  Vec3 operator-(const Vec3 &o) const { return {x - o.x, y - o.y, z - o.z}; }
  Vec3 operator*(S s) const { return {x * s, y * s, z * s}; }
  Vec3 operator/(S s) const { return {x / s, y / s, z / s}; }  // Eigen: coefficient-wise true division
  Vec3 cross(const Vec3 &o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
};

template <typename S> struct Mat3 {
  S m[3][3];
  static Mat3 identity() {
    Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = (i == j) ? S(1) : S(0);
    return r;
  }
  static Mat3 zero() { Mat3 r; std::memset(r.m, 0, sizeof(r.m)); return r; }
  // coefficient-wise product, k summed in ascending order (Eigen's lazy 3x3 product order)
  Vec3<S> operator*(const Vec3<S> &v) const {
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
            m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
  }
  Mat3 operator*(const Mat3 &o) const {
    Mat3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
      r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
    return r;
  }
  Mat3 operator*(S s) const { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[i][j] * s; return r; }
  Mat3 operator+(const Mat3 &o) const { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[i][j] + o.m[i][j]; return r; }
  Mat3 transpose() const { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i]; return r; }
};

template <typename S> inline Mat3<S> hat(const Vec3<S> &w) {
  Mat3<S> r = Mat3<S>::zero();
  r.m[0][1] = -w.z; r.m[0][2] = w.y;
  r.m[1][0] = w.z;  r.m[1][2] = -w.x;
  r.m[2][0] = -w.y; r.m[2][1] = w.x;
  return r;
}

template <typename S> struct Quat {
  S x, y, z, w;
  Quat() : x(0), y(0), z(0), w(1) {}
  Quat(S w_, S x_, S y_, S z_) : x(x_), y(y_), z(z_), w(w_) {}
  Quat operator*(const Quat &b) const {  // Hamilton product (Eigen ordering of terms)
    return Quat(w * b.w - x * b.x - y * b.y - z * b.z,
                w * b.x + x * b.w + y * b.z - z * b.y,
                w * b.y + y * b.w + z * b.x - x * b.z,
                w * b.z + z * b.w + x * b.y - y * b.x);
  }
  Quat conjugate() const { return Quat(w, -x, -y, -z); }
  S squaredNorm() const { return x * x + y * y + z * z + w * w; }
  void normalize() { S n = std::sqrt(squaredNorm()); x /= n; y /= n; z /= n; w /= n; }
  // Eigen's Quaternion::toRotationMatrix
  Mat3<S> toRotationMatrix() const {
    Mat3<S> r;
    const S tx = S(2) * x, ty = S(2) * y, tz = S(2) * z;
    const S twx = tx * w, twy = ty * w, twz = tz * w;
    const S txx = tx * x, txy = ty * x, txz = tz * x;
    const S tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r.m[0][0] = S(1) - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz; r.m[1][1] = S(1) - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = S(1) - (txx + tyy);
    return r;
  }
  Vec3<S> rotate(const Vec3<S> &v) const { return toRotationMatrix() * v; }
};

// SO3 exponential as a unit quaternion (Sophus expAndTheta).
template <typename S> inline Quat<S> so3_exp(const Vec3<S> &omega, S *theta_out) {
  const S theta = omega.norm();
  *theta_out = theta;
  const S half = S(0.5) * theta;
  S imag;
  const S real = std::cos(half);
  if (theta < Eps<S>::v) {
    const S t2 = theta * theta, t4 = t2 * t2;
    imag = S(0.5) - S(1.0 / 48.0) * t2 + S(1.0 / 3840.0) * t4;
  } else {
    imag = std::sin(half) / theta;
  }
  return Quat<S>(real, imag * omega.x, imag * omega.y, imag * omega.z);
}

// SO3 logarithm (Sophus logAndTheta).
template <typename S> inline Vec3<S> so3_log(const Quat<S> &q, S *theta_out = nullptr) {
  const S n2 = q.x * q.x + q.y * q.y + q.z * q.z;
  const S n = std::sqrt(n2);
  const S w = q.w;
  S two_atan_nbyw_by_n;
  if (n < Eps<S>::v) {
    two_atan_nbyw_by_n = S(2) / w - S(2) * n2 / (w * w * w);
  } else if (std::fabs(w) < Eps<S>::v) {
    two_atan_nbyw_by_n = (w > 0 ? S(M_PI) : -S(M_PI)) / n;
  } else {
    two_atan_nbyw_by_n = S(2) * std::atan(n / w) / n;
  }
  if (theta_out) *theta_out = two_atan_nbyw_by_n * n;
  return {two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z};
}

template <typename S> struct SE3 {
  Quat<S> q;
  Vec3<S> t;
  SE3() {}
  SE3(const Quat<S> &q_, const Vec3<S> &t_) : q(q_), t(t_) {}
  Mat3<S> rotationMatrix() const { return q.toRotationMatrix(); }
  SE3 operator*(const SE3 &o) const {
    SE3 r;
    r.q = q * o.q;
    r.q.normalize();
    r.t = t + q.rotate(o.t);
    return r;
  }
  SE3 inverse() const {
    SE3 r;
    r.q = q.conjugate();
    r.t = r.q.rotate(t * S(-1));
    return r;
  }
  template <typename T> SE3<T> cast() const {
    SE3<T> r;
    r.q = Quat<T>(T(q.w), T(q.x), T(q.y), T(q.z));
    r.t = Vec3<T>(T(t.x), T(t.y), T(t.z));
    return r;
  }
  // tangent = (upsilon, omega)
  static SE3 exp(const S tangent[6]) {
    const Vec3<S> ups(tangent[0], tangent[1], tangent[2]);
    const Vec3<S> om(tangent[3], tangent[4], tangent[5]);
    S theta;
    const Quat<S> q = so3_exp(om, &theta);
    const Mat3<S> Om = hat(om);
    const Mat3<S> Om2 = Om * Om;
    Mat3<S> V;
    if (theta < Eps<S>::v) {
      V = q.toRotationMatrix();
    } else {
      const S t2 = theta * theta;
      V = Mat3<S>::identity() + Om * ((S(1) - std::cos(theta)) / t2) + Om2 * ((theta - std::sin(theta)) / (t2 * theta));
    }
    return SE3(q, V * ups);
  }
  void log(S tangent[6]) const {
    S theta;
    const Vec3<S> om = so3_log(q, &theta);
    Mat3<S> Vinv;
    const Mat3<S> Om = hat(om);
    if (theta < Eps<S>::v) {
      Vinv = Mat3<S>::identity() + Om * S(-0.5) + (Om * Om) * S(1.0 / 12.0);
    } else {
      Vinv = Mat3<S>::identity() + Om * S(-0.5) +
             (Om * Om) * ((S(1) - theta / (S(2) * std::tan(theta / S(2)))) / (theta * theta));
    }
    const Vec3<S> u = Vinv * t;
    tangent[0] = u.x; tangent[1] = u.y; tangent[2] = u.z;
    tangent[3] = om.x; tangent[4] = om.y; tangent[5] = om.z;
  }
};

// Sim3 = (scale * R, t).  Kept as unit quaternion + explicit scale.
template <typename S> struct Sim3 {
  Quat<S> q;
  Vec3<S> t;
  S s;
  Sim3() : s(1) {}
  Sim3(const Quat<S> &q_, const Vec3<S> &t_, S s_) : q(q_), t(t_), s(s_) {}
  Mat3<S> rotationMatrix() const { return q.toRotationMatrix(); }
  Mat3<S> rxso3Matrix() const { return q.toRotationMatrix() * s; }
  Sim3 operator*(const Sim3 &o) const {
    Sim3 r;
    r.q = q * o.q;
    r.q.normalize();
    r.s = s * o.s;
    r.t = t + q.rotate(o.t) * s;
    return r;
  }
  Vec3<S> apply(const Vec3<S> &p) const { return q.rotate(p) * s + t; }
  Sim3 inverse() const {
    Sim3 r;
    r.q = q.conjugate();
    r.s = S(1) / s;
    r.t = r.q.rotate(t * S(-1)) * r.s;
    return r;
  }
  template <typename T> Sim3<T> cast() const {
    Sim3<T> r;
    r.q = Quat<T>(T(q.w), T(q.x), T(q.y), T(q.z));
    r.t = Vec3<T>(T(t.x), T(t.y), T(t.z));
    r.s = T(s);
    return r;
  }
  // tangent = (upsilon, omega, sigma);  Sophus Sim3Group::exp with calcW.
  static Sim3 exp(const S tangent[7]) {
    const Vec3<S> ups(tangent[0], tangent[1], tangent[2]);
    const Vec3<S> om(tangent[3], tangent[4], tangent[5]);
    const S sigma = tangent[6];
    S theta;
    const Quat<S> q = so3_exp(om, &theta);
    const S scale = std::exp(sigma);
    const Mat3<S> Om = hat(om);
    const Mat3<S> Om2 = Om * Om;
    S A, B, C;
    if (std::fabs(sigma) < Eps<S>::v) {
      C = S(1);
      if (std::fabs(theta) < Eps<S>::v) { A = S(0.5); B = S(1.0 / 6.0); }
      else { const S t2 = theta * theta; A = (S(1) - std::cos(theta)) / t2; B = (theta - std::sin(theta)) / (t2 * theta); }
    } else {
      C = (scale - S(1)) / sigma;
      if (std::fabs(theta) < Eps<S>::v) {
        const S s2 = sigma * sigma;
        A = ((sigma - S(1)) * scale + S(1)) / s2;
        B = ((S(0.5) * s2 - sigma + S(1)) * scale - S(1)) / (s2 * sigma);
      } else {
        const S t2 = theta * theta;
        const S a = scale * std::sin(theta), b = scale * std::cos(theta), c = t2 + sigma * sigma;
        A = (a * sigma + (S(1) - b) * theta) / (theta * c);
        B = (C - ((b - S(1)) * sigma + a * theta) / c) * S(1) / t2;
      }
    }
    const Mat3<S> W = Om * A + Om2 * B + Mat3<S>::identity() * C;
    return Sim3(q, W * ups, scale);
  }
};

template <typename S> inline SE3<S> se3FromSim3(const Sim3<S> &s) { return SE3<S>(s.q, s.t); }
template <typename S> inline Sim3<S> sim3FromSE3(const SE3<S> &e, S scale) { return Sim3<S>(e.q, e.t, scale); }

// LDL^T solve of the symmetric positive semi-definite N x N system A x = b.  No pivoting (Eigen's ldlt() pivots, which only
// reorders roundoff for these Gram matrices), but rank deficiency is handled the way Eigen::LDLT does it: a pivot with
// |d| <= 1 / highest() leaves its column undivided, and solve() applies the pseudo-inverse of D (that component is 0).
// A Gram matrix with a zero diagonal entry has a zero row and column, so the result equals Eigen's: e.g. Sim3 tracking
// when no warped point has a depth residual (scale row of A = 0) takes a finite step with inc[6] = 0 instead of NaN.
template <typename S> inline S ldlt_tolerance();
template <> inline float ldlt_tolerance<float>() { return 1.0f / 3.402823466e+38f; }
template <> inline double ldlt_tolerance<double>() { return 1.0 / 1.7976931348623157e+308; }
template <typename S, int N> inline void ldlt_solve(const S A_in[N][N], const S b[N], S x[N]) {
  S L[N][N];
  S D[N];
  const S tol = ldlt_tolerance<S>();
  for (int j = 0; j < N; j++) {
    S d = A_in[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    const bool pivotValid = std::fabs(d) > tol;
    for (int i = j + 1; i < N; i++) {
      S v = A_in[i][j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * D[k];
      L[i][j] = pivotValid ? v / d : v;
    }
  }
  S y[N];
  for (int i = 0; i < N; i++) {
    S v = b[i];
    for (int k = 0; k < i; k++) v -= L[i][k] * y[k];
    y[i] = v;
  }
  for (int i = 0; i < N; i++) y[i] = std::fabs(D[i]) > tol ? y[i] / D[i] : S(0);
  for (int i = N - 1; i >= 0; i--) {
    S v = y[i];
    for (int k = i + 1; k < N; k++) v -= L[k][i] * x[k];
    x[i] = v;
  }
}

}  // namespace lsdo
