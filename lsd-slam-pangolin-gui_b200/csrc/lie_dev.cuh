// lie_dev.cuh -- SO3/SE3/Sim3 exponentials, quaternion algebra and small LDL^T solves used by the
// on-device Levenberg-Marquardt step.  Replaces the Sophus (SE3f::exp, Sim3d::exp, operator*) and
// Eigen (Matrix::ldlt().solve) calls inside [UP] SE3Tracker::trackFrame / Sim3Tracker::trackFrameSim3.
// Tangent order (translation, rotation[, log scale]); quaternions (x,y,z,w); left-multiplicative updates.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lsd {

template <typename S> struct QuatT {
  S x, y, z, w;
};

template <typename S> __host__ __device__ inline QuatT<S> qmul(const QuatT<S> &a, const QuatT<S> &b) {
  QuatT<S> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

template <typename S> __host__ __device__ inline void qnormalize(QuatT<S> &q) {
  const S n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// Eigen Quaternion::toRotationMatrix (row-major 3x3 out)
template <typename S> __host__ __device__ inline void qtoR(const QuatT<S> &q, S R[9]) {
  const S tx = S(2) * q.x, ty = S(2) * q.y, tz = S(2) * q.z;
  const S twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const S txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const S tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = S(1) - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = S(1) - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = S(1) - (txx + tyy);
}

template <typename S> __host__ __device__ inline void mat3vec(const S R[9], const S v[3], S o[3]) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}

template <typename S> struct LieEps;
template <> struct LieEps<float> { __host__ __device__ static float v() { return 1e-5f; } };
template <> struct LieEps<double> { __host__ __device__ static double v() { return 1e-10; } };

// SO3 exp -> unit quaternion
template <typename S> __host__ __device__ inline QuatT<S> so3_exp(const S om[3], S *theta_out) {
  const S theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  *theta_out = theta;
  const S half = S(0.5) * theta;
  S imag;
  const S real = cos(half);
  if (theta < LieEps<S>::v()) {
    const S t2 = theta * theta, t4 = t2 * t2;
    imag = S(0.5) - S(1.0 / 48.0) * t2 + S(1.0 / 3840.0) * t4;
  } else {
    imag = sin(half) / theta;
  }
  QuatT<S> q;
  q.w = real; q.x = imag * om[0]; q.y = imag * om[1]; q.z = imag * om[2];
  return q;
}

// hat(om) and hat(om)^2, row-major
template <typename S> __host__ __device__ inline void hat2(const S om[3], S Om[9], S Om2[9]) {
  Om[0] = 0; Om[1] = -om[2]; Om[2] = om[1];
  Om[3] = om[2]; Om[4] = 0; Om[5] = -om[0];
  Om[6] = -om[1]; Om[7] = om[0]; Om[8] = 0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
}

// SE3 exp(inc) * (q, t)  -> (q', t')
template <typename S>
__host__ __device__ inline void se3_exp_compose(const S inc[6], const QuatT<S> &q, const S t[3], QuatT<S> &qo, S to[3]) {
  S theta;
  const S om[3] = {inc[3], inc[4], inc[5]};
  const QuatT<S> dq = so3_exp(om, &theta);
  S Om[9], Om2[9], V[9];
  hat2(om, Om, Om2);
  if (theta < LieEps<S>::v()) {
    qtoR(dq, V);
  } else {
    const S t2 = theta * theta;
    const S a = (S(1) - cos(theta)) / t2, b = (theta - sin(theta)) / (t2 * theta);
#pragma unroll
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? S(1) : S(0)) + Om[i] * a + Om2[i] * b;
  }
  const S ups[3] = {inc[0], inc[1], inc[2]};
  S dt[3];
  mat3vec(V, ups, dt);
  qo = qmul(dq, q);
  qnormalize(qo);
  S dR[9], rt[3];
  qtoR(dq, dR);
  mat3vec(dR, t, rt);
  to[0] = dt[0] + rt[0]; to[1] = dt[1] + rt[1]; to[2] = dt[2] + rt[2];
}

// Sim3 exp(inc) * (q, t, s)
template <typename S>
__host__ __device__ inline void sim3_exp_compose(const S inc[7], const QuatT<S> &q, const S t[3], S s, QuatT<S> &qo, S to[3],
                                                 S &so) {
  S theta;
  const S om[3] = {inc[3], inc[4], inc[5]};
  const S sigma = inc[6];
  const QuatT<S> dq = so3_exp(om, &theta);
  const S scale = exp(sigma);
  S Om[9], Om2[9], Wm[9];
  hat2(om, Om, Om2);
  S A, B, Cc;
  if (fabs(sigma) < LieEps<S>::v()) {
    Cc = S(1);
    if (fabs(theta) < LieEps<S>::v()) { A = S(0.5); B = S(1.0 / 6.0); }
    else { const S t2 = theta * theta; A = (S(1) - cos(theta)) / t2; B = (theta - sin(theta)) / (t2 * theta); }
  } else {
    Cc = (scale - S(1)) / sigma;
    if (fabs(theta) < LieEps<S>::v()) {
      const S s2 = sigma * sigma;
      A = ((sigma - S(1)) * scale + S(1)) / s2;
      B = ((S(0.5) * s2 - sigma + S(1)) * scale - S(1)) / (s2 * sigma);
    } else {
      const S t2 = theta * theta;
      const S a = scale * sin(theta), b = scale * cos(theta), c = t2 + sigma * sigma;
      A = (a * sigma + (S(1) - b) * theta) / (theta * c);
      B = (Cc - ((b - S(1)) * sigma + a * theta) / c) * S(1) / t2;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) Wm[i] = Om[i] * A + Om2[i] * B + ((i % 4 == 0) ? Cc : S(0));
  const S ups[3] = {inc[0], inc[1], inc[2]};
  S dt[3];
  mat3vec(Wm, ups, dt);
  qo = qmul(dq, q);
  qnormalize(qo);
  so = scale * s;
  S dR[9], rt[3];
  qtoR(dq, dR);
  mat3vec(dR, t, rt);
  to[0] = dt[0] + rt[0] * scale; to[1] = dt[1] + rt[1] * scale; to[2] = dt[2] + rt[2] * scale;
}

// LDL^T solve (A symmetric positive semi-definite, row-major N x N).  No pivoting, but rank deficiency is handled the way
// Eigen::LDLT (upstream's A.ldlt().solve(b)) handles it: a pivot with |d| <= 1 / highest() leaves its column undivided and the
// solve applies the pseudo-inverse of D (that component of x is 0).  For a Gram matrix a zero diagonal entry means a zero
// row and column, so this equals Eigen's pivoted result: Sim3 tracking with no depth residual under the warped points
// (scale row of A = 0) takes a finite step with inc[6] = 0 instead of NaN.
template <typename S> struct LdltTol;
template <> struct LdltTol<float> { __host__ __device__ static float v() { return 1.0f / 3.402823466e+38f; } };
template <> struct LdltTol<double> { __host__ __device__ static double v() { return 1.0 / 1.7976931348623157e+308; } };

template <typename S, int N> __host__ __device__ inline void ldlt_solve(const S *A, const S *b, S *x) {
  S L[N * N];
  S D[N];
  const S tol = LdltTol<S>::v();
#pragma unroll
  for (int j = 0; j < N; j++) {
    S d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j * N + k] * L[j * N + k] * D[k];
    D[j] = d;
    const bool pivotValid = fabs(d) > tol;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      S v = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; k++) v -= L[i * N + k] * L[j * N + k] * D[k];
      L[i * N + j] = pivotValid ? v / d : v;
    }
  }
  S y[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    S v = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) v -= L[i * N + k] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < N; i++) y[i] = fabs(D[i]) > tol ? y[i] / D[i] : S(0);
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    S v = y[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) v -= L[k * N + i] * x[k];
    x[i] = v;
  }
}

}  // namespace lsd
