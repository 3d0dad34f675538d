// oracle/sim3.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Restates upstream Tracking/Sim3Tracker.cpp + least_squares.cpp (LGS4 / LGS6 / LGS7) of the lsd-slam
// core (un-vendored: /root/reference/fips.yml:1-4) per SURVEY.md 3.6, 8a B8-B11 and Appendix A.4.
// The reference reaches this code through SlamSystem's constraint-search thread; the
// reference repo itself only consumes the resulting graph poses
// (lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:148-169).
#include <cmath>
#include <cstring>

#include "lsd_oracle.hpp"

namespace lsdo {

Sim3Tracker::Sim3Tracker(int w, int h) : w0(w), h0(h) {
  const size_t n = (size_t)w * h;
  for (auto *b : {&buf_warped_residual, &buf_warped_dx, &buf_warped_dy, &buf_warped_x, &buf_warped_y, &buf_warped_z, &buf_d,
                  &buf_residual_d, &buf_idepthVar, &buf_warped_idepthVar, &buf_weight_p, &buf_weight_d})
    b->assign(n, 0.0f);
}

// Eigen::Quaternionf::setFromTwoVectors(a, b).toRotationMatrix() for non-antiparallel unit-length-able vectors
static Mat3<float> rotationFromTwoVectors(const Vec3<float> &a, const Vec3<float> &b) {
  const float na = a.norm(), nb = b.norm();
  const Vec3<float> v0(a.x / na, a.y / na, a.z / na), v1(b.x / nb, b.y / nb, b.z / nb);
  const float c = v1.dot(v0);
  if (c < -1.0f + 1e-5f) {
    // antiparallel (camera looking backwards): any axis orthogonal to v0; pick via cross with the x / y axis.
    // DECISION: upstream uses an SVD here; this branch is unreachable for trackable poses.
    Vec3<float> ax = v0.cross(Vec3<float>(1, 0, 0));
    if (ax.norm() < 1e-3f) ax = v0.cross(Vec3<float>(0, 1, 0));
    const float n = ax.norm();
    return Quat<float>(0.0f, ax.x / n, ax.y / n, ax.z / n).toRotationMatrix();
  }
  const Vec3<float> axis = v0.cross(v1);
  const float s = sqrtf((1.0f + c) * 2.0f);
  const float invs = 1.0f / s;
  return Quat<float>(s * 0.5f, axis.x * invs, axis.y * invs, axis.z * invs).toRotationMatrix();
}

// Sim3Tracker::calcSim3Buffers
void Sim3Tracker::calcSim3Buffers(TrackingReference *reference, Frame *frame, const Sim3<double> &referenceToFrame, int level) {
  const int w = frame->w[level], h = frame->h[level];
  const float fx_l = frame->fx[level], fy_l = frame->fy[level], cx_l = frame->cx[level], cy_l = frame->cy[level];
  const Mat3<double> Rd = referenceToFrame.rotationMatrix();
  Mat3<float> rotMat, rotMatUnscaled;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      rotMat.m[i][j] = (float)(Rd.m[i][j] * referenceToFrame.s);  // rxso3().matrix().cast<float>()
      rotMatUnscaled.m[i][j] = (float)Rd.m[i][j];
    }
  const Vec3<float> transVec((float)referenceToFrame.t.x, (float)referenceToFrame.t.y, (float)referenceToFrame.t.z);

  // rotation around the optical axis, for rotating the reference gradients into the frame (ESM)
  const Vec3<float> forwardVector(0, 0, -1);
  const Vec3<float> rotatedForwardVector = rotMatUnscaled * forwardVector;
  const Mat3<float> rollMat = rotationFromTwoVectors(rotatedForwardVector, forwardVector) * rotMatUnscaled;
  const float xRoll0 = rollMat.m[0][0], xRoll1 = rollMat.m[0][1], yRoll0 = rollMat.m[1][0], yRoll1 = rollMat.m[1][1];

  frame->requireGradients(level);
  frame->requireIDepth(level);
  const float *refPoint = reference->posData[level].data();
  const float *refColVar = reference->colorAndVarData[level].data();
  const float *refGrad = reference->gradData[level].data();
  const int refNum = reference->numData[level];
  const float *frame_idepth = frame->idepth[level].data();
  const float *frame_idepthVar = frame->idepthVar[level].data();
  const float *frame_gradients = frame->grad[level].data();

  const bool ex = (mode == ReduceMode::EXACT);
  Acc sxx_(ex), syy_(ex), sx_(ex), sy_(ex), sw_(ex), usageCount(ex);
  int idx = 0;
  for (int i = 0; i < refNum; i++) {
    const Vec3<float> p(refPoint[3 * i], refPoint[3 * i + 1], refPoint[3 * i + 2]);
    const Vec3<float> Wxp = rotMat * p + transVec;
    const float u_new = (Wxp.x / Wxp.z) * fx_l + cx_l;
    const float v_new = (Wxp.y / Wxp.z) * fy_l + cy_l;
    if (!(u_new > 1 && v_new > 1 && u_new < w - 2 && v_new < h - 2)) continue;
    buf_warped_x[idx] = Wxp.x;
    buf_warped_y[idx] = Wxp.y;
    buf_warped_z[idx] = Wxp.z;
    float resInterp[3];
    getInterpolatedElement4N(frame_gradients, u_new, v_new, w, 3, resInterp);
    // USE_ESM_TRACKING == 1
    const float rotatedGradX = xRoll0 * refGrad[2 * i] + xRoll1 * refGrad[2 * i + 1];
    const float rotatedGradY = yRoll0 * refGrad[2 * i] + yRoll1 * refGrad[2 * i + 1];
    buf_warped_dx[idx] = fx_l * 0.5f * (resInterp[0] + rotatedGradX);
    buf_warped_dy[idx] = fy_l * 0.5f * (resInterp[1] + rotatedGradY);

    const float c1 = affineEstimation_a * refColVar[2 * i] + affineEstimation_b;
    const float c2 = resInterp[2];
    const float residual_p = c1 - c2;
    const float weight = fabsf(residual_p) < 2.0f ? 1 : 2.0f / fabsf(residual_p);
    sxx_.add(c1 * c1 * weight);
    syy_.add(c2 * c2 * weight);
    sx_.add(c1 * weight);
    sy_.add(c2 * weight);
    sw_.add(weight);
    buf_warped_residual[idx] = residual_p;
    buf_idepthVar[idx] = refColVar[2 * i + 1];

    // Sim3 only: depth residual against the frame's own (nearest-pixel) inverse depth
    const int idx_rounded = (int)(u_new + 0.5f) + w * (int)(v_new + 0.5f);
    const float var_frameDepth = frame_idepthVar[idx_rounded];
    const float ref_idepth = 1.0f / Wxp.z;
    buf_d[idx] = 1.0f / p.z;
    if (var_frameDepth > 0) {
      buf_residual_d[idx] = ref_idepth - frame_idepth[idx_rounded];
      buf_warped_idepthVar[idx] = var_frameDepth;
    } else {
      buf_residual_d[idx] = -1;
      buf_warped_idepthVar[idx] = -1;
    }
    idx++;
    const float depthChange = p.z / Wxp.z;
    usageCount.add(depthChange < 1 ? depthChange : 1);
  }
  buf_warped_size = idx;
  pointUsage = usageCount.get() / (float)refNum;
  if (ex) {
    const double Sxx = sxx_.d, Syy = syy_.d, Sx = sx_.d, Sy = sy_.d, Sw = sw_.d;
    const double a = std::sqrt((Syy - Sy * Sy / Sw) / (Sxx - Sx * Sx / Sw));
    affineEstimation_a_lastIt = (float)a;
    affineEstimation_b_lastIt = (float)((Sy - a * Sx) / Sw);
  } else {
    const float sxx = sxx_.get(), syy = syy_.get(), sx = sx_.get(), sy = sy_.get(), sw = sw_.get();
    affineEstimation_a_lastIt = sqrtf((syy - sy * sy / sw) / (sxx - sx * sx / sw));
    affineEstimation_b_lastIt = (sy - affineEstimation_a_lastIt * sx) / sw;
  }
}

// Sim3Tracker::calcSim3WeightsAndResidual
Sim3ResidualStruct Sim3Tracker::calcSim3WeightsAndResidual(const Sim3<double> &referenceToFrame) {
  const float tx = (float)referenceToFrame.t.x, ty = (float)referenceToFrame.t.y, tz = (float)referenceToFrame.t.z;
  Sim3ResidualStruct sumRes;
  const bool ex = (mode == ReduceMode::EXACT);
  Acc sumD(ex), sumP(ex);
  for (int i = 0; i < buf_warped_size; i++) {
    const float px = buf_warped_x[i], py = buf_warped_y[i], pz = buf_warped_z[i];
    const float d = buf_d[i];
    const float rp = buf_warped_residual[i], rd = buf_residual_d[i];
    const float gx = buf_warped_dx[i], gy = buf_warped_dy[i];
    const float s = settings.var_weight * buf_idepthVar[i];
    const float sv = settings.var_weight * buf_warped_idepthVar[i];
    const float g0 = (tx * pz - tz * px) / (pz * pz * d);
    const float g1 = (ty * pz - tz * py) / (pz * pz * d);
    const float g2 = (pz - tz) / (pz * pz * d);
    const float drpdd = gx * g0 + gy * g1;
    const float w_p = 1.0f / (CAMERA_PIXEL_NOISE2 + s * drpdd * drpdd);
    const float w_d = 1.0f / (sv + g2 * g2 * s);
    const float weighted_rd = fabsf(rd * sqrtf(w_d));
    const float weighted_rp = fabsf(rp * sqrtf(w_p));
    const float weighted_abs_res = sv > 0 ? weighted_rd + weighted_rp : weighted_rp;
    const float wh = fabsf(weighted_abs_res < settings.huber_d ? 1 : settings.huber_d / weighted_abs_res);
    if (sv > 0) {
      sumD.add(wh * w_d * rd * rd);
      sumRes.numTermsD++;
    }
    sumP.add(wh * w_p * rp * rp);
    sumRes.numTermsP++;
    buf_weight_p[i] = wh * w_p;
    buf_weight_d[i] = sv > 0 ? wh * w_d : 0;
  }
  sumRes.sumResD = sumD.get();
  sumRes.sumResP = sumP.get();
  sumRes.mean = (sumRes.sumResD + sumRes.sumResP) / (sumRes.numTermsD + sumRes.numTermsP);
  sumRes.meanD = sumRes.sumResD / sumRes.numTermsD;
  sumRes.meanP = sumRes.sumResP / sumRes.numTermsP;
  return sumRes;
}

// Sim3Tracker::calcSim3LGS: LGS6 on (v, r_p, w_p), LGS4 on (v4, r_d, w_d), finishNoDivide, LGS7::initializeFrom.
// b holds ls7.b (= -sum J r w); numConstraints = ls6.num_constraints + ls4.num_constraints = 2 * buf_warped_size.
void Sim3Tracker::calcSim3LGS(float A[7][7], float b[7], int *numConstraints) {
  const bool ex = (mode == ReduceMode::EXACT);
  float A6[21] = {0}, b6[6] = {0}, A4[10] = {0}, b4[4] = {0};
  double A6d[21] = {0}, b6d[6] = {0}, A4d[10] = {0}, b4d[4] = {0};
  for (int i = 0; i < buf_warped_size; i++) {
    const float px = buf_warped_x[i], py = buf_warped_y[i], pz = buf_warped_z[i];
    const float wp = buf_weight_p[i], wd = buf_weight_d[i];
    const float rp = buf_warped_residual[i], rd = buf_residual_d[i];
    const float gx = buf_warped_dx[i], gy = buf_warped_dy[i];
    const float z = 1.0f / pz;
    const float z_sqr = 1.0f / (pz * pz);
    float v[6], v4[4];
    v[0] = z * gx + 0;
    v[1] = 0 + z * gy;
    v[2] = (-px * z_sqr) * gx + (-py * z_sqr) * gy;
    v[3] = (float)((double)((-px * py * z_sqr) * gx) + (-(1.0 + (double)(py * py * z_sqr))) * (double)gy);
    v[4] = (float)((1.0 + (double)(px * px * z_sqr)) * (double)gx + (double)((px * py * z_sqr) * gy));
    v[5] = (-py * z) * gx + (px * z) * gy;
    v4[0] = z_sqr;
    v4[1] = z_sqr * py;
    v4[2] = -z_sqr * px;
    v4[3] = z;
    int k = 0;
    const float rpw = rp * wp, rdw = rd * wd;
    for (int a = 0; a < 6; a++) {
      const float wa = v[a] * wp;
      for (int c = a; c < 6; c++, k++) {
        if (ex) A6d[k] += (double)(wa * v[c]); else A6[k] += wa * v[c];
      }
      if (ex) b6d[a] -= (double)(v[a] * rpw); else b6[a] -= v[a] * rpw;
    }
    k = 0;
    for (int a = 0; a < 4; a++) {
      const float wa = v4[a] * wd;
      for (int c = a; c < 4; c++, k++) {
        if (ex) A4d[k] += (double)(wa * v4[c]); else A4[k] += wa * v4[c];
      }
      if (ex) b4d[a] -= (double)(v4[a] * rdw); else b4[a] -= v4[a] * rdw;
    }
  }
  if (ex) {
    for (int k = 0; k < 21; k++) A6[k] = (float)A6d[k];
    for (int k = 0; k < 6; k++) b6[k] = (float)b6d[k];
    for (int k = 0; k < 10; k++) A4[k] = (float)A4d[k];
    for (int k = 0; k < 4; k++) b4[k] = (float)b4d[k];
  }
  std::memset(A, 0, sizeof(float) * 49);
  std::memset(b, 0, sizeof(float) * 7);
  int k = 0;
  for (int a = 0; a < 6; a++) {
    for (int c = a; c < 6; c++, k++) A[a][c] = A[c][a] = A6[k];
    b[a] = b6[a];
  }
  const int remap[4] = {2, 3, 4, 6};
  k = 0;
  for (int a = 0; a < 4; a++) {
    for (int c = a; c < 4; c++, k++) {
      A[remap[a]][remap[c]] += A4[k];
      if (c != a) A[remap[c]][remap[a]] += A4[k];
    }
    b[remap[a]] += b4[a];
  }
  *numConstraints = 2 * buf_warped_size;
}

// Sim3Tracker::trackFrameSim3
Sim3<double> Sim3Tracker::trackFrameSim3(TrackingReference *reference, Frame *frame, const Sim3<double> &frameToReference_initialEstimate,
                                         int startLevel, int finalLevel) {
  diverged = false;
  affineEstimation_a = 1;
  affineEstimation_b = 0;
  trace.clear();
  Sim3<double> referenceToFrame = frameToReference_initialEstimate.inverse();
  float A7[7][7], b7[7];
  int nc = 0;
  std::memset(A7, 0, sizeof(A7));
  std::memset(b7, 0, sizeof(b7));
  Sim3ResidualStruct finalResidual;
  bool warp_update_up_to_date = false;
  auto tooFew = [&](int lvl) {
    return buf_warped_size < 0.5 * MIN_GOODPERALL_PIXEL_ABSMIN * (w0 >> lvl) * (h0 >> lvl) || buf_warped_size < 10;
  };

  for (int lvl = startLevel; lvl >= finalLevel; lvl--) {
    if (settings.maxItsPerLvl[lvl] == 0) continue;
    reference->makePointCloud(lvl);
    calcSim3Buffers(reference, frame, referenceToFrame, lvl);
    if (tooFew(lvl)) {
      diverged = true;
      return Sim3<double>();
    }
    Sim3ResidualStruct lastErr = calcSim3WeightsAndResidual(referenceToFrame);
    trace.push_back({lvl, -1, lastErr.mean, 0.0f, buf_warped_size});
    affineEstimation_a = affineEstimation_a_lastIt;
    affineEstimation_b = affineEstimation_b_lastIt;
    float LM_lambda = settings.lambdaInitial[lvl];
    warp_update_up_to_date = false;
    for (int iteration = 0; iteration < settings.maxItsPerLvl[lvl]; iteration++) {
      calcSim3LGS(A7, b7, &nc);
      warp_update_up_to_date = true;
      int incTry = 0;
      while (true) {
        float Al[7][7], rhs[7], inc[7];
        for (int i = 0; i < 7; i++) {
          rhs[i] = -b7[i] / nc;
          for (int j = 0; j < 7; j++) Al[i][j] = A7[i][j] / nc;
        }
        for (int i = 0; i < 7; i++) Al[i][i] *= 1 + LM_lambda;
        ldlt_solve<float, 7>(Al, rhs, inc);
        incTry++;
        float absInc = 0;
        for (int i = 0; i < 7; i++) absInc += inc[i] * inc[i];
        if (!(absInc >= 0 && absInc < 1)) {  // tracking diverged
          std::memset(lastSim3Hessian, 0, sizeof(lastSim3Hessian));
          return Sim3<double>();
        }
        double incd[7];
        for (int i = 0; i < 7; i++) incd[i] = (double)inc[i];
        const Sim3<double> new_referenceToFrame = Sim3<double>::exp(incd) * referenceToFrame;
        calcSim3Buffers(reference, frame, new_referenceToFrame, lvl);
        if (tooFew(lvl)) {
          diverged = true;
          return Sim3<double>();
        }
        const Sim3ResidualStruct error = calcSim3WeightsAndResidual(new_referenceToFrame);
        if (error.mean < lastErr.mean) {
          trace.push_back({lvl, 1, error.mean, LM_lambda, buf_warped_size});
          referenceToFrame = new_referenceToFrame;
          warp_update_up_to_date = false;
          affineEstimation_a = affineEstimation_a_lastIt;
          affineEstimation_b = affineEstimation_b_lastIt;
          if (error.mean / lastErr.mean > settings.convergenceEps[lvl]) iteration = settings.maxItsPerLvl[lvl];
          finalResidual = lastErr = error;
          if (LM_lambda <= 0.2f) LM_lambda = 0; else LM_lambda *= settings.lambdaSuccessFac;
          break;
        } else {
          trace.push_back({lvl, 0, error.mean, LM_lambda, buf_warped_size});
          if (!(absInc > settings.stepSizeMin[lvl])) {
            iteration = settings.maxItsPerLvl[lvl];
            break;
          }
          if (LM_lambda == 0) LM_lambda = 0.2f; else LM_lambda *= std::pow(settings.lambdaFailFac, (float)incTry);
        }
      }
    }
  }

  if (!warp_update_up_to_date) {
    reference->makePointCloud(finalLevel);
    calcSim3Buffers(reference, frame, referenceToFrame, finalLevel);
    finalResidual = calcSim3WeightsAndResidual(referenceToFrame);
    calcSim3LGS(A7, b7, &nc);
  }
  std::memcpy(lastSim3Hessian, A7, sizeof(A7));
  if (referenceToFrame.s <= 0) {
    diverged = true;
    return Sim3<double>();
  }
  lastResidual = finalResidual.mean;
  lastDepthResidual = finalResidual.meanD;
  lastPhotometricResidual = finalResidual.meanP;
  return referenceToFrame.inverse();
}

}  // namespace lsdo
