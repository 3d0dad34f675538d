// oracle/tracking.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Restates upstream Tracking/TrackingReference.cpp, Tracking/SE3Tracker.cpp and
// Tracking/least_squares.cpp (lsd-slam core, un-vendored) per SURVEY.md 3.3, A.2, A.3.
#include <cmath>
#include <cstring>

#include "lsd_oracle.hpp"

namespace lsdo {

// ---------------------------------------------------------------------------------------
// TrackingReference
// ---------------------------------------------------------------------------------------
void TrackingReference::importFrame(Frame *kf) {
  keyframe = kf;
  frameID = kf->id;
  invalidate();
}

void TrackingReference::invalidate() {
  for (int l = 0; l < NL; l++) numData[l] = 0;
}

// makePointCloud: x outer, y inner (column-major emission order), skip var<=0 || idepth==0.
void TrackingReference::makePointCloud(int level) {
  if (numData[level] > 0) return;
  Frame *kf = keyframe;
  kf->requireIDepth(level);
  kf->requireGradients(level);
  const int W = kf->w[level], H = kf->h[level];
  const float fxi = kf->fxi[level], fyi = kf->fyi[level], cxi = kf->cxi[level], cyi = kf->cyi[level];
  const float *id = kf->idepth[level].data();
  const float *var = kf->idepthVar[level].data();
  const float *col = kf->image[level].data();
  const float *g = kf->grad[level].data();
  posData[level].resize((size_t)3 * W * H);
  gradData[level].resize((size_t)2 * W * H);
  colorAndVarData[level].resize((size_t)2 * W * H);
  pointPosInXYGrid[level].resize((size_t)W * H);
  int n = 0;
  for (int x = 1; x < W - 1; x++)
    for (int y = 1; y < H - 1; y++) {
      const int idx = x + y * W;
      if (var[idx] <= 0 || id[idx] == 0) continue;
      const float inv = 1.0f / id[idx];
      posData[level][3 * n + 0] = inv * (fxi * x + cxi);
      posData[level][3 * n + 1] = inv * (fyi * y + cyi);
      posData[level][3 * n + 2] = inv * 1.0f;
      gradData[level][2 * n + 0] = g[4 * idx + 0];
      gradData[level][2 * n + 1] = g[4 * idx + 1];
      colorAndVarData[level][2 * n + 0] = col[idx];
      colorAndVarData[level][2 * n + 1] = var[idx];
      pointPosInXYGrid[level][n] = idx;
      n++;
    }
  numData[level] = n;
}

// ---------------------------------------------------------------------------------------
// SE3Tracker
// ---------------------------------------------------------------------------------------
SE3Tracker::SE3Tracker(int w, int h) : w0(w), h0(h) {
  const size_t n = (size_t)w * h;
  for (auto *b : {&buf_warped_residual, &buf_warped_dx, &buf_warped_dy, &buf_warped_x, &buf_warped_y, &buf_warped_z,
                  &buf_d, &buf_idepthVar, &buf_weight_p})
    b->assign(n, 0.0f);
}

// SE3Tracker::calcResidualAndBuffers (scalar variant; the SSE build calls the same body).
float SE3Tracker::calcResidualAndBuffers(const float *refPoint, const float *refColVar, const int *idxBuf, int refNum,
                                         Frame *frame, const SE3<float> &referenceToFrame, int level) {
  const int w = frame->w[level], h = frame->h[level];
  const float fx_l = frame->fx[level], fy_l = frame->fy[level], cx_l = frame->cx[level], cy_l = frame->cy[level];
  const Mat3<float> rotMat = referenceToFrame.rotationMatrix();
  const Vec3<float> transVec = referenceToFrame.t;
  frame->requireGradients(level);
  const float *frame_gradients = frame->grad[level].data();

  int idx = 0;
  const bool ex = (mode == ReduceMode::EXACT);
  Acc sumResUnweighted(ex);
  uint8_t *isGoodOutBuffer = idxBuf != nullptr ? frame->refPixelWasGoodBuf() : nullptr;
  int goodCount = 0, badCount = 0;
  Acc sumSignedRes(ex);
  Acc sxx_(ex), syy_(ex), sx_(ex), sy_(ex), sw_(ex);
  Acc usageCount(ex);

  for (int i = 0; i < refNum; i++) {
    const Vec3<float> p(refPoint[3 * i], refPoint[3 * i + 1], refPoint[3 * i + 2]);
    const Vec3<float> Wxp = rotMat * p + transVec;
    const float u_new = (Wxp.x / Wxp.z) * fx_l + cx_l;
    const float v_new = (Wxp.y / Wxp.z) * fy_l + cy_l;
    // inverse test to exclude NaNs
    if (!(u_new > 1 && v_new > 1 && u_new < w - 2 && v_new < h - 2)) {
      if (isGoodOutBuffer) isGoodOutBuffer[idxBuf[i]] = 0;
      continue;
    }
    float resInterp[3];
    getInterpolatedElement4N(frame_gradients, u_new, v_new, w, 3, resInterp);
    const float c1 = affineEstimation_a * refColVar[2 * i] + affineEstimation_b;
    const float c2 = resInterp[2];
    const float residual = c1 - c2;
    const float weight = fabsf(residual) < 5.0f ? 1 : 5.0f / fabsf(residual);
    sxx_.add(c1 * c1 * weight);
    syy_.add(c2 * c2 * weight);
    sx_.add(c1 * weight);
    sy_.add(c2 * weight);
    sw_.add(weight);
    const bool isGood = residual * residual /
                            (MAX_DIFF_CONSTANT + MAX_DIFF_GRAD_MULT * (resInterp[0] * resInterp[0] + resInterp[1] * resInterp[1])) <
                        1;
    if (isGoodOutBuffer) isGoodOutBuffer[idxBuf[i]] = isGood;
    buf_warped_x[idx] = Wxp.x;
    buf_warped_y[idx] = Wxp.y;
    buf_warped_z[idx] = Wxp.z;
    buf_warped_dx[idx] = fx_l * resInterp[0];
    buf_warped_dy[idx] = fy_l * resInterp[1];
    buf_warped_residual[idx] = residual;
    buf_d[idx] = 1.0f / p.z;
    buf_idepthVar[idx] = refColVar[2 * i + 1];
    idx++;
    if (isGood) {
      sumResUnweighted.add(residual * residual);
      sumSignedRes.add(residual);
      goodCount++;
    } else {
      badCount++;
    }
    const float depthChange = p.z / Wxp.z;  // larger depth => pixel "smaller" => count it less
    usageCount.add(depthChange < 1 ? depthChange : 1);
  }
  const float sxx = sxx_.get(), syy = syy_.get(), sx = sx_.get(), sy = sy_.get(), sw = sw_.get();
  buf_warped_size = idx;
  pointUsage = usageCount.get() / (float)refNum;
  lastGoodCount = goodCount;
  lastBadCount = badCount;
  lastMeanRes = sumSignedRes.get() / goodCount;
  if (ex) {
    // EXACT: the closed form cancels ~20x (a) / ~100x (b); evaluate it in fp64 from the fp64 sums, round once
    const double Sxx = sxx_.d, Syy = syy_.d, Sx = sx_.d, Sy = sy_.d, Sw = sw_.d;
    const double a = std::sqrt((Syy - Sy * Sy / Sw) / (Sxx - Sx * Sx / Sw));
    affineEstimation_a_lastIt = (float)a;
    affineEstimation_b_lastIt = (float)((Sy - a * Sx) / Sw);
  } else {
    affineEstimation_a_lastIt = sqrtf((syy - sy * sy / sw) / (sxx - sx * sx / sw));
    affineEstimation_b_lastIt = (sy - affineEstimation_a_lastIt * sx) / sw;
  }
  return sumResUnweighted.get() / goodCount;
}

// SE3Tracker::calcWeightsAndResidual.  ReduceMode::SSE4 mirrors the lane order of the SSE
// build (4 interleaved partial sums over i&~3, scalar tail, lanes summed 0+1+2+3).
// NOTE: upstream's SSE body additionally uses RCPPS approximations; only the ORDER is mirrored.
float SE3Tracker::calcWeightsAndResidual(const SE3<float> &referenceToFrame) {
  const float tx = referenceToFrame.t.x, ty = referenceToFrame.t.y, tz = referenceToFrame.t.z;
  float lanes[4] = {0, 0, 0, 0};
  float sumRes = 0;
  double sumResD = 0;
  const int n = buf_warped_size;
  const int n4 = (mode == ReduceMode::SSE4) ? (n & ~3) : 0;
  for (int i = 0; i < n; i++) {
    const float px = buf_warped_x[i], py = buf_warped_y[i], pz = buf_warped_z[i];
    const float d = buf_d[i];
    const float rp = buf_warped_residual[i];
    const float gx = buf_warped_dx[i], gy = buf_warped_dy[i];
    const float s = settings.var_weight * buf_idepthVar[i];
    const float g0 = (tx * pz - tz * px) / (pz * pz * d);
    const float g1 = (ty * pz - tz * py) / (pz * pz * d);
    const float drpdd = gx * g0 + gy * g1;
    const float w_p = 1.0f / (CAMERA_PIXEL_NOISE2 + s * drpdd * drpdd);
    const float weighted_rp = fabsf(rp * sqrtf(w_p));
    const float wh = fabsf(weighted_rp < (settings.huber_d / 2) ? 1 : (settings.huber_d / 2) / weighted_rp);
    const float term = wh * w_p * rp * rp;
    if (mode == ReduceMode::EXACT) sumResD += (double)term;
    else if (i < n4) lanes[i & 3] += term; else sumRes += term;
    buf_weight_p[i] = wh * w_p;
  }
  if (mode == ReduceMode::EXACT) sumRes = (float)sumResD;
  if (mode == ReduceMode::SSE4) sumRes = (((lanes[0] + lanes[1]) + lanes[2]) + lanes[3]) + sumRes;
  return sumRes / n;
}

// SE3Tracker::calculateWarpUpdate + NormalEquationsLeastSquares::{update,finish}.
// b accumulates -J*(res*weight) so that the caller solves A*inc = -b (see trackFrame).
void SE3Tracker::calculateWarpUpdate(float A[6][6], float b[6], float *error) {
  const int n = buf_warped_size;
  const int nl = (mode == ReduceMode::SSE4) ? 4 : 1;
  const int n4 = (mode == ReduceMode::SSE4) ? (n & ~3) : 0;
  float Aacc[5][21], bacc[5][6], eacc[5];
  double Ad[21] = {0}, bd[6] = {0}, ed = 0;
  const bool ex = (mode == ReduceMode::EXACT);
  std::memset(Aacc, 0, sizeof(Aacc));
  std::memset(bacc, 0, sizeof(bacc));
  std::memset(eacc, 0, sizeof(eacc));
  for (int i = 0; i < n; i++) {
    const float px = buf_warped_x[i], py = buf_warped_y[i], pz = buf_warped_z[i];
    const float r = buf_warped_residual[i];
    const float gx = buf_warped_dx[i], gy = buf_warped_dy[i];
    const float z = 1.0f / pz;
    const float z_sqr = 1.0f / (pz * pz);
    float v[6];
    v[0] = z * gx + 0;
    v[1] = 0 + z * gy;
    v[2] = (-px * z_sqr) * gx + (-py * z_sqr) * gy;
    // the `1.0 +` literals are doubles upstream: these two rows are evaluated in fp64 and rounded once
    v[3] = (float)((double)((-px * py * z_sqr) * gx) + (-(1.0 + (double)(py * py * z_sqr))) * (double)gy);
    v[4] = (float)((1.0 + (double)(px * px * z_sqr)) * (double)gx + (double)((px * py * z_sqr) * gy));
    v[5] = (-py * z) * gx + (px * z) * gy;
    const float wgt = buf_weight_p[i];
    const int lane = (i < n4) ? (i & 3) : (nl == 4 ? 4 : 0);
    int k = 0;
    const float rw = r * wgt;
    if (ex) {
      for (int a = 0; a < 6; a++) {
        const float wa = v[a] * wgt;
        for (int c = a; c < 6; c++) Ad[k++] += (double)(wa * v[c]);
      }
      for (int a = 0; a < 6; a++) bd[a] -= (double)(v[a] * rw);
      ed += (double)(r * r * wgt);
      continue;
    }
    for (int a = 0; a < 6; a++) {
      const float wa = v[a] * wgt;
      for (int c = a; c < 6; c++) Aacc[lane][k++] += wa * v[c];
    }
    for (int a = 0; a < 6; a++) bacc[lane][a] -= v[a] * rw;
    eacc[lane] += r * r * wgt;
  }
  if (ex) {
    for (int k = 0; k < 21; k++) Aacc[0][k] = (float)Ad[k];
    for (int k = 0; k < 6; k++) bacc[0][k] = (float)bd[k];
    eacc[0] = (float)ed;
  }
  float Asum[21], bsum[6], esum;
  if (nl == 4) {
    for (int k = 0; k < 21; k++) Asum[k] = (((Aacc[0][k] + Aacc[1][k]) + Aacc[2][k]) + Aacc[3][k]) + Aacc[4][k];
    for (int k = 0; k < 6; k++) bsum[k] = (((bacc[0][k] + bacc[1][k]) + bacc[2][k]) + bacc[3][k]) + bacc[4][k];
    esum = (((eacc[0] + eacc[1]) + eacc[2]) + eacc[3]) + eacc[4];
  } else {
    for (int k = 0; k < 21; k++) Asum[k] = Aacc[0][k];
    for (int k = 0; k < 6; k++) bsum[k] = bacc[0][k];
    esum = eacc[0];
  }
  const float nf = (float)n;
  int k = 0;
  for (int a = 0; a < 6; a++)
    for (int c = a; c < 6; c++) {
      A[a][c] = A[c][a] = Asum[k++] / nf;
    }
  for (int a = 0; a < 6; a++) b[a] = bsum[a] / nf;
  *error = esum / nf;
}

// SE3Tracker::trackFrame -- coarse-to-fine Levenberg-Marquardt.
SE3<double> SE3Tracker::trackFrame(TrackingReference *reference, Frame *frame,
                                   const SE3<double> &frameToReference_initialEstimate) {
  diverged = false;
  trackingWasGood = true;
  affineEstimation_a = 1;
  affineEstimation_b = 0;
  trace.clear();

  SE3<float> referenceToFrame = frameToReference_initialEstimate.inverse().cast<float>();
  for (int l = 0; l < NL; l++) numCalcResidualCalls[l] = numCalcWarpUpdateCalls[l] = 0;
  float last_residual = 0;

  for (int lvl = SE3TRACKING_MAX_LEVEL - 1; lvl >= SE3TRACKING_MIN_LEVEL; lvl--) {
    reference->makePointCloud(lvl);
    const float *pos = reference->posData[lvl].data();
    const float *cv = reference->colorAndVarData[lvl].data();
    const int *idxb = (lvl == SE3TRACKING_MIN_LEVEL) ? reference->pointPosInXYGrid[lvl].data() : nullptr;
    const int num = reference->numData[lvl];

    calcResidualAndBuffers(pos, cv, idxb, num, frame, referenceToFrame, lvl);
    if (buf_warped_size < MIN_GOODPERALL_PIXEL_ABSMIN * (w0 >> lvl) * (h0 >> lvl)) {
      diverged = true;
      trackingWasGood = false;
      return SE3<double>();
    }
    affineEstimation_a = affineEstimation_a_lastIt;
    affineEstimation_b = affineEstimation_b_lastIt;
    float lastErr = calcWeightsAndResidual(referenceToFrame);
    numCalcResidualCalls[lvl]++;
    trace.push_back({lvl, -1, lastErr, 0.0f, buf_warped_size});
    float LM_lambda = settings.lambdaInitial[lvl];

    for (int iteration = 0; iteration < settings.maxItsPerLvl[lvl]; iteration++) {
      float A[6][6], b[6], lsErr;
      calculateWarpUpdate(A, b, &lsErr);
      numCalcWarpUpdateCalls[lvl]++;
      iterationNumber = iteration;
      int incTry = 0;
      while (true) {
        float Al[6][6], nb[6], inc[6];
        for (int i = 0; i < 6; i++) {
          nb[i] = -b[i];
          for (int j = 0; j < 6; j++) Al[i][j] = A[i][j];
        }
        for (int i = 0; i < 6; i++) Al[i][i] *= 1 + LM_lambda;
        ldlt_solve<float, 6>(Al, nb, inc);
        incTry++;
        const SE3<float> new_referenceToFrame = SE3<float>::exp(inc) * referenceToFrame;

        calcResidualAndBuffers(pos, cv, idxb, num, frame, new_referenceToFrame, lvl);
        if (buf_warped_size < MIN_GOODPERALL_PIXEL_ABSMIN * (w0 >> lvl) * (h0 >> lvl)) {
          diverged = true;
          trackingWasGood = false;
          return SE3<double>();
        }
        const float error = calcWeightsAndResidual(new_referenceToFrame);
        numCalcResidualCalls[lvl]++;

        if (error < lastErr) {
          trace.push_back({lvl, 1, error, LM_lambda, buf_warped_size});
          referenceToFrame = new_referenceToFrame;
          affineEstimation_a = affineEstimation_a_lastIt;
          affineEstimation_b = affineEstimation_b_lastIt;
          if (error / lastErr > settings.convergenceEps[lvl]) iteration = settings.maxItsPerLvl[lvl];
          last_residual = lastErr = error;
          if (LM_lambda <= 0.2f) LM_lambda = 0; else LM_lambda *= settings.lambdaSuccessFac;
          break;
        } else {
          trace.push_back({lvl, 0, error, LM_lambda, buf_warped_size});
          float inc2 = 0;
          for (int i = 0; i < 6; i++) inc2 += inc[i] * inc[i];
          if (!(inc2 > settings.stepSizeMin[lvl])) {
            iteration = settings.maxItsPerLvl[lvl];
            break;
          }
          if (LM_lambda == 0) LM_lambda = 0.2f; else LM_lambda *= std::pow(settings.lambdaFailFac, (float)incTry);
        }
      }
    }
  }

  lastResidual = last_residual;
  trackingWasGood = !diverged &&
                    lastGoodCount / (frame->w[SE3TRACKING_MIN_LEVEL] * frame->h[SE3TRACKING_MIN_LEVEL]) > MIN_GOODPERALL_PIXEL &&
                    lastGoodCount / (lastGoodCount + lastBadCount) > MIN_GOODPERGOODBAD_PIXEL;
  if (trackingWasGood) reference->keyframe->numFramesTrackedOnThis++;
  frame->initialTrackedResidual = lastResidual / pointUsage;
  const SE3<double> frameToRef = referenceToFrame.inverse().cast<double>();
  frame->thisToParent_raw = sim3FromSE3<double>(frameToRef, 1.0);
  frame->trackingParentId = reference->keyframe->id;
  return frameToRef;
}

// SE3Tracker::checkPermaRefOverlap: mean min(1, z_ref / z') over points projecting inside the image at
// QUICK_KF_CHECK_LVL.
float SE3Tracker::checkPermaRefOverlap(Frame *reference, TrackingReference *permaRef, const SE3<double> &referenceToFrameOrg) {
  const SE3<float> referenceToFrame = referenceToFrameOrg.cast<float>();
  const int lvl = QUICK_KF_CHECK_LVL;
  permaRef->makePointCloud(lvl);
  const int w2 = reference->w[lvl] - 1, h2 = reference->h[lvl] - 1;
  const float fx_l = reference->fx[lvl], fy_l = reference->fy[lvl], cx_l = reference->cx[lvl], cy_l = reference->cy[lvl];
  const Mat3<float> rotMat = referenceToFrame.rotationMatrix();
  const Vec3<float> transVec = referenceToFrame.t;
  const float *pos = permaRef->posData[lvl].data();
  const int n = permaRef->numData[lvl];
  float usageCount = 0;
  for (int i = 0; i < n; i++) {
    const Vec3<float> p(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    const Vec3<float> Wxp = rotMat * p + transVec;
    const float u_new = (Wxp.x / Wxp.z) * fx_l + cx_l;
    const float v_new = (Wxp.y / Wxp.z) * fy_l + cy_l;
    if (u_new > 0 && v_new > 0 && u_new < w2 && v_new < h2) {
      const float depthChange = p.z / Wxp.z;
      usageCount += depthChange < 1 ? depthChange : 1;
    }
  }
  pointUsage = usageCount / (float)n;
  return pointUsage;
}

// SE3Tracker::trackFrameOnPermaref: single-level quick test track (maxIts 5, stepSizeMin 1e-3,
// convergenceEps 0.98 would be set by the caller in settings for QUICK_KF_CHECK_LVL).
SE3<double> SE3Tracker::trackFrameOnPermaref(Frame *reference, TrackingReference *permaRef, Frame *frame,
                                             const SE3<double> &referenceToFrameOrg) {
  (void)reference;
  SE3<float> referenceToFrame = referenceToFrameOrg.cast<float>();
  affineEstimation_a = 1;
  affineEstimation_b = 0;
  diverged = false;
  trackingWasGood = true;
  trace.clear();
  const int lvl = QUICK_KF_CHECK_LVL;
  permaRef->makePointCloud(lvl);
  const float *pos = permaRef->posData[lvl].data();
  const float *cv = permaRef->colorAndVarData[lvl].data();
  const int num = permaRef->numData[lvl];

  calcResidualAndBuffers(pos, cv, nullptr, num, frame, referenceToFrame, lvl);
  if (buf_warped_size < MIN_GOODPERALL_PIXEL_ABSMIN * (w0 >> lvl) * (h0 >> lvl)) {
    diverged = true;
    trackingWasGood = false;
    return SE3<double>();
  }
  affineEstimation_a = affineEstimation_a_lastIt;
  affineEstimation_b = affineEstimation_b_lastIt;
  float lastErr = calcWeightsAndResidual(referenceToFrame);
  trace.push_back({lvl, -1, lastErr, 0.0f, buf_warped_size});
  float LM_lambda = settings.lambdaInitial[lvl];
  for (int iteration = 0; iteration < settings.maxItsPerLvl[lvl]; iteration++) {
    float A[6][6], b[6], lsErr;
    calculateWarpUpdate(A, b, &lsErr);
    int incTry = 0;
    while (true) {
      float Al[6][6], nb[6], inc[6];
      for (int i = 0; i < 6; i++) { nb[i] = -b[i]; for (int j = 0; j < 6; j++) Al[i][j] = A[i][j]; }
      for (int i = 0; i < 6; i++) Al[i][i] *= 1 + LM_lambda;
      ldlt_solve<float, 6>(Al, nb, inc);
      incTry++;
      const SE3<float> new_referenceToFrame = SE3<float>::exp(inc) * referenceToFrame;
      calcResidualAndBuffers(pos, cv, nullptr, num, frame, new_referenceToFrame, lvl);
      if (buf_warped_size < MIN_GOODPERALL_PIXEL_ABSMIN * (w0 >> lvl) * (h0 >> lvl)) {
        diverged = true;
        trackingWasGood = false;
        return SE3<double>();
      }
      const float error = calcWeightsAndResidual(new_referenceToFrame);
      if (error < lastErr) {
        trace.push_back({lvl, 1, error, LM_lambda, buf_warped_size});
        referenceToFrame = new_referenceToFrame;
        affineEstimation_a = affineEstimation_a_lastIt;
        affineEstimation_b = affineEstimation_b_lastIt;
        if (error / lastErr > settings.convergenceEps[lvl]) iteration = settings.maxItsPerLvl[lvl];
        lastErr = error;
        lastResidual = lastErr;
        if (LM_lambda <= 0.2f) LM_lambda = 0; else LM_lambda *= settings.lambdaSuccessFac;
        break;
      } else {
        trace.push_back({lvl, 0, error, LM_lambda, buf_warped_size});
        float inc2 = 0;
        for (int i = 0; i < 6; i++) inc2 += inc[i] * inc[i];
        if (!(inc2 > settings.stepSizeMin[lvl])) { iteration = settings.maxItsPerLvl[lvl]; break; }
        if (LM_lambda == 0) LM_lambda = 0.2f; else LM_lambda *= std::pow(settings.lambdaFailFac, (float)incTry);
      }
    }
  }
  lastResidual = lastErr;
  trackingWasGood = !diverged &&
                    lastGoodCount / (frame->w[QUICK_KF_CHECK_LVL] * frame->h[QUICK_KF_CHECK_LVL]) > MIN_GOODPERALL_PIXEL &&
                    lastGoodCount / (lastGoodCount + lastBadCount) > MIN_GOODPERGOODBAD_PIXEL;
  return referenceToFrame.cast<double>();
}

}  // namespace lsdo
