// oracle/depthmap.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Restates upstream DepthEstimation/DepthMap.cpp + DepthMapPixelHypothesis.cpp (lsd-slam core,
// un-vendored: /root/reference/fips.yml:1-4) per SURVEY.md 3.4, 3.5, 8a C1-C10 and Appendix A.5-A.9.
// The consumer-side contract is evidenced in the reference at
//   lib/GUI.cpp:104-108                              (updateDepthImage: w*h*3 RGB of debugPlotDepthMap)
//   lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:56-79 (idepth / idepthVar left behind by setDepth).
//
// DECISION (documented in DESIGN.md "Deviations from SURVEY"): SURVEY.md 8a-K lists
// VAL_SUM_MIN_FOR_CREATE = -1 and VAL_SUM_MIN_FOR_KEEP = 0.  With -1 every textured invalid pixel without
// a single valid 5x5 neighbour would be "filled" with 0/0 = NaN; the published upstream settings.h values
// are 30 (create) / 24 (keep) / 100 (unblacklist).  Both thresholds are run-time settings of the oracle
// and of the product (DepthSettings / lsd_depth_settings) and default to the published 30 / 24 / 100.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "lsd_oracle.hpp"

namespace lsdo {

DepthMap::DepthMap(int w, int h, float fx_, float fy_, float cx_, float cy_) : width(w), height(h), fx(fx_), fy(fy_), cx(cx_), cy(cy_) {
  fxi = 1.0f / fx;
  fyi = 1.0f / fy;
  cxi = -cx / fx;
  cyi = -cy / fy;
  currentDepthMap.assign((size_t)w * h, Hypothesis());
  otherDepthMap.assign((size_t)w * h, Hypothesis());
  validityIntegralBuffer.assign((size_t)w * h, 0);
}

// IndexThreadReduce::reduce(f, first, end, step): workers pull [y, y+step) chunks.  Results never depend
// on the schedule: every stage writes only its own pixel and reads a snapshot.
void DepthMap::parallelRows(int first, int end, int step, const std::function<void(int, int)> &f) {
  if (numThreads <= 1 || step <= 0) {
    f(first, end);
    return;
  }
  std::atomic<int> next(first);
  std::vector<std::thread> pool;
  for (int t = 0; t < numThreads; t++)
    pool.emplace_back([&]() {
      for (;;) {
        const int y0 = next.fetch_add(step);
        if (y0 >= end) break;
        f(y0, std::min(end, y0 + step));
      }
    });
  for (auto &th : pool) th.join();
}

// DepthMap::initializeFromGTDepth
void DepthMap::initializeFromGTDepth(Frame *new_frame) {
  activeKeyFrame = new_frame;
  activeKeyFrameIsReactivated = false;
  const float *idepth = new_frame->idepth[0].data();
  for (int i = 0; i < width * height; i++) {
    const float v = idepth[i];
    if (!std::isnan(v) && v > 0)
      currentDepthMap[i] = Hypothesis(v, v, VAR_GT_INIT_INITIAL, VAR_GT_INIT_INITIAL, 20);
    else {
      currentDepthMap[i].isValid = false;
      currentDepthMap[i].blacklisted = 0;
    }
  }
  activeKeyFrame->setDepth(currentDepthMap.data());
}

// DepthMap::initializeRandomly: libc rand() consumed in raster order at pixels with maxGrad > MIN_ABS_GRAD_CREATE
void DepthMap::initializeRandomly(Frame *new_frame) {
  activeKeyFrame = new_frame;
  activeKeyFrameIsReactivated = false;
  new_frame->requireMaxGradients(0);
  const float *maxGradients = new_frame->maxGrad[0].data();
  for (int y = 1; y < height - 1; y++)
    for (int x = 1; x < width - 1; x++) {
      Hypothesis &t = currentDepthMap[x + y * width];
      if (maxGradients[x + y * width] > MIN_USE_GRAD) {
        const float idepth = 0.5f + 1.0f * ((rand() % 100001) / 100000.0f);
        t = Hypothesis(idepth, idepth, VAR_RANDOM_INIT_INITIAL, VAR_RANDOM_INIT_INITIAL, 20);
      } else {
        t.isValid = false;
        t.blacklisted = 0;
      }
    }
  activeKeyFrame->setDepth(currentDepthMap.data());
}

// DECISION: inject an explicit hypothesis map (parity runs cannot rely on rand() equivalence)
void DepthMap::initializeFromMap(Frame *kf, const Hypothesis *map) {
  activeKeyFrame = kf;
  activeKeyFrameIsReactivated = false;
  std::copy(map, map + (size_t)width * height, currentDepthMap.begin());
}

// ---------------------------------------------------------------------------------------
// observeDepth (A.5)
// ---------------------------------------------------------------------------------------
void DepthMap::observeDepth() {
  activeKeyFrame->requireMaxGradients(0);
  activeKeyFrame->requireGradients(0);
  parallelRows(3, height - 3, 10, [this](int y0, int y1) { observeDepthRow(y0, y1); });
}

void DepthMap::observeDepthRow(int yMin, int yMax) {
  const float *keyFrameMaxGradBuf = activeKeyFrame->maxGrad[0].data();
  for (int y = yMin; y < yMax; y++)
    for (int x = 3; x < width - 3; x++) {
      const int idx = x + y * width;
      Hypothesis *target = &currentDepthMap[idx];
      const bool hasHypothesis = target->isValid;
      if (hasHypothesis && keyFrameMaxGradBuf[idx] < MIN_USE_GRAD) {  // MIN_ABS_GRAD_DECREASE
        target->isValid = false;
        continue;
      }
      if (keyFrameMaxGradBuf[idx] < MIN_USE_GRAD || target->blacklisted < settings.minBlacklist) continue;
      if (!hasHypothesis)
        observeDepthCreate(x, y, idx);
      else
        observeDepthUpdate(x, y, idx, keyFrameMaxGradBuf);
    }
}

static inline bool trackedMaskRejects(const Frame *refFrame, const Frame *activeKeyFrame, int x, int y, int width) {
  if (refFrame->trackingParentId != activeKeyFrame->id) return false;
  if (refFrame->refPixelWasGood.empty()) return false;  // refPixelWasGoodNoCreate() == 0
  return !refFrame->refPixelWasGood[(x >> SE3TRACKING_MIN_LEVEL) + (width >> SE3TRACKING_MIN_LEVEL) * (y >> SE3TRACKING_MIN_LEVEL)];
}

bool DepthMap::observeDepthCreate(int x, int y, int idx) {
  Hypothesis *target = &currentDepthMap[idx];
  Frame *refFrame = activeKeyFrameIsReactivated ? newest_referenceFrame : oldest_referenceFrame;
  if (trackedMaskRejects(refFrame, activeKeyFrame, x, y, width)) return false;
  float epx, epy;
  if (!makeAndCheckEPL(x, y, refFrame, &epx, &epy)) return false;
  float result_idepth = 0, result_var = 0, result_eplLength = 0;
  const float error = doLineStereo((float)x, (float)y, epx, epy, 0.0f, 1.0f, 1.0f / MIN_DEPTH, refFrame, refFrame->image[0].data(),
                                   result_idepth, result_var, result_eplLength);
  if (error == -3 || error == -2) target->blacklisted--;
  if (error < 0 || result_var > MAX_VAR) return false;
  result_idepth = UNZERO(result_idepth);
  *target = Hypothesis(result_idepth, result_var, VALIDITY_COUNTER_INITIAL_OBSERVE);
  return true;
}

bool DepthMap::observeDepthUpdate(int x, int y, int idx, const float *keyFrameMaxGradBuf) {
  Hypothesis *target = &currentDepthMap[idx];
  Frame *refFrame;
  if (!activeKeyFrameIsReactivated) {
    const int k = (int)target->nextStereoFrameMinID - referenceFrameByID_offset;
    if (k >= (int)referenceFrameByID.size()) return false;
    refFrame = (k < 0) ? oldest_referenceFrame : referenceFrameByID[k];
  } else {
    refFrame = newest_referenceFrame;
  }
  if (trackedMaskRejects(refFrame, activeKeyFrame, x, y, width)) return false;
  float epx, epy;
  if (!makeAndCheckEPL(x, y, refFrame, &epx, &epy)) return false;

  const float sv = sqrtf(target->idepth_var_smoothed);
  float min_idepth = target->idepth_smoothed - sv * STEREO_EPL_VAR_FAC;
  float max_idepth = target->idepth_smoothed + sv * STEREO_EPL_VAR_FAC;
  if (min_idepth < 0) min_idepth = 0;
  if (max_idepth > 1 / MIN_DEPTH) max_idepth = 1 / MIN_DEPTH;

  float result_idepth = 0, result_var = 0, result_eplLength = 0;
  const float error = doLineStereo((float)x, (float)y, epx, epy, min_idepth, target->idepth_smoothed, max_idepth, refFrame,
                                   refFrame->image[0].data(), result_idepth, result_var, result_eplLength);
  const float diff = result_idepth - target->idepth_smoothed;

  if (error == -1) return false;  // out of bounds: try again later
  if (error == -2) {              // not good for stereo
    target->validity_counter -= VALIDITY_COUNTER_DEC;
    if (target->validity_counter < 0) target->validity_counter = 0;
    target->nextStereoFrameMinID = 0;
    target->idepth_var *= FAIL_VAR_INC_FAC;
    if (target->idepth_var > MAX_VAR) {
      target->isValid = false;
      target->blacklisted--;
    }
    return false;
  }
  if (error == -3 || error == -4) return false;
  if (DIFF_FAC_OBSERVE * diff * diff > result_var + target->idepth_var_smoothed) {  // inconsistent
    target->idepth_var *= FAIL_VAR_INC_FAC;
    if (target->idepth_var > MAX_VAR) target->isValid = false;
    return false;
  }
  // textbook EKF update
  float id_var = target->idepth_var * SUCC_VAR_INC_FAC;
  const float w = result_var / (result_var + id_var);
  const float new_idepth = (1 - w) * result_idepth + w * target->idepth;
  target->idepth = UNZERO(new_idepth);
  id_var = id_var * w;
  if (id_var < target->idepth_var) target->idepth_var = id_var;
  target->validity_counter += VALIDITY_COUNTER_INC;
  const float absGrad = keyFrameMaxGradBuf[idx];
  const float cap = VALIDITY_COUNTER_MAX + absGrad * (VALIDITY_COUNTER_MAX_VARIABLE) / 255.0f;
  if (target->validity_counter > cap) target->validity_counter = (int)cap;
  if (result_eplLength < MIN_EPL_LENGTH_CROP) {  // skip ahead
    float inc = activeKeyFrame->numFramesTrackedOnThis / (float)(activeKeyFrame->numMappedOnThis + 5);
    if (inc < 3) inc = 3;
    inc += ((int)(result_eplLength * 10000) % 2);
    if (result_eplLength < 0.5f * MIN_EPL_LENGTH_CROP) inc *= 3;
    target->nextStereoFrameMinID = refFrame->id + inc;
  }
  return true;
}

// makeAndCheckEPL (A.6)
bool DepthMap::makeAndCheckEPL(int x, int y, const Frame *ref, float *pepx, float *pepy) {
  const int idx = x + y * width;
  const float *I = activeKeyFrame->image[0].data();
  const float epx = -fx * ref->thisToOther_t.x + ref->thisToOther_t.z * (x - cx);
  const float epy = -fy * ref->thisToOther_t.y + ref->thisToOther_t.z * (y - cy);
  if (std::isnan(epx + epy)) return false;
  const float eplLengthSquared = epx * epx + epy * epy;
  if (eplLengthSquared < MIN_EPL_LENGTH_SQUARED) return false;
  const float gx = I[idx + 1] - I[idx - 1];
  const float gy = I[idx + width] - I[idx - width];
  float eplGradSquared = gx * epx + gy * epy;
  eplGradSquared = eplGradSquared * eplGradSquared / eplLengthSquared;
  if (eplGradSquared < MIN_EPL_GRAD_SQUARED) return false;
  if (eplGradSquared / (gx * gx + gy * gy) < MIN_EPL_ANGLE_SQUARED) return false;
  const float fac = GRADIENT_SAMPLE_DIST / sqrtf(eplLengthSquared);
  *pepx = epx * fac;
  *pepy = epy * fac;
  return true;
}

// doLineStereo (A.7).  Returns best error (>= 0) or -1 oob / -2 unclear winner / -3 error too large / -4 NaN.
float DepthMap::doLineStereo(float u, float v, float epxn, float epyn, float min_idepth, float prior_idepth, float max_idepth,
                             const Frame *referenceFrame, const float *referenceFrameImage, float &result_idepth,
                             float &result_var, float &result_eplLength) {
  const float *activeKeyFrameImageData = activeKeyFrame->image[0].data();
  const Vec3<float> KinvP(fxi * u + cxi, fyi * v + cyi, 1.0f);
  const Vec3<float> pInf = referenceFrame->K_otherToThis_R * KinvP;
  const Vec3<float> Kt = referenceFrame->K_otherToThis_t;
  const Vec3<float> pReal = pInf / prior_idepth + Kt;
  const float rescaleFactor = pReal.z * prior_idepth;

  const float firstX = u - 2 * epxn * rescaleFactor, firstY = v - 2 * epyn * rescaleFactor;
  const float lastX = u + 2 * epxn * rescaleFactor, lastY = v + 2 * epyn * rescaleFactor;
  if (firstX <= 0 || firstX >= width - 2 || firstY <= 0 || firstY >= height - 2 || lastX <= 0 || lastX >= width - 2 || lastY <= 0 ||
      lastY >= height - 2)
    return -1;
  if (!(rescaleFactor > 0.7f && rescaleFactor < 1.4f)) return -1;

  const float realVal_p1 = getInterpolatedElement(activeKeyFrameImageData, u + epxn * rescaleFactor, v + epyn * rescaleFactor, width);
  const float realVal_m1 = getInterpolatedElement(activeKeyFrameImageData, u - epxn * rescaleFactor, v - epyn * rescaleFactor, width);
  const float realVal = getInterpolatedElement(activeKeyFrameImageData, u, v, width);
  const float realVal_m2 = getInterpolatedElement(activeKeyFrameImageData, u - 2 * epxn * rescaleFactor, v - 2 * epyn * rescaleFactor, width);
  const float realVal_p2 = getInterpolatedElement(activeKeyFrameImageData, u + 2 * epxn * rescaleFactor, v + 2 * epyn * rescaleFactor, width);

  Vec3<float> pClose = pInf + Kt * max_idepth;
  if (pClose.z < 0.001f) {  // assumed close point lies behind the image
    max_idepth = (0.001f - pInf.z) / Kt.z;
    pClose = pInf + Kt * max_idepth;
  }
  pClose = pClose / pClose.z;
  Vec3<float> pFar = pInf + Kt * min_idepth;
  if (pFar.z < 0.001f || max_idepth < min_idepth) return -1;
  pFar = pFar / pFar.z;
  if (std::isnan(pFar.x + pClose.x)) return -4;

  float incx = pClose.x - pFar.x, incy = pClose.y - pFar.y;
  const float eplLength = sqrtf(incx * incx + incy * incy);
  // upstream writes `!eplLength > 0 || std::isinf(eplLength)`: (!eplLength) > 0 is true only for eplLength == 0
  if (eplLength == 0 || std::isinf(eplLength)) return -4;
  if (eplLength > MAX_EPL_LENGTH_CROP) {
    pClose.x = pFar.x + incx * MAX_EPL_LENGTH_CROP / eplLength;
    pClose.y = pFar.y + incy * MAX_EPL_LENGTH_CROP / eplLength;
  }
  incx *= GRADIENT_SAMPLE_DIST / eplLength;
  incy *= GRADIENT_SAMPLE_DIST / eplLength;
  pFar.x -= incx; pFar.y -= incy;
  pClose.x += incx; pClose.y += incy;
  if (eplLength < MIN_EPL_LENGTH_CROP) {
    const float pad = (MIN_EPL_LENGTH_CROP - eplLength) / 2.0f;
    pFar.x -= incx * pad; pFar.y -= incy * pad;
    pClose.x += incx * pad; pClose.y += incy * pad;
  }
  const float B = SAMPLE_POINT_TO_BORDER;
  if (pFar.x <= B || pFar.x >= width - B || pFar.y <= B || pFar.y >= height - B) return -1;
  if (pClose.x <= B || pClose.x >= width - B || pClose.y <= B || pClose.y >= height - B) {
    if (pClose.x <= B) {
      const float toAdd = (B - pClose.x) / incx;
      pClose.x += toAdd * incx; pClose.y += toAdd * incy;
    } else if (pClose.x >= width - B) {
      const float toAdd = (width - B - pClose.x) / incx;
      pClose.x += toAdd * incx; pClose.y += toAdd * incy;
    }
    if (pClose.y <= B) {
      const float toAdd = (B - pClose.y) / incy;
      pClose.x += toAdd * incx; pClose.y += toAdd * incy;
    } else if (pClose.y >= height - B) {
      const float toAdd = (height - B - pClose.y) / incy;
      pClose.x += toAdd * incx; pClose.y += toAdd * incy;
    }
    const float fincx = pClose.x - pFar.x, fincy = pClose.y - pFar.y;
    const float newEplLength = sqrtf(fincx * fincx + fincy * fincy);
    if (pClose.x <= B || pClose.x >= width - B || pClose.y <= B || pClose.y >= height - B || newEplLength < 8.0f) return -1;
  }

  float cpx = pFar.x, cpy = pFar.y;
  float val_cp_m2 = getInterpolatedElement(referenceFrameImage, cpx - 2.0f * incx, cpy - 2.0f * incy, width);
  float val_cp_m1 = getInterpolatedElement(referenceFrameImage, cpx - incx, cpy - incy, width);
  float val_cp = getInterpolatedElement(referenceFrameImage, cpx, cpy, width);
  float val_cp_p1 = getInterpolatedElement(referenceFrameImage, cpx + incx, cpy + incy, width);
  float val_cp_p2;

  int loopCounter = 0;
  float best_match_x = -1, best_match_y = -1;
  float best_match_err = INFINITY, second_best_match_err = INFINITY;  // upstream: float = 1e50 -> +inf
  float best_match_errPre = NAN, best_match_errPost = NAN, best_match_DiffErrPre = NAN, best_match_DiffErrPost = NAN;
  bool bestWasLastLoop = false;
  float eeLast = -1;
  float e1A = NAN, e1B = NAN, e2A = NAN, e2B = NAN, e3A = NAN, e3B = NAN, e4A = NAN, e4B = NAN, e5A = NAN, e5B = NAN;
  int loopCBest = -1, loopCSecond = -1;
  while (((incx < 0) == (cpx > pClose.x) && (incy < 0) == (cpy > pClose.y)) || loopCounter == 0) {
    val_cp_p2 = getInterpolatedElement(referenceFrameImage, cpx + 2 * incx, cpy + 2 * incy, width);
    float ee = 0;
    if (loopCounter % 2 == 0) {
      e1A = val_cp_p2 - realVal_p2; ee += e1A * e1A;
      e2A = val_cp_p1 - realVal_p1; ee += e2A * e2A;
      e3A = val_cp - realVal;       ee += e3A * e3A;
      e4A = val_cp_m1 - realVal_m1; ee += e4A * e4A;
      e5A = val_cp_m2 - realVal_m2; ee += e5A * e5A;
    } else {
      e1B = val_cp_p2 - realVal_p2; ee += e1B * e1B;
      e2B = val_cp_p1 - realVal_p1; ee += e2B * e2B;
      e3B = val_cp - realVal;       ee += e3B * e3B;
      e4B = val_cp_m1 - realVal_m1; ee += e4B * e4B;
      e5B = val_cp_m2 - realVal_m2; ee += e5B * e5B;
    }
    if (ee < best_match_err) {
      second_best_match_err = best_match_err;
      loopCSecond = loopCBest;
      best_match_err = ee;
      loopCBest = loopCounter;
      best_match_errPre = eeLast;
      best_match_DiffErrPre = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
      best_match_errPost = -1;
      best_match_DiffErrPost = -1;
      best_match_x = cpx;
      best_match_y = cpy;
      bestWasLastLoop = true;
    } else {
      if (bestWasLastLoop) {
        best_match_errPost = ee;
        best_match_DiffErrPost = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
        bestWasLastLoop = false;
      }
      if (ee < second_best_match_err) {
        second_best_match_err = ee;
        loopCSecond = loopCounter;
      }
    }
    eeLast = ee;
    val_cp_m2 = val_cp_m1; val_cp_m1 = val_cp; val_cp = val_cp_p1; val_cp_p1 = val_cp_p2;
    cpx += incx;
    cpy += incy;
    loopCounter++;
  }

  if (best_match_err > 4.0f * MAX_ERROR_STEREO) return -3;
  if (std::abs(loopCBest - loopCSecond) > 1.0f && MIN_DISTANCE_ERROR_STEREO * best_match_err > second_best_match_err) return -2;

  bool didSubpixel = false;
  {  // useSubpixelStereo
    const float gradPre_pre = -(best_match_errPre - best_match_DiffErrPre);
    const float gradPre_this = +(best_match_err - best_match_DiffErrPre);
    const float gradPost_this = -(best_match_err - best_match_DiffErrPost);
    const float gradPost_post = +(best_match_errPost - best_match_DiffErrPost);
    bool interpPost = false, interpPre = false;
    if (best_match_errPre < 0 || best_match_errPost < 0) {
      // one is out of bounds: no interpolation
    } else if ((gradPost_this < 0) ^ (gradPre_this < 0)) {
      // zero crossing exactly in between
    } else if ((gradPre_pre < 0) ^ (gradPre_this < 0)) {
      if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
        // two crossings
      } else
        interpPre = true;
    } else if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
      interpPost = true;
    }
    if (interpPre) {
      const float d = gradPre_this / (gradPre_this - gradPre_pre);
      best_match_x -= d * incx;
      best_match_y -= d * incy;
      best_match_err = best_match_err - 2 * d * gradPre_this - (gradPre_pre - gradPre_this) * d * d;
      didSubpixel = true;
    } else if (interpPost) {
      const float d = gradPost_this / (gradPost_this - gradPost_post);
      best_match_x += d * incx;
      best_match_y += d * incy;
      best_match_err = best_match_err + 2 * d * gradPost_this + (gradPost_post - gradPost_this) * d * d;
      didSubpixel = true;
    }
  }

  const float sampleDist = GRADIENT_SAMPLE_DIST * rescaleFactor;
  float gradAlongLine = 0;
  float tmp = realVal_p2 - realVal_p1; gradAlongLine += tmp * tmp;
  tmp = realVal_p1 - realVal;          gradAlongLine += tmp * tmp;
  tmp = realVal - realVal_m1;          gradAlongLine += tmp * tmp;
  tmp = realVal_m1 - realVal_m2;       gradAlongLine += tmp * tmp;
  gradAlongLine /= sampleDist * sampleDist;
  if (best_match_err > MAX_ERROR_STEREO + sqrtf(gradAlongLine) * 20) return -3;

  float idnew_best_match, alpha;
  const Vec3<float> &t = referenceFrame->otherToThis_t;
  if (incx * incx > incy * incy) {
    const float oldX = fxi * best_match_x + cxi;
    const float nominator = (oldX * t.z - t.x);
    const float dot0 = KinvP.dot(referenceFrame->otherToThis_R_row0);
    const float dot2 = KinvP.dot(referenceFrame->otherToThis_R_row2);
    idnew_best_match = (dot0 - oldX * dot2) / nominator;
    alpha = incx * fxi * (dot0 * t.z - dot2 * t.x) / (nominator * nominator);
  } else {
    const float oldY = fyi * best_match_y + cyi;
    const float nominator = (oldY * t.z - t.y);
    const float dot1 = KinvP.dot(referenceFrame->otherToThis_R_row1);
    const float dot2 = KinvP.dot(referenceFrame->otherToThis_R_row2);
    idnew_best_match = (dot1 - oldY * dot2) / nominator;
    alpha = incy * fyi * (dot1 * t.z - dot2 * t.y) / (nominator * nominator);
  }
  // allowNegativeIdepths == true: negative results are kept

  const float photoDispError = 4.0f * CAMERA_PIXEL_NOISE2 / (gradAlongLine + DIVISION_EPS);
  const float trackingErrorFac = 0.25f * (1.0f + referenceFrame->initialTrackedResidual);
  float gradsInterp[2];
  getInterpolatedElement4N(activeKeyFrame->grad[0].data(), u, v, width, 2, gradsInterp);
  float geoDispError = (gradsInterp[0] * epxn + gradsInterp[1] * epyn) + DIVISION_EPS;
  geoDispError = trackingErrorFac * trackingErrorFac * (gradsInterp[0] * gradsInterp[0] + gradsInterp[1] * gradsInterp[1]) /
                 (geoDispError * geoDispError);
  result_var = alpha * alpha * ((didSubpixel ? 0.05f : 0.5f) * sampleDist * sampleDist + geoDispError + photoDispError);
  result_idepth = idnew_best_match;
  result_eplLength = eplLength;
  return best_match_err;
}

// ---------------------------------------------------------------------------------------
// propagateDepth (C7 / A.9): raster-order forward warp with order-dependent merge / occlusion
// ---------------------------------------------------------------------------------------
void DepthMap::propagateDepth(Frame *new_keyframe) {
  for (auto &pt : otherDepthMap) {
    pt.isValid = false;
    pt.blacklisted = 0;
  }
  const SE3<double> oldToNew_SE3 = se3FromSim3(new_keyframe->thisToParent_raw).inverse();
  const Vec3<double> td = oldToNew_SE3.t;
  const Mat3<double> Rd = oldToNew_SE3.rotationMatrix();
  const Vec3<float> trafoInv_t((float)td.x, (float)td.y, (float)td.z);
  Mat3<float> trafoInv_R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) trafoInv_R.m[i][j] = (float)Rd.m[i][j];

  const uint8_t *trackingWasGood =
      (new_keyframe->trackingParentId == activeKeyFrame->id && !new_keyframe->refPixelWasGood.empty()) ? new_keyframe->refPixelWasGood.data()
                                                                                                        : nullptr;
  const float *activeKFImageData = activeKeyFrame->image[0].data();
  new_keyframe->requireMaxGradients(0);
  const float *newKFMaxGrad = new_keyframe->maxGrad[0].data();
  const float *newKFImageData = new_keyframe->image[0].data();

  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const Hypothesis *source = &currentDepthMap[x + y * width];
      if (!source->isValid) continue;
      const Vec3<float> pn = (trafoInv_R * Vec3<float>(x * fxi + cxi, y * fyi + cyi, 1.0f)) / source->idepth_smoothed + trafoInv_t;
      const float new_idepth = 1.0f / pn.z;
      const float u_new = pn.x * new_idepth * fx + cx;
      const float v_new = pn.y * new_idepth * fy + cy;
      if (!(u_new > 2.1f && v_new > 2.1f && u_new < width - 3.1f && v_new < height - 3.1f)) continue;
      const int newIDX = (int)(u_new + 0.5f) + ((int)(v_new + 0.5f)) * width;
      const float destAbsGrad = newKFMaxGrad[newIDX];
      if (trackingWasGood != nullptr) {
        if (!trackingWasGood[(x >> SE3TRACKING_MIN_LEVEL) + (width >> SE3TRACKING_MIN_LEVEL) * (y >> SE3TRACKING_MIN_LEVEL)] ||
            destAbsGrad < MIN_USE_GRAD)
          continue;
      } else {
        const float sourceColor = activeKFImageData[x + y * width];
        const float destColor = getInterpolatedElement(newKFImageData, u_new, v_new, width);
        const float residual = destColor - sourceColor;
        if (residual * residual / (MAX_DIFF_CONSTANT + MAX_DIFF_GRAD_MULT * destAbsGrad * destAbsGrad) > 1.0f || destAbsGrad < MIN_USE_GRAD)
          continue;
      }
      Hypothesis *targetBest = &otherDepthMap[newIDX];
      float idepth_ratio_4 = new_idepth / source->idepth_smoothed;
      idepth_ratio_4 *= idepth_ratio_4;
      idepth_ratio_4 *= idepth_ratio_4;
      const float new_var = idepth_ratio_4 * source->idepth_var;
      if (targetBest->isValid) {  // occlusion: one of the two gets removed
        const float diff = targetBest->idepth - new_idepth;
        if (DIFF_FAC_PROP_MERGE * diff * diff > new_var + targetBest->idepth_var) {
          if (new_idepth < targetBest->idepth) continue;
          targetBest->isValid = false;
        }
      }
      if (!targetBest->isValid) {
        *targetBest = Hypothesis(new_idepth, new_var, source->validity_counter);
      } else {
        const float w = new_var / (targetBest->idepth_var + new_var);
        const float merged_new_idepth = w * targetBest->idepth + (1.0f - w) * new_idepth;
        int merged_validity = source->validity_counter + targetBest->validity_counter;
        if (merged_validity > VALIDITY_COUNTER_MAX + VALIDITY_COUNTER_MAX_VARIABLE) merged_validity = VALIDITY_COUNTER_MAX + VALIDITY_COUNTER_MAX_VARIABLE;
        *targetBest = Hypothesis(merged_new_idepth, 1.0f / (1.0f / targetBest->idepth_var + 1.0f / new_var), merged_validity);
      }
    }
  std::swap(currentDepthMap, otherDepthMap);
}

// ---------------------------------------------------------------------------------------
// regularizeDepthMap (C8)
// ---------------------------------------------------------------------------------------
void DepthMap::regularizeDepthMapRow(bool removeOcclusions, int validityTH, int yMin, int yMax) {
  const int regularize_radius = 2;
  const float regDistVar = REG_DIST_VAR;
  for (int y = yMin; y < yMax; y++)
    for (int x = regularize_radius; x < width - regularize_radius; x++) {
      Hypothesis *dest = &currentDepthMap[x + y * width];
      const Hypothesis *destRead = &otherDepthMap[x + y * width];
      if (!destRead->isValid) continue;
      float sum = 0, val_sum = 0, sumIvar = 0;
      int numOccluding = 0, numNotOccluding = 0;
      for (int dx = -regularize_radius; dx <= regularize_radius; dx++)
        for (int dy = -regularize_radius; dy <= regularize_radius; dy++) {
          const Hypothesis *source = destRead + dx + dy * width;
          if (!source->isValid) continue;
          const float diff = source->idepth - destRead->idepth;
          if (DIFF_FAC_SMOOTHING * diff * diff > source->idepth_var + destRead->idepth_var) {
            if (removeOcclusions && source->idepth > destRead->idepth) numOccluding++;
            continue;
          }
          val_sum += source->validity_counter;
          if (removeOcclusions) numNotOccluding++;
          const float distFac = (float)(dx * dx + dy * dy) * regDistVar;
          const float ivar = 1.0f / (source->idepth_var + distFac);
          sum += source->idepth * ivar;
          sumIvar += ivar;
        }
      if (val_sum < validityTH) {
        dest->isValid = false;
        dest->blacklisted--;
        continue;
      }
      if (removeOcclusions && numOccluding > numNotOccluding) {
        dest->isValid = false;
        continue;
      }
      sum = sum / sumIvar;
      sum = UNZERO(sum);
      dest->idepth_smoothed = sum;
      dest->idepth_var_smoothed = 1.0f / sumIvar;
    }
}

void DepthMap::regularizeDepthMap(bool removeOcclusions, int validityTH) {
  otherDepthMap = currentDepthMap;  // memcpy(otherDepthMap, currentDepthMap)
  parallelRows(2, height - 2, 10,
               [this, removeOcclusions, validityTH](int y0, int y1) { regularizeDepthMapRow(removeOcclusions, validityTH, y0, y1); });
}

// ---------------------------------------------------------------------------------------
// regularizeDepthMapFillHoles (C9)
// ---------------------------------------------------------------------------------------
void DepthMap::buildRegIntegralBuffer() {
  parallelRows(0, height, 10, [this](int y0, int y1) {
    for (int y = y0; y < y1; y++) {
      int s = 0;
      for (int x = 0; x < width; x++) {
        const Hypothesis &p = currentDepthMap[x + y * width];
        if (p.isValid) s += p.validity_counter;
        validityIntegralBuffer[x + y * width] = s;
      }
    }
  });
  const int wh = width * height;
  for (int idx = width; idx < wh; idx++) validityIntegralBuffer[idx] += validityIntegralBuffer[idx - width];
}

void DepthMap::regularizeDepthMapFillHolesRow(int yMin, int yMax) {
  const float *keyFrameMaxGradBuf = activeKeyFrame->maxGrad[0].data();
  for (int y = yMin; y < yMax; y++)
    for (int x = 3; x < width - 2; x++) {
      const int idx = x + y * width;
      const Hypothesis *dest = &otherDepthMap[idx];
      if (dest->isValid) continue;
      if (keyFrameMaxGradBuf[idx] < MIN_USE_GRAD) continue;
      const int *io = validityIntegralBuffer.data() + idx;
      const int val = io[2 + 2 * width] - io[2 - 3 * width] - io[-3 + 2 * width] + io[-3 - 3 * width];
      if ((dest->blacklisted >= settings.minBlacklist && val > settings.valSumMinForCreate) || val > settings.valSumMinForUnblacklist) {
        float sumIdepthObs = 0, sumIVarObs = 0;
        for (int yy = y - 2; yy <= y + 2; yy++)
          for (int xx = x - 2; xx <= x + 2; xx++) {
            const Hypothesis *source = &otherDepthMap[xx + yy * width];
            if (!source->isValid) continue;
            sumIdepthObs += source->idepth / source->idepth_var;
            sumIVarObs += 1.0f / source->idepth_var;
          }
        float idepthObs = sumIdepthObs / sumIVarObs;
        idepthObs = UNZERO(idepthObs);
        currentDepthMap[idx] = Hypothesis(idepthObs, VAR_RANDOM_INIT_INITIAL, 0);
      }
    }
}

void DepthMap::regularizeDepthMapFillHoles() {
  activeKeyFrame->requireMaxGradients(0);
  buildRegIntegralBuffer();
  otherDepthMap = currentDepthMap;
  parallelRows(3, height - 2, 10, [this](int y0, int y1) { regularizeDepthMapFillHolesRow(y0, y1); });
}

// ---------------------------------------------------------------------------------------
// drivers (3.4, 3.5)
// ---------------------------------------------------------------------------------------
void DepthMap::updateKeyframe(const std::deque<Frame *> &referenceFrames) {
  oldest_referenceFrame = referenceFrames.front();
  newest_referenceFrame = referenceFrames.back();
  referenceFrameByID.clear();
  referenceFrameByID_offset = oldest_referenceFrame->id;
  for (Frame *frame : referenceFrames) {
    // refToKf: frames tracked on the active keyframe carry it in thisToParent_raw (the world-pose branch of
    // upstream needs the pose graph, which is out of scope: SURVEY.md 2b U11)
    const Sim3<double> refToKf = frame->thisToParent_raw;
    frame->prepareForStereoWith(activeKeyFrame, refToKf, 0);
    while ((int)referenceFrameByID.size() + referenceFrameByID_offset <= frame->id) referenceFrameByID.push_back(frame);
  }
  observeDepth();
  regularizeDepthMapFillHoles();
  regularizeDepthMap(false, settings.valSumMinForKeep);
  if (!activeKeyFrame->depthHasBeenUpdatedFlag) activeKeyFrame->setDepth(currentDepthMap.data());
  activeKeyFrame->numMappedOnThis++;
  activeKeyFrame->numMappedOnThisTotal++;
}

void DepthMap::createKeyFrame(Frame *new_keyframe) {
  const SE3<double> oldToNew_SE3 = se3FromSim3(new_keyframe->thisToParent_raw).inverse();
  propagateDepth(new_keyframe);
  activeKeyFrame = new_keyframe;
  activeKeyFrameIsReactivated = false;
  regularizeDepthMap(true, settings.valSumMinForKeep);
  regularizeDepthMapFillHoles();
  regularizeDepthMap(false, settings.valSumMinForKeep);

  // make mean inverse depth be one
  float sumIdepth = 0, numIdepth = 0;
  double sumD = 0;
  for (const Hypothesis &s : currentDepthMap) {
    if (!s.isValid) continue;
    sumIdepth += s.idepth_smoothed;
    sumD += (double)s.idepth_smoothed;
    numIdepth++;
  }
  if (g_exactSums) sumIdepth = (float)sumD;
  const float rescaleFactor = numIdepth / sumIdepth;
  const float rescaleFactor2 = rescaleFactor * rescaleFactor;
  for (Hypothesis &s : currentDepthMap) {
    if (!s.isValid) continue;
    s.idepth *= rescaleFactor;
    s.idepth_smoothed *= rescaleFactor;
    s.idepth_var *= rescaleFactor2;
    s.idepth_var_smoothed *= rescaleFactor2;
  }
  lastRescaleFactor = rescaleFactor;
  activeKeyFrame->thisToParent_raw = sim3FromSE3<double>(oldToNew_SE3.inverse(), (double)rescaleFactor);
  activeKeyFrame->setDepth(currentDepthMap.data());
}

void DepthMap::finalizeKeyFrame() {
  regularizeDepthMapFillHoles();
  regularizeDepthMap(false, settings.valSumMinForKeep);
  activeKeyFrame->setDepth(currentDepthMap.data());
}

// DepthMap::debugPlotDepthMap + DepthMapPixelHypothesis::getVisualizationColor (debugDisplay == 0):
// grey keyframe image, valid pixels overwritten with the idepth_smoothed rainbow; channel order as stored
// upstream (cv::Vec3b(255-r, 255-g, 255-b)), consumed as w*h*3 bytes by lib/GUI.cpp:104-108.
void DepthMap::debugPlotDepthMap(uint8_t *rgb) const {
  const float *I = activeKeyFrame->image[0].data();
  for (int i = 0; i < width * height; i++) {
    float gv = rintf(I[i]);  // cv::Mat::convertTo(CV_8UC1): saturate_cast<uchar>(round-half-even)
    const uint8_t g8 = (uint8_t)(gv < 0 ? 0 : (gv > 255 ? 255 : gv));
    rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = g8;
    const Hypothesis &p = currentDepthMap[i];
    if (!p.isValid) continue;
    const float id = p.idepth_smoothed;
    if (id < 0) {
      rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 255;
      continue;
    }
    float r = (0 - id) * 255 / 1.0f; if (r < 0) r = -r;
    float g = (1 - id) * 255 / 1.0f; if (g < 0) g = -g;
    float b = (2 - id) * 255 / 1.0f; if (b < 0) b = -b;
    const uint8_t rc = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    const uint8_t gc = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
    const uint8_t bc = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
    rgb[3 * i] = 255 - rc;
    rgb[3 * i + 1] = 255 - gc;
    rgb[3 * i + 2] = 255 - bc;
  }
}

}  // namespace lsdo
