// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Flat C entry points so tests / bench.py's cpu_baseline leg can drive the oracle via ctypes.
#include <chrono>
#include <cstdlib>
#include <deque>
#include <cstring>
#include <functional>
#include <thread>

#include "lsd_oracle.hpp"

using namespace lsdo;

extern "C" {

struct lsdo_se3_result {
  double frameToRef[7];  // qx qy qz qw tx ty tz
  float lastResidual, lastMeanRes, pointUsage, lastGoodCount, lastBadCount;
  float affine_a, affine_b, initialTrackedResidual;
  int diverged, trackingWasGood;
  int numResidualCalls[5], numWarpUpdateCalls[5];
  int traceLen;
};

struct lsdo_trace_entry {
  int level, accepted;
  float error, lambda;
  int bufSize;
};

enum { LSDO_IMAGE = 0, LSDO_GRADIENTS = 1, LSDO_MAXGRAD = 2, LSDO_IDEPTH = 3, LSDO_IDEPTHVAR = 4, LSDO_MASK = 5 };

void *lsdo_frame_create(int id, int w, int h, float fx, float fy, float cx, float cy, const uint8_t *img) {
  return new Frame(id, w, h, fx, fy, cx, cy, img);
}
void lsdo_frame_destroy(void *f) { delete (Frame *)f; }

void lsdo_frame_build_pyramids(void *fp) {
  Frame *f = (Frame *)fp;
  for (int l = 0; l < NL; l++) { f->requireImage(l); f->requireGradients(l); }
  f->requireMaxGradients(0);
}

int lsdo_frame_get(void *fp, int field, int level, void *dst) {
  Frame *f = (Frame *)fp;
  const size_t n = (size_t)f->w[level] * f->h[level];
  switch (field) {
    case LSDO_IMAGE: f->requireImage(level); std::memcpy(dst, f->image[level].data(), n * 4); return 0;
    case LSDO_GRADIENTS: f->requireGradients(level); std::memcpy(dst, f->grad[level].data(), n * 16); return 0;
    case LSDO_MAXGRAD: f->requireMaxGradients(level); std::memcpy(dst, f->maxGrad[level].data(), n * 4); return 0;
    case LSDO_IDEPTH: if (!f->hasIDepthBeenSet) return -1; f->requireIDepth(level); std::memcpy(dst, f->idepth[level].data(), n * 4); return 0;
    case LSDO_IDEPTHVAR: if (!f->hasIDepthBeenSet) return -1; f->requireIDepth(level); std::memcpy(dst, f->idepthVar[level].data(), n * 4); return 0;
    case LSDO_MASK: {
      const size_t m = (size_t)f->w[1] * f->h[1];
      std::memcpy(dst, f->refPixelWasGoodBuf(), m);
      return 0;
    }
  }
  return -2;
}
int lsdo_frame_num_mappable(void *fp) { Frame *f = (Frame *)fp; f->requireMaxGradients(0); return f->numMappablePixels; }
float lsdo_frame_mean_idepth(void *fp) { return ((Frame *)fp)->meanIdepth; }
int lsdo_frame_num_points(void *fp) { return ((Frame *)fp)->numPoints; }
void lsdo_frame_set_idepth(void *fp, const float *id, const float *var) { ((Frame *)fp)->setIDepthRaw(id, var); }
void lsdo_frame_set_depth_gt(void *fp, const float *depth, float cov) { ((Frame *)fp)->setDepthFromGroundTruth(depth, cov); }
void lsdo_frame_set_track_meta(void *fp, float initialTrackedResidual, int parentId, const double toParent[8]) {
  Frame *f = (Frame *)fp;
  f->initialTrackedResidual = initialTrackedResidual;
  f->trackingParentId = parentId;
  f->thisToParent_raw = Sim3<double>(Quat<double>(toParent[3], toParent[0], toParent[1], toParent[2]),
                                     Vec3<double>(toParent[4], toParent[5], toParent[6]), toParent[7]);
}
void lsdo_frame_set_mask(void *fp, const uint8_t *mask) {
  Frame *f = (Frame *)fp;
  uint8_t *m = f->refPixelWasGoodBuf();
  std::memcpy(m, mask, (size_t)f->w[1] * f->h[1]);
}
void lsdo_frame_set_counters(void *fp, int tracked, int mapped) {
  ((Frame *)fp)->numFramesTrackedOnThis = tracked;
  ((Frame *)fp)->numMappedOnThis = mapped;
}

void *lsdo_ref_create(void *kf) {
  auto *r = new TrackingReference();
  r->importFrame((Frame *)kf);
  return r;
}
void lsdo_ref_destroy(void *r) { delete (TrackingReference *)r; }
int lsdo_ref_num(void *rp, int level) {
  auto *r = (TrackingReference *)rp;
  r->makePointCloud(level);
  return r->numData[level];
}
// copies the point cloud of a level: pos (3n), grad (2n), colorVar (2n), idx (n); any may be null
int lsdo_ref_get(void *rp, int level, float *pos, float *grad, float *colvar, int *idx) {
  auto *r = (TrackingReference *)rp;
  r->makePointCloud(level);
  const int n = r->numData[level];
  if (pos) std::memcpy(pos, r->posData[level].data(), (size_t)n * 12);
  if (grad) std::memcpy(grad, r->gradData[level].data(), (size_t)n * 8);
  if (colvar) std::memcpy(colvar, r->colorAndVarData[level].data(), (size_t)n * 8);
  if (idx) std::memcpy(idx, r->pointPosInXYGrid[level].data(), (size_t)n * 4);
  return n;
}

static SE3<double> pose_in(const double p[7]) {
  return SE3<double>(Quat<double>(p[3], p[0], p[1], p[2]), Vec3<double>(p[4], p[5], p[6]));
}
static void pose_out(const SE3<double> &s, double p[7]) {
  p[0] = s.q.x; p[1] = s.q.y; p[2] = s.q.z; p[3] = s.q.w; p[4] = s.t.x; p[5] = s.t.y; p[6] = s.t.z;
}

static void fill_result(const SE3Tracker &t, const Frame *frame, const SE3<double> &res, lsdo_se3_result *out) {
  pose_out(res, out->frameToRef);
  out->lastResidual = t.lastResidual;
  out->lastMeanRes = t.lastMeanRes;
  out->pointUsage = t.pointUsage;
  out->lastGoodCount = t.lastGoodCount;
  out->lastBadCount = t.lastBadCount;
  out->affine_a = t.affineEstimation_a;
  out->affine_b = t.affineEstimation_b;
  out->initialTrackedResidual = frame->initialTrackedResidual;
  out->diverged = t.diverged;
  out->trackingWasGood = t.trackingWasGood;
  for (int l = 0; l < 5; l++) {
    out->numResidualCalls[l] = t.numCalcResidualCalls[l];
    out->numWarpUpdateCalls[l] = t.numCalcWarpUpdateCalls[l];
  }
  out->traceLen = (int)t.trace.size();
}

// SE3Tracker::trackFrame.  mode: 0 scalar, 1 sse4-order.  trace may be null.
int lsdo_se3_track(void *refp, void *framep, const double init[7], int mode, lsdo_se3_result *out, lsdo_trace_entry *trace,
                   int traceCap) {
  auto *ref = (TrackingReference *)refp;
  Frame *frame = (Frame *)framep;
  SE3Tracker t(frame->w[0], frame->h[0]);
  t.mode = (ReduceMode)mode;
  const SE3<double> res = t.trackFrame(ref, frame, pose_in(init));
  fill_result(t, frame, res, out);
  if (trace)
    for (int i = 0; i < (int)t.trace.size() && i < traceCap; i++)
      trace[i] = {t.trace[i].level, t.trace[i].accepted, t.trace[i].error, t.trace[i].lambda, t.trace[i].bufSize};
  return 0;
}

// SE3Tracker::trackFrameOnPermaref with the "test track" settings of DenseDepthTrackerSettings (maxItsTestTrack 5,
// stepSizeMinTestTrack 1e-3, convergenceEpsTestTrack 0.98, lambdaInitialTestTrack 0) at QUICK_KF_CHECK_LVL.
// out->frameToRef receives the returned referenceToFrame.
int lsdo_se3_track_permaref(void *refp, void *framep, const double init_refToFrame[7], int mode, lsdo_se3_result *out,
                            lsdo_trace_entry *trace, int traceCap) {
  auto *ref = (TrackingReference *)refp;
  Frame *frame = (Frame *)framep;
  SE3Tracker t(frame->w[0], frame->h[0]);
  t.mode = (ReduceMode)mode;
  t.settings.maxItsPerLvl[QUICK_KF_CHECK_LVL] = 5;
  t.settings.stepSizeMin[QUICK_KF_CHECK_LVL] = 1e-3f;
  t.settings.convergenceEps[QUICK_KF_CHECK_LVL] = 0.98f;
  t.settings.lambdaInitial[QUICK_KF_CHECK_LVL] = 0;
  const float itr = frame->initialTrackedResidual;
  const SE3<double> res = t.trackFrameOnPermaref(ref->keyframe, ref, frame, pose_in(init_refToFrame));
  fill_result(t, frame, res, out);
  out->initialTrackedResidual = itr;
  if (trace)
    for (int i = 0; i < (int)t.trace.size() && i < traceCap; i++)
      trace[i] = {t.trace[i].level, t.trace[i].accepted, t.trace[i].error, t.trace[i].lambda, t.trace[i].bufSize};
  return 0;
}

// SE3Tracker::checkPermaRefOverlap
float lsdo_check_permaref_overlap(void *refp, const double refToFrame[7]) {
  auto *ref = (TrackingReference *)refp;
  SE3Tracker t(ref->keyframe->w[0], ref->keyframe->h[0]);
  return t.checkPermaRefOverlap(ref->keyframe, ref, pose_in(refToFrame));
}

// One LM evaluation at a fixed pose (B3+B4+B5) -- used by the parity tests to compare the fused GPU
// evaluation against the three reference passes.  out38: see tests/test_se3_eval.py for the order.
int lsdo_se3_eval(void *refp, void *framep, const double refToFrame[7], int level, float affine_a, float affine_b, int mode,
                  float *A36, float *b6, float *scalars /*[12]*/) {
  auto *ref = (TrackingReference *)refp;
  Frame *frame = (Frame *)framep;
  SE3Tracker t(frame->w[0], frame->h[0]);
  t.mode = (ReduceMode)mode;
  t.affineEstimation_a = affine_a;
  t.affineEstimation_b = affine_b;
  ref->makePointCloud(level);
  const SE3<float> pose = pose_in(refToFrame).cast<float>();
  const int *idxb = (level == SE3TRACKING_MIN_LEVEL) ? ref->pointPosInXYGrid[level].data() : nullptr;
  const float meanRes2 = t.calcResidualAndBuffers(ref->posData[level].data(), ref->colorAndVarData[level].data(), idxb,
                                                  ref->numData[level], frame, pose, level);
  const float err = t.calcWeightsAndResidual(pose);
  float A[6][6], b[6], lsErr;
  t.calculateWarpUpdate(A, b, &lsErr);
  for (int i = 0; i < 6; i++) { b6[i] = b[i]; for (int j = 0; j < 6; j++) A36[6 * i + j] = A[i][j]; }
  scalars[0] = err;
  scalars[1] = meanRes2;
  scalars[2] = (float)t.buf_warped_size;
  scalars[3] = t.lastGoodCount;
  scalars[4] = t.lastBadCount;
  scalars[5] = t.pointUsage;
  scalars[6] = t.lastMeanRes;
  scalars[7] = t.affineEstimation_a_lastIt;
  scalars[8] = t.affineEstimation_b_lastIt;
  scalars[9] = lsErr;
  scalars[10] = 0;
  scalars[11] = 0;
  return 0;
}

// Timed batch for the CPU baseline: tracks n independent (ref, frame) pairs on `threads` host threads.
// Returns wall seconds.  Each worker owns its tracker (the reference has one tracking thread per SlamSystem;
// independent pairs over cores is the most generous CPU arm).
double lsdo_se3_track_batch(int n, void **refs, void **frames, const double *inits /*n*7*/, int mode, int threads,
                            lsdo_se3_result *outs) {
  if (threads < 1) threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int tid = 0; tid < threads; tid++) {
    pool.emplace_back([=]() {
      if (n == 0) return;
      Frame *f0 = (Frame *)frames[0];
      SE3Tracker t(f0->w[0], f0->h[0]);
      t.mode = (ReduceMode)mode;
      for (int i = tid; i < n; i += threads) {
        Frame *frame = (Frame *)frames[i];
        const SE3<double> res = t.trackFrame((TrackingReference *)refs[i], frame, pose_in(inits + 7 * i));
        fill_result(t, frame, res, &outs[i]);
      }
    });
  }
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Builds n (keyframe, frame, reference) triples on `threads` host threads (setup for the timed baseline):
// pyramids for both frames, keyframe idepth installed, point clouds of levels 1..4 prebuilt.
void lsdo_make_pairs(int n, const uint8_t *const *kf_imgs, const uint8_t *const *fr_imgs, const float *const *idepth,
                     const float *const *var, int w, int h, const float K[4], int threads, void **out_kf, void **out_fr,
                     void **out_ref) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int tid = 0; tid < threads; tid++) {
    pool.emplace_back([=]() {
      for (int i = tid; i < n; i += threads) {
        Frame *kf = new Frame(2 * i, w, h, K[0], K[1], K[2], K[3], kf_imgs[i]);
        Frame *fr = new Frame(2 * i + 1, w, h, K[0], K[1], K[2], K[3], fr_imgs[i]);
        for (int l = 0; l < NL; l++) { kf->requireImage(l); kf->requireGradients(l); fr->requireImage(l); fr->requireGradients(l); }
        kf->setIDepthRaw(idepth[i], var[i]);
        auto *r = new TrackingReference();
        r->importFrame(kf);
        for (int l = 1; l < NL; l++) r->makePointCloud(l);
        out_kf[i] = kf; out_fr[i] = fr; out_ref[i] = r;
      }
    });
  }
  for (auto &th : pool) th.join();
}


// Frees everything the timed tracking loop never reads (the keyframe's planes once its point clouds exist; the new
// frame's images and level-0 gradients): 1000 prepared pairs then take ~3 GB of host memory instead of ~20 GB, so that the
// CPU arm can run the FULL BASELINE configs[1] batch.
void lsdo_pairs_trim(int n, void **kfs, void **frs) {
  auto drop = [](std::vector<float> &v) { std::vector<float>().swap(v); };
  for (int i = 0; i < n; i++) {
    Frame *kf = (Frame *)kfs[i], *fr = (Frame *)frs[i];
    for (int l = 0; l < NL; l++) {
      drop(kf->image[l]); drop(kf->grad[l]); drop(kf->maxGrad[l]); drop(kf->idepth[l]); drop(kf->idepthVar[l]);
      drop(fr->image[l]); drop(fr->maxGrad[l]);
    }
    drop(fr->grad[0]);
  }
}

// ---- Sim3Tracker ---------------------------------------------------------------------------------
struct lsdo_sim3_result {
  double frameToRef[8];  // qx qy qz qw tx ty tz scale
  float hessian[49];
  float lastResidual, lastDepthResidual, lastPhotometricResidual, pointUsage, affine_a, affine_b;
  int diverged, traceLen;
};

static Sim3<double> sim3_in(const double p[8]) {
  return Sim3<double>(Quat<double>(p[3], p[0], p[1], p[2]), Vec3<double>(p[4], p[5], p[6]), p[7]);
}
static void sim3_out(const Sim3<double> &s, double p[8]) {
  p[0] = s.q.x; p[1] = s.q.y; p[2] = s.q.z; p[3] = s.q.w; p[4] = s.t.x; p[5] = s.t.y; p[6] = s.t.z; p[7] = s.s;
}
static void fill_sim3(const Sim3Tracker &t, const Sim3<double> &res, lsdo_sim3_result *out) {
  sim3_out(res, out->frameToRef);
  std::memcpy(out->hessian, t.lastSim3Hessian, sizeof(out->hessian));
  out->lastResidual = t.lastResidual;
  out->lastDepthResidual = t.lastDepthResidual;
  out->lastPhotometricResidual = t.lastPhotometricResidual;
  out->pointUsage = t.pointUsage;
  out->affine_a = t.affineEstimation_a;
  out->affine_b = t.affineEstimation_b;
  out->diverged = t.diverged;
  out->traceLen = (int)t.trace.size();
}

// Sim3Tracker::trackFrameSim3(reference, frame, frameToReference_initialEstimate, startLevel, finalLevel)
int lsdo_sim3_track(void *refp, void *framep, const double init[8], int startLevel, int finalLevel, int mode, lsdo_sim3_result *out,
                    lsdo_trace_entry *trace, int traceCap) {
  auto *ref = (TrackingReference *)refp;
  Frame *frame = (Frame *)framep;
  Sim3Tracker t(frame->w[0], frame->h[0]);
  t.mode = (ReduceMode)mode;
  const Sim3<double> res = t.trackFrameSim3(ref, frame, sim3_in(init), startLevel, finalLevel);
  fill_sim3(t, res, out);
  if (trace)
    for (int i = 0; i < (int)t.trace.size() && i < traceCap; i++)
      trace[i] = {t.trace[i].level, t.trace[i].accepted, t.trace[i].error, t.trace[i].lambda, t.trace[i].bufSize};
  return 0;
}

// n independent Sim3 tracks on `threads` host threads; returns wall seconds
double lsdo_sim3_track_batch(int n, void **refs, void **frames, const double *inits /*n*8*/, int startLevel, int finalLevel, int mode,
                             int threads, lsdo_sim3_result *outs) {
  if (threads < 1) threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int tid = 0; tid < threads; tid++) {
    pool.emplace_back([=]() {
      if (n == 0) return;
      Frame *f0 = (Frame *)frames[0];
      Sim3Tracker t(f0->w[0], f0->h[0]);
      t.mode = (ReduceMode)mode;
      for (int i = tid; i < n; i += threads) {
        const Sim3<double> res = t.trackFrameSim3((TrackingReference *)refs[i], (Frame *)frames[i], sim3_in(inits + 8 * i), startLevel, finalLevel);
        fill_sim3(t, res, &outs[i]);
      }
    });
  }
  for (auto &th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- DepthMap -----------------------------------------------------------------------------------
static double secs_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void *lsdo_depthmap_create(int w, int h, float fx, float fy, float cx, float cy, int threads) {
  auto *d = new DepthMap(w, h, fx, fy, cx, cy);
  d->numThreads = threads;
  return d;
}
void lsdo_depthmap_destroy(void *d) { delete (DepthMap *)d; }
void lsdo_depthmap_set_thresholds(void *dp, int create, int keep, int unblacklist, int minBlacklist) {
  auto *d = (DepthMap *)dp;
  d->settings.valSumMinForCreate = create;
  d->settings.valSumMinForKeep = keep;
  d->settings.valSumMinForUnblacklist = unblacklist;
  d->settings.minBlacklist = minBlacklist;
}
void lsdo_depthmap_init_gt(void *d, void *frame) { ((DepthMap *)d)->initializeFromGTDepth((Frame *)frame); }
void lsdo_depthmap_init_random(void *d, void *frame, unsigned seed) {
  srand(seed);
  ((DepthMap *)d)->initializeRandomly((Frame *)frame);
}
void lsdo_depthmap_init_map(void *d, void *frame, const Hypothesis *map) { ((DepthMap *)d)->initializeFromMap((Frame *)frame, map); }
void lsdo_depthmap_set_reactivated(void *d, int v) { ((DepthMap *)d)->activeKeyFrameIsReactivated = v != 0; }
void lsdo_depthmap_read(void *dp, Hypothesis *out) {
  auto *d = (DepthMap *)dp;
  std::memcpy(out, d->currentDepthMap.data(), d->currentDepthMap.size() * sizeof(Hypothesis));
}
void lsdo_depthmap_write(void *dp, const Hypothesis *in) {
  auto *d = (DepthMap *)dp;
  std::memcpy(d->currentDepthMap.data(), in, d->currentDepthMap.size() * sizeof(Hypothesis));
}
float lsdo_depthmap_last_rescale(void *dp) { return ((DepthMap *)dp)->lastRescaleFactor; }

// DepthMap::updateKeyframe(std::deque<std::shared_ptr<Frame>>); returns wall seconds
double lsdo_depthmap_update_keyframe(void *dp, int n, void **frames, double *stageSecs /*4: observe, fillHoles, regularize, setDepth*/) {
  auto *d = (DepthMap *)dp;
  std::deque<Frame *> refs;
  for (int i = 0; i < n; i++) refs.push_back((Frame *)frames[i]);
  auto t0 = std::chrono::steady_clock::now();
  d->updateKeyframe(refs);
  (void)stageSecs;
  return secs_since(t0);
}
double lsdo_depthmap_create_keyframe(void *dp, void *newKf) {
  auto t0 = std::chrono::steady_clock::now();
  ((DepthMap *)dp)->createKeyFrame((Frame *)newKf);
  return secs_since(t0);
}
void lsdo_depthmap_finalize(void *dp) { ((DepthMap *)dp)->finalizeKeyFrame(); }

// individual stages (parity + timing): 0 observeDepth (refs must have been prepared by a prior
// lsdo_depthmap_prepare), 1 regularizeDepthMapFillHoles, 2 regularizeDepthMap(arg1 = removeOcclusions, arg2 = validityTH),
// 3 propagateDepth(frame), 4 activeKeyFrame->setDepth + idepth pyramids
double lsdo_depthmap_stage(void *dp, int stage, int arg1, int arg2, void *frame) {
  auto *d = (DepthMap *)dp;
  auto t0 = std::chrono::steady_clock::now();
  switch (stage) {
    case 0: d->observeDepth(); break;
    case 1: d->regularizeDepthMapFillHoles(); break;
    case 2: d->regularizeDepthMap(arg1 != 0, arg2); break;
    case 3: d->propagateDepth((Frame *)frame); d->activeKeyFrame = (Frame *)frame; d->activeKeyFrameIsReactivated = false; break;
    case 4:
      d->activeKeyFrame->setDepth(d->currentDepthMap.data());
      for (int l = 1; l < NL; l++) d->activeKeyFrame->requireIDepth(l);
      break;
    default: return -1;
  }
  return secs_since(t0);
}
// the bookkeeping part of updateKeyframe (prepareForStereoWith + referenceFrameByID) without running the stages
void lsdo_depthmap_prepare(void *dp, int n, void **frames) {
  auto *d = (DepthMap *)dp;
  d->oldest_referenceFrame = (Frame *)frames[0];
  d->newest_referenceFrame = (Frame *)frames[n - 1];
  d->referenceFrameByID.clear();
  d->referenceFrameByID_offset = d->oldest_referenceFrame->id;
  for (int i = 0; i < n; i++) {
    Frame *f = (Frame *)frames[i];
    f->prepareForStereoWith(d->activeKeyFrame, f->thisToParent_raw, 0);
    while ((int)d->referenceFrameByID.size() + d->referenceFrameByID_offset <= f->id) d->referenceFrameByID.push_back(f);
  }
}
void lsdo_depthmap_debug_rgb(void *dp, uint8_t *rgb) { ((DepthMap *)dp)->debugPlotDepthMap(rgb); }

// one doLineStereo call (unit tests against analytic disparity); out3 = {idepth, var, eplLength}; ref must be prepared
float lsdo_line_stereo(void *dp, void *ref, float u, float v, float epxn, float epyn, float min_id, float prior_id, float max_id,
                       float *out3) {
  auto *d = (DepthMap *)dp;
  Frame *r = (Frame *)ref;
  d->activeKeyFrame->requireGradients(0);
  float id = 0, var = 0, len = 0;
  const float e = d->doLineStereo(u, v, epxn, epyn, min_id, prior_id, max_id, r, r->image[0].data(), id, var, len);
  out3[0] = id; out3[1] = var; out3[2] = len;
  return e;
}
int lsdo_make_epl(void *dp, void *ref, int x, int y, float *ep2) {
  return ((DepthMap *)dp)->makeAndCheckEPL(x, y, (Frame *)ref, ep2, ep2 + 1) ? 1 : 0;
}
// frame state the depth map reads / writes
void lsdo_frame_get_pose(void *fp, double out8[8]) { sim3_out(((Frame *)fp)->thisToParent_raw, out8); }
void lsdo_frame_get_counters(void *fp, int *out3) {
  Frame *f = (Frame *)fp;
  out3[0] = f->numFramesTrackedOnThis; out3[1] = f->numMappedOnThis; out3[2] = f->numMappedOnThisTotal;
}
void lsdo_frame_set_flags(void *fp, int depthHasBeenUpdated) { ((Frame *)fp)->depthHasBeenUpdatedFlag = depthHasBeenUpdated != 0; }
int lsdo_frame_get_flags(void *fp) { return ((Frame *)fp)->depthHasBeenUpdatedFlag ? 1 : 0; }
void lsdo_frame_clear_mask(void *fp) { ((Frame *)fp)->refPixelWasGood.clear(); }

void lsdo_set_exact_sums(int v) { g_exactSums = v != 0; }

int lsdo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
