// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Flat C entry points so tests / bench.py's cpu_baseline leg can drive the oracle via ctypes.
#include <chrono>
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

int lsdo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
