// slam.cu -- lock-step tracking + mapping driver over the hot path: the part of [UP]
// lsd_slam::SlamSystem that the reference application drives, with the semantics it uses when
// Conf().runRealTime == false -- `system->nextImage(idx, image, camera)` blocks until the frame has been tracked
// AND mapped (/root/reference/lib/App/InputThread.cpp:70-71).  Host code only: every piece of arithmetic is a call
// into the C ABI of this library (lsd_frame_create, lsd_se3_track, lsd_depth_update_keyframe, ...).
//
//   nextImage      -> first frame: randomInit (or gtDepthInit when depth is supplied);
//                     otherwise SlamSystem::trackFrame (TrackingReference::importFrame when the keyframe or its depth
//                     changed, SE3Tracker::trackFrame from the last frame-to-keyframe pose) followed by one
//                     doMappingIteration: updateKeyframe, or finishCurrentKeyframe + createNewCurrentKeyframe when the
//                     keyframe-selection score asks for a new one.
//   status         -> what the reference's output wrappers read per frame: camToWorld (publishPose,
//                     PangolinOutputIOWrapper / TextOutputIOWrapper.cpp:104-117), thisToParent_raw, keyframe flag.
// Pose-graph optimisation, loop closure and relocalisation are not here: they stay on the reference's CPU code
// (BASELINE.json north_star); a lost frame is reported and dropped.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.cuh"
#include "lie_dev.cuh"

namespace lsd {

// util/settings.h
static const float KF_DIST_WEIGHT = 4.0f, KF_USAGE_WEIGHT = 3.0f;
static const int INITIALIZATION_PHASE_COUNT = 5, MIN_NUM_MAPPED = 5;

static void sim3_identity(double p[8]) {
  p[0] = p[1] = p[2] = 0; p[3] = 1;
  p[4] = p[5] = p[6] = 0; p[7] = 1;
}

// a * b for Sim3 {qx,qy,qz,qw,tx,ty,tz,s}
static void sim3_mul(const double a[8], const double b[8], double o[8]) {
  QuatT<double> qa = {a[0], a[1], a[2], a[3]}, qb = {b[0], b[1], b[2], b[3]};
  QuatT<double> q = qmul(qa, qb);
  qnormalize(q);
  double R[9], rt[3];
  qtoR(qa, R);
  mat3vec(R, b + 4, rt);
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  for (int i = 0; i < 3; i++) o[4 + i] = a[4 + i] + a[7] * rt[i];
  o[7] = a[7] * b[7];
}

}  // namespace lsd

using namespace lsd;

struct lsd_slam_system {
  lsd_ctx *ctx;
  lsd_depthmap *dm;
  lsd_frame *kf;       // current keyframe
  lsd_ref *ref;        // tracking reference of `kf`
  int refKfId;
  std::vector<lsd_frame *> keyframes;  // finished keyframes stay alive (upstream keeps them in the KeyFrameGraph)
  double kfWorld[8];   // camToWorld of the current keyframe
  double lastToKf[8];  // last tracked frame -> current keyframe
  int nKeyframes, tracked, lost;
  float kfMeanIdepth;
  bool kfMeanValid;
  int keepFinishedKeyframes;
  bool pipelined;        // see lsd_slam_set_pipelined
  lsd_undistorter *und;  // optional: images arrive distorted, as at InputThread.cpp:59-62
  double stageSec[5];    // wall time spent in: frame ingest, reference import, tracking, updateKeyframe, keyframe switch
};

extern "C" {

// [UP] TrackableKeyFrameSearch::getRefFrameScore(distanceSquared, usage) with KFDistWeight 4, KFUsageWeight 3 (float arithmetic,
// upstream's operation order)
float lsd_slam_ref_frame_score(float distanceSquared, float usage) {
  return distanceSquared * KF_DIST_WEIGHT * KF_DIST_WEIGHT + (1 - usage) * (1 - usage) * KF_USAGE_WEIGHT * KF_USAGE_WEIGHT;
}

int lsd_slam_create(lsd_ctx *ctx, lsd_slam_system **out) {
  LSD_ARG(ctx && out);
  lsd_slam_system *s = new lsd_slam_system();
  s->ctx = ctx;
  s->dm = nullptr;
  s->kf = nullptr;
  s->ref = nullptr;
  s->refKfId = -1;
  sim3_identity(s->kfWorld);
  sim3_identity(s->lastToKf);
  s->nKeyframes = s->tracked = s->lost = 0;
  s->kfMeanIdepth = 0;
  s->kfMeanValid = false;
  s->keepFinishedKeyframes = 1;
  {
    const char *e = std::getenv("LSD_B200_SLAM_PIPELINE");  // experiments: 0 = every stage synchronises on its own
    s->pipelined = !(e && e[0] == '0');
  }
  s->und = nullptr;
  for (double &v : s->stageSec) v = 0;
  int rc = lsd_depthmap_create(ctx, &s->dm);
  if (rc) { delete s; return rc; }
  // the pipelined driver keeps two pointer lists of a frame in flight in the context's pinned table (frames at offset 0,
  // tracking references at 8192): sized here once, so that no list is ever moved while a copy still reads it
  if ((rc = ensure_table(ctx, 65536))) { lsd_depthmap_destroy(ctx, s->dm); delete s; return rc; }
  *out = s;
  return LSD_OK;
}

int lsd_slam_destroy(lsd_slam_system *s) {
  if (!s) return LSD_OK;
  ctx_finish_pending(s->ctx);
  if (s->ref) lsd_ref_release(s->ctx, s->ref);
  for (lsd_frame *f : s->keyframes) lsd_frame_release(s->ctx, f);
  if (s->kf) lsd_frame_release(s->ctx, s->kf);
  if (s->dm) lsd_depthmap_destroy(s->ctx, s->dm);
  delete s;
  return LSD_OK;
}

int lsd_slam_set_keep_keyframes(lsd_slam_system *s, int keep) {
  LSD_ARG(s);
  s->keepFinishedKeyframes = keep;
  return LSD_OK;
}

int lsd_slam_set_pipelined(lsd_slam_system *s, int enable) {
  LSD_ARG(s);
  s->pipelined = enable != 0;
  return LSD_OK;
}

int lsd_slam_set_undistorter(lsd_slam_system *s, lsd_undistorter *und) {
  LSD_ARG(s);
  s->und = und;
  return LSD_OK;
}

static float slam_min_val(const lsd_slam_system *s);

static int slam_new_frame(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, unsigned flags, lsd_frame **f) {
  if (s->und) return lsd_frame_create_undistorted(s->ctx, s->und, id, image, pitch, flags, nullptr, f);
  return lsd_frame_create(s->ctx, id, image, pitch, flags, f);
}

static void fill_status(lsd_slam_system *s, int id, int tracked, int isKeyframe, const double toKf[8], const lsd_se3_result *r, float score,
                        lsd_slam_status *st) {
  if (!st) return;
  std::memset(st, 0, sizeof(*st));
  st->frameId = id;
  st->tracked = tracked;
  st->isKeyframe = isKeyframe;
  st->numKeyframes = s->nKeyframes;
  st->currentKeyframeId = s->kf ? s->kf->id : -1;
  st->keyframeScore = score;
  std::memcpy(st->thisToParent_raw, toKf, sizeof(double) * 8);
  sim3_mul(s->kfWorld, toKf, st->camToWorld);
  st->keyframeRescale = 1.0;
  if (r) {
    st->pointUsage = r->pointUsage;
    st->lastResidual = r->lastResidual;
    st->trackingWasGood = r->trackingWasGood;
    st->diverged = r->diverged;
  }
}

static int first_keyframe(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, const float *depth, lsd_slam_status *st) {
  LSD_ARG(s && image);
  if (s->kf) { set_error("SlamSystem already initialised (fullReset = destroy + create)"); return LSD_ERR_STATE; }
  lsd_frame *kf = nullptr;
  int rc = slam_new_frame(s, id, image, pitch, LSD_BUILD_MAXGRAD0 | LSD_BUILD_GRAD0, &kf);
  if (rc) return rc;
  if (depth) {  // SlamSystem::gtDepthInit: Frame::setDepthFromGroundTruth + DepthMap::initializeFromGTDepth
    rc = lsd_frame_set_depth_from_gt(s->ctx, kf, depth, 1.0f);
    if (!rc) rc = lsd_depth_initialize_from_gt(s->ctx, s->dm, kf);
  } else {  // SlamSystem::randomInit
    rc = lsd_depth_initialize_randomly(s->ctx, s->dm, kf);
  }
  if (rc) {
    lsd_frame_release(s->ctx, kf);
    return rc;
  }
  s->kf = kf;
  s->nKeyframes = 1;
  s->kfMeanValid = false;
  sim3_identity(s->kfWorld);
  sim3_identity(s->lastToKf);
  double id8[8];
  sim3_identity(id8);
  fill_status(s, id, 1, 1, id8, nullptr, 0.0f, st);
  return LSD_OK;
}

int lsd_slam_gt_depth_init(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, const float *depth, lsd_slam_status *st) {
  LSD_ARG(depth);
  return first_keyframe(s, id, image, pitch, depth, st);
}

int lsd_slam_random_init(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, lsd_slam_status *st) {
  return first_keyframe(s, id, image, pitch, nullptr, st);
}

int lsd_slam_next_image(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, lsd_slam_status *st) {
  LSD_ARG(s && image);
  if (!s->kf) return first_keyframe(s, id, image, pitch, nullptr, st);  // SlamSystem::nextImage: randomInit on the first image
  lsd_ctx *ctx = s->ctx;
  int rc;
  lsd_frame *f = nullptr;
  typedef std::chrono::steady_clock clk;
  clk::time_point t0 = clk::now();
  auto lap = [&](int k) {
    const clk::time_point t1 = clk::now();
    s->stageSec[k] += std::chrono::duration<double>(t1 - t0).count();
    t0 = t1;
  };
  // Pipelined (default): frame ingest, reference import and the tracker are queued back to back and synchronised ONCE, by the
  // tracker's result; updateKeyframe is queued and finished by the next call, AFTER that call has staged its image into pinned
  // memory -- the mapping kernels of frame k run while the host copies frame k + 1.  Same kernels in the same stream order:
  // not a bit changes.  The caller still sees blocking semantics: every entry point that reads a result of the mapping
  // (or rewrites a table it reads) finishes it first (ctx_finish_pending).
  struct DeferGuard {
    lsd_ctx *ctx;
    ~DeferGuard() { ctx->deferSync = false; }
  } deferGuard = {ctx};
  ctx->deferSync = s->pipelined;
  // a tracked frame needs image + gradient pyramids only; maxGradients(0) / gradients(0) are built lazily if it is promoted
  // to keyframe (propagateDepth asks for them)
  if ((rc = slam_new_frame(s, id, image, pitch, LSD_BUILD_TRACKING, &f))) return rc;
  lap(0);
  struct FrameGuard {  // the new frame is dropped on every error exit below
    lsd_ctx *ctx;
    lsd_frame *&f;
    bool armed;
    ~FrameGuard() { if (armed && f) lsd_frame_release(ctx, f); }
  } guard = {ctx, f, true};

  // ---- SlamSystem::trackFrame
  if (!s->ref || s->refKfId != s->kf->id || s->kf->depthHasBeenUpdatedFlag) {
    if (s->ref) lsd_ref_release(ctx, s->ref);
    s->ref = nullptr;
    if ((rc = lsd_ref_create(ctx, s->kf, &s->ref))) return rc;
    s->refKfId = s->kf->id;
    s->kf->depthHasBeenUpdatedFlag = false;
  }
  lap(1);
  lsd_se3_result res;
  if ((rc = lsd_se3_track(ctx, s->ref, f, s->lastToKf, &res, nullptr))) return rc;
  lap(2);
  if (res.diverged || !res.trackingWasGood) {  // upstream hands over to the Relocalizer (out of scope): drop the frame
    s->lost++;
    fill_status(s, id, 0, 0, s->lastToKf, &res, 0.0f, st);
    return LSD_OK;  // guard releases the frame
  }
  s->tracked++;
  double toKf[8];
  std::memcpy(toKf, res.frameToRef, sizeof(double) * 7);
  toKf[7] = 1.0;
  std::memcpy(s->lastToKf, toKf, sizeof(toKf));

  // ---- keyframe selection (A.10): distance weighted by the keyframe's mean inverse depth + point usage
  bool create = false;
  float score = 0.0f;
  if (s->kf->numMappedOnThis > MIN_NUM_MAPPED) {
    if (!s->kfMeanValid) {
      if ((rc = lsd_frame_mean_idepth(ctx, s->kf, &s->kfMeanIdepth, nullptr))) return rc;
      s->kfMeanValid = true;
    }
    const double m = (double)s->kfMeanIdepth;
    const double d[3] = {toKf[4] * m, toKf[5] * m, toKf[6] * m};
    const float minVal = slam_min_val(s);
    // TrackableKeyFrameSearch::getRefFrameScore: distSq * KFDistWeight^2 + (1 - usage)^2 * KFUsageWeight^2  (16 and 9)
    score = lsd_slam_ref_frame_score((float)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), res.pointUsage);
    create = score > minVal;
  }

  // ---- one blocking mapping iteration
  if (create) {
    // the frame is published with the pose it was tracked at (keyframe pose * frame-to-keyframe, scale 1): upstream's
    // publishTrackedFrame runs before the mapping thread promotes the frame and folds the rescale factor in
    fill_status(s, id, 1, 1, toKf, &res, score, st);
    if ((rc = lsd_depth_finalize_keyframe(ctx, s->dm))) return rc;          // finishCurrentKeyframe
    if ((rc = lsd_depth_create_keyframe(ctx, s->dm, f, nullptr))) return rc;  // createNewCurrentKeyframe
    double world[8];
    sim3_mul(s->kfWorld, f->thisToParent_raw, world);  // thisToParent_raw now carries the rescale factor
    std::memcpy(s->kfWorld, world, sizeof(world));
    lsd_ref_release(ctx, s->ref);  // the tracking reference belongs to the finished keyframe
    s->ref = nullptr;
    if (s->keepFinishedKeyframes) s->keyframes.push_back(s->kf);
    else lsd_frame_release(ctx, s->kf);
    s->kf = f;
    guard.armed = false;  // the frame lives on as the current keyframe
    s->nKeyframes++;
    s->kfMeanValid = false;
    sim3_identity(s->lastToKf);
    if (st) {
      st->numKeyframes = s->nKeyframes;
      st->currentKeyframeId = s->kf->id;
      st->keyframeRescale = f->thisToParent_raw[7];
    }
    lap(4);
  } else {
    const bool setsDepth = !s->kf->depthHasBeenUpdatedFlag;  // updateKeyframe runs setDepth only when the flag is clear
    if ((rc = lsd_depth_update_keyframe(ctx, s->dm, 1, &f, nullptr))) return rc;
    if (setsDepth) s->kfMeanValid = false;
    fill_status(s, id, 1, 0, toKf, &res, score, st);
    lsd_frame_release(ctx, f);
    guard.armed = false;
    lap(3);
  }
  return LSD_OK;
}

// minVal of SlamSystem::trackFrame's keyframe decision
static float slam_min_val(const lsd_slam_system *s) {
  float minVal = std::fmin(0.2f + s->nKeyframes * 0.8f / INITIALIZATION_PHASE_COUNT, 1.0f);
  if (s->nKeyframes < INITIALIZATION_PHASE_COUNT) minVal *= 0.7f;
  return minVal;
}

// n live sequences on one context, one image each: the stages of lsd_slam_next_image, each batched over the sequences.
int lsd_slam_next_image_batch(int n, lsd_slam_system *const *sys, const int *ids, const uint8_t *const *images, size_t pitch,
                              lsd_slam_status *st) {
  LSD_ARG(n >= 1 && sys && ids && images && st);
  lsd_ctx *ctx = sys[0] ? sys[0]->ctx : nullptr;
  LSD_ARG(ctx);
  for (int i = 0; i < n; i++) {
    LSD_ARG(sys[i] && images[i]);
    LSD_ARG(sys[i]->ctx == ctx);
    if (!sys[i]->kf || sys[i]->und) {
      set_error("lsd_slam_next_image_batch: every system must be initialised and have no undistorter attached");
      return LSD_ERR_STATE;
    }
    for (int j = 0; j < i; j++) LSD_ARG(sys[j] != sys[i]);
  }
  typedef std::chrono::steady_clock clk;
  clk::time_point t0 = clk::now();
  auto lap = [&](int k) {
    const clk::time_point t1 = clk::now();
    const double dt = std::chrono::duration<double>(t1 - t0).count() / n;
    for (int i = 0; i < n; i++) sys[i]->stageSec[k] += dt;
    t0 = t1;
  };
  int rc;
  // pipelined like lsd_slam_next_image: one synchronisation (the tracker's) for ingest + import + tracking, and the batched
  // updateKeyframe -- queued LAST, after the keyframe switches -- is finished by the next call, after it has staged its images
  struct DeferGuard {
    lsd_ctx *ctx;
    ~DeferGuard() { ctx->deferSync = false; }
  } deferGuard = {ctx};
  bool pipelined = true;
  for (int i = 0; i < n; i++) pipelined = pipelined && sys[i]->pipelined;
  ctx->deferSync = pipelined;
  std::vector<lsd_frame *> fr(n, nullptr);
  struct Guard {  // frames that nobody has taken over are dropped on every exit
    lsd_ctx *ctx;
    std::vector<lsd_frame *> &f;
    ~Guard() { for (lsd_frame *x : f) if (x) lsd_frame_release(ctx, x); }
  } guard = {ctx, fr};
  if ((rc = lsd_frame_create_batch(ctx, n, ids, images, pitch, LSD_BUILD_TRACKING, fr.data()))) return rc;
  lap(0);

  // ---- SlamSystem::trackFrame: TrackingReference::importFrame where the keyframe or its depth changed
  {
    std::vector<lsd_frame *> kfs;
    std::vector<int> who;
    for (int i = 0; i < n; i++) {
      lsd_slam_system *s = sys[i];
      if (!s->ref || s->refKfId != s->kf->id || s->kf->depthHasBeenUpdatedFlag) {
        if (s->ref) lsd_ref_release(ctx, s->ref);
        s->ref = nullptr;
        kfs.push_back(s->kf);
        who.push_back(i);
      }
    }
    if (!kfs.empty()) {
      std::vector<lsd_ref *> refs(kfs.size(), nullptr);
      if ((rc = lsd_ref_create_batch(ctx, (int)kfs.size(), kfs.data(), refs.data()))) return rc;
      for (size_t k = 0; k < who.size(); k++) {
        lsd_slam_system *s = sys[who[k]];
        s->ref = refs[k];
        s->refKfId = s->kf->id;
        s->kf->depthHasBeenUpdatedFlag = false;
      }
    }
  }
  lap(1);
  std::vector<lsd_ref *> refs(n);
  std::vector<double> inits(7 * (size_t)n);
  std::vector<lsd_se3_result> res(n);
  for (int i = 0; i < n; i++) {
    refs[i] = sys[i]->ref;
    std::memcpy(&inits[7 * (size_t)i], sys[i]->lastToKf, sizeof(double) * 7);
  }
  if ((rc = lsd_se3_track_batch(ctx, n, refs.data(), fr.data(), inits.data(), res.data(), nullptr))) return rc;
  lap(2);

  // ---- per sequence: lost / keyframe decision.  The mean inverse depths the decision needs are fetched in one batch.
  std::vector<double> toKf(8 * (size_t)n);
  std::vector<int> needMean;
  for (int i = 0; i < n; i++) {
    lsd_slam_system *s = sys[i];
    if (res[i].diverged || !res[i].trackingWasGood) continue;
    if (s->kf->numMappedOnThis > MIN_NUM_MAPPED && !s->kfMeanValid) needMean.push_back(i);
  }
  if (!needMean.empty()) {
    std::vector<lsd_frame *> kfs;
    std::vector<float> means(needMean.size());
    for (int i : needMean) kfs.push_back(sys[i]->kf);
    if ((rc = lsd_frame_mean_idepth_batch(ctx, (int)kfs.size(), kfs.data(), means.data(), nullptr))) return rc;
    for (size_t k = 0; k < needMean.size(); k++) {
      sys[needMean[k]]->kfMeanIdepth = means[k];
      sys[needMean[k]]->kfMeanValid = true;
    }
  }
  std::vector<int> upd, sw;  // sequences that update their keyframe / that switch to a new one
  std::vector<float> score(n, 0.0f);
  for (int i = 0; i < n; i++) {
    lsd_slam_system *s = sys[i];
    if (res[i].diverged || !res[i].trackingWasGood) {
      s->lost++;
      fill_status(s, ids[i], 0, 0, s->lastToKf, &res[i], 0.0f, &st[i]);
      continue;  // the guard releases the frame
    }
    s->tracked++;
    double *p = &toKf[8 * (size_t)i];
    std::memcpy(p, res[i].frameToRef, sizeof(double) * 7);
    p[7] = 1.0;
    std::memcpy(s->lastToKf, p, sizeof(double) * 8);
    bool create = false;
    if (s->kf->numMappedOnThis > MIN_NUM_MAPPED) {
      const double m = (double)s->kfMeanIdepth;
      const double d[3] = {p[4] * m, p[5] * m, p[6] * m};
      score[i] = lsd_slam_ref_frame_score((float)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), res[i].pointUsage);
      create = score[i] > slam_min_val(s);
    }
    (create ? sw : upd).push_back(i);
  }
  lap(3);

  // ---- one mapping iteration per sequence, batched by kind: keyframe switches first (they synchronise), updates last (deferred)
  if (!sw.empty()) {
    std::vector<lsd_depthmap *> dms;
    std::vector<lsd_frame *> frames;
    for (int i : sw) {
      dms.push_back(sys[i]->dm);
      frames.push_back(fr[i]);
      fill_status(sys[i], ids[i], 1, 1, &toKf[8 * (size_t)i], &res[i], score[i], &st[i]);  // published with the pose it was tracked at
    }
    if ((rc = lsd_depth_finalize_keyframe_batch(ctx, (int)sw.size(), dms.data()))) return rc;
    if ((rc = lsd_depth_create_keyframe_batch(ctx, (int)sw.size(), dms.data(), frames.data(), nullptr))) return rc;
    for (int i : sw) {
      lsd_slam_system *s = sys[i];
      lsd_frame *f = fr[i];
      double world[8];
      sim3_mul(s->kfWorld, f->thisToParent_raw, world);
      std::memcpy(s->kfWorld, world, sizeof(world));
      lsd_ref_release(ctx, s->ref);
      s->ref = nullptr;
      if (s->keepFinishedKeyframes) s->keyframes.push_back(s->kf);
      else lsd_frame_release(ctx, s->kf);
      s->kf = f;
      fr[i] = nullptr;  // lives on as the current keyframe
      s->nKeyframes++;
      s->kfMeanValid = false;
      sim3_identity(s->lastToKf);
      st[i].numKeyframes = s->nKeyframes;
      st[i].currentKeyframeId = s->kf->id;
      st[i].keyframeRescale = f->thisToParent_raw[7];
    }
    lap(4);
  }
  if (!upd.empty()) {
    std::vector<lsd_depthmap *> dms;
    std::vector<lsd_frame *> frames;
    std::vector<char> setsDepth;
    for (int i : upd) {
      dms.push_back(sys[i]->dm);
      frames.push_back(fr[i]);
      setsDepth.push_back(!sys[i]->kf->depthHasBeenUpdatedFlag);
    }
    if ((rc = lsd_depth_update_keyframe_batch(ctx, (int)upd.size(), dms.data(), frames.data()))) return rc;
    for (size_t k = 0; k < upd.size(); k++) {
      const int i = upd[k];
      if (setsDepth[k]) sys[i]->kfMeanValid = false;
      fill_status(sys[i], ids[i], 1, 0, &toKf[8 * (size_t)i], &res[i], score[i], &st[i]);
    }
    lap(3);
  }
  return LSD_OK;  // the guard releases every frame that did not become a keyframe
}

int lsd_slam_current_keyframe(lsd_slam_system *s, lsd_frame **kf, lsd_depthmap **dm) {
  LSD_ARG(s);
  if (kf) *kf = s->kf;
  if (dm) *dm = s->dm;
  return LSD_OK;
}

int lsd_slam_stage_seconds(lsd_slam_system *s, double out[5]) {
  LSD_ARG(s && out);
  for (int i = 0; i < 5; i++) out[i] = s->stageSec[i];
  return LSD_OK;
}

int lsd_slam_counters(lsd_slam_system *s, int *tracked, int *lost, int *keyframes) {
  LSD_ARG(s);
  if (tracked) *tracked = s->tracked;
  if (lost) *lost = s->lost;
  if (keyframes) *keyframes = s->nKeyframes;
  return LSD_OK;
}

// one line of the reference's pose.txt: "id,tx,ty,tz,rawtx,rawty,rawtz\n" with ostream's default float formatting
// (/root/reference/lib/Pangolin_IOWrapper/TextOutputIOWrapper.cpp:100-120)
int lsd_slam_pose_line(const lsd_slam_status *st, char *buf, size_t n) {
  LSD_ARG(st && buf && n > 0);
  const int k = std::snprintf(buf, n, "%d,%g,%g,%g,%g,%g,%g\n", st->frameId, st->camToWorld[4], st->camToWorld[5], st->camToWorld[6],
                              st->thisToParent_raw[4], st->thisToParent_raw[5], st->thisToParent_raw[6]);
  LSD_ARG(k > 0 && (size_t)k < n);
  return LSD_OK;
}

}  // extern "C"
