/* lsd_b200.h -- C ABI of the B200-native LSD-SLAM hot path (liblsd_b200.so).
 *
 * The reference application (apl-ocean-engineering/lsd-slam-pangolin-gui) has no FFI for this
 * path: it links the C++ classes of the un-vendored `lsd-slam` core (fips.yml:1-4) and drives
 * them through lsd_slam::SlamSystem (tools/LSD.cpp:102, lib/App/InputThread.cpp:71).  The
 * drop-in boundary is therefore the set of lsd-slam core methods named in BASELINE.json;
 * each entry point below cites the core method it replaces ([UP] = upstream lsd-slam file,
 * source absent from /root/reference) and the reference-side line that evidences its contract.
 * C++ adapters with the upstream class signatures sit on top of this ABI in
 * lsd-slam-pangolin-gui_b200/host/ (see INTEGRATION.md).
 *
 * Conventions: every function returns 0 on success, <0 on error (no exceptions cross the ABI;
 * lsd_last_error() gives the text).  Handles are opaque and owned by their context.  All host
 * buffers are caller-owned.  Calls are blocking unless named *_async.  One context per calling
 * thread (upstream runs tracking / mapping / constraint search concurrently); frame and
 * reference handles may be shared read-only between contexts on the same device.
 *
 * Poses: SE3 = double[7] {qx,qy,qz,qw,tx,ty,tz} (Sophus::SE3d::data() order);
 *        Sim3 = double[8] {qx,qy,qz,qw,tx,ty,tz,scale}.
 */
#ifndef LSD_B200_H
#define LSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSD_PYRAMID_LEVELS 5

typedef struct lsd_ctx lsd_ctx;           /* one device, one stream, one set of scratch buffers   */
typedef struct lsd_frame lsd_frame;       /* [UP] lsd_slam::Frame -- device-resident pyramids       */
typedef struct lsd_ref lsd_ref;           /* [UP] lsd_slam::TrackingReference -- per-level points   */
typedef struct lsd_depthmap lsd_depthmap; /* [UP] lsd_slam::DepthMap -- SoA hypothesis planes       */
typedef struct lsd_undistorter lsd_undistorter; /* libvideoio::Undistorter -- OpenCV fixed-point remap maps */

enum lsd_status {
  LSD_OK = 0,
  LSD_ERR_ARG = -1,
  LSD_ERR_CUDA = -2,
  LSD_ERR_STATE = -3,
  LSD_ERR_NOMEM = -4
};

/* Frame::image/gradients/maxGradients/idepth/idepthVar(level), refPixelWasGood()
 * (read by the reference at lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:56-79). */
enum lsd_field {
  LSD_FIELD_IMAGE = 0,      /* float  [h_l*w_l]                                  */
  LSD_FIELD_GRADIENTS = 1,  /* float4 [h_l*w_l]  (gx, gy, I, 0)                  */
  LSD_FIELD_MAXGRAD = 2,    /* float  [h_l*w_l]  (level 0 only is ever consumed)  */
  LSD_FIELD_IDEPTH = 3,     /* float  [h_l*w_l]                                  */
  LSD_FIELD_IDEPTHVAR = 4,  /* float  [h_l*w_l]                                  */
  LSD_FIELD_MASK = 5        /* uint8  [(h>>1)*(w>>1)]  refPixelWasGood            */
};

/* Tracker tunables: [UP] DenseDepthTrackerSettings + the util/settings.h constants the kernels use. */
typedef struct lsd_tracker_settings {
  float lambdaSuccessFac, lambdaFailFac;
  float stepSizeMin[LSD_PYRAMID_LEVELS];
  float convergenceEps[LSD_PYRAMID_LEVELS];
  int maxItsPerLvl[LSD_PYRAMID_LEVELS];
  float lambdaInitial[LSD_PYRAMID_LEVELS];
  float var_weight, huber_d;
} lsd_tracker_settings;

/* What [UP] SE3Tracker exposes as members after trackFrame. */
typedef struct lsd_se3_result {
  double frameToRef[7];
  float lastResidual, lastMeanRes, pointUsage, lastGoodCount, lastBadCount;
  float affine_a, affine_b, initialTrackedResidual;
  int diverged, trackingWasGood;
  int numResidualCalls[LSD_PYRAMID_LEVELS], numWarpUpdateCalls[LSD_PYRAMID_LEVELS];
  int traceLen;
} lsd_se3_result;

/* One LM evaluation, recorded when a trace buffer is supplied (parity tests). */
typedef struct lsd_trace_entry {
  int level, accepted; /* -1 first evaluation of a level, 0 rejected, 1 accepted */
  float error, lambda;
  int bufSize;
} lsd_trace_entry;
#define LSD_TRACE_CAP 512

/* ---- context ------------------------------------------------------------------------------- */
const char *lsd_last_error(void);
int lsd_version(void);
/* stream: a cudaStream_t to run on (0/NULL: the context creates its own non-blocking stream).
 * K = {fx, fy, cx, cy} at level 0 ([UP] Frame ctor takes the undistorted camera matrix; the
 * reference passes undistorter->getCamera(), lib/App/InputThread.cpp:71). */
int lsd_ctx_create(int device, int width, int height, const float K[4], void *stream, lsd_ctx **out);
int lsd_ctx_destroy(lsd_ctx *ctx);
int lsd_ctx_synchronize(lsd_ctx *ctx);
void *lsd_ctx_stream(lsd_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py gpu_launches) */
long long lsd_ctx_launch_count(lsd_ctx *ctx);
int lsd_default_tracker_settings(lsd_tracker_settings *s);
int lsd_ctx_set_se3_settings(lsd_ctx *ctx, const lsd_tracker_settings *s);
/* scheduling knob of the persistent tracker: 4096-point records per work item (0 = automatic).  Never
 * changes a result (the summation order is fixed by the records), only latency vs throughput. */
int lsd_ctx_set_se3_work_item_records(lsd_ctx *ctx, int records);
/* Points per partial record of the SE3 tracker (0 = default 4096; multiple of 128).  A record is reduced by one CTA and
 * records are summed in order, so this value DEFINES the fp32 summation order: for a given value results are
 * bit-identical for every batch size and scheduling; between values they differ by reassociation only.  Large records
 * maximise batch throughput; a context that tracks ONE live sequence (SlamSystem's tracking thread) wants small records
 * (512): an evaluation then spreads over 8x the CTAs (measured: 0.51 -> 0.34 ms per tracked frame; the 1000-pair batch would
 * go from 4.1 to > 6 ms). */
int lsd_ctx_set_se3_record_points(lsd_ctx *ctx, int points);
/* Record size per pyramid level: points[l] for level l = 1 .. LSD_PYRAMID_LEVELS - 1 (points[0] is ignored: level 0 is never
 * tracked); 0 = the context-wide value of lsd_ctx_set_se3_record_points.  A live context wants every evaluation to be ONE
 * record per 128-thread group of its cluster with as few points per thread as the level allows, e.g. {0, 768, 256, 128, 128}
 * at 640x480: a coarse level then costs one point per thread instead of four.  Like the context-wide value this defines the
 * summation order and nothing else. */
int lsd_ctx_set_se3_record_points_per_level(lsd_ctx *ctx, const int *points);
/* Convenience for the context of SlamSystem's tracking thread (one frame per call): picks the per-level record sizes for this
 * image size -- ceil(0.6 * pixels(level) / 64) rounded up to a multiple of 128, i.e. {768, 256, 128, 128} at 640x480 and
 * {2944, 768, 256, 128} at 1280x960: one record per 128-thread group of a 16-CTA cluster as long as at most 60 % of a level's
 * pixels carry depth.  enable = 0 returns to the context-wide record size. */
int lsd_ctx_set_live_tracking(lsd_ctx *ctx, int enable);
/* SE3 tracking calls with at most `pairs` pairs run on the live kernel: ONE thread-block cluster per pair (16 CTAs x 512 threads
 * where the device co-schedules them, else 8), LM state in the leader CTA's shared memory, header and partial records exchanged
 * over distributed shared memory, two cluster barriers per LM evaluation instead of the work queue's global-memory hand-offs.
 * This is the shape of SlamSystem::trackFrame (one frame at a time).  Same records, same summation order: for a given record
 * size both kernels return identical bits.  -1 = default (2 while level 1 has at most 131072 pixels -- up to 724x724 images --,
 * else 0: a larger level 1 is better spread over the whole device by the work-queue kernel), 0 = always the work-queue kernel;
 * at most 8 pairs per live launch (larger batches take the work-queue kernel whatever the setting). */
int lsd_ctx_set_se3_live_pairs(lsd_ctx *ctx, int pairs);
/* Depth-map stencil kernels: bit 0 of `mask` = regularizeDepthMap, bit 1 = regularizeDepthMapFillHoles.  A set bit makes the
 * kernel fetch the halo tile of each CTA with the TMA unit (cp.async.bulk.tensor, out-of-map cells zero-filled by the
 * hardware), a clear bit with 16-byte vector loads.  Results are bit-identical; the default (1) is the faster choice per kernel
 * on B200.  Environment LSD_B200_STENCIL_TMA=<mask> overrides the default at context creation.  On a driver
 * without cuTensorMapEncodeTiled the mask is forced to 0 (vector loads). */
int lsd_ctx_set_stencil_tma(lsd_ctx *ctx, int mask);
/* pairs in flight inside one lsd_se3_track_batch launch (0 = default): bounds the working set to what L2 holds.
 * Scheduling only: results are bit-identical for every value. */
int lsd_ctx_set_se3_active_pairs(lsd_ctx *ctx, int pairs);

/* ---- Frame ---------------------------------------------------------------------------------- */
#define LSD_BUILD_TRACKING 0u /* image L0-4 + gradients L1-4: what a tracked frame needs        */
#define LSD_BUILD_MAXGRAD0 1u /* + maxGradients(0): what a keyframe / stereo KF needs            */
#define LSD_BUILD_GRAD0 2u    /* + gradients(0)                                                  */
/* [UP] Frame::Frame(id, w, h, K, timestamp, const unsigned char* image) + buildImage /
 * buildGradients / buildMaxGradients for all levels.  `image` is the 8-bit grey frame the
 * reference hands SlamSystem::nextImage (lib/App/InputThread.cpp:59,65,71); pitch in bytes. */
int lsd_frame_create(lsd_ctx *ctx, int id, const uint8_t *image, size_t pitch, unsigned flags, lsd_frame **out);
/* batch form: n frames from n host images in one pass (pinned staging, one launch per stage) */
int lsd_frame_create_batch(lsd_ctx *ctx, int n, const int *ids, const uint8_t *const *images, size_t pitch,
                           unsigned flags, lsd_frame **out);
/* same, source images already in device memory (contiguous n * h * w bytes) */
int lsd_frame_create_batch_device(lsd_ctx *ctx, int n, const int *ids, const void *d_images, unsigned flags,
                                  lsd_frame **out);
int lsd_frame_release(lsd_ctx *ctx, lsd_frame *f);
int lsd_frame_release_batch(lsd_ctx *ctx, int n, lsd_frame **f);
/* lazily builds whatever `field` at `level` needs, then copies it to host memory */
int lsd_frame_read(lsd_ctx *ctx, lsd_frame *f, int field, int level, void *dst);
int lsd_frame_num_mappable_pixels(lsd_ctx *ctx, lsd_frame *f, int *out); /* [UP] Frame::numMappablePixels */
/* [UP] Frame::setDepthFromGroundTruth(const float* depth, float cov_scale) */
int lsd_frame_set_depth_from_gt(lsd_ctx *ctx, lsd_frame *f, const float *depth, float cov_scale);
/* install level-0 idepth / idepthVar planes directly (what Frame::setDepth leaves behind) */
int lsd_frame_set_idepth(lsd_ctx *ctx, lsd_frame *f, const float *idepth, const float *idepthVar);
int lsd_frame_set_idepth_batch_device(lsd_ctx *ctx, int n, lsd_frame *const *f, const void *d_idepth,
                                      const void *d_idepthVar);
int lsd_frame_mean_idepth(lsd_ctx *ctx, lsd_frame *f, float *meanIdepth, int *numPoints);
/* n frames, one host synchronisation */
int lsd_frame_mean_idepth_batch(lsd_ctx *ctx, int n, lsd_frame *const *f, float *meanIdepth /* n or NULL */, int *numPoints /* n or NULL */);

/* ---- TrackingReference ---------------------------------------------------------------------- */
/* [UP] TrackingReference::importFrame + makePointCloud(level) for levels 1..4 */
int lsd_ref_create(lsd_ctx *ctx, lsd_frame *keyframe, lsd_ref **out);
int lsd_ref_create_batch(lsd_ctx *ctx, int n, lsd_frame *const *keyframes, lsd_ref **out);
int lsd_ref_release(lsd_ctx *ctx, lsd_ref *r);
int lsd_ref_num_data(lsd_ctx *ctx, lsd_ref *r, int level, int *out); /* [UP] numData[level] */
/* copies the level's point cloud in the reference's layout (pos 3f, grad 2f, colorAndVar 2f, idx);
 * emission order is row-major here (upstream: column-major) -- see DESIGN.md */
int lsd_ref_read(lsd_ctx *ctx, lsd_ref *r, int level, float *pos, float *grad, float *colorAndVar, int *idx);

/* ---- SE3Tracker ----------------------------------------------------------------------------- */
/* [UP] SE3Tracker::trackFrame(TrackingReference*, Frame*, const SE3& frameToReference_initialEstimate).
 * Writes refPixelWasGood of `frame`, frame->initialTrackedResidual, pose->thisToParent_raw. */
int lsd_se3_track(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double init_frameToRef[7], lsd_se3_result *result,
                  lsd_trace_entry *trace /* LSD_TRACE_CAP entries or NULL */);
/* n independent (ref, frame) pairs in ONE persistent launch (config 2 of BASELINE.json) */
int lsd_se3_track_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames,
                        const double *init_frameToRef /* n*7 */, lsd_se3_result *results,
                        lsd_trace_entry *traces /* n*LSD_TRACE_CAP or NULL */);
/* host images in, poses out.  Three streams: the copy engine moves 48-frame chunks, the context's stream builds their pyramids
 * and feeds the pairs into ONE persistent tracker that was started up front on a third stream and leaves one CTA slot per SM to
 * the ingest kernels (LSD_B200_E2E_STREAM=0: one tracker launch per 250-frame chunk).  Same poses as lsd_se3_track_batch. */
int lsd_se3_track_images_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, const uint8_t *const *images, size_t pitch,
                               const double *init_frameToRef, lsd_se3_result *results);
/* scheduling of lsd_se3_track_images_batch (never changes a result): frames per H2D copy / ingest launch (0 = default);
 * streamed = 1 one persistent tracker fed chunk by chunk, 0 one tracker launch per chunk, -1 automatic (streamed unless
 * kernels are known to be serialised: CUDA_LAUNCH_BLOCKING, profiler / sanitizer injection); watchdogSeconds (0 = keep, default
 * 2 s): a streamed tracker that waits that long for work stops itself and the batch is re-run with an ordinary launch. */
int lsd_ctx_set_image_pipeline(lsd_ctx *ctx, int chunkFrames, int streamed, double watchdogSeconds);
/* [UP] SE3Tracker::trackFrameOnPermaref(reference, frame, referenceToFrame): the quick single-level test track
 * (QUICK_KF_CHECK_LVL = 4) the constraint search and the Relocalizer run against a keyframe's permanent reference,
 * n candidates in ONE launch.  init and results[i].frameToRef both hold referenceToFrame
 * (upstream returns it un-inverted); the frames, their masks and the keyframes' counters are left untouched.
 * Settings: the "test track" members of DenseDepthTrackerSettings (lsd_default_permaref_settings). */
int lsd_default_permaref_settings(lsd_tracker_settings *s);
int lsd_ctx_set_permaref_settings(lsd_ctx *ctx, const lsd_tracker_settings *s);
int lsd_se3_track_permaref_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames,
                                 const double *init_refToFrame /* n*7 */, lsd_se3_result *results,
                                 lsd_trace_entry *traces /* n*LSD_TRACE_CAP or NULL */);
/* [UP] SE3Tracker::checkPermaRefOverlap: mean min(1, z_ref / z') over the level-4 points that land inside the image */
int lsd_se3_check_permaref_overlap_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, const double *refToFrame /* n*7 */,
                                         float *pointUsage /* n */);
/* one fused LM evaluation (calcResidualAndBuffers + calcWeightsAndResidual + calculateWarpUpdate)
 * at a fixed pose; A36/b6 as after NormalEquationsLeastSquares::finish(); scalars[12] =
 * {error, meanSqRes, bufSize, good, bad, pointUsage, meanRes, a_lastIt, b_lastIt, lsError, 0, 0} */
int lsd_se3_eval(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double refToFrame[7], int level, float affine_a,
                 float affine_b, float *A36, float *b6, float *scalars);
/* algorithmic bytes and evaluation count of the last lsd_se3_track* call on this ctx */
int lsd_se3_last_stats(lsd_ctx *ctx, double *algorithmic_bytes, long long *evaluations, float *kernel_ms);

/* ---- Sim3Tracker ------------------------------------------------------------------------------------ */
/* What [UP] Sim3Tracker exposes as members after trackFrameSim3. */
typedef struct lsd_sim3_result {
  double frameToRef[8];      /* Sim3 {qx,qy,qz,qw,tx,ty,tz,scale}; identity when diverged (upstream returns Sim3()) */
  float lastSim3Hessian[49]; /* 7x7, row-major: ls7.A of the last LGS (not divided by num_constraints)           */
  float lastResidual, lastDepthResidual, lastPhotometricResidual, pointUsage, affine_a, affine_b;
  int diverged;
  int numResidualCalls[LSD_PYRAMID_LEVELS], numWarpUpdateCalls[LSD_PYRAMID_LEVELS];
  int traceLen;
} lsd_sim3_result;

int lsd_ctx_set_sim3_settings(lsd_ctx *ctx, const lsd_tracker_settings *s);
/* Points per partial record of the Sim3 tracker (0 = default 1024; multiple of 128): a record is reduced by one CTA and the
 * records are summed in order, so this value DEFINES the fp32 summation order (see lsd_ctx_set_se3_record_points). */
int lsd_ctx_set_sim3_record_points(lsd_ctx *ctx, int points);
/* [UP] Sim3Tracker::trackFrameSim3(TrackingReference*, Frame*, const Sim3& frameToReference_initialEstimate,
 * int startLevel, int finalLevel).  `frame` must carry depth (it is a keyframe): its idepth pyramid is read. */
int lsd_sim3_track(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double init_frameToRef[8], int startLevel, int finalLevel,
                   lsd_sim3_result *result, lsd_trace_entry *trace /* LSD_TRACE_CAP entries or NULL */);
/* n independent tracks in ONE persistent launch (the constraint search of SlamSystem::findConstraintsForNewKeyFrames:
 * BASELINE.json configs[3]); results are bit-identical for every batch composition */
int lsd_sim3_track_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef /* n*8 */,
                         int startLevel, int finalLevel, lsd_sim3_result *results, lsd_trace_entry *traces /* n*LSD_TRACE_CAP or NULL */);

/* [UP] SlamSystem::tryTrackSim3 chains trackFrameSim3 calls over level ranges ([4,3], [2], [1] in testConstraint), each starting
 * from the previous call's result.  Here every track runs its whole chain inside the ONE persistent launch: stage k+1 starts
 * from stage k's frameToRef exactly as a separate call would (bit-identical), a track that diverged / came back with an empty
 * information matrix (upstream's rejection test) stops and reports diverged for its remaining stages.
 * results: nStages * n, stage-major (results[k * n + i] = what the k-th call of track i would have returned). */
int lsd_sim3_track_stages_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef /* n*8 */,
                                int nStages /* <= 4 */, const int *startLevels, const int *finalLevels, lsd_sim3_result *results);

/* ---- frame bookkeeping the mapping side reads ---------------------------------------------------- */
/* What SE3Tracker::trackFrame leaves on a tracked Frame ([UP] frame->pose->thisToParent_raw,
 * trackingParent, initialTrackedResidual); settable directly for frames whose pose comes from elsewhere
 * (ground truth, pose graph).  toParent = Sim3 double[8] {qx,qy,qz,qw,tx,ty,tz,scale}. */
int lsd_frame_set_tracking_meta(lsd_ctx *ctx, lsd_frame *f, int parentId, const double toParent[8],
                                float initialTrackedResidual);
int lsd_frame_get_tracking_meta(lsd_ctx *ctx, lsd_frame *f, int *parentId, double toParent[8], float *initialTrackedResidual);
/* refPixelWasGood(): install (mask != NULL, (h>>1)*(w>>1) bytes) or drop (NULL: refPixelWasGoodNoCreate() == 0) */
int lsd_frame_set_mask(lsd_ctx *ctx, lsd_frame *f, const uint8_t *mask);
/* [UP] Frame::numFramesTrackedOnThis / numMappedOnThis (feed observeDepth's skip-ahead) */
int lsd_frame_set_counters(lsd_ctx *ctx, lsd_frame *f, int numFramesTrackedOnThis, int numMappedOnThis);
int lsd_frame_get_counters(lsd_ctx *ctx, lsd_frame *f, int *numFramesTrackedOnThis, int *numMappedOnThis);
int lsd_frame_set_depth_updated_flag(lsd_ctx *ctx, lsd_frame *f, int depthHasBeenUpdatedFlag);
int lsd_frame_get_depth_updated_flag(lsd_ctx *ctx, lsd_frame *f, int *depthHasBeenUpdatedFlag);

/* ---- DepthMap -------------------------------------------------------------------------------------- */
/* [UP] DepthMapPixelHypothesis, upstream's 32-byte AoS layout; the device keeps SoA planes. */
typedef struct lsd_hypothesis {
  uint8_t isValid, pad_[3];
  int32_t blacklisted;
  float nextStereoFrameMinID;
  int32_t validity_counter;
  float idepth, idepth_var, idepth_smoothed, idepth_var_smoothed;
} lsd_hypothesis;

/* util/settings.h thresholds that are run-time settable here (DESIGN.md "Deviations from SURVEY") */
typedef struct lsd_depth_settings {
  int valSumMinForCreate;      /* VAL_SUM_MIN_FOR_CREATE      default 30  */
  int valSumMinForKeep;        /* VAL_SUM_MIN_FOR_KEEP        default 24  */
  int valSumMinForUnblacklist; /* VAL_SUM_MIN_FOR_UNBLACKLIST default 100 */
  int minBlacklist;            /* MIN_BLACKLIST               default -1  */
} lsd_depth_settings;

enum lsd_depth_stage {
  LSD_STAGE_OBSERVE = 0,    /* DepthMap::observeDepth (references from the last lsd_depth_prepare)       */
  LSD_STAGE_FILL_HOLES = 1, /* DepthMap::regularizeDepthMapFillHoles                                      */
  LSD_STAGE_REGULARIZE = 2, /* DepthMap::regularizeDepthMap(arg1 = removeOcclusions, arg2 = validityTH)   */
  LSD_STAGE_PROPAGATE = 3,  /* DepthMap::propagateDepth(frame); the frame becomes the active keyframe      */
  LSD_STAGE_SET_DEPTH = 4   /* activeKeyFrame->setDepth(currentDepthMap) + idepth pyramids                */
};

int lsd_depthmap_create(lsd_ctx *ctx, lsd_depthmap **out); /* [UP] DepthMap::DepthMap(w, h, K) */
int lsd_depthmap_destroy(lsd_ctx *ctx, lsd_depthmap *dm);
int lsd_default_depth_settings(lsd_depth_settings *s);
int lsd_depthmap_set_settings(lsd_ctx *ctx, lsd_depthmap *dm, const lsd_depth_settings *s);
/* [UP] DepthMap::initializeFromGTDepth(Frame*): the frame must carry depth (lsd_frame_set_depth_from_gt) */
int lsd_depth_initialize_from_gt(lsd_ctx *ctx, lsd_depthmap *dm, lsd_frame *kf);
/* [UP] DepthMap::initializeRandomly(Frame*): consumes libc rand() in raster order exactly like upstream
 * (one draw per pixel with maxGradients > MIN_ABS_GRAD_CREATE, x in [1,w-1), y in [1,h-1)) */
int lsd_depth_initialize_randomly(lsd_ctx *ctx, lsd_depthmap *dm, lsd_frame *kf);
/* install an explicit hypothesis map (w*h lsd_hypothesis) for `kf` -- [UP] DepthMap::setFromExistingKF /
 * re-activation data, and the way parity runs inject identical state into the oracle and the device */
int lsd_depth_initialize_from_map(lsd_ctx *ctx, lsd_depthmap *dm, lsd_frame *kf, const lsd_hypothesis *map, int reactivated);
/* [UP] DepthMap::updateKeyframe(std::deque<std::shared_ptr<Frame>> referenceFrames).  refToKf: n*8 Sim3
 * (frame -> active keyframe) or NULL to use each frame's thisToParent_raw (frames tracked on the active keyframe). */
int lsd_depth_update_keyframe(lsd_ctx *ctx, lsd_depthmap *dm, int n, lsd_frame *const *referenceFrames, const double *refToKf);
/* [UP] DepthMap::createKeyFrame(Frame* new_keyframe): propagateDepth + regularize(true) + fillHoles +
 * regularize(false) + mean-idepth normalisation + setDepth; rescaleFactor is folded into new_keyframe's
 * thisToParent_raw exactly like upstream and also returned. */
int lsd_depth_create_keyframe(lsd_ctx *ctx, lsd_depthmap *dm, lsd_frame *new_keyframe, float *rescaleFactor);
int lsd_depth_finalize_keyframe(lsd_ctx *ctx, lsd_depthmap *dm); /* [UP] DepthMap::finalizeKeyFrame */
/* The three calls above on n independent depth maps of one context (the maps of n live sequences): every stage is ONE launch
 * over all maps and the host synchronises once per call.  Map i is handled exactly as by the single-map call (bit-identical
 * results); updateKeyframe takes one reference frame per map (the frame just tracked on that map's keyframe).  All maps must
 * share the same lsd_depth_settings. */
int lsd_depth_update_keyframe_batch(lsd_ctx *ctx, int n, lsd_depthmap *const *dms, lsd_frame *const *referenceFrames);
int lsd_depth_create_keyframe_batch(lsd_ctx *ctx, int n, lsd_depthmap *const *dms, lsd_frame *const *new_keyframes,
                                    float *rescaleFactors /* n or NULL */);
int lsd_depth_finalize_keyframe_batch(lsd_ctx *ctx, int n, lsd_depthmap *const *dms);
int lsd_depth_read(lsd_ctx *ctx, lsd_depthmap *dm, lsd_hypothesis *dst); /* currentDepthMap, w*h entries */
/* [UP] DepthMap::debugPlotDepthMap -> the w*h*3 bytes lib/GUI.cpp:104-108 (updateDepthImage) consumes */
int lsd_depth_debug_rgb(lsd_ctx *ctx, lsd_depthmap *dm, uint8_t *rgb);
/* stage-level entry points (parity tests, per-kernel timing: BASELINE.json configs[2]) */
int lsd_depth_prepare(lsd_ctx *ctx, lsd_depthmap *dm, int n, lsd_frame *const *referenceFrames, const double *refToKf);
int lsd_depth_stage(lsd_ctx *ctx, lsd_depthmap *dm, int stage, int arg1, int arg2, lsd_frame *frame);
/* the same stage on n independent depth maps in ONE set of launches (blockIdx.z = map); frames: n or NULL */
/* device time (CUDA events on the context's stream) of the kernels of the last lsd_depth_stage* call */
int lsd_ctx_last_stage_ms(lsd_ctx *ctx, float *ms);
int lsd_depth_stage_batch(lsd_ctx *ctx, int n, lsd_depthmap *const *dms, int stage, int arg1, int arg2, lsd_frame *const *frames);

/* ---- keyframe publish / point-cloud extraction ---------------------------------------- */
/* InputPointDense, /root/reference/lib/Pangolin_IOWrapper/Keyframe.h:16-21 */
typedef struct lsd_input_point_dense {
  float idepth, idepth_var;
  unsigned char color[4];
} lsd_input_point_dense;
/* Keyframe::MyVertex, Keyframe.h:47-51 */
typedef struct lsd_vertex {
  float point[3];
  unsigned char color[4];
} lsd_vertex;
/* the locals of Keyframe::computeVbo (Keyframe.h:79-84).  contractFma = 1 evaluates x*fxi+cxi as one FMA, which is
 * what gcc emits for the reference's Release flags (-march=native, CMakeLists.txt:59) on an FMA host; 0 = IEEE. */
typedef struct lsd_vbo_params {
  float scaledTH, absTH; /* my_scaledTH 1e-3, my_absTH 1e-1 */
  int minNearSupport;    /* 9 */
  int sparsifyFactor;    /* 1 (the reference's constant; anything else is rejected) */
  int contractFma;
} lsd_vbo_params;
int lsd_default_vbo_params(lsd_vbo_params *p);
/* PangolinOutputIOWrapper::publishKeyframe's pack loop (PangolinOutputIOWrapper.cpp:69-89): writes w_l*h_l records
 * into dst (what the reference stores in Keyframe::pointData).  A frame without depth yields zeros. */
int lsd_frame_publish_keyframe(lsd_ctx *ctx, lsd_frame *f, int level, lsd_input_point_dense *dst);
/* Keyframe::computeVbo (Keyframe.h:66-158) straight from the frame's planes: dst receives `*points` vertices in the
 * reference's raster order (capacity w_l*h_l).  camToWorldScale = Keyframe::camToWorld.scale(). */
int lsd_keyframe_compute_vbo(lsd_ctx *ctx, lsd_frame *f, int level, float camToWorldScale, const lsd_vbo_params *params,
                             lsd_vertex *dst, int *points);
/* n keyframes in ONE launch.  d_vertices: device memory for n * w_l*h_l vertices (e.g. a mapped GL buffer; keyframe i
 * starts at i*w_l*h_l) or NULL to use context scratch; dst: n host pointers (entries may be NULL) or NULL. */
int lsd_keyframe_compute_vbo_batch(lsd_ctx *ctx, int n, lsd_frame *const *frames, int level, const float *camToWorldScale,
                                   const lsd_vbo_params *params, void *d_vertices, lsd_vertex *const *dst, int *points);

/* ---- lock-step SlamSystem driver -------------------------------------------------------- */
/* The part of [UP] lsd_slam::SlamSystem the reference application drives (new SlamSystem() tools/LSD.cpp:102,
 * system->nextImage(idx, image, camera) lib/App/InputThread.cpp:71) with its `runRealTime == false` semantics:
 * nextImage returns once the frame is tracked AND mapped.  Tracking, mapping and keyframe selection run through the
 * entry points above; pose graph, loop closure and relocalisation stay on the reference's CPU code. */
typedef struct lsd_slam_system lsd_slam_system;
typedef struct lsd_slam_status {
  int frameId;
  int tracked;            /* 0: tracking lost on this frame (upstream would start the Relocalizer) */
  int isKeyframe;         /* this frame became the current keyframe */
  int numKeyframes, currentKeyframeId;
  int trackingWasGood, diverged;
  float pointUsage, lastResidual, keyframeScore;
  double camToWorld[8];       /* Frame::getCamToWorld(): what publishPose / pose.txt columns 2-4 carry */
  double thisToParent_raw[8]; /* frame -> keyframe (pose.txt columns 5-7)                              */
  double keyframeRescale;     /* isKeyframe: the mean-idepth rescale createKeyFrame folded into the new keyframe's pose */
} lsd_slam_status;
int lsd_slam_create(lsd_ctx *ctx, lsd_slam_system **out); /* SlamSystem::SlamSystem() */
/* [UP] TrackableKeyFrameSearch::getRefFrameScore(distanceSquared, usage) = distSq*KFDistWeight^2 + (1-usage)^2*KFUsageWeight^2
 * (KFDistWeight 4, KFUsageWeight 3): the closeness score nextImage compares with minVal to decide on a new keyframe */
float lsd_slam_ref_frame_score(float distanceSquared, float usage);
int lsd_slam_destroy(lsd_slam_system *s);                 /* fullReset() = destroy + create */
/* keep finished keyframes alive (upstream: KeyFrameGraph::keyframesAll) -- default 1; 0 frees them for long benches */
int lsd_slam_set_keep_keyframes(lsd_slam_system *s, int keep);
/* Pipelined stages inside lsd_slam_next_image (default 1): ingest, TrackingReference::importFrame and the tracker are queued back
 * to back and synchronised once, by the tracker's result; updateKeyframe is queued and completed by the next call into this
 * context that needs one of its results or rewrites a table it reads (at the latest the next image, after that image has been
 * staged into pinned memory).  Stream order is unchanged, so results are bit-identical; the caller-visible semantics stay
 * blocking.  0 = every stage synchronises on its own (the mapping of frame k is complete when the call returns). */
int lsd_slam_set_pipelined(lsd_slam_system *s, int enable);
/* optional: images handed to nextImage are still distorted and go through `und` first (undistort + ingest fused:
 * lib/App/InputThread.cpp:59-71 in one call); NULL = images are already undistorted */
int lsd_slam_set_undistorter(lsd_slam_system *s, lsd_undistorter *und);
/* SlamSystem::gtDepthInit / randomInit on the first image (nextImage on an empty system = randomInit) */
int lsd_slam_gt_depth_init(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, const float *depth, lsd_slam_status *st);
int lsd_slam_random_init(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, lsd_slam_status *st);
/* SlamSystem::nextImage: 8-bit grey image as handed over at lib/App/InputThread.cpp:71 */
int lsd_slam_next_image(lsd_slam_system *s, int id, const uint8_t *image, size_t pitch, lsd_slam_status *st);
/* n independent live sequences that share ONE context advance by one image each: SlamSystem::nextImage for every system, with
 * every stage batched over the sequences (one H2D + ingest, one tracking-reference import, ONE tracker launch for the n
 * pairs, one set of DepthMap launches for the sequences that update their keyframe and one for those that switch it).
 * Per sequence the results are bit-identical to n separate lsd_slam_next_image calls; what changes is that the GPU sees
 * n x the work per launch (a single 640x480 sequence keeps ~2 % of a B200 busy).  Every system must be initialised
 * (gt_depth_init / random_init) and have no undistorter attached. */
int lsd_slam_next_image_batch(int n, lsd_slam_system *const *systems, const int *ids, const uint8_t *const *images, size_t pitch,
                              lsd_slam_status *st /* n */);
int lsd_slam_current_keyframe(lsd_slam_system *s, lsd_frame **kf, lsd_depthmap **dm);
int lsd_slam_counters(lsd_slam_system *s, int *tracked, int *lost, int *keyframes);
/* host wall time accumulated inside nextImage, by stage: {frame ingest, reference import, tracking, updateKeyframe
 * (incl. keyframe selection), keyframe switch (finalize + createKeyFrame)} */
int lsd_slam_stage_seconds(lsd_slam_system *s, double out[5]);
/* "id,tx,ty,tz,rawtx,rawty,rawtz\n" as written by TextOutputIOWrapper::publishTrackedFrame
 * (lib/Pangolin_IOWrapper/TextOutputIOWrapper.cpp:100-120) */
int lsd_slam_pose_line(const lsd_slam_status *st, char *buf, size_t n);

/* ---- undistortion in front of the ingest ------------------------------------------------ */
/* libvideoio::Undistorter as the reference uses it: created from the calibration file (tools/LSD.cpp:88),
 * undistorter->undistort(image, imageUndist) per frame (lib/App/InputThread.cpp:61-65).  Its arithmetic is OpenCV's
 * fixed-point remap: map1 = int16 (x, y) pairs, map2 = uint16 sub-pixel index (fy*32 + fx), as produced by
 * cv::initUndistortRectifyMap(..., CV_16SC2); output size = the context's width x height. */
int lsd_undistorter_create_from_maps(lsd_ctx *ctx, int inWidth, int inHeight, const int16_t *map1, const uint16_t *map2,
                                     lsd_undistorter **out);
/* builds the maps itself: K = {fx, fy, cx, cy} of the distorted camera, dist = {k1, k2, p1, p2, k3}, Kout = the
 * undistorted camera (what undistorter->getCamera() returns and the context was created with) */
int lsd_undistorter_create_opencv(lsd_ctx *ctx, int inWidth, int inHeight, const double K[4], const double dist[5],
                                  const double Kout[4], lsd_undistorter **out);
int lsd_undistorter_destroy(lsd_ctx *ctx, lsd_undistorter *u);
int lsd_undistorter_maps(lsd_ctx *ctx, lsd_undistorter *u, int16_t *map1, uint16_t *map2);
/* Undistorter::undistort: 8-bit grey in (inWidth x inHeight, pitch in bytes), 8-bit grey out (width x height) */
int lsd_undistort(lsd_ctx *ctx, lsd_undistorter *u, const uint8_t *image, size_t pitch, uint8_t *undistorted);
/* undistort + Frame construction without the host round trip; `undistorted` (optional) receives the image the GUI
 * shows (output->updateLiveImage(imageUndist), InputThread.cpp:78) */
int lsd_frame_create_undistorted(lsd_ctx *ctx, lsd_undistorter *u, int id, const uint8_t *image, size_t pitch, unsigned flags,
                                 uint8_t *undistorted, lsd_frame **out);
int lsd_frame_create_undistorted_batch(lsd_ctx *ctx, lsd_undistorter *u, int n, const int *ids, const uint8_t *const *images,
                                       size_t pitch, unsigned flags, uint8_t *const *undistorted, lsd_frame **out);

#ifdef __cplusplus
}
#endif
#endif /* LSD_B200_H */
