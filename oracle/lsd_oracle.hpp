// oracle/lsd_oracle.hpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// CPU restatement ("oracle") of the LSD-SLAM per-pixel direct-alignment and depth
// filtering hot path that apl-ocean-engineering/lsd-slam-pangolin-gui drives through
// lsd_slam::SlamSystem (/root/reference/tools/LSD.cpp:102, lib/App/InputThread.cpp:71).
//
// WHY "PARITY UNPINNED": the arithmetic lives in the un-vendored fips import
//   lsd-slam = https://github.com/apl-ocean-engineering/lsd-slam.git  (branch `unstable`,
//   no commit pinned: /root/reference/fips.yml:1-4)
// which is absent from /root/reference and cannot be fetched (no network).  The reference
// holds no golden vector / fixture / known-answer test for this path (its only test is
// test/unit/test_test.cpp:4-7, ASSERT_TRUE(true)).  This file therefore restates the
// PUBLISHED algorithm (Engel, Schoeps, Cremers: "LSD-SLAM", ECCV 2014; upstream files
// DataStructures/Frame.cpp, Tracking/{SE3Tracker,Sim3Tracker,TrackingReference,
// least_squares}.cpp, DepthEstimation/DepthMap.cpp, util/globalFuncs.h, util/settings.h) as
// specified in SURVEY.md section 8a + Appendix A, and is anchored on analytic ground truth
// (tests/test_oracle_*.py).  Decisions taken where the spec is silent are marked DECISION.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this code.  The product (lsd-slam-pangolin-gui_b200/) never includes or links it.
//
// Build:  parity  g++ -O2 -ffp-contract=off -fno-fast-math   (oracle/Makefile)
//         timing  g++ -O3 -march=x86-64-v3                   (mirrors the reference's Release
//                 flags -O3 -march=native -DENABLE_SSE, /root/reference/CMakeLists.txt:59)
#pragma once
#include <atomic>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <vector>

#include "lie.hpp"

namespace lsdo {

constexpr int NL = 5;  // PYRAMID_LEVELS

// ---- util/settings.h constants -------------------------------------
constexpr int SE3TRACKING_MAX_LEVEL = 5;
constexpr int SE3TRACKING_MIN_LEVEL = 1;
constexpr int SIM3TRACKING_MAX_LEVEL = 5;
constexpr int SIM3TRACKING_MIN_LEVEL = 1;
constexpr int QUICK_KF_CHECK_LVL = 4;
constexpr float MIN_USE_GRAD = 5.0f;  // minUseGrad = MIN_ABS_GRAD_CREATE = MIN_ABS_GRAD_DECREASE
constexpr float CAMERA_PIXEL_NOISE2 = 4.0f * 4.0f;
constexpr float MIN_DEPTH = 0.05f;
constexpr float MAX_VAR = 0.5f * 0.5f;
constexpr float VAR_RANDOM_INIT_INITIAL = 0.5f * MAX_VAR;
constexpr float VAR_GT_INIT_INITIAL = 0.01f * 0.01f;
constexpr float SUCC_VAR_INC_FAC = 1.01f;
constexpr float FAIL_VAR_INC_FAC = 1.1f;
constexpr int VALIDITY_COUNTER_MAX = 5;
constexpr int VALIDITY_COUNTER_MAX_VARIABLE = 250;
constexpr int VALIDITY_COUNTER_INC = 5;
constexpr int VALIDITY_COUNTER_DEC = 5;
constexpr int VALIDITY_COUNTER_INITIAL_OBSERVE = 5;
constexpr int VAL_SUM_MIN_FOR_CREATE = -1;
constexpr int VAL_SUM_MIN_FOR_KEEP = 0;
constexpr int VAL_SUM_MIN_FOR_UNBLACKLIST = 100;
constexpr int MIN_BLACKLIST = -1;
constexpr float REG_DIST_VAR = 0.075f * 0.075f;  // * depthSmoothingFactor^2 (=1)
constexpr float DIFF_FAC_SMOOTHING = 1.0f;
constexpr float DIFF_FAC_OBSERVE = 1.0f;
constexpr float DIFF_FAC_PROP_MERGE = 1.0f;
constexpr float STEREO_EPL_VAR_FAC = 2.0f;
constexpr float GRADIENT_SAMPLE_DIST = 1.0f;
constexpr float SAMPLE_POINT_TO_BORDER = 7.0f;
constexpr float MIN_EPL_LENGTH_SQUARED = 1.0f;
constexpr float MIN_EPL_GRAD_SQUARED = 2.0f * 2.0f;
constexpr float MIN_EPL_ANGLE_SQUARED = 0.3f * 0.3f;
constexpr float MIN_EPL_LENGTH_CROP = 3.0f;
constexpr float MAX_EPL_LENGTH_CROP = 30.0f;
constexpr float MAX_ERROR_STEREO = 1300.0f;
constexpr float MIN_DISTANCE_ERROR_STEREO = 1.5f;
constexpr float MAX_DIFF_CONSTANT = 40.0f * 40.0f;
constexpr float MAX_DIFF_GRAD_MULT = 0.5f * 0.5f;
constexpr float MIN_GOODPERGOODBAD_PIXEL = 0.5f;
constexpr float MIN_GOODPERALL_PIXEL = 0.04f;
constexpr float MIN_GOODPERALL_PIXEL_ABSMIN = 0.01f;
constexpr float DIVISION_EPS = 1e-10f;
inline float UNZERO(float v) { return v < 0 ? (v > -1e-10f ? -1e-10f : v) : (v < 1e-10f ? 1e-10f : v); }

struct TrackerSettings {  // DenseDepthTrackerSettings
  float lambdaSuccessFac = 0.5f;
  float lambdaFailFac = 2.0f;
  float stepSizeMin[NL] = {1e-8f, 1e-8f, 1e-8f, 1e-8f, 1e-8f};
  float convergenceEps[NL] = {0.999f, 0.999f, 0.999f, 0.999f, 0.999f};
  int maxItsPerLvl[NL] = {5, 20, 50, 100, 100};
  float lambdaInitial[NL] = {0, 0, 0, 0, 0};
  float var_weight = 1.0f;
  float huber_d = 3.0f;
};

// ---- DepthMapPixelHypothesis (32-byte AoS) --------------------------
struct Hypothesis {
  bool isValid = false;
  int blacklisted = 0;
  float nextStereoFrameMinID = 0;
  int validity_counter = 0;
  float idepth = 0, idepth_var = 0, idepth_smoothed = 0, idepth_var_smoothed = 0;
  Hypothesis() {}
  Hypothesis(float id, float id_s, float var, float var_s, int val)
      : isValid(true), blacklisted(0), nextStereoFrameMinID(0), validity_counter(val), idepth(id), idepth_var(var),
        idepth_smoothed(id_s), idepth_var_smoothed(var_s) {}
  Hypothesis(float id, float var, int val)
      : isValid(true), blacklisted(0), nextStereoFrameMinID(0), validity_counter(val), idepth(id), idepth_var(var),
        idepth_smoothed(-1), idepth_var_smoothed(-1) {}
};

// DECISION: when set, the two whole-map sums of the depth map (createKeyFrame's sumIdepth and Frame::setDepth's
// sumIdepth) are accumulated in fp64 and rounded once -- the order-independent definition of the same
// quantity (upstream: sequential fp32).  Parity tests switch it on so that maps can be compared bit for bit.
extern bool g_exactSums;

// ---- Frame (DataStructures/Frame.cpp) -----------------------------------------------
struct Frame {
  int id = 0;
  int w[NL], h[NL];
  float fx[NL], fy[NL], cx[NL], cy[NL], fxi[NL], fyi[NL], cxi[NL], cyi[NL];
  std::vector<float> image[NL];     // f32
  std::vector<float> grad[NL];      // 4 floats / px: gx, gy, I, 0
  std::vector<float> maxGrad[NL];   // f32
  std::vector<float> idepth[NL], idepthVar[NL];
  bool imageValid[NL] = {}, gradValid[NL] = {}, maxGradValid[NL] = {}, idepthValid[NL] = {};
  bool hasIDepthBeenSet = false;
  std::vector<uint8_t> refPixelWasGood;  // (w>>1)*(h>>1), created 0xFF on first use
  int numMappablePixels = -1;
  float meanIdepth = 1;
  int numPoints = 0;
  float initialTrackedResidual = 0;
  int numFramesTrackedOnThis = 0, numMappedOnThis = 0, numMappedOnThisTotal = 0;
  int trackingParentId = -1;
  Sim3<double> thisToParent_raw;
  bool depthHasBeenUpdatedFlag = false;
  bool isReactivated = false;  // DepthMap: activeKeyFramelock re-activation

  // prepareForStereoWith cache (A7)
  int referenceID = -1, referenceLevel = -1;
  float distSquared = 0;
  Mat3<float> K_otherToThis_R;
  Vec3<float> K_otherToThis_t, otherToThis_t, K_thisToOther_t, thisToOther_t;
  Mat3<float> thisToOther_R;
  Vec3<float> otherToThis_R_row0, otherToThis_R_row1, otherToThis_R_row2;

  Frame(int id, int width, int height, float fx, float fy, float cx, float cy, const uint8_t *img);
  void buildImage(int level);
  void buildGradients(int level);
  void buildMaxGradients(int level);
  void buildIDepthAndIDepthVar(int level);
  void requireImage(int l) { if (!imageValid[l]) buildImage(l); }
  void requireGradients(int l) { if (!gradValid[l]) buildGradients(l); }
  void requireMaxGradients(int l) { if (!maxGradValid[l]) buildMaxGradients(l); }
  void requireIDepth(int l) { if (!idepthValid[l]) buildIDepthAndIDepthVar(l); }
  void setDepth(const Hypothesis *map);
  void setDepthFromGroundTruth(const float *depth, float cov_scale);
  // DECISION: convenience used by the microbench configs -- install a semi-dense level-0
  // idepth/var map directly (what setDepth produces from a converged DepthMap).
  void setIDepthRaw(const float *idepth, const float *var);
  uint8_t *refPixelWasGoodBuf();
  void prepareForStereoWith(const Frame *other, const Sim3<double> &thisToOther, int level);
};

// ---- TrackingReference (Tracking/TrackingReference.cpp) -----------------------------
struct TrackingReference {
  Frame *keyframe = nullptr;
  int frameID = -1;
  std::vector<float> posData[NL];          // 3 / point
  std::vector<float> gradData[NL];         // 2 / point
  std::vector<float> colorAndVarData[NL];  // 2 / point
  std::vector<int> pointPosInXYGrid[NL];   // x + y*w
  int numData[NL] = {};
  void importFrame(Frame *kf);
  void makePointCloud(int level);
  void invalidate();
};

// SCALAR: sequential fp32 sums in emission order (upstream's scalar build).
// SSE4:   4 interleaved fp32 lanes + scalar tail (lane ORDER of upstream's -DENABLE_SSE build).
// EXACT:  identical per-point fp32 values, but every cross-point sum accumulated in fp64 and rounded
//         once -- the order-independent definition of the same algorithm.  Used to separate the
//         oracle's own fp32 summation noise from genuine differences (tests/test_gpu_se3.py).
enum class ReduceMode { SCALAR = 0, SSE4 = 1, EXACT = 2 };

struct Acc {  // one running sum in the selected mode
  float f = 0;
  double d = 0;
  bool exact;
  explicit Acc(bool e = false) : exact(e) {}
  inline void add(float v) { if (exact) d += (double)v; else f += v; }
  inline float get() const { return exact ? (float)d : f; }
};

struct LMTraceEntry {  // one LM evaluation (B3+B4), for per-iteration parity checks
  int level;
  int accepted;  // -1: first evaluation of a level; 0 rejected; 1 accepted
  float error;
  float lambda;
  int bufSize;
};

// ---- SE3Tracker (Tracking/SE3Tracker.cpp) --------------------------------------------
struct SE3Tracker {
  TrackerSettings settings;
  ReduceMode mode = ReduceMode::SCALAR;
  int w0, h0;
  // buffers (SoA, w0*h0)
  std::vector<float> buf_warped_residual, buf_warped_dx, buf_warped_dy, buf_warped_x, buf_warped_y, buf_warped_z, buf_d,
      buf_idepthVar, buf_weight_p;
  int buf_warped_size = 0;
  // outputs
  float pointUsage = 0, lastGoodCount = 0, lastMeanRes = 0, lastBadCount = 0, lastResidual = 0;
  float affineEstimation_a = 1, affineEstimation_b = 0, affineEstimation_a_lastIt = 1, affineEstimation_b_lastIt = 0;
  bool diverged = false, trackingWasGood = false;
  int iterationNumber = 0;
  int numCalcResidualCalls[NL] = {}, numCalcWarpUpdateCalls[NL] = {};
  std::vector<LMTraceEntry> trace;

  SE3Tracker(int w, int h);
  SE3<double> trackFrame(TrackingReference *ref, Frame *frame, const SE3<double> &frameToReference_initialEstimate);
  SE3<double> trackFrameOnPermaref(Frame *reference, TrackingReference *permaRef, Frame *frame,
                                   const SE3<double> &referenceToFrameOrg);
  float checkPermaRefOverlap(Frame *reference, TrackingReference *permaRef, const SE3<double> &referenceToFrameOrg);

  float calcResidualAndBuffers(const float *refPoint, const float *refColVar, const int *idxBuf, int refNum, Frame *frame,
                               const SE3<float> &referenceToFrame, int level);
  float calcWeightsAndResidual(const SE3<float> &referenceToFrame);
  // returns A (6x6), b (6) after finish(); error
  void calculateWarpUpdate(float A[6][6], float b[6], float *error);
};

// ---- Sim3Tracker (Tracking/Sim3Tracker.cpp) ------------------------------------------
struct Sim3ResidualStruct {
  float sumResD = 0, sumResP = 0;
  int numTermsD = 0, numTermsP = 0;
  float meanD = 0, meanP = 0, mean = 0;
};

struct Sim3Tracker {
  TrackerSettings settings;
  ReduceMode mode = ReduceMode::SCALAR;
  int w0, h0;
  std::vector<float> buf_warped_residual, buf_warped_weights, buf_warped_dx, buf_warped_dy, buf_warped_x, buf_warped_y,
      buf_warped_z, buf_d, buf_residual_d, buf_idepthVar, buf_warped_idepthVar, buf_weight_p, buf_weight_d,
      buf_weight_Huber, buf_weight_VarP, buf_weight_VarD;
  int buf_warped_size = 0;
  float pointUsage = 0, lastResidual = 0, lastDepthResidual = 0, lastPhotometricResidual = 0;
  float affineEstimation_a = 1, affineEstimation_b = 0, affineEstimation_a_lastIt = 1, affineEstimation_b_lastIt = 0;
  bool diverged = false;
  float lastSim3Hessian[7][7] = {};
  std::vector<LMTraceEntry> trace;

  Sim3Tracker(int w, int h);
  Sim3<double> trackFrameSim3(TrackingReference *ref, Frame *frame, const Sim3<double> &frameToReference_initialEstimate,
                              int startLevel, int finalLevel);
  void calcSim3Buffers(TrackingReference *ref, Frame *frame, const Sim3<double> &referenceToFrame, int level);
  Sim3ResidualStruct calcSim3WeightsAndResidual(const Sim3<double> &referenceToFrame);
  void calcSim3LGS(float A[7][7], float b[7], int *numConstraints);
};

// ---- DepthMap (DepthEstimation/DepthMap.cpp) -----------------------------------------
struct DepthMapStats {
  int created = 0, updated = 0, killed = 0, skipped = 0;
};

// util/settings.h thresholds that SURVEY.md 8a-K and the published upstream header disagree on (see the
// DECISION note at the top of depthmap.cpp): run-time settings, defaults = published upstream values.
struct DepthSettings {
  int valSumMinForCreate = 30;       // VAL_SUM_MIN_FOR_CREATE
  int valSumMinForKeep = 24;         // VAL_SUM_MIN_FOR_KEEP
  int valSumMinForUnblacklist = 100; // VAL_SUM_MIN_FOR_UNBLACKLIST
  int minBlacklist = -1;             // MIN_BLACKLIST
};

struct DepthMap {
  DepthSettings settings;
  float lastRescaleFactor = 1;
  int width, height;
  float fx, fy, cx, cy, fxi, fyi, cxi, cyi;
  std::vector<Hypothesis> currentDepthMap, otherDepthMap;
  std::vector<int> validityIntegralBuffer;
  Frame *activeKeyFrame = nullptr;
  bool activeKeyFrameIsReactivated = false;
  Frame *oldest_referenceFrame = nullptr, *newest_referenceFrame = nullptr;
  std::vector<Frame *> referenceFrameByID;
  int referenceFrameByID_offset = 0;
  int numThreads = 1;  // IndexThreadReduce workers (MAPPING_THREADS = 4 upstream)

  DepthMap(int w, int h, float fx, float fy, float cx, float cy);
  void initializeFromGTDepth(Frame *new_frame);
  void initializeRandomly(Frame *new_frame);
  void initializeFromMap(Frame *kf, const Hypothesis *map);  // DECISION: inject an explicit map (replaces rand())
  void updateKeyframe(const std::deque<Frame *> &referenceFrames);
  void createKeyFrame(Frame *new_keyframe);
  void finalizeKeyFrame();

  // stages (public so tests can time / compare them individually)
  void observeDepth();
  void observeDepthRow(int yMin, int yMax);
  bool observeDepthCreate(int x, int y, int idx);
  bool observeDepthUpdate(int x, int y, int idx, const float *keyFrameMaxGradBuf);
  bool makeAndCheckEPL(int x, int y, const Frame *ref, float *pepx, float *pepy);
  float doLineStereo(float u, float v, float epxn, float epyn, float min_idepth, float prior_idepth, float max_idepth,
                     const Frame *referenceFrame, const float *referenceFrameImage, float &result_idepth,
                     float &result_var, float &result_eplLength);
  void propagateDepth(Frame *new_keyframe);
  void regularizeDepthMap(bool removeOcclusions, int validityTH);
  void regularizeDepthMapRow(bool removeOcclusions, int validityTH, int yMin, int yMax);
  void regularizeDepthMapFillHoles();
  void regularizeDepthMapFillHolesRow(int yMin, int yMax);
  void buildRegIntegralBuffer();
  void debugPlotDepthMap(uint8_t *rgb) const;
  void parallelRows(int first, int end, int step, const std::function<void(int, int)> &f);
};

// bilinear samplers (util/globalFuncs.h)
inline float getInterpolatedElement(const float *mat, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float *bp = mat + ix + iy * width;
  return dxdy * bp[1 + width] + (dy - dxdy) * bp[width] + (dx - dxdy) * bp[1] + (1 - dx - dy + dxdy) * bp[0];
}
inline void getInterpolatedElement4N(const float *mat4, float x, float y, int width, int n, float *out) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float *bp = mat4 + 4 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int c = 0; c < n; c++)
    out[c] = w11 * bp[4 * (1 + width) + c] + w01 * bp[4 * width + c] + w10 * bp[4 + c] + w00 * bp[c];
}

}  // namespace lsdo
