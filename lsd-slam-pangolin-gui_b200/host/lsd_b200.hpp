// lsd_b200.hpp -- header-only C++11 host adapter: the lsd-slam core classes the reference links
// (lsd_slam::Frame, TrackingReference, SE3Tracker, Sim3Tracker, DepthMap; SURVEY.md 8b "upper face")
// re-expressed over the C ABI of liblsd_b200.so (include/lsd_b200.h).  Method names, argument meaning and
// result members follow upstream ([UP] = un-vendored lsd-slam core; /root/reference/fips.yml:1-4), so a
// SlamSystem built against these classes drives the B200 kernels instead of the CPU/SSE loops:
//   tools/LSD.cpp:102              new SlamSystem()            -> owns one lsd_b200::Context per worker thread
//   lib/App/InputThread.cpp:71     system->nextImage(...)      -> Frame(id, w, h, K, timestamp, image)
//   PangolinOutputIOWrapper.cpp:56-79   f->image/idepth/idepthVar(level)  -> Frame accessors (lazy D2H)
//   TextOutputIOWrapper.cpp:104-117     kf->pose->thisToParent_raw        -> Frame::thisToParent_raw()
//
// Pose types: plain structs SE3 / Sim3 in Sophus' data() order.  With -DLSD_B200_WITH_SOPHUS the converting
// constructors from / to Sophus::SE3d / Sophus::Sim3d are enabled (needs an lsd-slam checkout's SophusUtil.h).
// Errors: upstream aborts through g3log CHECK; here every failed C call throws lsd_b200::Error.
#pragma once
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lsd_b200.h"

#ifdef LSD_B200_WITH_SOPHUS
#include "util/SophusUtil.h"
#endif
// LSD_B200_LSDSLAM_COMPAT (set by host/compat/DataStructures/Frame.h): Frame additionally carries everything the reference's
// consumers read through the lsd-slam headers -- fx/fy/cx/cy(level), getCamToWorld(), getActiveLock(), pose->thisToParent_raw
// (/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:50-63, TextOutputIOWrapper.cpp:107,114) -- in the types
// those files expect (Sophus::Sim3d, boost::shared_lock).
#ifdef LSD_B200_LSDSLAM_COMPAT
#include <Eigen/Core>
#include <boost/thread/shared_mutex.hpp>
#include <cmath>

#include "sophus/sim3.hpp"
#endif

namespace lsd_b200 {

struct Error : std::runtime_error {
  explicit Error(const std::string &m) : std::runtime_error(m) {}
};
inline void check(int rc) {
  if (rc != LSD_OK) throw Error(std::string("liblsd_b200: ") + lsd_last_error());
}

struct SE3 {  // Sophus::SE3d::data(): unit quaternion (x,y,z,w), translation
  double d[7];
  SE3() : d{0, 0, 0, 1, 0, 0, 0} {}
#ifdef LSD_B200_WITH_SOPHUS
  SE3(const Sophus::SE3d &s) { std::memcpy(d, s.data(), sizeof(d)); }
  operator Sophus::SE3d() const {
    return Sophus::SE3d(Eigen::Quaterniond(d[3], d[0], d[1], d[2]), Eigen::Vector3d(d[4], d[5], d[6]));
  }
#endif
};
struct Sim3 {  // quaternion (x,y,z,w), translation, scale
  double d[8];
  Sim3() : d{0, 0, 0, 1, 0, 0, 0, 1} {}
#ifdef LSD_B200_WITH_SOPHUS
  Sim3(const Sophus::Sim3d &s) {
    const Eigen::Quaterniond q(s.rotationMatrix());
    d[0] = q.x(); d[1] = q.y(); d[2] = q.z(); d[3] = q.w();
    d[4] = s.translation()[0]; d[5] = s.translation()[1]; d[6] = s.translation()[2];
    d[7] = s.scale();
  }
  operator Sophus::Sim3d() const {
    Sophus::Sim3d r(Sophus::RxSO3d(d[7], Eigen::Quaterniond(d[3], d[0], d[1], d[2]).toRotationMatrix()),
                    Eigen::Vector3d(d[4], d[5], d[6]));
    return r;
  }
#endif
};

// One device + stream + scratch.  Upstream runs tracking, mapping and constraint search on separate threads:
// give each its own Context (frames / references may be shared read-only between contexts of one device).
class Context {
 public:
  Context(int width, int height, float fx, float fy, float cx, float cy, int device = 0, void *stream = nullptr) : w_(width), h_(height) {
    const float K[4] = {fx, fy, cx, cy};
    check(lsd_ctx_create(device, width, height, K, stream, &c_));
    // [UP] Frame::initialize: fx_l = fx_{l-1} * 0.5, cx_l = (cx_0 + 0.5) / 2^l - 0.5 (the values the device uses, csrc/api.cu)
    for (int l = 0; l < LSD_PYRAMID_LEVELS; l++) {
      fx_[l] = l ? (float)(fx_[l - 1] * 0.5) : fx;
      fy_[l] = l ? (float)(fy_[l - 1] * 0.5) : fy;
      cx_[l] = l ? (float)((cx + 0.5) / (1 << l) - 0.5) : cx;
      cy_[l] = l ? (float)((cy + 0.5) / (1 << l) - 0.5) : cy;
    }
    if (!current()) current() = this;
  }
  ~Context() {
    if (current() == this) current() = nullptr;
    lsd_ctx_destroy(c_);
  }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  lsd_ctx *c() const { return c_; }
  int width() const { return w_; }
  int height() const { return h_; }
  float fx(int level = 0) const { return fx_[level]; }
  float fy(int level = 0) const { return fy_[level]; }
  float cx(int level = 0) const { return cx_[level]; }
  float cy(int level = 0) const { return cy_[level]; }
  // The context of the calling thread (one context per worker thread, include/lsd_b200.h): what the upstream-shaped Frame
  // constructor, which carries no context argument, builds its frame on.  The first context a thread creates becomes current.
  static Context *&current() {
    static thread_local Context *cur = nullptr;
    return cur;
  }
  void makeCurrent() { current() = this; }
  // The context of SlamSystem's tracking thread (one frame per tracking call): per-level record sizes for the live tracker
  // kernel -- one thread-block cluster per (reference, frame) pair (include/lsd_b200.h: lsd_ctx_set_live_tracking)
  void setLiveTracking(bool on = true) { check(lsd_ctx_set_live_tracking(c_, on ? 1 : 0)); }

 private:
  lsd_ctx *c_ = nullptr;
  int w_, h_;
  float fx_[LSD_PYRAMID_LEVELS], fy_[LSD_PYRAMID_LEVELS], cx_[LSD_PYRAMID_LEVELS], cy_[LSD_PYRAMID_LEVELS];
};

#ifdef LSD_B200_LSDSLAM_COMPAT
inline Sophus::Sim3d toSophus(const double d[8]) {
  return Sophus::Sim3d(Sophus::RxSO3d(d[7], Eigen::Quaterniond(d[3], d[0], d[1], d[2])), Eigen::Vector3d(d[4], d[5], d[6]));
}
inline void fromSophus(const Sophus::Sim3d &s, double d[8]) {
  const Eigen::Quaterniond q = s.quaternion();  // RxSO3: squared norm = scale
  const double n = std::sqrt(q.squaredNorm());
  d[0] = q.x() / n; d[1] = q.y() / n; d[2] = q.z() / n; d[3] = q.w() / n;
  d[4] = s.translation()[0]; d[5] = s.translation()[1]; d[6] = s.translation()[2];
  d[7] = s.scale();
}
// [UP] lsd_slam::FramePoseStruct: frame -> tracking parent chain; the pose graph (camToWorld_new, graphVertex) stays upstream
struct FramePoseStruct {
  FramePoseStruct *trackingParent = nullptr;
  Sophus::Sim3d thisToParent_raw;
  int frameID = -1;
  bool isRegisteredToGraph = false;
  Sophus::Sim3d camToWorld;  // absolute pose of a registered keyframe (set by the owner of the graph)
  Sophus::Sim3d getCamToWorld() const {
    if (isRegisteredToGraph || !trackingParent) return camToWorld;
    return trackingParent->getCamToWorld() * thisToParent_raw;
  }
};
#endif

// [UP] lsd_slam::Frame.  Pyramids live on the device; accessors copy a level to a host cache on first use.
class Frame {
 public:
  typedef std::shared_ptr<Frame> SharedPtr;
  // [UP] Frame(int id, int width, int height, const Eigen::Matrix3f& K, double timestamp, const unsigned char* image)
  Frame(Context &ctx, int id, double timestamp, const unsigned char *image, size_t pitch = 0, bool keyframeCandidate = true)
      : ctx_(ctx), id_(id), timestamp_(timestamp) {
    check(lsd_frame_create(ctx.c(), id, image, pitch ? pitch : (size_t)ctx.width(), keyframeCandidate ? LSD_BUILD_MAXGRAD0 : LSD_BUILD_TRACKING, &f_));
  }
#ifdef LSD_B200_LSDSLAM_COMPAT
  // [UP] Frame(int id, int width, int height, const Eigen::Matrix3f& K, double timestamp, const unsigned char* image): the frame is
  // built on the calling thread's current context, which must have been created for the same image size and camera
  Frame(int id, int width, int height, const Eigen::Matrix3f &K, double timestamp, const unsigned char *image)
      : ctx_(requireContext(width, height, K(0, 0), K(1, 1), K(0, 2), K(1, 2))), id_(id), timestamp_(timestamp) {
    check(lsd_frame_create(ctx_.c(), id, image, (size_t)width, LSD_BUILD_MAXGRAD0, &f_));
    pose_.frameID = id;
  }
#endif
  ~Frame() { lsd_frame_release(ctx_.c(), f_); }
  Frame(const Frame &) = delete;
  Frame &operator=(const Frame &) = delete;

  int id() const { return id_; }
  double timestamp() const { return timestamp_; }
  int width(int level = 0) const { return ctx_.width() >> level; }
  int height(int level = 0) const { return ctx_.height() >> level; }
  const float *image(int level = 0) { return plane(LSD_FIELD_IMAGE, level, 1); }
  const float *gradients(int level = 0) { return plane(LSD_FIELD_GRADIENTS, level, 4); }  // Eigen::Vector4f per pixel
  const float *maxGradients(int level = 0) { return plane(LSD_FIELD_MAXGRAD, level, 1); }
  const float *idepth(int level = 0) { return plane(LSD_FIELD_IDEPTH, level, 1); }
  const float *idepthVar(int level = 0) { return plane(LSD_FIELD_IDEPTHVAR, level, 1); }
  bool hasIDepthBeenSet() {
    float m;
    int n;
    return lsd_frame_mean_idepth(ctx_.c(), f_, &m, &n) == LSD_OK;
  }
  void setDepthFromGroundTruth(const float *depth, float cov_scale = 1.0f) {
    check(lsd_frame_set_depth_from_gt(ctx_.c(), f_, depth, cov_scale));
    invalidate();
  }
  Sim3 thisToParent_raw() const {  // [UP] pose->thisToParent_raw
    Sim3 s;
    int pid;
    float r;
    check(lsd_frame_get_tracking_meta(ctx_.c(), f_, &pid, s.d, &r));
    return s;
  }
  float initialTrackedResidual() const {
    Sim3 s;
    int pid;
    float r;
    check(lsd_frame_get_tracking_meta(ctx_.c(), f_, &pid, s.d, &r));
    return r;
  }
  void invalidate() { cache_.clear(); }  // device planes changed (setDepth, tracking mask)
  lsd_frame *handle() const { return f_; }
  Context &context() const { return ctx_; }
  // per-level intrinsics, as [UP] Frame::fx(level) ... (read at PangolinOutputIOWrapper.cpp:60-63)
  float fx(int level = 0) const { return ctx_.fx(level); }
  float fy(int level = 0) const { return ctx_.fy(level); }
  float cx(int level = 0) const { return ctx_.cx(level); }
  float cy(int level = 0) const { return ctx_.cy(level); }
#ifdef LSD_B200_LSDSLAM_COMPAT
  FramePoseStruct *pose = &pose_;  // [UP] Frame::pose (TextOutputIOWrapper.cpp:114 reads pose->thisToParent_raw)
  Sophus::Sim3d getCamToWorld() const { return pose_.getCamToWorld(); }  // [UP] Frame::getCamToWorld()
  // [UP] Frame::getActiveLock(): readers hold it while they copy planes; the mapping side takes the unique lock around setDepth
  boost::shared_lock<boost::shared_mutex> getActiveLock() { return boost::shared_lock<boost::shared_mutex>(activeMutex); }
  boost::shared_mutex activeMutex;
  // pulls thisToParent_raw (written on the device-side handle by SE3Tracker::trackFrame / DepthMap::createKeyFrame) into pose
  void syncPoseFromDevice(FramePoseStruct *parent) {
    const Sim3 s = thisToParent_raw();
    pose_.thisToParent_raw = toSophus(s.d);
    if (parent) pose_.trackingParent = parent;
  }
#endif

 private:
#ifdef LSD_B200_LSDSLAM_COMPAT
  static Context &requireContext(int width, int height, float fx, float fy, float cx, float cy) {
    Context *c = Context::current();
    if (!c) throw Error("lsd_b200::Frame: no current Context on this thread (create one, or call Context::makeCurrent())");
    if (c->width() != width || c->height() != height || c->fx() != fx || c->fy() != fy || c->cx() != cx || c->cy() != cy)
      throw Error("lsd_b200::Frame: image size / camera matrix differ from the current Context's");
    return *c;
  }
  FramePoseStruct pose_;
#endif
  const float *plane(int field, int level, int comps) {
    const int key = field * 8 + level;
    for (auto &e : cache_)
      if (e.first == key) return e.second.data();
    cache_.emplace_back(key, std::vector<float>((size_t)width(level) * height(level) * comps));
    check(lsd_frame_read(ctx_.c(), f_, field, level, cache_.back().second.data()));
    return cache_.back().second.data();
  }
  Context &ctx_;
  lsd_frame *f_ = nullptr;
  int id_;
  double timestamp_;
  std::vector<std::pair<int, std::vector<float>>> cache_;
};

// [UP] lsd_slam::TrackingReference
class TrackingReference {
 public:
  explicit TrackingReference(Context &ctx) : ctx_(ctx) {}
  ~TrackingReference() { invalidate(); }
  void importFrame(Frame *source) {  // + makePointCloud(level) for levels 1..4, on device
    invalidate();
    keyframe = source;
    frameID = source->id();
    check(lsd_ref_create(ctx_.c(), source->handle(), &r_));
  }
  void invalidate() {
    if (r_) lsd_ref_release(ctx_.c(), r_);
    r_ = nullptr;
    keyframe = nullptr;
  }
  int numData(int level) const {
    int n = 0;
    check(lsd_ref_num_data(ctx_.c(), r_, level, &n));
    return n;
  }
  lsd_ref *handle() const { return r_; }
  Frame *keyframe = nullptr;
  int frameID = -1;

 private:
  Context &ctx_;
  lsd_ref *r_ = nullptr;
};

// [UP] lsd_slam::SE3Tracker: public members keep their upstream names.
class SE3Tracker {
 public:
  explicit SE3Tracker(Context &ctx) : ctx_(ctx) { lsd_default_tracker_settings(&settings); }
  // SE3 trackFrame(TrackingReference* reference, Frame* frame, const SE3& frameToReference_initialEstimate)
  SE3 trackFrame(TrackingReference *reference, Frame *frame, const SE3 &frameToReference_initialEstimate) {
    check(lsd_ctx_set_se3_settings(ctx_.c(), &settings));
    lsd_se3_result r;
    check(lsd_se3_track(ctx_.c(), reference->handle(), frame->handle(), frameToReference_initialEstimate.d, &r, nullptr));
    diverged = r.diverged != 0;
    trackingWasGood = r.trackingWasGood != 0;
    lastResidual = r.lastResidual;
    lastMeanRes = r.lastMeanRes;
    pointUsage = r.pointUsage;
    lastGoodCount = r.lastGoodCount;
    lastBadCount = r.lastBadCount;
    affineEstimation_a = r.affine_a;
    affineEstimation_b = r.affine_b;
    frame->invalidate();
#ifdef LSD_B200_LSDSLAM_COMPAT
    if (!diverged) frame->syncPoseFromDevice(reference->keyframe ? reference->keyframe->pose : nullptr);
#endif
    SE3 out;
    std::memcpy(out.d, r.frameToRef, sizeof(out.d));
    return out;
  }
  lsd_tracker_settings settings;  // [UP] DenseDepthTrackerSettings
  float pointUsage = 0, lastGoodCount = 0, lastMeanRes = 0, lastBadCount = 0, lastResidual = 0;
  float affineEstimation_a = 1, affineEstimation_b = 0;
  bool diverged = false, trackingWasGood = false;

 private:
  Context &ctx_;
};

// [UP] lsd_slam::Sim3Tracker
class Sim3Tracker {
 public:
  explicit Sim3Tracker(Context &ctx) : ctx_(ctx) { lsd_default_tracker_settings(&settings); }
  // Sim3 trackFrameSim3(TrackingReference* reference, Frame* frame, const Sim3& frameToReference_initialEstimate, int startLevel, int finalLevel)
  Sim3 trackFrameSim3(TrackingReference *reference, Frame *frame, const Sim3 &frameToReference_initialEstimate, int startLevel,
                      int finalLevel) {
    check(lsd_ctx_set_sim3_settings(ctx_.c(), &settings));
    lsd_sim3_result r;
    check(lsd_sim3_track(ctx_.c(), reference->handle(), frame->handle(), frameToReference_initialEstimate.d, startLevel, finalLevel, &r, nullptr));
    diverged = r.diverged != 0;
    lastResidual = r.lastResidual;
    lastDepthResidual = r.lastDepthResidual;
    lastPhotometricResidual = r.lastPhotometricResidual;
    pointUsage = r.pointUsage;
    affineEstimation_a = r.affine_a;
    affineEstimation_b = r.affine_b;
    std::memcpy(lastSim3Hessian, r.lastSim3Hessian, sizeof(lastSim3Hessian));
    Sim3 out;
    std::memcpy(out.d, r.frameToRef, sizeof(out.d));
    return out;
  }
  lsd_tracker_settings settings;
  float lastSim3Hessian[49] = {0};  // Matrix7x7, row-major
  float pointUsage = 0, lastResidual = 0, lastDepthResidual = 0, lastPhotometricResidual = 0;
  float affineEstimation_a = 1, affineEstimation_b = 0;
  bool diverged = false;

 private:
  Context &ctx_;
};

// [UP] lsd_slam::DepthMap
class DepthMap {
 public:
  explicit DepthMap(Context &ctx) : ctx_(ctx) { check(lsd_depthmap_create(ctx.c(), &d_)); }
  ~DepthMap() { lsd_depthmap_destroy(ctx_.c(), d_); }
  DepthMap(const DepthMap &) = delete;
  DepthMap &operator=(const DepthMap &) = delete;
  void initializeFromGTDepth(Frame *new_frame) { check(lsd_depth_initialize_from_gt(ctx_.c(), d_, new_frame->handle())); new_frame->invalidate(); }
  void initializeRandomly(Frame *new_frame) { check(lsd_depth_initialize_randomly(ctx_.c(), d_, new_frame->handle())); new_frame->invalidate(); }
  // void updateKeyframe(std::deque< std::shared_ptr<Frame> > referenceFrames)
  void updateKeyframe(std::deque<std::shared_ptr<Frame>> referenceFrames) {
    std::vector<lsd_frame *> h;
    for (auto &f : referenceFrames) h.push_back(f->handle());
    check(lsd_depth_update_keyframe(ctx_.c(), d_, (int)h.size(), h.data(), nullptr));
  }
  // void createKeyFrame(Frame* new_keyframe)
  void createKeyFrame(Frame *new_keyframe) {
    float rescale = 1;
    check(lsd_depth_create_keyframe(ctx_.c(), d_, new_keyframe->handle(), &rescale));
    new_keyframe->invalidate();
#ifdef LSD_B200_LSDSLAM_COMPAT
    new_keyframe->syncPoseFromDevice(nullptr);  // thisToParent_raw now carries the mean-idepth rescale factor
#endif
  }
  void finalizeKeyFrame() { check(lsd_depth_finalize_keyframe(ctx_.c(), d_)); }
  // int debugPlotDepthMap(): fills the RGB image handed to OutputIOWrapper::updateDepthImage (lib/GUI.cpp:104-108)
  void debugPlotDepthMap(unsigned char *rgb) { check(lsd_depth_debug_rgb(ctx_.c(), d_, rgb)); }
  void readHypotheses(lsd_hypothesis *dst) { check(lsd_depth_read(ctx_.c(), d_, dst)); }
  lsd_depthmap *handle() const { return d_; }

 private:
  Context &ctx_;
  lsd_depthmap *d_ = nullptr;
};

// libvideoio::Undistorter as used at tools/LSD.cpp:88 / lib/App/InputThread.cpp:62 (OpenCV fixed-point remap maps)
class Undistorter {
 public:
  // maps as an OpenCV undistorter holds them: map1 CV_16SC2 (x, y), map2 CV_16UC1 (sub-pixel table index)
  Undistorter(Context &ctx, int inWidth, int inHeight, const int16_t *map1, const uint16_t *map2) : ctx_(ctx) {
    check(lsd_undistorter_create_from_maps(ctx.c(), inWidth, inHeight, map1, map2, &u_));
  }
  // K = {fx, fy, cx, cy} of the distorted camera, dist = {k1, k2, p1, p2, k3}, Kout = getCamera() of the output
  Undistorter(Context &ctx, int inWidth, int inHeight, const double K[4], const double dist[5], const double Kout[4]) : ctx_(ctx) {
    check(lsd_undistorter_create_opencv(ctx.c(), inWidth, inHeight, K, dist, Kout, &u_));
  }
  ~Undistorter() { lsd_undistorter_destroy(ctx_.c(), u_); }
  Undistorter(const Undistorter &) = delete;
  Undistorter &operator=(const Undistorter &) = delete;
  // void undistort(const cv::Mat &image, cv::OutputArray result): 8-bit grey in / out
  void undistort(const unsigned char *image, size_t pitch, unsigned char *result) { check(lsd_undistort(ctx_.c(), u_, image, pitch, result)); }
  lsd_undistorter *handle() const { return u_; }

 private:
  Context &ctx_;
  lsd_undistorter *u_ = nullptr;
};

// lib/Pangolin_IOWrapper/Keyframe.h: the two loops that touch every pixel of a published keyframe
struct KeyframePublisher {
  // PangolinOutputIOWrapper::publishKeyframe (PangolinOutputIOWrapper.cpp:69-89): fills Keyframe::pointData
  static void pack(Context &ctx, Frame &f, int publishLvl, unsigned char *pointData) {
    check(lsd_frame_publish_keyframe(ctx.c(), f.handle(), publishLvl, reinterpret_cast<lsd_input_point_dense *>(pointData)));
  }
  // Keyframe::computeVbo (Keyframe.h:66-158): fills the MyVertex buffer handed to glBufferData; returns `points`
  static int computeVbo(Context &ctx, Frame &f, int publishLvl, float camToWorldScale, lsd_vertex *vertices) {
    int points = 0;
    check(lsd_keyframe_compute_vbo(ctx.c(), f.handle(), publishLvl, camToWorldScale, nullptr, vertices, &points));
    return points;
  }
};

// [UP] lsd_slam::SlamSystem, lock-step (Conf().runRealTime == false): tools/LSD.cpp:102, lib/App/InputThread.cpp:71
class SlamSystem {
 public:
  // a lock-step system tracks one frame per call: its context gets the live-tracker configuration
  explicit SlamSystem(Context &ctx) : ctx_(ctx) {
    ctx.setLiveTracking(true);
    check(lsd_slam_create(ctx.c(), &s_));
  }
  // 0: every stage of nextImage synchronises on its own (default: stages queued back to back, updateKeyframe finished by the next call)
  void setPipelined(bool on) { check(lsd_slam_set_pipelined(s_, on ? 1 : 0)); }
  ~SlamSystem() { lsd_slam_destroy(s_); }
  SlamSystem(const SlamSystem &) = delete;
  SlamSystem &operator=(const SlamSystem &) = delete;
  void setUndistorter(Undistorter *u) { check(lsd_slam_set_undistorter(s_, u ? u->handle() : nullptr)); }
  void gtDepthInit(int id, const unsigned char *image, size_t pitch, const float *depth) { check(lsd_slam_gt_depth_init(s_, id, image, pitch, depth, &last)); }
  void randomInit(int id, const unsigned char *image, size_t pitch) { check(lsd_slam_random_init(s_, id, image, pitch, &last)); }
  // void nextImage(unsigned int id, const cv::Mat &img, const Camera &cam): returns after tracking AND mapping
  void nextImage(int id, const unsigned char *image, size_t pitch) { check(lsd_slam_next_image(s_, id, image, pitch, &last)); }
  // SlamSystem* fullReset(): a fresh system on the same context (lib/App/InputThread.cpp:86)
  SlamSystem *fullReset() { return new SlamSystem(ctx_); }
  std::string poseLine() const {  // TextOutputIOWrapper::publishTrackedFrame's line
    char buf[256];
    check(lsd_slam_pose_line(&last, buf, sizeof(buf)));
    return buf;
  }
  lsd_slam_status last = lsd_slam_status();

 private:
  Context &ctx_;
  lsd_slam_system *s_ = nullptr;
};

}  // namespace lsd_b200
