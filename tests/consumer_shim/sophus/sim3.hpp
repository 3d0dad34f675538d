// stand-in for <sophus/sim3.hpp> (Sophus of the lsd-slam era: Sim3Group / RxSO3Group): the members the reference's output
// wrappers, lib/Pangolin_IOWrapper/Keyframe.h and the B200 adapter touch.  data() layout as in Sophus: the RxSO3 quaternion
// (x, y, z, w; squared norm = scale) followed by the translation -- 7 scalars, what publishKeyframeGraph memcpy's.
#pragma once
#include <cmath>

#include "../Eigen/Core"
namespace Sophus {
typedef Eigen::Matrix4f Matrix4f;
template <typename S> class RxSO3Group {
 public:
  RxSO3Group() {}
  RxSO3Group(S scale, const Eigen::Quaternion<S> &unit) {
    const S r = std::sqrt(scale);
    q_ = Eigen::Quaternion<S>(unit.w() * r, unit.x() * r, unit.y() * r, unit.z() * r);
  }
  explicit RxSO3Group(const Eigen::Quaternion<S> &scaled) : q_(scaled) {}
  const Eigen::Quaternion<S> &quaternion() const { return q_; }
  S scale() const { return q_.squaredNorm(); }

 private:
  Eigen::Quaternion<S> q_;
};
template <typename S> class Sim3Group {
 public:
  Sim3Group() { for (int i = 0; i < 7; i++) d_[i] = S(i == 3 ? 1 : 0); }
  Sim3Group(const RxSO3Group<S> &r, const Eigen::Matrix<S, 3, 1> &t) {
    d_[0] = r.quaternion().x(); d_[1] = r.quaternion().y(); d_[2] = r.quaternion().z(); d_[3] = r.quaternion().w();
    d_[4] = t[0]; d_[5] = t[1]; d_[6] = t[2];
  }
  Eigen::Quaternion<S> quaternion() const { return Eigen::Quaternion<S>(d_[3], d_[0], d_[1], d_[2]); }
  RxSO3Group<S> rxso3() const { return RxSO3Group<S>(quaternion()); }
  Eigen::Matrix<S, 3, 1> translation() const { return Eigen::Matrix<S, 3, 1>(d_[4], d_[5], d_[6]); }
  S scale() const { return quaternion().squaredNorm(); }
  S *data() { return d_; }
  const S *data() const { return d_; }
  template <typename T> Sim3Group<T> cast() const {
    Sim3Group<T> r;
    for (int i = 0; i < 7; i++) r.data()[i] = (T)d_[i];
    return r;
  }
  Sim3Group operator*(const Sim3Group &b) const {
    const Eigen::Matrix<S, 3, 1> rt = quaternion().rotateScale(b.translation());
    return Sim3Group(RxSO3Group<S>(quaternion() * b.quaternion()), Eigen::Matrix<S, 3, 1>(d_[4] + rt[0], d_[5] + rt[1], d_[6] + rt[2]));
  }
  Matrix4f matrix() const { return Matrix4f(); }  // only the GL drawing code of Keyframe.h reads it

 private:
  S d_[7];
};
typedef Sim3Group<float> Sim3f;
typedef Sim3Group<double> Sim3d;
typedef RxSO3Group<float> RxSO3f;
typedef RxSO3Group<double> RxSO3d;
}  // namespace Sophus
