// IOWrapper/OutputIOWrapper.h of the lsd-slam core: the 12 virtuals the reference's wrappers override
// (/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.h:28-63, TextOutputIOWrapper.h:20-52)
#pragma once
#include <Eigen/Core>
#include <memory>
#include <opencv2/core.hpp>
#include <string>
#include <vector>

#include "../DataStructures/Frame.h"
#include "../GlobalMapping/KeyFrameGraph.h"
#include "../util/SophusUtil.h"
#include "g3log/g3log.hpp"
namespace lsd_slam {
class OutputIOWrapper {
 public:
  virtual ~OutputIOWrapper() {}
  virtual void publishPose(const Sophus::Sim3f &pose) = 0;
  virtual void publishKeyframeGraph(const std::shared_ptr<KeyFrameGraph> &graph) = 0;
  virtual void publishPointCloud(const std::shared_ptr<KeyFrameGraph> &graph) = 0;
  virtual void publishPointCloud(const Frame::SharedPtr &kf) = 0;
  virtual void publishKeyframe(const Frame::SharedPtr &kf) = 0;
  virtual void updateDepthImage(unsigned char *data) = 0;
  virtual void publishTrackedFrame(const Frame::SharedPtr &kf) = 0;
  virtual void publishTrajectory(std::vector<Eigen::Matrix<float, 3, 1>> trajectory, std::string identifier) = 0;
  virtual void publishTrajectoryIncrement(Eigen::Matrix<float, 3, 1> pt, std::string identifier) = 0;
  virtual void publishDebugInfo(Eigen::Matrix<float, 20, 1> data) = 0;
  virtual void updateFrameNumber(int) = 0;
  virtual void updateLiveImage(const cv::Mat &img) = 0;
};
}  // namespace lsd_slam
