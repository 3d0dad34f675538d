// GlobalMapping/KeyFrameGraph.h of the lsd-slam core, as far as the output wrappers read it
// (/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:150-164): keyframesAll + its shared mutex.  The graph
// itself (edges, optimisation) stays on the reference's CPU code.
#pragma once
#include <boost/thread/shared_mutex.hpp>
#include <memory>
#include <vector>

#include "../DataStructures/Frame.h"
namespace lsd_slam {
class KeyFrame {
 public:
  typedef std::shared_ptr<KeyFrame> SharedPtr;
  explicit KeyFrame(const Frame::SharedPtr &f) : frame_(f) {}
  int id() const { return frame_->id(); }
  const Frame::SharedPtr &frame() const { return frame_; }

 private:
  Frame::SharedPtr frame_;
};
class KeyFrameGraph {
 public:
  std::vector<KeyFrame::SharedPtr> keyframesAll;
  boost::shared_mutex keyframesAllMutex;
};
}  // namespace lsd_slam
