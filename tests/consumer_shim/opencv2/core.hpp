// stand-in for <opencv2/core.hpp>: cv::Mat as far as OutputIOWrapper::updateLiveImage(const cv::Mat&) needs it
#pragma once
namespace cv {
struct Mat {
  unsigned char *data = nullptr;
  int rows = 0, cols = 0;
};
}  // namespace cv
