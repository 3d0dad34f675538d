// stand-in for libvideoio/types/Camera.h
#pragma once
namespace libvideoio {
struct Camera {
  float fx = 0, fy = 0, cx = 0, cy = 0;
};
}  // namespace libvideoio
