// util/Configuration.h of the lsd-slam core: the Conf() singleton the reference application configures
// (/root/reference/tools/LSD.cpp:97-100, lib/App/InputThread.cpp:32,47,94)
#pragma once
#include "libvideoio/types/ImageSize.h"
namespace lsd_slam {
struct Configuration {
  libvideoio::ImageSize slamImageSize;
  bool runRealTime = true;
  bool stopOnFailedRead = true;
  void setSlamImageSize(const libvideoio::ImageSize &sz) { slamImageSize = sz; }
};
inline Configuration &Conf() {
  static Configuration c;
  return c;
}
}  // namespace lsd_slam
