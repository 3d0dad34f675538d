// DataStructures/Frame.h of the lsd-slam core -> the device-resident frame of liblsd_b200
#pragma once
#define LSD_B200_LSDSLAM_COMPAT 1
#include "../../lsd_b200.hpp"
namespace lsd_slam {
using lsd_b200::Frame;
using lsd_b200::FramePoseStruct;
}  // namespace lsd_slam
