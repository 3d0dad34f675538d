// util/SophusUtil.h of the lsd-slam core: the Sophus typedefs every consumer uses
#pragma once
#include "sophus/sim3.hpp"
typedef Sophus::Sim3d Sim3;
