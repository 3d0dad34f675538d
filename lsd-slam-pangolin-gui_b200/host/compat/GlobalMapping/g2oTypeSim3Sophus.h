// GlobalMapping/g2oTypeSim3Sophus.h of the lsd-slam core: g2o vertex / edge types of the pose graph.  The pose graph stays on
// the reference's CPU code (BASELINE.json north_star); the output wrappers include this header but use nothing of it.
#pragma once
