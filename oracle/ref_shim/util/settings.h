// oracle/ref_shim/util/settings.h -- TEST INFRASTRUCTURE ONLY (oracle/_ref recipe).
// Keyframe.h includes the lsd-slam core's util/settings.h but uses nothing from it.
#pragma once
