"""lsd_b200 -- host-side mirror (Python/ctypes) of the B200-native LSD-SLAM hot path.

The product is lsd-slam-pangolin-gui_b200/liblsd_b200.so (hand-written sm_100a kernels behind the
C ABI in include/lsd_b200.h).  This package only binds it; importing it does not require a GPU, but
every compute call does (there is no CPU fallback).
"""
from .binding import (BUILD_GRAD0, BUILD_MAXGRAD0, BUILD_TRACKING, FIELD_GRADIENTS, FIELD_IDEPTH, FIELD_IDEPTHVAR,
                      FIELD_IMAGE, FIELD_MASK, FIELD_MAXGRAD, HYP_DTYPE, LIB_PATH, STAGE_FILL_HOLES, STAGE_OBSERVE, STAGE_PROPAGATE,
                      STAGE_REGULARIZE, STAGE_SET_DEPTH, SYMBOLS, POINT_DTYPE, VERTEX_DTYPE, VboParams, SlamStatus, SlamSystem, Undistorter, Context, DepthMap, Frame, LsdError, Ref, load)

__all__ = ["Context", "Frame", "Ref", "DepthMap", "HYP_DTYPE", "STAGE_OBSERVE", "STAGE_FILL_HOLES", "STAGE_REGULARIZE",
           "STAGE_PROPAGATE", "STAGE_SET_DEPTH", "LsdError", "load", "LIB_PATH", "SYMBOLS", "FIELD_IMAGE", "FIELD_GRADIENTS",
           "FIELD_MAXGRAD", "FIELD_IDEPTH", "FIELD_IDEPTHVAR", "FIELD_MASK", "BUILD_TRACKING", "BUILD_MAXGRAD0",
           "BUILD_GRAD0", "POINT_DTYPE", "VERTEX_DTYPE", "VboParams", "SlamSystem", "SlamStatus", "Undistorter"]
