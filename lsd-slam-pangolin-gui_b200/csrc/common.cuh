// common.cuh -- shared device/host declarations of liblsd_b200 (sm_100a only, no fallback path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/lsd_b200.h"

#define NL LSD_PYRAMID_LEVELS

namespace lsd {

void set_error(const std::string &msg);

#define LSD_CUDA(call)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      ::lsd::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ + ":" +    \
                       std::to_string(__LINE__));                                                       \
      return LSD_ERR_CUDA;                                                                              \
    }                                                                                                   \
  } while (0)

#define LSD_ARG(cond)                                                                     \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::lsd::set_error(std::string("bad argument: ") + #cond + " @" + __FILE__ + ":" +    \
                       std::to_string(__LINE__));                                         \
      return LSD_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)

// util/settings.h constants used by the kernels
#define LSD_MIN_USE_GRAD 5.0f
#define LSD_CAMERA_PIXEL_NOISE2 16.0f
#define LSD_MAX_DIFF_CONSTANT 1600.0f
#define LSD_MAX_DIFF_GRAD_MULT 0.25f
#define LSD_MIN_GOODPERGOODBAD_PIXEL 0.5f
#define LSD_MIN_GOODPERALL_PIXEL 0.04f
#define LSD_MIN_GOODPERALL_PIXEL_ABSMIN 0.01f
#define LSD_VAR_GT_INIT_INITIAL (0.01f * 0.01f)
#define LSD_SE3TRACKING_MIN_LEVEL 1
#define LSD_SE3TRACKING_MAX_LEVEL 5
#define LSD_QUICK_KF_CHECK_LVL 4

// Per-level pinhole intrinsics (Frame::initialize): fx_l = fx_{l-1}/2, cx_l = (cx_0+.5)/2^l - .5
struct Intrinsics {
  int w[NL], h[NL];
  float fx[NL], fy[NL], cx[NL], cy[NL], fxi[NL], fyi[NL], cxi[NL], cyi[NL];
};

// Byte offsets of every plane inside one frame slab (one cudaMalloc per frame, pooled).
struct FrameLayout {
  size_t img[NL];     // float
  size_t grad[NL];    // float4 (gx, gy, I, 0)
  size_t maxgrad;     // float, level 0
  size_t idepth[NL];  // float
  size_t idvar[NL];   // float
  size_t mask;        // uint8 (w>>1)*(h>>1)
  size_t total;
};

}  // namespace lsd


#ifdef __CUDACC__
// Decoupled look-back over raster-ordered chunks (single-pass ordered compaction; used by k_make_pointcloud and
// k_vbo_extract).  state[c] = flag << 30 | count, flag 1 = the chunk's own count, 2 = inclusive prefix; all zero before
// the launch.  Chunk ids must be handed out by an atomic ticket so that every predecessor is resident.  Called by ALL
// 32 lanes of one warp with the chunk's own `total`; returns the number of items in the chunks before `chunk`.
#define LSD_LB_SHIFT 30
#define LSD_LB_MASK ((1u << LSD_LB_SHIFT) - 1u)
__device__ __forceinline__ unsigned lookback_exclusive(unsigned *state, unsigned chunk, unsigned total, int lane) {
  volatile unsigned *vs = state;
  unsigned excl = 0;
  if (chunk > 0) {
    if (lane == 0) vs[chunk] = (1u << LSD_LB_SHIFT) | total;
    int idx = (int)chunk - 1 - lane;
    while (true) {
      unsigned v = idx >= 0 ? vs[idx] : (2u << LSD_LB_SHIFT);
      while (__any_sync(0xffffffffu, (v >> LSD_LB_SHIFT) == 0u)) v = idx >= 0 ? vs[idx] : (2u << LSD_LB_SHIFT);
      const unsigned inclMask = __ballot_sync(0xffffffffu, (v >> LSD_LB_SHIFT) == 2u);
      const int first = inclMask ? __ffs(inclMask) - 1 : 32;  // nearest predecessor that holds an inclusive prefix
      unsigned c = lane <= first ? (v & LSD_LB_MASK) : 0u;
#pragma unroll
      for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      excl += c;
      if (inclMask) break;
      idx -= 32;
    }
  }
  if (lane == 0) vs[chunk] = (2u << LSD_LB_SHIFT) | (excl + total);
  return excl;
}
#endif

#ifndef LSD_STENCIL_TMA_DEFAULT
#define LSD_STENCIL_TMA_DEFAULT 1  // bit 0: regularizeDepthMap, bit 1: fillHoles (r02l: 0.283 vs 0.285 ms and 0.235 vs 0.202 ms per 64 keyframes)
#endif

struct lsd_frame {
  int id;
  uint8_t *slab;       // device
  unsigned built;      // bit0 tracking planes, bit1 maxgrad0, bit2 grad0, bit3 idepth L0, bit4 idepth L1-4, bit5 mask init
  int numMappable;     // -1 until maxgrad0 built and read
  int *d_numMappable;  // device counter (inside slab tail)
  float initialTrackedResidual;
  double thisToParent_raw[8];
  int trackingParentId;
  float meanIdepth;
  int numPoints;
  bool meanValid;  // meanIdepth / numPoints describe the current level-0 idepth planes
  int numFramesTrackedOnThis, numMappedOnThis, numMappedOnThisTotal;
  bool depthHasBeenUpdatedFlag;
};

enum {
  FB_TRACKING = 1u,
  FB_MAXGRAD0 = 2u,
  FB_GRAD0 = 4u,
  FB_IDEPTH0 = 8u,
  FB_IDEPTH_PYR = 16u,
  FB_MASK = 32u
};

// One tracking-reference point (16 B, one LDG.128): what TrackingReference::makePointCloud emits,
// with pos = invDepth * (fxi*x+cxi, fyi*y+cyi, 1) recomputed on the fly.  invDepth is upstream's own
// intermediate 1.0f / idepth (IEEE division, done once in k_make_pointcloud), so the position is
// bit-identical to the stored posData and the trackers save one division per point per evaluation.
struct __align__(16) RefPoint {
  uint32_t xy;  // x | y << 16
  float invDepth;
  float color;
  float var;
};

struct lsd_ref {
  lsd_frame *keyframe;
  int frameID;
  uint8_t *slab;         // device: per level RefPoint[N_l] + float2 grad[N_l]
  size_t offPts[NL], offGrad[NL];
  int *d_num;            // device int[NL]
  int num[NL];           // host copy (valid when numValid)
  bool numValid;
};
