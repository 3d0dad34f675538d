// depth.cuh -- device-side layout of [UP] lsd_slam::DepthMap, internal.
//
// Upstream keeps two 32-byte AoS maps (currentDepthMap / otherDepthMap) and memcpy's one onto the
// other before every stencil pass.  Here a hypothesis is split into 4-byte SoA planes (24 B/px):
//   meta  uint32  bit 0 isValid | bits 1..15 validity_counter | bits 16..31 blacklisted (int16)
//   next  float   nextStereoFrameMinID
//   idepth, var, ids (idepth_smoothed), vars (idepth_var_smoothed)
// meta / idepth / var exist twice: stencil kernels read one copy and write the other (the snapshot
// semantics of upstream's memcpy, without the copy).
#pragma once
#include "ctx.cuh"

namespace lsd {

__host__ __device__ __forceinline__ uint32_t dm_pack(bool valid, int validity, int blacklisted) {
  validity = validity < 0 ? 0 : (validity > 0x7fff ? 0x7fff : validity);
  blacklisted = blacklisted < -32768 ? -32768 : (blacklisted > 32767 ? 32767 : blacklisted);
  return (valid ? 1u : 0u) | ((uint32_t)validity << 1) | ((uint32_t)(uint16_t)(int16_t)blacklisted << 16);
}
__host__ __device__ __forceinline__ bool dm_valid(uint32_t m) { return (m & 1u) != 0; }
__host__ __device__ __forceinline__ int dm_validity(uint32_t m) { return (int)((m >> 1) & 0x7fffu); }
__host__ __device__ __forceinline__ int dm_black(uint32_t m) { return (int)(int16_t)(uint16_t)(m >> 16); }

// What Frame::prepareForStereoWith caches for one reference frame (A7), plus what observeDepth reads of it.
struct StereoRef {
  const float *img;     // reference frame image, level 0
  const uint8_t *mask;  // refPixelWasGood of the frame if it was tracked on the active keyframe, else null
  int id;
  float initialTrackedResidual;
  float KR[9];   // K_otherToThis_R
  float Kt[3];   // K_otherToThis_t
  float t_o2t[3];  // otherToThis_t
  float t_t2o[3];  // thisToOther_t
  float row0[3], row1[3], row2[3];  // otherToThis_R_row0..2
};

struct DepthDesc {
  // hypothesis planes
  uint32_t *meta, *metaOut;
  float *idepth, *var, *idepthOut, *varOut;
  float *next, *ids, *vars;
  // active keyframe
  const float *kfImg;
  const float4 *kfGrad;
  const float *kfMaxGrad;
  // observeDepth
  const StereoRef *refs;
  const int *refById;  // frame id - refByIdOffset -> index into refs
  int nRefs, refByIdSize, refByIdOffset, reactivated;
  float numTrackedOverMapped;  // numFramesTrackedOnThis / (float)(numMappedOnThis + 5)
  // propagateDepth
  const float *newImg, *newMaxGrad;
  const uint8_t *newMask;  // trackingWasGood or null
  float R[9], t[3];        // oldToNew
  unsigned *cnt;                 // per TARGET pixel: number of sources that arrived
  unsigned *ovfHead, *ovfNext;   // overflow list of a target's fifth and later sources (head per target, link per source)
  float4 *rec;                   // per SOURCE pixel that went to an overflow list: (new_idepth, new_var, validity, source index)
  float4 *tgt[4];                // per TARGET pixel: the records of its rank-0 .. rank-3 arrivals
  // TMA descriptors (CUtensorMap, in device memory) of the CURRENT meta / idepth / var planes for the two stencil tile shapes:
  // [0..2] the 36x36 tile of regularizeDepthMap, [3..5] the 36x12 tile of regularizeDepthMapFillHoles
  const void *tmap[6];
  // setDepth target
  float *frIdepth, *frVar;
  double *sums;  // [4]: sum ids (valid), n valid, sum ids (valid && ids >= -0.05), n
  float rescale; // createKeyFrame: applied when != 0
  int validityTH;
};

}  // namespace lsd

// Host object behind the opaque handle.
struct lsd_depthmap {
  uint8_t *slab;  // device
  uint32_t *meta[2];
  float *idepth[2], *var[2];
  float *next, *ids, *vars;
  int mi, di;  // current copies
  unsigned *cnt, *ovfHead, *ovfNext, *cursor;
  float4 *rec;
  float4 *tgt[4];
  double *sums;
  lsd_frame *activeKeyFrame;
  bool reactivated;
  lsd_depth_settings settings;
  float lastRescale;
  // per-call device tables (refs, refById, desc)
  uint8_t *d_tab, *h_tab;  // h_tab pinned
  size_t tabBytes;
  uint8_t *d_tmaps;  // 12 CUtensorMap (128 B each): [tile shape][plane: meta, idepth, var][copy]
};
