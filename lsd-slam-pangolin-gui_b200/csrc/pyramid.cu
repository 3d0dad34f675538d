// pyramid.cu -- Frame pyramids on device, batched over frames.
//
// Replaces [UP] Frame::Frame(u8 image) + Frame::buildImage / buildGradients / buildMaxGradients /
// buildIDepthAndIDepthVar (lsd-slam core DataStructures/Frame.cpp; the reference consumes the
// results at lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:56-79).  All results are
// bit-exact against the oracle: box means of u8 data are exact in fp32, gradients are exact,
// |g| uses IEEE sqrt and the file is compiled with -fmad=false so dx*dx+dy*dy is not contracted.
#include "ctx.cuh"

namespace lsd {

#define TILE_W 64
#define TILE_H 16

// ---------------------------------------------------------------------------------------
// k_ingest: u8 level 0 -> float image levels 0..4 in one pass (one 64x16 tile per CTA).
// Each thread converts 4 adjacent pixels (uchar4 load, float4 store); levels 1..4 are
// reduced hierarchically in shared memory, so the u8 source is read exactly once.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ingest(const uint8_t *__restrict__ src, size_t srcPitch, size_t srcFrameStride,
                                                uint8_t *const *__restrict__ slabs, FrameLayout lay, int W, int H) {
  __shared__ float s0[TILE_H][TILE_W];
  __shared__ float s1[TILE_H / 2][TILE_W / 2];
  __shared__ float s2[TILE_H / 4][TILE_W / 4];
  __shared__ float s3[TILE_H / 8][TILE_W / 8];
  const int f = blockIdx.z;
  const uint8_t *img = src + (size_t)f * srcFrameStride;
  uint8_t *slab = slabs[f];
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const int t = threadIdx.x;
  {
    const int tx = (t & 15) * 4, ty = t >> 4;  // 16 threads x 4 px per row, 16 rows
    const int x = x0 + tx, y = y0 + ty;
    float4 v = make_float4(0, 0, 0, 0);
    if (x < W && y < H) {
      const uchar4 p = *reinterpret_cast<const uchar4 *>(img + (size_t)y * srcPitch + x);
      v = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
      *reinterpret_cast<float4 *>(reinterpret_cast<float *>(slab + lay.img[0]) + (size_t)y * W + x) = v;
    }
    s0[ty][tx] = v.x; s0[ty][tx + 1] = v.y; s0[ty][tx + 2] = v.z; s0[ty][tx + 3] = v.w;
  }
  __syncthreads();
  {  // level 1: 32 x 8 = 256 px
    const int tx = t & 31, ty = t >> 5;
    const float v = (s0[2 * ty][2 * tx] + s0[2 * ty][2 * tx + 1] + s0[2 * ty + 1][2 * tx] + s0[2 * ty + 1][2 * tx + 1]) * 0.25f;
    s1[ty][tx] = v;
    const int x = (x0 >> 1) + tx, y = (y0 >> 1) + ty;
    if (x < (W >> 1) && y < (H >> 1)) reinterpret_cast<float *>(slab + lay.img[1])[(size_t)y * (W >> 1) + x] = v;
  }
  __syncthreads();
  if (t < 64) {  // level 2: 16 x 4
    const int tx = t & 15, ty = t >> 4;
    const float v = (s1[2 * ty][2 * tx] + s1[2 * ty][2 * tx + 1] + s1[2 * ty + 1][2 * tx] + s1[2 * ty + 1][2 * tx + 1]) * 0.25f;
    s2[ty][tx] = v;
    const int x = (x0 >> 2) + tx, y = (y0 >> 2) + ty;
    if (x < (W >> 2) && y < (H >> 2)) reinterpret_cast<float *>(slab + lay.img[2])[(size_t)y * (W >> 2) + x] = v;
  }
  __syncthreads();
  if (t < 16) {  // level 3: 8 x 2
    const int tx = t & 7, ty = t >> 3;
    const float v = (s2[2 * ty][2 * tx] + s2[2 * ty][2 * tx + 1] + s2[2 * ty + 1][2 * tx] + s2[2 * ty + 1][2 * tx + 1]) * 0.25f;
    s3[ty][tx] = v;
    const int x = (x0 >> 3) + tx, y = (y0 >> 3) + ty;
    if (x < (W >> 3) && y < (H >> 3)) reinterpret_cast<float *>(slab + lay.img[3])[(size_t)y * (W >> 3) + x] = v;
  }
  __syncthreads();
  if (t < 4) {  // level 4: 4 x 1
    const int tx = t;
    const float v = (s3[0][2 * tx] + s3[0][2 * tx + 1] + s3[1][2 * tx] + s3[1][2 * tx + 1]) * 0.25f;
    const int x = (x0 >> 4) + tx, y = (y0 >> 4);
    if (x < (W >> 4) && y < (H >> 4)) reinterpret_cast<float *>(slab + lay.img[4])[(size_t)y * (W >> 4) + x] = v;
  }
}

void launch_ingest(lsd_ctx *ctx, const uint8_t *d_src, size_t srcPitch, size_t srcFrameStride, uint8_t *const *d_slabs, int n,
                   cudaStream_t st) {
  dim3 grid((ctx->w + TILE_W - 1) / TILE_W, (ctx->h + TILE_H - 1) / TILE_H, n);
  k_ingest<<<grid, 256, 0, st>>>(d_src, srcPitch, srcFrameStride, d_slabs, ctx->lay, ctx->w, ctx->h);
  ctx->launches++;
}

// ---------------------------------------------------------------------------------------
// k_gradients: (gx, gy, I, 0) for the LINEAR index range [w, w(h-1)); rows 0 and h-1 are
// zero (upstream leaves them unwritten); x=0 / x=w-1 use the wrapped linear neighbours.
// One launch covers levels lvlLo..lvlHi of n frames.
// ---------------------------------------------------------------------------------------
struct LevelSpan {
  int start[NL + 1];  // prefix of pixel counts over the covered levels
  int lvlLo, lvlHi;
};

// initMask: the level-1 threads also create the frame's refPixelWasGood plane (0xFF, Frame::refPixelWasGood()): a new frame
// then needs no mask-initialisation launch before it is tracked
__global__ void __launch_bounds__(256) k_gradients(uint8_t *const *__restrict__ slabs, FrameLayout lay, Intrinsics K, LevelSpan sp, int initMask) {
  const int f = blockIdx.y;
  uint8_t *slab = slabs[f];
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= sp.start[sp.lvlHi - sp.lvlLo + 1]) return;
  int l = sp.lvlLo;
  while (g >= sp.start[l - sp.lvlLo + 1]) l++;
  const int i = g - sp.start[l - sp.lvlLo];
  const int W = K.w[l], H = K.h[l];
  const float *I = reinterpret_cast<const float *>(slab + lay.img[l]);
  float4 out = make_float4(0, 0, 0, 0);
  if (i >= W && i < W * (H - 1)) {
    out.x = 0.5f * (__ldg(I + i + 1) - __ldg(I + i - 1));
    out.y = 0.5f * (__ldg(I + i + W) - __ldg(I + i - W));
    out.z = __ldg(I + i);
  }
  reinterpret_cast<float4 *>(slab + lay.grad[l])[i] = out;
  if (initMask && l == 1) (slab + lay.mask)[i] = 0xFF;
}

void launch_gradients(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, int lvlLo, int lvlHi, cudaStream_t st, bool initMask) {
  LevelSpan sp;
  sp.lvlLo = lvlLo;
  sp.lvlHi = lvlHi;
  int acc = 0;
  for (int l = lvlLo; l <= lvlHi; l++) {
    sp.start[l - lvlLo] = acc;
    acc += ctx->K.w[l] * ctx->K.h[l];
  }
  sp.start[lvlHi - lvlLo + 1] = acc;
  dim3 grid((acc + 255) / 256, n);
  k_gradients<<<grid, 256, 0, st>>>(d_slabs, ctx->lay, ctx->K, sp, (initMask && lvlLo <= 1 && lvlHi >= 1) ? 1 : 0);
  ctx->launches++;
}

// ---------------------------------------------------------------------------------------
// k_maxgrad0: |g| -> vertical 3-max -> horizontal 3-max, fused through shared memory, with
// the upstream linear-index ranges ([w, w(h-1)) for |g|; [w+1, w(h-1)-1) for both max
// passes; everything unwritten is 0).  Counts numMappablePixels (out >= 5 inside the range).
// ---------------------------------------------------------------------------------------
#define MG_TX 32
#define MG_TY 8

__device__ __forceinline__ float absgrad_lin(const float *__restrict__ I, int k, int W, int H) {
  if (k < W || k >= W * (H - 1)) return 0.0f;
  const float dx = 0.5f * (__ldg(I + k + 1) - __ldg(I + k - 1));
  const float dy = 0.5f * (__ldg(I + k + W) - __ldg(I + k - W));
  return sqrtf(dx * dx + dy * dy);
}

__global__ void __launch_bounds__(MG_TX *MG_TY) k_maxgrad0(uint8_t *const *__restrict__ slabs, FrameLayout lay, int W, int H) {
  __shared__ float sm[MG_TY + 2][MG_TX + 2];
  __shared__ float st[MG_TY][MG_TX + 2];
  const int f = blockIdx.z;
  uint8_t *slab = slabs[f];
  const float *I = reinterpret_cast<const float *>(slab + lay.img[0]);
  float *out = reinterpret_cast<float *>(slab + lay.maxgrad);
  const int x0 = blockIdx.x * MG_TX, y0 = blockIdx.y * MG_TY;
  const int t = threadIdx.y * MG_TX + threadIdx.x;
  const int lo = W + 1, hi = W * (H - 1) - 1;  // [lo, hi)
  for (int c = t; c < (MG_TY + 2) * (MG_TX + 2); c += MG_TX * MG_TY) {
    const int cy = c / (MG_TX + 2), cx = c % (MG_TX + 2);
    const int k = (y0 + cy - 1) * W + (x0 + cx - 1);  // linear semantics: cx-1 == -1 wraps to the previous row
    sm[cy][cx] = absgrad_lin(I, k, W, H);
  }
  __syncthreads();
  for (int c = t; c < MG_TY * (MG_TX + 2); c += MG_TX * MG_TY) {
    const int cy = c / (MG_TX + 2), cx = c % (MG_TX + 2);
    const int k = (y0 + cy) * W + (x0 + cx - 1);
    float v = 0.0f;
    if (k >= lo && k < hi) {
      float g1 = sm[cy][cx];
      const float g2 = sm[cy + 1][cx];
      if (g1 < g2) g1 = g2;
      const float g3 = sm[cy + 2][cx];
      v = (g1 < g3) ? g3 : g1;
    }
    st[cy][cx] = v;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  int mappable = 0;
  if (x < W && y < H) {
    const int i = y * W + x;
    float r;
    if (i >= lo && i < hi) {
      float g1 = st[threadIdx.y][threadIdx.x];
      const float g2 = st[threadIdx.y][threadIdx.x + 1];
      if (g1 < g2) g1 = g2;
      const float g3 = st[threadIdx.y][threadIdx.x + 2];
      r = (g1 < g3) ? g3 : g1;
      mappable = r >= LSD_MIN_USE_GRAD;
    } else {
      r = sm[threadIdx.y + 1][threadIdx.x + 1];  // i == w or i == w(h-1)-1 keep |g|; border rows are 0
    }
    out[i] = r;
  }
  const int cnt = __syncthreads_count(mappable);
  if (t == 0 && cnt) atomicAdd(reinterpret_cast<int *>(slab + lay.total - 16), cnt);
}

__global__ void k_zero_counters(uint8_t *const *__restrict__ slabs, FrameLayout lay, int n) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n) *reinterpret_cast<int *>(slabs[f] + lay.total - 16) = 0;
}

void launch_maxgrad0(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st) {
  k_zero_counters<<<(n + 127) / 128, 128, 0, st>>>(d_slabs, ctx->lay, n);
  dim3 grid((ctx->w + MG_TX - 1) / MG_TX, (ctx->h + MG_TY - 1) / MG_TY, n);
  k_maxgrad0<<<grid, dim3(MG_TX, MG_TY), 0, st>>>(d_slabs, ctx->lay, ctx->w, ctx->h);
  ctx->launches += 2;
}

// ---------------------------------------------------------------------------------------
// k_idepth_pyramid: Frame::buildIDepthAndIDepthVar for levels 1..4 in one pass (64x16 tile).
// Children order (2x,2y),(2x+1,2y),(2x,2y+1),(2x+1,2y+1); only var > 0 children fuse.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fuse4(const float id[4], const float var[4], float &oid, float &ovar) {
  float idepthSumsSum = 0, ivarSumsSum = 0;
  int num = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (var[k] > 0) {
      const float ivar = 1.0f / var[k];
      ivarSumsSum += ivar;
      idepthSumsSum += ivar * id[k];
      num++;
    }
  }
  if (num > 0) {
    const float depth = ivarSumsSum / idepthSumsSum;
    oid = 1.0f / depth;
    ovar = num / ivarSumsSum;
  } else {
    oid = -1;
    ovar = -1;
  }
}

// FROM_MAP: level 0 does not exist yet -- it is Frame::setDepth of the keyframe's depth map (hypothesis planes meta / idepth_smoothed
// / idepth_var_smoothed), computed here, written to the level-0 planes and consumed from registers: one pass over the map instead
// of k_depth_set_depth followed by a re-read of the two planes it wrote.
// STATS: Frame::setDepth's meanIdepth / numPoints (mean of idepth over the level-0 pixels with idepthVar > 0) ride along: every
// CTA reduces its tile in a fixed order (fp64), the last CTA of a frame (ticket) sums the per-CTA partials in a fixed order.
// scratch: per frame a 16-byte ticket + 16 bytes per CTA (stats_stride); out2: (mean, count as int bits) per frame.
#ifndef IDP_MINB
#define IDP_MINB 8  // 32 registers, no spills: setDepth + pyramids 0.127 -> 0.110 ms per 64 keyframes (r02za)
#endif
template <bool FROM_MAP, bool ONE = false>
__global__ void __launch_bounds__(256, IDP_MINB) k_idepth_pyramid(uint8_t *const *__restrict__ slabs, const IdepthMapSrc *__restrict__ srcs,
                                                        FrameLayout lay, int W, int H, uint8_t *__restrict__ statScratch, size_t statStride,
                                                        float *__restrict__ statOut2, uint8_t *slab0, const IdepthMapSrc src0) {
  // ONE: a single frame whose slab / map planes ride in the kernel parameters (per-frame setDepth: no pointer-table uploads)
  __shared__ float a1[TILE_H / 2][TILE_W / 2], b1[TILE_H / 2][TILE_W / 2];
  __shared__ double s_wsum[8];
  __shared__ int s_wcnt[8];
  __shared__ float a2[TILE_H / 4][TILE_W / 4], b2[TILE_H / 4][TILE_W / 4];
  __shared__ float a3[TILE_H / 8][TILE_W / 8], b3[TILE_H / 8][TILE_W / 8];
  const int f = blockIdx.z;
  uint8_t *slab = ONE ? slab0 : slabs[f];
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const int t = threadIdx.x;
  {
    const int tx = t & 31, ty = t >> 5;
    const int x = (x0 >> 1) + tx, y = (y0 >> 1) + ty;
    float oid = -1, ovar = -1;
    double ssum = 0;
    int scnt = 0;
    if (x < (W >> 1) && y < (H >> 1)) {
      float *ID = reinterpret_cast<float *>(slab + lay.idepth[0]);
      float *VR = reinterpret_cast<float *>(slab + lay.idvar[0]);
      const size_t base = (size_t)(2 * y) * W + 2 * x;
      float2 i0, i1, v0, v1;
      if (FROM_MAP) {
        const IdepthMapSrc S = ONE ? src0 : srcs[f];
        const uint2 m0 = *reinterpret_cast<const uint2 *>(S.meta + base), m1 = *reinterpret_cast<const uint2 *>(S.meta + base + W);
        i0 = *reinterpret_cast<const float2 *>(S.ids + base); i1 = *reinterpret_cast<const float2 *>(S.ids + base + W);
        v0 = *reinterpret_cast<const float2 *>(S.vars + base); v1 = *reinterpret_cast<const float2 *>(S.vars + base + W);
        // Frame::setDepth: valid hypotheses with idepth_smoothed >= -0.05 carry (idepth_smoothed, idepth_var_smoothed), the rest (-1, -1)
        if (!((m0.x & 1u) && i0.x >= -0.05f)) i0.x = v0.x = -1;
        if (!((m0.y & 1u) && i0.y >= -0.05f)) i0.y = v0.y = -1;
        if (!((m1.x & 1u) && i1.x >= -0.05f)) i1.x = v1.x = -1;
        if (!((m1.y & 1u) && i1.y >= -0.05f)) i1.y = v1.y = -1;
        *reinterpret_cast<float2 *>(ID + base) = i0; *reinterpret_cast<float2 *>(ID + base + W) = i1;
        *reinterpret_cast<float2 *>(VR + base) = v0; *reinterpret_cast<float2 *>(VR + base + W) = v1;
      } else {
        i0 = *reinterpret_cast<const float2 *>(ID + base); i1 = *reinterpret_cast<const float2 *>(ID + base + W);
        v0 = *reinterpret_cast<const float2 *>(VR + base); v1 = *reinterpret_cast<const float2 *>(VR + base + W);
      }
      const float id[4] = {i0.x, i0.y, i1.x, i1.y}, var[4] = {v0.x, v0.y, v1.x, v1.y};
      fuse4(id, var, oid, ovar);
      reinterpret_cast<float *>(slab + lay.idepth[1])[(size_t)y * (W >> 1) + x] = oid;
      reinterpret_cast<float *>(slab + lay.idvar[1])[(size_t)y * (W >> 1) + x] = ovar;
      if (statScratch) {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (var[k] > 0) {
            ssum += (double)id[k];
            scnt++;
          }
      }
    }
    a1[ty][tx] = oid;
    b1[ty][tx] = ovar;
    if (statScratch) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        ssum += __shfl_down_sync(0xffffffffu, ssum, o);
        scnt += __shfl_down_sync(0xffffffffu, scnt, o);
      }
      if ((t & 31) == 0) {
        s_wsum[t >> 5] = ssum;
        s_wcnt[t >> 5] = scnt;
      }
    }
  }
  __syncthreads();
  if (statScratch && t < 32) {  // warp 0: CTA partial, ticket, and -- in the frame's last CTA -- the total
    unsigned *ticket = reinterpret_cast<unsigned *>(statScratch + statStride * f);
    double *partial = reinterpret_cast<double *>(statScratch + statStride * f + 16);
    const unsigned nCta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    unsigned last = 0;
    if (t == 0) {
      double S = 0;
      int Cn = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        S += s_wsum[k];
        Cn += s_wcnt[k];
      }
      partial[2 * cta] = S;
      partial[2 * cta + 1] = (double)Cn;
      // gpu-scope acquire-release ticket: releases this CTA's partial, and the CTA that takes the last ticket acquires all others
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(ticket), "r"(1u) : "memory");
      last = old == nCta - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    __syncwarp();  // lane 0's acquire precedes the other lanes' L2 reads of the partials
    if (last) {
      double S = 0, Cn = 0;
      for (unsigned k = t; k < nCta; k += 32) {  // lane-strided, then a fixed shuffle tree: a pure function of the partials
        S += __ldcg(partial + 2 * k);
        Cn += __ldcg(partial + 2 * k + 1);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        S += __shfl_down_sync(0xffffffffu, S, o);
        Cn += __shfl_down_sync(0xffffffffu, Cn, o);
      }
      if (t == 0) {
        statOut2[2 * f] = (float)S / (float)Cn;  // upstream: float sum / float count
        statOut2[2 * f + 1] = __int_as_float((int)Cn);
        *ticket = 0;  // ready for the next call
      }
    }
  }
  if (t < 64) {
    const int tx = t & 15, ty = t >> 4;
    const float id[4] = {a1[2 * ty][2 * tx], a1[2 * ty][2 * tx + 1], a1[2 * ty + 1][2 * tx], a1[2 * ty + 1][2 * tx + 1]};
    const float var[4] = {b1[2 * ty][2 * tx], b1[2 * ty][2 * tx + 1], b1[2 * ty + 1][2 * tx], b1[2 * ty + 1][2 * tx + 1]};
    float oid, ovar;
    fuse4(id, var, oid, ovar);
    a2[ty][tx] = oid;
    b2[ty][tx] = ovar;
    const int x = (x0 >> 2) + tx, y = (y0 >> 2) + ty;
    if (x < (W >> 2) && y < (H >> 2)) {
      reinterpret_cast<float *>(slab + lay.idepth[2])[(size_t)y * (W >> 2) + x] = oid;
      reinterpret_cast<float *>(slab + lay.idvar[2])[(size_t)y * (W >> 2) + x] = ovar;
    }
  }
  __syncthreads();
  if (t < 16) {
    const int tx = t & 7, ty = t >> 3;
    const float id[4] = {a2[2 * ty][2 * tx], a2[2 * ty][2 * tx + 1], a2[2 * ty + 1][2 * tx], a2[2 * ty + 1][2 * tx + 1]};
    const float var[4] = {b2[2 * ty][2 * tx], b2[2 * ty][2 * tx + 1], b2[2 * ty + 1][2 * tx], b2[2 * ty + 1][2 * tx + 1]};
    float oid, ovar;
    fuse4(id, var, oid, ovar);
    a3[ty][tx] = oid;
    b3[ty][tx] = ovar;
    const int x = (x0 >> 3) + tx, y = (y0 >> 3) + ty;
    if (x < (W >> 3) && y < (H >> 3)) {
      reinterpret_cast<float *>(slab + lay.idepth[3])[(size_t)y * (W >> 3) + x] = oid;
      reinterpret_cast<float *>(slab + lay.idvar[3])[(size_t)y * (W >> 3) + x] = ovar;
    }
  }
  __syncthreads();
  if (t < 4) {
    const int tx = t;
    const float id[4] = {a3[0][2 * tx], a3[0][2 * tx + 1], a3[1][2 * tx], a3[1][2 * tx + 1]};
    const float var[4] = {b3[0][2 * tx], b3[0][2 * tx + 1], b3[1][2 * tx], b3[1][2 * tx + 1]};
    float oid, ovar;
    fuse4(id, var, oid, ovar);
    const int x = (x0 >> 4) + tx, y = (y0 >> 4);
    if (x < (W >> 4) && y < (H >> 4)) {
      reinterpret_cast<float *>(slab + lay.idepth[4])[(size_t)y * (W >> 4) + x] = oid;
      reinterpret_cast<float *>(slab + lay.idvar[4])[(size_t)y * (W >> 4) + x] = ovar;
    }
  }
}

static size_t stats_stride(const lsd_ctx *ctx);

// d_statOut2 != nullptr: Frame::setDepth's (meanIdepth, numPoints) of every frame as well (ensure_stats_scratch(ctx, n) first)
void launch_idepth_pyramid(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st, float *d_statOut2) {
  dim3 grid((ctx->w + TILE_W - 1) / TILE_W, (ctx->h + TILE_H - 1) / TILE_H, n);
  const IdepthMapSrc none = {nullptr, nullptr, nullptr};
  k_idepth_pyramid<false><<<grid, 256, 0, st>>>(d_slabs, nullptr, ctx->lay, ctx->w, ctx->h, d_statOut2 ? ctx->d_stats : nullptr,
                                                stats_stride(ctx), d_statOut2, nullptr, none);
  ctx->launches++;
}

// Frame::setDepth(depth map) + buildIDepthAndIDepthVar in one pass (d_srcs[i]: the hypothesis planes of keyframe i's map)
void launch_set_depth_and_pyramid(lsd_ctx *ctx, uint8_t *const *d_slabs, const IdepthMapSrc *d_srcs, int n, cudaStream_t st,
                                  float *d_statOut2) {
  dim3 grid((ctx->w + TILE_W - 1) / TILE_W, (ctx->h + TILE_H - 1) / TILE_H, n);
  const IdepthMapSrc none = {nullptr, nullptr, nullptr};
  k_idepth_pyramid<true><<<grid, 256, 0, st>>>(d_slabs, d_srcs, ctx->lay, ctx->w, ctx->h, d_statOut2 ? ctx->d_stats : nullptr,
                                               stats_stride(ctx), d_statOut2, nullptr, none);
  ctx->launches++;
}

// the same for ONE keyframe, slab and map planes passed in the kernel parameters (no pointer tables to upload)
void launch_set_depth_and_pyramid_one(lsd_ctx *ctx, uint8_t *slab, const IdepthMapSrc &src, cudaStream_t st, float *d_statOut2) {
  dim3 grid((ctx->w + TILE_W - 1) / TILE_W, (ctx->h + TILE_H - 1) / TILE_H, 1);
  k_idepth_pyramid<true, true><<<grid, 256, 0, st>>>(nullptr, nullptr, ctx->lay, ctx->w, ctx->h, d_statOut2 ? ctx->d_stats : nullptr,
                                               stats_stride(ctx), d_statOut2, slab, src);
  ctx->launches++;
}

// Frame::setDepthFromGroundTruth
__global__ void k_set_depth_gt(uint8_t *slab, FrameLayout lay, const float *__restrict__ depth, float var, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float d = depth[i];
  float id = -1, v = -1;
  if (d > 0) {
    id = 1.0f / d;
    v = var;
  }
  reinterpret_cast<float *>(slab + lay.idepth[0])[i] = id;
  reinterpret_cast<float *>(slab + lay.idvar[0])[i] = v;
}

void launch_set_depth_gt(lsd_ctx *ctx, uint8_t *slab, const float *d_depth, float cov, cudaStream_t st) {
  const int N = ctx->w * ctx->h;
  k_set_depth_gt<<<(N + 255) / 256, 256, 0, st>>>(slab, ctx->lay, d_depth, LSD_VAR_GT_INIT_INITIAL * cov, N);
  ctx->launches++;
}

// Frame::refPixelWasGood(): created as 0xFF
__global__ void k_mask_init(uint8_t *const *__restrict__ slabs, FrameLayout lay, int words) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < words) reinterpret_cast<uint32_t *>(slabs[blockIdx.y] + lay.mask)[i] = 0xFFFFFFFFu;
}

void launch_mask_init(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st) {
  const int words = (ctx->K.w[1] * ctx->K.h[1] + 3) / 4;
  dim3 grid((words + 255) / 256, n);
  k_mask_init<<<grid, 256, 0, st>>>(d_slabs, ctx->lay, words);
  ctx->launches++;
}

// meanIdepth / numPoints of a level-0 idepth plane (Frame::setDepth bookkeeping, read by the keyframe-selection score on
// every tracked frame).  One CTA per SM: per-thread fp64 partial sums over a grid-strided slice, fixed-order block tree,
// per-CTA partials to global memory, and the CTA that takes the last ticket adds them in CTA order -- deterministic and
// order-independent to fp32 rounding (a single 1024-thread CTA took 46 us per frame on the live path).
#define STATS_THREADS 256
// blockIdx.y = frame: n frames in one launch, each with its own ticket / partial-sum scratch and the SAME decomposition as a
// single-frame launch (gridDim.x CTAs, grid-strided slices), so a frame's result does not depend on how many ride along.
__global__ void __launch_bounds__(STATS_THREADS) k_idepth_stats(uint8_t *const *__restrict__ slabs, const uint8_t *slab0, FrameLayout lay, int N,
                                                                uint8_t *__restrict__ scratch, size_t scratchStride, float *__restrict__ out2) {
  __shared__ double ssum[STATS_THREADS / 32];
  __shared__ int scnt[STATS_THREADS / 32];
  __shared__ bool sLast;
  const uint8_t *slab = slabs ? slabs[blockIdx.y] : slab0;
  unsigned *ticket = reinterpret_cast<unsigned *>(scratch + scratchStride * blockIdx.y);
  double *partial = reinterpret_cast<double *>(scratch + scratchStride * blockIdx.y + 16);
  out2 += 2 * blockIdx.y;
  const float *ID = reinterpret_cast<const float *>(slab + lay.idepth[0]);
  const float *VR = reinterpret_cast<const float *>(slab + lay.idvar[0]);
  double s = 0;
  int c = 0;
  // 16-byte vectors of both planes, two in flight per thread (N % 4 == 0: w and h are multiples of 16); the fp64 sum of fp32
  // values is exact far beyond the final fp32 rounding, so the grouping does not show in the result
  const float4 *ID4 = reinterpret_cast<const float4 *>(ID), *VR4 = reinterpret_cast<const float4 *>(VR);
  const int N4 = N >> 2, stride = gridDim.x * STATS_THREADS;
  for (int i = blockIdx.x * STATS_THREADS + threadIdx.x; i < N4; i += 2 * stride) {
    const bool two = i + stride < N4;
    const float4 v0 = VR4[i], d0 = ID4[i];
    const float4 v1 = two ? VR4[i + stride] : make_float4(0, 0, 0, 0), d1 = two ? ID4[i + stride] : make_float4(0, 0, 0, 0);
    if (v0.x > 0) { s += (double)d0.x; c++; }
    if (v0.y > 0) { s += (double)d0.y; c++; }
    if (v0.z > 0) { s += (double)d0.z; c++; }
    if (v0.w > 0) { s += (double)d0.w; c++; }
    if (v1.x > 0) { s += (double)d1.x; c++; }
    if (v1.y > 0) { s += (double)d1.y; c++; }
    if (v1.z > 0) { s += (double)d1.z; c++; }
    if (v1.w > 0) { s += (double)d1.w; c++; }
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    scnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0;
    int Cn = 0;
    for (int k = 0; k < STATS_THREADS / 32; k++) {
      S += ssum[k];
      Cn += scnt[k];
    }
    partial[2 * blockIdx.x] = S;
    partial[2 * blockIdx.x + 1] = (double)Cn;
    __threadfence();
    sLast = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (sLast && threadIdx.x == 0) {
    __threadfence();
    double S = 0, Cn = 0;
    for (unsigned k = 0; k < gridDim.x; k++) {
      S += __ldcg(partial + 2 * k);
      Cn += __ldcg(partial + 2 * k + 1);
    }
    out2[0] = (float)S / (float)Cn;  // upstream: float sum / float count
    out2[1] = __int_as_float((int)Cn);
    *ticket = 0;  // ready for the next call
  }
}

// scratch: 16 + 16 bytes per CTA of a frame, zero-initialised once, owned by the context
// (k_idepth_stats runs numSMs CTAs per frame, k_idepth_pyramid one per 64x16 tile)
static size_t stats_stride(const lsd_ctx *ctx) {
  const size_t tiles = (size_t)((ctx->w + TILE_W - 1) / TILE_W) * ((ctx->h + TILE_H - 1) / TILE_H);
  const size_t ctas = tiles > (size_t)ctx->numSMs ? tiles : (size_t)ctx->numSMs;
  return (16 + 16 * ctas + 255) / 256 * 256;
}

int ensure_stats_scratch(lsd_ctx *ctx, int frames) {
  if (frames <= ctx->statsFrames) return LSD_OK;
  if (ctx->d_stats) cudaFree(ctx->d_stats);
  ctx->d_stats = nullptr;
  ctx->statsFrames = 0;
  const int cap = frames < 8 ? 8 : frames * 2;
  LSD_CUDA(cudaMalloc(&ctx->d_stats, stats_stride(ctx) * (size_t)cap));
  LSD_CUDA(cudaMemsetAsync(ctx->d_stats, 0, stats_stride(ctx) * (size_t)cap, ctx->stream));
  ctx->statsFrames = cap;
  return LSD_OK;
}

void launch_idepth_stats(lsd_ctx *ctx, uint8_t *slab, float *d_out2, cudaStream_t st) {
  k_idepth_stats<<<ctx->numSMs, STATS_THREADS, 0, st>>>(nullptr, slab, ctx->lay, ctx->w * ctx->h, ctx->d_stats, stats_stride(ctx), d_out2);
  ctx->launches++;
}

// n frames (d_slabs: device pointer list) in one launch; d_out2: 2 floats per frame
void launch_idepth_stats_batch(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, float *d_out2, cudaStream_t st) {
  k_idepth_stats<<<dim3(ctx->numSMs, n), STATS_THREADS, 0, st>>>(d_slabs, nullptr, ctx->lay, ctx->w * ctx->h, ctx->d_stats, stats_stride(ctx), d_out2);
  ctx->launches++;
}

}  // namespace lsd
