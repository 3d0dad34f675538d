// trackref.cu -- [UP] TrackingReference::makePointCloud(level) for levels 1..4.
//
// Ordered stream compaction of the keyframe's semi-dense pixels (var > 0 && idepth != 0,
// x in [1,w-1), y in [1,h-1)).  numData[level] is bit-exact with upstream.  Emission order is
// ROW-MAJOR here (upstream walks x outer / y inner): consecutive points are then neighbours in
// x, so the 16-byte bilinear taps of a warp in the tracker fall into the same 128-byte lines.
// Order only affects fp32 summation order (hard part 6).
//
// One point = 16 B {x|y<<16, 1/idepth, colour, var}; the 3-D position
// (1/idepth)*(fxi*x+cxi, fyi*y+cyi, 1) is recomputed by the tracker with the same operations,
// so it is bit-identical to upstream's stored posData.  grad (Sim3 only) is a separate float2 plane.
#include "ctx.cuh"

namespace lsd {

struct PCOffsets {
  size_t pts[NL], grad[NL];
};

__global__ void __launch_bounds__(1024) k_make_pointcloud(uint8_t *const *__restrict__ kfSlabs, uint8_t *const *__restrict__ refSlabs,
                                                          int *const *__restrict__ nums, FrameLayout lay, Intrinsics K,
                                                          PCOffsets off) {
  __shared__ int warpTot[32];
  __shared__ int chunkBase;
  const int level = 1 + blockIdx.x;
  const int f = blockIdx.y;
  const uint8_t *kf = kfSlabs[f];
  uint8_t *rs = refSlabs[f];
  const int W = K.w[level], H = K.h[level], N = W * H;
  const float *ID = reinterpret_cast<const float *>(kf + lay.idepth[level]);
  const float *VR = reinterpret_cast<const float *>(kf + lay.idvar[level]);
  const float4 *G = reinterpret_cast<const float4 *>(kf + lay.grad[level]);
  RefPoint *pts = reinterpret_cast<RefPoint *>(rs + off.pts[level]);
  float2 *gr = reinterpret_cast<float2 *>(rs + off.grad[level]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) chunkBase = 0;
  __syncthreads();
  for (int i0 = 0; i0 < N; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    bool keep = false;
    float id = 0, var = 0;
    int x = 0, y = 0;
    if (i < N) {
      y = i / W;
      x = i - y * W;
      if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
        id = ID[i];
        var = VR[i];
        keep = !(var <= 0 || id == 0);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int inWarp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warpTot[wid] = __popc(bal);
    __syncthreads();
    int v = 0;
    if (wid == 0) {
      v = warpTot[lane];
      int s = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += n;
      }
      warpTot[lane] = s - v;  // exclusive
      if (lane == 31) v = s;  // total in lane 31
    }
    __syncthreads();
    const int base = chunkBase;
    if (keep) {
      const int o = base + warpTot[wid] + inWarp;
      const float4 g = G[i];
      RefPoint p;
      p.xy = (uint32_t)x | ((uint32_t)y << 16);
      p.invDepth = 1.0f / id;  // == posData z (upstream: pos = (1/idepth) * (...))
      p.color = g.z;  // gradients.z == image(level)
      p.var = var;
      pts[o] = p;
      gr[o] = make_float2(g.x, g.y);
    }
    __syncthreads();
    if (threadIdx.x == 31) chunkBase = base + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) nums[f][level] = chunkBase;
}

void launch_make_pointcloud(lsd_ctx *ctx, uint8_t *const *d_kfSlabs, uint8_t *const *d_refSlabs, int *const *d_nums, int n,
                            const size_t *offPts, const size_t *offGrad, cudaStream_t st) {
  PCOffsets off;
  for (int l = 0; l < NL; l++) {
    off.pts[l] = offPts[l];
    off.grad[l] = offGrad[l];
  }
  dim3 grid(NL - 1, n);
  k_make_pointcloud<<<grid, 1024, 0, st>>>(d_kfSlabs, d_refSlabs, d_nums, ctx->lay, ctx->K, off);
  ctx->launches++;
}

}  // namespace lsd
