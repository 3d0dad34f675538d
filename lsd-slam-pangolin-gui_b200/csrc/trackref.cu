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

// chunk geometry: level l owns grid columns [first[l], first[l+1]); a chunk is PC_CHUNK consecutive pixels
#define PC_CHUNK 1024
struct PCChunks {
  int first[NL + 1];
};

// One CTA per 1024-pixel raster chunk of one (keyframe, level); the chunk's output offset comes from a decoupled
// look-back over the level's earlier chunks (common.cuh), so a single keyframe is compacted by ~100 CTAs at once
// (the live pipeline re-imports the reference after every depth update) and the emission order stays raster order.
// Look-back state lives behind the reference's counters: nums[f][64 + l] = chunk ticket of level l,
// nums[f][64 + NL + first[l] + c] = state of chunk c; zeroed by the host before the launch.
__global__ void __launch_bounds__(PC_CHUNK) k_make_pointcloud(uint8_t *const *__restrict__ kfSlabs, uint8_t *const *__restrict__ refSlabs,
                                                              int *const *__restrict__ nums, FrameLayout lay, Intrinsics K,
                                                              PCOffsets off, PCChunks ch) {
  __shared__ int warpTot[32];
  __shared__ unsigned s_chunk, s_base;
  int level = 1;
  while (level < NL - 1 && (int)blockIdx.x >= ch.first[level + 1]) level++;
  const int f = blockIdx.y;
  const uint8_t *kf = kfSlabs[f];
  uint8_t *rs = refSlabs[f];
  unsigned *lb = reinterpret_cast<unsigned *>(nums[f]) + 64;
  unsigned *state = lb + NL + ch.first[level];
  const int nChunks = ch.first[level + 1] - ch.first[level];
  const int W = K.w[level], H = K.h[level], N = W * H;
  const float *ID = reinterpret_cast<const float *>(kf + lay.idepth[level]);
  const float *VR = reinterpret_cast<const float *>(kf + lay.idvar[level]);
  const float4 *G = reinterpret_cast<const float4 *>(kf + lay.grad[level]);
  RefPoint *pts = reinterpret_cast<RefPoint *>(rs + off.pts[level]);
  float2 *gr = reinterpret_cast<float2 *>(rs + off.grad[level]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_chunk = atomicAdd(lb + level, 1u);
  __syncthreads();
  const unsigned chunk = s_chunk;
  const int i = (int)chunk * PC_CHUNK + threadIdx.x;
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
  if (wid == 0) {
    const int v = warpTot[lane];
    int s = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    warpTot[lane] = s - v;  // exclusive
    const unsigned total = (unsigned)__shfl_sync(0xffffffffu, s, 31);
    const unsigned excl = lookback_exclusive(state, chunk, total, lane);
    if (lane == 0) {
      s_base = excl;
      if ((int)chunk == nChunks - 1) nums[f][level] = (int)(excl + total);
    }
  }
  __syncthreads();
  if (keep) {
    const int o = (int)s_base + warpTot[wid] + inWarp;
    const float4 g = G[i];
    RefPoint p;
    p.xy = (uint32_t)x | ((uint32_t)y << 16);
    p.invDepth = 1.0f / id;  // == posData z (upstream: pos = (1/idepth) * (...))
    p.color = g.z;  // gradients.z == image(level)
    p.var = var;
    pts[o] = p;
    gr[o] = make_float2(g.x, g.y);
  }
}

int pointcloud_state_words(const lsd_ctx *ctx) {
  int total = 0;
  for (int l = 1; l < NL; l++) total += (ctx->K.w[l] * ctx->K.h[l] + PC_CHUNK - 1) / PC_CHUNK;
  return 64 + NL + total;
}

// numData + look-back state of n references, zeroed in one launch (one memset per reference cost ~3 us of launch time each)
__global__ void k_zero_ref_state(int *const *__restrict__ nums, int words) {
  int *p = nums[blockIdx.x];
  for (int i = threadIdx.x; i < words; i += blockDim.x) p[i] = 0;
}

void launch_make_pointcloud(lsd_ctx *ctx, uint8_t *const *d_kfSlabs, uint8_t *const *d_refSlabs, int *const *d_nums, int n,
                            const size_t *offPts, const size_t *offGrad, cudaStream_t st) {
  k_zero_ref_state<<<n, 128, 0, st>>>(d_nums, pointcloud_state_words(ctx));
  ctx->launches++;
  PCOffsets off;
  for (int l = 0; l < NL; l++) {
    off.pts[l] = offPts[l];
    off.grad[l] = offGrad[l];
  }
  PCChunks ch;
  ch.first[0] = ch.first[1] = 0;
  for (int l = 1; l < NL; l++) ch.first[l + 1] = ch.first[l] + (ctx->K.w[l] * ctx->K.h[l] + PC_CHUNK - 1) / PC_CHUNK;
  dim3 grid(ch.first[NL], n);
  k_make_pointcloud<<<grid, PC_CHUNK, 0, st>>>(d_kfSlabs, d_refSlabs, d_nums, ctx->lay, ctx->K, off, ch);
  ctx->launches++;
}

}  // namespace lsd
