// keyframe.cu -- keyframe publish + point-cloud (VBO) extraction on device.
//
// Replaces, for the step that follows the hot path on every keyframe:
//   * PangolinOutputIOWrapper::publishKeyframe's pack loop
//     (/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:69-89): idepth / idepthVar / image of the
//     publish level -> 12-byte InputPointDense records (Keyframe.h:16-21)                       -> k_publish_pack
//   * Keyframe::computeVbo (/root/reference/lib/Pangolin_IOWrapper/Keyframe.h:66-158): variance / scale filter,
//     3x3 near-support test, back-projection, raster-ordered compaction into MyVertex[] (Keyframe.h:47-51)
//                                                                                                -> k_vbo_extract
// The reference runs (1) on a core worker thread and (2) on the GL thread over the packed copy; here (2) reads the
// frame's planes directly (the pack is only produced when a host consumer asks for it), so a keyframe costs
// 12 B/px of reads + 16 B per emitted vertex.
//
// k_vbo_extract is ONE pass: a CTA owns a 4096-pixel raster chunk (256 threads x four float4 groups of each plane), counts
// its survivors, and obtains its output offset by decoupled look-back over the chunks before it (chunk ids are
// handed out by an atomic ticket, so every predecessor is already resident and the spin cannot deadlock).  The
// vertices therefore land in exactly the reference's raster order and `points` is exact.  blockIdx.y = keyframe.
// Arithmetic is statement-for-statement the reference's (library built with -fmad=false; IEEE division), so the
// vertex buffer is compared bit for bit against the reference's own Keyframe.h (tests/test_gpu_keyframe.py).
#include <cstring>

#include "ctx.cuh"

namespace lsd {

#define VBO_THREADS 256
#define VBO_PX_PER_THREAD 4
#define VBO_SUB 4  // 1024-pixel sub-chunks per CTA: one ticket and one look-back per 4096 pixels
#define VBO_CHUNK (VBO_THREADS * VBO_PX_PER_THREAD * VBO_SUB)
#define VBO_VALUE_MASK LSD_LB_MASK

struct VboJob {
  const float *idepth, *var, *img;
  uint4 *out;        // MyVertex[capacity = w*h]
  unsigned *state;   // one word per chunk: flag << 30 | count;  state[nChunks] is the chunk ticket
  int *points;       // emitted vertex count
  float scale;       // camToWorld.scale()
};

struct VboK {
  int W, H, nChunks;
  float fxi, fyi, cxi, cyi;
  float scaledTH, absTH;
  int minNearSupport, contractFma;
};

// per-pixel filter (Keyframe.h:95-134) on 4 consecutive pixels of one row.  Every load of the thread is issued up front
// and unconditionally (one memory round trip instead of three dependent ones): 3 rows of idepth, var, image.
__device__ __forceinline__ unsigned vbo_filter4(const VboJob &J, const VboK &P, int p0, int N, float depthK[4], unsigned &colPacked) {
  unsigned keep = 0;
  colPacked = 0;
  if (p0 >= N) return 0;
  const int y = p0 / P.W, x0 = p0 - y * P.W;
  if (y < 1 || y >= P.H - 1) return 0;
  float nb[3][6];  // rows y-1, y, y+1 x columns x0-1 .. x0+4 of idepth
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const float *row = J.idepth + p0 + (r - 1) * P.W;
    const float4 q = __ldg(reinterpret_cast<const float4 *>(row));
    nb[r][1] = q.x; nb[r][2] = q.y; nb[r][3] = q.z; nb[r][4] = q.w;
    nb[r][0] = x0 > 0 ? __ldg(row - 1) : 0.0f;          // unused when x0 == 0 (pixel x = 0 is never emitted)
    nb[r][5] = x0 + 4 < P.W ? __ldg(row + 4) : 0.0f;    // unused when x0 + 3 == W - 1
  }
  const float4 v4 = __ldg(reinterpret_cast<const float4 *>(J.var + p0));
  const float4 c4 = __ldg(reinterpret_cast<const float4 *>(J.img + p0));
  const float vc[4] = {v4.x, v4.y, v4.z, v4.w};
  // publishKeyframe: float -> unsigned char (truncation)
  colPacked = (unsigned)(unsigned char)c4.x | ((unsigned)(unsigned char)c4.y << 8) | ((unsigned)(unsigned char)c4.z << 16) |
              ((unsigned)(unsigned char)c4.w << 24);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = x0 + j;
    const float idc = nb[1][j + 1];
    if (x < 1 || x >= P.W - 1) continue;
    if (idc <= 0) continue;
    const float depth = 1 / idc;
    float depth4 = depth * depth;
    depth4 *= depth4;
    if (vc[j] * depth4 > P.scaledTH) continue;
    if (vc[j] * depth4 * J.scale * J.scale > P.absTH) continue;
    if (P.minNearSupport > 1) {
      int nearSupport = 0;
#pragma unroll
      for (int dx = 0; dx < 3; dx++)
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
          const float nid = nb[dy][j + dx];
          if (nid > 0) {
            const float diff = nid - 1.0f / depth;
            if (diff * diff < 2 * vc[j]) nearSupport++;
          }
        }
      if (nearSupport < P.minNearSupport) continue;
    }
    keep |= 1u << j;
    depthK[j] = depth;
  }
  return keep;
}

__global__ void __launch_bounds__(VBO_THREADS) k_vbo_extract(const VboJob *__restrict__ jobs, const VboK P) {
  __shared__ unsigned s_chunk, s_base;
  __shared__ unsigned s_cnt[VBO_SUB][VBO_THREADS / 32];  // survivors per (sub-chunk, warp); exclusive offsets after the scan
  const VboJob J = jobs[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_chunk = atomicAdd(J.state + P.nChunks, 1u);
  __syncthreads();
  const unsigned chunk = s_chunk;
  const int N = P.W * P.H;

  float depthK[VBO_SUB][4];
  unsigned keep[VBO_SUB], col[VBO_SUB], inclW[VBO_SUB];
#pragma unroll
  for (int sb = 0; sb < VBO_SUB; sb++) {
    const int p0 = (int)chunk * VBO_CHUNK + sb * (VBO_THREADS * VBO_PX_PER_THREAD) + tid * VBO_PX_PER_THREAD;
    keep[sb] = vbo_filter4(J, P, p0, N, depthK[sb], col[sb]);
    unsigned incl = __popc(keep[sb]);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    inclW[sb] = incl;
    if (lane == 31) s_cnt[sb][warp] = incl;
  }
  __syncthreads();

  // ---- raster order inside the chunk = sub-chunk, then warp: scan the 32 (sub-chunk, warp) totals, then one
  // ---- decoupled look-back for the whole 4096-pixel chunk (warp 0)
  if (warp == 0) {
    unsigned *flat = &s_cnt[0][0];  // VBO_SUB * 8 == 32 entries
    const unsigned v = flat[lane];
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    flat[lane] = incl - v;
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned excl = lookback_exclusive(J.state, chunk, total, lane);
    if (lane == 0) {
      s_base = excl;
      if ((int)chunk == P.nChunks - 1) *J.points = (int)(excl + total);
    }
  }
  __syncthreads();

  // ---- emit (Keyframe.h:136-143)
#pragma unroll
  for (int sb = 0; sb < VBO_SUB; sb++) {
    if (!keep[sb]) continue;
    const int p0 = (int)chunk * VBO_CHUNK + sb * (VBO_THREADS * VBO_PX_PER_THREAD) + tid * VBO_PX_PER_THREAD;
    const int y = p0 / P.W, x0 = p0 - y * P.W;
    unsigned o = s_base + s_cnt[sb][warp] + (inclW[sb] - __popc(keep[sb]));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (!(keep[sb] >> j & 1u)) continue;
      const int x = x0 + j;
      const float depth = depthK[sb][j];
      float px, py;
      if (P.contractFma) {
        px = __fmaf_rn((float)x, P.fxi, P.cxi) * depth;
        py = __fmaf_rn((float)y, P.fyi, P.cyi) * depth;
      } else {
        px = (x * P.fxi + P.cxi) * depth;
        py = (y * P.fyi + P.cyi) * depth;
      }
      const unsigned g = (col[sb] >> (8 * j)) & 0xffu;
      uint4 vtx;
      vtx.x = __float_as_uint(px);
      vtx.y = __float_as_uint(py);
      vtx.z = __float_as_uint(depth);
      vtx.w = g | (g << 8) | (g << 16) | (100u << 24);  // color[0..2] = b, g, r (all the grey value), color[3] = 100
      J.out[o++] = vtx;
    }
  }
}

// InputPointDense[] of one frame level (3 words per pixel)
__global__ void __launch_bounds__(256) k_publish_pack(const float *__restrict__ idepth, const float *__restrict__ var,
                                                      const float *__restrict__ img, unsigned *__restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned g = (unsigned)(unsigned char)__ldg(img + i);
  out[3 * i] = __float_as_uint(__ldg(idepth + i));
  out[3 * i + 1] = __float_as_uint(__ldg(var + i));
  out[3 * i + 2] = g | (g << 8) | (g << 16) | (g << 24);
}

static size_t kalign(size_t v) { return (v + 255) / 256 * 256; }

static int require_depth(lsd_ctx *ctx, lsd_frame *f, int level) {
  if (!(f->built & FB_IDEPTH0)) {
    set_error("frame " + std::to_string(f->id) + " has no depth (hasIDepthBeenSet() == false)");
    return LSD_ERR_STATE;
  }
  if (level > 0) return frame_ensure_built(ctx, f, FB_IDEPTH_PYR);
  return LSD_OK;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

int lsd_default_vbo_params(lsd_vbo_params *p) {
  LSD_ARG(p);
  p->scaledTH = 1e-3;  // my_scaledTH, Keyframe.h:79 (float initialised from the double literal)
  p->absTH = 1e-1;     // my_absTH, Keyframe.h:80
  p->minNearSupport = 9;
  p->sparsifyFactor = 1;
  p->contractFma = 0;
  return LSD_OK;
}

int lsd_frame_publish_keyframe(lsd_ctx *ctx, lsd_frame *f, int level, lsd_input_point_dense *dst) {
  LSD_ARG(ctx && f && dst);
  LSD_ARG(level >= 0 && level < NL);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const int N = ctx->K.w[level] * ctx->K.h[level];
  if (!(f->built & FB_IDEPTH0)) {  // the reference publishes the buffer as allocated and logs a warning; zero-filled here
    std::memset(dst, 0, sizeof(lsd_input_point_dense) * (size_t)N);
    return LSD_OK;
  }
  int rc = require_depth(ctx, f, level);
  if (rc) return rc;
  rc = ensure_stage(ctx, 0, sizeof(lsd_input_point_dense) * (size_t)N);
  if (rc) return rc;
  const FrameLayout &L = ctx->lay;
  k_publish_pack<<<(N + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const float *>(f->slab + L.idepth[level]),
                                                          reinterpret_cast<const float *>(f->slab + L.idvar[level]),
                                                          reinterpret_cast<const float *>(f->slab + L.img[level]),
                                                          reinterpret_cast<unsigned *>(ctx->d_stage), N);
  ctx->launches++;
  LSD_CUDA(cudaGetLastError());
  LSD_CUDA(cudaMemcpyAsync(dst, ctx->d_stage, sizeof(lsd_input_point_dense) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}

int lsd_keyframe_compute_vbo_batch(lsd_ctx *ctx, int n, lsd_frame *const *frames, int level, const float *camToWorldScale,
                                   const lsd_vbo_params *params, void *d_vertices, lsd_vertex *const *dst, int *points) {
  LSD_ARG(ctx && frames && camToWorldScale && points && n >= 1);
  LSD_ARG(level >= 0 && level < NL);
  lsd_vbo_params prm;
  lsd_default_vbo_params(&prm);
  if (params) prm = *params;
  if (prm.sparsifyFactor != 1) {  // the reference's constant (Keyframe.h:82); >1 would draw libc rand() per pixel on the host
    set_error("computeVbo: sparsifyFactor != 1 is not supported (reference constant my_sparsifyFactor = 1)");
    return LSD_ERR_ARG;
  }
  LSD_CUDA(cudaSetDevice(ctx->device));
  const int W = ctx->K.w[level], H = ctx->K.h[level], N = W * H;
  LSD_ARG(W % 4 == 0);
  const int nChunks = (N + VBO_CHUNK - 1) / VBO_CHUNK;
  LSD_ARG((size_t)N <= VBO_VALUE_MASK);
  int rc;
  for (int i = 0; i < n; i++) {
    LSD_ARG(frames[i]);
    if ((rc = require_depth(ctx, frames[i], level))) return rc;
  }
  // scratch: [vertices n*N*16 (unless caller-provided)] [state n*(nChunks+1) words] [points n ints]
  const size_t vtxBytes = d_vertices ? 0 : kalign((size_t)n * N * sizeof(lsd_vertex));
  const size_t stateBytes = kalign(sizeof(unsigned) * (size_t)n * (nChunks + 1));
  const size_t cntBytes = kalign(sizeof(int) * (size_t)n);
  if ((rc = ensure_stage(ctx, cntBytes, vtxBytes + stateBytes + cntBytes))) return rc;
  uint8_t *base = ctx->d_stage;
  uint4 *d_vtx = d_vertices ? reinterpret_cast<uint4 *>(d_vertices) : reinterpret_cast<uint4 *>(base);
  unsigned *d_state = reinterpret_cast<unsigned *>(base + vtxBytes);
  int *d_points = reinterpret_cast<int *>(base + vtxBytes + stateBytes);
  if ((rc = ensure_table(ctx, kalign(sizeof(VboJob) * (size_t)n)))) return rc;
  VboJob *h = reinterpret_cast<VboJob *>(ctx->h_table);
  const FrameLayout &L = ctx->lay;
  for (int i = 0; i < n; i++) {
    h[i].idepth = reinterpret_cast<const float *>(frames[i]->slab + L.idepth[level]);
    h[i].var = reinterpret_cast<const float *>(frames[i]->slab + L.idvar[level]);
    h[i].img = reinterpret_cast<const float *>(frames[i]->slab + L.img[level]);
    h[i].out = d_vtx + (size_t)i * N;
    h[i].state = d_state + (size_t)i * (nChunks + 1);
    h[i].points = d_points + i;
    h[i].scale = camToWorldScale[i];
  }
  VboK P;
  P.W = W; P.H = H; P.nChunks = nChunks;
  {  // Keyframe.h:87-90, from the float intrinsics the Keyframe message carries
    const float fx = ctx->K.fx[level], fy = ctx->K.fy[level], cx = ctx->K.cx[level], cy = ctx->K.cy[level];
    P.fxi = 1 / fx; P.fyi = 1 / fy;
    P.cxi = -cx / fx; P.cyi = -cy / fy;
  }
  P.scaledTH = prm.scaledTH; P.absTH = prm.absTH;
  P.minNearSupport = prm.minNearSupport; P.contractFma = prm.contractFma;
  cudaStream_t st = ctx->stream;
  LSD_CUDA(cudaMemcpyAsync(ctx->d_table, h, sizeof(VboJob) * (size_t)n, cudaMemcpyHostToDevice, st));
  LSD_CUDA(cudaMemsetAsync(d_state, 0, stateBytes + cntBytes, st));
  LSD_CUDA(cudaEventRecord(ctx->evA, st));
  k_vbo_extract<<<dim3(nChunks, n), VBO_THREADS, 0, st>>>(reinterpret_cast<const VboJob *>(ctx->d_table), P);
  ctx->launches++;
  LSD_CUDA(cudaEventRecord(ctx->evB, st));
  ctx->stageTimed = true;
  LSD_CUDA(cudaGetLastError());
  LSD_CUDA(cudaMemcpyAsync(ctx->h_stage, d_points, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  std::memcpy(points, ctx->h_stage, sizeof(int) * (size_t)n);
  if (dst) {
    for (int i = 0; i < n; i++)
      if (dst[i] && points[i] > 0)
        LSD_CUDA(cudaMemcpyAsync(dst[i], d_vtx + (size_t)i * N, sizeof(lsd_vertex) * (size_t)points[i], cudaMemcpyDeviceToHost, st));
    LSD_CUDA(cudaStreamSynchronize(st));
  }
  return LSD_OK;
}

int lsd_keyframe_compute_vbo(lsd_ctx *ctx, lsd_frame *f, int level, float camToWorldScale, const lsd_vbo_params *params, lsd_vertex *dst,
                             int *points) {
  LSD_ARG(dst);
  return lsd_keyframe_compute_vbo_batch(ctx, 1, &f, level, &camToWorldScale, params, nullptr, &dst, points);
}

}  // extern "C"
