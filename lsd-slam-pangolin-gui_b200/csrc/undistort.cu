// undistort.cu -- image undistortion on device, fused in front of the frame ingest.
//
// Replaces the step immediately before the hot path: `undistorter->undistort(image, imageUndist)`
// (/root/reference/lib/App/InputThread.cpp:61-65; the undistorter is libvideoio::Undistorter, created at
// tools/LSD.cpp:88 from a calibration file such as d2_camera.xml).  libvideoio is an un-vendored import; its arithmetic
// is OpenCV's -- cv::initUndistortRectifyMap(..., CV_16SC2) once and cv::remap(..., INTER_LINEAR) per frame -- and that
// is what is reproduced here, bit for bit:
//   * maps: either handed over by the caller (the int16 x/y + uint16 sub-pixel-index maps an OpenCV undistorter
//     holds), or built on the host in fp64 with OpenCV's operation order (a one-off per run);
//   * k_remap_u8: one thread per output pixel; map1 (4 B) + map2 (2 B) streamed, four 8-bit taps gathered from the
//     distorted frame (smooth maps: a warp's taps fall in 1-3 lines), 15-bit fixed-point weights that are exact
//     integers on the 32x32 sub-pixel grid, (sum + 2^14) >> 15; BORDER_CONSTANT 0.  blockIdx.y = image.
// The undistorted 8-bit image stays on the device and goes straight into k_ingest (pyramids); the host copy the GUI
// shows (output->updateLiveImage(imageUndist), InputThread.cpp:78) is optional.
#include <cmath>
#include <cstring>

#include "ctx.cuh"

struct lsd_undistorter {
  int inW, inH, outW, outH;
  short2 *d_map1;
  uint16_t *d_map2;
  uint8_t *d_src, *d_dst;  // staging for n images
  uint8_t *h_src;          // pinned
  int cap;                 // images the staging holds
};

namespace lsd {

__global__ void __launch_bounds__(256) k_remap_u8(const uint8_t *__restrict__ src, int sw, int sh, size_t srcStride,
                                                  const short2 *__restrict__ map1, const uint16_t *__restrict__ map2, int N,
                                                  uint8_t *__restrict__ dst) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= N) return;
  const uint8_t *s = src + (size_t)blockIdx.y * srcStride;
  const short2 m = __ldg(map1 + o);
  const int f = __ldg(map2 + o) & 1023;
  const int fx = f & 31, fy = f >> 5;
  const int sx = m.x, sy = m.y;
  const bool x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw;
  const bool y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
  const uint8_t *p = s + (ptrdiff_t)sy * sw + sx;
  const int p00 = (x0 && y0) ? __ldg(p) : 0, p01 = (x1 && y0) ? __ldg(p + 1) : 0;
  const int p10 = (x0 && y1) ? __ldg(p + sw) : 0, p11 = (x1 && y1) ? __ldg(p + sw + 1) : 0;
  const int acc = (32 - fx) * (32 - fy) * 32 * p00 + fx * (32 - fy) * 32 * p01 + (32 - fx) * fy * 32 * p10 + fx * fy * 32 * p11;
  dst[(size_t)blockIdx.y * N + o] = (uint8_t)((acc + (1 << 14)) >> 15);
}

// cv::initUndistortRectifyMap(K, dist, I, Kout, (w, h), CV_16SC2): fp64, OpenCV's operation order
static void init_maps_host(const double K[4], const double dist[5], const double Kout[4], int w, int h, short2 *map1, uint16_t *map2) {
  const double fx = K[0], fy = K[1], u0 = K[2], v0 = K[3];
  const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
  const double A[9] = {Kout[0], 0, Kout[2], 0, Kout[1], Kout[3], 0, 0, 1};  // (Kout * R)^-1 via the adjugate
  const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
  const double d = 1.0 / det;
  const double ir[9] = {(A[4] * A[8] - A[5] * A[7]) * d, (A[2] * A[7] - A[1] * A[8]) * d, (A[1] * A[5] - A[2] * A[4]) * d,
                        (A[5] * A[6] - A[3] * A[8]) * d, (A[0] * A[8] - A[2] * A[6]) * d, (A[2] * A[3] - A[0] * A[5]) * d,
                        (A[3] * A[7] - A[4] * A[6]) * d, (A[1] * A[6] - A[0] * A[7]) * d, (A[0] * A[4] - A[1] * A[3]) * d};
  auto sat = [](double v) { return v > 2147483647.0 ? 2147483647.0 : (v < -2147483648.0 ? -2147483648.0 : v); };
  for (int i = 0; i < h; i++) {
    double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
    for (int j = 0; j < w; j++, _x += ir[0], _y += ir[3], _w += ir[6]) {
      const double ww = 1. / _w, x = _x * ww, y = _y * ww;
      const double x2 = x * x, y2 = y * y;
      const double r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
      const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
      const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
      const double u = fx * xd + u0, v = fy * yd + v0;
      const int iu = (int)std::lrint(sat(u * 32)), iv = (int)std::lrint(sat(v * 32));  // cvRound
      const size_t o = (size_t)i * w + j;
      map1[o].x = (short)(iu >> 5);
      map1[o].y = (short)(iv >> 5);
      map2[o] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    }
  }
}

static int und_reserve(lsd_undistorter *u, int n) {
  if (n <= u->cap) return LSD_OK;
  if (u->d_src) cudaFree(u->d_src);
  if (u->d_dst) cudaFree(u->d_dst);
  if (u->h_src) cudaFreeHost(u->h_src);
  u->d_src = u->d_dst = u->h_src = nullptr;
  u->cap = 0;
  const size_t inB = (size_t)u->inW * u->inH, outB = (size_t)u->outW * u->outH;
  LSD_CUDA(cudaMalloc(&u->d_src, inB * n));
  LSD_CUDA(cudaMalloc(&u->d_dst, outB * n));
  LSD_CUDA(cudaMallocHost(&u->h_src, inB * n));
  u->cap = n;
  return LSD_OK;
}

// H2D of n distorted images + one remap launch; the undistorted images are left in u->d_dst (n * outW * outH bytes)
static int und_run(lsd_ctx *ctx, lsd_undistorter *u, int n, const uint8_t *const *images, size_t pitch) {
  LSD_ARG(pitch >= (size_t)u->inW);
  int rc = und_reserve(u, n);
  if (rc) return rc;
  const size_t inB = (size_t)u->inW * u->inH;
  for (int i = 0; i < n; i++) {
    LSD_ARG(images[i]);
    uint8_t *dst = u->h_src + inB * i;
    if (pitch == (size_t)u->inW) std::memcpy(dst, images[i], inB);
    else for (int y = 0; y < u->inH; y++) std::memcpy(dst + (size_t)y * u->inW, images[i] + (size_t)y * pitch, u->inW);
  }
  cudaStream_t st = ctx->stream;
  LSD_CUDA(cudaMemcpyAsync(u->d_src, u->h_src, inB * n, cudaMemcpyHostToDevice, st));
  const int N = u->outW * u->outH;
  LSD_CUDA(cudaEventRecord(ctx->evA, st));
  k_remap_u8<<<dim3((N + 255) / 256, n), 256, 0, st>>>(u->d_src, u->inW, u->inH, inB, u->d_map1, u->d_map2, N, u->d_dst);
  ctx->launches++;
  LSD_CUDA(cudaEventRecord(ctx->evB, st));
  ctx->stageTimed = true;
  LSD_CUDA(cudaGetLastError());
  return LSD_OK;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

int lsd_undistorter_destroy(lsd_ctx *ctx, lsd_undistorter *u);

int lsd_undistorter_create_from_maps(lsd_ctx *ctx, int inWidth, int inHeight, const int16_t *map1, const uint16_t *map2,
                                     lsd_undistorter **out) {
  LSD_ARG(ctx && map1 && map2 && out && inWidth > 0 && inHeight > 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  lsd_undistorter *u = new lsd_undistorter();
  std::memset(u, 0, sizeof(*u));
  u->inW = inWidth; u->inH = inHeight;
  u->outW = ctx->w; u->outH = ctx->h;
  const size_t N = (size_t)u->outW * u->outH;
  auto upload = [&]() -> int {
    LSD_CUDA(cudaMalloc(&u->d_map1, N * sizeof(short2)));
    LSD_CUDA(cudaMalloc(&u->d_map2, N * sizeof(uint16_t)));
    LSD_CUDA(cudaMemcpyAsync(u->d_map1, map1, N * sizeof(short2), cudaMemcpyHostToDevice, ctx->stream));
    LSD_CUDA(cudaMemcpyAsync(u->d_map2, map2, N * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    LSD_CUDA(cudaStreamSynchronize(ctx->stream));
    return LSD_OK;
  };
  const int rc = upload();
  if (rc) {
    lsd_undistorter_destroy(ctx, u);
    return rc;
  }
  *out = u;
  return LSD_OK;
}

int lsd_undistorter_create_opencv(lsd_ctx *ctx, int inWidth, int inHeight, const double K[4], const double dist[5], const double Kout[4],
                                  lsd_undistorter **out) {
  LSD_ARG(ctx && K && dist && Kout && out);
  const size_t N = (size_t)ctx->w * ctx->h;
  std::vector<short2> m1(N);
  std::vector<uint16_t> m2(N);
  init_maps_host(K, dist, Kout, ctx->w, ctx->h, m1.data(), m2.data());
  return lsd_undistorter_create_from_maps(ctx, inWidth, inHeight, reinterpret_cast<const int16_t *>(m1.data()), m2.data(), out);
}

int lsd_undistorter_destroy(lsd_ctx *ctx, lsd_undistorter *u) {
  LSD_ARG(ctx);
  if (!u) return LSD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (u->d_map1) cudaFree(u->d_map1);
  if (u->d_map2) cudaFree(u->d_map2);
  if (u->d_src) cudaFree(u->d_src);
  if (u->d_dst) cudaFree(u->d_dst);
  if (u->h_src) cudaFreeHost(u->h_src);
  delete u;
  return LSD_OK;
}

int lsd_undistorter_maps(lsd_ctx *ctx, lsd_undistorter *u, int16_t *map1, uint16_t *map2) {
  LSD_ARG(ctx && u && map1 && map2);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t N = (size_t)u->outW * u->outH;
  LSD_CUDA(cudaMemcpyAsync(map1, u->d_map1, N * sizeof(short2), cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaMemcpyAsync(map2, u->d_map2, N * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}

int lsd_undistort(lsd_ctx *ctx, lsd_undistorter *u, const uint8_t *image, size_t pitch, uint8_t *undistorted) {
  LSD_ARG(ctx && u && image && undistorted);
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = und_run(ctx, u, 1, &image, pitch);
  if (rc) return rc;
  LSD_CUDA(cudaMemcpyAsync(undistorted, u->d_dst, (size_t)u->outW * u->outH, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}

int lsd_frame_create_undistorted_batch(lsd_ctx *ctx, lsd_undistorter *u, int n, const int *ids, const uint8_t *const *images, size_t pitch,
                                       unsigned flags, uint8_t *const *undistorted, lsd_frame **out) {
  LSD_ARG(ctx && u && images && out && n >= 1);
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = und_run(ctx, u, n, images, pitch);
  if (rc) return rc;
  const size_t outB = (size_t)u->outW * u->outH;
  if (undistorted)
    for (int i = 0; i < n; i++)
      if (undistorted[i]) LSD_CUDA(cudaMemcpyAsync(undistorted[i], u->d_dst + outB * i, outB, cudaMemcpyDeviceToHost, ctx->stream));
  return lsd_frame_create_batch_device(ctx, n, ids, u->d_dst, flags, out);  // synchronises
}

int lsd_frame_create_undistorted(lsd_ctx *ctx, lsd_undistorter *u, int id, const uint8_t *image, size_t pitch, unsigned flags,
                                 uint8_t *undistorted, lsd_frame **out) {
  return lsd_frame_create_undistorted_batch(ctx, u, 1, &id, &image, pitch, flags, undistorted ? &undistorted : nullptr, out);
}

}  // extern "C"
