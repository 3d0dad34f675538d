// oracle/undistort.cpp -- TEST INFRASTRUCTURE ONLY.  PINNED ON THE DEPENDENCY'S OWN OUTPUT (see below).
//
// CPU restatement of the step immediately BEFORE the hot path:
//   undistorter->undistort(image, imageUndist)            /root/reference/lib/App/InputThread.cpp:61-65
// The undistorter is libvideoio::Undistorter (libvideoio/Undistorter.h, un-vendored fips import, no version pinned:
// /root/reference/fips.yml), created by UndistorterFactory::getUndistorterFromFile (tools/LSD.cpp:88) from a calibration
// such as /root/reference/d2_camera.xml.  Its arithmetic is OpenCV's: cv::initUndistortRectifyMap(K, dist, I, K_out,
// size, CV_16SC2, map1, map2) once, then cv::remap(src, dst, map1, map2, INTER_LINEAR) per frame (the lsd_slam
// "UndistorterOpenCV" lineage).  OpenCV's C++ headers are absent here, but its Python build (cv2 4.13.0) is installed:
// scripts/make_golden_undistort.py records cv2's maps and remapped images into tests/golden/undistort.npz and
// tests/test_oracle_undistort.py requires this restatement to reproduce them bit for bit (live against cv2 as well
// when it can be imported).
//
// Published algorithm restated (OpenCV imgproc, undistort.dispatch.cpp / imgwarp.cpp):
//   maps : iR = (K_out * R)^-1 by the 3x3 adjugate (R = I); per row _x,_y,_w start at (i*iR[1]+iR[2], ...) and advance by
//          (iR[0], iR[3], iR[6]) per column; x = _x/_w, y = _y/_w; kr = 1 + ((k3 r2 + k2) r2 + k1) r2;
//          xd = x kr + p1 2xy + p2 (r2 + 2x^2), yd = y kr + p1 (r2 + 2y^2) + p2 2xy; u = fx xd + cx, v = fy yd + cy;
//          fixed point: iu = cvRound(u*32), iv = cvRound(v*32); map1 = (iu>>5, iv>>5) int16, map2 = (iv&31)*32 + (iu&31).
//   remap: 8-bit bilinear with 15-bit fixed-point weights; on the 32x32 sub-pixel grid these are exact integers
//          (32-fx)(32-fy)*32 ..., so dst = (sum w_k p_k + 2^14) >> 15.  BORDER_CONSTANT, value 0: taps outside the
//          source contribute 0; a footprint entirely outside gives 0.
#include <cmath>
#include <cstdint>
#include <cstring>

extern "C" {

// K = {fx, fy, cx, cy}; dist = {k1, k2, p1, p2, k3}; map1: 2*w*h int16 (x, y interleaved); map2: w*h uint16
void lsdo_init_undistort_rectify_map(const double K[4], const double dist[5], const double Kout[4], int w, int h, int16_t *map1,
                                     uint16_t *map2) {
  const double fx = K[0], fy = K[1], u0 = K[2], v0 = K[3];
  const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
  // inverse of A = [a 0 c; 0 b d; 0 0 1] by the adjugate / determinant (cv::Matx33d::inv)
  const double A[9] = {Kout[0], 0, Kout[2], 0, Kout[1], Kout[3], 0, 0, 1};
  const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
  const double d = 1.0 / det;
  double ir[9];
  ir[0] = (A[4] * A[8] - A[5] * A[7]) * d;
  ir[1] = (A[2] * A[7] - A[1] * A[8]) * d;
  ir[2] = (A[1] * A[5] - A[2] * A[4]) * d;
  ir[3] = (A[5] * A[6] - A[3] * A[8]) * d;
  ir[4] = (A[0] * A[8] - A[2] * A[6]) * d;
  ir[5] = (A[2] * A[3] - A[0] * A[5]) * d;
  ir[6] = (A[3] * A[7] - A[4] * A[6]) * d;
  ir[7] = (A[1] * A[6] - A[0] * A[7]) * d;
  ir[8] = (A[0] * A[4] - A[1] * A[3]) * d;
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
      double su = u * 32, sv = v * 32;  // saturate_cast<int>(double) = cvRound with saturation
      if (su > 2147483647.0) su = 2147483647.0;
      if (su < -2147483648.0) su = -2147483648.0;
      if (sv > 2147483647.0) sv = 2147483647.0;
      if (sv < -2147483648.0) sv = -2147483648.0;
      const int iu = (int)std::lrint(su), iv = (int)std::lrint(sv);
      map1[2 * ((size_t)i * w + j)] = (int16_t)(iu >> 5);
      map1[2 * ((size_t)i * w + j) + 1] = (int16_t)(iv >> 5);
      map2[(size_t)i * w + j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    }
  }
}

// cv::remap(src, dst, map1 (CV_16SC2), map2 (CV_16UC1), INTER_LINEAR, BORDER_CONSTANT, 0) for CV_8UC1
void lsdo_remap_u8(const uint8_t *src, int sw, int sh, size_t spitch, const int16_t *map1, const uint16_t *map2, int w, int h, uint8_t *dst) {
  for (int i = 0; i < h; i++)
    for (int j = 0; j < w; j++) {
      const size_t o = (size_t)i * w + j;
      const int sx = map1[2 * o], sy = map1[2 * o + 1];
      const int f = map2[o] & 1023, fxq = f & 31, fyq = f >> 5;
      const int w00 = (32 - fxq) * (32 - fyq) * 32, w01 = fxq * (32 - fyq) * 32, w10 = (32 - fxq) * fyq * 32, w11 = fxq * fyq * 32;
      auto px = [&](int x, int y) -> int { return (x >= 0 && x < sw && y >= 0 && y < sh) ? src[(size_t)y * spitch + x] : 0; };
      const int acc = w00 * px(sx, sy) + w01 * px(sx + 1, sy) + w10 * px(sx, sy + 1) + w11 * px(sx + 1, sy + 1);
      dst[o] = (uint8_t)((acc + (1 << 14)) >> 15);
    }
}

}  // extern "C"
