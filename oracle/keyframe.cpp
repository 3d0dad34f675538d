// oracle/keyframe.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY PINNED (this file only; see below).
//
// CPU restatement of the keyframe-publish step that follows the hot path on every keyframe:
//   (1) PangolinOutputIOWrapper::publishKeyframe's pack loop
//       (/root/reference/lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:69-89): idepth / idepthVar / image of
//       the publish level (publishLvl = 0, PangolinOutputIOWrapper.cpp:21) -> 12-byte InputPointDense records
//       (Keyframe.h:16-21), the grey value truncated to unsigned char and replicated four times;
//   (2) Keyframe::computeVbo (/root/reference/lib/Pangolin_IOWrapper/Keyframe.h:66-158): per interior pixel, in
//       raster order, keep the point when idepth > 0, var*depth^4 <= 1e-3, var*depth^4*scale^2 <= 1e-1 and all nine
//       3x3 neighbours support it (|idepth_n - 1/depth|^2 < 2*var); emit MyVertex{(x*fxi+cxi)*depth,
//       (y*fyi+cyi)*depth, depth, colour(b,g,r,100)} (Keyframe.h:47-51,136-142).
//
// Unlike the rest of oracle/, THIS restatement is pinned: both functions live in /root/reference, and
// oracle/_ref/libref_keyframe.so is the reference's own Keyframe.h compiled from where it lies (oracle/Makefile,
// oracle/ref_keyframe.cpp).  tests/test_oracle_keyframe.py requires restatement == reference bit for bit whenever
// /root/reference (or the prebuilt _ref library) is present, and against the committed vectors in tests/golden/.
//
// Floating point: the reference's Release flags include -march=native (CMakeLists.txt:59), under which gcc may
// contract x*fxi+cxi into an FMA.  `contractFma` selects that variant; 0 is the IEEE (uncontracted) evaluation a
// plain `g++ -O2` build of the reference produces, which is what oracle/_ref is.
#include <cmath>
#include <cstdint>
#include <cstring>

extern "C" {

struct lsdo_input_point_dense {  // Keyframe.h:16-21
  float idepth, idepth_var;
  unsigned char color[4];
};
struct lsdo_vertex {  // Keyframe.h:47-51
  float point[3];
  unsigned char color[4];
};

// (1) publishKeyframe pack.  hasIDepth == 0 reproduces the reference's "frame has no depth" branch: the
// buffer it hands the GUI is then left as allocated (PangolinOutputIOWrapper.cpp:89-91); here it is zero-filled.
void lsdo_publish_keyframe_pack(const float *idepth, const float *idepthVar, const float *image, int n, int hasIDepth,
                                lsdo_input_point_dense *out) {
  if (!hasIDepth) {
    std::memset(out, 0, sizeof(lsdo_input_point_dense) * (size_t)n);
    return;
  }
  for (int i = 0; i < n; i++) {
    lsdo_input_point_dense p;
    p.idepth = idepth[i];
    p.idepth_var = idepthVar[i];
    const unsigned char g = (unsigned char)image[i];  // float -> unsigned char: truncation
    p.color[0] = p.color[1] = p.color[2] = p.color[3] = g;
    out[i] = p;
  }
}

// (2) computeVbo.  Thresholds are the reference's constants; sparsifyFactor is fixed at 1 there (rand() never drawn).
int lsdo_compute_vbo(const lsdo_input_point_dense *in, int width, int height, float fx, float fy, float cx, float cy,
                     float camToWorldScale, float scaledTH, float absTH, int minNearSupport, int contractFma, lsdo_vertex *out) {
  const float fxi = 1 / fx, fyi = 1 / fy;
  const float cxi = -cx / fx, cyi = -cy / fy;
  int points = 0;
  for (int y = 1; y < height - 1; y++) {
    const lsdo_input_point_dense *row = in + (size_t)y * width;
    for (int x = 1; x < width - 1; x++) {
      const lsdo_input_point_dense &c = row[x];
      if (c.idepth <= 0) continue;
      const float depth = 1 / c.idepth;
      float depth4 = depth * depth;
      depth4 *= depth4;
      if (c.idepth_var * depth4 > scaledTH) continue;
      if (c.idepth_var * depth4 * camToWorldScale * camToWorldScale > absTH) continue;
      if (minNearSupport > 1) {
        int support = 0;
        for (int dx = -1; dx <= 1; dx++)
          for (int dy = -1; dy <= 1; dy++) {
            const float nid = in[(size_t)(y + dy) * width + (x + dx)].idepth;
            if (nid > 0) {
              const float diff = nid - 1.0f / depth;
              if (diff * diff < 2 * c.idepth_var) support++;
            }
          }
        if (support < minNearSupport) continue;
      }
      lsdo_vertex &v = out[points++];
      if (contractFma) {
        v.point[0] = std::fmaf((float)x, fxi, cxi) * depth;
        v.point[1] = std::fmaf((float)y, fyi, cyi) * depth;
      } else {
        v.point[0] = (x * fxi + cxi) * depth;
        v.point[1] = (y * fyi + cyi) * depth;
      }
      v.point[2] = depth;
      v.color[3] = 100;
      v.color[2] = c.color[0];
      v.color[1] = c.color[1];
      v.color[0] = c.color[2];
    }
  }
  return points;
}

}  // extern "C"
