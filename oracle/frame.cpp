// oracle/frame.cpp -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lsd_oracle.hpp).
// Restates upstream DataStructures/Frame.cpp (lsd-slam core, un-vendored) per SURVEY.md A.1.
// The consumer-side contract these arrays must honour is evidenced in the reference at
// lib/Pangolin_IOWrapper/PangolinOutputIOWrapper.cpp:56-79 (width/height/fx.. (level),
// image(level), idepth(level), idepthVar(level)).
#include <cmath>
#include <cstring>

#include "lsd_oracle.hpp"

namespace lsdo {

bool g_exactSums = false;

// Frame::Frame / Frame::initialize.  Input contract: 8-bit grey image
// (/root/reference/lib/App/InputThread.cpp:59,65,71).
Frame::Frame(int id_, int width, int height, float fx0, float fy0, float cx0, float cy0, const uint8_t *img) : id(id_) {
  for (int l = 0; l < NL; l++) {
    w[l] = width >> l;
    h[l] = height >> l;
    if (l == 0) {
      fx[0] = fx0; fy[0] = fy0; cx[0] = cx0; cy[0] = cy0;
    } else {
      fx[l] = fx[l - 1] * 0.5;
      fy[l] = fy[l - 1] * 0.5;
      cx[l] = (cx[0] + 0.5) / ((int)1 << l) - 0.5;  // double arithmetic, rounded on store (upstream literals)
      cy[l] = (cy[0] + 0.5) / ((int)1 << l) - 0.5;
    }
    // DECISION: K^-1 entries in closed form (upstream takes Eigen's 3x3 inverse of K_l; equal up to 1 ulp).
    fxi[l] = 1.0f / fx[l];
    fyi[l] = 1.0f / fy[l];
    cxi[l] = -cx[l] / fx[l];
    cyi[l] = -cy[l] / fy[l];
  }
  image[0].resize((size_t)width * height);
  for (size_t i = 0; i < image[0].size(); i++) image[0][i] = (float)img[i];
  imageValid[0] = true;
}

// Frame::buildImage: 2x2 box mean.  Exact in fp32 for u8 input through level 4.
void Frame::buildImage(int level) {
  if (level == 0) return;
  requireImage(level - 1);
  const int sw = w[level - 1];
  const int dw = w[level], dh = h[level];
  const float *src = image[level - 1].data();
  image[level].resize((size_t)dw * dh);
  float *dst = image[level].data();
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      const float *s = src + 2 * x + 2 * y * sw;
      dst[x + y * dw] = (s[0] + s[1] + s[sw] + s[1 + sw]) * 0.25f;
    }
  imageValid[level] = true;
}

// Frame::buildGradients: central differences over the LINEAR index range [w, w(h-1)); rows 0
// and h-1 stay untouched (DECISION: zero), x=0 / x=w-1 use the wrapped linear neighbours.
void Frame::buildGradients(int level) {
  requireImage(level);
  const int W = w[level], H = h[level];
  grad[level].assign((size_t)4 * W * H, 0.0f);
  const float *I = image[level].data();
  float *g = grad[level].data();
  for (int i = W; i < W * (H - 1); i++) {
    g[4 * i + 0] = 0.5f * (I[i + 1] - I[i - 1]);
    g[4 * i + 1] = 0.5f * (I[i + W] - I[i - W]);
    g[4 * i + 2] = I[i];
  }
  gradValid[level] = true;
}

// Frame::buildMaxGradients: |g|, then vertical 3-max into a temp, then horizontal 3-max back.
// Unwritten entries are defined as 0 (upstream reads recycled allocator memory there).
void Frame::buildMaxGradients(int level) {
  requireGradients(level);
  const int W = w[level], H = h[level];
  maxGrad[level].assign((size_t)W * H, 0.0f);
  std::vector<float> tmp((size_t)W * H, 0.0f);
  float *m = maxGrad[level].data();
  const float *g = grad[level].data();
  for (int i = W; i < W * (H - 1); i++) {
    const float dx = g[4 * i], dy = g[4 * i + 1];
    m[i] = sqrtf(dx * dx + dy * dy);
  }
  for (int i = W + 1; i < W * (H - 1) - 1; i++) {
    float g1 = m[i - W];
    const float g2 = m[i];
    if (g1 < g2) g1 = g2;
    const float g3 = m[i + W];
    tmp[i] = (g1 < g3) ? g3 : g1;
  }
  int mappable = 0;
  for (int i = W + 1; i < W * (H - 1) - 1; i++) {
    float g1 = tmp[i - 1];
    const float g2 = tmp[i];
    if (g1 < g2) g1 = g2;
    const float g3 = tmp[i + 1];
    const float r = (g1 < g3) ? g3 : g1;
    m[i] = r;
    if (r >= MIN_USE_GRAD) mappable++;
  }
  if (level == 0) numMappablePixels = mappable;
  maxGradValid[level] = true;
}

// Frame::buildIDepthAndIDepthVar: inverse-variance weighted 2x2 fusion of valid children.
void Frame::buildIDepthAndIDepthVar(int level) {
  if (level == 0) return;
  requireIDepth(level - 1);
  const int sw = w[level - 1];
  const int W = w[level], H = h[level];
  const float *ids = idepth[level - 1].data();
  const float *vs = idepthVar[level - 1].data();
  idepth[level].resize((size_t)W * H);
  idepthVar[level].resize((size_t)W * H);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int idx = 2 * (x + y * sw);
      float idepthSumsSum = 0, ivarSumsSum = 0;
      int num = 0;
      const int offs[4] = {0, 1, sw, sw + 1};
      for (int k = 0; k < 4; k++) {
        const float var = vs[idx + offs[k]];
        if (var > 0) {
          const float ivar = 1.0f / var;
          ivarSumsSum += ivar;
          idepthSumsSum += ivar * ids[idx + offs[k]];
          num++;
        }
      }
      if (num > 0) {
        const float depth = ivarSumsSum / idepthSumsSum;
        idepth[level][x + y * W] = 1.0f / depth;
        idepthVar[level][x + y * W] = num / ivarSumsSum;
      } else {
        idepth[level][x + y * W] = -1;
        idepthVar[level][x + y * W] = -1;
      }
    }
  idepthValid[level] = true;
}

// Frame::setDepth(const DepthMapPixelHypothesis*)
void Frame::setDepth(const Hypothesis *map) {
  const int N = w[0] * h[0];
  idepth[0].resize(N);
  idepthVar[0].resize(N);
  float numIdepth = 0, sumIdepth = 0;
  double sumD = 0;
  for (int i = 0; i < N; i++) {
    if (map[i].isValid && map[i].idepth_smoothed >= -0.05f) {
      idepth[0][i] = map[i].idepth_smoothed;
      idepthVar[0][i] = map[i].idepth_var_smoothed;
      numIdepth++;
      sumIdepth += map[i].idepth_smoothed;
      sumD += (double)map[i].idepth_smoothed;
    } else {
      idepth[0][i] = -1;
      idepthVar[0][i] = -1;
    }
  }
  if (g_exactSums) sumIdepth = (float)sumD;
  meanIdepth = sumIdepth / numIdepth;
  numPoints = (int)numIdepth;
  idepthValid[0] = true;
  for (int l = 1; l < NL; l++) idepthValid[l] = false;
  hasIDepthBeenSet = true;
  depthHasBeenUpdatedFlag = true;
}

// Frame::setDepthFromGroundTruth(const float* depth, float cov_scale)
void Frame::setDepthFromGroundTruth(const float *depth, float cov_scale) {
  const int N = w[0] * h[0];
  idepth[0].resize(N);
  idepthVar[0].resize(N);
  for (int i = 0; i < N; i++) {
    if (depth[i] > 0) {
      idepth[0][i] = 1.0f / depth[i];
      idepthVar[0][i] = VAR_GT_INIT_INITIAL * cov_scale;
    } else {
      idepth[0][i] = -1;
      idepthVar[0][i] = -1;
    }
  }
  idepthValid[0] = true;
  for (int l = 1; l < NL; l++) idepthValid[l] = false;
  hasIDepthBeenSet = true;
}

void Frame::setIDepthRaw(const float *id, const float *var) {
  const int N = w[0] * h[0];
  idepth[0].assign(id, id + N);
  idepthVar[0].assign(var, var + N);
  idepthValid[0] = true;
  for (int l = 1; l < NL; l++) idepthValid[l] = false;
  hasIDepthBeenSet = true;
}

// Frame::refPixelWasGood(): (w>>1)(h>>1) bools, memset 0xFF on creation.
uint8_t *Frame::refPixelWasGoodBuf() {
  if (refPixelWasGood.empty())
    refPixelWasGood.assign((size_t)w[SE3TRACKING_MIN_LEVEL] * h[SE3TRACKING_MIN_LEVEL], 0xFF);
  return refPixelWasGood.data();
}

// Frame::prepareForStereoWith(other, thisToOther, K, level)
void Frame::prepareForStereoWith(const Frame *other, const Sim3<double> &thisToOther, int level) {
  const Sim3<float> otherToThis = thisToOther.inverse().cast<float>();
  Mat3<float> K = Mat3<float>::zero();
  K.m[0][0] = fx[level]; K.m[1][1] = fy[level]; K.m[0][2] = cx[level]; K.m[1][2] = cy[level]; K.m[2][2] = 1;
  // K * R * s  -- upstream: K * otherToThis.rotationMatrix().cast<float>() * otherToThis.scale()
  Mat3<double> Rd = thisToOther.inverse().rotationMatrix();
  Mat3<float> Rf;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rf.m[i][j] = (float)Rd.m[i][j];
  K_otherToThis_R = (K * Rf) * otherToThis.s;
  otherToThis_t = otherToThis.t;
  K_otherToThis_t = K * otherToThis_t;

  const Sim3<float> t2o = thisToOther.cast<float>();
  thisToOther_t = t2o.t;
  K_thisToOther_t = K * thisToOther_t;
  Mat3<double> R2d = thisToOther.rotationMatrix();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) thisToOther_R.m[i][j] = (float)R2d.m[i][j] * t2o.s;
  otherToThis_R_row0 = Vec3<float>(thisToOther_R.m[0][0], thisToOther_R.m[1][0], thisToOther_R.m[2][0]);
  otherToThis_R_row1 = Vec3<float>(thisToOther_R.m[0][1], thisToOther_R.m[1][1], thisToOther_R.m[2][1]);
  otherToThis_R_row2 = Vec3<float>(thisToOther_R.m[0][2], thisToOther_R.m[1][2], thisToOther_R.m[2][2]);
  distSquared = otherToThis.t.dot(otherToThis.t);
  referenceID = other->id;
  referenceLevel = level;
}

}  // namespace lsdo
