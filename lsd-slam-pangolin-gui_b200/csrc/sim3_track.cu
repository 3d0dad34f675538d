// sim3_track.cu -- [UP] Sim3Tracker::trackFrameSim3: 7-DoF
// photometric + depth alignment between keyframes, the unit of work of the constraint search
// (BASELINE.json configs[3]: a new keyframe against 64 candidates, sharded over GPUs with no collective).
//
// Reference structure (lsd-slam core Tracking/Sim3Tracker.cpp, un-vendored): per LM evaluation
// calcSim3Buffers (12 SoA buffers) -> calcSim3WeightsAndResidual -> calcSim3LGS (LGS6 + LGS4 -> LGS7), with
// the 7x7 LDLT and the fp64 Sim3 exponential on the host in between.
//
// B200 structure: ONE THREAD-BLOCK CLUSTER (S3_CL = 4 CTAs x 256 threads) PER TRACK, the whole coarse-to-fine LM
// loop on device.  Every evaluation is cut into S3_CL contiguous parts, one per CTA of the cluster; each CTA
// fuses the three reference passes per point (no buffers are materialised), reduces its 41 normal-equation
// terms + residual sums in a fixed order, and CTA 0 gathers the S3_CL partials over distributed shared memory
// in rank order, runs accept / reject, the 7x7 LDL^T solve and the fp64 Sim3 exponential, and broadcasts the
// next pose into every CTA's shared memory.  Two cluster barriers per evaluation, no global-memory round
// trip, no host involvement, bit-reproducible (the decomposition never depends on the batch).
// Compiled -fmad=false: per-point values are IEEE-identical to the oracle's; only summation order differs.
#include <cooperative_groups.h>

#include <cmath>
#include <cstring>

#include "ctx.cuh"
#include "lie_dev.cuh"

namespace cg = cooperative_groups;

namespace lsd {

#ifndef S3_THREADS
#define S3_THREADS 256
#endif
// CTAs per cluster = per track.  Measured on B200 (64 candidates x 2 directions, levels 4->1): 8 CTAs 2.96 ms, 4 CTAs
// 2.70 ms, 2 CTAs 3.22 ms, 1 CTA 6.2 ms (profiles/r01j_sim3_cluster_sweep.txt)
#ifndef S3_CL
#define S3_CL 4
#endif
// resident CTAs per SM the register allocation is capped for (221 registers uncapped = ONE 256-thread CTA per SM, i.e.
// 18 clusters on the whole GPU; the cap trades a few spills in the single-thread LM step for 2-4x the clusters in flight)
#ifndef S3_MINB
#define S3_MINB 2
#endif
// fp32 sums
enum { Q_A6 = 0, Q_B6 = 21, Q_A4 = 27, Q_B4 = 37, Q_RD = 41, Q_RP = 42, Q_USAGE = 43, Q_ND = 44, Q_CNT = 45, S3_NF = 46 };
#define S3_ND 5  // fp64 affine-lighting sums (sxx, syy, sx, sy, sw), see se3_track.cu

struct Sim3Job {
  const RefPoint *pts[NL];
  const float2 *rgrad[NL];
  const float4 *fgrad[NL];
  const float *fid[NL], *fvar[NL];
  const int *d_num;
  double init[8];  // referenceToFrame (qx,qy,qz,qw,tx,ty,tz,s)
};

struct Sim3Out {
  double frameToRef[8];
  float H[49];
  float lastResidual, lastDepthResidual, lastPhotometricResidual, pointUsage, affine_a, affine_b;
  int diverged;
  int nRes[NL], nWarp[NL];
  int traceLen;
  int n[NL];
};

struct Sim3Params {
  Intrinsics K;
  lsd_tracker_settings s;
  int startLevel, finalLevel;
};

// what every CTA of the cluster needs for one evaluation
struct S3Cmd {
  float Rs[9], t[3];  // scaled rotation (rxso3) and translation, float
  float roll[4];      // xRoll0, xRoll1, yRoll0, yRoll1
  float a, b;
  int level;
  int op;  // 0 evaluate, 1 finished
};

struct S3Res {
  float sumResD, sumResP;
  int numTermsD, numTermsP;
  float meanD, meanP, mean;
};

struct S3State {
  double q[4], t[3], s;     // current referenceToFrame
  double qt[4], tt[3], st;  // trial
  float a, b;
  int level, phase, iteration, incTry;
  bool upToDate;
  float lambda, absInc;
  S3Res lastErr, finalRes;
  float sums[41];  // A6, b6, A4, b4 of the evaluation the current outer iteration started from (undivided)
  int nc;
};

__device__ void s3_make_cmd(const double q[4], const double t[3], double s, float a, float b, int level, S3Cmd &c) {
  QuatT<double> qq = {q[0], q[1], q[2], q[3]};
  double Rd[9];
  qtoR(qq, Rd);
  float Ru[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    c.Rs[i] = (float)(Rd[i] * s);  // rxso3().matrix().cast<float>()
    Ru[i] = (float)Rd[i];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) c.t[i] = (float)t[i];
  // rotation about the optical axis: shortest rotation taking R*(0,0,-1) back to (0,0,-1), times R
  const float fwd[3] = {0, 0, -1};
  float rf[3];
  mat3vec(Ru, fwd, rf);
  const float na = sqrtf(rf[0] * rf[0] + rf[1] * rf[1] + rf[2] * rf[2]), nb = 1.0f;
  const float v0[3] = {rf[0] / na, rf[1] / na, rf[2] / na}, v1[3] = {fwd[0] / nb, fwd[1] / nb, fwd[2] / nb};
  const float cth = v1[0] * v0[0] + v1[1] * v0[1] + v1[2] * v0[2];
  QuatT<float> qr;
  if (cth < -1.0f + 1e-5f) {
    float ax[3] = {0 * 0 - v0[2] * 0, v0[2] * 1 - v0[0] * 0, v0[0] * 0 - v0[1] * 1};  // v0 x (1,0,0)
    float n = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    if (n < 1e-3f) {
      ax[0] = v0[1] * 0 - v0[2] * 1; ax[1] = v0[2] * 0 - v0[0] * 0; ax[2] = v0[0] * 1 - v0[1] * 0;  // v0 x (0,1,0)
      n = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    }
    qr.w = 0; qr.x = ax[0] / n; qr.y = ax[1] / n; qr.z = ax[2] / n;
  } else {
    const float axis[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
    const float sq = sqrtf((1.0f + cth) * 2.0f);
    const float invs = 1.0f / sq;
    qr.w = sq * 0.5f; qr.x = axis[0] * invs; qr.y = axis[1] * invs; qr.z = axis[2] * invs;
  }
  float Rb[9];
  qtoR(qr, Rb);
  // rollMat = Rb * Ru: rows 0 and 1, columns 0 and 1
  c.roll[0] = Rb[0] * Ru[0] + Rb[1] * Ru[3] + Rb[2] * Ru[6];
  c.roll[1] = Rb[0] * Ru[1] + Rb[1] * Ru[4] + Rb[2] * Ru[7];
  c.roll[2] = Rb[3] * Ru[0] + Rb[4] * Ru[3] + Rb[5] * Ru[6];
  c.roll[3] = Rb[3] * Ru[1] + Rb[4] * Ru[4] + Rb[5] * Ru[7];
  c.a = a;
  c.b = b;
  c.level = level;
  c.op = 0;
}

// Per-thread accumulators.  S3_SMEM_ACC = 1 keeps the 46 fp32 sums of a thread in ITS column of the shared-memory
// reduction buffer (conflict-free: consecutive threads, consecutive banks) instead of in registers: the register budget
// drops by ~50, twice the CTAs are resident, and the sequence of additions per thread -- hence every bit of the result
// -- is unchanged (the block reduction then reads the buffer in place).
#ifndef S3_SMEM_ACC
#define S3_SMEM_ACC 1
#endif
#if S3_SMEM_ACC
struct S3Acc {
  float *col;  // &sm.f[0][threadIdx.x]
  __device__ __forceinline__ float &operator[](int j) const { return col[j * S3_THREADS]; }
};
#else
typedef float *S3Acc;
#endif

// One reference point: its warp and the loads issued for it (stage A), consumed by stage B.  Running stage A of point
// k+1 before stage B of point k (a two-deep software pipeline) was measured SLOWER (2.29 vs 2.13 ms: the 18 extra live
// registers spill), so the two stages run back to back; only the 24-byte point record is fetched one iteration ahead.
struct S3Warp {
  float Wx, Wy, Wz, pz, u_new, v_new;
  float4 p00, p10, p01, p11;
  float var_frameDepth, id_frameDepth;
  bool inside;
};

__device__ __forceinline__ void s3_stage_a(const float4 raw, const S3Cmd &c, const Sim3Params &prm, int W, int H,
                                           const float4 *__restrict__ G, const float *__restrict__ FID, const float *__restrict__ FVAR,
                                           S3Warp &w) {
  const int lvl = c.level;
  const float fx_l = prm.K.fx[lvl], fy_l = prm.K.fy[lvl], cx_l = prm.K.cx[lvl], cy_l = prm.K.cy[lvl];
  const uint32_t xy = __float_as_uint(raw.x);
  const int x = xy & 0xffff, y = xy >> 16;
  const float inv = raw.y;  // RefPoint::invDepth
  const float px = inv * (prm.K.fxi[lvl] * x + prm.K.cxi[lvl]);
  const float py = inv * (prm.K.fyi[lvl] * y + prm.K.cyi[lvl]);
  const float pz = inv * 1.0f;
  w.pz = pz;
  w.Wx = (c.Rs[0] * px + c.Rs[1] * py + c.Rs[2] * pz) + c.t[0];
  w.Wy = (c.Rs[3] * px + c.Rs[4] * py + c.Rs[5] * pz) + c.t[1];
  w.Wz = (c.Rs[6] * px + c.Rs[7] * py + c.Rs[8] * pz) + c.t[2];
  w.u_new = (w.Wx / w.Wz) * fx_l + cx_l;
  w.v_new = (w.Wy / w.Wz) * fy_l + cy_l;
  w.inside = w.u_new > 1 && w.v_new > 1 && w.u_new < W - 2 && w.v_new < H - 2;
  if (!w.inside) return;
  // getInterpolatedElement43 taps + the frame's own (idepth, var) at the nearest pixel, all in ONE memory round trip
  // (ncu source view of the first version: point record, taps, var, idepth were four serial HBM latencies per point)
  const int ix = (int)w.u_new, iy = (int)w.v_new;
  const float4 *bp = G + ix + iy * W;
  w.p00 = __ldg(bp); w.p10 = __ldg(bp + 1); w.p01 = __ldg(bp + W); w.p11 = __ldg(bp + 1 + W);
  const int idx_rounded = (int)(w.u_new + 0.5f) + W * (int)(w.v_new + 0.5f);
  w.var_frameDepth = __ldg(FVAR + idx_rounded);
  w.id_frameDepth = __ldg(FID + idx_rounded);
}

__device__ __forceinline__ void s3_stage_b(const float4 raw, const float2 rg, const S3Warp &w, const S3Cmd &c, const Sim3Params &prm,
                                           S3Acc acc, double dacc[S3_ND]) {
  if (!w.inside) return;
  const int lvl = c.level;
  const float fx_l = prm.K.fx[lvl], fy_l = prm.K.fy[lvl];
  const float Wx = w.Wx, Wy = w.Wy, Wz = w.Wz, pz = w.pz, u_new = w.u_new, v_new = w.v_new;
  const float4 p00 = w.p00, p10 = w.p10, p01 = w.p01, p11 = w.p11;
  const float var_frameDepth = w.var_frameDepth, id_frameDepth = w.id_frameDepth;
  const int ix = (int)u_new, iy = (int)v_new;
  const float dx = u_new - ix, dy = v_new - iy, dxdy = dx * dy;
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  const float gxI = w11 * p11.x + w01 * p01.x + w10 * p10.x + w00 * p00.x;
  const float gyI = w11 * p11.y + w01 * p01.y + w10 * p10.y + w00 * p00.y;
  const float cI = w11 * p11.z + w01 * p01.z + w10 * p10.z + w00 * p00.z;
  // USE_ESM_TRACKING: mean of the frame gradient and the rolled reference gradient
  const float rotatedGradX = c.roll[0] * rg.x + c.roll[1] * rg.y;
  const float rotatedGradY = c.roll[2] * rg.x + c.roll[3] * rg.y;
  const float gx = fx_l * 0.5f * (gxI + rotatedGradX);
  const float gy = fy_l * 0.5f * (gyI + rotatedGradY);

  const float c1 = c.a * raw.z + c.b;
  const float c2 = cI;
  const float rp = c1 - c2;
  const float weight = fabsf(rp) < 2.0f ? 1 : 2.0f / fabsf(rp);
  dacc[0] += (double)(c1 * c1 * weight);
  dacc[1] += (double)(c2 * c2 * weight);
  dacc[2] += (double)(c1 * weight);
  dacc[3] += (double)(c2 * weight);
  dacc[4] += (double)weight;

  // depth residual against the frame's own inverse depth (nearest pixel)
  const float ref_idepth = 1.0f / Wz;
  const float d = 1.0f / pz;
  float rd, svw;
  if (var_frameDepth > 0) {
    rd = ref_idepth - id_frameDepth;
    svw = var_frameDepth;
  } else {
    rd = -1;
    svw = -1;
  }
  acc[Q_CNT] += 1.0f;
  const float depthChange = pz / Wz;
  acc[Q_USAGE] += depthChange < 1 ? depthChange : 1;

  // calcSim3WeightsAndResidual
  const float s = prm.s.var_weight * raw.w;
  const float sv = prm.s.var_weight * svw;
  const float g0 = (c.t[0] * Wz - c.t[2] * Wx) / (Wz * Wz * d);
  const float g1 = (c.t[1] * Wz - c.t[2] * Wy) / (Wz * Wz * d);
  const float g2 = (Wz - c.t[2]) / (Wz * Wz * d);
  const float drpdd = gx * g0 + gy * g1;
  const float w_p = 1.0f / (LSD_CAMERA_PIXEL_NOISE2 + s * drpdd * drpdd);
  const float w_d = 1.0f / (sv + g2 * g2 * s);
  const float weighted_rd = fabsf(rd * sqrtf(w_d));
  const float weighted_rp = fabsf(rp * sqrtf(w_p));
  const float weighted_abs_res = sv > 0 ? weighted_rd + weighted_rp : weighted_rp;
  const float wh = fabsf(weighted_abs_res < prm.s.huber_d ? 1 : prm.s.huber_d / weighted_abs_res);
  float wd = 0;
  if (sv > 0) {
    acc[Q_RD] += wh * w_d * rd * rd;
    acc[Q_ND] += 1.0f;
    wd = wh * w_d;
  }
  acc[Q_RP] += wh * w_p * rp * rp;
  const float wp = wh * w_p;

  // calcSim3LGS
  const float z = 1.0f / Wz;
  const float z_sqr = 1.0f / (Wz * Wz);
  float v[6], v4[4];
  v[0] = z * gx + 0;
  v[1] = 0 + z * gy;
  v[2] = (-Wx * z_sqr) * gx + (-Wy * z_sqr) * gy;
  v[3] = (float)((double)((-Wx * Wy * z_sqr) * gx) + (-(1.0 + (double)(Wy * Wy * z_sqr))) * (double)gy);
  v[4] = (float)((1.0 + (double)(Wx * Wx * z_sqr)) * (double)gx + (double)((Wx * Wy * z_sqr) * gy));
  v[5] = (-Wy * z) * gx + (Wx * z) * gy;
  v4[0] = z_sqr;
  v4[1] = z_sqr * Wy;
  v4[2] = -z_sqr * Wx;
  v4[3] = z;
  int k = 0;
  const float rpw = rp * wp, rdw = rd * wd;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const float wa = v[a] * wp;
#pragma unroll
    for (int cc = a; cc < 6; cc++, k++) acc[Q_A6 + k] = fmaf(wa, v[cc], acc[Q_A6 + k]);
    acc[Q_B6 + a] = fmaf(v[a], rpw, acc[Q_B6 + a]);
  }
  k = 0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const float wa = v4[a] * wd;
#pragma unroll
    for (int cc = a; cc < 4; cc++, k++) acc[Q_A4 + k] = fmaf(wa, v4[cc], acc[Q_A4 + k]);
    acc[Q_B4 + a] = fmaf(v4[a], rdw, acc[Q_B4 + a]);
  }
}

__device__ __forceinline__ void s3_point(const float4 raw, const float2 rg, const S3Cmd &c, const Sim3Params &prm, int W, int H,
                                         const float4 *__restrict__ G, const float *__restrict__ FID, const float *__restrict__ FVAR,
                                         S3Acc acc, double dacc[S3_ND]) {
  S3Warp w;
  s3_stage_a(raw, c, prm, W, H, G, FID, FVAR, w);
  s3_stage_b(raw, rg, w, c, prm, acc, dacc);
}

// S3_DSHUF = 1: the five fp64 affine sums are reduced with warp shuffles + an 8-entry table per sum instead of a
// [5][256] double buffer (10 KB): the dynamic shared memory of a CTA drops to 46 KB so that FOUR CTAs fit on an SM.
#ifndef S3_DSHUF
#define S3_DSHUF 1
#endif
struct S3Smem {
  float f[S3_NF][S3_THREADS];
#if !S3_DSHUF
  double d[S3_ND][S3_THREADS];
#endif
};

__device__ __forceinline__ bool s3_too_few(int size, int lvl, const Sim3Params &prm) {
  return size < 0.5 * LSD_MIN_GOODPERALL_PIXEL_ABSMIN * prm.K.w[lvl] * prm.K.h[lvl] || size < 10;
}

__device__ void s3_identity_out(Sim3Out *O) {
  O->frameToRef[0] = O->frameToRef[1] = O->frameToRef[2] = 0; O->frameToRef[3] = 1;
  O->frameToRef[4] = O->frameToRef[5] = O->frameToRef[6] = 0; O->frameToRef[7] = 1;
}

// Writes the outputs of a finished track.  H = ls7.A (undivided), scattered as NormalEquationsLeastSquares7::initializeFrom.
__device__ void s3_finish(const S3State &S, Sim3Out *O) {
  float A[49];
#pragma unroll
  for (int i = 0; i < 49; i++) A[i] = 0;
  int k = 0;
  for (int a = 0; a < 6; a++)
    for (int c = a; c < 6; c++, k++) A[a * 7 + c] = A[c * 7 + a] = S.sums[Q_A6 + k];
  const int remap[4] = {2, 3, 4, 6};
  k = 0;
  for (int a = 0; a < 4; a++)
    for (int c = a; c < 4; c++, k++) {
      A[remap[a] * 7 + remap[c]] += S.sums[Q_A4 + k];
      if (c != a) A[remap[c] * 7 + remap[a]] += S.sums[Q_A4 + k];
    }
  for (int i = 0; i < 49; i++) O->H[i] = A[i];
  if (S.s <= 0) {
    O->diverged = 1;
    s3_identity_out(O);
    return;
  }
  O->lastResidual = S.finalRes.mean;
  O->lastDepthResidual = S.finalRes.meanD;
  O->lastPhotometricResidual = S.finalRes.meanP;
  // referenceToFrame.inverse()
  QuatT<double> qc = {-S.q[0], -S.q[1], -S.q[2], S.q[3]};
  double R[9], nt[3] = {S.t[0] * -1.0, S.t[1] * -1.0, S.t[2] * -1.0}, rt[3];
  qtoR(qc, R);
  mat3vec(R, nt, rt);
  const double si = 1.0 / S.s;
  O->frameToRef[0] = qc.x; O->frameToRef[1] = qc.y; O->frameToRef[2] = qc.z; O->frameToRef[3] = qc.w;
  O->frameToRef[4] = rt[0] * si; O->frameToRef[5] = rt[1] * si; O->frameToRef[6] = rt[2] * si;
  O->frameToRef[7] = si;
}

// The LM state machine after one evaluation.  Returns false when the track is finished.
__device__ bool s3_step(const Sim3Job *J, S3State &S, Sim3Out *O, const float *tot, const double *dtot, const Sim3Params &prm,
                        lsd_trace_entry *trace, S3Cmd &next) {
  const int lvl = S.level;
  const int size = (int)tot[Q_CNT];
  O->pointUsage = tot[Q_USAGE] / (float)O->n[lvl];
  const double sxx = dtot[0], syy = dtot[1], sx = dtot[2], sy = dtot[3], sw = dtot[4];
  const double aLd = sqrt((syy - sy * sy / sw) / (sxx - sx * sx / sw));
  const float aL = (float)aLd, bL = (float)((sy - aLd * sx) / sw);
  S3Res err;
  err.sumResD = tot[Q_RD];
  err.sumResP = tot[Q_RP];
  err.numTermsD = (int)tot[Q_ND];
  err.numTermsP = size;
  err.mean = (err.sumResD + err.sumResP) / (err.numTermsD + err.numTermsP);
  err.meanD = err.sumResD / err.numTermsD;
  err.meanP = err.sumResP / err.numTermsP;

  if (S.phase == 2) {  // the re-evaluation upstream runs when the last step was accepted (!warp_update_up_to_date)
    S.finalRes = err;
#pragma unroll
    for (int k = 0; k < 41; k++) S.sums[k] = tot[k];
    O->affine_a = S.a;
    O->affine_b = S.b;
    s3_finish(S, O);
    return false;
  }
  if (s3_too_few(size, lvl, prm)) {
    O->diverged = 1;
    s3_identity_out(O);
    return false;
  }
  O->nRes[lvl]++;
  const int maxIts = prm.s.maxItsPerLvl[lvl];
  bool take = false;
  int accepted;
  float traceLambda = S.lambda;
  if (S.phase == 0) {
    S.lastErr = err;
    S.a = aL;
    S.b = bL;
    S.lambda = prm.s.lambdaInitial[lvl];
    S.iteration = 0;
    S.upToDate = false;
    accepted = -1;
    traceLambda = 0.0f;
    take = true;
  } else if (err.mean < S.lastErr.mean) {
    accepted = 1;
#pragma unroll
    for (int i = 0; i < 4; i++) S.q[i] = S.qt[i];
#pragma unroll
    for (int i = 0; i < 3; i++) S.t[i] = S.tt[i];
    S.s = S.st;
    S.upToDate = false;
    S.a = aL;
    S.b = bL;
    if (err.mean / S.lastErr.mean > prm.s.convergenceEps[lvl]) S.iteration = maxIts;
    S.finalRes = S.lastErr = err;
    if (S.lambda <= 0.2f) S.lambda = 0; else S.lambda *= prm.s.lambdaSuccessFac;
    S.iteration++;
    take = true;
  } else {
    accepted = 0;
    if (!(S.absInc > prm.s.stepSizeMin[lvl])) {
      S.iteration = maxIts + 1;
    } else if (S.lambda == 0) {
      S.lambda = 0.2f;
    } else {
      float f = 1.0f;
      for (int k = 0; k < S.incTry; k++) f *= prm.s.lambdaFailFac;
      S.lambda *= f;
    }
  }
  if (trace && O->traceLen < LSD_TRACE_CAP) trace[O->traceLen] = {lvl, accepted, err.mean, traceLambda, size};
  O->traceLen++;
  O->affine_a = S.a;
  O->affine_b = S.b;

  if (S.iteration >= maxIts) {
    // next level with iterations (upstream `continue`s over levels whose maxItsPerLvl is 0)
    int nl = lvl - 1;
    while (nl >= prm.finalLevel && prm.s.maxItsPerLvl[nl] == 0) nl--;
    if (nl >= prm.finalLevel) {
      S.level = nl;
      S.phase = 0;
      if (O->n[nl] == 0) {
        O->diverged = 1;
        s3_identity_out(O);
        return false;
      }
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, nl, next);
      return true;
    }
    if (!S.upToDate) {
      S.phase = 2;
      S.level = prm.finalLevel;
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, prm.finalLevel, next);
      return true;
    }
    s3_finish(S, O);
    return false;
  }
  if (take) {
#pragma unroll
    for (int k = 0; k < 41; k++) S.sums[k] = tot[k];
    S.nc = 2 * size;
    O->nWarp[lvl]++;
    S.incTry = 0;
    S.upToDate = true;
  }
  // A = ls7.A / num_constraints; b = -ls7.b / num_constraints; A(i,i) *= 1 + lambda; inc = A.ldlt().solve(b)
  float A[49], rhs[7], inc[7];
  {
#pragma unroll
    for (int i = 0; i < 49; i++) A[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) rhs[i] = 0;
    int k = 0;
    for (int a = 0; a < 6; a++) {
      for (int c = a; c < 6; c++, k++) A[a * 7 + c] = A[c * 7 + a] = S.sums[Q_A6 + k];
      rhs[a] = S.sums[Q_B6 + a];
    }
    const int remap[4] = {2, 3, 4, 6};
    k = 0;
    for (int a = 0; a < 4; a++) {
      for (int c = a; c < 4; c++, k++) {
        A[remap[a] * 7 + remap[c]] += S.sums[Q_A4 + k];
        if (c != a) A[remap[c] * 7 + remap[a]] += S.sums[Q_A4 + k];
      }
      rhs[remap[a]] += S.sums[Q_B4 + a];
    }
    const float nc = (float)S.nc;
    for (int i = 0; i < 49; i++) A[i] = A[i] / nc;
    for (int i = 0; i < 7; i++) rhs[i] = rhs[i] / nc;
    const float lam1 = 1 + S.lambda;
    for (int i = 0; i < 7; i++) A[i * 7 + i] *= lam1;
  }
  ldlt_solve<float, 7>(A, rhs, inc);
  S.incTry++;
  float absInc = 0;
  for (int i = 0; i < 7; i++) absInc += inc[i] * inc[i];
  S.absInc = absInc;
  if (!(absInc >= 0 && absInc < 1)) {  // upstream: lastSim3Hessian.setZero(); return Sim3(); (diverged is NOT set)
    for (int i = 0; i < 49; i++) O->H[i] = 0;
    s3_identity_out(O);
    return false;
  }
  double incd[7];
  for (int i = 0; i < 7; i++) incd[i] = (double)inc[i];
  QuatT<double> qc = {S.q[0], S.q[1], S.q[2], S.q[3]}, qn;
  double tn[3], sn;
  sim3_exp_compose<double>(incd, qc, S.t, S.s, qn, tn, sn);
  S.qt[0] = qn.x; S.qt[1] = qn.y; S.qt[2] = qn.z; S.qt[3] = qn.w;
  S.tt[0] = tn[0]; S.tt[1] = tn[1]; S.tt[2] = tn[2];
  S.st = sn;
  S.phase = 1;
  s3_make_cmd(S.qt, S.tt, S.st, S.a, S.b, lvl, next);
  return true;
}

__global__ void __cluster_dims__(S3_CL, 1, 1) __launch_bounds__(S3_THREADS, S3_MINB)
k_sim3_track(const Sim3Job *__restrict__ jobs, Sim3Out *__restrict__ outs, const __grid_constant__ Sim3Params prm,
             lsd_trace_entry *__restrict__ traces) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int jobIdx = blockIdx.x / S3_CL;
  const Sim3Job *J = jobs + jobIdx;
  Sim3Out *O = outs + jobIdx;
  extern __shared__ __align__(16) unsigned char s3_dyn_smem[];  // 56 KB: above the static limit
  S3Smem &sm = *reinterpret_cast<S3Smem *>(s3_dyn_smem);
  __shared__ S3Cmd cmd;
  __shared__ float part[S3_NF];
  __shared__ double dpart[S3_ND];
  __shared__ double dwarp[S3_ND][S3_THREADS / 32];
  __shared__ float tot[S3_NF];
  __shared__ double dtot[S3_ND];
  __shared__ S3State S;  // used by thread 0 of rank 0 only

  if (rank == 0 && threadIdx.x == 0) {
    memset(&S, 0, sizeof(S));
    for (int i = 0; i < 4; i++) S.q[i] = J->init[i];
    for (int i = 0; i < 3; i++) S.t[i] = J->init[4 + i];
    S.s = J->init[7];
    S.a = 1;
    S.b = 0;
    memset(O, 0, sizeof(Sim3Out));
    for (int l = 0; l < NL; l++) O->n[l] = J->d_num[l];
    O->affine_a = 1;
    S3Cmd first;
    int lvl = prm.startLevel;
    while (lvl >= prm.finalLevel && prm.s.maxItsPerLvl[lvl] == 0) lvl--;
    if (lvl < prm.finalLevel) {
      // no level has iterations: upstream still evaluates once at finalLevel (!warp_update_up_to_date)
      S.phase = 2;
      S.level = prm.finalLevel;
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, prm.finalLevel, first);
    } else if (O->n[lvl] == 0) {
      O->diverged = 1;
      s3_identity_out(O);
      first.op = 1;
    } else {
      S.level = lvl;
      S.phase = 0;
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, lvl, first);
    }
    for (unsigned r = 0; r < S3_CL; r++) *cluster.map_shared_rank(&cmd, r) = first;
  }
  cluster.sync();

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (;;) {
    if (cmd.op != 0) break;
    const int lvl = cmd.level;
    const int n = J->d_num[lvl];
    const int W = prm.K.w[lvl], H = prm.K.h[lvl];
    // contiguous part of this CTA (multiple of 32 points)
    const int per = (((n + S3_CL - 1) / S3_CL) + 31) & ~31;
    const int begin = min(n, (int)rank * per), end = min(n, begin + per);
#if S3_SMEM_ACC
    S3Acc acc = {&sm.f[0][threadIdx.x]};
#else
    float accReg[S3_NF];
    S3Acc acc = accReg;
#endif
    double dacc[S3_ND];
#pragma unroll
    for (int j = 0; j < S3_NF; j++) acc[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < S3_ND; j++) dacc[j] = 0.0;
    const float4 *pts4 = reinterpret_cast<const float4 *>(J->pts[lvl]);
    const float2 *rg = J->rgrad[lvl];
    {  // the next point's record (reference point + gradient) is loaded while the current point is evaluated
      const float4 *G = J->fgrad[lvl];
      const float *FID = J->fid[lvl], *FVAR = J->fvar[lvl];
      int i = begin + threadIdx.x;
      float4 rawC = make_float4(0, 0, 0, 0);
      float2 rgC = make_float2(0, 0);
      if (i < end) { rawC = __ldg(pts4 + i); rgC = __ldg(rg + i); }
      for (; i < end; i += S3_THREADS) {
        float4 rawN = rawC;
        float2 rgN = rgC;
        if (i + S3_THREADS < end) { rawN = __ldg(pts4 + i + S3_THREADS); rgN = __ldg(rg + i + S3_THREADS); }
        s3_point(rawC, rgC, cmd, prm, W, H, G, FID, FVAR, acc, dacc);
        rawC = rawN;
        rgC = rgN;
      }
    }
    // block reduction in a fixed order (same scheme as the SE3 tracker)
#if !S3_SMEM_ACC
#pragma unroll
    for (int j = 0; j < S3_NF; j++) sm.f[j][threadIdx.x] = acc[j];
#endif
#if S3_DSHUF
#pragma unroll
    for (int j = 0; j < S3_ND; j++) {
      double v = dacc[j];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) dwarp[j][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < S3_ND) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < S3_THREADS / 32; k++) v += dwarp[threadIdx.x][k];
      dpart[threadIdx.x] = v;
    }
    for (int row = wid; row < S3_NF; row += S3_THREADS / 32) {
#else
#pragma unroll
    for (int j = 0; j < S3_ND; j++) sm.d[j][threadIdx.x] = dacc[j];
    __syncthreads();
    for (int row = wid; row < S3_NF + S3_ND; row += S3_THREADS / 32) {
#endif
      if (row < S3_NF) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < S3_THREADS / 32; k++) v += sm.f[row][lane + 32 * k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) part[row] = v;
      }
#if !S3_DSHUF
      else {
        const int r = row - S3_NF;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < S3_THREADS / 32; k++) v += sm.d[r][lane + 32 * k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) dpart[r] = v;
      }
#endif
    }
    cluster.sync();  // every CTA's partial is in its shared memory
    if (rank == 0) {
      if (threadIdx.x < S3_NF) {
        float v = 0.0f;
        for (unsigned r = 0; r < S3_CL; r++) v += *cluster.map_shared_rank(&part[threadIdx.x], r);
        tot[threadIdx.x] = v;
      } else if (threadIdx.x < S3_NF + S3_ND) {
        const int j = threadIdx.x - S3_NF;
        double v = 0.0;
        for (unsigned r = 0; r < S3_CL; r++) v += *cluster.map_shared_rank(&dpart[j], r);
        dtot[j] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        S3Cmd next;
        next.op = 1;
        const bool more = s3_step(J, S, O, tot, dtot, prm, traces ? traces + (size_t)jobIdx * LSD_TRACE_CAP : nullptr, next);
        if (!more) next.op = 1;
        for (unsigned r = 0; r < S3_CL; r++) *cluster.map_shared_rank(&cmd, r) = next;
      }
    }
    cluster.sync();  // next command visible everywhere; partials may be overwritten
  }
}

static void sim3_inverse_host(const double p[8], double o[8]) {
  QuatT<double> qc = {-p[0], -p[1], -p[2], p[3]};
  double R[9], nt[3] = {p[4] * -1.0, p[5] * -1.0, p[6] * -1.0}, rt[3];
  qtoR(qc, R);
  mat3vec(R, nt, rt);
  const double si = 1.0 / p[7];
  o[0] = qc.x; o[1] = qc.y; o[2] = qc.z; o[3] = qc.w;
  o[4] = rt[0] * si; o[5] = rt[1] * si; o[6] = rt[2] * si;
  o[7] = si;
}

int sim3_track_batch_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init,
                          int startLevel, int finalLevel, lsd_sim3_result *results, lsd_trace_entry *traces) {
  if (n == 0) return LSD_OK;
  LSD_ARG(startLevel >= finalLevel && finalLevel >= 1 && startLevel < NL);
  cudaStream_t st = ctx->stream;
  const FrameLayout &lay = ctx->lay;
  // frames need their idepth pyramid (frame->idepth(level), idepthVar(level))
  for (int i = 0; i < n; i++) {
    LSD_ARG(refs[i] && frames[i]);
    LSD_ARG(frames[i]->built & FB_TRACKING);
    if (!(frames[i]->built & FB_IDEPTH0)) {
      set_error("trackFrameSim3: frame has no depth");
      return LSD_ERR_STATE;
    }
    int rc = frame_ensure_built(ctx, frames[i], FB_IDEPTH_PYR);
    if (rc) return rc;
  }
  const size_t jobBytes = sizeof(Sim3Job) * (size_t)n, outBytes = sizeof(Sim3Out) * (size_t)n;
  const size_t trBytes = traces ? sizeof(lsd_trace_entry) * LSD_TRACE_CAP * (size_t)n : 0;
  const size_t off1 = (jobBytes + 255) / 256 * 256, off2 = off1 + (outBytes + 255) / 256 * 256;
  int rc = ensure_stage(ctx, off2, off2 + trBytes);
  if (rc) return rc;
  Sim3Job *hj = reinterpret_cast<Sim3Job *>(ctx->h_stage);
  for (int i = 0; i < n; i++) {
    Sim3Job &J = hj[i];
    std::memset(&J, 0, sizeof(J));
    for (int l = 0; l < NL; l++) {
      J.pts[l] = reinterpret_cast<const RefPoint *>(refs[i]->slab + refs[i]->offPts[l]);
      J.rgrad[l] = reinterpret_cast<const float2 *>(refs[i]->slab + refs[i]->offGrad[l]);
      J.fgrad[l] = reinterpret_cast<const float4 *>(frames[i]->slab + lay.grad[l]);
      J.fid[l] = reinterpret_cast<const float *>(frames[i]->slab + lay.idepth[l]);
      J.fvar[l] = reinterpret_cast<const float *>(frames[i]->slab + lay.idvar[l]);
    }
    J.d_num = refs[i]->d_num;
    sim3_inverse_host(init + 8 * (size_t)i, J.init);
  }
  Sim3Params prm;
  prm.K = ctx->K;
  prm.s = ctx->sim3;
  prm.startLevel = startLevel;
  prm.finalLevel = finalLevel;
  Sim3Job *dj = reinterpret_cast<Sim3Job *>(ctx->d_stage);
  Sim3Out *dout = reinterpret_cast<Sim3Out *>(ctx->d_stage + off1);
  lsd_trace_entry *dtr = traces ? reinterpret_cast<lsd_trace_entry *>(ctx->d_stage + off2) : nullptr;
  LSD_CUDA(cudaMemcpyAsync(dj, hj, jobBytes, cudaMemcpyHostToDevice, st));
  LSD_CUDA(cudaEventRecord(ctx->evA, st));
  // per device, not per process: set on every call (cheap) so that a second context on another device works too
  LSD_CUDA(cudaFuncSetAttribute(k_sim3_track, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S3Smem)));
  k_sim3_track<<<n * S3_CL, S3_THREADS, sizeof(S3Smem), st>>>(dj, dout, prm, dtr);
  LSD_CUDA(cudaGetLastError());
  ctx->launches++;
  LSD_CUDA(cudaEventRecord(ctx->evB, st));
  Sim3Out *ho = reinterpret_cast<Sim3Out *>(ctx->h_stage + off1);
  LSD_CUDA(cudaMemcpyAsync(ho, dout, outBytes, cudaMemcpyDeviceToHost, st));
  if (traces) LSD_CUDA(cudaMemcpyAsync(traces, dtr, trBytes, cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
  double bytes = 0;
  long long evals = 0;
  for (int i = 0; i < n; i++) {
    const Sim3Out &o = ho[i];
    lsd_sim3_result &r = results[i];
    for (int k = 0; k < 8; k++) r.frameToRef[k] = o.frameToRef[k];
    std::memcpy(r.lastSim3Hessian, o.H, sizeof(o.H));
    r.lastResidual = o.lastResidual;
    r.lastDepthResidual = o.lastDepthResidual;
    r.lastPhotometricResidual = o.lastPhotometricResidual;
    r.pointUsage = o.pointUsage;
    r.affine_a = o.affine_a;
    r.affine_b = o.affine_b;
    r.diverged = o.diverged;
    r.traceLen = o.traceLen;
    for (int l = 0; l < NL; l++) {
      r.numResidualCalls[l] = o.nRes[l];
      r.numWarpUpdateCalls[l] = o.nWarp[l];
      // per evaluation 20 n + 16 min(4n, N) (as SE3) + 8 n (refGrad) + 8 min(n, N) (frame idepth/var) + 140
      const double N = (double)ctx->K.w[l] * ctx->K.h[l], nn = o.n[l];
      const double per = 20.0 * nn + 16.0 * (4.0 * nn < N ? 4.0 * nn : N) + 8.0 * nn + 8.0 * (nn < N ? nn : N) + 140.0;
      bytes += o.nRes[l] * per;
      evals += o.nRes[l];
      refs[i]->num[l] = o.n[l];
    }
    refs[i]->numValid = true;
  }
  ctx->lastAlgBytes = bytes;
  ctx->lastEvals = evals;
  ctx->lastKernelMs = ms;
  return LSD_OK;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

int lsd_ctx_set_sim3_settings(lsd_ctx *ctx, const lsd_tracker_settings *s) {
  LSD_ARG(ctx && s);
  ctx->sim3 = *s;
  return LSD_OK;
}

int lsd_sim3_track_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef,
                         int startLevel, int finalLevel, lsd_sim3_result *results, lsd_trace_entry *traces) {
  LSD_ARG(ctx && refs && frames && init_frameToRef && results && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return sim3_track_batch_impl(ctx, n, refs, frames, init_frameToRef, startLevel, finalLevel, results, traces);
}

int lsd_sim3_track(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double init_frameToRef[8], int startLevel, int finalLevel,
                   lsd_sim3_result *result, lsd_trace_entry *trace) {
  return lsd_sim3_track_batch(ctx, 1, &ref, &frame, init_frameToRef, startLevel, finalLevel, result, trace);
}

}  // extern "C"
