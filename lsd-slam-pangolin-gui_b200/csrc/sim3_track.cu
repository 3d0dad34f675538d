// sim3_track.cu -- [UP] Sim3Tracker::trackFrameSim3: 7-DoF
// photometric + depth alignment between keyframes, the unit of work of the constraint search
// (BASELINE.json configs[3]: a new keyframe against 64 candidates, sharded over GPUs with no collective).
//
// Reference structure (lsd-slam core Tracking/Sim3Tracker.cpp, un-vendored): per LM evaluation
// calcSim3Buffers (12 SoA buffers) -> calcSim3WeightsAndResidual -> calcSim3LGS (LGS6 + LGS4 -> LGS7), with
// the 7x7 LDLT and the fp64 Sim3 exponential on the host in between.
//
// B200 structure (round 2; round 1 ran one 4-CTA cluster per track with two cluster barriers per evaluation and sat at
// 0.17 of the HBM roofline): the persistent work-queue scheme of the SE3 tracker (se3_track.cu).
//  * the three reference passes are fused into one per-point evaluation that never materialises the buffers;
//  * an evaluation is cut into RECORDS of recPoints consecutive points (default 1024: a constraint search is at most a few
//    hundred tracks, so an evaluation is spread over many CTAs instead of four); every record is reduced by one CTA in a
//    fixed order and the records are summed in record order: bit-reproducible, independent of batch and scheduling;
//  * work items are (track, record); CTAs of one grid-resident kernel pull them from a device ring queue.  The CTA that
//    completes a track's last record runs accept / reject, the 7x7 LDL^T and the fp64 Sim3 exponential and publishes the
//    next evaluation.  Tracks advance independently: no cluster barrier, no grid barrier, no host round trip;
//  * per point, both global-memory latencies (point record; four gradient taps + the frame's own idepth / var at the
//    nearest pixel) are in flight two points ahead (cp.async into per-thread shared-memory slots).
// Arithmetic tiers as in the SE3 tracker: everything a discrete output depends on (warp, projection, in-image test, the
// bilinear sample, both residuals and the depth-validity tests) is IEEE-identical to the oracle (-fmad=false, correctly
// rounded division); weights, Huber factors and Jacobian rows -- which only enter sums whose order already differs from the
// reference's -- use fmaf and MUFU reciprocal / rsqrt (2^-22 relative, far inside the 1e-4 residual tolerance).
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"
#include "lie_dev.cuh"
#include "reduce.cuh"

namespace lsd {

#ifndef S3_THREADS
#define S3_THREADS 128
#endif
#ifndef S3_D
#define S3_D 2            // software-pipeline depth of s3_eval_range
#endif
#ifndef S3_MINB
#define S3_MINB 3         // resident CTAs per SM the register budget is sized for
#endif
#define S3_REC_DEFAULT 1024
// fp32 sums
enum { Q_A6 = 0, Q_B6 = 21, Q_A4 = 27, Q_B4 = 37, Q_RD = 41, Q_RP = 42, Q_USAGE = 43, Q_ND = 44, Q_CNT = 45, S3_NF = 46 };
#define S3_ND 5  // fp64 affine-lighting sums (sxx, syy, sx, sy, sw), see se3_track.cu
#define S3_NRED 64  // floats per partial record: 5 doubles + 46 floats + pad (256 B)

// Written by the host before the launch: immutable on the device.
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
  int nStages;                      // chained trackFrameSim3 calls per track (tryTrackSim3's [4,3], [2], [1] = 3)
  int startLevel[4], finalLevel[4];
  int recPoints;  // points per partial record: defines the summation order
  int maxRecs;    // per-track stride of the partial records
};
#define S3_MAX_STAGES 4

// what every CTA working on a track's current evaluation needs: the first 96 bytes of the track state
struct S3Cmd {
  float Rs[9], t[3];  // scaled rotation (rxso3) and translation, float
  float roll[4];      // xRoll0, xRoll1, yRoll0, yRoll1
  float a, b;
  int level;
  int op;    // 0 evaluate, 1 finished
  int nPts;  // numData[level]
  int pad_[3];
};
static_assert(sizeof(S3Cmd) == 96, "evaluation header = 6 x int4");

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
  float Adiv[36];  // the same, assembled into the 7x7 system and divided by num_constraints: 28 upper-triangle entries (row-major)
                   // + 7 right-hand sides -- kept so that a rejected step re-solves with a new lambda without re-dividing
  // outputs that accumulate over the track: kept here (the state is only ever read through L2) and written to Sim3Out once,
  // when the track finishes -- a read-modify-write of Sim3Out from whichever SM runs the LM step could hit a stale L1 line
  int n[NL], nRes[NL], nWarp[NL];
  int traceLen;
  float pointUsage;
  int stage;  // which of the chained trackFrameSim3 calls this track is in
};
static_assert(sizeof(S3State) % 16 == 0, "S3State must be int4-copyable");

// Mutable per-track state in global memory: only ever read through L2 (__ldcg) inside the persistent kernel.
struct __align__(16) S3Track {
  S3Cmd cmd;
  unsigned done;  // records of the evaluation in flight that have been reduced
  int pad_[3];
  S3State st;
};
static_assert(sizeof(S3Track) % 16 == 0 && offsetof(S3Track, st) % 16 == 0, "S3Track must be int4-copyable");

struct S3Queue {
  unsigned long long *slots;  // {sequence : 32, item : 32}; valid for ticket T when sequence == T / cap + 1
  unsigned *head, *tail;
  int *remaining;  // tracks not finished yet
  unsigned cap;    // power of two >= tracks * maxRecs
  int nTracks;
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


__device__ __forceinline__ unsigned s3_atom_add_acq_rel(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// publish the records of a track's next evaluation (its state has been stored already)
__device__ void s3_push(const S3Queue &q, int track, int nRecs) {
  __threadfence();
  const unsigned base = atomicAdd(q.tail, (unsigned)nRecs);
  for (int c = 0; c < nRecs; c++) {
    const unsigned t = base + c;
    const unsigned long long v = ((unsigned long long)(t / q.cap + 1) << 32) | (unsigned)((track << 12) | c);
    *reinterpret_cast<volatile unsigned long long *>(&q.slots[t & (q.cap - 1)]) = v;
  }
}

__device__ __forceinline__ void s3_cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void s3_cp_async4(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void s3_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void s3_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct S3Const {  // per-evaluation constants in registers
  float Rs[9], t[3], roll[4];
  float a, b;
  float fx, fy, cx, cy, fxi, fyi, cxi, cyi;
  float var_weight, huber;
  int W, H;
};

// what the warp stage of a point hands to its accumulate stage
struct S3Pending {
  float Wx, Wy, Wz, pz, dx, dy, color, var;
  float2 rg;
  int st;  // 1: loads in flight, 0: projects outside the image, -1: no point
};

// per-thread staging slots: [stage][tap][thread] float4 (conflict-free rows) and [stage][idepth|var][thread] float
struct S3Slots {
  float4 taps[S3_D * 4 * S3_THREADS];
  float dv[S3_D * 2 * S3_THREADS];
};

// EXACT tier: warp, projection, in-image test.  The four bilinear taps and the frame's own (idepth, var) at the nearest
// pixel go straight from global memory into this thread's slot.
__device__ __forceinline__ void s3_warp_point(const float4 raw, const float2 rg, const S3Const &c, const float4 *__restrict__ G,
                                              const float *__restrict__ FID, const float *__restrict__ FVAR, float4 *tapSlot,
                                              float *dvSlot, S3Pending &w) {
  const uint32_t xy = __float_as_uint(raw.x);
  const int x = xy & 0xffff, y = xy >> 16;
  const float inv = raw.y;  // RefPoint::invDepth
  const float px = inv * (c.fxi * x + c.cxi);
  const float py = inv * (c.fyi * y + c.cyi);
  const float pz = inv * 1.0f;
  w.pz = pz;
  w.Wx = (c.Rs[0] * px + c.Rs[1] * py + c.Rs[2] * pz) + c.t[0];
  w.Wy = (c.Rs[3] * px + c.Rs[4] * py + c.Rs[5] * pz) + c.t[1];
  w.Wz = (c.Rs[6] * px + c.Rs[7] * py + c.Rs[8] * pz) + c.t[2];
  w.color = raw.z;
  w.var = raw.w;
  w.rg = rg;
  const float u_new = (w.Wx / w.Wz) * c.fx + c.cx;
  const float v_new = (w.Wy / w.Wz) * c.fy + c.cy;
  if (!(u_new > 1 && v_new > 1 && u_new < c.W - 2 && v_new < c.H - 2)) {
    w.st = 0;
    return;
  }
  const int ix = (int)u_new, iy = (int)v_new;
  w.dx = u_new - ix;
  w.dy = v_new - iy;
  w.st = 1;
  const float4 *bp = G + (ix + iy * c.W);
  s3_cp_async16(tapSlot, bp);
  s3_cp_async16(tapSlot + S3_THREADS, bp + 1);
  s3_cp_async16(tapSlot + 2 * S3_THREADS, bp + c.W);
  s3_cp_async16(tapSlot + 3 * S3_THREADS, bp + c.W + 1);
  const int idx_rounded = (int)(u_new + 0.5f) + c.W * (int)(v_new + 0.5f);
  s3_cp_async4(dvSlot, FID + idx_rounded);
  s3_cp_async4(dvSlot + S3_THREADS, FVAR + idx_rounded);
}

__device__ __forceinline__ float s3_rcp(float x) { return __fdividef(1.0f, x); }

__device__ __forceinline__ void s3_accumulate(const S3Pending &w, const float4 p00, const float4 p10, const float4 p01, const float4 p11,
                                              const float id_frameDepth, const float var_frameDepth, const S3Const &c,
                                              float acc[S3_NF], double dacc[S3_ND]) {
  // ---- EXACT: getInterpolatedElement43 (this weight form and summation order), both residuals, the validity tests
  const float dxdy = w.dx * w.dy;
  const float w11 = dxdy, w01 = w.dy - dxdy, w10 = w.dx - dxdy, w00 = 1 - w.dx - w.dy + dxdy;
  const float gxI = w11 * p11.x + w01 * p01.x + w10 * p10.x + w00 * p00.x;
  const float gyI = w11 * p11.y + w01 * p01.y + w10 * p10.y + w00 * p00.y;
  const float cI = w11 * p11.z + w01 * p01.z + w10 * p10.z + w00 * p00.z;
  const float Wx = w.Wx, Wy = w.Wy, Wz = w.Wz, pz = w.pz;
  const float c1 = c.a * w.color + c.b;
  const float c2 = cI;
  const float rp = c1 - c2;
  const float z = 1.0f / Wz;  // ref_idepth (IEEE: rd below is a difference of two nearly equal inverse depths)
  const bool hasDepth = var_frameDepth > 0;
  const float rd = hasDepth ? z - id_frameDepth : -1.0f;
  acc[Q_CNT] += 1.0f;
  if (hasDepth) acc[Q_ND] += 1.0f;

  // ---- RELAXED from here on
  const float arp = fabsf(rp);
  const float weight = arp < 2.0f ? 1.0f : 2.0f * s3_rcp(arp);  // affine-lighting Huber weight (k = 2)
  const float c1w = c1 * weight, c2w = c2 * weight;
  dacc[0] += (double)(c1 * c1w);
  dacc[1] += (double)(c2 * c2w);
  dacc[2] += (double)c1w;
  dacc[3] += (double)c2w;
  dacc[4] += (double)weight;
  const float depthChange = pz * z;
  acc[Q_USAGE] += depthChange < 1 ? depthChange : 1;

  // USE_ESM_TRACKING: mean of the frame gradient and the rolled reference gradient
  const float rotatedGradX = fmaf(c.roll[0], w.rg.x, c.roll[1] * w.rg.y);
  const float rotatedGradY = fmaf(c.roll[2], w.rg.x, c.roll[3] * w.rg.y);
  const float gx = c.fx * 0.5f * (gxI + rotatedGradX);
  const float gy = c.fy * 0.5f * (gyI + rotatedGradY);

  // calcSim3WeightsAndResidual: g = (.) / (z'^2 d) with d = 1 / p_z  =>  (.) * p_z / z'^2
  const float z_sqr = z * z;
  const float kk = z_sqr * pz;
  const float g0 = fmaf(c.t[0], Wz, -c.t[2] * Wx) * kk;
  const float g1 = fmaf(c.t[1], Wz, -c.t[2] * Wy) * kk;
  const float g2 = (Wz - c.t[2]) * kk;
  const float drpdd = fmaf(gx, g0, gy * g1);
  const float s = c.var_weight * w.var;
  const float rsp = rsqrtf(fmaf(s * drpdd, drpdd, LSD_CAMERA_PIXEL_NOISE2));  // sqrt(w_p)
  const float w_p = rsp * rsp;
  const float weighted_rp = arp * rsp;
  float w_d = 0.0f, weighted_rd = 0.0f;
  if (hasDepth) {
    const float sv = c.var_weight * var_frameDepth;
    const float rsd = rsqrtf(fmaf(g2 * g2, s, sv));  // sqrt(w_d)
    w_d = rsd * rsd;
    weighted_rd = fabsf(rd) * rsd;
  }
  const float war = weighted_rd + weighted_rp;
  const float wh = war < c.huber ? 1.0f : c.huber * s3_rcp(war);
  const float wd = wh * w_d;  // 0 without a depth residual
  const float wp = wh * w_p;
  acc[Q_RD] = fmaf(wd * rd, rd, acc[Q_RD]);
  acc[Q_RP] = fmaf(wp * rp, rp, acc[Q_RP]);

  // calcSim3LGS (regrouped as in the SE3 tracker)
  float v[6], v4[4];
  v[0] = z * gx;
  v[1] = z * gy;
  v[2] = -z * fmaf(Wx, v[0], Wy * v[1]);
  v[3] = fmaf(Wy, v[2], -gy);
  v[4] = fmaf(-Wx, v[2], gx);
  v[5] = fmaf(Wx, v[1], -Wy * v[0]);
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

// All points [begin, end) of one record.  Thread t takes points begin + t + m * S3_THREADS in order of m (part of the summation
// order).  Software pipeline as in se3_track.cu: the point record one stage ahead of its warp stage, the taps S3_D - 1 points
// ahead of their accumulate stage; a thread only reads slots it filled itself, so there is no CTA barrier inside the loop.
__device__ __forceinline__ void s3_eval_range(const RefPoint *__restrict__ pts, const float2 *__restrict__ rgrad, int begin, int end,
                                              const float4 *__restrict__ G, const float *__restrict__ FID, const float *__restrict__ FVAR,
                                              const S3Const &c, float acc[S3_NF], double dacc[S3_ND], S3Slots &slots) {
  const float4 *pts4 = reinterpret_cast<const float4 *>(pts);
  float4 *myTap = slots.taps + threadIdx.x;
  float *myDv = slots.dv + threadIdx.x;
  S3Pending pd[S3_D];
  int iLoad = begin + threadIdx.x;
  float4 rawNext = make_float4(0, 0, 0, 0);
  float2 rgNext = make_float2(0, 0);
  if (iLoad < end) { rawNext = __ldg(pts4 + iLoad); rgNext = __ldg(rgrad + iLoad); }
  auto issue = [&](const int s) {
    const float4 raw = rawNext;
    const float2 rg = rgNext;
    const bool have = iLoad < end;
    iLoad += S3_THREADS;
    if (iLoad < end) { rawNext = __ldg(pts4 + iLoad); rgNext = __ldg(rgrad + iLoad); }
    if (have) s3_warp_point(raw, rg, c, G, FID, FVAR, myTap + s * 4 * S3_THREADS, myDv + s * 2 * S3_THREADS, pd[s]);
    else pd[s].st = -1;
    s3_cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < S3_D - 1; s++) issue(s);
  const int steps = (end - begin + S3_THREADS - 1) / S3_THREADS;  // CTA-uniform
  for (int m0 = 0; m0 < steps; m0 += S3_D) {
#pragma unroll
    for (int s = 0; s < S3_D; s++) {
      issue((s + S3_D - 1) % S3_D);
      s3_cp_async_wait<S3_D - 1>();
      if (pd[s].st > 0) {
        const float4 *sl = myTap + s * 4 * S3_THREADS;
        const float *dv = myDv + s * 2 * S3_THREADS;
        s3_accumulate(pd[s], sl[0], sl[S3_THREADS], sl[2 * S3_THREADS], sl[3 * S3_THREADS], dv[0], dv[S3_THREADS], c, acc, dacc);
      }
    }
  }
  s3_cp_async_wait<0>();
}

typedef RecordRed<S3_NF, S3_ND, S3_THREADS / 32> S3Red;
struct S3Smem {
  S3Slots slots;
  S3Red red;
};

// fixed-order block reduction of the per-thread sums into one partial record (reduce.cuh: the SE3 tracker's scheme)
__device__ __forceinline__ void s3_block_reduce_store(const float (&acc)[S3_NF], const double (&dacc)[S3_ND], float *dst, S3Smem &smu) {
  reduce_record<S3_NF, S3_ND, S3_THREADS / 32>(acc, dacc, dst, smu.red, threadIdx.x, 0, [] { __syncthreads(); });
  __syncthreads();  // the record (shared or global memory) is complete for the whole CTA; the scratch is free again
}

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
// the accumulated outputs of a finished track (see S3State)
__device__ void s3_flush(const S3State &S, Sim3Out *O) {
  O->pointUsage = S.pointUsage;
  O->affine_a = S.a;
  O->affine_b = S.b;
  O->traceLen = S.traceLen;
  for (int l = 0; l < NL; l++) {
    O->n[l] = S.n[l];
    O->nRes[l] = S.nRes[l];
    O->nWarp[l] = S.nWarp[l];
  }
}

// What two other warps of the CTA compute WHILE thread 0 runs the LM decision logic of s3_step (same operations on the same
// operands as the one-thread path, so not a bit changes; the serial tail of an evaluation gets shorter):
//   pre[0..34]  = the 7x7 normal equations of THIS evaluation: ls6 + the remapped ls4 terms, divided by num_constraints
//                 (28 upper-triangle entries row by row, then the 7 right-hand sides): helper warp 1, one or two entries per lane
//   pre[36..37] = affine-lighting estimate (fp64 square root + four fp64 divisions): one lane of helper warp 2
// `ready[k] == seq` publishes part k for LM step number `seq` of this CTA (seq only grows: no reset).
struct S3Pre {
  volatile float pre[40];
  volatile int ready[2];
};
__device__ __forceinline__ void s3_affine(const double *dtot, float *aL, float *bL) {
  const double sxx = dtot[0], syy = dtot[1], sx = dtot[2], sy = dtot[3], sw = dtot[4];
  const double aLd = sqrt((syy - sy * sy / sw) / (sxx - sx * sx / sw));
  *aL = (float)aLd;
  *bL = (float)((sy - aLd * sx) / sw);
}
// entry e of the divided system: e < 28 -> A(a, c) with a <= c in row-major upper-triangle order; e >= 28 -> rhs[e - 28]
__device__ __forceinline__ float s3_system_entry(const float *sums, const int e, const float nc) {
  int a, c;
  if (e < 28) {
    a = 0;
    int rem = e;
    while (rem >= 7 - a) { rem -= 7 - a; a++; }
    c = a + rem;
  } else {
    a = c = e - 28;
  }
  // index of row / column x inside the 4-parameter system (2, 3, 4, 6 -> 0, 1, 2, 3), -1 when x is not part of it
  const int m4a = a == 2 ? 0 : a == 3 ? 1 : a == 4 ? 2 : a == 6 ? 3 : -1;
  const int m4c = c == 2 ? 0 : c == 3 ? 1 : c == 4 ? 2 : c == 6 ? 3 : -1;
  float v = 0.0f;
  if (e < 28) {
    if (c < 6) v = sums[Q_A6 + a * 6 - a * (a - 1) / 2 + (c - a)];
    if (m4a >= 0 && m4c >= 0) v += sums[Q_A4 + m4a * 4 - m4a * (m4a - 1) / 2 + (m4c - m4a)];
  } else {
    if (a < 6) v = sums[Q_B6 + a];
    if (m4a >= 0) v += sums[Q_B4 + m4a];
  }
  return v / nc;
}
__device__ __forceinline__ void s3_prework(const float *tot, const double *dtot, S3Pre *P, const int seq, const int t) {
  if (t >= 32 && t < 64) {
    const int lane = t - 32;
    const float nc = (float)(2 * (int)tot[Q_CNT]);
    P->pre[lane] = s3_system_entry(tot, lane, nc);
    if (lane + 32 < 35) P->pre[lane + 32] = s3_system_entry(tot, lane + 32, nc);
    __syncwarp();
    __threadfence_block();
    if (lane == 0) P->ready[0] = seq;
  } else if (t == 64) {
    float aL, bL;
    s3_affine(dtot, &aL, &bL);
    P->pre[36] = aL;
    P->pre[37] = bL;
    __threadfence_block();
    P->ready[1] = seq;
  }
}

__device__ bool s3_step(const Sim3Job *J, S3State &S, Sim3Out *O, const float *tot, const double *dtot, const Sim3Params &prm,
                        lsd_trace_entry *trace, S3Cmd &next, const S3Pre *P = nullptr, const int seq = 0) {
  const int lvl = S.level;
  const int size = (int)tot[Q_CNT];
  S.pointUsage = tot[Q_USAGE] / (float)S.n[lvl];
  // the affine estimate is only consumed when the step is accepted (or the level starts): fetched from the helper at that point
  auto affine = [&](float &aL, float &bL) {
    if (P) {
      while (P->ready[1] != seq) {}
      aL = P->pre[36];
      bL = P->pre[37];
    } else {
      s3_affine(dtot, &aL, &bL);
    }
  };
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
    s3_finish(S, O);
    return false;
  }
  if (s3_too_few(size, lvl, prm)) {
    O->diverged = 1;
    s3_identity_out(O);
    return false;
  }
  S.nRes[lvl]++;
  const int maxIts = prm.s.maxItsPerLvl[lvl];
  bool take = false;
  int accepted;
  float traceLambda = S.lambda;
  if (S.phase == 0) {
    S.lastErr = err;
    affine(S.a, S.b);
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
    affine(S.a, S.b);
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
  if (trace && S.traceLen < LSD_TRACE_CAP) trace[S.traceLen] = {lvl, accepted, err.mean, traceLambda, size};
  S.traceLen++;

  if (S.iteration >= maxIts) {
    // next level with iterations (upstream `continue`s over levels whose maxItsPerLvl is 0)
    int nl = lvl - 1;
    while (nl >= prm.finalLevel[S.stage] && prm.s.maxItsPerLvl[nl] == 0) nl--;
    if (nl >= prm.finalLevel[S.stage]) {
      S.level = nl;
      S.phase = 0;
      if (S.n[nl] == 0) {
        O->diverged = 1;
        s3_identity_out(O);
        return false;
      }
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, nl, next);
      return true;
    }
    if (!S.upToDate) {
      S.phase = 2;
      S.level = prm.finalLevel[S.stage];
      s3_make_cmd(S.q, S.t, S.s, S.a, S.b, prm.finalLevel[S.stage], next);
      return true;
    }
    s3_finish(S, O);
    return false;
  }
  if (take) {
#pragma unroll
    for (int k = 0; k < 41; k++) S.sums[k] = tot[k];
    S.nc = 2 * size;
    S.nWarp[lvl]++;
    S.incTry = 0;
    S.upToDate = true;
    // A = ls7.A / num_constraints; b = -ls7.b / num_constraints (ls7 = ls6 + the remapped ls4 terms)
    if (P) {
      while (P->ready[0] != seq) {}
#pragma unroll
      for (int e = 0; e < 35; e++) S.Adiv[e] = P->pre[e];
    } else {
      const float nc = (float)S.nc;
      for (int e = 0; e < 35; e++) S.Adiv[e] = s3_system_entry(S.sums, e, nc);
    }
  }
  // A(i,i) *= 1 + lambda; inc = A.ldlt().solve(b)
  float A[49], rhs[7], inc[7];
  {
    int e = 0;
#pragma unroll
    for (int a = 0; a < 7; a++)
#pragma unroll
      for (int c = a; c < 7; c++, e++) A[a * 7 + c] = A[c * 7 + a] = S.Adiv[e];
#pragma unroll
    for (int i = 0; i < 7; i++) rhs[i] = S.Adiv[28 + i];
    const float lam1 = 1 + S.lambda;
#pragma unroll
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


// Starts stage `stage` of a track from the pose in S (q, t, s): the part of Sim3Tracker::trackFrameSim3 before the first
// evaluation.  Returns false when there is nothing to evaluate (the track's outputs for this stage are then complete).
__device__ bool s3_begin_stage(S3State &S, Sim3Out *O, const Sim3Params &prm, int stage, S3Cmd &first) {
  S.stage = stage;
  S.a = 1;
  S.b = 0;
  S.phase = S.iteration = S.incTry = 0;
  S.upToDate = false;
  S.lambda = S.absInc = 0;
  S.traceLen = 0;
  S.pointUsage = 0;
  for (int l = 0; l < NL; l++) S.nRes[l] = S.nWarp[l] = 0;
  memset(O, 0, sizeof(Sim3Out));
  memset(&first, 0, sizeof(first));
  first.op = 0;
  const int startLevel = prm.startLevel[stage], finalLevel = prm.finalLevel[stage];
  int lvl = startLevel;
  while (lvl >= finalLevel && prm.s.maxItsPerLvl[lvl] == 0) lvl--;
  if (lvl < finalLevel) {
    // no level has iterations: upstream still evaluates once at finalLevel (!warp_update_up_to_date)
    S.phase = 2;
    S.level = finalLevel;
    s3_make_cmd(S.q, S.t, S.s, S.a, S.b, finalLevel, first);
  } else {
    S.level = lvl;
    S.phase = 0;
    s3_make_cmd(S.q, S.t, S.s, S.a, S.b, lvl, first);
  }
  if (S.n[first.level] == 0) {  // an empty cloud at the first level to evaluate
    O->diverged = 1;
    s3_identity_out(O);
    first.op = 1;
    return false;
  }
  first.nPts = S.n[first.level];
  return true;
}

// The next stage starts from the previous stage's RESULT the way the host would hand it over: frameToReference (what
// trackFrameSim3 returned, fp64) inverted back into referenceToFrame -- the same two inversions, so a chained stage is
// bit-identical to a separate call.
__device__ void s3_pose_from_result(S3State &S, const Sim3Out *O) {
  const double *p = O->frameToRef;
  QuatT<double> qc = {-p[0], -p[1], -p[2], p[3]};
  double R[9], nt[3] = {p[4] * -1.0, p[5] * -1.0, p[6] * -1.0}, rt[3];
  qtoR(qc, R);
  mat3vec(R, nt, rt);
  const double si = 1.0 / p[7];
  S.q[0] = qc.x; S.q[1] = qc.y; S.q[2] = qc.z; S.q[3] = qc.w;
  S.t[0] = rt[0] * si; S.t[1] = rt[1] * si; S.t[2] = rt[2] * si;
  S.s = si;
}

__global__ void __launch_bounds__(S3_THREADS, S3_MINB)
k_sim3_track(const Sim3Job *__restrict__ jobs, S3Track *tracks, Sim3Out *outs, float *partials, const S3Queue q,
             const __grid_constant__ Sim3Params prm, lsd_trace_entry *traces) {
  __shared__ __align__(16) S3Smem sm;
  __shared__ __align__(16) float srec[S3_NRED];  // the partial record of a single-record evaluation never leaves the CTA
  __shared__ float stot[S3_NF];
  __shared__ double sdtot[S3_ND];
  __shared__ int sCode, sIsLast, sMore;
  __shared__ __align__(16) S3State sState;
  __shared__ __align__(16) S3Cmd sNext;
  __shared__ S3Pre sPre;
  int lmSeq = 0;  // LM steps this CTA has run (CTA-uniform)
  if (threadIdx.x < 2) sPre.ready[threadIdx.x] = 0;

  unsigned ticket = 0;
  if (threadIdx.x == 0) ticket = atomicAdd(q.head, 1u);
  for (;;) {
    if (threadIdx.x == 0) {  // fetch the next work item: thread 0 spins on its ticket's slot
      const unsigned slot = ticket & (q.cap - 1), seq = ticket / q.cap + 1;
      const volatile unsigned long long *sp = reinterpret_cast<const volatile unsigned long long *>(&q.slots[slot]);
      int code = -1;
      for (;;) {
        const unsigned long long v = *sp;
        if ((unsigned)(v >> 32) == seq) {
          code = (int)(unsigned)v;
          break;
        }
        if (*reinterpret_cast<const volatile int *>(q.remaining) <= 0) break;
        __nanosleep(40);
      }
      sCode = code;
      if (code >= 0) ticket = atomicAdd(q.head, 1u);
    }
    __syncthreads();
    const int code = sCode;
    if (code < 0) break;
    const int track = code >> 12;
    int rec = code & 0xfff;
    const Sim3Job *J = jobs + track;
    S3Track *T = tracks + track;
    // evaluation header through L2: 5 x LDG.128 + the point count
    S3Const c;
    int lvl, n;
    {
      const int4 *hp = reinterpret_cast<const int4 *>(T);
      const int4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1), h2 = __ldcg(hp + 2), h3 = __ldcg(hp + 3), h4 = __ldcg(hp + 4);
      c.Rs[0] = __int_as_float(h0.x); c.Rs[1] = __int_as_float(h0.y); c.Rs[2] = __int_as_float(h0.z); c.Rs[3] = __int_as_float(h0.w);
      c.Rs[4] = __int_as_float(h1.x); c.Rs[5] = __int_as_float(h1.y); c.Rs[6] = __int_as_float(h1.z); c.Rs[7] = __int_as_float(h1.w);
      c.Rs[8] = __int_as_float(h2.x); c.t[0] = __int_as_float(h2.y); c.t[1] = __int_as_float(h2.z); c.t[2] = __int_as_float(h2.w);
      c.roll[0] = __int_as_float(h3.x); c.roll[1] = __int_as_float(h3.y); c.roll[2] = __int_as_float(h3.z); c.roll[3] = __int_as_float(h3.w);
      c.a = __int_as_float(h4.x); c.b = __int_as_float(h4.y);
      lvl = h4.z;
      n = __ldcg(&T->cmd.nPts);
    }
    c.var_weight = prm.s.var_weight;
    c.huber = prm.s.huber_d;
    bool haveState = false;  // sState holds this track's state (true while the CTA keeps evaluating the same track)
    // A track whose next evaluation is a single record stays on this CTA: no queue hop, no state round trip through global
    // memory.  The coarse levels of a constraint search are exactly that: tens of tiny evaluations whose cost was all latency.
    for (;;) {
      c.fx = prm.K.fx[lvl]; c.fy = prm.K.fy[lvl]; c.cx = prm.K.cx[lvl]; c.cy = prm.K.cy[lvl];
      c.fxi = prm.K.fxi[lvl]; c.fyi = prm.K.fyi[lvl]; c.cxi = prm.K.cxi[lvl]; c.cyi = prm.K.cyi[lvl];
      c.W = prm.K.w[lvl];
      c.H = prm.K.h[lvl];
      float acc[S3_NF];
      double dacc[S3_ND];
#pragma unroll
      for (int j = 0; j < S3_NF; j++) acc[j] = 0.0f;
#pragma unroll
      for (int j = 0; j < S3_ND; j++) dacc[j] = 0.0;
      const int nRecs = (n + prm.recPoints - 1) / prm.recPoints;
      const int begin = rec * prm.recPoints, end = min(n, begin + prm.recPoints);
      s3_eval_range(J->pts[lvl], J->rgrad[lvl], begin, end, J->fgrad[lvl], J->fid[lvl], J->fvar[lvl], c, acc, dacc, sm.slots);
      float *recBase = partials + (size_t)track * prm.maxRecs * S3_NRED;
      const bool single = nRecs == 1;
      s3_block_reduce_store(acc, dacc, single ? srec : recBase + (size_t)rec * S3_NRED, sm);
      if (!single) {
        if (threadIdx.x == 0) sIsLast = (s3_atom_add_acq_rel(&T->done, 1u) == (unsigned)(nRecs - 1));
        __syncthreads();
        if (!sIsLast) break;
      }
      // ---- this CTA completed the evaluation: totals (records in record order), LM step
      if (threadIdx.x < S3_ND) {
        double s = 0.0;
        if (single) {
          s = reinterpret_cast<const double *>(srec)[threadIdx.x];
        } else {
          const double *src = reinterpret_cast<const double *>(recBase) + threadIdx.x;
          for (int r = 0; r < nRecs; r++) s += __ldcg(src + (size_t)r * (S3_NRED / 2));
        }
        sdtot[threadIdx.x] = s;
      } else if (threadIdx.x < S3_ND + S3_NF) {
        const int j = threadIdx.x - S3_ND;
        float s = 0.0f;
        if (single) {
          s = srec[2 * S3_ND + j];
        } else {
          const float *src = recBase + 2 * S3_ND + j;
          for (int r = 0; r < nRecs; r++) s += __ldcg(src + (size_t)r * S3_NRED);
        }
        stot[j] = s;
      } else if (!haveState && threadIdx.x >= 64 && threadIdx.x < 64 + (int)(sizeof(S3State) / 16)) {
        const int k = threadIdx.x - 64;
        reinterpret_cast<int4 *>(&sState)[k] = __ldcg(reinterpret_cast<const int4 *>(&T->st) + k);
      }
      __syncthreads();
      haveState = true;
      lmSeq++;
      if (threadIdx.x != 0) s3_prework(stot, sdtot, &sPre, lmSeq, threadIdx.x);
      if (threadIdx.x == 0) {
        S3Cmd next;
        next.op = 1;
        next.nPts = 0;
        Sim3Out *O = outs + (size_t)sState.stage * q.nTracks + track;
        bool more = s3_step(J, sState, O, stot, sdtot, prm, traces ? traces + (size_t)track * LSD_TRACE_CAP : nullptr, next, &sPre, lmSeq);
        if (more) {
          next.nPts = sState.n[next.level];
        } else {
          s3_flush(sState, O);
          // chained stages (tryTrackSim3 at [4,3], [2], [1]): the next stage starts from this stage's result; a track that
          // diverged or returned the identity ends here and its later stages report the same
          int stage = sState.stage;
          while (!more && stage + 1 < prm.nStages) {
            Sim3Out *On = outs + (size_t)(stage + 1) * q.nTracks + track;
            // upstream's own test in SlamSystem::tryTrackSim3: diverged, a degenerate scale, or an empty information matrix
            const bool dead = O->diverged || !(O->frameToRef[7] < 1e10) || !(O->frameToRef[7] > 1e-10) || O->H[0] == 0.0f || O->H[48] == 0.0f;
            if (dead) {
              *On = *O;
              On->diverged = 1;
              stage++;
              O = On;
              continue;
            }
            s3_pose_from_result(sState, O);
            more = s3_begin_stage(sState, On, prm, stage + 1, next);
            if (!more) s3_flush(sState, On);
            stage++;
            O = On;
          }
          if (!more) next.op = 1;
        }
        sNext = next;
        sMore = more ? 1 : 0;
      }
      __syncthreads();
      if (!sMore) {
        if (threadIdx.x == 0) {
          __threadfence();  // the track's outputs precede the completion count
          atomicSub(q.remaining, 1);
        }
        break;
      }
      const int nextRecs = (sNext.nPts + prm.recPoints - 1) / prm.recPoints;
      if (nextRecs == 1) {  // stay on this track: the header comes from shared memory
#pragma unroll
        for (int k = 0; k < 9; k++) c.Rs[k] = sNext.Rs[k];
#pragma unroll
        for (int k = 0; k < 3; k++) c.t[k] = sNext.t[k];
#pragma unroll
        for (int k = 0; k < 4; k++) c.roll[k] = sNext.roll[k];
        c.a = sNext.a;
        c.b = sNext.b;
        lvl = sNext.level;
        n = sNext.nPts;
        rec = 0;
        __syncthreads();  // every thread has read sNext / stot before the next step overwrites them
        continue;
      }
      // store the state and the next evaluation's header, then publish its records
      if (threadIdx.x < (int)(sizeof(S3State) / 16))
        reinterpret_cast<int4 *>(&T->st)[threadIdx.x] = reinterpret_cast<const int4 *>(&sState)[threadIdx.x];
      else if (threadIdx.x >= 64 && threadIdx.x < 64 + (int)(sizeof(S3Cmd) / 16))
        reinterpret_cast<int4 *>(&T->cmd)[threadIdx.x - 64] = reinterpret_cast<const int4 *>(&sNext)[threadIdx.x - 64];
      __syncthreads();
      if (threadIdx.x == 0) {
        T->done = 0;
        s3_push(q, track, nextRecs);
      }
      break;
    }
    __syncthreads();  // sCode / sIsLast / sm are reused by the next item
  }
}

// Initial state of every track and its first evaluation.
__global__ void k_sim3_init(const Sim3Job *__restrict__ jobs, S3Track *__restrict__ tracks, Sim3Out *__restrict__ outs, int n,
                            const S3Queue q, const Sim3Params prm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Sim3Job *J = jobs + i;
  S3Track *T = tracks + i;
  S3State S;
  memset(&S, 0, sizeof(S));
  for (int k = 0; k < 4; k++) S.q[k] = J->init[k];
  for (int k = 0; k < 3; k++) S.t[k] = J->init[4 + k];
  S.s = J->init[7];
  for (int l = 0; l < NL; l++) S.n[l] = J->d_num[l];
  S3Cmd first;
  const bool more = s3_begin_stage(S, outs + i, prm, 0, first);
  if (!more) {  // nothing to evaluate: stage 0 reports diverged, and so do the chained stages
    s3_flush(S, outs + i);
    for (int stage = 1; stage < prm.nStages; stage++) outs[(size_t)stage * n + i] = outs[i];
  }
  T->cmd = first;
  T->done = 0;
  T->st = S;
  if (more) s3_push(q, i, (first.nPts + prm.recPoints - 1) / prm.recPoints);
  else atomicSub(q.remaining, 1);
}

__global__ void k_sim3_reset(unsigned *ctrs, unsigned n) {
  ctrs[0] = 0u;
  ctrs[1] = 0u;
  ctrs[2] = n;
  ctrs[3] = 0u;
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

struct Sim3ScratchImpl {
  S3Track *d_tracks = nullptr;
  float *d_partials = nullptr;
  unsigned long long *d_slots = nullptr;
  unsigned *d_ctrs = nullptr;  // head, tail, remaining, pad
  int cap = 0, maxRecs = 0;
  unsigned qcap = 0;
  int gridBlocks = 0;
};

}  // namespace lsd

struct Sim3Scratch : lsd::Sim3ScratchImpl {};

namespace lsd {

void sim3_scratch_free(lsd_ctx *ctx) {
  Sim3Scratch *s = ctx->sim3s;
  if (!s) return;
  cudaFree(s->d_tracks);
  cudaFree(s->d_partials);
  cudaFree(s->d_slots);
  cudaFree(s->d_ctrs);
  delete s;
  ctx->sim3s = nullptr;
}

static int sim3_scratch_ensure(lsd_ctx *ctx, int n, int recPoints) {
  if (!ctx->sim3s) ctx->sim3s = new Sim3Scratch();
  Sim3Scratch *s = ctx->sim3s;
  const int maxRecs = (ctx->K.w[1] * ctx->K.h[1] + recPoints - 1) / recPoints;
  if (n > s->cap || maxRecs != s->maxRecs) {
    cudaFree(s->d_tracks);
    cudaFree(s->d_partials);
    cudaFree(s->d_slots);
    s->d_tracks = nullptr; s->d_partials = nullptr; s->d_slots = nullptr;
    int cap = n < s->cap ? s->cap : n;
    if (cap < 16) cap = 16;
    LSD_CUDA(cudaMalloc(&s->d_tracks, sizeof(S3Track) * (size_t)cap));
    LSD_CUDA(cudaMalloc(&s->d_partials, sizeof(float) * S3_NRED * (size_t)cap * maxRecs));
    unsigned need = (unsigned)cap * (unsigned)maxRecs + 1024u, qcap = 1;
    while (qcap < need) qcap <<= 1;
    LSD_CUDA(cudaMalloc(&s->d_slots, sizeof(unsigned long long) * qcap));
    s->qcap = qcap;
    s->cap = cap;
    s->maxRecs = maxRecs;
  }
  if (!s->d_ctrs) LSD_CUDA(cudaMalloc(&s->d_ctrs, sizeof(unsigned) * 4));
  if (!s->gridBlocks) {
    int perSM = 0;
    LSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_sim3_track, S3_THREADS, 0));
    if (perSM < 1) {
      set_error("k_sim3_track cannot be resident");
      return LSD_ERR_CUDA;
    }
    s->gridBlocks = perSM * ctx->numSMs;  // every CTA resident: spinning consumers never starve producers
  }
  return LSD_OK;
}

int sim3_track_batch_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init, int nStages,
                          const int *startLevels, const int *finalLevels, lsd_sim3_result *results, lsd_trace_entry *traces) {
  if (n == 0) return LSD_OK;
  LSD_ARG(nStages >= 1 && nStages <= S3_MAX_STAGES && startLevels && finalLevels);
  LSD_ARG(nStages == 1 || traces == nullptr);  // LM traces are per trackFrameSim3 call
  for (int k = 0; k < nStages; k++) LSD_ARG(startLevels[k] >= finalLevels[k] && finalLevels[k] >= 1 && startLevels[k] < NL);
  LSD_ARG(n < (1 << 19));
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
  const int recPoints = ctx->sim3RecordPoints > 0 ? ctx->sim3RecordPoints : S3_REC_DEFAULT;
  int rc = sim3_scratch_ensure(ctx, n, recPoints);
  if (rc) return rc;
  Sim3Scratch *sc = ctx->sim3s;
  const size_t jobBytes = sizeof(Sim3Job) * (size_t)n, outBytes = sizeof(Sim3Out) * (size_t)n * nStages;
  const size_t trBytes = traces ? sizeof(lsd_trace_entry) * LSD_TRACE_CAP * (size_t)n : 0;
  const size_t off1 = (jobBytes + 255) / 256 * 256, off2 = off1 + (outBytes + 255) / 256 * 256;
  rc = ensure_stage(ctx, off2, off2 + trBytes);
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
  prm.nStages = nStages;
  for (int k = 0; k < S3_MAX_STAGES; k++) {
    prm.startLevel[k] = k < nStages ? startLevels[k] : 0;
    prm.finalLevel[k] = k < nStages ? finalLevels[k] : 0;
  }
  prm.recPoints = recPoints;
  prm.maxRecs = sc->maxRecs;
  S3Queue q;
  q.slots = sc->d_slots;
  q.head = sc->d_ctrs;
  q.tail = sc->d_ctrs + 1;
  q.remaining = reinterpret_cast<int *>(sc->d_ctrs + 2);
  q.cap = sc->qcap;
  q.nTracks = n;
  Sim3Job *dj = reinterpret_cast<Sim3Job *>(ctx->d_stage);
  Sim3Out *dout = reinterpret_cast<Sim3Out *>(ctx->d_stage + off1);
  lsd_trace_entry *dtr = traces ? reinterpret_cast<lsd_trace_entry *>(ctx->d_stage + off2) : nullptr;
  LSD_CUDA(cudaMemcpyAsync(dj, hj, jobBytes, cudaMemcpyHostToDevice, st));
  LSD_CUDA(cudaEventRecord(ctx->evA, st));
  LSD_CUDA(cudaMemsetAsync(sc->d_slots, 0, sizeof(unsigned long long) * q.cap, st));
  k_sim3_reset<<<1, 1, 0, st>>>(sc->d_ctrs, (unsigned)n);
  k_sim3_init<<<(n + 63) / 64, 64, 0, st>>>(dj, sc->d_tracks, dout, n, q, prm);
  // every launched CTA polls the queue while idle: a small batch gets only as many CTAs as it can have records in flight
  const long long useful = (long long)n * prm.maxRecs;
  int grid = (int)(useful < (long long)sc->gridBlocks ? (useful < 32 ? 32 : useful) : sc->gridBlocks);
  static const int envGrid = getenv("LSD_B200_SIM3_GRID") ? atoi(getenv("LSD_B200_SIM3_GRID")) : 0;  // experiments only
  if (envGrid > 0 && envGrid <= sc->gridBlocks) grid = envGrid;
  k_sim3_track<<<grid, S3_THREADS, 0, st>>>(dj, sc->d_tracks, dout, sc->d_partials, q, prm, dtr);
  LSD_CUDA(cudaGetLastError());
  ctx->launches += 3;
  LSD_CUDA(cudaEventRecord(ctx->evB, st));
  Sim3Out *ho = reinterpret_cast<Sim3Out *>(ctx->h_stage + off1);
  LSD_CUDA(cudaMemcpyAsync(ho, dout, outBytes, cudaMemcpyDeviceToHost, st));
  if (traces) LSD_CUDA(cudaMemcpyAsync(traces, dtr, trBytes, cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
  double bytes = 0;
  long long evals = 0;
  for (int i = 0; i < n * nStages; i++) {  // stage-major: results[stage * n + track]
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
      if (i < n) refs[i]->num[l] = o.n[l];
    }
    if (i < n) refs[i]->numValid = true;
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

int lsd_ctx_set_sim3_record_points(lsd_ctx *ctx, int points) {
  LSD_ARG(ctx);
  LSD_ARG(points == 0 || (points >= 128 && points % 128 == 0 && points <= (1 << 20)));
  LSD_ARG(points == 0 || (ctx->K.w[1] * ctx->K.h[1] + points - 1) / points <= 4096);  // work-item codes carry 12 bits of record index
  ctx->sim3RecordPoints = points;
  return LSD_OK;
}

int lsd_sim3_track_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef,
                         int startLevel, int finalLevel, lsd_sim3_result *results, lsd_trace_entry *traces) {
  LSD_ARG(ctx && refs && frames && init_frameToRef && results && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return sim3_track_batch_impl(ctx, n, refs, frames, init_frameToRef, 1, &startLevel, &finalLevel, results, traces);
}

int lsd_sim3_track_stages_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef,
                                int nStages, const int *startLevels, const int *finalLevels, lsd_sim3_result *results) {
  LSD_ARG(ctx && refs && frames && init_frameToRef && results && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return sim3_track_batch_impl(ctx, n, refs, frames, init_frameToRef, nStages, startLevels, finalLevels, results, nullptr);
}

int lsd_sim3_track(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double init_frameToRef[8], int startLevel, int finalLevel,
                   lsd_sim3_result *result, lsd_trace_entry *trace) {
  return lsd_sim3_track_batch(ctx, 1, &ref, &frame, init_frameToRef, startLevel, finalLevel, result, trace);
}

}  // extern "C"
