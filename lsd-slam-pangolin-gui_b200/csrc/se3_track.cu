// se3_track.cu -- [UP] SE3Tracker::trackFrame as ONE persistent kernel over a batch of independent
// (TrackingReference, Frame) pairs (BASELINE.json config 2).
//
// Reference structure (lsd-slam core Tracking/SE3Tracker.cpp, un-vendored): per LM evaluation three
// CPU passes -- calcResidualAndBuffers (8 SoA buffers out), calcWeightsAndResidual, and (per outer
// iteration) calculateWarpUpdate -- with a 6x6 LDLT + SE3 exp on the host in between.
//
// B200 structure:
//  * the three passes are fused into one per-point evaluation that never materialises the buffers;
//    every evaluation also reduces the 21+6 normal-equation terms (they are only CONSUMED if the step
//    is accepted, exactly when upstream would call calculateWarpUpdate on the same buffers);
//  * an evaluation is cut into RECORDS of SE3_REC consecutive points; every record is reduced by one CTA
//    in a fixed order (thread-strided partial sums, then a fixed shared-memory tree) and the records are
//    summed in record order, so every sum is a pure function of the inputs: bit-reproducible run to
//    run and independent of batch size, work-item size and scheduling;
//  * work items are (pair, run of consecutive records); a grid-resident kernel pulls items from a device
//    ring queue.  The CTA that completes a pair's last item sums the records,
//    runs the LM accept/reject logic, the 6x6 LDL^T solve
//    and the SE3 exponential on device, and pushes the pair's next evaluation into the queue.  Pairs
//    advance independently: no host round trip and no grid-wide barrier anywhere in a track;
//  * results are bit-reproducible run to run and independent of the batch composition.
//
// Compiled with -fmad=false: everything a discrete output depends on (warp, projection, bilinear sample,
// residual, isGood) is bit-identical to the oracle's (-ffp-contract=off); weights, Jacobian rows and the
// accumulated terms use explicit fmaf and MUFU approximations (see accumulate_point).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>

#include "ctx.cuh"
#include "lie_dev.cuh"
#include "reduce.cuh"

namespace lsd {

#ifndef SE3_THREADS
#define SE3_THREADS 128   // threads per CTA (small CTAs: an item's fetch / reduce barriers stall fewer warps)
#endif
#ifndef SE3_D
#define SE3_D 2           // per-thread software-pipeline depth of eval_range (points whose taps are in flight + 1)
#endif
#ifndef SE3_MINB
#define SE3_MINB 5        // resident CTAs per SM the register budget is sized for (r02n, 1000 pairs: 4.37 ms at 4 / 128 registers,
                          // 4.16 ms at 5 / 96 registers, 4.46 ms at 6 / 80, 6.66 ms at 8 / 64)
#endif
#ifndef SE3_REC
#define SE3_REC 4096      // points per partial record: FIXED, it defines the summation order (see above)
#endif
#define SE3_DEFAULT_ACTIVE 1000000  // pairs in flight (lsd_ctx_set_se3_active_pairs)
#define SE3_NRED 44       // floats per partial record: 5 doubles (affine sums) + 33 floats + pad
#define SE3_NF 33         // fp32 sums per record
#define SE3_ND 5          // fp64 sums per record

// fp32 sums (index into acc[] / the float part of a record, which starts at float offset 2*SE3_ND)
enum { R_A = 0, R_B = 21, R_SUMRES = 27, R_SUMUNW = 28, R_SUMSGN = 29, R_USAGE = 30, R_GOOD = 31, R_BAD = 32 };
// fp64 sums: the affine-lighting estimate sqrt((syy - sy^2/sw)/(sxx - sx^2/sw)) cancels ~20x (a) and
// ~100x (b = mean_y - a*mean_x), so its five sums are carried in fp64 end to end (5 DADD per point);
// everything else is consumed without cancellation and stays fp32.
enum { D_SXX = 0, D_SYY = 1, D_SX = 2, D_SY = 3, D_SW = 4 };

// Written by the host only, before any kernel that reads it is launched: immutable on the device, read with plain loads.
// (numData lives in the pair's SE3State, which is only ever read through L2: in the streamed host-image path k_se3_init fills
// it while the persistent tracker is already running, and a plain load could be served from a stale L1 line.)
struct SE3Pair {
  const RefPoint *pts[NL];
  const float4 *fgrad[NL];
  const int *d_num;  // ref numData[NL] on device
  uint8_t *mask;
  float q0[4], t0[3];  // initial referenceToFrame (float)
  int pad_;
};

// Mutable LM state: other SMs update it between evaluations, so inside the persistent kernel it is
// only read through L2 (__ldcg) -- never through a possibly stale L1 line.
struct __align__(16) SE3State {
  // --- evaluation header: what every CTA working on this pair's current evaluation needs (64 B) ---
  float R[9], t[3];  // pose of the evaluation in flight
  float aff_a, aff_b;
  int level;
  int nPts;  // numData[level]: every CTA derives the records / work items of the evaluation from it
  // --- LM bookkeeping (touched only by the thread running the LM step) ---
  float q_try[4];
  float q_cur[4], t_cur[3];
  int phase, iteration, incTry;
  float lambda, lastErr, last_residual;
  float A[21], b[6], inc[6];
  unsigned done;
  int finished, diverged, trackingWasGood;
  float pointUsage, good, bad, meanRes, aff_a_lastIt, aff_b_lastIt;
  int bufSize;
  int nResCalls[NL], nWarpCalls[NL];
  int traceLen;
  float outq[4], outt[3];  // frameToRef (float)
  float initialTrackedResidual;
  int n[NL];  // numData of the reference (read from the reference's device counters by k_se3_init)
  int nChunks;  // work items of the evaluation in flight
  int pad_[1];
};
static_assert(sizeof(SE3State) % 16 == 0, "SE3State must be int4-copyable");
static_assert(offsetof(SE3State, q_try) == 64, "evaluation header must be the first 64 bytes");

__device__ __forceinline__ void state_load(SE3State *dst, const SE3State *src) {
  const int4 *s4 = reinterpret_cast<const int4 *>(src);
  int4 *d4 = reinterpret_cast<int4 *>(dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(SE3State) / 16); i++) d4[i] = __ldcg(s4 + i);
}
__device__ __forceinline__ void state_store(SE3State *dst, const SE3State *src) {
  const int4 *s4 = reinterpret_cast<const int4 *>(src);
  int4 *d4 = reinterpret_cast<int4 *>(dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(SE3State) / 16); i++) d4[i] = s4[i];
}

// Device work queue: ring of 64-bit slots {sequence : 32, item code : 32}.  A slot is valid for ticket T
// when its sequence equals T / cap + 1.  At most n * maxChunks items are outstanding (one evaluation per
// pair), and cap >= that, so a slot is never overwritten before it has been consumed.
struct SE3Queue {
  unsigned long long *slots;
  unsigned *head;  // consumer tickets
  unsigned *tail;  // producer reservations
  int *remaining;  // pairs not finished yet
  unsigned *nextPair;  // admission: next pair to start when an active one finishes
  int *starved;        // set by the watchdog of a streamed launch (see SE3Params::watchdogNs)
  int nPairs;
  unsigned cap;    // power of two
};

struct SE3Params {
  Intrinsics K;
  lsd_tracker_settings s;
  int maxChunks;           // per-pair stride of the partial records (records of the largest tracked level)
  int recsPerItem;         // records per work item: scheduling granularity only, never changes a result
  int minLevel, maxLevel;  // SE3TRACKING_MIN_LEVEL, SE3TRACKING_MAX_LEVEL-1
  int recPointsLvl[NL];    // points per partial record at each level (lsd_ctx_set_se3_record_points[_per_level]; default SE3_REC)
  int permaref;            // SE3Tracker::trackFrameOnPermaref: single level, no frame side effects, referenceToFrame returned
  // Streamed launches only (0 otherwise): a CTA that has waited this long for a work item while pairs are still outstanding
  // gives up and stops the whole launch.  The streamed host-image path starts the tracker BEFORE its producers (ingest kernels
  // on another stream) and relies on them becoming co-resident; where the platform serialises kernels instead (profilers,
  // sanitizers, CUDA_LAUNCH_BLOCKING, an exhausted SM) the tracker would otherwise spin forever.  The host then re-runs the
  // batch with ordinary per-launch scheduling.
  unsigned long long watchdogNs;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// gpu-scope acquire-release fetch-add: releases this CTA's record stores (ordered before it by the CTA barrier)
// and, for the CTA that takes the last ticket, acquires every other CTA's.
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// Publish the chunks of the pair's next evaluation.  The caller has stored the pair's state already.
__device__ void q_push(const SE3Queue &q, int pairIdx, int nch) {
  __threadfence();  // state (and everything before) visible before any consumer can see the items
  const unsigned base = atomicAdd(q.tail, (unsigned)nch);
  for (int c = 0; c < nch; c++) {
    const unsigned t = base + c;
    const unsigned long long v = ((unsigned long long)(t / q.cap + 1) << 32) | (unsigned)((pairIdx << 12) | c);
    *reinterpret_cast<volatile unsigned long long *>(&q.slots[t & (q.cap - 1)]) = v;
  }
}

__device__ __forceinline__ void set_eval_pose(SE3State *S, const float q[4], const float t[3]) {
  QuatT<float> qq = {q[0], q[1], q[2], q[3]};
  float R[9];
  qtoR(qq, R);
#pragma unroll
  for (int i = 0; i < 9; i++) S->R[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) S->t[i] = t[i];
}

__device__ void mark_diverged(SE3State *S) {
  S->diverged = 1;
  S->trackingWasGood = 0;
  S->outq[0] = S->outq[1] = S->outq[2] = 0; S->outq[3] = 1;  // upstream returns SE3()
  S->outt[0] = S->outt[1] = S->outt[2] = 0;
  S->finished = 1;
}

__device__ void finish_pair(SE3State *S, const SE3Params &prm) {
  // trackingWasGood / outputs exactly as at the end of SE3Tracker::trackFrame
  const int l1 = prm.minLevel;
  const float lastGood = S->good, lastBad = S->bad;
  S->trackingWasGood = !S->diverged && lastGood / (prm.K.w[l1] * prm.K.h[l1]) > LSD_MIN_GOODPERALL_PIXEL &&
                       lastGood / (lastGood + lastBad) > LSD_MIN_GOODPERGOODBAD_PIXEL;
  S->initialTrackedResidual = S->last_residual / S->pointUsage;
  if (prm.permaref) {  // trackFrameOnPermaref returns referenceToFrame itself
#pragma unroll
    for (int i = 0; i < 4; i++) S->outq[i] = S->q_cur[i];
#pragma unroll
    for (int i = 0; i < 3; i++) S->outt[i] = S->t_cur[i];
    S->finished = 1;
    return;
  }
  // frameToRef = referenceToFrame.inverse()  (float)
  QuatT<float> qc = {-S->q_cur[0], -S->q_cur[1], -S->q_cur[2], S->q_cur[3]};
  float R[9], nt[3] = {-S->t_cur[0], -S->t_cur[1], -S->t_cur[2]}, ot[3];
  qtoR(qc, R);
  mat3vec(R, nt, ot);
  S->outq[0] = qc.x; S->outq[1] = qc.y; S->outq[2] = qc.z; S->outq[3] = qc.w;
  S->outt[0] = ot[0]; S->outt[1] = ot[1]; S->outt[2] = ot[2];
  S->finished = 1;
}

// S: the CTA's shared-memory copy of the state (or a thread-local one in k_se3_init).  Returns the number of chunks to publish (0: the pair is finished).
__device__ int start_level(SE3State *S, int level, const SE3Params &prm) {
  S->level = level;
  S->nPts = S->n[level];
  S->phase = 0;
  set_eval_pose(S, S->q_cur, S->t_cur);
  if (S->n[level] == 0) {  // calcResidualAndBuffers on an empty cloud: buf_warped_size 0 < 1% => diverged
    mark_diverged(S);
    return 0;
  }
  const int nRecs = (S->n[level] + prm.recPointsLvl[level] - 1) / prm.recPointsLvl[level];
  S->nChunks = (nRecs + prm.recsPerItem - 1) / prm.recsPerItem;  // work items of this evaluation
  S->done = 0;
  return S->nChunks;
}

// closed form of the affine-lighting estimate, evaluated in fp64 (upstream: fp32; mathematically identical, see DESIGN.md "affine lighting")
__device__ __forceinline__ void affine_estimate(const double *dtot, float *aL, float *bL) {
  const double sxx = dtot[D_SXX], syy = dtot[D_SYY], sx = dtot[D_SX], sy = dtot[D_SY], sw = dtot[D_SW];
  const double aLd = sqrt((syy - sy * sy / sw) / (sxx - sx * sx / sw));
  *aL = (float)aLd;
  *bL = (float)((sy - aLd * sx) / sw);
}

// What two other warps of the CTA compute WHILE thread 0 runs the LM decision logic (lm_step consumes it when it gets there):
//   pre[0..26] = the 21 + 6 normal-equation sums divided by num_constraints (27 IEEE divisions, one per lane of helper warp 1)
//   pre[27..28] = affine-lighting estimate (fp64 square root + four fp64 divisions, one lane of helper warp 2)
// Same operations on the same operands as the one-thread path: not a bit changes, the serial tail of an evaluation gets shorter.
// `ready[k] == seq` publishes part k of evaluation number `seq` of this CTA (seq only ever grows: no reset, no ABA).
struct LmPre {
  volatile float pre[32];
  volatile int ready[2];
};
__device__ __forceinline__ void lm_prework(const float *tot, const double *dtot, LmPre *P, const int seq, const int t) {
  if (t >= 32 && t < 64) {
    const int k = t - 32;
    if (k < 27) {
      const float nf = (float)(int)(tot[R_GOOD] + tot[R_BAD]);
      P->pre[k] = tot[R_A + k] / nf;  // R_A .. R_B + 5 are contiguous
    }
    __syncwarp();
    __threadfence_block();
    if (k == 0) P->ready[0] = seq;
  } else if (t == 64) {
    float aL, bL;
    affine_estimate(dtot, &aL, &bL);
    P->pre[27] = aL;
    P->pre[28] = bL;
    __threadfence_block();
    P->ready[1] = seq;
  }
}

// The LM state machine, run by one thread after the last chunk of an evaluation (tot = summed partials).
// Returns the number of chunks of the next evaluation (0: pair finished).  P (optional): see lm_prework.
__device__ int lm_step(SE3State *S, const float *tot, const double *dtot, const SE3Params &prm, lsd_trace_entry *trace, const LmPre *P = nullptr,
                       const int seq = 0) {
  const int lvl = S->level;
  const float good = tot[R_GOOD], bad = tot[R_BAD];
  const int size = (int)(good + bad);
  S->bufSize = size;
  S->good = good;
  S->bad = bad;
  S->pointUsage = tot[R_USAGE] / (float)S->n[lvl];
  S->meanRes = tot[R_SUMSGN] / good;
  bool takeAffine = false;  // the step is accepted (or the level starts): aff_a / aff_b follow this evaluation's estimate
  auto affine = [&]() {     // runs once, right before lm_step returns: by then the helper warp has long finished
    float aL, bL;
    if (P) {
      while (P->ready[1] != seq) {}
      aL = P->pre[27];
      bL = P->pre[28];
    } else {
      affine_estimate(dtot, &aL, &bL);
    }
    S->aff_a_lastIt = aL;
    S->aff_b_lastIt = bL;
    if (takeAffine) {
      S->aff_a = aL;
      S->aff_b = bL;
    }
  };

  if (size < LSD_MIN_GOODPERALL_PIXEL_ABSMIN * prm.K.w[lvl] * prm.K.h[lvl]) {
    affine();
    mark_diverged(S);
    return 0;
  }
  const float error = tot[R_SUMRES] / (float)size;
  S->nResCalls[lvl]++;
  const int maxIts = prm.s.maxItsPerLvl[lvl];

  bool takeNormalEq = false;  // begin a new outer iteration with this evaluation's A, b
  int accepted;
  float traceLambda = S->lambda;
  if (S->phase == 0) {
    takeAffine = true;
    S->lastErr = error;
    S->lambda = prm.s.lambdaInitial[lvl];
    S->iteration = 0;
    accepted = -1;
    traceLambda = 0.0f;
    takeNormalEq = true;
  } else if (error < S->lastErr) {
    accepted = 1;
#pragma unroll
    for (int i = 0; i < 4; i++) S->q_cur[i] = S->q_try[i];
#pragma unroll
    for (int i = 0; i < 3; i++) S->t_cur[i] = S->t[i];
    takeAffine = true;
    if (error / S->lastErr > prm.s.convergenceEps[lvl]) S->iteration = maxIts;
    S->last_residual = S->lastErr = error;
    if (S->lambda <= 0.2f) S->lambda = 0; else S->lambda *= prm.s.lambdaSuccessFac;
    S->iteration++;
    takeNormalEq = true;
  } else {
    accepted = 0;
    float inc2 = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) inc2 += S->inc[i] * S->inc[i];
    if (!(inc2 > prm.s.stepSizeMin[lvl])) {
      S->iteration = maxIts + 1;  // level ends
    } else if (S->lambda == 0) {
      S->lambda = 0.2f;
    } else {
      float f = 1.0f;  // std::pow(lambdaFailFac, incTry): exact for the default factor 2
      for (int k = 0; k < S->incTry; k++) f *= prm.s.lambdaFailFac;
      S->lambda *= f;
    }
  }
  if (trace && S->traceLen < LSD_TRACE_CAP) trace[S->traceLen] = {lvl, accepted, error, traceLambda, size};
  S->traceLen++;

  if (S->iteration >= maxIts) {
    affine();  // the next level's header carries aff_a / aff_b
    if (lvl - 1 < prm.minLevel) {
      finish_pair(S, prm);
      return 0;
    }
    return start_level(S, lvl - 1, prm);
  }
  if (takeNormalEq) {  // NormalEquationsLeastSquares::finish(): divide by num_constraints
    if (P) {
      while (P->ready[0] != seq) {}
#pragma unroll
      for (int k = 0; k < 21; k++) S->A[k] = P->pre[k];
#pragma unroll
      for (int k = 0; k < 6; k++) S->b[k] = P->pre[21 + k];
    } else {
      const float nf = (float)size;
#pragma unroll
      for (int k = 0; k < 21; k++) S->A[k] = tot[R_A + k] / nf;
#pragma unroll
      for (int k = 0; k < 6; k++) S->b[k] = tot[R_B + k] / nf;
    }
    S->nWarpCalls[lvl]++;
    S->incTry = 0;
  }
  // solve (A + lambda*diag(A)) inc = b   (b already holds +sum w r J / n, i.e. "-ls.b")
  float Al[36], rhs[6], inc[6];
  {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int c = a; c < 6; c++) {
        Al[a * 6 + c] = Al[c * 6 + a] = S->A[k++];
      }
    const float lam1 = 1 + S->lambda;
#pragma unroll
    for (int a = 0; a < 6; a++) {
      Al[a * 6 + a] *= lam1;
      rhs[a] = S->b[a];
    }
  }
  ldlt_solve<float, 6>(Al, rhs, inc);
  S->incTry++;
#pragma unroll
  for (int i = 0; i < 6; i++) S->inc[i] = inc[i];
  QuatT<float> qc = {S->q_cur[0], S->q_cur[1], S->q_cur[2], S->q_cur[3]}, qn;
  float tn[3];
  se3_exp_compose<float>(inc, qc, S->t_cur, qn, tn);
  S->q_try[0] = qn.x; S->q_try[1] = qn.y; S->q_try[2] = qn.z; S->q_try[3] = qn.w;
  set_eval_pose(S, S->q_try, tn);
  S->phase = 1;
  S->done = 0;  // same level => same nChunks
  affine();
  return S->nChunks;
}

// ---------------------------------------------------------------------------------------------
// Fused calcResidualAndBuffers + calcWeightsAndResidual + calculateWarpUpdate, split in a warp stage and an
// accumulate stage that are software-pipelined per thread (eval_range).
// ---------------------------------------------------------------------------------------------
struct EvalConst {
  float R[9], t[3];
  float a, b;
  float fx, fy, cx, cy, fxi, fyi, cxi, cyi;
  float var_weight, huber_half;
  int W, H;
};

// What the warp stage of one point hands to its accumulate stage (kept in registers while the taps fly).
struct Pending {
  float Wx, Wy, Wz, pz, dx, dy, color, var;
  int midx;  // x + y*W (mask index)
  int st;    // 1: taps in flight, 0: projects outside the image, -1: no point (past the end)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Warp stage: TrackingReference position -> warp -> projection -> in-image test; when the point is inside, its four
// bilinear taps (16 B each) are sent straight from global memory into this thread's shared-memory slot
// (cp.async: no register is held while they are in flight).  EXACT tier (see below).
__device__ __forceinline__ void warp_point(const float4 raw, const EvalConst &c, const float4 *__restrict__ G, float4 *slot,
                                           Pending &w) {
  const uint32_t xy = __float_as_uint(raw.x);
  const int x = xy & 0xffff, y = xy >> 16;
  const float inv = raw.y;  // RefPoint::invDepth; pos = (1/idepth) * (fxi*x+cxi, fyi*y+cyi, 1): TrackingReference::makePointCloud
  const float px = inv * (c.fxi * x + c.cxi);
  const float py = inv * (c.fyi * y + c.cyi);
  const float pz = inv * 1.0f;
  w.Wx = (c.R[0] * px + c.R[1] * py + c.R[2] * pz) + c.t[0];
  w.Wy = (c.R[3] * px + c.R[4] * py + c.R[5] * pz) + c.t[1];
  w.Wz = (c.R[6] * px + c.R[7] * py + c.R[8] * pz) + c.t[2];
  w.pz = pz;
  w.color = raw.z;
  w.var = raw.w;
  const float u_new = (w.Wx / w.Wz) * c.fx + c.cx;
  const float v_new = (w.Wy / w.Wz) * c.fy + c.cy;
  w.midx = x + y * c.W;
  if (!(u_new > 1 && v_new > 1 && u_new < c.W - 2 && v_new < c.H - 2)) {  // inverse test excludes NaN
    w.st = 0;
    return;
  }
  const int ix = (int)u_new, iy = (int)v_new;
  w.dx = u_new - ix;
  w.dy = v_new - iy;
  w.st = 1;
  const float4 *bp = G + (ix + iy * c.W);
  cp_async16(slot, bp);
  cp_async16(slot + SE3_THREADS, bp + 1);
  cp_async16(slot + 2 * SE3_THREADS, bp + c.W);
  cp_async16(slot + 3 * SE3_THREADS, bp + c.W + 1);
}

// Per-point arithmetic comes in two tiers.
//  EXACT tier (plain operators under -fmad=false, IEEE division): everything a discrete output depends on --
//    the warp, the projection and in-image test (buf_warped_size), the bilinear sample, the residual and the
//    isGood test (good / bad counts, refPixelWasGood).  Bit-identical to the oracle at a given pose.
//  RELAXED tier (explicit fmaf, MUFU reciprocal / rsqrt, algebraically regrouped): Huber / variance weights,
//    Jacobian rows and every accumulated term.  These only enter sums whose order already differs from the
//    reference's; their 2^-22 relative error is far inside the 1e-4 residual / 1e-5 pose tolerances.
__device__ __forceinline__ float fast_rcp(float x) { return __fdividef(1.0f, x); }

__device__ __forceinline__ void accumulate_point(const Pending &w, const float4 p00, const float4 p10, const float4 p01,
                                                 const float4 p11, const EvalConst &c, uint8_t *__restrict__ mask,
                                                 float acc[SE3_NF], double dacc[SE3_ND]) {
  // ---- EXACT: getInterpolatedElement43 (this exact weight form and summation order), residual, isGood
  const float dxdy = w.dx * w.dy;
  const float w11 = dxdy, w01 = w.dy - dxdy, w10 = w.dx - dxdy, w00 = 1 - w.dx - w.dy + dxdy;
  const float gxI = w11 * p11.x + w01 * p01.x + w10 * p10.x + w00 * p00.x;
  const float gyI = w11 * p11.y + w01 * p01.y + w10 * p10.y + w00 * p00.y;
  const float cI = w11 * p11.z + w01 * p01.z + w10 * p10.z + w00 * p00.z;
  const float Wx = w.Wx, Wy = w.Wy, Wz = w.Wz, pz = w.pz;
  const float c1 = c.a * w.color + c.b;
  const float c2 = cI;
  const float residual = c1 - c2;
  const float r2 = residual * residual;
  // isGood = fl(r2 / D) < 1 with D = MAX_DIFF_CONSTANT + MAX_DIFF_GRAD_MULT * |g|^2 >= 1600.  The correctly rounded
  // quotient of two floats is below 1 exactly when r2 < D: r2 >= D gives a quotient >= 1, and r2 < D means
  // r2 <= pred(D) <= D (1 - 2^-24), strictly below the rounding midpoint D (1 - 2^-25) of [pred(1), 1].
  // So the comparison equals upstream's division bit for bit (NaN: both false).
  const float D = LSD_MAX_DIFF_CONSTANT + LSD_MAX_DIFF_GRAD_MULT * (gxI * gxI + gyI * gyI);
  const bool isGood = r2 < D;
  if (mask) mask[w.midx] = isGood;
  if (isGood) {
    acc[R_SUMUNW] += r2;
    acc[R_SUMSGN] += residual;
    acc[R_GOOD] += 1.0f;
  } else {
    acc[R_BAD] += 1.0f;
  }

  // ---- RELAXED from here on
  const float ar = fabsf(residual);
  const float weight = ar < 5.0f ? 1.0f : 5.0f * fast_rcp(ar);  // affine-lighting Huber weight
  const float c1w = c1 * weight, c2w = c2 * weight;
  dacc[D_SXX] += (double)(c1 * c1w);
  dacc[D_SYY] += (double)(c2 * c2w);
  dacc[D_SX] += (double)c1w;
  dacc[D_SY] += (double)c2w;
  dacc[D_SW] += (double)weight;

  const float z = fast_rcp(Wz);
  const float depthChange = pz * z;
  acc[R_USAGE] += depthChange < 1 ? depthChange : 1;

  // calcWeightsAndResidual: g0 = (tx z' - tz x') / (z'^2 d), d = 1 / p_z  =>  g0 = (tx z' - tz x') * (p_z / z'^2)
  const float gx = c.fx * gxI, gy = c.fy * gyI;
  const float z_sqr = z * z;
  const float kk = z_sqr * pz;
  const float g0 = fmaf(c.t[0], Wz, -c.t[2] * Wx) * kk;
  const float g1 = fmaf(c.t[1], Wz, -c.t[2] * Wy) * kk;
  const float drpdd = fmaf(gx, g0, gy * g1);
  const float s = c.var_weight * w.var;
  const float rs = rsqrtf(fmaf(s * drpdd, drpdd, LSD_CAMERA_PIXEL_NOISE2));  // sqrt(w_p)
  const float w_p = rs * rs;
  const float weighted_rp = ar * rs;
  const float wh = weighted_rp < c.huber_half ? 1.0f : c.huber_half * fast_rcp(weighted_rp);
  const float wgt = wh * w_p;
  acc[R_SUMRES] = fmaf(wgt, r2, acc[R_SUMRES]);

  // calculateWarpUpdate, regrouped:  v2 = -z (x' v0 + y' v1),  v3 = y' v2 - gy,  v4 = gx - x' v2,  v5 = x' v1 - y' v0
  float v[6];
  v[0] = z * gx;
  v[1] = z * gy;
  v[2] = -z * fmaf(Wx, v[0], Wy * v[1]);
  v[3] = fmaf(Wy, v[2], -gy);
  v[4] = fmaf(-Wx, v[2], gx);
  v[5] = fmaf(Wx, v[1], -Wy * v[0]);
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const float wa = v[a] * wgt;
#pragma unroll
    for (int cc = a; cc < 6; cc++) {
      acc[R_A + k] = fmaf(wa, v[cc], acc[R_A + k]);
      k++;
    }
  }
  const float rw = residual * wgt;
#pragma unroll
  for (int a = 0; a < 6; a++) acc[R_B + a] = fmaf(v[a], rw, acc[R_B + a]);
}

// All points [begin, end) of one evaluation handled by this CTA.  Thread t takes points begin + t + m * SE3_THREADS
// in order of m (this order is part of the summation order).  Software pipeline, SE3_D stages deep, per thread:
//   point m + SE3_D       : its 16-byte record is loaded into a register (one stage ahead of its warp stage)
//   point m + SE3_D - 1   : warp stage -> four cp.async taps into this thread's slot (m + SE3_D - 1) % SE3_D
//   point m               : cp.async.wait_group(SE3_D - 1), then the accumulate stage reads its slot
// so both global-memory latencies of a point (record, taps) overlap the arithmetic of the SE3_D - 1 points
// before it.  A thread only ever reads slots it filled itself: no CTA barrier inside the loop.
__device__ __forceinline__ void eval_range(const RefPoint *__restrict__ pts, int begin, int end, const float4 *__restrict__ G,
                                           uint8_t *__restrict__ mask, const EvalConst &c, float acc[SE3_NF],
                                           double dacc[SE3_ND], float4 *tapbuf, const int tid) {
  // tid: the thread's index inside the SE3_THREADS-wide group that works on this range (the whole CTA in k_se3_track; one of
  // four groups of a CTA in k_se3_track_live)
  const float4 *pts4 = reinterpret_cast<const float4 *>(pts);
  float4 *mySlot = tapbuf + tid;
  Pending pd[SE3_D];
  int iLoad = begin + tid;
  float4 rawNext = (iLoad < end) ? __ldg(pts4 + iLoad) : make_float4(0, 0, 0, 0);
  auto issue = [&](const int s) {
    const float4 raw = rawNext;
    const bool have = iLoad < end;
    iLoad += SE3_THREADS;
    if (iLoad < end) rawNext = __ldg(pts4 + iLoad);
    if (have) warp_point(raw, c, G, mySlot + s * 4 * SE3_THREADS, pd[s]);
    else pd[s].st = -1;
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < SE3_D - 1; s++) issue(s);
  const int steps = (end - begin + SE3_THREADS - 1) / SE3_THREADS;  // CTA-uniform
  for (int m0 = 0; m0 < steps; m0 += SE3_D) {
#pragma unroll
    for (int s = 0; s < SE3_D; s++) {
      issue((s + SE3_D - 1) % SE3_D);
      cp_async_wait<SE3_D - 1>();
      if (pd[s].st > 0) {
        const float4 *sl = mySlot + s * 4 * SE3_THREADS;
        accumulate_point(pd[s], sl[0], sl[SE3_THREADS], sl[2 * SE3_THREADS], sl[3 * SE3_THREADS], c, mask, acc, dacc);
      } else if (pd[s].st == 0 && mask) {
        mask[pd[s].midx] = 0;
      }
    }
  }
  cp_async_wait<0>();
}

// The reduction of one record (reduce.cuh): 128 threads x (33 fp32 + 5 fp64) partial sums -> 38 numbers.  Both tracker kernels use
// this one function, so they keep returning identical bits.
typedef RecordRed<SE3_NF, SE3_ND, SE3_THREADS / 32> SE3Red;  // per 128-thread group: the warps' sums of two consecutive records
struct SE3Smem {
  float4 taps[SE3_D * 4 * SE3_THREADS];  // [stage][tap][thread]: a warp's LDS.128 / cp.async rows are conflict-free
  SE3Red red;
};

__device__ __forceinline__ void block_reduce_store(const float (&acc)[SE3_NF], const double (&dacc)[SE3_ND], float *__restrict__ dst,
                                                   SE3Smem &smu, const int parity) {
  reduce_record<SE3_NF, SE3_ND, SE3_THREADS / 32>(acc, dacc, dst, smu.red, threadIdx.x, parity, [] { __syncthreads(); });
  // No fence here: the record is published by the release-ordered completion ticket below (the CTA barrier
  // orders these stores before thread 0's gpu-scope release; a per-thread __threadfence would also flush L1
  // -- CCTL.IVALL -- after every record and throw away the tap locality of the next one).
  __syncthreads();
}

__device__ __forceinline__ void load_eval_const(const SE3Params &prm, int level, const int4 h0, const int4 h1, const int4 h2,
                                                const int4 h3, EvalConst &c) {
  c.R[0] = __int_as_float(h0.x); c.R[1] = __int_as_float(h0.y); c.R[2] = __int_as_float(h0.z); c.R[3] = __int_as_float(h0.w);
  c.R[4] = __int_as_float(h1.x); c.R[5] = __int_as_float(h1.y); c.R[6] = __int_as_float(h1.z); c.R[7] = __int_as_float(h1.w);
  c.R[8] = __int_as_float(h2.x); c.t[0] = __int_as_float(h2.y); c.t[1] = __int_as_float(h2.z); c.t[2] = __int_as_float(h2.w);
  c.a = __int_as_float(h3.x);
  c.b = __int_as_float(h3.y);
  c.fx = prm.K.fx[level]; c.fy = prm.K.fy[level]; c.cx = prm.K.cx[level]; c.cy = prm.K.cy[level];
  c.fxi = prm.K.fxi[level]; c.fyi = prm.K.fyi[level]; c.cxi = prm.K.cxi[level]; c.cyi = prm.K.cyi[level];
  c.var_weight = prm.s.var_weight;
  c.huber_half = prm.s.huber_d / 2;
  c.W = prm.K.w[level];
  c.H = prm.K.h[level];
}

__global__ void __launch_bounds__(SE3_THREADS, SE3_MINB)
k_se3_track(const SE3Pair *__restrict__ pairs, SE3State *states, float *partials, const SE3Queue q,
            const __grid_constant__ SE3Params prm, lsd_trace_entry *traces) {
  __shared__ __align__(16) SE3Smem sm;
  __shared__ float stot[SE3_NF];
  __shared__ double sdtot[SE3_ND];
  __shared__ int sCode, sIsLast, sNext;
  __shared__ __align__(16) SE3State sState;  // the LM step works on a shared-memory copy (a thread-local one lived in local memory)
  __shared__ LmPre sPre;
  int lmSeq = 0;  // LM steps this CTA has run (CTA-uniform)
  if (threadIdx.x < 2) sPre.ready[threadIdx.x] = 0;

  unsigned ticket = 0;
  if (threadIdx.x == 0) ticket = atomicAdd(q.head, 1u);
  for (;;) {
    // ---- fetch the next work item (thread 0 spins on its ticket's slot) ----
    if (threadIdx.x == 0) {
      const unsigned slot = ticket & (q.cap - 1), seq = ticket / q.cap + 1;
      const volatile unsigned long long *sp = reinterpret_cast<const volatile unsigned long long *>(&q.slots[slot]);
      int code = -1;
      unsigned long long waitStart = 0;
      for (;;) {
        const unsigned long long v = *sp;
        if ((unsigned)(v >> 32) == seq) {
          code = (int)(unsigned)v;
          break;
        }
        if (*reinterpret_cast<const volatile int *>(q.remaining) <= 0) break;
        if (prm.watchdogNs) {
          const unsigned long long now = global_timer_ns();
          if (!waitStart) waitStart = now;
          else if (now - waitStart > prm.watchdogNs) {  // producers never became resident: stop the launch (see SE3Params)
            atomicExch(q.starved, 1);
            atomicExch(q.remaining, -(1 << 30));
            __threadfence();
            break;
          }
        }
        __nanosleep(40);
      }
      sCode = code;
      if (code >= 0) ticket = atomicAdd(q.head, 1u);  // next ticket is requested now, consumed after this item
    }
    __syncthreads();
    const int code = sCode;
    if (code < 0) break;
    const int pairIdx = code >> 12, chunk0 = code & 0xfff;
    const SE3Pair *P = pairs + pairIdx;
    SE3State *S = states + pairIdx;
    // evaluation header: 4 x LDG.128 through L2
    const int4 *hp = reinterpret_cast<const int4 *>(S);
    int4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1), h2 = __ldcg(hp + 2), h3 = __ldcg(hp + 3);
    int chunk = chunk0;
    bool haveState = false;  // sState holds this pair's state (true while this CTA keeps evaluating the same pair)
    // A pair whose next evaluation is ONE work item stays on this CTA: no queue hop, no state round trip through global
    // memory.  The coarse levels (a few hundred points, a third of all evaluations) are exactly that.
    for (;;) {
      const int level = h3.z, n = h3.w;
      EvalConst c;
      load_eval_const(prm, level, h0, h1, h2, h3, c);
      uint8_t *mask = (level == prm.minLevel && !prm.permaref) ? P->mask : nullptr;  // permaref: idxBuf == nullptr upstream

      float acc[SE3_NF];
      double dacc[SE3_ND];
#pragma unroll
      for (int j = 0; j < SE3_NF; j++) acc[j] = 0.0f;
#pragma unroll
      for (int j = 0; j < SE3_ND; j++) dacc[j] = 0.0;
      const int recPoints = prm.recPointsLvl[level];
      const int nRecs = (n + recPoints - 1) / recPoints;
      const int nch = (nRecs + prm.recsPerItem - 1) / prm.recsPerItem;
      const int rec0 = chunk * prm.recsPerItem, rec1 = min(nRecs, rec0 + prm.recsPerItem);
      for (int rec = rec0; rec < rec1; rec++) {
        if (rec > rec0) {
#pragma unroll
          for (int j = 0; j < SE3_NF; j++) acc[j] = 0.0f;
#pragma unroll
          for (int j = 0; j < SE3_ND; j++) dacc[j] = 0.0;
        }
        const int begin = rec * recPoints;
        const int end = min(n, begin + recPoints);
        eval_range(P->pts[level], begin, end, P->fgrad[level], mask, c, acc, dacc, sm.taps, threadIdx.x);
        float *dst = partials + ((size_t)pairIdx * prm.maxChunks + rec) * SE3_NRED;
        block_reduce_store(acc, dacc, dst, sm, rec & 1);
      }
      if (nch > 1) {
        if (threadIdx.x == 0) sIsLast = (atom_add_acq_rel(&S->done, 1u) == (unsigned)(nch - 1));
        __syncthreads();
        if (!sIsLast) break;
      }
      const float *recBase = partials + (size_t)pairIdx * prm.maxChunks * SE3_NRED;
      if (threadIdx.x < SE3_ND) {
        const double *src = reinterpret_cast<const double *>(recBase) + threadIdx.x;
        double s = 0.0;
        for (int cidx = 0; cidx < nRecs; cidx++) s += __ldcg(src + (size_t)cidx * (SE3_NRED / 2));
        sdtot[threadIdx.x] = s;
      } else if (threadIdx.x < SE3_ND + SE3_NF) {
        const int j = threadIdx.x - SE3_ND;
        const float *src = recBase + 2 * SE3_ND + j;
        float s = 0.0f;
        for (int cidx = 0; cidx < nRecs; cidx++) s += __ldcg(src + (size_t)cidx * SE3_NRED);
        stot[j] = s;
      } else if (!haveState && threadIdx.x >= 64 && threadIdx.x < 64 + (int)(sizeof(SE3State) / 16)) {
        // the pair's state comes in with one LDG.128 per thread of warps 2-3 while warps 0-1 sum the records
        const int k = threadIdx.x - 64;
        reinterpret_cast<int4 *>(&sState)[k] = __ldcg(reinterpret_cast<const int4 *>(S) + k);
      }
      __syncthreads();
      haveState = true;
      lmSeq++;
      if (threadIdx.x == 0) sNext = lm_step(&sState, stot, sdtot, prm, traces ? traces + (size_t)pairIdx * LSD_TRACE_CAP : nullptr, &sPre, lmSeq);
      else lm_prework(stot, sdtot, &sPre, lmSeq, threadIdx.x);
      __syncthreads();
      if (sNext == 1) {  // the next evaluation is a single work item: keep it (the header is the first 64 bytes of the state)
        const int4 *sp4 = reinterpret_cast<const int4 *>(&sState);
        h0 = sp4[0]; h1 = sp4[1]; h2 = sp4[2]; h3 = sp4[3];
        chunk = 0;
        __syncthreads();  // every thread has read sNext / the header before the next LM step rewrites them
        continue;
      }
      if (threadIdx.x < (int)(sizeof(SE3State) / 16))
        reinterpret_cast<int4 *>(S)[threadIdx.x] = reinterpret_cast<const int4 *>(&sState)[threadIdx.x];
      __syncthreads();  // the state stores precede thread 0's fence + publication below
      if (threadIdx.x == 0) {
        const int next = sNext;
        if (next > 0) {
          q_push(q, pairIdx, next);
        } else {
          // Admission control: the number of pairs in flight is bounded so that their level data stays in L2
          // between consecutive LM evaluations; a finished pair hands its slot to the next waiting one.
          for (;;) {
            const unsigned cand = atomicAdd(q.nextPair, 1u);
            if (cand >= (unsigned)q.nPairs) break;
            SE3State *S2 = states + cand;
            const int nch2 = __ldcg(&S2->nChunks);
            if (nch2 > 0 && !__ldcg(&S2->finished)) {  // initialised by k_se3_init, not started yet
              q_push(q, (int)cand, nch2);
              break;
            }
            atomicSub(q.remaining, 1);  // a pair that diverged at initialisation: nothing to run
          }
          __threadfence();
          atomicSub(q.remaining, 1);
        }
      }
      break;
    }
    __syncthreads();  // sCode / sIsLast / sm are reused by the next item
  }
}

// Build the initial state of every pair and publish the first evaluations (level maxLevel).
// `base`: index of the first pair this launch initialises inside the arrays / the queue's pair numbering (0 for a whole batch;
// the streamed host-image path feeds one chunk at a time into a tracker that is already running).
__global__ void k_se3_init(const SE3Pair *__restrict__ pairs, SE3State *__restrict__ states, int n, const SE3Queue q, SE3Params prm,
                           int active, int base) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n) return;
  const int i = base + li;
  const SE3Pair *P = pairs + i;
  SE3State L;
  memset(&L, 0, sizeof(L));
  for (int l = 0; l < NL; l++) L.n[l] = P->d_num[l];
  for (int k = 0; k < 4; k++) L.q_cur[k] = P->q0[k];
  for (int k = 0; k < 3; k++) L.t_cur[k] = P->t0[k];
  L.aff_a = 1;
  L.aff_a_lastIt = 1;
  L.trackingWasGood = 1;
  const int next = start_level(&L, prm.maxLevel, prm);
  state_store(states + i, &L);
  if (li < active) {  // the rest is admitted by finishing pairs (k_se3_track)
    if (next > 0) q_push(q, i, next);
    else atomicSub(q.remaining, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// The live tracker: ONE pair per thread-block cluster (SlamSystem's tracking thread has exactly one frame to track).
// The work-queue kernel above buys batch throughput with global-memory hand-offs: every evaluation of a pair costs a
// fence + queue publication, a poll, a header load through L2, a release ticket per record and a state round trip --
// about 11 us of fixed latency per evaluation, 30 evaluations per frame.  Here the pair lives in one cluster:
//   * the LM state sits in the leader CTA's shared memory; the other CTAs read the evaluation header from it over DSMEM;
//   * a CTA is LIVE_GROUPS groups of SE3_THREADS threads; group g of CTA r takes records myGroup, myGroup + G, ... and
//     reduces each with the SAME thread-strided order and the same reduction as k_se3_track (reduce_record, named barrier per group),
//     into the CTA's own shared memory;
//   * after one cluster barrier the leader gathers the records over DSMEM, sums them in record order and runs lm_step;
//     a second cluster barrier publishes the next header.
// Two hardware barriers per evaluation instead of five global-memory round trips.  Record boundaries, the order inside a
// record and the order of the records are those of k_se3_track, so for a given record size both kernels return the same
// bits (tests/test_gpu_se3.py::test_live_cluster_kernel_matches_queue_kernel).
// ---------------------------------------------------------------------------------------------
#define LIVE_GROUPS 4
#define LIVE_THREADS (LIVE_GROUPS * SE3_THREADS)
#define LIVE_MAX_PAIRS 8  // pairs of one live launch (their descriptors are kernel parameters)
struct SE3PairPack {
  SE3Pair p[LIVE_MAX_PAIRS];
};

__device__ __forceinline__ void group_bar(const int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(SE3_THREADS) : "memory");
}

#ifndef LIVE_NP
#define LIVE_NP 2  // points of a thread whose record loads and taps are in flight TOGETHER (r03d: 4 spills at 128 registers, 0.268 vs 0.245 ms)
#endif
struct LiveSmem {
  float4 taps[LIVE_NP * 4 * SE3_THREADS];  // [point][tap][thread]
  SE3Red red;
};

// eval_range for the live kernel.  Same points per thread in the same order (thread t: begin + t + m * SE3_THREADS, m ascending),
// so the same bits; but a live evaluation is a few points per thread and all latency, so instead of a steady-state software
// pipeline the thread requests LIVE_NP records at once, then all their taps at once, and only then accumulates them in order:
// two memory round trips per batch instead of one per point.
__device__ __forceinline__ void eval_range_live(const RefPoint *__restrict__ pts, int begin, int end, const float4 *__restrict__ G,
                                                uint8_t *__restrict__ mask, const EvalConst &c, float acc[SE3_NF], double dacc[SE3_ND],
                                                float4 *tapbuf, const int tid) {
  const float4 *pts4 = reinterpret_cast<const float4 *>(pts);
  float4 *mySlot = tapbuf + tid;
  for (int base = begin; base < end; base += LIVE_NP * SE3_THREADS) {  // group-uniform trip count
    float4 raw[LIVE_NP];
#pragma unroll
    for (int s = 0; s < LIVE_NP; s++) {
      const int i = base + tid + s * SE3_THREADS;
      raw[s] = i < end ? __ldg(pts4 + i) : make_float4(0, 0, 0, 0);
    }
    Pending pd[LIVE_NP];
#pragma unroll
    for (int s = 0; s < LIVE_NP; s++) {
      if (base + tid + s * SE3_THREADS < end) warp_point(raw[s], c, G, mySlot + s * 4 * SE3_THREADS, pd[s]);
      else pd[s].st = -1;
    }
    cp_async_commit();
    cp_async_wait<0>();
#pragma unroll
    for (int s = 0; s < LIVE_NP; s++) {
      if (pd[s].st > 0) {
        const float4 *sl = mySlot + s * 4 * SE3_THREADS;
        accumulate_point(pd[s], sl[0], sl[SE3_THREADS], sl[2 * SE3_THREADS], sl[3 * SE3_THREADS], c, mask, acc, dacc);
      } else if (pd[s].st == 0 && mask) {
        mask[pd[s].midx] = 0;
      }
    }
  }
}

#ifdef SE3_LIVE_TIMING  // variant builds only: where a live evaluation spends its time (leader thread 0, nanoseconds)
__device__ unsigned long long g_liveNs[64];  // [0..7]: evaluations, lm_step, wait at [A], whole kernel; [8 l + k]: level l -- count, [A] -> [B],
                                             // header, first record, its reduction, further records, wait at [B], gather + sum
#define LIVE_T(var) const unsigned long long var = global_timer_ns()
#else
#define LIVE_T(var)
#endif

// Cluster barrier that orders SHARED memory only.  barrier.cluster.arrive.release would put a MEMBAR.ALL.GPU in front of every
// arrive (a cluster-scope release has to push this thread's global stores to L2: ~1 us with the level-1 mask stores in flight),
// and nothing this kernel exchanges between CTAs lives in global memory.  What is exchanged -- the leader's header, every CTA's
// records -- is written to the writer's OWN shared memory: the CTA-scope fence below (MEMBAR.SC.CTA, ~40 cycles) retires those
// stores into the SM's shared memory before the thread arrives, and a peer's distributed-shared-memory load is issued only after
// its wait has completed, so it reads the retired value.  (The wait is an acquire: ptxas emits CCTL.IVALL with it.)
#ifndef LIVE_RELEASE_BARRIER
__device__ __forceinline__ void cluster_barrier() {
  __threadfence_block();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#else
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#endif
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// generic address of `p` (a shared-memory object of this CTA) inside CTA `rank` of the cluster
template <typename T> __device__ __forceinline__ T *dsmem_ptr(T *p, const unsigned rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
  return reinterpret_cast<T *>(out);
}

// recsPerCta: record slots in every CTA's shared memory (>= ceil(maxChunks / (clusterSize * LIVE_GROUPS)) * LIVE_GROUPS)
__global__ void __launch_bounds__(LIVE_THREADS, 1)
k_se3_track_live(const __grid_constant__ SE3PairPack pk, SE3State *states, const __grid_constant__ SE3Params prm, lsd_trace_entry *traces,
                 const int recsPerCta) {
  extern __shared__ __align__(16) unsigned char dsm[];
  LiveSmem *gsm = reinterpret_cast<LiveSmem *>(dsm);                                  // one per group
  float *recs = reinterpret_cast<float *>(dsm + sizeof(LiveSmem) * LIVE_GROUPS);      // [recsPerCta][SE3_NRED]
  __shared__ __align__(16) SE3State sState;
  __shared__ float stot[SE3_NF];
  __shared__ double sdtot[SE3_ND];
  __shared__ int sNext;
  __shared__ LmPre sPre;
  int lmSeq = 0;
  if (threadIdx.x < 2) sPre.ready[threadIdx.x] = 0;

  const unsigned CL = cluster_size(), rank = cluster_rank();
  const int pairIdx = blockIdx.x / CL;
  const SE3Pair *P = &pk.p[pairIdx];  // the pair descriptors ride in the kernel parameters: no upload before a live launch
  const int g = threadIdx.x / SE3_THREADS, tid = threadIdx.x % SE3_THREADS;
  const int G = (int)CL * LIVE_GROUPS, myGroup = (int)rank * LIVE_GROUPS + g;

  if (rank == 0 && threadIdx.x == 0) {  // k_se3_init
    SE3State *L = &sState;
    memset(L, 0, sizeof(SE3State));
    for (int l = 0; l < NL; l++) L->n[l] = P->d_num[l];
    for (int k = 0; k < 4; k++) L->q_cur[k] = P->q0[k];
    for (int k = 0; k < 3; k++) L->t_cur[k] = P->t0[k];
    L->aff_a = 1;
    L->aff_a_lastIt = 1;
    L->trackingWasGood = 1;
    sNext = start_level(L, prm.maxLevel, prm);
  }
  const int4 *leadHdr = reinterpret_cast<const int4 *>(dsmem_ptr(&sState, 0));
  const volatile int *leadNext = dsmem_ptr(&sNext, 0);

  LIVE_T(tStart);
  for (;;) {
    LIVE_T(tA0);
    cluster_barrier();  // [A] the leader's header / sNext are final; the leader has consumed the previous records
    LIVE_T(tA);
    if (*leadNext <= 0) break;
    const int4 h0 = leadHdr[0], h1 = leadHdr[1], h2 = leadHdr[2], h3 = leadHdr[3];
    const int level = h3.z, n = h3.w;
    EvalConst c;
    load_eval_const(prm, level, h0, h1, h2, h3, c);
    uint8_t *mask = (level == prm.minLevel && !prm.permaref) ? P->mask : nullptr;
    const int recPoints = prm.recPointsLvl[level];
    const int nRecs = (n + recPoints - 1) / recPoints;
#ifdef SE3_LIVE_TIMING
    unsigned long long tH = global_timer_ns(), tE = tH, tR = tH;
#endif
    int slot = g;
    for (int rec = myGroup; rec < nRecs; rec += G, slot += LIVE_GROUPS) {
      float acc[SE3_NF];
      double dacc[SE3_ND];
#pragma unroll
      for (int j = 0; j < SE3_NF; j++) acc[j] = 0.0f;
#pragma unroll
      for (int j = 0; j < SE3_ND; j++) dacc[j] = 0.0;
      const int begin = rec * recPoints;
      const int end = min(n, begin + recPoints);
      eval_range_live(P->pts[level], begin, end, P->fgrad[level], mask, c, acc, dacc, gsm[g].taps, tid);
#ifdef SE3_LIVE_TIMING
      if (rec == myGroup) tE = global_timer_ns();
#endif
      reduce_record<SE3_NF, SE3_ND, SE3_THREADS / 32>(acc, dacc, recs + (size_t)slot * SE3_NRED, gsm[g].red, tid, (slot / LIVE_GROUPS) & 1, [g] { group_bar(1 + g); });
#ifdef SE3_LIVE_TIMING
      if (rec == myGroup) tR = global_timer_ns();
#endif
    }
    LIVE_T(tD);
    cluster_barrier();  // [B] every record of the evaluation is in its CTA's shared memory
    LIVE_T(tB);
    if (rank == 0) {
      // gather: every thread fetches ONE float4 of one record over DSMEM (a thread that issues several distributed-shared-memory
      // loads gets them back one after the other -- r03h: 16 loads per thread cost 1.3 us, one load per thread 0.3 us) into the
      // groups' scratch; then 38 threads sum their column in record order, 16 local loads in flight ahead of the adds
      constexpr int CAP = (int)(sizeof(LiveSmem) * LIVE_GROUPS / (SE3_NRED * sizeof(float)));  // records the scratch holds
      float *buf = reinterpret_cast<float *>(dsm);
      const int gShift = __ffs(G) - 1;  // G = 32 or 64
      double ds = 0.0;
      float fs = 0.0f;
      for (int r0 = 0; r0 < nRecs; r0 += CAP) {
        const int m = min(CAP, nRecs - r0);
        for (int i = threadIdx.x; i < m * (SE3_NRED / 4); i += LIVE_THREADS) {
          const int rl = i / (SE3_NRED / 4), q4 = i - rl * (SE3_NRED / 4);
          const int r = r0 + rl, grp = r & (G - 1), round = r >> gShift;
          const float *src = dsmem_ptr(recs, (unsigned)(grp / LIVE_GROUPS)) + (round * LIVE_GROUPS + grp % LIVE_GROUPS) * SE3_NRED;
          reinterpret_cast<float4 *>(buf)[i] = reinterpret_cast<const float4 *>(src)[q4];
        }
        __syncthreads();
        if (threadIdx.x < SE3_ND) {
          const double *src = reinterpret_cast<const double *>(buf) + threadIdx.x;
          for (int k0 = 0; k0 < m; k0 += 16) {
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = (k0 + k < m) ? src[(k0 + k) * (SE3_NRED / 2)] : 0.0;
#pragma unroll
            for (int k = 0; k < 16; k++)
              if (k0 + k < m) ds += v[k];
          }
        } else if (threadIdx.x < SE3_ND + SE3_NF) {
          const float *src = buf + 2 * SE3_ND + (threadIdx.x - SE3_ND);
          for (int k0 = 0; k0 < m; k0 += 16) {
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = (k0 + k < m) ? src[(k0 + k) * SE3_NRED] : 0.0f;
#pragma unroll
            for (int k = 0; k < 16; k++)
              if (k0 + k < m) fs += v[k];
          }
        }
        if (r0 + CAP < nRecs) __syncthreads();  // the scratch is refilled
      }
      if (threadIdx.x < SE3_ND) sdtot[threadIdx.x] = ds;
      else if (threadIdx.x < SE3_ND + SE3_NF) stot[threadIdx.x - SE3_ND] = fs;
      __syncthreads();
      LIVE_T(tS);
      lmSeq++;
      if (threadIdx.x == 0) sNext = lm_step(&sState, stot, sdtot, prm, traces ? traces + (size_t)pairIdx * LSD_TRACE_CAP : nullptr, &sPre, lmSeq);
      else lm_prework(stot, sdtot, &sPre, lmSeq, threadIdx.x);
#ifdef SE3_LIVE_TIMING
      if (threadIdx.x == 0) {
        const unsigned long long tL = global_timer_ns();
        g_liveNs[0] += 1; g_liveNs[1] += tL - tS; g_liveNs[2] += tA - tA0;
        unsigned long long *L = g_liveNs + 8 * level;
        L[0] += 1; L[1] += tB - tA; L[2] += tH - tA; L[3] += tE - tH; L[4] += tR - tE; L[5] += tD - tR; L[6] += tB - tD; L[7] += tS - tB;
      }
#endif
    }
  }
#ifdef SE3_LIVE_TIMING
  if (rank == 0 && threadIdx.x == 0) g_liveNs[3] += global_timer_ns() - tStart;
#endif
  if (rank == 0 && threadIdx.x < (int)(sizeof(SE3State) / 16))
    reinterpret_cast<int4 *>(states + pairIdx)[threadIdx.x] = reinterpret_cast<const int4 *>(&sState)[threadIdx.x];
  cluster_barrier();  // nobody leaves while a peer may still read its shared memory
}

struct SE3ScratchImpl {
  SE3Pair *d_pairs = nullptr;
  SE3Pair *h_pairs = nullptr;  // pinned
  SE3State *d_states = nullptr;
  SE3State *h_states = nullptr;  // pinned
  uint8_t **d_maskTab = nullptr, **h_maskTab = nullptr;  // frames whose refPixelWasGood mask is created by this call (pinned / device)
  int cap = 0;
  float *d_partials = nullptr;
  int maxChunks = 0;
  unsigned long long *d_slots = nullptr;
  unsigned qcap = 0;
  unsigned *d_ctrs = nullptr;  // head, tail, remaining, nextPair, starved, pad[3]
  lsd_trace_entry *d_traces = nullptr;
  size_t tracesBytes = 0;
  int gridBlocks = 0;
  int liveCluster = -1;  // -1: not probed; 0: k_se3_track_live cannot run here; else CTAs per cluster (16 where the device schedules it, else 8)
  int liveSmemSet = 0;   // dynamic shared memory the kernel attribute currently allows
};

}  // namespace lsd

struct SE3Scratch : lsd::SE3ScratchImpl {};

namespace lsd {

void se3_scratch_free(lsd_ctx *ctx) {
  SE3Scratch *s = ctx->se3s;
  if (!s) return;
  cudaFree(s->d_pairs);
  cudaFreeHost(s->h_pairs);
  cudaFree(s->d_states);
  cudaFreeHost(s->h_states);
  cudaFree(s->d_maskTab);
  cudaFreeHost(s->h_maskTab);
  cudaFree(s->d_partials);
  cudaFree(s->d_slots);
  cudaFree(s->d_ctrs);
  cudaFree(s->d_traces);
  delete s;
  ctx->se3s = nullptr;
}

// points per partial record at `level`: the per-level setting, else the context-wide one, else the default
int se3_record_points(const lsd_ctx *ctx, int level) {
  if (ctx->se3RecordPointsLvl[level] > 0) return ctx->se3RecordPointsLvl[level];
  return ctx->se3RecordPoints > 0 ? ctx->se3RecordPoints : SE3_REC;
}

static int se3_scratch_ensure(lsd_ctx *ctx, int n, bool wantTrace) {
  if (!ctx->se3s) ctx->se3s = new SE3Scratch();
  SE3Scratch *s = ctx->se3s;
  int maxChunks = 1;  // records of the level that has the most of them (worst case: every pixel of the level carries depth)
  for (int l = 1; l < NL; l++) {
    const int rp = se3_record_points(ctx, l);
    const int c = (ctx->K.w[l] * ctx->K.h[l] + rp - 1) / rp;
    if (c > maxChunks) maxChunks = c;
  }
  if (n > s->cap || maxChunks != s->maxChunks) {
    if (n < s->cap) n = s->cap;
    cudaFree(s->d_pairs);
    cudaFreeHost(s->h_pairs);
    cudaFree(s->d_states);
    cudaFreeHost(s->h_states);
    cudaFree(s->d_maskTab);
    cudaFreeHost(s->h_maskTab);
    cudaFree(s->d_partials);
    cudaFree(s->d_slots);
    int cap = n < 16 ? 16 : n;
    LSD_CUDA(cudaMalloc(&s->d_pairs, sizeof(SE3Pair) * cap));
    LSD_CUDA(cudaMallocHost(&s->h_pairs, sizeof(SE3Pair) * cap));
    LSD_CUDA(cudaMalloc(&s->d_states, sizeof(SE3State) * cap));
    LSD_CUDA(cudaMallocHost(&s->h_states, sizeof(SE3State) * cap));
    LSD_CUDA(cudaMalloc(&s->d_maskTab, sizeof(uint8_t *) * cap));
    LSD_CUDA(cudaMallocHost(&s->h_maskTab, sizeof(uint8_t *) * cap));
    LSD_CUDA(cudaMalloc(&s->d_partials, (size_t)cap * maxChunks * SE3_NRED * sizeof(float)));
    s->maxChunks = maxChunks;
    unsigned need = (unsigned)cap * (unsigned)maxChunks + 4096u, qcap = 1;
    while (qcap < need) qcap <<= 1;
    LSD_CUDA(cudaMalloc(&s->d_slots, sizeof(unsigned long long) * qcap));
    s->qcap = qcap;
    s->cap = cap;
    if (s->d_traces) {
      cudaFree(s->d_traces);
      s->d_traces = nullptr;
      s->tracesBytes = 0;
    }
  }
  if (!s->d_ctrs) LSD_CUDA(cudaMalloc(&s->d_ctrs, sizeof(unsigned) * 8));
  if (wantTrace && s->tracesBytes < sizeof(lsd_trace_entry) * LSD_TRACE_CAP * (size_t)s->cap) {
    cudaFree(s->d_traces);
    s->tracesBytes = sizeof(lsd_trace_entry) * LSD_TRACE_CAP * (size_t)s->cap;
    LSD_CUDA(cudaMalloc(&s->d_traces, s->tracesBytes));
  }
  if (!s->gridBlocks) {
    int perSM = 0;
    LSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_se3_track, SE3_THREADS, 0));
    if (perSM < 1) {
      set_error("k_se3_track cannot be resident");
      return LSD_ERR_CUDA;
    }
    s->gridBlocks = perSM * ctx->numSMs;  // every CTA resident: spinning consumers never starve producers
  }
  return LSD_OK;
}

static SE3Params make_params(lsd_ctx *ctx, int nPairs) {
  SE3Params prm;
  prm.K = ctx->K;
  prm.s = ctx->se3Permaref ? ctx->permaref : ctx->se3;
  prm.maxChunks = ctx->se3s->maxChunks;
  // Work-item size: one 4096-point record per item measured best at every batch size on B200 (profiles/);
  // the knob stays for experiments.
  prm.recsPerItem = ctx->se3RecsPerItem > 0 ? ctx->se3RecsPerItem : 1;
  prm.minLevel = LSD_SE3TRACKING_MIN_LEVEL;
  prm.maxLevel = LSD_SE3TRACKING_MAX_LEVEL - 1;
  for (int l = 0; l < NL; l++) prm.recPointsLvl[l] = se3_record_points(ctx, l);
  prm.permaref = ctx->se3Permaref ? 1 : 0;
  if (prm.permaref) prm.minLevel = prm.maxLevel = LSD_QUICK_KF_CHECK_LVL;
  prm.watchdogNs = 0;
  return prm;
}

// double frameToRef (Sophus data order) -> float referenceToFrame (q, t): inverse in fp64, then cast
static void invert_pose_to_float(const double p[7], float q[4], float t[3]) {
  QuatT<double> qc = {-p[0], -p[1], -p[2], p[3]};
  double R[9], nt[3] = {-p[4], -p[5], -p[6]}, ot[3];
  qtoR(qc, R);
  mat3vec(R, nt, ot);
  q[0] = (float)qc.x; q[1] = (float)qc.y; q[2] = (float)qc.z; q[3] = (float)qc.w;
  t[0] = (float)ot[0]; t[1] = (float)ot[1]; t[2] = (float)ot[2];
}

static double alg_bytes_level(const lsd_ctx *ctx, int l, int n) {
  // 20 n + 16 min(4n, N_l) + (idx 4n + isGood n at level 1) + 108
  const double N = (double)ctx->K.w[l] * ctx->K.h[l];
  double taps = 4.0 * n < N ? 4.0 * n : N;
  return 20.0 * n + 16.0 * taps + (l == LSD_SE3TRACKING_MIN_LEVEL ? 5.0 * n : 0.0) + 108.0;
}

static int init_masks(lsd_ctx *ctx, int n, lsd_frame *const *frames, cudaStream_t st) {
  // masks are created 0xFF on first use (Frame::refPixelWasGood()).  The pointer list has its own pinned / device table
  // in the tracker scratch (sized for the batch by se3_scratch_ensure), so no host synchronisation is needed here: the
  // table is next written by the next tracking call, which starts after this call's final synchronisation.
  SE3Scratch *s = ctx->se3s;
  int m = 0;
  for (int i = 0; i < n; i++)
    if (!(frames[i]->built & FB_MASK)) {
      s->h_maskTab[m++] = frames[i]->slab;
      frames[i]->built |= FB_MASK;
    }
  if (m == 0) return LSD_OK;
  if (m == 1 && st == ctx->stream) {  // one frame: its slab's entry in the context's device-resident pointer table (no upload;
                                       // entries are written in the order of the context's own stream)
    uint8_t *const *e = reinterpret_cast<uint8_t *const *>(ptr_table_entry(ctx, s->h_maskTab[0]));
    if (e) {
      launch_mask_init(ctx, e, 1, st);
      return LSD_OK;
    }
  }
  LSD_CUDA(cudaMemcpyAsync(s->d_maskTab, s->h_maskTab, sizeof(uint8_t *) * (size_t)m, cudaMemcpyHostToDevice, st));
  launch_mask_init(ctx, s->d_maskTab, m, st);
  return LSD_OK;
}

// resets the queue counters of one launch (head, tail, remaining, nextPair)
__global__ void k_se3_reset(unsigned *ctrs, unsigned n, unsigned active) {
  ctrs[0] = 0u;
  ctrs[1] = 0u;
  ctrs[2] = n;
  ctrs[3] = active;
  ctrs[4] = 0u;
}

// ---- the batch API in three steps, so that the host-image pipeline can queue several launches without a host
// ---- synchronisation in between:  prepare (pair table for ALL n pairs, one upload)  ->  launch(i0, m) ...  ->  collect
int se3_prepare(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init, bool wantTrace,
                cudaStream_t st, bool upload) {
  LSD_ARG(n < (1 << 19));
  int rc = se3_scratch_ensure(ctx, n, wantTrace);
  if (rc) return rc;
  SE3Scratch *s = ctx->se3s;
  const FrameLayout &lay = ctx->lay;
  for (int i = 0; i < n; i++) {
    LSD_ARG(refs[i] && frames[i]);
    LSD_ARG(frames[i]->built & FB_TRACKING);
    SE3Pair &P = s->h_pairs[i];
    std::memset(&P, 0, sizeof(P));
    for (int l = 0; l < NL; l++) {
      P.pts[l] = reinterpret_cast<const RefPoint *>(refs[i]->slab + refs[i]->offPts[l]);
      P.fgrad[l] = reinterpret_cast<const float4 *>(frames[i]->slab + lay.grad[l]);
    }
    P.d_num = refs[i]->d_num;
    P.mask = frames[i]->slab + lay.mask;
    if (ctx->se3Permaref) {  // the caller hands referenceToFrame: cast only
      for (int k = 0; k < 4; k++) P.q0[k] = (float)init[7 * i + k];
      for (int k = 0; k < 3; k++) P.t0[k] = (float)init[7 * i + 4 + k];
    } else {
      invert_pose_to_float(init + 7 * i, P.q0, P.t0);
    }
  }
  // (a live launch takes the descriptors as kernel parameters: se3_track_batch_impl uploads them only for the work-queue kernel)
  if (upload) LSD_CUDA(cudaMemcpyAsync(s->d_pairs, s->h_pairs, sizeof(SE3Pair) * n, cudaMemcpyHostToDevice, st));
  return LSD_OK;
}

static SE3Queue make_queue(SE3Scratch *s, int nPairs) {
  SE3Queue q;
  q.slots = s->d_slots;
  q.head = s->d_ctrs;
  q.tail = s->d_ctrs + 1;
  q.remaining = reinterpret_cast<int *>(s->d_ctrs + 2);
  q.nextPair = s->d_ctrs + 3;
  q.starved = reinterpret_cast<int *>(s->d_ctrs + 4);
  q.nPairs = nPairs;
  q.cap = s->qcap;
  return q;
}

// tracks pairs [i0, i0 + m) of the prepared table; everything it touches (queue, counters, partial records, states) is
// private to the launch or indexed by pair, and all of it is ordered on `st`
int se3_launch(lsd_ctx *ctx, int i0, int m, bool wantTrace, cudaStream_t st) {
  SE3Scratch *s = ctx->se3s;
  SE3Params prm = make_params(ctx, m);
  const SE3Queue q = make_queue(s, m);
  int active = ctx->se3ActivePairs > 0 ? ctx->se3ActivePairs : SE3_DEFAULT_ACTIVE;
  if (active > m) active = m;
  LSD_CUDA(cudaMemsetAsync(s->d_slots, 0, sizeof(unsigned long long) * q.cap, st));
  k_se3_reset<<<1, 1, 0, st>>>(s->d_ctrs, (unsigned)m, (unsigned)active);
  k_se3_init<<<(m + 127) / 128, 128, 0, st>>>(s->d_pairs + i0, s->d_states + i0, m, q, prm, active, 0);
  lsd_trace_entry *d_tr = wantTrace ? s->d_traces + (size_t)i0 * LSD_TRACE_CAP : nullptr;
  // every launched CTA polls the queue while idle: a small batch gets only as many CTAs as it can ever have work items in
  // flight (one evaluation per pair, at most maxChunks items each), not the whole machine
  const long long useful = (long long)m * prm.maxChunks;
  int grid = (int)(useful < (long long)s->gridBlocks ? (useful < 32 ? 32 : useful) : s->gridBlocks);
  static const int envGrid = getenv("LSD_B200_SE3_GRID") ? atoi(getenv("LSD_B200_SE3_GRID")) : 0;  // experiments only
  if (envGrid > 0 && envGrid <= s->gridBlocks) grid = envGrid;
  k_se3_track<<<grid, SE3_THREADS, 0, st>>>(s->d_pairs + i0, s->d_states + i0, s->d_partials + (size_t)i0 * prm.maxChunks * SE3_NRED,
                                                    q, prm, d_tr);
  LSD_CUDA(cudaGetLastError());
  ctx->launches += 3;
  return LSD_OK;
}

// ---- the live tracker: one cluster per pair (k_se3_track_live).  Returns 1 in *used when the launch was made, 0 when the
// ---- batch is not eligible (too many pairs, clusters unavailable, records of a level do not fit): the caller then takes se3_launch.
static int live_smem_bytes(const SE3Params &prm, int cluster, int *recsPerCta) {
  const int G = cluster * LIVE_GROUPS;
  *recsPerCta = LIVE_GROUPS * ((prm.maxChunks + G - 1) / G);
  return (int)(sizeof(LiveSmem) * LIVE_GROUPS + (size_t)*recsPerCta * SE3_NRED * sizeof(float));
}

static int se3_launch_live(lsd_ctx *ctx, int m, bool wantTrace, cudaStream_t st, int *used) {
  *used = 0;
  SE3Scratch *s = ctx->se3s;
  // default: up to two pairs, and only while level 1 is small enough for one cluster (r03h, 1280x960: level 1 has ~140 k points,
  // 17 per thread of a 16-CTA cluster; the work-queue kernel spreads them over the whole device: 1892 vs 1811 fps)
  const int limit = ctx->se3LivePairs < 0 ? (ctx->K.w[1] * ctx->K.h[1] <= 131072 ? 2 : 0) : ctx->se3LivePairs;
  if (m > limit || m > LIVE_MAX_PAIRS || s->liveCluster == 0) return LSD_OK;
  SE3Params prm = make_params(ctx, m);
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  cfg.blockDim = dim3(LIVE_THREADS, 1, 1);
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (s->liveCluster < 0) {  // probe once per context: the largest cluster this device can co-schedule
    static const int envCl = getenv("LSD_B200_SE3_LIVE_CLUSTER") ? atoi(getenv("LSD_B200_SE3_LIVE_CLUSTER")) : 0;  // experiments: 8 or 16
    s->liveCluster = 0;
    cudaFuncSetAttribute(k_se3_track_live, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cl = (envCl == 8 ? 8 : 16); cl >= 8; cl -= 8) {
      int recsPerCta = 0;
      const int bytes = live_smem_bytes(prm, cl, &recsPerCta);
      if (bytes > 220 * 1024) continue;
      // the cap is set ONCE per context (= per device), to the most any image size may ask for: later launches do not touch it
      if (cudaFuncSetAttribute(k_se3_track_live, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) continue;
      cfg.gridDim = dim3(cl, 1, 1);
      cfg.dynamicSmemBytes = bytes;
      attr[0].val.clusterDim.x = cl;
      attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
      int nClusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nClusters, k_se3_track_live, &cfg) == cudaSuccess && nClusters >= 1) {
        s->liveCluster = cl;
        break;
      }
    }
    cudaGetLastError();  // a failed probe is not an error of the call
    if (s->liveCluster == 0) return LSD_OK;
  }
  const int cl = s->liveCluster;
  int recsPerCta = 0;
  const int bytes = live_smem_bytes(prm, cl, &recsPerCta);
  if (bytes > 220 * 1024) return LSD_OK;  // a record size that small for this image size: the queue kernel takes it
  cfg.gridDim = dim3(cl * m, 1, 1);
  cfg.dynamicSmemBytes = bytes;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  lsd_trace_entry *d_tr = wantTrace ? s->d_traces : nullptr;
  SE3PairPack pk;
  std::memset(&pk, 0, sizeof(pk));
  for (int i = 0; i < m; i++) pk.p[i] = s->h_pairs[i];
  SE3State *d_states = s->d_states;
  LSD_CUDA(cudaLaunchKernelEx(&cfg, k_se3_track_live, pk, d_states, prm, d_tr, recsPerCta));
  ctx->launches += 1;
  *used = 1;
  return LSD_OK;
}

// ---- streamed variant for the host-image path: ONE persistent launch for all n prepared pairs, started before any
// ---- frame has arrived; chunks of pairs are fed into its queue (k_se3_init with a base index) as their frames have been
// ---- ingested on another stream.  The tracker leaves one CTA slot per SM free so that the ingest kernels of later chunks
// ---- can become resident next to it (a tracker that filled the machine would wait forever for work nobody can produce).
int se3_stream_begin(lsd_ctx *ctx, int n, cudaStream_t trackSt, cudaEvent_t armed) {
  SE3Scratch *s = ctx->se3s;
  SE3Params prm = make_params(ctx, n);
  prm.watchdogNs = ctx->streamWatchdogNs;
  const SE3Queue q = make_queue(s, n);
  LSD_CUDA(cudaMemsetAsync(s->d_slots, 0, sizeof(unsigned long long) * q.cap, trackSt));
  k_se3_reset<<<1, 1, 0, trackSt>>>(s->d_ctrs, (unsigned)n, (unsigned)n);  // remaining = n; admission is by feeding, not by nextPair
  LSD_CUDA(cudaEventRecord(armed, trackSt));  // feeders may push from here on (recorded BEFORE the persistent kernel)
  const int grid = s->gridBlocks - ctx->numSMs > ctx->numSMs ? s->gridBlocks - ctx->numSMs : ctx->numSMs;
  k_se3_track<<<grid, SE3_THREADS, 0, trackSt>>>(s->d_pairs, s->d_states, s->d_partials, q, prm, nullptr);
  LSD_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return LSD_OK;
}

int se3_stream_feed(lsd_ctx *ctx, int i0, int m, int n, cudaStream_t st) {
  SE3Scratch *s = ctx->se3s;
  SE3Params prm = make_params(ctx, n);
  const SE3Queue q = make_queue(s, n);
  k_se3_init<<<(m + 127) / 128, 128, 0, st>>>(s->d_pairs, s->d_states, m, q, prm, m, i0);
  LSD_CUDA(cudaGetLastError());
  ctx->launches++;
  return LSD_OK;
}

// Error path of the streamed host-image pipeline: makes the persistent tracker drain (remaining = 0 is its exit condition)
// and waits for it, so that no kernel is left spinning when the entry point returns early.  Ordered on a stream of its own:
// the tracker's stream is occupied by the tracker itself and the feeding stream may be blocked behind a failed step.
int se3_stream_abort(lsd_ctx *ctx, cudaStream_t trackSt, cudaStream_t sideSt) {
  SE3Scratch *s = ctx->se3s;
  if (!s || !s->d_ctrs) return LSD_OK;
  cudaMemsetAsync(s->d_ctrs + 2, 0, sizeof(unsigned), sideSt);
  cudaStreamSynchronize(sideSt);
  cudaStreamSynchronize(trackSt);
  return LSD_OK;
}

// 1 when the watchdog of the last streamed launch fired (the launch was stopped before every pair finished)
int se3_stream_starved(lsd_ctx *ctx, cudaStream_t st, int *starved) {
  SE3Scratch *s = ctx->se3s;
  int h = 0;
  LSD_CUDA(cudaMemcpyAsync(&h, s->d_ctrs + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  *starved = h;
  return LSD_OK;
}

int se3_collect(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, lsd_se3_result *results, lsd_trace_entry *traces,
                cudaStream_t st, float kernelMs) {
  SE3Scratch *s = ctx->se3s;
  LSD_CUDA(cudaMemcpyAsync(s->h_states, s->d_states, sizeof(SE3State) * n, cudaMemcpyDeviceToHost, st));
  if (traces)
    LSD_CUDA(cudaMemcpyAsync(traces, s->d_traces, sizeof(lsd_trace_entry) * LSD_TRACE_CAP * (size_t)n, cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  double bytes = 0;
  long long evals = 0;
  for (int i = 0; i < n; i++) {
    const SE3State &P = s->h_states[i];
    lsd_se3_result &r = results[i];
    r.frameToRef[0] = P.outq[0]; r.frameToRef[1] = P.outq[1]; r.frameToRef[2] = P.outq[2]; r.frameToRef[3] = P.outq[3];
    r.frameToRef[4] = P.outt[0]; r.frameToRef[5] = P.outt[1]; r.frameToRef[6] = P.outt[2];
    r.lastResidual = P.last_residual;
    r.lastMeanRes = P.meanRes;
    r.pointUsage = P.pointUsage;
    r.lastGoodCount = P.good;
    r.lastBadCount = P.bad;
    r.affine_a = P.aff_a;
    r.affine_b = P.aff_b;
    r.initialTrackedResidual = P.initialTrackedResidual;
    r.diverged = P.diverged;
    r.trackingWasGood = P.trackingWasGood;
    for (int l = 0; l < NL; l++) {
      r.numResidualCalls[l] = P.nResCalls[l];
      r.numWarpUpdateCalls[l] = P.nWarpCalls[l];
      bytes += P.nResCalls[l] * alg_bytes_level(ctx, l, P.n[l]);
      evals += P.nResCalls[l];
      refs[i]->num[l] = P.n[l];
    }
    refs[i]->numValid = true;
    r.traceLen = P.traceLen;
    // Frame bookkeeping done by SE3Tracker::trackFrame
    if (ctx->se3Permaref) continue;  // trackFrameOnPermaref leaves the frame and the keyframe counters alone
    lsd_frame *f = frames[i];
    if (!P.diverged) {
      f->initialTrackedResidual = P.initialTrackedResidual;
      for (int k = 0; k < 7; k++) f->thisToParent_raw[k] = r.frameToRef[k];
      f->thisToParent_raw[7] = 1.0;
      f->trackingParentId = refs[i]->frameID;
    }
    if (P.trackingWasGood) refs[i]->keyframe->numFramesTrackedOnThis++;
  }
#ifdef SE3_LIVE_TIMING
  {
    static int calls = 0;
    if (++calls % 100 == 0) {
      unsigned long long h[64];
      cudaMemcpyFromSymbol(h, g_liveNs, sizeof(h));
      const double e = (double)(h[0] ? h[0] : 1);
      std::fprintf(stderr, "[live timing] %d calls, %.1f evaluations/call, lm_step %.2f us, wait at [A] %.2f us per evaluation; kernel %.1f us/call\n", calls,
                   e / calls, 1e-3 * h[1] / e, 1e-3 * h[2] / e, 1e-3 * h[3] / calls);
      for (int l = 1; l < 5; l++) {
        const unsigned long long *L = h + 8 * l;
        const double c = (double)(L[0] ? L[0] : 1);
        std::fprintf(stderr, "[live timing]   level %d: %.1f evaluations/call, [A]->[B] %.2f us = header %.2f + first record %.2f + its reduction %.2f + further records %.2f "
                     "+ wait at [B] %.2f; gather+sum %.2f us\n", l, L[0] / (double)calls, 1e-3 * L[1] / c, 1e-3 * L[2] / c, 1e-3 * L[3] / c, 1e-3 * L[4] / c,
                     1e-3 * L[5] / c, 1e-3 * L[6] / c, 1e-3 * L[7] / c);
      }
    }
  }
#endif
  ctx->lastAlgBytes = bytes;
  ctx->lastEvals = evals;
  ctx->lastKernelMs = kernelMs;
  return LSD_OK;
}

int se3_track_batch_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init,
                         lsd_se3_result *results, lsd_trace_entry *traces, cudaStream_t st) {
  if (n == 0) return LSD_OK;
  int rc = se3_prepare(ctx, n, refs, frames, init, traces != nullptr, st, false);
  if (rc) return rc;
  if (!ctx->se3Permaref) {
    rc = init_masks(ctx, n, frames, st);
    if (rc) return rc;
  }
  LSD_CUDA(cudaEventRecord(ctx->evA, st));
  int live = 0;
  rc = se3_launch_live(ctx, n, traces != nullptr, st, &live);
  if (rc) return rc;
  if (!live) {
    LSD_CUDA(cudaMemcpyAsync(ctx->se3s->d_pairs, ctx->se3s->h_pairs, sizeof(SE3Pair) * n, cudaMemcpyHostToDevice, st));
    rc = se3_launch(ctx, 0, n, traces != nullptr, st);
  }
  if (rc) return rc;
  LSD_CUDA(cudaEventRecord(ctx->evB, st));
  // one host synchronisation per call: the read-back of the states is queued behind the kernels and se3_collect waits for it;
  // both events have completed by then
  rc = se3_collect(ctx, n, refs, frames, results, traces, st, 0.0f);
  if (rc) return rc;
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
  ctx->lastKernelMs = ms;
  return LSD_OK;
}

// ---------------------------------------------------------------------------------------
// SE3Tracker::checkPermaRefOverlap (B7): mean of min(1, z_ref / z') over the level-4 points that project inside
// (0, w-1) x (0, h-1).  One CTA per candidate; per-thread strided fp32 sums and a fixed tree: deterministic.
// ---------------------------------------------------------------------------------------
struct OverlapJob {
  const RefPoint *pts;
  const int *d_num;
  float R[9], t[3];
};

__global__ void __launch_bounds__(128) k_permaref_overlap(const OverlapJob *__restrict__ jobs, float *__restrict__ out, const Intrinsics K) {
  __shared__ float swarp[4];
  const OverlapJob &J = jobs[blockIdx.x];
  const int lvl = LSD_QUICK_KF_CHECK_LVL;
  const int n = J.d_num[lvl];
  const float w2 = (float)(K.w[lvl] - 1), h2 = (float)(K.h[lvl] - 1);
  const float fx = K.fx[lvl], fy = K.fy[lvl], cx = K.cx[lvl], cy = K.cy[lvl];
  float usage = 0.0f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const RefPoint p = J.pts[i];
    const int x = p.xy & 0xffff, y = p.xy >> 16;
    const float px = p.invDepth * (K.fxi[lvl] * x + K.cxi[lvl]), py = p.invDepth * (K.fyi[lvl] * y + K.cyi[lvl]), pz = p.invDepth * 1.0f;
    const float Wx = (J.R[0] * px + J.R[1] * py + J.R[2] * pz) + J.t[0];
    const float Wy = (J.R[3] * px + J.R[4] * py + J.R[5] * pz) + J.t[1];
    const float Wz = (J.R[6] * px + J.R[7] * py + J.R[8] * pz) + J.t[2];
    const float u_new = (Wx / Wz) * fx + cx, v_new = (Wy / Wz) * fy + cy;
    if (u_new > 0 && v_new > 0 && u_new < w2 && v_new < h2) {
      const float depthChange = pz / Wz;
      usage += depthChange < 1 ? depthChange : 1;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) usage += __shfl_down_sync(0xffffffffu, usage, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = usage;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = n > 0 ? (((swarp[0] + swarp[1]) + swarp[2]) + swarp[3]) / (float)n : 0.0f;
}

int se3_permaref_overlap_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, const double *refToFrame, float *pointUsage) {
  if (n == 0) return LSD_OK;
  cudaStream_t st = ctx->stream;
  const size_t jobBytes = (sizeof(OverlapJob) * (size_t)n + 255) / 256 * 256;
  int rc = ensure_table(ctx, jobBytes + sizeof(float) * (size_t)n);
  if (rc) return rc;
  OverlapJob *h = reinterpret_cast<OverlapJob *>(ctx->h_table);
  for (int i = 0; i < n; i++) {
    LSD_ARG(refs[i]);
    h[i].pts = reinterpret_cast<const RefPoint *>(refs[i]->slab + refs[i]->offPts[LSD_QUICK_KF_CHECK_LVL]);
    h[i].d_num = refs[i]->d_num;
    const double *p = refToFrame + 7 * (size_t)i;
    QuatT<float> q = {(float)p[0], (float)p[1], (float)p[2], (float)p[3]};  // SE3d -> SE3f, then rotationMatrix()
    qtoR(q, h[i].R);
    for (int k = 0; k < 3; k++) h[i].t[k] = (float)p[4 + k];
  }
  float *d_out = reinterpret_cast<float *>((char *)ctx->d_table + jobBytes);
  float *h_out = reinterpret_cast<float *>((char *)ctx->h_table + jobBytes);
  LSD_CUDA(cudaMemcpyAsync(ctx->d_table, h, sizeof(OverlapJob) * (size_t)n, cudaMemcpyHostToDevice, st));
  k_permaref_overlap<<<n, 128, 0, st>>>(reinterpret_cast<const OverlapJob *>(ctx->d_table), d_out, ctx->K);
  ctx->launches++;
  LSD_CUDA(cudaGetLastError());
  LSD_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  std::memcpy(pointUsage, h_out, sizeof(float) * (size_t)n);
  return LSD_OK;
}

// ---------------------------------------------------------------------------------------
// Single fused evaluation at a fixed pose (parity probe for B3+B4+B5).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SE3_THREADS)
k_se3_eval_once(const SE3Pair *__restrict__ P, const SE3State *__restrict__ S, float *__restrict__ partials, int level,
                const __grid_constant__ SE3Params prm) {
  __shared__ __align__(16) SE3Smem sm;
  const int4 *hp = reinterpret_cast<const int4 *>(S);
  const int4 h0 = __ldcg(hp), h1 = __ldcg(hp + 1), h2 = __ldcg(hp + 2), h3 = __ldcg(hp + 3);
  EvalConst c;
  load_eval_const(prm, level, h0, h1, h2, h3, c);
  const int n = P->d_num[level];
  uint8_t *mask = (level == prm.minLevel) ? P->mask : nullptr;
  float acc[SE3_NF];
  double dacc[SE3_ND];
#pragma unroll
  for (int j = 0; j < SE3_NF; j++) acc[j] = 0.0f;
#pragma unroll
  for (int j = 0; j < SE3_ND; j++) dacc[j] = 0.0;
  const int begin = blockIdx.x * prm.recPointsLvl[level];
  const int end = min(n, begin + prm.recPointsLvl[level]);
  eval_range(P->pts[level], begin, end, P->fgrad[level], mask, c, acc, dacc, sm.taps, threadIdx.x);
  block_reduce_store(acc, dacc, partials + (size_t)blockIdx.x * SE3_NRED, sm, 0);
}

int se3_eval_impl(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double refToFrame[7], int level, float a, float b,
                  float *A36, float *b6, float *scalars) {
  LSD_ARG(level >= 1 && level < NL);
  int rc = se3_scratch_ensure(ctx, 1, false);
  if (rc) return rc;
  SE3Scratch *s = ctx->se3s;
  cudaStream_t st = ctx->stream;
  const FrameLayout &lay = ctx->lay;
  rc = init_masks(ctx, 1, &frame, st);
  if (rc) return rc;
  SE3Pair &P = s->h_pairs[0];
  std::memset(&P, 0, sizeof(P));
  for (int l = 0; l < NL; l++) {
    P.pts[l] = reinterpret_cast<const RefPoint *>(ref->slab + ref->offPts[l]);
    P.fgrad[l] = reinterpret_cast<const float4 *>(frame->slab + lay.grad[l]);
  }
  P.d_num = ref->d_num;
  P.mask = frame->slab + lay.mask;
  SE3State &S = s->h_states[0];
  std::memset(&S, 0, sizeof(S));
  QuatT<float> q = {(float)refToFrame[0], (float)refToFrame[1], (float)refToFrame[2], (float)refToFrame[3]};
  qtoR(q, S.R);
  for (int k = 0; k < 3; k++) S.t[k] = (float)refToFrame[4 + k];
  S.aff_a = a;
  S.aff_b = b;
  LSD_CUDA(cudaMemcpyAsync(s->d_pairs, &P, sizeof(SE3Pair), cudaMemcpyHostToDevice, st));
  LSD_CUDA(cudaMemcpyAsync(s->d_states, &S, sizeof(SE3State), cudaMemcpyHostToDevice, st));
  SE3Params prm = make_params(ctx, 1);
  const int nblk = prm.maxChunks;
  k_se3_eval_once<<<nblk, SE3_THREADS, 0, st>>>(s->d_pairs, s->d_states, s->d_partials, level, prm);
  ctx->launches++;
  std::vector<float> h((size_t)nblk * SE3_NRED);
  int hnum[NL];
  LSD_CUDA(cudaMemcpyAsync(h.data(), s->d_partials, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaMemcpyAsync(hnum, ref->d_num, sizeof(hnum), cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  float tot[SE3_NF] = {0};
  double dtot[SE3_ND] = {0};
  const int nch = (hnum[level] + prm.recPointsLvl[level] - 1) / prm.recPointsLvl[level];
  for (int cidx = 0; cidx < nch; cidx++) {
    for (int j = 0; j < SE3_ND; j++) dtot[j] += reinterpret_cast<const double *>(&h[(size_t)cidx * SE3_NRED])[j];
    for (int j = 0; j < SE3_NF; j++) tot[j] += h[(size_t)cidx * SE3_NRED + 2 * SE3_ND + j];
  }
  const float size = tot[R_GOOD] + tot[R_BAD];
  int k = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      A36[i * 6 + j] = A36[j * 6 + i] = tot[R_A + k] / size;
      k++;
    }
  for (int i = 0; i < 6; i++) b6[i] = -tot[R_B + i] / size;
  const double sxx = dtot[D_SXX], syy = dtot[D_SYY], sx = dtot[D_SX], sy = dtot[D_SY], sw = dtot[D_SW];
  const double aL = std::sqrt((syy - sy * sy / sw) / (sxx - sx * sx / sw));
  scalars[0] = tot[R_SUMRES] / size;
  scalars[1] = tot[R_SUMUNW] / tot[R_GOOD];
  scalars[2] = size;
  scalars[3] = tot[R_GOOD];
  scalars[4] = tot[R_BAD];
  scalars[5] = tot[R_USAGE] / (float)hnum[level];
  scalars[6] = tot[R_SUMSGN] / tot[R_GOOD];
  scalars[7] = (float)aL;
  scalars[8] = (float)((sy - aL * sx) / sw);
  scalars[9] = tot[R_SUMRES] / size;
  scalars[10] = scalars[11] = 0;
  return LSD_OK;
}

}  // namespace lsd
