// api.cu -- extern "C" surface of liblsd_b200.so (include/lsd_b200.h): contexts, device-resident
// frames, tracking references and the SE3 tracker entry points.  No CPU fallback exists: every
// entry point either runs the sm_100a kernels or returns an error.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "ctx.cuh"

// A few persistent host threads (created on first use): copying a batch of pageable images into the pinned staging buffer is
// the host-side cost of every multi-frame entry point (one memcpy thread moves ~10 GB/s; 32 VGA frames are 10 MB).
struct HostPool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cvWork, cvDone;
  std::function<void(int)> fn;
  int nTasks = 0, generation = 0, pending = 0;
  std::atomic<int> next{0};
  bool stop = false;
  explicit HostPool(int threads) {
    for (int t = 0; t < threads; t++)
      th.emplace_back([this]() {
        int seen = 0;
        for (;;) {
          {
            std::unique_lock<std::mutex> l(mu);
            cvWork.wait(l, [&]() { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
          }
          work();
          std::unique_lock<std::mutex> l(mu);
          if (--pending == 0) cvDone.notify_all();
        }
      });
  }
  ~HostPool() {
    {
      std::unique_lock<std::mutex> l(mu);
      stop = true;
    }
    cvWork.notify_all();
    for (std::thread &t : th) t.join();
  }
  void work() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= nTasks) return;
      fn(i);
    }
  }
  // runs f(0..n-1) on the pool + the calling thread; returns when all are done
  void run(int n, const std::function<void(int)> &f) {
    if (n <= 0) return;
    if (th.empty() || n == 1) {
      for (int i = 0; i < n; i++) f(i);
      return;
    }
    {
      std::unique_lock<std::mutex> l(mu);
      fn = f;
      nTasks = n;
      next.store(0);
      pending = (int)th.size();
      generation++;
    }
    cvWork.notify_all();
    work();
    std::unique_lock<std::mutex> l(mu);
    cvDone.wait(l, [&]() { return pending == 0; });
  }
};


// Copy into a pinned staging buffer with NON-TEMPORAL stores.  A buffer that host threads have just written with ordinary stores
// sits dirty in their caches, and the DMA engine then reads it at 9-14 GB/s instead of 45-52 GB/s (measured on the pool's B200
// boxes, scripts/probe/h2d_probe.cu: 9.8 MB written by 8 threads, then cudaMemcpyAsync); streaming stores leave nothing in the
// caches to snoop.  dst is 16-byte aligned (cudaMallocHost, frame sizes are multiples of 256).
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
static void stage_copy(uint8_t *dst, const uint8_t *src, size_t n) {
  size_t i = 0;
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
      _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
  }
  if (i < n) std::memcpy(dst + i, src + i, n - i);
  _mm_sfence();  // the streamed lines are globally visible before the copy engine is started
#else
  std::memcpy(dst, src, n);
#endif
}

namespace lsd {

static HostPool *host_pool(lsd_ctx *ctx) {
  if (!ctx->pool) {
    static const int envT = getenv("LSD_B200_STAGE_THREADS") ? atoi(getenv("LSD_B200_STAGE_THREADS")) : 0;
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = envT > 0 ? envT : (int)(hw / 2 < 1 ? 1 : (hw / 2 > 8 ? 8 : hw / 2));
    ctx->pool = new HostPool(T - 1);  // the calling thread is the T-th worker
  }
  return ctx->pool;
}

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void make_intrinsics(int w, int h, const float K[4], Intrinsics &I) {
  for (int l = 0; l < NL; l++) {
    I.w[l] = w >> l;
    I.h[l] = h >> l;
    if (l == 0) {
      I.fx[0] = K[0]; I.fy[0] = K[1]; I.cx[0] = K[2]; I.cy[0] = K[3];
    } else {
      // Frame::initialize: fx_l = fx_{l-1} * 0.5; cx_l = (cx_0 + 0.5) / 2^l - 0.5 (double literals, rounded on store)
      I.fx[l] = (float)(I.fx[l - 1] * 0.5);
      I.fy[l] = (float)(I.fy[l - 1] * 0.5);
      I.cx[l] = (float)((I.cx[0] + 0.5) / ((int)1 << l) - 0.5);
      I.cy[l] = (float)((I.cy[0] + 0.5) / ((int)1 << l) - 0.5);
    }
    I.fxi[l] = 1.0f / I.fx[l];
    I.fyi[l] = 1.0f / I.fy[l];
    I.cxi[l] = -I.cx[l] / I.fx[l];
    I.cyi[l] = -I.cy[l] / I.fy[l];
  }
}

static void make_layout(const Intrinsics &I, FrameLayout &L) {
  size_t off = 0;
  for (int l = 0; l < NL; l++) { L.img[l] = off; off = align_up(off + (size_t)I.w[l] * I.h[l] * 4, 256); }
  for (int l = 0; l < NL; l++) { L.grad[l] = off; off = align_up(off + (size_t)I.w[l] * I.h[l] * 16, 256); }
  L.maxgrad = off; off = align_up(off + (size_t)I.w[0] * I.h[0] * 4, 256);
  for (int l = 0; l < NL; l++) { L.idepth[l] = off; off = align_up(off + (size_t)I.w[l] * I.h[l] * 4, 256); }
  for (int l = 0; l < NL; l++) { L.idvar[l] = off; off = align_up(off + (size_t)I.w[l] * I.h[l] * 4, 256); }
  L.mask = off; off = align_up(off + (size_t)I.w[1] * I.h[1], 256);
  off += 256;  // tail: counters (numMappablePixels at total-16)
  L.total = off;
}

int ensure_stage(lsd_ctx *ctx, size_t hostBytes, size_t devBytes) {
  if (hostBytes > ctx->h_stageBytes) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stageBytes = 0;
    LSD_CUDA(cudaMallocHost(&ctx->h_stage, hostBytes));
    ctx->h_stageBytes = hostBytes;
  }
  if (devBytes > ctx->d_stageBytes) {
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->d_stageBytes = 0;
    LSD_CUDA(cudaMalloc(&ctx->d_stage, devBytes));
    ctx->d_stageBytes = devBytes;
  }
  return LSD_OK;
}

int ctx_finish_pending(lsd_ctx *ctx) {
  if (!ctx->pendingSync) return LSD_OK;
  ctx->pendingSync = false;
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  resolve_pending_means(ctx);
  return LSD_OK;
}

int ensure_table(lsd_ctx *ctx, size_t bytes) {
  if (ctx->pendingSync) {  // every writer of the pinned table passes through here first
    int rc = ctx_finish_pending(ctx);
    if (rc) return rc;
  }
  if (bytes > ctx->tableBytes) {
    if (ctx->h_table) cudaFreeHost(ctx->h_table);
    if (ctx->d_table) cudaFree(ctx->d_table);
    ctx->h_table = ctx->d_table = nullptr;
    ctx->tableBytes = 0;
    size_t nb = align_up(bytes < 4096 ? 4096 : bytes * 2, 4096);
    LSD_CUDA(cudaMallocHost(&ctx->h_table, nb));
    LSD_CUDA(cudaMalloc(&ctx->d_table, nb));
    ctx->tableBytes = nb;
  }
  return LSD_OK;
}

#define LSD_PTR_TABLE_CAP 8192
int ptr_table_register(lsd_ctx *ctx, const void *p) {
  if (!ctx->d_ptrTable) {
    LSD_CUDA(cudaMalloc(&ctx->d_ptrTable, sizeof(void *) * LSD_PTR_TABLE_CAP));
    LSD_CUDA(cudaMallocHost(&ctx->h_ptrTable, sizeof(void *) * LSD_PTR_TABLE_CAP));
  }
  if (ctx->ptrTableCount >= LSD_PTR_TABLE_CAP) return LSD_OK;  // such a slab simply keeps using uploaded lists
  const int idx = ctx->ptrTableCount++;
  // Once per slab allocation, never on the per-frame path (slabs are pooled).  Queued on the context's stream out of a pinned
  // mirror that is never rewritten: every consumer of the entry runs on that stream, after the copy.  (A blocking cudaMemcpy from
  // pageable memory may return before its DMA has landed, and nothing would order it against the context's non-blocking stream.)
  ctx->h_ptrTable[idx] = const_cast<void *>(p);
  LSD_CUDA(cudaMemcpyAsync(ctx->d_ptrTable + idx, ctx->h_ptrTable + idx, sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
  ctx->ptrIndex[p] = idx;
  return LSD_OK;
}
void *const *ptr_table_entry(const lsd_ctx *ctx, const void *p) {
  const auto it = ctx->ptrIndex.find(p);
  return it == ctx->ptrIndex.end() ? nullptr : ctx->d_ptrTable + it->second;
}

static int alloc_frame_slab(lsd_ctx *ctx, uint8_t **out) {
  if (!ctx->frameSlabPool.empty()) {
    *out = ctx->frameSlabPool.back();
    ctx->frameSlabPool.pop_back();
    return LSD_OK;
  }
  LSD_CUDA(cudaMalloc(out, ctx->lay.total));
  return ptr_table_register(ctx, *out);
}

static lsd_frame *new_frame(int id, uint8_t *slab) {
  lsd_frame *f = new lsd_frame();
  std::memset(f, 0, sizeof(*f));
  f->id = id;
  f->slab = slab;
  f->numMappable = -1;
  f->trackingParentId = -1;
  f->thisToParent_raw[3] = 1.0;
  f->thisToParent_raw[7] = 1.0;
  f->meanIdepth = 1.0f;
  return f;
}

// upload a pointer list to d_table (blocking: h_table is reused)
static int upload_ptrs(lsd_ctx *ctx, const std::vector<void *> &v, cudaStream_t st, size_t tableOffset = 0) {
  int rc = ensure_table(ctx, tableOffset + v.size() * sizeof(void *));
  if (rc) return rc;
  std::memcpy((char *)ctx->h_table + tableOffset, v.data(), v.size() * sizeof(void *));
  LSD_CUDA(cudaMemcpyAsync((char *)ctx->d_table + tableOffset, (char *)ctx->h_table + tableOffset, v.size() * sizeof(void *),
                           cudaMemcpyHostToDevice, st));
  return LSD_OK;
}

// builds the requested planes of n frames whose slabs are listed in d_slabs[0..n)
static void build_planes(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, unsigned flags, cudaStream_t st) {
  launch_gradients(ctx, d_slabs, n, (flags & LSD_BUILD_GRAD0) ? 0 : 1, NL - 1, st, true);  // + the refPixelWasGood plane (FB_MASK)
  if (flags & LSD_BUILD_MAXGRAD0) launch_maxgrad0(ctx, d_slabs, n, st);
}

// Frame::setDepth's bookkeeping (meanIdepth, numPoints) is produced by the setDepth / pyramid launch itself (k_idepth_pyramid):
// prepare_mean_idepth sizes the buffers before that launch, schedule_mean_idepth queues the read-back behind it.
int prepare_mean_idepth(lsd_ctx *ctx, int n, float **d_out2) {
  if (ctx->pendingSync) {
    int rc = ctx_finish_pending(ctx);
    if (rc) return rc;
  }
  ctx->pendingMeans.clear();  // leftovers of a call that failed before its synchronisation
  if (n > ctx->meansCap) {
    if (ctx->d_means) cudaFree(ctx->d_means);
    if (ctx->h_means) cudaFreeHost(ctx->h_means);
    ctx->d_means = ctx->h_means = nullptr;
    ctx->meansCap = 0;
    const int cap = n < 32 ? 32 : 2 * n;
    LSD_CUDA(cudaMalloc(&ctx->d_means, 8 * (size_t)cap));
    LSD_CUDA(cudaMallocHost(&ctx->h_means, 8 * (size_t)cap));
    ctx->meansCap = cap;
  }
  int rc = ensure_stats_scratch(ctx, n);
  if (rc) return rc;
  *d_out2 = ctx->d_means;
  return LSD_OK;
}

int schedule_mean_idepth(lsd_ctx *ctx, int n, lsd_frame *const *frames, cudaStream_t st) {
  LSD_CUDA(cudaMemcpyAsync(ctx->h_means, ctx->d_means, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
  ctx->pendingMeans.assign(frames, frames + n);
  return LSD_OK;
}

void resolve_pending_means(lsd_ctx *ctx) {
  for (size_t i = 0; i < ctx->pendingMeans.size(); i++) {
    lsd_frame *f = ctx->pendingMeans[i];
    f->meanIdepth = ctx->h_means[2 * i];
    std::memcpy(&f->numPoints, &ctx->h_means[2 * i + 1], 4);
    f->meanValid = true;
  }
  ctx->pendingMeans.clear();
}

int frame_ensure_built(lsd_ctx *ctx, lsd_frame *f, unsigned need) {
  const unsigned missing = need & ~f->built;
  if (!missing) return LSD_OK;
  cudaStream_t st = ctx->stream;
  std::vector<void *> v{f->slab};
  int rc = upload_ptrs(ctx, v, st);
  if (rc) return rc;
  uint8_t *const *d_slabs = reinterpret_cast<uint8_t *const *>(ctx->d_table);
  if (missing & FB_GRAD0) {
    launch_gradients(ctx, d_slabs, 1, 0, 0, st);
    f->built |= FB_GRAD0;
  }
  if (missing & FB_MAXGRAD0) {
    launch_maxgrad0(ctx, d_slabs, 1, st);
    f->built |= FB_MAXGRAD0;
  }
  if (missing & FB_IDEPTH_PYR) {
    if (!(f->built & FB_IDEPTH0)) {
      set_error("frame has no depth (hasIDepthBeenSet() == false)");
      return LSD_ERR_STATE;
    }
    launch_idepth_pyramid(ctx, d_slabs, 1, st);
    f->built |= FB_IDEPTH_PYR;
  }
  if (missing & FB_MASK) {
    launch_mask_init(ctx, d_slabs, 1, st);
    f->built |= FB_MASK;
  }
  LSD_CUDA(cudaStreamSynchronize(st));
  return LSD_OK;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

const char *lsd_last_error(void) { return g_err.c_str(); }
int lsd_version(void) { return 100; }

int lsd_default_tracker_settings(lsd_tracker_settings *s) {
  LSD_ARG(s);
  s->lambdaSuccessFac = 0.5f;
  s->lambdaFailFac = 2.0f;
  const int its[NL] = {5, 20, 50, 100, 100};
  for (int l = 0; l < NL; l++) {
    s->stepSizeMin[l] = 1e-8f;
    s->convergenceEps[l] = 0.999f;
    s->maxItsPerLvl[l] = its[l];
    s->lambdaInitial[l] = 0;
  }
  s->var_weight = 1.0f;
  s->huber_d = 3.0f;
  return LSD_OK;
}

// the "test track" members of [UP] DenseDepthTrackerSettings: maxItsTestTrack 5, stepSizeMinTestTrack 1e-3,
// convergenceEpsTestTrack 0.98, lambdaInitialTestTrack 0 -- installed at every level (only QUICK_KF_CHECK_LVL is used)
int lsd_default_permaref_settings(lsd_tracker_settings *s) {
  int rc = lsd_default_tracker_settings(s);
  if (rc) return rc;
  for (int l = 0; l < NL; l++) {
    s->stepSizeMin[l] = 1e-3f;
    s->convergenceEps[l] = 0.98f;
    s->maxItsPerLvl[l] = 5;
    s->lambdaInitial[l] = 0;
  }
  return LSD_OK;
}

int lsd_ctx_create(int device, int width, int height, const float K[4], void *stream, lsd_ctx **out) {
  LSD_ARG(out && K);
  LSD_ARG(width > 0 && height > 0 && width % 16 == 0 && height % 16 == 0);
  LSD_ARG(width <= 65535 && height <= 65535);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (liblsd_b200 has no CPU fallback)");
    return LSD_ERR_CUDA;
  }
  LSD_ARG(device >= 0 && device < ndev);
  LSD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LSD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("liblsd_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    return LSD_ERR_CUDA;
  }
  lsd_ctx *ctx = new lsd_ctx();
  ctx->device = device;
  ctx->numSMs = prop.multiProcessorCount;
  ctx->w = width;
  ctx->h = height;
  make_intrinsics(width, height, K, ctx->K);
  make_layout(ctx->K, ctx->lay);
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
    ctx->ownStream = false;
  } else {
    LSD_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->ownStream = true;
  }
  LSD_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  LSD_CUDA(cudaStreamCreateWithFlags(&ctx->trackStream, cudaStreamNonBlocking));
  LSD_CUDA(cudaEventCreate(&ctx->evA));
  LSD_CUDA(cudaEventCreate(&ctx->evB));
  for (int i = 0; i < 4; i++) LSD_CUDA(cudaEventCreateWithFlags(&ctx->evPipe[i], cudaEventDisableTiming));
  ctx->launches = 0;
  lsd_default_tracker_settings(&ctx->se3);
  lsd_default_tracker_settings(&ctx->sim3);
  lsd_default_permaref_settings(&ctx->permaref);
  ctx->se3Permaref = false;
  ctx->se3RecsPerItem = 0;
  ctx->se3RecordPoints = 0;
  ctx->deferSync = ctx->pendingSync = false;
  for (int &v : ctx->se3RecordPointsLvl) v = 0;
  ctx->se3LivePairs = std::getenv("LSD_B200_SE3_LIVE_PAIRS") ? std::atoi(std::getenv("LSD_B200_SE3_LIVE_PAIRS")) : -1;
  ctx->tmaUnavailable = false;
  {
    const char *e = std::getenv("LSD_B200_STENCIL_TMA");
    ctx->stencilTma = e ? (std::atoi(e) & 3) : LSD_STENCIL_TMA_DEFAULT;
  }
  ctx->imageChunk = 0;
  ctx->imageStreamed = -1;
  ctx->streamWatchdogNs = 2000000000ull;
  ctx->d_stats = nullptr;
  ctx->d_ptrTable = nullptr;
  ctx->h_ptrTable = nullptr;
  ctx->ptrTableCount = 0;
  ctx->statsFrames = 0;
  ctx->d_means = ctx->h_means = nullptr;
  ctx->meansCap = 0;
  {
    const int rc = ensure_stats_scratch(ctx, 1);
    if (rc) return rc;
  }
  ctx->se3ActivePairs = 0;
  ctx->refSlabBytes = 0;
  ctx->h_stage = ctx->d_stage = nullptr;
  ctx->h_stageBytes = ctx->d_stageBytes = 0;
  ctx->h_table = ctx->d_table = nullptr;
  ctx->tableBytes = 0;
  ctx->descSlot = 0;
  ctx->stageTimed = false;
  ctx->pool = nullptr;
  ctx->se3s = nullptr;
  ctx->sim3s = nullptr;
  ctx->sim3RecordPoints = 0;
  ctx->lastAlgBytes = 0;
  ctx->lastEvals = 0;
  ctx->lastKernelMs = 0;
  *out = ctx;
  return LSD_OK;
}

int lsd_ctx_destroy(lsd_ctx *ctx) {
  if (!ctx) return LSD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  se3_scratch_free(ctx);
  sim3_scratch_free(ctx);
  delete ctx->pool;
  if (ctx->d_ptrTable) cudaFree(ctx->d_ptrTable);
  if (ctx->h_ptrTable) cudaFreeHost(ctx->h_ptrTable);
  for (auto p : ctx->frameSlabPool) cudaFree(p);
  for (auto p : ctx->refSlabPool) cudaFree(p);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->h_table) cudaFreeHost(ctx->h_table);
  if (ctx->d_table) cudaFree(ctx->d_table);
  if (ctx->d_stats) cudaFree(ctx->d_stats);
  if (ctx->d_means) cudaFree(ctx->d_means);
  if (ctx->h_means) cudaFreeHost(ctx->h_means);
  cudaEventDestroy(ctx->evA);
  cudaEventDestroy(ctx->evB);
  for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->evPipe[i]);
  cudaStreamDestroy(ctx->copyStream);
  cudaStreamDestroy(ctx->trackStream);
  if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LSD_OK;
}

int lsd_ctx_synchronize(lsd_ctx *ctx) {
  LSD_ARG(ctx);
  if (ctx->pendingSync) return ctx_finish_pending(ctx);
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}
void *lsd_ctx_stream(lsd_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
long long lsd_ctx_launch_count(lsd_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lsd_ctx_set_se3_settings(lsd_ctx *ctx, const lsd_tracker_settings *s) {
  LSD_ARG(ctx && s);
  ctx->se3 = *s;
  return LSD_OK;
}

int lsd_ctx_set_se3_work_item_records(lsd_ctx *ctx, int records) {
  LSD_ARG(ctx && records >= 0 && records <= 64);
  ctx->se3RecsPerItem = records;
  return LSD_OK;
}

int lsd_ctx_set_se3_record_points(lsd_ctx *ctx, int points) {
  LSD_ARG(ctx);
  LSD_ARG(points == 0 || (points >= 128 && points % 128 == 0 && points <= (1 << 20)));
  LSD_ARG(points == 0 || (ctx->K.w[1] * ctx->K.h[1] + points - 1) / points <= 4096);  // work-item codes carry 12 bits of record index
  ctx->se3RecordPoints = points;
  return LSD_OK;
}

int lsd_ctx_set_se3_record_points_per_level(lsd_ctx *ctx, const int *points) {
  LSD_ARG(ctx && points);
  for (int l = 1; l < NL; l++) {
    const int p = points[l];
    LSD_ARG(p == 0 || (p >= 128 && p % 128 == 0 && p <= (1 << 20)));
    LSD_ARG(p == 0 || (ctx->K.w[l] * ctx->K.h[l] + p - 1) / p <= 4096);  // work-item codes carry 12 bits of record index
  }
  for (int l = 0; l < NL; l++) ctx->se3RecordPointsLvl[l] = l ? points[l] : 0;
  return LSD_OK;
}

int lsd_ctx_set_live_tracking(lsd_ctx *ctx, int enable) {
  LSD_ARG(ctx);
  if (!enable) {
    for (int &v : ctx->se3RecordPointsLvl) v = 0;
    return LSD_OK;
  }
  // one record per 128-thread group of a 16-CTA cluster (64 groups) while at most 60 % of a level's pixels carry depth (a
  // semi-dense keyframe of the bench scenes: 45 %); a denser level simply takes a second round
  for (int l = 1; l < NL; l++) {
    const long long px = (long long)ctx->K.w[l] * ctx->K.h[l];
    long long p = (px * 60 / 100 + 63) / 64;
    p = (p + 127) / 128 * 128;
    if (p < 128) p = 128;
    while ((px + p - 1) / p > 4096) p += 128;
    ctx->se3RecordPointsLvl[l] = (int)p;
  }
  ctx->se3RecordPointsLvl[0] = 0;
  return LSD_OK;
}

int lsd_ctx_set_se3_live_pairs(lsd_ctx *ctx, int pairs) {
  LSD_ARG(ctx && pairs >= -1);
  ctx->se3LivePairs = pairs;
  return LSD_OK;
}

int lsd_ctx_set_stencil_tma(lsd_ctx *ctx, int enable) {  // `enable`: the mask of include/lsd_b200.h
  LSD_ARG(ctx);
  ctx->stencilTma = ctx->tmaUnavailable ? 0 : (enable & 3);
  return LSD_OK;
}

int lsd_ctx_set_se3_active_pairs(lsd_ctx *ctx, int pairs) {
  LSD_ARG(ctx && pairs >= 0);
  ctx->se3ActivePairs = pairs;
  return LSD_OK;
}

// ---- frames ---------------------------------------------------------------------------------

int lsd_frame_create_batch_device(lsd_ctx *ctx, int n, const int *ids, const void *d_images, unsigned flags, lsd_frame **out) {
  LSD_ARG(ctx && out && d_images && n >= 0);
  if (n == 0) return LSD_OK;
  LSD_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::vector<void *> slabs(n);
  for (int i = 0; i < n; i++) {
    uint8_t *s = nullptr;
    int rc = alloc_frame_slab(ctx, &s);
    if (rc) return rc;
    slabs[i] = s;
    out[i] = new_frame(ids ? ids[i] : i, s);
  }
  static const bool trace = std::getenv("LSD_B200_TRACE") != nullptr;
  static int traced = 0;
  static cudaEvent_t tev[3];
  static bool tevMade = false;
  if (trace && !tevMade) {
    for (int i = 0; i < 3; i++) cudaEventCreate(&tev[i]);
    tevMade = true;
  }
  const auto t0 = std::chrono::steady_clock::now();
  // one frame: its slab's entry in the device-resident pointer table; several: an uploaded list
  uint8_t *const *d_slabs = n == 1 ? reinterpret_cast<uint8_t *const *>(ptr_table_entry(ctx, slabs[0])) : nullptr;
  if (!d_slabs) {
    int rc = upload_ptrs(ctx, slabs, st);
    if (rc) return rc;
    d_slabs = reinterpret_cast<uint8_t *const *>(ctx->d_table);
  }
  if (trace) cudaEventRecord(tev[0], st);
  launch_ingest(ctx, (const uint8_t *)d_images, ctx->w, (size_t)ctx->w * ctx->h, d_slabs, n, st);
  if (trace) cudaEventRecord(tev[1], st);
  build_planes(ctx, d_slabs, n, flags, st);
  if (trace) cudaEventRecord(tev[2], st);
  const auto t1 = std::chrono::steady_clock::now();
  if (!ctx->deferSync) LSD_CUDA(cudaStreamSynchronize(st));  // pipelined driver: the planes are consumed in stream order by the tracker
  if (trace && n > 1 && traced++ < 12) {
    const auto t2 = std::chrono::steady_clock::now();
    float msI = 0, msG = 0;
    cudaEventElapsedTime(&msI, tev[0], tev[1]);
    cudaEventElapsedTime(&msG, tev[1], tev[2]);
    std::fprintf(stderr, "[lsd_b200] frame_create_batch_device n=%d: enqueue %.1f us, wait %.1f us; device: ingest %.1f us, planes %.1f us\n", n,
                 1e6 * std::chrono::duration<double>(t1 - t0).count(), 1e6 * std::chrono::duration<double>(t2 - t1).count(), 1e3 * msI, 1e3 * msG);
  }
  for (int i = 0; i < n; i++)
    out[i]->built = FB_TRACKING | FB_MASK | ((flags & LSD_BUILD_MAXGRAD0) ? FB_MAXGRAD0 : 0) | ((flags & LSD_BUILD_GRAD0) ? FB_GRAD0 : 0);
  return LSD_OK;
}

int lsd_frame_create_batch(lsd_ctx *ctx, int n, const int *ids, const uint8_t *const *images, size_t pitch, unsigned flags,
                           lsd_frame **out) {
  LSD_ARG(ctx && out && images && n >= 0);
  LSD_ARG(pitch >= (size_t)ctx->w);
  if (n == 0) return LSD_OK;
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t fbytes = (size_t)ctx->w * ctx->h;
  // stage in chunks of <= 64 frames so pinned/device staging stays small
  const int CH = 64;
  int rc = ensure_stage(ctx, fbytes * (n < CH ? n : CH), fbytes * (n < CH ? n : CH));
  if (rc) return rc;
  for (int i = 0; i < n; i++) LSD_ARG(images[i]);
  static const bool trace = std::getenv("LSD_B200_TRACE") != nullptr;  // host-side timing of the ingest steps on stderr
  static int traced = 0, traced1 = 0;
  // the pinned staging buffer is reused chunk by chunk: only a single-chunk call may leave its copy unsynchronised
  struct DeferRestore {
    lsd_ctx *ctx;
    bool saved;
    ~DeferRestore() { ctx->deferSync = saved; }
  } deferRestore = {ctx, ctx->deferSync};
  if (n > CH) ctx->deferSync = false;
  for (int i0 = 0; i0 < n; i0 += CH) {
    const int m = (n - i0) < CH ? (n - i0) : CH;
    const auto t0 = std::chrono::steady_clock::now();
    // staged and copied in up to four sub-chunks: the copy engine moves sub-chunk k while the host threads stage k + 1
    const int SUB = m >= 8 ? (m + 3) / 4 : m;
    static const int splitOne = getenv("LSD_B200_STAGE_SPLIT") ? atoi(getenv("LSD_B200_STAGE_SPLIT")) : 1;  // r02v: 55 us (4 parts) vs 57 us (1): not worth waking the pool
    if (m == 1 && splitOne > 1 && ctx->h >= 4 * splitOne) {
      // one live frame: its rows are staged by a few pool threads at once (32 us of single-thread copy for a VGA frame otherwise)
      const uint8_t *src = images[i0];
      uint8_t *dst = ctx->h_stage;
      const int rowsPer = (ctx->h + splitOne - 1) / splitOne;
      host_pool(ctx)->run(splitOne, [&](int part) {
        const int y0 = part * rowsPer, y1 = (y0 + rowsPer) < ctx->h ? (y0 + rowsPer) : ctx->h;
        if (pitch == (size_t)ctx->w) {
          if (y1 > y0) stage_copy(dst + (size_t)y0 * ctx->w, src + (size_t)y0 * ctx->w, (size_t)(y1 - y0) * ctx->w);
        } else {
          for (int y = y0; y < y1; y++) stage_copy(dst + (size_t)y * ctx->w, src + (size_t)y * pitch, ctx->w);
        }
      });
      LSD_CUDA(cudaMemcpyAsync(ctx->d_stage, ctx->h_stage, fbytes, cudaMemcpyHostToDevice, ctx->stream));
    } else
    for (int j0 = 0; j0 < m; j0 += SUB) {
      const int mm = (m - j0) < SUB ? (m - j0) : SUB;
      host_pool(ctx)->run(mm, [&](int i) {
        const uint8_t *src = images[i0 + j0 + i];
        uint8_t *dst = ctx->h_stage + fbytes * (j0 + i);
        if (pitch == (size_t)ctx->w) stage_copy(dst, src, fbytes);
        else for (int y = 0; y < ctx->h; y++) stage_copy(dst + (size_t)y * ctx->w, src + (size_t)y * pitch, ctx->w);
      });
      LSD_CUDA(cudaMemcpyAsync(ctx->d_stage + fbytes * j0, ctx->h_stage + fbytes * j0, fbytes * mm, cudaMemcpyHostToDevice, ctx->stream));
    }
    const auto t1 = std::chrono::steady_clock::now();
    rc = lsd_frame_create_batch_device(ctx, m, ids ? ids + i0 : nullptr, ctx->d_stage, flags, out + i0);
    if (rc) return rc;
    if (!ids) for (int i = 0; i < m; i++) out[i0 + i]->id = i0 + i;
    if (trace && (m > 1 ? traced++ < 12 : traced1++ < 3)) {
      const auto t2 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[lsd_b200] frame_create_batch n=%d: stage %.1f us (%d pool threads), h2d + kernels + sync %.1f us\n", m,
                   1e6 * std::chrono::duration<double>(t1 - t0).count(), (int)host_pool(ctx)->th.size() + 1,
                   1e6 * std::chrono::duration<double>(t2 - t1).count());
    }
  }
  return LSD_OK;
}

int lsd_frame_create(lsd_ctx *ctx, int id, const uint8_t *image, size_t pitch, unsigned flags, lsd_frame **out) {
  const uint8_t *imgs[1] = {image};
  return lsd_frame_create_batch(ctx, 1, &id, imgs, pitch, flags, out);
}

int lsd_frame_release(lsd_ctx *ctx, lsd_frame *f) {
  LSD_ARG(ctx);
  if (!f) return LSD_OK;
  ctx->frameSlabPool.push_back(f->slab);
  delete f;
  return LSD_OK;
}

int lsd_frame_release_batch(lsd_ctx *ctx, int n, lsd_frame **f) {
  LSD_ARG(ctx && (f || n == 0));
  for (int i = 0; i < n; i++) {
    lsd_frame_release(ctx, f[i]);
    f[i] = nullptr;
  }
  return LSD_OK;
}

int lsd_frame_read(lsd_ctx *ctx, lsd_frame *f, int field, int level, void *dst) {
  LSD_ARG(ctx && f && dst);
  LSD_ARG(level >= 0 && level < NL);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t N = (size_t)ctx->K.w[level] * ctx->K.h[level];
  const FrameLayout &L = ctx->lay;
  const uint8_t *src = nullptr;
  size_t bytes = 0;
  int rc = LSD_OK;
  switch (field) {
    case LSD_FIELD_IMAGE: src = f->slab + L.img[level]; bytes = N * 4; break;
    case LSD_FIELD_GRADIENTS:
      if (level == 0) rc = frame_ensure_built(ctx, f, FB_GRAD0);
      src = f->slab + L.grad[level]; bytes = N * 16; break;
    case LSD_FIELD_MAXGRAD:
      LSD_ARG(level == 0);
      rc = frame_ensure_built(ctx, f, FB_MAXGRAD0);
      src = f->slab + L.maxgrad; bytes = N * 4; break;
    case LSD_FIELD_IDEPTH:
    case LSD_FIELD_IDEPTHVAR:
      if (!(f->built & FB_IDEPTH0)) { set_error("frame has no depth"); return LSD_ERR_STATE; }
      if (level > 0) rc = frame_ensure_built(ctx, f, FB_IDEPTH_PYR);
      src = f->slab + (field == LSD_FIELD_IDEPTH ? L.idepth[level] : L.idvar[level]); bytes = N * 4; break;
    case LSD_FIELD_MASK:
      rc = frame_ensure_built(ctx, f, FB_MASK);
      src = f->slab + L.mask; bytes = (size_t)ctx->K.w[1] * ctx->K.h[1]; break;
    default: LSD_ARG(!"unknown field");
  }
  if (rc) return rc;
  LSD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}

int lsd_frame_num_mappable_pixels(lsd_ctx *ctx, lsd_frame *f, int *out) {
  LSD_ARG(ctx && f && out);
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = frame_ensure_built(ctx, f, FB_MAXGRAD0);
  if (rc) return rc;
  if (f->numMappable < 0) {
    LSD_CUDA(cudaMemcpyAsync(&f->numMappable, f->slab + ctx->lay.total - 16, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *out = f->numMappable;
  return LSD_OK;
}

int lsd_frame_set_depth_from_gt(lsd_ctx *ctx, lsd_frame *f, const float *depth, float cov_scale) {
  LSD_ARG(ctx && f && depth);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->w * ctx->h * 4;
  int rc = ensure_stage(ctx, bytes, bytes);
  if (rc) return rc;
  std::memcpy(ctx->h_stage, depth, bytes);
  LSD_CUDA(cudaMemcpyAsync(ctx->d_stage, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
  launch_set_depth_gt(ctx, f->slab, reinterpret_cast<const float *>(ctx->d_stage), cov_scale, ctx->stream);
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  f->built = (f->built | FB_IDEPTH0) & ~FB_IDEPTH_PYR;
  f->meanValid = false;
  return LSD_OK;
}

int lsd_frame_set_idepth(lsd_ctx *ctx, lsd_frame *f, const float *idepth, const float *idepthVar) {
  LSD_ARG(ctx && f && idepth && idepthVar);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->w * ctx->h * 4;
  int rc = ensure_stage(ctx, 2 * bytes, 0);
  if (rc) return rc;
  std::memcpy(ctx->h_stage, idepth, bytes);
  std::memcpy(ctx->h_stage + bytes, idepthVar, bytes);
  LSD_CUDA(cudaMemcpyAsync(f->slab + ctx->lay.idepth[0], ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LSD_CUDA(cudaMemcpyAsync(f->slab + ctx->lay.idvar[0], ctx->h_stage + bytes, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  f->built = (f->built | FB_IDEPTH0) & ~FB_IDEPTH_PYR;
  f->meanValid = false;
  return LSD_OK;
}

int lsd_frame_set_idepth_batch_device(lsd_ctx *ctx, int n, lsd_frame *const *f, const void *d_idepth, const void *d_idepthVar) {
  LSD_ARG(ctx && f && d_idepth && d_idepthVar);
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->w * ctx->h * 4;
  for (int i = 0; i < n; i++) {
    LSD_CUDA(cudaMemcpyAsync(f[i]->slab + ctx->lay.idepth[0], (const char *)d_idepth + bytes * i, bytes, cudaMemcpyDeviceToDevice,
                             ctx->stream));
    LSD_CUDA(cudaMemcpyAsync(f[i]->slab + ctx->lay.idvar[0], (const char *)d_idepthVar + bytes * i, bytes, cudaMemcpyDeviceToDevice,
                             ctx->stream));
    f[i]->built = (f[i]->built | FB_IDEPTH0) & ~FB_IDEPTH_PYR;
    f[i]->meanValid = false;
  }
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSD_OK;
}

int lsd_frame_mean_idepth(lsd_ctx *ctx, lsd_frame *f, float *meanIdepth, int *numPoints) {
  LSD_ARG(ctx && f);
  if (!(f->built & FB_IDEPTH0)) { set_error("frame has no depth"); return LSD_ERR_STATE; }
  if (!f->meanValid && ctx->pendingSync) {  // the deferred updateKeyframe may carry exactly this value
    int rc0 = ctx_finish_pending(ctx);
    if (rc0) return rc0;
  }
  if (f->meanValid) {  // computed alongside the setDepth that produced the planes
    if (meanIdepth) *meanIdepth = f->meanIdepth;
    if (numPoints) *numPoints = f->numPoints;
    return LSD_OK;
  }
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_table(ctx, 64);
  if (rc) return rc;
  launch_idepth_stats(ctx, f->slab, reinterpret_cast<float *>(ctx->d_table), ctx->stream);
  float h[2];
  LSD_CUDA(cudaMemcpyAsync(h, ctx->d_table, 8, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  f->meanIdepth = h[0];
  std::memcpy(&f->numPoints, &h[1], 4);
  f->meanValid = true;
  if (meanIdepth) *meanIdepth = f->meanIdepth;
  if (numPoints) *numPoints = f->numPoints;
  return LSD_OK;
}

int lsd_frame_mean_idepth_batch(lsd_ctx *ctx, int n, lsd_frame *const *f, float *meanIdepth, int *numPoints) {
  LSD_ARG(ctx && f && n >= 0);
  if (n == 0) return LSD_OK;
  LSD_CUDA(cudaSetDevice(ctx->device));
  bool allCached = true;
  for (int i = 0; i < n; i++) {
    LSD_ARG(f[i]);
    if (!(f[i]->built & FB_IDEPTH0)) { set_error("frame has no depth"); return LSD_ERR_STATE; }
    allCached = allCached && f[i]->meanValid;
  }
  if (allCached) {
    for (int i = 0; i < n; i++) {
      if (meanIdepth) meanIdepth[i] = f[i]->meanIdepth;
      if (numPoints) numPoints[i] = f[i]->numPoints;
    }
    return LSD_OK;
  }
  // table: n slab pointers, then n x 2 floats of results; one launch for all frames (blockIdx.y = frame)
  const size_t offOut = (sizeof(void *) * (size_t)n + 255) / 256 * 256;
  int rc = ensure_table(ctx, offOut + 8 * (size_t)n);
  if (rc) return rc;
  rc = ensure_stats_scratch(ctx, n);
  if (rc) return rc;
  void **hp = reinterpret_cast<void **>(ctx->h_table);
  for (int i = 0; i < n; i++) hp[i] = f[i]->slab;
  LSD_CUDA(cudaMemcpyAsync(ctx->d_table, ctx->h_table, sizeof(void *) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  launch_idepth_stats_batch(ctx, reinterpret_cast<uint8_t *const *>(ctx->d_table), n, reinterpret_cast<float *>((char *)ctx->d_table + offOut),
                            ctx->stream);
  LSD_CUDA(cudaMemcpyAsync((char *)ctx->h_table + offOut, (char *)ctx->d_table + offOut, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  const float *h = reinterpret_cast<const float *>((char *)ctx->h_table + offOut);
  for (int i = 0; i < n; i++) {
    f[i]->meanIdepth = h[2 * i];
    std::memcpy(&f[i]->numPoints, &h[2 * i + 1], 4);
    f[i]->meanValid = true;
    if (meanIdepth) meanIdepth[i] = f[i]->meanIdepth;
    if (numPoints) numPoints[i] = f[i]->numPoints;
  }
  return LSD_OK;
}

int lsd_frame_set_tracking_meta(lsd_ctx *ctx, lsd_frame *f, int parentId, const double toParent[8], float initialTrackedResidual) {
  LSD_ARG(ctx && f && toParent);
  f->trackingParentId = parentId;
  for (int k = 0; k < 8; k++) f->thisToParent_raw[k] = toParent[k];
  f->initialTrackedResidual = initialTrackedResidual;
  return LSD_OK;
}

int lsd_frame_get_tracking_meta(lsd_ctx *ctx, lsd_frame *f, int *parentId, double toParent[8], float *initialTrackedResidual) {
  LSD_ARG(ctx && f);
  if (parentId) *parentId = f->trackingParentId;
  if (toParent) for (int k = 0; k < 8; k++) toParent[k] = f->thisToParent_raw[k];
  if (initialTrackedResidual) *initialTrackedResidual = f->initialTrackedResidual;
  return LSD_OK;
}

int lsd_frame_set_mask(lsd_ctx *ctx, lsd_frame *f, const uint8_t *mask) {
  LSD_ARG(ctx && f);
  if (!mask) {
    f->built &= ~FB_MASK;
    return LSD_OK;
  }
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->K.w[1] * ctx->K.h[1];
  int rc = ensure_stage(ctx, bytes, 0);
  if (rc) return rc;
  std::memcpy(ctx->h_stage, mask, bytes);
  LSD_CUDA(cudaMemcpyAsync(f->slab + ctx->lay.mask, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  f->built |= FB_MASK;
  return LSD_OK;
}

int lsd_frame_set_counters(lsd_ctx *ctx, lsd_frame *f, int numFramesTrackedOnThis, int numMappedOnThis) {
  LSD_ARG(ctx && f);
  f->numFramesTrackedOnThis = numFramesTrackedOnThis;
  f->numMappedOnThis = numMappedOnThis;
  return LSD_OK;
}

int lsd_frame_get_counters(lsd_ctx *ctx, lsd_frame *f, int *numFramesTrackedOnThis, int *numMappedOnThis) {
  LSD_ARG(ctx && f);
  if (numFramesTrackedOnThis) *numFramesTrackedOnThis = f->numFramesTrackedOnThis;
  if (numMappedOnThis) *numMappedOnThis = f->numMappedOnThis;
  return LSD_OK;
}

int lsd_frame_set_depth_updated_flag(lsd_ctx *ctx, lsd_frame *f, int flag) {
  LSD_ARG(ctx && f);
  f->depthHasBeenUpdatedFlag = flag != 0;
  return LSD_OK;
}

int lsd_frame_get_depth_updated_flag(lsd_ctx *ctx, lsd_frame *f, int *flag) {
  LSD_ARG(ctx && f && flag);
  *flag = f->depthHasBeenUpdatedFlag ? 1 : 0;
  return LSD_OK;
}

// ---- tracking references ----------------------------------------------------------------------

int lsd_ref_create_batch(lsd_ctx *ctx, int n, lsd_frame *const *keyframes, lsd_ref **out) {
  LSD_ARG(ctx && keyframes && out && n >= 0);
  if (n == 0) return LSD_OK;
  LSD_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // slab layout: per level RefPoint[N_l] then float2[N_l], then int[NL] counters
  size_t offPts[NL], offGrad[NL], off = 0;
  for (int l = 0; l < NL; l++) {
    const size_t N = (l == 0) ? 0 : (size_t)ctx->K.w[l] * ctx->K.h[l];
    offPts[l] = off; off = align_up(off + N * sizeof(RefPoint), 256);
    offGrad[l] = off; off = align_up(off + N * sizeof(float2), 256);
  }
  const size_t offNum = off;
  const size_t numBytes = sizeof(int) * (size_t)pointcloud_state_words(ctx);  // numData + look-back state of k_make_pointcloud
  off += align_up(numBytes, 256);
  ctx->refSlabBytes = off;
  // frames that still need their idepth pyramid
  std::vector<void *> needPyr;
  for (int i = 0; i < n; i++) {
    LSD_ARG(keyframes[i]);
    if (!(keyframes[i]->built & FB_IDEPTH0)) { set_error("keyframe has no depth"); return LSD_ERR_STATE; }
    if (!(keyframes[i]->built & FB_IDEPTH_PYR)) needPyr.push_back(keyframes[i]->slab);
  }
  if (!needPyr.empty()) {
    // pipelined driver: offset 0 of the pinned table may still be read by the copy of the frame list queued just before
    const size_t pyrOff = ctx->deferSync ? 16384 : 0;
    int rc = upload_ptrs(ctx, needPyr, st, pyrOff);
    if (rc) return rc;
    launch_idepth_pyramid(ctx, reinterpret_cast<uint8_t *const *>((char *)ctx->d_table + pyrOff), (int)needPyr.size(), st);
    LSD_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) keyframes[i]->built |= FB_IDEPTH_PYR;
  }
  std::vector<void *> tab(3 * (size_t)n);
  for (int i = 0; i < n; i++) {
    uint8_t *slab = nullptr;
    if (!ctx->refSlabPool.empty()) {
      slab = ctx->refSlabPool.back();
      ctx->refSlabPool.pop_back();
    } else {
      LSD_CUDA(cudaMalloc(&slab, ctx->refSlabBytes));
      int rcp = ptr_table_register(ctx, slab);
      if (!rcp) rcp = ptr_table_register(ctx, slab + offNum);  // the reference's counters (lsd_ref::d_num)
      if (rcp) return rcp;
    }
    lsd_ref *r = new lsd_ref();
    r->keyframe = keyframes[i];
    r->frameID = keyframes[i]->id;
    r->slab = slab;
    for (int l = 0; l < NL; l++) { r->offPts[l] = offPts[l]; r->offGrad[l] = offGrad[l]; r->num[l] = 0; }
    r->d_num = reinterpret_cast<int *>(slab + offNum);
    r->numValid = false;
    out[i] = r;
    tab[i] = keyframes[i]->slab;
    tab[n + i] = slab;
    tab[2 * (size_t)n + i] = r->d_num;
  }
  // pipelined driver: the frame created just before may still be waiting for ITS pointer list at offset 0 of the pinned table
  void *const *eK = nullptr, *const *eR = nullptr, *const *eN = nullptr;
  if (n == 1) {  // one reference: the three pointers are entries of the device-resident pointer table
    eK = ptr_table_entry(ctx, tab[0]);
    eR = ptr_table_entry(ctx, tab[1]);
    eN = ptr_table_entry(ctx, tab[2]);
  }
  if (eK && eR && eN) {
    launch_make_pointcloud(ctx, reinterpret_cast<uint8_t *const *>(eK), reinterpret_cast<uint8_t *const *>(eR), reinterpret_cast<int *const *>(eN), 1,
                           offPts, offGrad, st);
  } else {
    const size_t tabOff = ctx->deferSync ? 8192 : 0;  // room for the pointers of 1024 frames in front
    int rc = upload_ptrs(ctx, tab, st, tabOff);
    if (rc) return rc;
    void **d = reinterpret_cast<void **>((char *)ctx->d_table + tabOff);
    launch_make_pointcloud(ctx, reinterpret_cast<uint8_t *const *>(d), reinterpret_cast<uint8_t *const *>(d + n),
                           reinterpret_cast<int *const *>(d + 2 * (size_t)n), n, offPts, offGrad, st);
  }
  if (!ctx->deferSync) LSD_CUDA(cudaStreamSynchronize(st));
  return LSD_OK;
}

int lsd_ref_create(lsd_ctx *ctx, lsd_frame *keyframe, lsd_ref **out) { return lsd_ref_create_batch(ctx, 1, &keyframe, out); }

int lsd_ref_release(lsd_ctx *ctx, lsd_ref *r) {
  LSD_ARG(ctx);
  if (!r) return LSD_OK;
  ctx->refSlabPool.push_back(r->slab);
  delete r;
  return LSD_OK;
}

static int ref_fetch_nums(lsd_ctx *ctx, lsd_ref *r) {
  if (r->numValid) return LSD_OK;
  LSD_CUDA(cudaMemcpyAsync(r->num, r->d_num, sizeof(int) * NL, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  r->numValid = true;
  return LSD_OK;
}

int lsd_ref_num_data(lsd_ctx *ctx, lsd_ref *r, int level, int *out) {
  LSD_ARG(ctx && r && out && level >= 1 && level < NL);
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = ref_fetch_nums(ctx, r);
  if (rc) return rc;
  *out = r->num[level];
  return LSD_OK;
}

int lsd_ref_read(lsd_ctx *ctx, lsd_ref *r, int level, float *pos, float *grad, float *colorAndVar, int *idx) {
  LSD_ARG(ctx && r && level >= 1 && level < NL);
  LSD_CUDA(cudaSetDevice(ctx->device));
  int rc = ref_fetch_nums(ctx, r);
  if (rc) return rc;
  const int n = r->num[level];
  std::vector<RefPoint> pts(n);
  std::vector<float2> g(n);
  LSD_CUDA(cudaMemcpyAsync(pts.data(), r->slab + r->offPts[level], sizeof(RefPoint) * n, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaMemcpyAsync(g.data(), r->slab + r->offGrad[level], sizeof(float2) * n, cudaMemcpyDeviceToHost, ctx->stream));
  LSD_CUDA(cudaStreamSynchronize(ctx->stream));
  const float fxi = ctx->K.fxi[level], fyi = ctx->K.fyi[level], cxi = ctx->K.cxi[level], cyi = ctx->K.cyi[level];
  const int W = ctx->K.w[level];
  for (int i = 0; i < n; i++) {
    const int x = pts[i].xy & 0xffff, y = pts[i].xy >> 16;
    if (pos) {  // same operations as the kernel (and upstream makePointCloud)
      volatile float inv = pts[i].invDepth;
      volatile float ax = fxi * x, ay = fyi * y;
      volatile float bx = ax + cxi, by = ay + cyi;
      pos[3 * i] = inv * bx;
      pos[3 * i + 1] = inv * by;
      pos[3 * i + 2] = inv * 1.0f;
    }
    if (grad) { grad[2 * i] = g[i].x; grad[2 * i + 1] = g[i].y; }
    if (colorAndVar) { colorAndVar[2 * i] = pts[i].color; colorAndVar[2 * i + 1] = pts[i].var; }
    if (idx) idx[i] = x + y * W;
  }
  return LSD_OK;
}

// ---- SE3 tracker --------------------------------------------------------------------------------

int lsd_se3_track_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_frameToRef,
                        lsd_se3_result *results, lsd_trace_entry *traces) {
  LSD_ARG(ctx && refs && frames && init_frameToRef && results && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return se3_track_batch_impl(ctx, n, refs, frames, init_frameToRef, results, traces, ctx->stream);
}

int lsd_ctx_set_permaref_settings(lsd_ctx *ctx, const lsd_tracker_settings *s) {
  LSD_ARG(ctx && s);
  ctx->permaref = *s;
  return LSD_OK;
}

int lsd_se3_track_permaref_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init_refToFrame,
                                 lsd_se3_result *results, lsd_trace_entry *traces) {
  LSD_ARG(ctx && refs && frames && init_refToFrame && results && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  ctx->se3Permaref = true;
  const int rc = se3_track_batch_impl(ctx, n, refs, frames, init_refToFrame, results, traces, ctx->stream);
  ctx->se3Permaref = false;
  return rc;
}

int lsd_se3_check_permaref_overlap_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, const double *refToFrame, float *pointUsage) {
  LSD_ARG(ctx && refs && refToFrame && pointUsage && n >= 0);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return se3_permaref_overlap_impl(ctx, n, refs, refToFrame, pointUsage);
}

int lsd_se3_track(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double init_frameToRef[7], lsd_se3_result *result,
                  lsd_trace_entry *trace) {
  return lsd_se3_track_batch(ctx, 1, &ref, &frame, init_frameToRef, result, trace);
}

int lsd_se3_eval(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double refToFrame[7], int level, float affine_a,
                 float affine_b, float *A36, float *b6, float *scalars) {
  LSD_ARG(ctx && ref && frame && refToFrame && A36 && b6 && scalars);
  LSD_CUDA(cudaSetDevice(ctx->device));
  return se3_eval_impl(ctx, ref, frame, refToFrame, level, affine_a, affine_b, A36, b6, scalars);
}

int lsd_se3_last_stats(lsd_ctx *ctx, double *algorithmic_bytes, long long *evaluations, float *kernel_ms) {
  LSD_ARG(ctx);
  if (algorithmic_bytes) *algorithmic_bytes = ctx->lastAlgBytes;
  if (evaluations) *evaluations = ctx->lastEvals;
  if (kernel_ms) *kernel_ms = ctx->lastKernelMs;
  return LSD_OK;
}

// Host images in, poses out: SE3Tracker::trackFrame for n freshly captured frames.  A fully asynchronous pipeline:
//   * every small host->device table (frame slab pointers, the tracker's pair table) is uploaded ONCE, before any
//     image copy is queued, so that no kernel launch ever waits behind a bulk transfer on the copy engine;
//   * images go up chunk by chunk on the copy stream (frames that are contiguous in host memory in one
//     cudaMemcpyAsync), double-buffered staging, event-ordered against the ingest that consumes them;
//   * per chunk the compute stream runs ingest -> gradients -> mask init -> tracker feed / launch back to back;
//   * the host synchronises once, at the end, and reads all results in one device->host copy.
// Streamed schedule (default): ONE persistent tracker is started up front on a third stream and fed chunk by chunk.  It needs
// the feeding kernels to become co-resident with it, which CUDA does not promise: where kernels are known to be serialised
// (CUDA_LAUNCH_BLOCKING, a profiler / sanitizer injected into the process) the per-chunk schedule is used instead, and a
// tracker that starves anyway stops itself (SE3Params::watchdogNs) and the batch is re-run with one ordinary launch.
static bool kernels_may_overlap() {
  const char *b = getenv("CUDA_LAUNCH_BLOCKING");
  if (b && atoi(b) != 0) return false;
  if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE"))
    return false;  // ncu / compute-sanitizer / nsys serialise kernels
  return true;
}

int lsd_ctx_set_image_pipeline(lsd_ctx *ctx, int chunkFrames, int streamed, double watchdogSeconds) {
  LSD_ARG(ctx && chunkFrames >= 0 && streamed >= -1 && streamed <= 1 && watchdogSeconds >= 0);
  ctx->imageChunk = chunkFrames;
  ctx->imageStreamed = streamed;
  if (watchdogSeconds > 0) ctx->streamWatchdogNs = (unsigned long long)(watchdogSeconds * 1e9);
  return LSD_OK;
}

int lsd_se3_track_images_batch(lsd_ctx *ctx, int n, lsd_ref *const *refs, const uint8_t *const *images, size_t pitch,
                               const double *init_frameToRef, lsd_se3_result *results) {
  LSD_ARG(ctx && refs && images && init_frameToRef && results && n >= 0);
  LSD_ARG(pitch >= (size_t)ctx->w);
  if (n == 0) return LSD_OK;
  LSD_CUDA(cudaSetDevice(ctx->device));
  const size_t fbytes = (size_t)ctx->w * ctx->h;
  static const int envChunk = getenv("LSD_B200_E2E_CHUNK") ? atoi(getenv("LSD_B200_E2E_CHUNK")) : 0;
  static const int envStream = getenv("LSD_B200_E2E_STREAM") ? atoi(getenv("LSD_B200_E2E_STREAM")) : -1;
  const int wantStream = ctx->imageStreamed >= 0 ? ctx->imageStreamed : envStream;
  const bool streamedCfg = wantStream >= 0 ? wantStream != 0 : kernels_may_overlap();
  // chunk = frames per H2D copy / ingest launch.  Measured on B200, 1000 pairs (profiles/): per-chunk tracker launches want big
  // chunks (250: 127 k frames/s); the streamed tracker wants small ones so that work arrives early and evenly
  // (250: 123 k, 125: 134 k, 63: 144 k, 48: 145 k, 32: 147 k, 20: 146 k)
  const int chunk = ctx->imageChunk > 0 ? ctx->imageChunk : (envChunk > 0 ? envChunk : (streamedCfg ? 48 : 250));
  const int CH = n < chunk ? n : chunk;
  int rc = ensure_stage(ctx, 0, 2 * fbytes * CH);
  if (rc) return rc;
  const int nChunks = (n + CH - 1) / CH;
  const bool streamed = streamedCfg && nChunks > 1;
  cudaStream_t st = ctx->stream;
  // Pageable caller memory (a cv::Mat handed to SlamSystem::nextImage is not pinned: lib/App/InputThread.cpp:58-71): a
  // cudaMemcpyAsync from it is staged by the driver on the calling thread, chunk by chunk, and blocks it (measured: 34 k
  // frames/s against 147 k from pinned memory).  Such images go through a pinned ring owned by the context instead, filled
  // by a few host threads while earlier chunks are on the copy engine.
  bool pageable = false;
  {
    cudaPointerAttributes a0, a1;
    const cudaError_t e0 = cudaPointerGetAttributes(&a0, images[0]), e1 = cudaPointerGetAttributes(&a1, images[n - 1]);
    if (e0 != cudaSuccess || e1 != cudaSuccess) cudaGetLastError();
    pageable = (e0 != cudaSuccess || a0.type == cudaMemoryTypeUnregistered) || (e1 != cudaSuccess || a1.type == cudaMemoryTypeUnregistered);
  }
  const int RING = 3;  // ring slots of CH frames each
  if (pageable) {
    rc = ensure_stage(ctx, (size_t)RING * fbytes * CH, 0);
    if (rc) return rc;
  }

  // everything this call creates is released on every exit path; a persistent tracker that was started is drained first
  struct Scope {
    lsd_ctx *ctx;
    std::vector<lsd_frame *> fr;
    std::vector<cudaEvent_t> copied, consumed;
    bool trackerRunning = false;
    std::vector<std::thread> stagers;  // pinned-ring staging of pageable images (below); joined before anything is released
    std::atomic<int> stageStop{0}, nextFrame{0}, freeUpTo{0};
    std::vector<std::atomic<int>> staged;
    ~Scope() {
      stageStop.store(1);
      for (std::thread &t : stagers) if (t.joinable()) t.join();
      if (trackerRunning) se3_stream_abort(ctx, ctx->trackStream, ctx->copyStream);
      cudaStreamSynchronize(ctx->copyStream);
      cudaStreamSynchronize(ctx->stream);
      for (cudaEvent_t e : copied) if (e) cudaEventDestroy(e);
      for (cudaEvent_t e : consumed) if (e) cudaEventDestroy(e);
      for (lsd_frame *f : fr) if (f) lsd_frame_release(ctx, f);
    }
  } sc;
  sc.ctx = ctx;
  sc.fr.assign(n, nullptr);
  std::vector<lsd_frame *> &fr = sc.fr;
  // frames + all tables first
  std::vector<void *> slabs(n);
  for (int i = 0; i < n; i++) {
    uint8_t *s = nullptr;
    rc = alloc_frame_slab(ctx, &s);
    if (rc) return rc;
    slabs[i] = s;
    fr[i] = new_frame(i, s);
    fr[i]->built = FB_TRACKING | FB_MASK;  // mask initialised below, in stream order
  }
  rc = upload_ptrs(ctx, slabs, st);
  if (rc) return rc;
  rc = se3_prepare(ctx, n, refs, fr.data(), init_frameToRef, false, st);
  if (rc) return rc;
  LSD_CUDA(cudaEventRecord(ctx->evPipe[0], st));
  LSD_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evPipe[0], 0));  // tables precede the bulk copies on the copy engine
  sc.copied.assign(nChunks, nullptr);
  sc.consumed.assign(nChunks, nullptr);
  std::vector<cudaEvent_t> &copied = sc.copied, &consumed = sc.consumed;
  for (int c = 0; c < nChunks; c++) {
    LSD_CUDA(cudaEventCreateWithFlags(&copied[c], cudaEventDisableTiming));
    LSD_CUDA(cudaEventCreateWithFlags(&consumed[c], cudaEventDisableTiming));
  }
  uint8_t *const *d_slabs = reinterpret_cast<uint8_t *const *>(ctx->d_table);
  // pinned-ring staging of pageable images: workers take frames in order; frame i may be written once the ring slot of its
  // chunk is free (freeUpTo), and chunk c may be copied once all of its frames are staged (staged[c])
  std::vector<std::atomic<int>> &staged = sc.staged;
  std::atomic<int> &nextFrame = sc.nextFrame, &freeUpTo = sc.freeUpTo;
  freeUpTo.store(RING - 1);
  if (pageable) {
    staged = std::vector<std::atomic<int>>(nChunks);
    for (auto &a : staged) a.store(0);
    static const int envT = getenv("LSD_B200_STAGE_THREADS") ? atoi(getenv("LSD_B200_STAGE_THREADS")) : 0;
    unsigned hw = std::thread::hardware_concurrency();
    int T = envT > 0 ? envT : (int)(hw / 2 < 1 ? 1 : (hw / 2 > 8 ? 8 : hw / 2));
    for (int i = 0; i < n; i++) LSD_ARG(images[i]);
    for (int t = 0; t < T; t++)
      sc.stagers.emplace_back([&, fbytes, CH]() {
        for (;;) {
          const int i = nextFrame.fetch_add(1);
          if (i >= n) return;
          const int c = i / CH;
          while (freeUpTo.load(std::memory_order_acquire) < c) {
            if (sc.stageStop.load()) return;
            std::this_thread::yield();
          }
          uint8_t *dst = ctx->h_stage + ((size_t)(c % RING) * CH + (size_t)(i - c * CH)) * fbytes;
          const uint8_t *src = images[i];
          if (pitch == (size_t)ctx->w) stage_copy(dst, src, fbytes);
          else for (int y = 0; y < ctx->h; y++) stage_copy(dst + (size_t)y * ctx->w, src + (size_t)y * pitch, ctx->w);
          staged[c].fetch_add(1, std::memory_order_release);
        }
      });
  }
  auto issue_copy = [&](int c) -> int {
    const int i0 = c * CH, m = (n - i0) < CH ? (n - i0) : CH;
    uint8_t *dst = ctx->d_stage + (size_t)(c & 1) * fbytes * CH;
    if (c >= 2) LSD_CUDA(cudaStreamWaitEvent(ctx->copyStream, consumed[c - 2], 0));  // staging half free again
    if (pageable) {
      while (staged[c].load(std::memory_order_acquire) < m) std::this_thread::yield();
      LSD_CUDA(cudaMemcpyAsync(dst, ctx->h_stage + (size_t)(c % RING) * CH * fbytes, fbytes * (size_t)m, cudaMemcpyHostToDevice, ctx->copyStream));
      LSD_CUDA(cudaEventRecord(copied[c], ctx->copyStream));
      if (c >= 1) {  // the ring slot of chunk c - 1 is free once its copy has completed: chunk c - 1 + RING may be staged
        LSD_CUDA(cudaEventSynchronize(copied[c - 1]));
        freeUpTo.store(c - 1 + RING, std::memory_order_release);
      }
      return LSD_OK;
    }
    int i = 0;
    while (i < m) {  // runs of frames that are back to back in host memory
      int j = i + 1;
      LSD_ARG(images[i0 + i]);
      if (pitch == (size_t)ctx->w) {
        while (j < m && images[i0 + j] == images[i0 + j - 1] + fbytes) j++;
        LSD_CUDA(cudaMemcpyAsync(dst + fbytes * i, images[i0 + i], fbytes * (size_t)(j - i), cudaMemcpyHostToDevice, ctx->copyStream));
      } else {
        LSD_CUDA(cudaMemcpy2DAsync(dst + fbytes * i, ctx->w, images[i0 + i], pitch, ctx->w, ctx->h, cudaMemcpyHostToDevice,
                                   ctx->copyStream));
      }
      i = j;
    }
    LSD_CUDA(cudaEventRecord(copied[c], ctx->copyStream));
    return LSD_OK;
  };
  if (streamed) {
    LSD_CUDA(cudaStreamWaitEvent(ctx->trackStream, ctx->evPipe[0], 0));  // pair table uploaded
    rc = se3_stream_begin(ctx, n, ctx->trackStream, ctx->evPipe[1]);
    if (rc) return rc;
    sc.trackerRunning = true;
    LSD_CUDA(cudaStreamWaitEvent(st, ctx->evPipe[1], 0));  // queue armed before the first feed
  }
  for (int c = 0; c < nChunks; c++) {
    const int i0 = c * CH, m = (n - i0) < CH ? (n - i0) : CH;
    rc = issue_copy(c);
    if (rc) return rc;
    LSD_CUDA(cudaStreamWaitEvent(st, copied[c], 0));
    const uint8_t *src = ctx->d_stage + (size_t)(c & 1) * fbytes * CH;
    launch_ingest(ctx, src, ctx->w, fbytes, d_slabs + i0, m, st);
    LSD_CUDA(cudaEventRecord(consumed[c], st));
    launch_gradients(ctx, d_slabs + i0, m, 1, NL - 1, st);
    launch_mask_init(ctx, d_slabs + i0, m, st);
    rc = streamed ? se3_stream_feed(ctx, i0, m, n, st) : se3_launch(ctx, i0, m, false, st);
    if (rc) return rc;
  }
  if (streamed) {
    LSD_CUDA(cudaEventRecord(ctx->evPipe[2], ctx->trackStream));  // completes when the persistent tracker has drained
    LSD_CUDA(cudaStreamWaitEvent(st, ctx->evPipe[2], 0));
    int starved = 0;
    rc = se3_stream_starved(ctx, st, &starved);  // synchronises: the tracker has exited
    sc.trackerRunning = false;
    if (rc) return rc;
    if (starved) {  // the producers never ran next to the tracker: every frame is ingested by now, track them in one launch
      rc = se3_launch(ctx, 0, n, false, st);
      if (rc) return rc;
    }
  }
  return se3_collect(ctx, n, refs, fr.data(), results, nullptr, st, 0.0f);
}

}  // extern "C"
