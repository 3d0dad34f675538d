// ctx.cuh -- lsd_ctx definition and kernel-launcher prototypes (internal).
#pragma once
#include <unordered_map>

#include "common.cuh"

struct SE3Scratch;   // se3_track.cu
struct HostPool;     // api.cu: a few host threads for staging pageable images into pinned memory
struct Sim3Scratch;  // sim3_track.cu
struct DepthScratch;

struct lsd_ctx {
  int device;
  int numSMs;
  int w, h;
  lsd::Intrinsics K;
  lsd::FrameLayout lay;
  cudaStream_t stream;
  bool ownStream;
  cudaStream_t copyStream;  // second stream for the pipelined host-image path
  cudaStream_t trackStream; // third stream: the persistent tracker of the streamed host-image path
  cudaEvent_t evA, evB, evPipe[4];
  long long launches;
  lsd_tracker_settings se3, sim3, permaref;
  bool se3Permaref;    // set for the duration of lsd_se3_track_permaref_batch
  int se3ActivePairs;  // 0: default; pairs in flight inside the persistent tracker (L2 residency)
  int se3RecsPerItem;  // 0: automatic (scheduling granularity only)
  bool tmaUnavailable; // cuTensorMapEncodeTiled missing or failing on this driver: the stencil kernels keep to vector loads
  int stencilTma;      // bit 0 / 1: regularizeDepthMap / fillHoles fetch their halo tiles with TMA (lsd_ctx_set_stencil_tma)
  int se3RecordPoints; // 0: default (4096); points per partial record = the summation order of the SE3 tracker
  int se3RecordPointsLvl[5];  // per level (0: se3RecordPoints applies): lsd_ctx_set_se3_record_points_per_level
  int se3LivePairs;    // -1: default (2); batches of at most this many pairs run on k_se3_track_live, one cluster per pair; 0: never
  int imageChunk;      // 0: default; frames per H2D copy / ingest launch of lsd_se3_track_images_batch
  int imageStreamed;   // -1: default (streamed when the platform runs kernels concurrently); 0 / 1: forced
  unsigned long long streamWatchdogNs;  // give-up time of the streamed tracker's work-item wait
  // pools
  std::vector<uint8_t *> frameSlabPool;
  std::vector<uint8_t *> refSlabPool;
  size_t refSlabBytes;
  // Device-resident table of the slab pointers this context has allocated (entry written once, when the slab is created): a
  // kernel that takes a pointer LIST (uint8_t *const *slabs) is handed the address of the one entry when it works on one frame --
  // the per-frame path uploads no pointer lists at all.
  void **d_ptrTable, **h_ptrTable;  // device table and its pinned mirror (an entry is written once and never changes)
  int ptrTableCount;
  std::unordered_map<const void *, int> ptrIndex;
  // staging for uploads
  uint8_t *h_stage;  // pinned
  size_t h_stageBytes;
  uint8_t *d_stage;
  size_t d_stageBytes;
  // small pinned / device tables for batched pointer lists
  void *h_table;
  void *d_table;
  size_t tableBytes;
  uint8_t *d_stats;  // k_idepth_stats: per frame a ticket (16 B) + per-CTA partial sums
  int statsFrames;   // frames d_stats has room for
  // Frame::setDepth bookkeeping (meanIdepth, numPoints) rides along with the setDepth launch: results land in pinned memory
  // and are attached to the frames after the call's own synchronisation (no extra launch + sync when the score asks for them)
  float *d_means, *h_means;
  int meansCap;
  std::vector<lsd_frame *> pendingMeans;
  // Pipelined lock-step driver (slam.cu): while deferSync is set, frame creation, reference import and updateKeyframe queue their
  // work and leave the final synchronisation to the next call that needs a result (the tracker's).  pendingSync = mapping work
  // of the last frame may still be in flight together with the pinned tables its uploads read from: everything that rewrites
  // such a table (ensure_table, the depth map's reference table, the mean-idepth read-back) finishes it first.
  bool deferSync, pendingSync;
  bool stageTimed;  // evA/evB bracket the kernels of the last depth stage
  int descSlot;  // rotating slot of the depth-map descriptor uploads (depth.cu)
  HostPool *pool;
  SE3Scratch *se3s;
  Sim3Scratch *sim3s;
  int sim3RecordPoints;  // 0: default (1024); points per partial record = the summation order of the Sim3 tracker
  // last-call stats
  double lastAlgBytes;
  long long lastEvals;
  float lastKernelMs;
};

namespace lsd {

int ensure_stage(lsd_ctx *ctx, size_t hostBytes, size_t devBytes);
int ensure_table(lsd_ctx *ctx, size_t bytes);
int ptr_table_register(lsd_ctx *ctx, const void *p);             // api.cu: after a cudaMalloc of a slab
void *const *ptr_table_entry(const lsd_ctx *ctx, const void *p);  // device address of the entry holding p, or nullptr
int ctx_finish_pending(lsd_ctx *ctx);  // api.cu: waits for deferred mapping work (no-op when there is none)
int frame_ensure_built(lsd_ctx *ctx, lsd_frame *f, unsigned need);  // api.cu: lazily builds planes (blocking)

// pyramid.cu
void launch_ingest(lsd_ctx *ctx, const uint8_t *d_src, size_t srcPitch, size_t srcFrameStride, uint8_t *const *d_slabs, int n,
                   cudaStream_t st);
void launch_gradients(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, int lvlLo, int lvlHi, cudaStream_t st, bool initMask = false);
void launch_maxgrad0(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st);
void launch_idepth_pyramid(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st, float *d_statOut2 = nullptr);
struct IdepthMapSrc {  // hypothesis planes Frame::setDepth reads (depth.cuh layout)
  const uint32_t *meta;
  const float *ids, *vars;
};
void launch_set_depth_and_pyramid(lsd_ctx *ctx, uint8_t *const *d_slabs, const IdepthMapSrc *d_srcs, int n, cudaStream_t st,
                                  float *d_statOut2 = nullptr);
void launch_set_depth_and_pyramid_one(lsd_ctx *ctx, uint8_t *slab, const IdepthMapSrc &src, cudaStream_t st, float *d_statOut2 = nullptr);
void launch_set_depth_gt(lsd_ctx *ctx, uint8_t *slab, const float *d_depth, float cov, cudaStream_t st);
void launch_mask_init(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, cudaStream_t st);
void launch_idepth_stats(lsd_ctx *ctx, uint8_t *slab, float *d_out2, cudaStream_t st);
void launch_idepth_stats_batch(lsd_ctx *ctx, uint8_t *const *d_slabs, int n, float *d_out2, cudaStream_t st);
int ensure_stats_scratch(lsd_ctx *ctx, int frames);
int prepare_mean_idepth(lsd_ctx *ctx, int n, float **d_out2);  // api.cu: before the setDepth launch
int schedule_mean_idepth(lsd_ctx *ctx, int n, lsd_frame *const *frames, cudaStream_t st);  // api.cu: after it
void resolve_pending_means(lsd_ctx *ctx);  // call after the stream has been synchronised

// trackref.cu
int pointcloud_state_words(const lsd_ctx *ctx);  // ints behind a reference's counters: numData[NL], pad, look-back state
void launch_make_pointcloud(lsd_ctx *ctx, uint8_t *const *d_kfSlabs, uint8_t *const *d_refSlabs, int *const *d_nums, int n,
                            const size_t *offPts, const size_t *offGrad, cudaStream_t st);

// se3_track.cu
int se3_track_batch_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init,
                         lsd_se3_result *results, lsd_trace_entry *traces, cudaStream_t st);
int se3_prepare(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, const double *init, bool wantTrace,
                cudaStream_t st, bool upload = true);
int se3_launch(lsd_ctx *ctx, int i0, int m, bool wantTrace, cudaStream_t st);
int se3_stream_begin(lsd_ctx *ctx, int n, cudaStream_t trackSt, cudaEvent_t armed);
int se3_stream_feed(lsd_ctx *ctx, int i0, int m, int n, cudaStream_t st);
int se3_stream_abort(lsd_ctx *ctx, cudaStream_t trackSt, cudaStream_t sideSt);
int se3_stream_starved(lsd_ctx *ctx, cudaStream_t st, int *starved);
int se3_collect(lsd_ctx *ctx, int n, lsd_ref *const *refs, lsd_frame *const *frames, lsd_se3_result *results, lsd_trace_entry *traces,
                cudaStream_t st, float kernelMs);
int se3_eval_impl(lsd_ctx *ctx, lsd_ref *ref, lsd_frame *frame, const double refToFrame[7], int level, float a, float b,
                  float *A36, float *b6, float *scalars);
int se3_record_points(const lsd_ctx *ctx, int level);
void se3_scratch_free(lsd_ctx *ctx);
void sim3_scratch_free(lsd_ctx *ctx);
int se3_permaref_overlap_impl(lsd_ctx *ctx, int n, lsd_ref *const *refs, const double *refToFrame, float *pointUsage);

}  // namespace lsd
