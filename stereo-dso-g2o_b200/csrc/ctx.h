// Context of the B200 hot path: calibration pyramid, device arenas, stream, settings.
// Replaces the reference's process-wide globals (util/globalCalib.cpp:33-46, util/settings.cpp).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <utility>
#include <cstdio>
#include "../../include/sdso_b200.h"
#include "common.cuh"

namespace sdso {

struct Frame {
  bool in_use = false;
  float4* tex[kPyrLevels] = {nullptr};  // one allocation, level pointers into it
  float* image = nullptr;               // level-0 intensity plane (the uploaded image)
  float ab_exposure = 1.0f;
  bool valid = false;
  bool plane_valid = false;             // `image` holds the level-0 intensities (true when the source was uploaded into it; else extracted on demand)
  unsigned gen = 0;                     // bumped by every makeImages on this slot (the selector keys its histograms on it)
  unsigned char* src8 = nullptr;        // device staging of an 8-bit source image (a slice of a batch arena)
  bool src8_owned = false;
  cudaEvent_t uploaded = nullptr;       // NOT owned: the event (ctx->upload_events ring) recorded on the copy stream after the upload batch this frame's source was in
  cudaEvent_t consumed = nullptr;       // NOT owned: the event (ctx->consume_events ring) recorded on the compute stream after the last makeImages that read src8 / image
  int pending_u8 = -1;                  // source format of the pending upload: -1 none, 0 float (in `image`), 1 uint8 (in `src8`)
};

struct HostCalib {  // per level, float, as util/globalCalib.cpp:48-108 computes them
  int levels = 0;
  int w[kPyrLevels], h[kPyrLevels];
  float fx[kPyrLevels], fy[kPyrLevels], cx[kPyrLevels], cy[kPyrLevels];
  float K[kPyrLevels][9], Ki[kPyrLevels][9];
  void set(int w0, int h0, float fx0, float fy0, float cx0, float cy0, bool decide_levels);
};

struct TrackerState;  // tracker.cu
struct BAState;       // ba.cu
struct TraceState;    // trace.cu
struct SelectorState; // pixel_select.cu
struct DistMapState;  // distmap.cu
struct UndistortState; // undistort.cu

}  // namespace sdso

namespace sdso { struct PeerExchange; }

struct sdso_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  sdso_settings S;
  sdso::HostCalib G;   // initial calibration == the reference's globals wG,hG,KG,...
  float baseline = 0;
  std::vector<sdso::Frame> frames;
  size_t tex_total = 0;  // float4 texels per frame over all levels
  float* staging = nullptr;  // pinned host staging for image upload
  // stream-ordered scratch of the per-call operators (set_ref, vertex / edge batches): a private pool that KEEPS its memory across
  // synchronisations (release threshold = max); the device's default pool hands everything back at every sync, which turns each
  // cudaMallocAsync of a key frame into a driver allocation (measured: sporadic 10-700 ms stalls in tracker_set_ref)
  cudaMemPool_t pool = nullptr;
  std::vector<void*> arenas;           // batch allocations of 8-bit staging
  cudaStream_t copy_stream = nullptr;  // H2D uploads that overlap the kernels of the previous step (sdso_upload_images_async)
  // one event per upload batch / per makeImages batch, shared by every frame of the batch. Rings: a slot that is re-recorded while an
  // old frame still points at it only makes that frame's wait more conservative (both streams complete in order).
  static constexpr int kEventRing = 64;
  cudaEvent_t upload_events[kEventRing] = {nullptr}, consume_events[kEventRing] = {nullptr};
  unsigned upload_seq = 0, consume_seq = 0;
  sdso::TrackerState* tracker = nullptr;
  sdso::BAState* ba = nullptr;
  sdso::TraceState* trace = nullptr;
  sdso::SelectorState* selector = nullptr;
  sdso::DistMapState* distmap = nullptr;
  sdso::UndistortState* undistort = nullptr;
  uint64_t launches = 0;
  // optional CUDA-event profiling of the two hot launches (bench.py roofline); see sdso_profile_*
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_track, ev_images;
  size_t ev_track_used = 0, ev_images_used = 0;
  int num_sms = 0;
  void* nccl_comm = nullptr;  // ncclComm_t of the sharded-BA allreduce (collective.cu), null unless sdso_nccl_init ran
  sdso::PeerExchange* peer = nullptr;   // one-shot allreduce over NVLink peer memory (collective.cu), null unless sdso_peer_connect ran
  int nccl_rank = 0, nccl_nranks = 1;
  std::string err;
};

namespace sdso {

inline int fail(sdso_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

// A context belongs to ONE device (sdso_ctx_create). Entry points that allocate or launch make that device current first, so two
// contexts on different GPUs can live in one process (the bench uses one process per GPU, where this is a no-op).
inline cudaError_t alloc_async(sdso_ctx* c, void** p, size_t bytes, cudaStream_t st) {
  return c->pool ? cudaMallocFromPoolAsync(p, bytes, c->pool, st) : cudaMallocAsync(p, bytes, st);
}
template <typename T>
inline cudaError_t alloc_async(sdso_ctx* c, T** p, size_t bytes, cudaStream_t st) { return alloc_async(c, reinterpret_cast<void**>(p), bytes, st); }

inline void enter(const sdso_ctx* c) {
  int cur = -1;
  if (c && cudaGetDevice(&cur) == cudaSuccess && cur != c->device) cudaSetDevice(c->device);
}

#define SDSO_CUDA(ctx, expr)                                                                       \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      return sdso::fail(ctx, SDSO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    }                                                                                              \
  } while (0)

#define SDSO_CHECK_LAUNCH(ctx)                                                                     \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return sdso::fail(ctx, SDSO_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// 3x3 float inverse by cofactors * (1/det), the same arithmetic Eigen's Matrix3f::inverse() performs,
// so K^-1 matches the reference's KiG / Ki bit for bit.
void inverse3f(const float m[9], float out[9]);

// host-side SE3 helpers in double (row-major 3x4)
void se3_exp(const double a[6], double T[12]);
void se3_mul(const double A[12], const double B[12], double C[12]);
void se3_inv(const double A[12], double B[12]);
void se3_log(const double T[12], double a[6]);
void se3_adj(const double T[12], double Ad[36]);  // 6x6 row-major
void so3_normalize(double T[12]);                 // rotation block back onto SO(3) through a unit quaternion (what Sophus' SE3 type guarantees)

// profiling helpers (ctx.cu)
void prof_begin(sdso_ctx* ctx, int which);
void prof_end(sdso_ctx* ctx, int which);
// make_images.cu
int make_images_launch(sdso_ctx* ctx, Frame& f, const float* dev_image, bool use_hcalib);
int make_images_batch_launch(sdso_ctx* ctx, int nb, Frame* const* frames, const void* const* srcs, bool src_u8, bool use_hcalib);
int ensure_intensity_plane(sdso_ctx* ctx, Frame& f);  // level-0 intensity plane for the kernels that read 4-byte pixels (epipolar search)
// tracker.cu
int tracker_create(sdso_ctx* ctx);
void tracker_destroy(sdso_ctx* ctx);
// ba.cu / trace.cu
int ba_create(sdso_ctx* ctx);
void ba_destroy(sdso_ctx* ctx);
void collective_destroy(sdso_ctx* ctx);
int trace_create(sdso_ctx* ctx);
void trace_destroy(sdso_ctx* ctx);
int selector_create(sdso_ctx* ctx);
void selector_destroy(sdso_ctx* ctx);
int distmap_create(sdso_ctx* ctx);
void distmap_destroy(sdso_ctx* ctx);
int undistort_create(sdso_ctx* ctx);
void undistort_destroy(sdso_ctx* ctx);

}  // namespace sdso
