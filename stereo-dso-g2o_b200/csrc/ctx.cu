// Context, frame arena, A1 entry points of the C ABI (include/sdso_b200.h).
#include "ctx.h"
#include <cstring>
#include <cmath>

namespace sdso {

int set_gamma_table(sdso_ctx* ctx, const float B[256]);

void inverse3f(const float m[9], float out[9]) {
  // cofactor(i,j) over cyclic indices; inverse(j,i) = cofactor(i,j) * (1/det); det expanded along column 0
  auto cof = [&](int i, int j) -> float {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
  };
  float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  float det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  float invdet = 1.0f / det;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[j * 3 + i] = cof(i, j) * invdet;
}

// util/globalCalib.cpp:48-108 (decide_levels) and CoarseTracker::makeK (CoarseTracker.cpp:108-136)
void HostCalib::set(int w0, int h0, float fx0, float fy0, float cx0, float cy0, bool decide_levels) {
  if (decide_levels) {
    int wl = w0, hl = h0;
    levels = 1;
    while (wl % 2 == 0 && hl % 2 == 0 && wl * hl > 5000 && levels < kPyrLevels) { wl /= 2; hl /= 2; levels++; }
  }
  w[0] = w0; h[0] = h0; fx[0] = fx0; fy[0] = fy0; cx[0] = cx0; cy[0] = cy0;
  for (int l = 1; l < kPyrLevels; l++) {
    w[l] = w0 >> l; h[l] = h0 >> l;
    fx[l] = (float)(fx[l - 1] * 0.5);
    fy[l] = (float)(fy[l - 1] * 0.5);
    cx[l] = (float)((cx[0] + 0.5) / ((int)1 << l) - 0.5);
    cy[l] = (float)((cy[0] + 0.5) / ((int)1 << l) - 0.5);
  }
  for (int l = 0; l < kPyrLevels; l++) {
    float Kl[9] = {fx[l], 0.f, cx[l], 0.f, fy[l], cy[l], 0.f, 0.f, 1.f};
    memcpy(K[l], Kl, sizeof(Kl));
    inverse3f(K[l], Ki[l]);
  }
}

// which: 0 = track kernel, 1 = makeImages kernels
void prof_begin(sdso_ctx* ctx, int which) {
  if (!ctx->profile) return;
  auto& v = which == 0 ? ctx->ev_track : ctx->ev_images;
  size_t& used = which == 0 ? ctx->ev_track_used : ctx->ev_images_used;
  if (used == v.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    v.push_back({a, b});
  }
  cudaEventRecord(v[used].first, ctx->stream);
}
void prof_end(sdso_ctx* ctx, int which) {
  if (!ctx->profile) return;
  auto& v = which == 0 ? ctx->ev_track : ctx->ev_images;
  size_t& used = which == 0 ? ctx->ev_track_used : ctx->ev_images_used;
  cudaEventRecord(v[used].second, ctx->stream);
  used++;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_profile_enable(sdso_ctx* ctx, int on) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  ctx->profile = on != 0;
  ctx->ev_track_used = ctx->ev_images_used = 0;
  return SDSO_OK;
}

int sdso_profile_read(sdso_ctx* ctx, double* track_ms, int* track_launches, double* images_ms, int* images_launches) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double t = 0, m = 0;
  for (size_t i = 0; i < ctx->ev_track_used; i++) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_track[i].first, ctx->ev_track[i].second); t += ms; }
  for (size_t i = 0; i < ctx->ev_images_used; i++) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_images[i].first, ctx->ev_images[i].second); m += ms; }
  if (track_ms) *track_ms = t;
  if (track_launches) *track_launches = (int)ctx->ev_track_used;
  if (images_ms) *images_ms = m;
  if (images_launches) *images_launches = (int)ctx->ev_images_used;
  ctx->ev_track_used = ctx->ev_images_used = 0;
  return SDSO_OK;
}

void sdso_default_settings(sdso_settings* s) {
  if (!s) return;
  s->huberTH = 9;
  s->coarseCutoffTH = 20;
  s->outlierTH = 12 * 12;
  s->outlierTHSumComponent = 50 * 50;
  s->overallEnergyTHWeight = 1;
  s->maxPixSearch = 0.027f;
  s->minTraceTestRadius = 2;
  s->trace_stepsize = 1.0f;
  s->trace_GNIterations = 3;
  s->trace_GNThreshold = 0.1f;
  s->trace_extraSlackOnTH = 1.2f;
  s->trace_slackInterval = 1.5f;
  s->trace_minImprovementFactor = 2;
  s->affineOptModeA = 0;
  s->affineOptModeB = 0;
  s->gammaWeightsPixelSelect = 1;
  s->g2o_stop_flag_persists = 1;
  s->cluster_size = 0;
  s->block_threads = 0;
  s->gather_batch = 0;
  s->idepthFixPrior = 50 * 50;
  s->idepthFixPriorMargFac = 600 * 600;
  s->initialRotPrior = 1e11f;
  s->initialTransPrior = 1e10f;
  s->initialAffBPrior = 1e14f;
  s->initialAffAPrior = 1e14f;
  s->initialCalibHessian = 5e9f;
  s->margWeightFac = 0.5f * 0.5f;
  s->solverModeDelta = 0.00001;
  s->minOptIterations = 1;
  s->thOptIterations = 1.2f;
  s->frameEnergyTHConstWeight = 0.5f;
  s->frameEnergyTHN = 0.7f;
  s->frameEnergyTHFacMedian = 1.5f;
  s->minGradHistCut = 0.5f;
  s->minGradHistAdd = 7;
  s->gradDownweightPerLevel = 0.75f;
  s->desiredImmatureDensity = 3000;
  s->minTraceQuality = 3;
  s->track_cache = 1;
}

int sdso_ctx_create(sdso_ctx** out, int device, int w, int h, const float K[4], float baseline, const sdso_settings* settings) {
  if (!out || !K || w < 16 || h < 16) return SDSO_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return SDSO_E_NODEVICE;  // no CPU fallback
  if (device < 0 || device >= ndev) return SDSO_E_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return SDSO_E_NODEVICE;
  sdso_ctx* c = new sdso_ctx();
  c->device = device;
  if (settings) c->S = *settings; else sdso_default_settings(&c->S);
  c->G.set(w, h, K[0], K[1], K[2], K[3], true);
  c->baseline = baseline;
  c->tex_total = 0;
  for (int l = 0; l < c->G.levels; l++) c->tex_total += (size_t)c->G.w[l] * c->G.h[l];
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return SDSO_E_CUDA; }
  c->num_sms = prop.multiProcessorCount;
  if (cudaMallocHost(&c->staging, (size_t)w * h * sizeof(float)) != cudaSuccess) { delete c; return SDSO_E_NOMEM; }
  {
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    if (cudaMemPoolCreate(&c->pool, &pp) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      c->pool = nullptr;   // fall back to the device's default pool
      cudaGetLastError();
    }
  }
  float B[256];
  for (int i = 0; i < 256; i++) B[i] = (float)i;
  int rc = set_gamma_table(c, B);
  if (rc == SDSO_OK) rc = tracker_create(c);
  if (rc == SDSO_OK) rc = ba_create(c);
  if (rc == SDSO_OK) rc = trace_create(c);
  if (rc == SDSO_OK) rc = selector_create(c);
  if (rc == SDSO_OK) rc = distmap_create(c);
  if (rc == SDSO_OK) rc = undistort_create(c);
  if (rc != SDSO_OK) { sdso_ctx_destroy(c); return rc; }
  *out = c;
  return SDSO_OK;
}

void sdso_ctx_destroy(sdso_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  collective_destroy(ctx);
  trace_destroy(ctx);
  selector_destroy(ctx);
  distmap_destroy(ctx);
  undistort_destroy(ctx);
  ba_destroy(ctx);
  tracker_destroy(ctx);
  for (auto& f : ctx->frames) {
    if (f.tex[0]) cudaFree(f.tex[0]);
    if (f.image) cudaFree(f.image);
    if (f.src8 && f.src8_owned) cudaFree(f.src8);
  }
  for (void* a : ctx->arenas) cudaFree(a);
  if (ctx->staging) cudaFreeHost(ctx->staging);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  for (int i = 0; i < sdso_ctx::kEventRing; i++) {
    if (ctx->upload_events[i]) cudaEventDestroy(ctx->upload_events[i]);
    if (ctx->consume_events[i]) cudaEventDestroy(ctx->consume_events[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& e : ctx->ev_track) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (auto& e : ctx->ev_images) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  delete ctx;
}

const char* sdso_last_error(const sdso_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sdso_set_stream(sdso_ctx* ctx, void* s) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  ctx->stream = (cudaStream_t)s;
  return SDSO_OK;
}
int sdso_synchronize(sdso_ctx* ctx) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}
int sdso_pyr_levels(const sdso_ctx* ctx) { return ctx ? ctx->G.levels : SDSO_E_INVALID; }
int sdso_level_size(const sdso_ctx* ctx, int lvl, int* w, int* h) {
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  if (w) *w = ctx->G.w[lvl];
  if (h) *h = ctx->G.h[lvl];
  return SDSO_OK;
}
int sdso_level_K(const sdso_ctx* ctx, int lvl, float K[9], float Ki[9]) {
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  if (K) memcpy(K, ctx->G.K[lvl], 9 * sizeof(float));
  if (Ki) memcpy(Ki, ctx->G.Ki[lvl], 9 * sizeof(float));
  return SDSO_OK;
}
uint64_t sdso_launch_count(const sdso_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sdso_set_gamma(sdso_ctx* ctx, const float B[256]) {
  sdso::enter(ctx);
  if (!ctx || !B) return SDSO_E_INVALID;
  return set_gamma_table(ctx, B);
}

int sdso_frame_create(sdso_ctx* ctx, int* frame_id) {
  sdso::enter(ctx);
  if (!ctx || !frame_id) return SDSO_E_INVALID;
  int id = -1;
  for (size_t i = 0; i < ctx->frames.size(); i++) if (!ctx->frames[i].in_use) { id = (int)i; break; }
  if (id < 0) {
    Frame f;
    float4* base = nullptr;
    SDSO_CUDA(ctx, cudaMalloc(&base, ctx->tex_total * sizeof(float4)));
    cudaError_t e = cudaMalloc(&f.image, ctx->tex_total * sizeof(float));
    if (e != cudaSuccess) { cudaFree(base); return fail(ctx, SDSO_E_CUDA, cudaGetErrorString(e)); }
    size_t off = 0;
    for (int l = 0; l < ctx->G.levels; l++) { f.tex[l] = base + off; off += (size_t)ctx->G.w[l] * ctx->G.h[l]; }
    ctx->frames.push_back(f);
    id = (int)ctx->frames.size() - 1;
  }
  ctx->frames[id].in_use = true;
  ctx->frames[id].valid = false;
  *frame_id = id;
  return SDSO_OK;
}

int sdso_frame_release(sdso_ctx* ctx, int frame_id) {
  sdso::enter(ctx);
  if (!ctx || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].in_use) return SDSO_E_INVALID;
  ctx->frames[frame_id].in_use = false;  // device memory is kept for reuse
  ctx->frames[frame_id].valid = false;
  return SDSO_OK;
}

int sdso_make_images(sdso_ctx* ctx, int frame_id, const float* host_image, float ab_exposure, int use_hcalib) {
  sdso::enter(ctx);
  if (!ctx || !host_image || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].in_use) return SDSO_E_INVALID;
  Frame& f = ctx->frames[frame_id];
  const size_t n = (size_t)ctx->G.w[0] * ctx->G.h[0];
  SDSO_CUDA(ctx, cudaMemcpyAsync(f.image, host_image, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  f.ab_exposure = ab_exposure;
  int rc = make_images_launch(ctx, f, f.image, use_hcalib != 0);
  if (rc) return rc;
  f.valid = true; f.gen++; f.plane_valid = true;
  return SDSO_OK;
}

int sdso_make_images_device(sdso_ctx* ctx, int frame_id, const float* device_image, float ab_exposure, int use_hcalib) {
  sdso::enter(ctx);
  if (!ctx || !device_image || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].in_use) return SDSO_E_INVALID;
  Frame& f = ctx->frames[frame_id];
  const size_t n = (size_t)ctx->G.w[0] * ctx->G.h[0];
  (void)n;  // the pyramid kernel copies an external level-0 image into the frame's own intensity plane
  f.ab_exposure = ab_exposure;
  int rc = make_images_launch(ctx, f, device_image, use_hcalib != 0);
  if (rc) return rc;
  f.valid = true; f.gen++; f.plane_valid = (device_image == f.image);
  return SDSO_OK;
}

static int check_batch(sdso_ctx* ctx, int nb, const int* frame_ids) {
  if (!ctx || nb < 0 || nb > 4096 || (nb > 0 && !frame_ids)) return SDSO_E_INVALID;
  for (int i = 0; i < nb; i++)
    if (frame_ids[i] < 0 || frame_ids[i] >= (int)ctx->frames.size() || !ctx->frames[frame_ids[i]].in_use) return SDSO_E_INVALID;
  return SDSO_OK;
}

// Asynchronous H2D upload of nb source images (float32 or uint8, w*h each, ideally pinned) on the context's copy stream.
// Returns immediately; sdso_make_images_uploaded makes the compute stream wait for exactly these copies.
int sdso_upload_images_async(sdso_ctx* ctx, int nb, const int* frame_ids, const void* const* host_images, int src_u8) {
  sdso::enter(ctx);
  int rc = check_batch(ctx, nb, frame_ids);
  if (rc) return rc;
  if (nb > 0 && !host_images) return SDSO_E_INVALID;
  if (!ctx->copy_stream) SDSO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  const size_t n = (size_t)ctx->G.w[0] * ctx->G.h[0];
  if (src_u8) {
    // frames that have no 8-bit staging yet get slices of ONE allocation, so a batch that is also contiguous on the host
    // (a ring buffer of camera frames) goes up as a single copy
    int missing = 0;
    for (int i = 0; i < nb; i++) if (!ctx->frames[frame_ids[i]].src8) missing++;
    if (missing > 0) {
      unsigned char* arena = nullptr;
      SDSO_CUDA(ctx, cudaMalloc(&arena, (size_t)missing * n));
      ctx->arenas.push_back(arena);
      int k = 0;
      for (int i = 0; i < nb; i++) { Frame& f = ctx->frames[frame_ids[i]]; if (!f.src8) { f.src8 = arena + (size_t)(k++) * n; f.src8_owned = false; } }
    }
  }
  const size_t bytes = src_u8 ? n : n * sizeof(float);
  // a makeImages still queued on the compute stream may be reading these staging buffers: the copies wait for it
  {
    cudaEvent_t seen[8]; int ns = 0;
    for (int i = 0; i < nb; i++) {
      cudaEvent_t e = ctx->frames[frame_ids[i]].consumed;
      if (!e) continue;
      bool dup = false;
      for (int k = 0; k < ns; k++) dup |= (seen[k] == e);
      if (dup) continue;
      SDSO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, e, 0));
      if (ns < 8) seen[ns++] = e;
    }
  }
  for (int i = 0; i < nb;) {
    Frame& f = ctx->frames[frame_ids[i]];
    unsigned char* dst = src_u8 ? f.src8 : (unsigned char*)f.image;
    const unsigned char* src = (const unsigned char*)host_images[i];
    int run = 1;  // longest run that is contiguous on both sides
    while (i + run < nb) {
      Frame& g = ctx->frames[frame_ids[i + run]];
      unsigned char* d2 = src_u8 ? g.src8 : (unsigned char*)g.image;
      if (d2 != dst + (size_t)run * bytes || (const unsigned char*)host_images[i + run] != src + (size_t)run * bytes) break;
      run++;
    }
    SDSO_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)run * bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    i += run;
  }
  if (nb == 0) return SDSO_OK;
  // one event for the whole batch (copies on one stream complete in order); EVERY frame of the batch points at it, so any subset
  // or reordering of the batch handed to sdso_make_images_uploaded waits for its copies
  cudaEvent_t& ev = ctx->upload_events[ctx->upload_seq++ % sdso_ctx::kEventRing];
  if (!ev) SDSO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SDSO_CUDA(ctx, cudaEventRecord(ev, ctx->copy_stream));
  for (int i = 0; i < nb; i++) {
    Frame& f = ctx->frames[frame_ids[i]];
    f.uploaded = ev;
    f.pending_u8 = src_u8 ? 1 : 0;
    f.valid = false;
  }
  return SDSO_OK;
}

// makeImages of nb frames whose sources were uploaded with sdso_upload_images_async: ONE pyramid + ONE gradient launch.
int sdso_make_images_uploaded(sdso_ctx* ctx, int nb, const int* frame_ids, const float* ab_exposure, int use_hcalib) {
  sdso::enter(ctx);
  int rc = check_batch(ctx, nb, frame_ids);
  if (rc) return rc;
  if (nb == 0) return SDSO_OK;
  std::vector<Frame*> fr(nb); std::vector<const void*> src(nb);
  const int u8 = ctx->frames[frame_ids[0]].pending_u8;
  for (int i = 0; i < nb; i++) {
    Frame& f = ctx->frames[frame_ids[i]];
    if (f.pending_u8 < 0 || f.pending_u8 != u8) return fail(ctx, SDSO_E_STATE, "make_images_uploaded: frame has no pending upload (or mixed source formats)");
    fr[i] = &f; src[i] = u8 ? (const void*)f.src8 : (const void*)f.image;
    f.ab_exposure = ab_exposure ? ab_exposure[i] : 1.0f;
  }
  {  // wait for every distinct upload batch these frames came from
    cudaEvent_t last = nullptr;
    for (int i = 0; i < nb; i++) {
      cudaEvent_t e = fr[i]->uploaded;
      if (!e) return fail(ctx, SDSO_E_STATE, "make_images_uploaded: frame has no upload event");
      if (e == last) continue;
      SDSO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
      last = e;
    }
  }
  for (int o = 0; o < nb; o += 128) {  // 128 frames per launch pair (kernel-parameter space)
    rc = make_images_batch_launch(ctx, nb - o < 128 ? nb - o : 128, fr.data() + o, src.data() + o, u8 != 0, use_hcalib != 0);
    if (rc) return rc;
  }
  {  // the next upload into these slots must not overwrite the staging while these launches read it
    cudaEvent_t& ev = ctx->consume_events[ctx->consume_seq++ % sdso_ctx::kEventRing];
    if (!ev) SDSO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SDSO_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    for (int i = 0; i < nb; i++) fr[i]->consumed = ev;
  }
  for (int i = 0; i < nb; i++) { fr[i]->valid = true; fr[i]->gen++; fr[i]->pending_u8 = -1; fr[i]->plane_valid = (!u8); }  // float uploads land in the plane
  return SDSO_OK;
}

// makeImages of nb frames from images already on the device (float32 or uint8): ONE pyramid + ONE gradient launch.
int sdso_make_images_batch_device(sdso_ctx* ctx, int nb, const int* frame_ids, const void* const* device_images, int src_u8,
                                  const float* ab_exposure, int use_hcalib) {
  sdso::enter(ctx);
  int rc = check_batch(ctx, nb, frame_ids);
  if (rc) return rc;
  if (nb == 0) return SDSO_OK;
  if (!device_images) return SDSO_E_INVALID;
  std::vector<Frame*> fr(nb);
  for (int i = 0; i < nb; i++) { fr[i] = &ctx->frames[frame_ids[i]]; fr[i]->ab_exposure = ab_exposure ? ab_exposure[i] : 1.0f; }
  for (int o = 0; o < nb; o += 128) {
    rc = make_images_batch_launch(ctx, nb - o < 128 ? nb - o : 128, fr.data() + o, device_images + o, src_u8 != 0, use_hcalib != 0);
    if (rc) return rc;
  }
  for (int i = 0; i < nb; i++) { fr[i]->valid = true; fr[i]->gen++; fr[i]->pending_u8 = -1; fr[i]->plane_valid = false; }
  return SDSO_OK;
}

int sdso_frame_download(sdso_ctx* ctx, int frame_id, int lvl, float* dI3, float* absgrad) {
  sdso::enter(ctx);
  if (!ctx || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].valid) return SDSO_E_INVALID;
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  const size_t n = (size_t)ctx->G.w[lvl] * ctx->G.h[lvl];
  std::vector<float4> tmp(n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->frames[frame_id].tex[lvl], n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; i++) {
    if (dI3) { dI3[3 * i] = tmp[i].x; dI3[3 * i + 1] = tmp[i].y; dI3[3 * i + 2] = tmp[i].z; }
    if (absgrad) absgrad[i] = tmp[i].w;
  }
  return SDSO_OK;
}

}  // extern "C"

namespace sdso {
__global__ void interp_kernel(const float4* tex, int width, const float2* xy, int n, float3* out, int bilin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 p = xy[i];
  out[i] = bilin ? interp33BiLin(tex, p.x, p.y, width) : interp33(tex, p.x, p.y, width);
}
}  // namespace sdso

extern "C" int sdso_interp33(sdso_ctx* ctx, int frame_id, int lvl, const float* xy, int n, float* out3, int bilin_variant) {
  sdso::enter(ctx);
  if (!ctx || !xy || !out3 || n < 0 || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].valid) return SDSO_E_INVALID;
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  if (n == 0) return SDSO_OK;
  float2* dxy = nullptr; float3* dout = nullptr;
  SDSO_CUDA(ctx, cudaMalloc(&dxy, n * sizeof(float2)));
  SDSO_CUDA(ctx, cudaMalloc(&dout, n * sizeof(float3)));
  SDSO_CUDA(ctx, cudaMemcpyAsync(dxy, xy, n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
  sdso::interp_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->frames[frame_id].tex[lvl], ctx->G.w[lvl], dxy, n, dout, bilin_variant);
  ctx->launches++;
  SDSO_CUDA(ctx, cudaMemcpyAsync(out3, dout, n * sizeof(float3), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(dxy); cudaFree(dout);
  return SDSO_OK;
}
