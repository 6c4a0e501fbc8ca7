// A4 — CoarseTracker::makeCoarseDepthL0 STEP1..STEP5 on the device (FullSystem/CoarseTracker.cpp:350-533;
// the first-frame variant :138-271 has the same steps). Builds the per-level tracking template
// pc_{u,v,idepth,color}[lvl] from n splats {u, v, idepth, weight} without a host round trip.
//
//   splat    : one thread per point, idepth*w and w into the level-0 maps (:350-354)
//   pool     : 2x2 sums up the pyramid, ((a+b)+c)+d as written (:360-386)
//   dilate   : diagonal 4-neighbours on levels 0-1 (:390-442), axis 4-neighbours above (:446-488),
//              reading the backup copy so the result is order independent
//   compact  : normalise + RASTER-ORDER stream compaction of the interior [2,w-2)x[2,h-2) (:492-533):
//              one warp per row counts (ballot/popc), a per-row prefix is summed, one warp per row writes.
//              Raster order is preserved exactly because calcRes samples its flow indicators at i%32==0
//              (:662) and float summation order downstream depends on it.
#include "tracker_state.h"

namespace sdso {

__device__ __forceinline__ int splat_pixel(const float4 p, int w0, int h0) {
  const int u = (int)(p.x + 0.5f), v = (int)(p.y + 0.5f);
  if (u < 0 || v < 0 || u >= w0 || v >= h0) return -1;  // the reference would write out of bounds
  return u + w0 * v;
}
// Splats that share a pixel must add in POINT order as the reference's loop does (:350-354); float atomics would add in arrival
// order. Pass 1 finds, per pixel, the lowest point index (owner) and the number of splats; pass 2: a pixel with one splat is
// written directly, a shared pixel is summed by its owner in index order (collisions are rare: a scan of the later points).
__global__ void splat_owner_kernel(const float4* __restrict__ pts, int n, int* owner, int* count, int w0, int h0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pix = splat_pixel(pts[i], w0, h0);
  if (pix < 0) return;
  atomicMin(&owner[pix], i);
  atomicAdd(&count[pix], 1);
}
__global__ void splat_kernel(const float4* __restrict__ pts, int n, const int* __restrict__ owner, const int* __restrict__ count,
                             float* idepth0, float* wsum0, int w0, int h0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int pix = splat_pixel(p, w0, h0);
  if (pix < 0 || owner[pix] != i) return;
  float id = p.z * p.w, ws = p.w;   // 0 + x == x
  int left = count[pix] - 1;
  for (int j = i + 1; j < n && left > 0; j++) {
    const float4 q = pts[j];
    if (splat_pixel(q, w0, h0) == pix) { id += q.z * q.w; ws += q.w; left--; }
  }
  idepth0[pix] = id;
  wsum0[pix] = ws;
}

__global__ void pool_kernel(const float* __restrict__ id_lm, const float* __restrict__ ws_lm, float* id_l, float* ws_l, int wl, int hl, int wlm1) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= wl || y >= hl) return;
  int bidx = 2 * x + 2 * y * wlm1;
  id_l[x + y * wl] = ((id_lm[bidx] + id_lm[bidx + 1]) + id_lm[bidx + wlm1]) + id_lm[bidx + wlm1 + 1];
  ws_l[x + y * wl] = ((ws_lm[bidx] + ws_lm[bidx + 1]) + ws_lm[bidx + wlm1]) + ws_lm[bidx + wlm1 + 1];
}

__global__ void dilate_kernel(float* idepthl, float* wsl, const float* __restrict__ bak, int wl, int hl, int diagonal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + wl;
  const int wh = wl * hl - wl;
  if (i >= wh) return;
  if (bak[i] <= 0) {
    int offs[4];
    if (diagonal) { offs[0] = 1 + wl; offs[1] = -1 - wl; offs[2] = wl - 1; offs[3] = -wl + 1; }
    else { offs[0] = 1; offs[1] = -1; offs[2] = wl; offs[3] = -wl; }
    float sum = 0, num = 0, numn = 0;
    const int npx = wl * hl;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = i + offs[k];
      // the reference reads one element before / past the map for the first / last pixel (heap garbage);
      // here an out-of-range neighbour counts as "no depth"
      float b = (j >= 0 && j < npx) ? bak[j] : 0.f;
      // idepthl is only read where bak>0 (never written by this kernel) and only written where bak<=0
      if (b > 0) { sum += idepthl[j]; num += b; numn++; }
    }
    if (numn > 0) { idepthl[i] = sum / numn; wsl[i] = num / numn; }
  }
}

__device__ __forceinline__ bool pc_candidate(const float* idepthl, const float* wsl, const float4* tex, int i, float& id_out, float& col_out) {
  float ws = wsl[i];
  if (!(ws > 0)) return false;
  float id = idepthl[i] / ws;
  float col = tex[i].x;
  id_out = id; col_out = col;
  return isfinite(col) && (id > 0);
}

// one warp per interior row: count valid pixels
__global__ void rowcount_kernel(const float* idepthl, const float* wsl, const float4* tex, int wl, int hl, int* rowcount) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int y = warp + 2;
  if (y >= hl - 2) return;
  int cnt = 0;
  for (int x0 = 2; x0 < wl - 2; x0 += 32) {
    int x = x0 + lane;
    bool ok = false;
    float a, b;
    if (x < wl - 2) ok = pc_candidate(idepthl, wsl, tex, x + y * wl, a, b);
    cnt += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) rowcount[y] = cnt;
}

__global__ void compact_kernel(const float* idepthl, const float* wsl, const float4* tex, int wl, int hl, const int* rowcount, float4* pc, int* total) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int y = warp + 2;
  if (y >= hl - 2) return;
  int base = 0;
  for (int r = 2 + lane; r < y; r += 32) base += rowcount[r];
  for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
  for (int x0 = 2; x0 < wl - 2; x0 += 32) {
    int x = x0 + lane;
    bool ok = false;
    float id = 0, col = 0;
    if (x < wl - 2) ok = pc_candidate(idepthl, wsl, tex, x + y * wl, id, col);
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) pc[base + __popc(m & ((1u << lane) - 1))] = make_float4((float)x, (float)y, id, col);
    base += __popc(m);
  }
  if (y == hl - 3 && lane == 0) *total = base;
}

}  // namespace sdso

using namespace sdso;

extern "C" int sdso_tracker_set_ref(sdso_ctx* ctx, int ref_frame, const float* uvidw, int n, const double ref_aff[2]) {
  sdso::enter(ctx);
  if (!ctx || n < 0 || (n > 0 && !uvidw) || !ref_aff) return SDSO_E_INVALID;
  if (ref_frame < 0 || ref_frame >= (int)ctx->frames.size() || !ctx->frames[ref_frame].valid) return fail(ctx, SDSO_E_INVALID, "bad ref_frame");
  TrackerState* t = ctx->tracker;
  const int L = ctx->G.levels;
  const int w0 = ctx->G.w[0], h0 = ctx->G.h[0];
  if (L < 1 || h0 < 6) return SDSO_E_INVALID;
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemsetAsync(t->idepth[0], 0, (size_t)w0 * h0 * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(t->wsum[0], 0, (size_t)w0 * h0 * sizeof(float), st));
  float4* dpts = nullptr;
  if (n > 0) {
    SDSO_CUDA(ctx, sdso::alloc_async(ctx, &dpts, (size_t)n * sizeof(float4), st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(dpts, uvidw, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
    int* owner = t->scan_tmp;                              // free until the row counts below
    int* count = reinterpret_cast<int*>(t->wsum_bak[0]);   // free until the dilation backup below
    SDSO_CUDA(ctx, cudaMemsetAsync(owner, 0x7f, (size_t)w0 * h0 * sizeof(int), st));
    SDSO_CUDA(ctx, cudaMemsetAsync(count, 0, (size_t)w0 * h0 * sizeof(int), st));
    splat_owner_kernel<<<(n + 255) / 256, 256, 0, st>>>(dpts, n, owner, count, w0, h0);
    SDSO_CHECK_LAUNCH(ctx);
    splat_kernel<<<(n + 255) / 256, 256, 0, st>>>(dpts, n, owner, count, t->idepth[0], t->wsum[0], w0, h0);
    SDSO_CHECK_LAUNCH(ctx);
  }
  for (int l = 1; l < L; l++) {
    int wl = ctx->G.w[l], hl = ctx->G.h[l];
    dim3 grid((wl + 127) / 128, hl);
    pool_kernel<<<grid, 128, 0, st>>>(t->idepth[l - 1], t->wsum[l - 1], t->idepth[l], t->wsum[l], wl, hl, ctx->G.w[l - 1]);
    SDSO_CHECK_LAUNCH(ctx);
  }
  for (int l = 0; l < L; l++) {
    int wl = ctx->G.w[l], hl = ctx->G.h[l];
    SDSO_CUDA(ctx, cudaMemcpyAsync(t->wsum_bak[l], t->wsum[l], (size_t)wl * hl * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int cnt = wl * hl - 2 * wl;
    if (cnt > 0) {
      dilate_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(t->idepth[l], t->wsum[l], t->wsum_bak[l], wl, hl, l < 2 ? 1 : 0);
      SDSO_CHECK_LAUNCH(ctx);
    }
  }
  SDSO_CUDA(ctx, cudaMemsetAsync(t->d_counts, 0, 64 * sizeof(int), st));
  int* rowcount = t->scan_tmp;
  for (int l = 0; l < L; l++) {
    int wl = ctx->G.w[l], hl = ctx->G.h[l];
    int rows = hl - 4;
    if (rows <= 0 || wl <= 4) continue;
    int blocks = (rows * 32 + 255) / 256;
    rowcount_kernel<<<blocks, 256, 0, st>>>(t->idepth[l], t->wsum[l], ctx->frames[ref_frame].tex[l], wl, hl, rowcount);
    SDSO_CHECK_LAUNCH(ctx);
    compact_kernel<<<blocks, 256, 0, st>>>(t->idepth[l], t->wsum[l], ctx->frames[ref_frame].tex[l], wl, hl, rowcount, t->pc[l], t->d_counts + l);
    SDSO_CHECK_LAUNCH(ctx);
    rowcount += hl;  // separate slice per level so the launches need no extra sync
  }
  int counts[64];
  SDSO_CUDA(ctx, cudaMemcpyAsync(counts, t->d_counts, 64 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (dpts) SDSO_CUDA(ctx, cudaFreeAsync(dpts, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  for (int l = 0; l < L; l++) t->pc_n[l] = counts[l];
  t->ref_frame = ref_frame;
  t->ref_exposure = ctx->frames[ref_frame].ab_exposure;
  t->ref_aff[0] = ref_aff[0]; t->ref_aff[1] = ref_aff[1];
  t->have_ref = true;
  return SDSO_OK;
}
