// Coarse distance map and the activation candidate filter on the device
// (CoarseDistanceMap::{makeDistanceMap, growDistBFS, addIntoDistFinal}, FullSystem/CoarseTracker.cpp:1216-1366;
//  FullSystem::activatePointsMT STEP2, FullSystem.cpp:838-901).
//
// growDistBFS is a level-synchronous flood on the level-1 grid: step k (1..39) gives value k to every cell above k that
// touches a cell of value k-1 — 8-neighbourhood for odd k, 4-neighbourhood for even k — and cells on the image border never
// expand. Two facts carry the device design (both checked against the sequential restatement in tests/):
//  * single source, interior cells: value(dx, dy) = min k with max(|dx|,|dy|) <= k and |dx|+|dy| <= k + ceil(k/2); for a set
//    of seeds the value of an interior cell is the minimum over the seeds, so the field does not depend on the order in
//    which addIntoDistFinal inserted them. Border cells take (value of an interior neighbour) + 1 where a diagonal step
//    needs an even neighbour value, which is not monotone in the seed set — they are replayed in insertion order.
//  * the flood itself runs tile-wise: a CTA floods a 32x32 tile with a halo of 8 cells for 8 steps in shared memory
//    (errors from outside the halo travel one cell per step), so the 39 steps are 5 launches.
// The candidate loop accepts candidate i iff the field of (projected active points + candidates accepted before i) at its
// cell, plus the sub-pixel fraction, reaches currentMinActDist * my_type. It is evaluated in rounds: a candidate decides as
// soon as every earlier candidate that could reach its cell has decided (the lowest undecided index always can), rejects
// early when an accepted earlier candidate is already too close.
#include "ctx.h"
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace sdso {

enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };  // ImmaturePoint.h:50-56

// what the candidate loop reads of an ImmaturePoint (packed on the host: 28 bytes instead of the 148-byte record)
struct DmCand { float u, v, idepth_min, idepth_max, quality, lastTracePixelInterval; int lastTraceStatus; };

struct DistMapState {
  int w1 = 0, h1 = 0;
  unsigned char* d_map = nullptr;   // value 0..39, 255 = the reference's 1000 ("farther than 39 steps")
  int* d_acc_min = nullptr;         // per cell: lowest index of an accepted candidate, INT_MAX if none
  unsigned* d_und = nullptr;        // per cell: (round << 20) | (0xFFFFF - lowest undecided index) of the latest round that marked it
  // candidates
  int cap = 0;
  DmCand* d_pts = nullptr; DmCand* h_pts = nullptr;   // device / pinned staging
  int *d_host = nullptr, *d_cell = nullptr, *d_need = nullptr, *d_state = nullptr, *d_verdict = nullptr;
  float* d_type = nullptr;
  float *d_KRKi = nullptr, *d_Kt = nullptr; unsigned char* d_flag = nullptr; int cap_hosts = 0;
  float* d_uvid = nullptr; int* d_pt_host = nullptr; int cap_pts = 0;
  int* d_counter = nullptr;         // undecided candidates after the last round
  int* h_counter = nullptr;         // pinned
  float* d_mapf = nullptr;          // float view for download
};

constexpr int kDmTile = 32, kDmHalo = 8, kDmSpan = kDmTile + 2 * kDmHalo;

__device__ __forceinline__ bool dm_border(int x, int y, int w1, int h1) { return x == 0 || y == 0 || x == w1 - 1 || y == h1 - 1; }

// first-reach step of (dx, dy) from an interior seed (see header); >= 40 means never (value stays 1000)
__device__ __forceinline__ int dm_steps(int dx, int dy) {
  dx = abs(dx); dy = abs(dy);
  const int m = max(dx, dy), s = dx + dy;
  // smallest k with k >= m and k + ceil(k/2) >= s
  int k = m;
  while (k + ((k + 1) >> 1) < s) k++;
  return k;
}

__global__ void dm_seed_kernel(int n, const int* __restrict__ pt_host, const float* __restrict__ uvid, const float* __restrict__ KRKi,
                               const float* __restrict__ Kt, int w1, int h1, unsigned char* map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* Kr = KRKi + 9 * pt_host[i]; const float* kt = Kt + 3 * pt_host[i];
  const float pu = uvid[3 * i], pv = uvid[3 * i + 1], id = uvid[3 * i + 2];
  const float p0 = (Kr[0] * pu + Kr[1] * pv + Kr[2] * 1.f) + kt[0] * id;
  const float p1 = (Kr[3] * pu + Kr[4] * pv + Kr[5] * 1.f) + kt[1] * id;
  const float p2 = (Kr[6] * pu + Kr[7] * pv + Kr[8] * 1.f) + kt[2] * id;
  const float fu = p0 / p2 + 0.5f, fv = p1 / p2 + 0.5f;
  if (!(fu > -2e9f && fu < 2e9f && fv > -2e9f && fv < 2e9f)) return;
  const int u = (int)fu, v = (int)fv;
  if (!(u > 0 && v > 0 && u < w1 && v < h1)) return;
  map[u + w1 * v] = 0;
}

__global__ void dm_seed_cells_kernel(int n, const int* __restrict__ uv, int w1, int h1, unsigned char* map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = uv[2 * i], v = uv[2 * i + 1];
  if (u >= 0 && v >= 0 && u < w1 && v < h1) map[u + w1 * v] = 0;
}

// steps k0 .. k0 + nsteps - 1 (nsteps <= 8) of growDistBFS for one 32x32 tile, in place
__global__ void __launch_bounds__(256) dm_flood_kernel(unsigned char* map, int w1, int h1, int k0, int nsteps) {
  __shared__ unsigned char t[kDmSpan][kDmSpan + 4];
  const int x0 = blockIdx.x * kDmTile - kDmHalo, y0 = blockIdx.y * kDmTile - kDmHalo;
  for (int c = threadIdx.x; c < kDmSpan * kDmSpan; c += blockDim.x) {
    const int lx = c % kDmSpan, ly = c / kDmSpan, x = x0 + lx, y = y0 + ly;
    t[ly][lx] = (x >= 0 && y >= 0 && x < w1 && y < h1) ? map[x + w1 * y] : 255;
  }
  __syncthreads();
  for (int k = k0; k < k0 + nsteps; k++) {
    const bool eight = (k & 1) != 0;
    for (int c = threadIdx.x; c < (kDmSpan - 2) * (kDmSpan - 2); c += blockDim.x) {
      const int lx = 1 + c % (kDmSpan - 2), ly = 1 + c / (kDmSpan - 2), x = x0 + lx, y = y0 + ly;
      if (x < 0 || y < 0 || x >= w1 || y >= h1 || t[ly][lx] <= k) continue;
      bool hit = false;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int dx = (j == 0 || j == 4 || j == 7) ? 1 : ((j == 1 || j == 5 || j == 6) ? -1 : 0);
        const int dy = (j == 2 || j == 4 || j == 5) ? 1 : ((j == 3 || j == 6 || j == 7) ? -1 : 0);
        if (j >= 4 && !eight) break;
        // a neighbour expands only if it holds k-1 and is not on the image border
        if (t[ly + dy][lx + dx] == k - 1 && !dm_border(x + dx, y + dy, w1, h1)) hit = true;
      }
      if (hit) t[ly][lx] = (unsigned char)k;   // in place: a concurrent reader sees either > k or k, never k-1
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < kDmTile * kDmTile; c += blockDim.x) {
    const int lx = kDmHalo + c % kDmTile, ly = kDmHalo + c / kDmTile, x = x0 + lx, y = y0 + ly;
    if (x < w1 && y < h1) map[x + w1 * y] = t[ly][lx];
  }
}

__global__ void dm_to_float_kernel(const unsigned char* __restrict__ map, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = map[i] == 255 ? 1000.f : (float)map[i];
}

// ---- candidate filter -----------------------------------------------------------------------------------------------
struct FilterParams {
  int n, w1, h1;
  const DmCand* pts;
  const int* host; const float* type;
  const float *KRKi, *Kt; const unsigned char* flagged;
  const unsigned char* map;   // field of the active points (makeDistanceMap)
  int *cell, *need, *state, *verdict;   // state: 0 undecided, 1 accepted, 2 rejected / settled without the field
  int* acc_min; unsigned* und;
  float minActDist, minTraceQuality;
  int* counter;
};

__global__ void dm_filter_prep_kernel(FilterParams F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F.n) return;
  const DmCand ph = F.pts[i];
  const int hst = F.host[i];
  int verdict = -1, cell = -1, need = 0;
  if (!isfinite(ph.idepth_max) || ph.lastTraceStatus == IPS_OUTLIER) verdict = 2;
  else {
    const bool canActivate = (ph.lastTraceStatus == IPS_GOOD || ph.lastTraceStatus == IPS_SKIPPED || ph.lastTraceStatus == IPS_BADCONDITION ||
                              ph.lastTraceStatus == IPS_OOB) &&
                             ph.lastTracePixelInterval < 8 && ph.quality > F.minTraceQuality && (ph.idepth_max + ph.idepth_min) > 0;
    if (!canActivate) verdict = (F.flagged[hst] || ph.lastTraceStatus == IPS_OOB) ? 2 : 0;
    else {
      const float* Kr = F.KRKi + 9 * hst; const float* kt = F.Kt + 3 * hst;
      const float id = 0.5f * (ph.idepth_max + ph.idepth_min);
      const float p0 = (Kr[0] * ph.u + Kr[1] * ph.v + Kr[2] * 1.f) + kt[0] * id;
      const float p1 = (Kr[3] * ph.u + Kr[4] * ph.v + Kr[5] * 1.f) + kt[1] * id;
      const float p2 = (Kr[6] * ph.u + Kr[7] * ph.v + Kr[8] * 1.f) + kt[2] * id;
      const float fu = p0 / p2 + 0.5f, fv = p1 / p2 + 0.5f;
      const bool conv = fu > -2e9f && fu < 2e9f && fv > -2e9f && fv < 2e9f;
      const int u = conv ? (int)fu : -1, v = conv ? (int)fv : -1;
      if (!(u > 0 && v > 0 && u < F.w1 && v < F.h1)) verdict = 2;
      else {
        cell = u + F.w1 * v;
        const float frac = p0 - floorf(p0), th = F.minActDist * F.type[i];
        // smallest field value that passes `dist >= currentMinActDist * my_type` (FullSystem.cpp:886-889)
        need = INT_MAX;
        for (int val = 0; val < 40; val++) if ((float)val + frac >= th) { need = val; break; }
        if (need == INT_MAX && 1000.f + frac >= th) need = 1000;
      }
    }
  }
  int state = 2;
  if (cell >= 0) {
    if (need <= 0) { verdict = 1; state = 1; atomicMin(&F.acc_min[cell], i); }
    else if (need == INT_MAX) verdict = 0;
    else {
      const int x = cell % F.w1, y = cell / F.w1;
      const int v0 = F.map[cell] == 255 ? 1000 : F.map[cell];
      // interior cells: the field only decreases as candidates are accepted, so a failing start value is final
      if (!dm_border(x, y, F.w1, F.h1) && v0 < need) verdict = 0;
      else { state = 0; verdict = 0; }
    }
  }
  F.cell[i] = cell; F.need[i] = need; F.state[i] = state; F.verdict[i] = verdict;
}

__device__ __forceinline__ int und_index(unsigned v, int round) { return (int)(v >> 20) == round ? (int)(0xFFFFFu - (v & 0xFFFFFu)) : INT_MAX; }

__global__ void dm_filter_mark_kernel(FilterParams F, int round) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *F.counter = 0;
  if (i >= F.n || F.state[i] != 0) return;
  atomicMax(&F.und[F.cell[i]], ((unsigned)round << 20) | (0xFFFFFu - (unsigned)i));
}

// value of border cell (x, y) after the seeds of the window with index < i were inserted in index order (whole warp:
// the lanes share the scan for the next seed, every lane keeps the same replay state)
__device__ int dm_border_replay(const FilterParams& F, int x, int y, int i, int R, int lane) {
  const int w1 = F.w1, h1 = F.h1;
  const int c = x + w1 * y;
  int val = F.map[c] == 255 ? 1000 : F.map[c];
  int Dq[8]; bool okq[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int dx = (j == 0 || j == 4 || j == 7) ? 1 : ((j == 1 || j == 5 || j == 6) ? -1 : 0);
    const int dy = (j == 2 || j == 4 || j == 5) ? 1 : ((j == 3 || j == 6 || j == 7) ? -1 : 0);
    const int qx = x + dx, qy = y + dy;
    okq[j] = qx >= 0 && qy >= 0 && qx < w1 && qy < h1 && !dm_border(qx, qy, w1, h1);
    Dq[j] = okq[j] ? (F.map[qx + w1 * qy] == 255 ? 1000 : F.map[qx + w1 * qy]) : 1000;
  }
  const int xa = max(0, x - R), xb = min(w1 - 1, x + R), ya = max(0, y - R), yb = min(h1 - 1, y + R);
  const int ww = xb - xa + 1, cells = ww * (yb - ya + 1);
  int last = -1;
  for (;;) {  // next accepted seed of the window in index order
    unsigned long long best = ~0ull;   // (index << 32) | cell
    for (int q = lane; q < cells; q += 32) {
      const int px = xa + q % ww, py = ya + q / ww;
      const int a = F.acc_min[px + w1 * py];
      if (a > last && a < i) { const unsigned long long v = ((unsigned long long)a << 32) | (unsigned)(px + w1 * py); if (v < best) best = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o); best = u < best ? u : best; }
    if (best == ~0ull) break;
    last = (int)(best >> 32);
    const int bc = (int)(best & 0xffffffffu), bx = bc % w1, by = bc / w1;
    if (bx == x && by == y) val = 0;
    if (!dm_border(bx, by, w1, h1)) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int dx = (j == 0 || j == 4 || j == 7) ? 1 : ((j == 1 || j == 5 || j == 6) ? -1 : 0);
        const int dy = (j == 2 || j == 4 || j == 5) ? 1 : ((j == 3 || j == 6 || j == 7) ? -1 : 0);
        if (!okq[j]) continue;
        const int f = dm_steps(x + dx - bx, y + dy - by);
        if (f < Dq[j]) Dq[j] = f;
        // the neighbour offers Dq + 1; a diagonal offer needs an 8-step, i.e. an odd step number
        const int k = Dq[j] + 1;
        if (k <= 39 && (j < 4 || (k & 1)) && k < val) val = k;
      }
    }
  }
  return val;
}

// one warp per candidate: the lanes share the window
__global__ void __launch_bounds__(256) dm_filter_decide_kernel(FilterParams F, int round) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= F.n || F.state[i] != 0) return;
  const int w1 = F.w1, h1 = F.h1, c = F.cell[i], need = F.need[i];
  const int x = c % w1, y = c / w1;
  const bool border = dm_border(x, y, w1, h1);
  const int reach = min(need, 40);   // a seed farther than 39 steps never arrives
  const int R = min(39, border ? need : need - 1);
  const int xa = max(0, x - R), xb = min(w1 - 1, x + R), ya = max(0, y - R), yb = min(h1 - 1, y + R);
  const int ww = xb - xa + 1, cells = ww * (yb - ya + 1);
  bool rejected = false, blocked = false;
  for (int q0 = 0; q0 < cells; q0 += 32) {
    const int qi = q0 + lane;
    bool rej = false, blk = false;
    if (qi < cells) {
      const int px = xa + qi % ww, py = ya + qi / ww, q = px + w1 * py;
      bool matters = true;
      if (!border) {
        matters = q == c || !(dm_border(px, py, w1, h1) || dm_steps(px - x, py - y) >= reach);   // can this cell pull the field below `need`?
        if (matters && F.acc_min[q] < i) rej = true;
      }
      if (matters && und_index(F.und[q], round) < i) blk = true;
    }
    rejected |= __any_sync(0xffffffffu, rej);
    blocked |= __any_sync(0xffffffffu, blk);
    if (rejected) break;
  }
  if (!rejected && blocked) { if (lane == 0) atomicAdd(F.counter, 1); return; }
  if (!rejected && border) rejected = dm_border_replay(F, x, y, i, R, lane) < need;
  if (lane != 0) return;
  if (rejected) { F.state[i] = 2; F.verdict[i] = 0; }
  else { F.state[i] = 1; F.verdict[i] = 1; atomicMin(&F.acc_min[c], i); }
}

__global__ void dm_filter_apply_kernel(FilterParams F, unsigned char* map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F.n && F.state[i] == 1) map[F.cell[i]] = 0;
}

int distmap_create(sdso_ctx* ctx) {
  DistMapState* s = new DistMapState();
  ctx->distmap = s;
  if (ctx->G.levels < 2) return SDSO_OK;   // no level 1: the entry points report it
  s->w1 = ctx->G.w[1]; s->h1 = ctx->G.h[1];
  const size_t n = (size_t)s->w1 * s->h1;
  SDSO_CUDA(ctx, cudaMalloc(&s->d_map, n));
  SDSO_CUDA(ctx, cudaMemset(s->d_map, 255, n));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_mapf, n * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_acc_min, n * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_und, n * sizeof(unsigned)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_counter, sizeof(int)));
  SDSO_CUDA(ctx, cudaMallocHost(&s->h_counter, sizeof(int)));
  return SDSO_OK;
}

void distmap_destroy(sdso_ctx* ctx) {
  DistMapState* s = ctx->distmap;
  if (!s) return;
  void* ptrs[] = {s->d_map, s->d_mapf, s->d_acc_min, s->d_und, s->d_counter, s->d_pts, s->d_host, s->d_cell, s->d_need, s->d_state, s->d_verdict,
                  s->d_type, s->d_KRKi, s->d_Kt, s->d_flag, s->d_uvid, s->d_pt_host};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (s->h_counter) cudaFreeHost(s->h_counter);
  if (s->h_pts) cudaFreeHost(s->h_pts);
  delete s;
  ctx->distmap = nullptr;
}

static int ensure_hosts(sdso_ctx* ctx, int n_hosts) {
  DistMapState* s = ctx->distmap;
  if (n_hosts <= s->cap_hosts) return SDSO_OK;
  if (s->d_KRKi) cudaFree(s->d_KRKi);
  if (s->d_Kt) cudaFree(s->d_Kt);
  if (s->d_flag) cudaFree(s->d_flag);
  s->d_KRKi = s->d_Kt = nullptr; s->d_flag = nullptr;
  const int cap = n_hosts < 16 ? 16 : n_hosts;
  SDSO_CUDA(ctx, cudaMalloc(&s->d_KRKi, cap * 9 * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_Kt, cap * 3 * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_flag, cap));
  s->cap_hosts = cap;
  return SDSO_OK;
}

static int launch_flood(sdso_ctx* ctx) {
  DistMapState* s = ctx->distmap;
  const dim3 grid((s->w1 + kDmTile - 1) / kDmTile, (s->h1 + kDmTile - 1) / kDmTile);
  for (int k0 = 1; k0 < 40; k0 += 8) {
    dm_flood_kernel<<<grid, 256, 0, ctx->stream>>>(s->d_map, s->w1, s->h1, k0, min(8, 40 - k0));
    SDSO_CHECK_LAUNCH(ctx);
  }
  return SDSO_OK;
}

static int download_map(sdso_ctx* ctx, float* map_out) {
  DistMapState* s = ctx->distmap;
  const int n = s->w1 * s->h1;
  if (map_out) {
    dm_to_float_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(s->d_map, n, s->d_mapf);
    SDSO_CHECK_LAUNCH(ctx);
    SDSO_CUDA(ctx, cudaMemcpyAsync(map_out, s->d_mapf, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_distmap_make(sdso_ctx* ctx, int n_hosts, const float* KRKi, const float* Kt, int n_pts, const int* pt_host, const float* pt_uvid, float* map_out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->distmap || n_hosts < 0 || n_pts < 0) return SDSO_E_INVALID;
  DistMapState* s = ctx->distmap;
  if (!s->d_map) return fail(ctx, SDSO_E_INVALID, "distmap: needs pyramid level 1");
  if (n_pts > 0 && (!KRKi || !Kt || !pt_host || !pt_uvid || n_hosts == 0)) return SDSO_E_INVALID;
  for (int i = 0; i < n_pts; i++) if (pt_host[i] < 0 || pt_host[i] >= n_hosts) return fail(ctx, SDSO_E_INVALID, "distmap: point host out of range");
  int rc = ensure_hosts(ctx, n_hosts);
  if (rc) return rc;
  if (n_pts > s->cap_pts) {
    if (s->d_uvid) cudaFree(s->d_uvid);
    if (s->d_pt_host) cudaFree(s->d_pt_host);
    s->d_uvid = nullptr; s->d_pt_host = nullptr;
    const int cap = n_pts < 8192 ? 8192 : n_pts;
    SDSO_CUDA(ctx, cudaMalloc(&s->d_uvid, (size_t)cap * 3 * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_pt_host, (size_t)cap * sizeof(int)));
    s->cap_pts = cap;
  }
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemsetAsync(s->d_map, 255, (size_t)s->w1 * s->h1, st));
  if (n_pts > 0) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_KRKi, KRKi, (size_t)n_hosts * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_Kt, Kt, (size_t)n_hosts * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_uvid, pt_uvid, (size_t)n_pts * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_pt_host, pt_host, (size_t)n_pts * sizeof(int), cudaMemcpyHostToDevice, st));
    dm_seed_kernel<<<(n_pts + 255) / 256, 256, 0, st>>>(n_pts, s->d_pt_host, s->d_uvid, s->d_KRKi, s->d_Kt, s->w1, s->h1, s->d_map);
    SDSO_CHECK_LAUNCH(ctx);
  }
  rc = launch_flood(ctx);
  if (rc) return rc;
  return download_map(ctx, map_out);
}

int sdso_distmap_add(sdso_ctx* ctx, int n, const int* uv, float* map_out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->distmap || n < 0 || (n > 0 && !uv)) return SDSO_E_INVALID;
  DistMapState* s = ctx->distmap;
  if (!s->d_map) return fail(ctx, SDSO_E_INVALID, "distmap: needs pyramid level 1");
  if (n > 0) {
    int* d_uv = nullptr;
    SDSO_CUDA(ctx, cudaMalloc(&d_uv, (size_t)n * 2 * sizeof(int)));
    SDSO_CUDA(ctx, cudaMemcpyAsync(d_uv, uv, (size_t)n * 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    dm_seed_cells_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, d_uv, s->w1, s->h1, s->d_map);
    ctx->launches++;
    int rc = launch_flood(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_uv);
    if (rc) return rc;
  }
  return download_map(ctx, map_out);
}

int sdso_activation_filter(sdso_ctx* ctx, int n_hosts, const float* KRKi, const float* Kt, const unsigned char* host_flagged, int n, const int* cand_host,
                           const sdso_immature_point* pts, const float* my_type, float currentMinActDist, int* verdict, int* rounds, float* map_out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->distmap || n < 0 || n_hosts < 0) return SDSO_E_INVALID;
  DistMapState* s = ctx->distmap;
  if (!s->d_map) return fail(ctx, SDSO_E_INVALID, "distmap: needs pyramid level 1");
  if (n > 0 && (!KRKi || !Kt || !host_flagged || !cand_host || !pts || !my_type || !verdict || n_hosts == 0)) return SDSO_E_INVALID;
  if (n > 0xFFFFF) return fail(ctx, SDSO_E_INVALID, "activation_filter: more than 2^20 - 1 candidates");
  for (int i = 0; i < n; i++) if (cand_host[i] < 0 || cand_host[i] >= n_hosts) return fail(ctx, SDSO_E_INVALID, "activation_filter: host out of range");
  if (rounds) *rounds = 0;
  if (n == 0) return download_map(ctx, map_out);
  int rc = ensure_hosts(ctx, n_hosts);
  if (rc) return rc;
  if (n > s->cap) {
    void* old[] = {s->d_pts, s->d_host, s->d_cell, s->d_need, s->d_state, s->d_verdict, s->d_type};
    for (void* p : old) if (p) cudaFree(p);
    s->d_pts = nullptr; s->d_host = s->d_cell = s->d_need = s->d_state = s->d_verdict = nullptr; s->d_type = nullptr;
    const int cap = n < 16384 ? 16384 : n;
    if (s->h_pts) cudaFreeHost(s->h_pts);
    s->h_pts = nullptr;
    SDSO_CUDA(ctx, cudaMalloc(&s->d_pts, (size_t)cap * sizeof(DmCand)));
    SDSO_CUDA(ctx, cudaMallocHost(&s->h_pts, (size_t)cap * sizeof(DmCand)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_host, (size_t)cap * sizeof(int)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_cell, (size_t)cap * sizeof(int)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_need, (size_t)cap * sizeof(int)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_state, (size_t)cap * sizeof(int)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_verdict, (size_t)cap * sizeof(int)));
    SDSO_CUDA(ctx, cudaMalloc(&s->d_type, (size_t)cap * sizeof(float)));
    s->cap = cap;
  }
  cudaStream_t st = ctx->stream;
  const size_t cells = (size_t)s->w1 * s->h1;
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_KRKi, KRKi, (size_t)n_hosts * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_Kt, Kt, (size_t)n_hosts * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_flag, host_flagged, (size_t)n_hosts, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < n; i++) {
    const sdso_immature_point& p = pts[i];
    s->h_pts[i] = DmCand{p.u, p.v, p.idepth_min, p.idepth_max, p.quality, p.lastTracePixelInterval, p.lastTraceStatus};
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_pts, s->h_pts, (size_t)n * sizeof(DmCand), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_host, cand_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_type, my_type, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(s->d_acc_min, 0x7f, cells * sizeof(int), st));   // 0x7f7f7f7f: above every candidate index
  SDSO_CUDA(ctx, cudaMemsetAsync(s->d_und, 0, cells * sizeof(unsigned), st));
  FilterParams F{};
  F.n = n; F.w1 = s->w1; F.h1 = s->h1; F.pts = s->d_pts; F.host = s->d_host; F.type = s->d_type; F.KRKi = s->d_KRKi; F.Kt = s->d_Kt; F.flagged = s->d_flag;
  F.map = s->d_map; F.cell = s->d_cell; F.need = s->d_need; F.state = s->d_state; F.verdict = s->d_verdict; F.acc_min = s->d_acc_min; F.und = s->d_und;
  F.minActDist = currentMinActDist; F.minTraceQuality = ctx->S.minTraceQuality; F.counter = s->d_counter;
  const int grid = (n + 127) / 128;
  dm_filter_prep_kernel<<<grid, 128, 0, st>>>(F);
  SDSO_CHECK_LAUNCH(ctx);
  int round = 1;
  for (;;) {
    // a few rounds per host check; a round with nothing undecided is a pair of empty launches
    for (int r = 0; r < 4; r++, round++) {
      dm_filter_mark_kernel<<<grid, 128, 0, st>>>(F, round);
      SDSO_CHECK_LAUNCH(ctx);
      dm_filter_decide_kernel<<<(n + 7) / 8, 256, 0, st>>>(F, round);
      SDSO_CHECK_LAUNCH(ctx);
    }
    SDSO_CUDA(ctx, cudaMemcpyAsync(s->h_counter, s->d_counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
    if (*s->h_counter == 0) break;
    if (round > 2000) return fail(ctx, SDSO_E_STATE, "activation_filter: rounds do not converge");
  }
  if (rounds) *rounds = round - 1;
  SDSO_CUDA(ctx, cudaMemcpyAsync(verdict, s->d_verdict, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  // the field after the loop: accepted candidates become seeds (interior cells do not depend on the insertion order)
  dm_filter_apply_kernel<<<grid, 128, 0, st>>>(F, s->d_map);
  SDSO_CHECK_LAUNCH(ctx);
  rc = launch_flood(ctx);
  if (rc) return rc;
  return download_map(ctx, map_out);
}

}  // extern "C"
