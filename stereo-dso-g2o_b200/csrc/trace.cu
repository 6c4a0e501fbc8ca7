// D1-D3, E3 — epipolar search for immature points on the device (FullSystem/ImmaturePoint.cpp:33-88, 94-451,
// 459-828; dso_g2o_edge.cpp:571-619; dso_g2o_vertex.cpp:73-88).
//
// One thread block per point (north_star): thread 0 runs the segment set-up (STEP1-2: project the inverse-depth
// interval, OOB / SKIPPED / BADCONDITION verdicts, errorInPixel), then every thread i < numSteps evaluates the
// 8-pixel Huber energy of search step i (8 x 4 gathers on the 4-byte intensity plane — the search needs no
// gradients, so it reads the planar level-0 image instead of the 16-byte texels), thread 0 picks best / second
// best (quality) and runs the <= 3 refinement iterations (texel gathers) and the new inverse-depth interval.
// Every float expression on the decision path keeps the reference's operand order (-fmad=false), including the
// repeated `ptx += dx` that positions step i, so status / bestIdx / numSteps are bit-exact.
#include "ctx.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace sdso {

enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };  // ImmaturePoint.h:50-56

struct TraceState {
  sdso_immature_point* d_pts = nullptr;   // staging of calls that carry host records
  int cap = 0;
  float* d_uv = nullptr;
  int* d_ok = nullptr;
  int cap_uv = 0;
  sdso_immature_point* d_pool = nullptr;  // device-resident records (sdso_immature_upload): traced in place, read back on demand
  int pool_cap = 0, pool_n = 0;
  int* d_xf_of = nullptr;                 // per-point host index of a multi-host call
  int xf_of_cap = 0;
  struct TraceXf* d_xf = nullptr;         // per-host transforms of the current call ([kMaxTraceHosts])
};
constexpr int kMaxTraceHosts = 16;

struct TraceXf {       // per host frame: hostToFrame_KRKi, hostToFrame_Kt, hostToFrame_affine (ImmaturePoint.h:90)
  float KRKi[9], Kt[3], aff[2];
  float bf;            // stereo: -K(0,0) * bl[0]
  float pad;
};
struct TraceParams {
  const float4* tex;   // level-0 texels of the searched frame
  const float* img;    // level-0 intensity plane of the searched frame
  int w, h;
  float maxPixSearch, huberTH, slackInterval, stepsize, minImprovementFactor, GNThreshold, extraSlackOnTH;
  int GNIterations, minTraceTestRadius;
};

__global__ void immature_init_kernel(const float4* __restrict__ tex, int width, const float* __restrict__ uv, int n, sdso_immature_point* out,
                                     int* ok, float outlierTHSumComponent, float outlierTH, float overallEnergyTHWeight) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sdso_immature_point p;
  p.u = uv[2 * i]; p.v = uv[2 * i + 1];
  p.idepth_min = 0; p.idepth_max = NAN;
  p.quality = 10000; p.energyTH = 0;
  p.u_stereo = p.u; p.v_stereo = p.v; p.idepth_min_stereo = 0; p.idepth_max_stereo = NAN; p.idepth_stereo = 0;
  p.lastTraceUV[0] = p.lastTraceUV[1] = 0; p.lastTracePixelInterval = 0;
  p.lastTraceStatus = IPS_UNINITIALIZED; p.bestIdx = -1; p.numSteps = 0;
  float g0 = 0, g1 = 0, g2 = 0, g3 = 0;
  int good = 1;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) { p.color[idx] = 0; p.weights[idx] = 0; }
  for (int idx = 0; idx < 8; idx++) {
    const float3 ptc = interp33BiLin(tex, p.u + kPatternP[idx][0], p.v + kPatternP[idx][1], width);
    p.color[idx] = ptc.x;
    if (!isfinite(ptc.x)) { p.energyTH = NAN; good = 0; break; }
    g0 += ptc.y * ptc.y; g1 += ptc.y * ptc.z; g2 += ptc.z * ptc.y; g3 += ptc.z * ptc.z;
    p.weights[idx] = sqrtf(outlierTHSumComponent / (outlierTHSumComponent + (ptc.y * ptc.y + ptc.z * ptc.z)));
  }
  p.gradH[0] = g0; p.gradH[1] = g1; p.gradH[2] = g2; p.gradH[3] = g3;
  if (good) {
    float e = 8 * outlierTH;
    e *= overallEnergyTHWeight * overallEnergyTHWeight;
    p.energyTH = e;
  }
  out[i] = p;
  ok[i] = good;
}

__device__ __forceinline__ bool inside_search(float u, float v, int w, int h) { return u > 4 && v > 4 && u < w - 5 && v < h - 5; }

constexpr int kTraceWarps = 4;      // points per CTA (one warp each)
constexpr int kMaxRounds = 4;       // ceil(99 / 32) search steps per lane

// One WARP per point. The segment set-up (STEP1-2) is a few dozen dependent scalar operations: every lane computes it
// redundantly (uniform control flow, no broadcast, no shared memory). The discrete search (STEP3) gives lane l the steps
// l, l + 32, ...: all 8 x 4 taps of a step are issued before the first is consumed, the arg-min (first minimum wins, as the
// reference's strict `<` scan does) and the masked second-best are warp reductions. In the refinement (STEP4) lane idx < 8
// gathers pattern pixel idx, so the 8 dependent gather round trips of an iteration collapse into one; the values are then
// broadcast and summed by every lane in the reference's pixel order, which keeps H, b and the energies bit-identical.
template <bool STEREO>
__global__ void __launch_bounds__(32 * kTraceWarps) trace_kernel(TraceParams T, const TraceXf* __restrict__ xfs, const int* __restrict__ xf_of,
                                                                  sdso_immature_point* pts, int n) {
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x * kTraceWarps + (threadIdx.x >> 5);
  if (pi >= n) return;
  sdso_immature_point& p = pts[pi];
  const TraceXf& X = xfs[xf_of ? xf_of[pi] : 0];
  const unsigned FULL = 0xffffffffu;
  int status = -1;
  float prv[3] = {0, 0, 0}, dx = 0, dy = 0, errorInPixel = 0, ptx0 = 0, pty0 = 0;
  int numSteps = 0;
  const int lastStatus = p.lastTraceStatus;
  float wUV0 = p.lastTraceUV[0], wUV1 = p.lastTraceUV[1], wInt = p.lastTracePixelInterval;   // values lane 0 writes back
  do {
    if (!STEREO && lastStatus == IPS_OOB) { status = IPS_OOB; break; }
    const float u0 = STEREO ? p.u_stereo : p.u, v0 = STEREO ? p.v_stereo : p.v;
    const float id_min = STEREO ? p.idepth_min_stereo : p.idepth_min, id_max = STEREO ? p.idepth_max_stereo : p.idepth_max;
    float pr[3], ptpMin[3], ptpMax[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pr[k] = X.KRKi[k * 3] * u0 + X.KRKi[k * 3 + 1] * v0 + X.KRKi[k * 3 + 2] * 1.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) ptpMin[k] = pr[k] + X.Kt[k] * id_min;
    const float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
    bool oob = !inside_search(uMin, vMin, T.w, T.h);
    float dist = 0, uMax = 0, vMax = 0;
    if (!oob) {
      if (isfinite(id_max)) {
#pragma unroll
        for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + X.Kt[k] * id_max;
        uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
        if (!inside_search(uMax, vMax, T.w, T.h)) oob = true;
        else {
          dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
          dist = sqrtf(dist);
          if (dist < T.slackInterval) {
            if (!STEREO) { wUV0 = (uMax + uMin) * 0.5f; wUV1 = (vMax + vMin) * 0.5f; wInt = dist; }
            status = IPS_SKIPPED;
            break;
          }
        }
      } else {
        dist = T.maxPixSearch;
#pragma unroll
        for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + X.Kt[k] * 0.01f;
        uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
        const float ddx = uMax - uMin, ddy = vMax - vMin;
        const float d = 1.0f / sqrtf(ddx * ddx + ddy * ddy);
        uMax = uMin + dist * ddx * d;
        vMax = vMin + dist * ddy * d;
        if (!inside_search(uMax, vMax, T.w, T.h)) oob = true;
      }
    }
    // scale-change test (:589-595; traceStereo tests the temporal idepth_min member, :197)
    if (!oob && !(p.idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) oob = true;
    if (oob) { wUV0 = wUV1 = -1; wInt = 0; status = IPS_OOB; break; }
    dx = T.stepsize * (uMax - uMin); dy = T.stepsize * (vMax - vMin);
    const float* g = p.gradH;
    const float a = (dx * g[0] + dy * g[2]) * dx + (dx * g[1] + dy * g[3]) * dy;
    const float b = (dy * g[0] + (-dx) * g[2]) * dy + (dy * g[1] + (-dx) * g[3]) * (-dx);
    errorInPixel = 0.2f + 0.2f * (a + b) / a;
    if (errorInPixel * T.minImprovementFactor > dist && isfinite(id_max)) {
      if (!STEREO) { wUV0 = (uMax + uMin) * 0.5f; wUV1 = (vMax + vMin) * 0.5f; wInt = dist; }
      status = IPS_BADCONDITION;
      break;
    }
    if (errorInPixel > 10) errorInPixel = 10;
    dx /= dist; dy /= dist;
    if (dist > T.maxPixSearch) { uMax = uMin + T.maxPixSearch * dx; vMax = vMin + T.maxPixSearch * dy; dist = T.maxPixSearch; }
    numSteps = (int)(1.9999f + dist / T.stepsize);
    const float randShift = uMin * 1000 - floorf(uMin * 1000);
    ptx0 = uMin - randShift * dx; pty0 = vMin - randShift * dy;
    if (!isfinite(dx) || !isfinite(dy)) { wUV0 = wUV1 = -1; wInt = 0; status = IPS_OOB; break; }
    if (numSteps >= 100) numSteps = 99;
    prv[0] = pr[0]; prv[1] = pr[1]; prv[2] = pr[2];
  } while (false);
  if (status >= 0) {   // finished in the set-up (uniform over the warp)
    if (lane == 0) {
      p.numSteps = 0; p.bestIdx = -1;
      p.lastTraceUV[0] = wUV0; p.lastTraceUV[1] = wUV1; p.lastTracePixelInterval = wInt;
      if (!(!STEREO && lastStatus == IPS_OOB)) p.lastTraceStatus = status;
    }
    return;
  }
  float rot[8][2];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    rot[idx][0] = X.KRKi[0] * kPatternP[idx][0] + X.KRKi[1] * kPatternP[idx][1];
    rot[idx][1] = X.KRKi[3] * kPatternP[idx][0] + X.KRKi[4] * kPatternP[idx][1];
  }
  float color[8], weights[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) { color[idx] = p.color[idx]; weights[idx] = p.weights[idx]; }
  const float aff0 = X.aff[0], aff1 = X.aff[1];
  // ---- STEP3: discrete search (:659-691): lane l evaluates steps l, l + 32, ... ----
  float err[kMaxRounds];
  float myBest = 1e10f, myU = 0, myV = 0;
  int myIdx = -1;
  {
    float ptx = ptx0, pty = pty0;
    for (int k = 0; k < lane; k++) { ptx += dx; pty += dy; }   // the reference positions step i by i additions
#pragma unroll
    for (int r = 0; r < kMaxRounds; r++) {
      const int i = lane + 32 * r;
      err[r] = 1e10f;   // (never selected: the scans below use strict <)
      if (32 * r < numSteps) {   // uniform
        if (i < numSteps) {
          float hit[8];
#pragma unroll
          for (int idx = 0; idx < 8; idx++) hit[idx] = interp31(T.img, (float)(ptx + rot[idx][0]), (float)(pty + rot[idx][1]), T.w);
          float energy = 0;
#pragma unroll
          for (int idx = 0; idx < 8; idx++) {
            if (!isfinite(hit[idx])) { energy += 1e5f; continue; }
            const float residual = hit[idx] - (float)(aff0 * color[idx] + aff1);
            const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
            energy += hw * residual * residual * (2 - hw);
          }
          err[r] = energy;
          if (energy < myBest) { myBest = energy; myU = ptx; myV = pty; myIdx = i; }
        }
        for (int k = 0; k < 32; k++) { ptx += dx; pty += dy; }
      }
    }
  }
  // arg-min with the reference's tie rule (first minimum wins): lexicographic (energy, index) minimum over the lanes
  float bestEnergy = myBest;
  int bestIdx = myIdx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oe = __shfl_xor_sync(FULL, bestEnergy, o);
    const int oi = __shfl_xor_sync(FULL, bestIdx, o);
    const bool take = (oi >= 0) && (bestIdx < 0 || oe < bestEnergy || (oe == bestEnergy && oi < bestIdx));
    if (take) { bestEnergy = oe; bestIdx = oi; }
  }
  if (bestIdx < 0) bestEnergy = 1e10f;
  const int src = bestIdx >= 0 ? (bestIdx & 31) : 0;
  float bestU = __shfl_sync(FULL, myU, src), bestV = __shfl_sync(FULL, myV, src);
  if (bestIdx < 0) { bestU = 0; bestV = 0; }
  // second best outside +-minTraceTestRadius (:694-701)
  float secondBest = 1e10f;
#pragma unroll
  for (int r = 0; r < kMaxRounds; r++) {
    const int i = lane + 32 * r;
    if (i < numSteps && (i < bestIdx - T.minTraceTestRadius || i > bestIdx + T.minTraceTestRadius) && err[r] < secondBest) secondBest = err[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const float ov = __shfl_xor_sync(FULL, secondBest, o); if (ov < secondBest) secondBest = ov; }
  const float newQuality = secondBest / bestEnergy;
  float quality = p.quality;
  if (newQuality < quality || numSteps > 10) quality = newQuality;
  // ---- STEP4: refinement: lane idx < 8 gathers pattern pixel idx; every lane sums in pixel order ----
  if (T.GNIterations > 0) bestEnergy = 1e5f;
  const int gl = lane & 7;
  float rx = 0, ry = 0;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) if (idx == gl) { rx = rot[idx][0]; ry = rot[idx][1]; }
  if (!STEREO) {  // traceOn (:707-779): damped (H = 1 + ..) GN with back-off
    float uBak = bestU, vBak = bestV, stepBack = 0;
    const float gnstepsize = 1;
    for (int it = 0; it < T.GNIterations; it++) {
      float H = 1, b = 0, energy = 0;
      const float3 mine = interp33(T.tex, (float)(bestU + rx), (float)(bestV + ry), T.w);
#pragma unroll
      for (int idx = 0; idx < 8; idx++) {
        const float hx = __shfl_sync(FULL, mine.x, idx), hy = __shfl_sync(FULL, mine.y, idx), hz = __shfl_sync(FULL, mine.z, idx);
        if (!isfinite(hx)) { energy += 1e5f; continue; }
        const float residual = hx - (aff0 * color[idx] + aff1);
        const float dResdDist = dx * hy + dy * hz;
        const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
        H += hw * dResdDist * dResdDist;
        b += hw * residual * dResdDist;
        energy += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
      }
      if (energy > bestEnergy) {
        stepBack *= 0.5f;
        bestU = uBak + stepBack * dx;
        bestV = vBak + stepBack * dy;
      } else {
        float step = -gnstepsize * b / H;
        if (step < -0.5f) step = -0.5f;
        else if (step > 0.5f) step = 0.5f;
        if (!isfinite(step)) step = 0;
        uBak = bestU; vBak = bestV; stepBack = step;
        bestU += step * dx; bestV += step * dy;
        bestEnergy = energy;
      }
      if (fabsf(stepBack) < T.GNThreshold) break;
    }
  } else {  // traceStereo (:309-411): g2o Gauss-Newton over EdgeTracePointUVDSO (E3), VertexUVDSO::oplusImpl clamp
    const double dxd = dx, dyd = dy;
    double U = bestU, V = bestV;
    const int wl = T.w - 3, hl = T.h - 3;
    for (int it = 0; it < T.GNIterations; it++) {
      float energy = 0;
      double H = 0, b = 0;
      const bool outside = (U - 2) < 0 || (U + 3) > wl || (V - 2) < 0 || (V + 3) > hl;  // util::CheckBoundary (dso_util.hpp:36-45)
      float3 mine = make_float3(0.f, 0.f, 0.f);
      if (!outside) mine = interp33(T.tex, (float)(U + rx), (float)(V + ry), T.w);
#pragma unroll
      for (int idx = 0; idx < 8; idx++) {
        const float hx = __shfl_sync(FULL, mine.x, idx), hy = __shfl_sync(FULL, mine.y, idx), hz = __shfl_sync(FULL, mine.z, idx);
        double err_ = 0, J = 0;
        if (!outside && isfinite(hx)) {
          err_ = hx - ((double)aff0 * (double)color[idx] + (double)aff1);
          J = dxd * hy + dyd * hz;
        }
        const float residual = (float)err_;
        const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
        energy += weights[idx] * weights[idx] * hw * residual * residual * (2 - hw);
        const double e2 = err_ * err_, dlt = T.huberTH;
        const double rho1 = (e2 <= dlt * dlt) ? 1.0 : dlt / sqrt(e2);  // RobustKernelHuber, first derivative
        H += rho1 * J * J;
        b -= rho1 * J * err_;
      }
      if (H > 0 && isfinite(H)) {
        double update = b / H;
        if (update < -0.5) update = -0.5;
        else if (update > 0.5) update = 0.5;
        else if (!isfinite(update)) update = 0;
        U += update * dxd; V += update * dyd;
      }
      if (!(energy > bestEnergy)) bestEnergy = energy;
    }
    bestU = (float)U; bestV = (float)V;
  }
  if (lane != 0) return;
  p.quality = quality; p.numSteps = numSteps; p.bestIdx = bestIdx;
  // ---- energy-based outlier (:781-793) ----
  if (!(bestEnergy < p.energyTH * T.extraSlackOnTH)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
    p.lastTraceStatus = (lastStatus == IPS_OUTLIER) ? IPS_OOB : IPS_OUTLIER;
    return;
  }
  // ---- STEP5: new interval (:795-827) ----
  const float e = errorInPixel;
  float id_min, id_max;
  if (dx * dx > dy * dy) {
    id_min = (prv[2] * (bestU - e * dx) - prv[0]) / (X.Kt[0] - X.Kt[2] * (bestU - e * dx));
    id_max = (prv[2] * (bestU + e * dx) - prv[0]) / (X.Kt[0] - X.Kt[2] * (bestU + e * dx));
  } else {
    id_min = (prv[2] * (bestV - e * dy) - prv[1]) / (X.Kt[1] - X.Kt[2] * (bestV - e * dy));
    id_max = (prv[2] * (bestV + e * dy) - prv[1]) / (X.Kt[1] - X.Kt[2] * (bestV + e * dy));
  }
  if (id_min > id_max) { const float q = id_min; id_min = id_max; id_max = q; }
  if (STEREO) { p.idepth_min_stereo = id_min; p.idepth_max_stereo = id_max; }
  else { p.idepth_min = id_min; p.idepth_max = id_max; }
  if (!isfinite(id_min) || !isfinite(id_max) || (id_max < 0)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
    p.lastTraceStatus = IPS_OUTLIER;
    return;
  }
  p.lastTracePixelInterval = 2 * e;
  p.lastTraceUV[0] = bestU; p.lastTraceUV[1] = bestV;
  if (STEREO) p.idepth_stereo = (p.u_stereo - bestU) / X.bf;
  p.lastTraceStatus = IPS_GOOD;
}

int trace_create(sdso_ctx* ctx) { ctx->trace = new TraceState(); return SDSO_OK; }
void trace_destroy(sdso_ctx* ctx) {
  if (!ctx->trace) return;
  TraceState* t = ctx->trace;
  void* ptrs[] = {t->d_pts, t->d_uv, t->d_ok, t->d_pool, t->d_xf_of, t->d_xf};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete ctx->trace;
  ctx->trace = nullptr;
}

static int ensure_pts(sdso_ctx* ctx, int n) {
  TraceState* t = ctx->trace;
  if (n > t->cap) {
    if (t->d_pts) cudaFree(t->d_pts);
    t->d_pts = nullptr;
    const int cap = n < 4096 ? 4096 : n;
    SDSO_CUDA(ctx, cudaMalloc(&t->d_pts, (size_t)cap * sizeof(sdso_immature_point)));
    t->cap = cap;
  }
  if (!t->d_xf) SDSO_CUDA(ctx, cudaMalloc(&t->d_xf, kMaxTraceHosts * sizeof(TraceXf)));
  return SDSO_OK;
}

static void fill_settings(const sdso_ctx* ctx, TraceParams& T) {
  const sdso_settings& S = ctx->S;
  T.w = ctx->G.w[0]; T.h = ctx->G.h[0];
  T.maxPixSearch = (ctx->G.w[0] + ctx->G.h[0]) * S.maxPixSearch;
  T.huberTH = S.huberTH; T.slackInterval = S.trace_slackInterval; T.stepsize = S.trace_stepsize;
  T.minImprovementFactor = S.trace_minImprovementFactor; T.GNThreshold = S.trace_GNThreshold; T.extraSlackOnTH = S.trace_extraSlackOnTH;
  T.GNIterations = S.trace_GNIterations; T.minTraceTestRadius = S.minTraceTestRadius;
}

__global__ void trace_status_kernel(const sdso_immature_point* __restrict__ pts, int n, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) status[i] = pts[i].lastTraceStatus;
}

// One launch for n points against nxf host transforms. pts != nullptr: host records, uploaded, traced, read back (the operator
// call of the reference's per-point loop). pts == nullptr: the device-resident pool is traced in place; only the statuses come
// back (if asked for).
static int run_trace(sdso_ctx* ctx, int frame, const TraceXf* xf, int nxf, const int* xf_of, bool stereo, int n, sdso_immature_point* pts, int* status) {
  if (frame < 0 || frame >= (int)ctx->frames.size() || !ctx->frames[frame].valid) return fail(ctx, SDSO_E_INVALID, "trace: invalid frame");
  if (n < 0 || nxf < 1 || nxf > kMaxTraceHosts) return SDSO_E_INVALID;
  TraceState* t = ctx->trace;
  if (!pts && n > t->pool_n) return fail(ctx, SDSO_E_STATE, "trace: more points than the resident pool holds (sdso_immature_upload)");
  if (n == 0) return SDSO_OK;
  int rc = ensure_pts(ctx, pts ? n : 1);
  if (rc) return rc;
  rc = ensure_intensity_plane(ctx, ctx->frames[frame]);
  if (rc) return rc;
  TraceParams T{};
  fill_settings(ctx, T);
  T.tex = ctx->frames[frame].tex[0]; T.img = ctx->frames[frame].image;
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_xf, xf, (size_t)nxf * sizeof(TraceXf), cudaMemcpyHostToDevice, st));
  const int* d_of = nullptr;
  if (xf_of) {
    if (n > t->xf_of_cap) {
      if (t->d_xf_of) cudaFree(t->d_xf_of);
      t->d_xf_of = nullptr;
      const int cap = n < 4096 ? 4096 : n;
      SDSO_CUDA(ctx, cudaMalloc(&t->d_xf_of, (size_t)cap * sizeof(int)));
      t->xf_of_cap = cap;
    }
    SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_xf_of, xf_of, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    d_of = t->d_xf_of;
  }
  sdso_immature_point* d = pts ? t->d_pts : t->d_pool;
  if (pts) SDSO_CUDA(ctx, cudaMemcpyAsync(d, pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyHostToDevice, st));
  const int blocks = (n + kTraceWarps - 1) / kTraceWarps;
  if (stereo) trace_kernel<true><<<blocks, 32 * kTraceWarps, 0, st>>>(T, t->d_xf, d_of, d, n);
  else trace_kernel<false><<<blocks, 32 * kTraceWarps, 0, st>>>(T, t->d_xf, d_of, d, n);
  SDSO_CHECK_LAUNCH(ctx);
  if (pts) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(pts, d, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
    if (status) for (int i = 0; i < n; i++) status[i] = pts[i].lastTraceStatus;
  } else if (status) {
    if (n > t->cap_uv) {
      if (t->d_uv) cudaFree(t->d_uv);
      if (t->d_ok) cudaFree(t->d_ok);
      t->d_uv = nullptr; t->d_ok = nullptr;
      SDSO_CUDA(ctx, cudaMalloc(&t->d_uv, (size_t)n * 2 * sizeof(float)));
      SDSO_CUDA(ctx, cudaMalloc(&t->d_ok, (size_t)n * sizeof(int)));
      t->cap_uv = n;
    }
    trace_status_kernel<<<(n + 255) / 256, 256, 0, st>>>(d, n, t->d_ok);
    SDSO_CHECK_LAUNCH(ctx);
    SDSO_CUDA(ctx, cudaMemcpyAsync(status, t->d_ok, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return SDSO_OK;
}

static void stereo_xf(const sdso_ctx* ctx, const float K[9], int mode_right, TraceXf& X) {
  const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++) X.KRKi[i] = I3[i];
  const float bl[3] = {mode_right ? -ctx->baseline : ctx->baseline, 0, 0};
  for (int r = 0; r < 3; r++) X.Kt[r] = K[r * 3] * bl[0] + K[r * 3 + 1] * bl[1] + K[r * 3 + 2] * bl[2];
  X.aff[0] = 1; X.aff[1] = 0;
  X.bf = -K[0] * bl[0];
  X.pad = 0;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_immature_init(sdso_ctx* ctx, int host_frame, int n, const float* uv, sdso_immature_point* out, int* ok) {
  sdso::enter(ctx);
  if (!ctx || host_frame < 0 || host_frame >= (int)ctx->frames.size() || !ctx->frames[host_frame].valid) return SDSO_E_INVALID;
  if (n < 0 || (n > 0 && (!uv || !out))) return SDSO_E_INVALID;
  if (n == 0) return SDSO_OK;
  int rc = ensure_pts(ctx, n);
  if (rc) return rc;
  TraceState* t = ctx->trace;
  if (n > t->cap_uv) {
    if (t->d_uv) cudaFree(t->d_uv);
    if (t->d_ok) cudaFree(t->d_ok);
    t->d_uv = nullptr; t->d_ok = nullptr;
    const int cap = n < 4096 ? 4096 : n;
    SDSO_CUDA(ctx, cudaMalloc(&t->d_uv, (size_t)cap * 2 * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&t->d_ok, (size_t)cap * sizeof(int)));
    t->cap_uv = cap;
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_uv, uv, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  immature_init_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->frames[host_frame].tex[0], ctx->G.w[0], t->d_uv, n, t->d_pts, t->d_ok,
                                                                  ctx->S.outlierTHSumComponent, ctx->S.outlierTH, ctx->S.overallEnergyTHWeight);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(out, t->d_pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int> okv(n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(okv.data(), t->d_ok, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ok) for (int i = 0; i < n; i++) ok[i] = okv[i];
  return SDSO_OK;
}

int sdso_trace_on(sdso_ctx* ctx, int frame, const float KRKi[9], const float Kt[3], const float aff[2], int n, sdso_immature_point* pts, int* status) {
  sdso::enter(ctx);
  if (!ctx || !KRKi || !Kt || !aff || (n > 0 && !pts)) return SDSO_E_INVALID;
  TraceXf X{};
  for (int i = 0; i < 9; i++) X.KRKi[i] = KRKi[i];
  for (int i = 0; i < 3; i++) X.Kt[i] = Kt[i];
  X.aff[0] = aff[0]; X.aff[1] = aff[1];
  return run_trace(ctx, frame, &X, 1, nullptr, false, n, pts, status);
}

int sdso_trace_stereo(sdso_ctx* ctx, int frame, const float K[9], int mode_right, int n, sdso_immature_point* pts, int* status) {
  sdso::enter(ctx);
  if (!ctx || !K || (n > 0 && !pts)) return SDSO_E_INVALID;
  TraceXf X{};
  stereo_xf(ctx, K, mode_right, X);
  return run_trace(ctx, frame, &X, 1, nullptr, true, n, pts, status);
}

int sdso_trace_on_hosts(sdso_ctx* ctx, int frame, int n_hosts, const float* KRKi, const float* Kt, const float* aff, int n, const int* host_of_point,
                        sdso_immature_point* pts, int* status) {
  sdso::enter(ctx);
  if (!ctx || !KRKi || !Kt || !aff || n_hosts < 1 || n_hosts > kMaxTraceHosts || (n > 0 && !host_of_point)) return SDSO_E_INVALID;
  for (int i = 0; i < n; i++) if (host_of_point[i] < 0 || host_of_point[i] >= n_hosts) return fail(ctx, SDSO_E_INVALID, "trace: host index out of range");
  TraceXf X[kMaxTraceHosts];
  memset(X, 0, sizeof(X));
  for (int h = 0; h < n_hosts; h++) {
    for (int i = 0; i < 9; i++) X[h].KRKi[i] = KRKi[9 * h + i];
    for (int i = 0; i < 3; i++) X[h].Kt[i] = Kt[3 * h + i];
    X[h].aff[0] = aff[2 * h]; X[h].aff[1] = aff[2 * h + 1];
  }
  return run_trace(ctx, frame, X, n_hosts, host_of_point, false, n, pts, status);
}

int sdso_trace_stereo_resident(sdso_ctx* ctx, int frame, const float K[9], int mode_right, int n, int* status) {
  sdso::enter(ctx);
  if (!ctx || !K) return SDSO_E_INVALID;
  TraceXf X{};
  stereo_xf(ctx, K, mode_right, X);
  return run_trace(ctx, frame, &X, 1, nullptr, true, n, nullptr, status);
}

int sdso_immature_upload(sdso_ctx* ctx, int n, const sdso_immature_point* pts) {
  sdso::enter(ctx);
  if (!ctx || n < 0 || (n > 0 && !pts)) return SDSO_E_INVALID;
  TraceState* t = ctx->trace;
  if (n > t->pool_cap) {
    if (t->d_pool) cudaFree(t->d_pool);
    t->d_pool = nullptr; t->pool_cap = 0;
    const int cap = n < 4096 ? 4096 : n;
    SDSO_CUDA(ctx, cudaMalloc(&t->d_pool, (size_t)cap * sizeof(sdso_immature_point)));
    t->pool_cap = cap;
  }
  if (n > 0) SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_pool, pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the caller's buffer may be reused
  t->pool_n = n;
  return SDSO_OK;
}

int sdso_immature_download(sdso_ctx* ctx, int first, int n, sdso_immature_point* pts) {
  sdso::enter(ctx);
  if (!ctx || first < 0 || n < 0 || (n > 0 && !pts)) return SDSO_E_INVALID;
  TraceState* t = ctx->trace;
  if (first + n > t->pool_n) return fail(ctx, SDSO_E_INVALID, "immature_download: range exceeds the resident pool");
  if (n > 0) SDSO_CUDA(ctx, cudaMemcpyAsync(pts, t->d_pool + first, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

}  // extern "C"
