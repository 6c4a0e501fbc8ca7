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
#include <vector>

namespace sdso {

enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };  // ImmaturePoint.h:50-56

struct TraceState {
  sdso_immature_point* d_pts = nullptr;
  int cap = 0;
  float* d_uv = nullptr;
  int* d_ok = nullptr;
  int cap_uv = 0;
};

struct TraceParams {
  const float4* tex;   // level-0 texels of the searched frame
  const float* img;    // level-0 intensity plane of the searched frame
  int w, h;
  float KRKi[9], Kt[3], aff[2];
  float bf;            // stereo: -K(0,0) * bl[0]
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

struct TraceShared {
  float pr[3], dx, dy, errorInPixel, ptx0, pty0;
  float rot[8][2];
  int numSteps, status;   // status >= 0: finished in the set-up
  float errors[100];
};

__device__ __forceinline__ bool inside_search(float u, float v, int w, int h) { return u > 4 && v > 4 && u < w - 5 && v < h - 5; }

template <bool STEREO>
__global__ void __launch_bounds__(128) trace_kernel(TraceParams T, sdso_immature_point* pts, int n) {
  __shared__ TraceShared S;
  sdso_immature_point& p = pts[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid == 0) {
    S.status = -1; S.numSteps = 0;
    p.numSteps = 0; p.bestIdx = -1;
    do {
      if (!STEREO && p.lastTraceStatus == IPS_OOB) { S.status = IPS_OOB; break; }
      const float u0 = STEREO ? p.u_stereo : p.u, v0 = STEREO ? p.v_stereo : p.v;
      const float id_min = STEREO ? p.idepth_min_stereo : p.idepth_min, id_max = STEREO ? p.idepth_max_stereo : p.idepth_max;
      float pr[3], ptpMin[3], ptpMax[3];
#pragma unroll
      for (int k = 0; k < 3; k++) pr[k] = T.KRKi[k * 3] * u0 + T.KRKi[k * 3 + 1] * v0 + T.KRKi[k * 3 + 2] * 1.0f;
#pragma unroll
      for (int k = 0; k < 3; k++) ptpMin[k] = pr[k] + T.Kt[k] * id_min;
      const float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
      bool oob = !inside_search(uMin, vMin, T.w, T.h);
      float dist = 0, uMax = 0, vMax = 0;
      if (!oob) {
        if (isfinite(id_max)) {
#pragma unroll
          for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + T.Kt[k] * id_max;
          uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
          if (!inside_search(uMax, vMax, T.w, T.h)) oob = true;
          else {
            dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
            dist = sqrtf(dist);
            if (dist < T.slackInterval) {
              if (!STEREO) { p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f; p.lastTracePixelInterval = dist; }
              S.status = p.lastTraceStatus = IPS_SKIPPED;
              break;
            }
          }
        } else {
          dist = T.maxPixSearch;
#pragma unroll
          for (int k = 0; k < 3; k++) ptpMax[k] = pr[k] + T.Kt[k] * 0.01f;
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
      if (oob) { p.lastTraceUV[0] = p.lastTraceUV[1] = -1; p.lastTracePixelInterval = 0; S.status = p.lastTraceStatus = IPS_OOB; break; }
      float dx = T.stepsize * (uMax - uMin), dy = T.stepsize * (vMax - vMin);
      const float* g = p.gradH;
      const float a = (dx * g[0] + dy * g[2]) * dx + (dx * g[1] + dy * g[3]) * dy;
      const float b = (dy * g[0] + (-dx) * g[2]) * dy + (dy * g[1] + (-dx) * g[3]) * (-dx);
      float errorInPixel = 0.2f + 0.2f * (a + b) / a;
      if (errorInPixel * T.minImprovementFactor > dist && isfinite(id_max)) {
        if (!STEREO) { p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f; p.lastTracePixelInterval = dist; }
        S.status = p.lastTraceStatus = IPS_BADCONDITION;
        break;
      }
      if (errorInPixel > 10) errorInPixel = 10;
      dx /= dist; dy /= dist;
      if (dist > T.maxPixSearch) { uMax = uMin + T.maxPixSearch * dx; vMax = vMin + T.maxPixSearch * dy; dist = T.maxPixSearch; }
      int numSteps = (int)(1.9999f + dist / T.stepsize);
      const float randShift = uMin * 1000 - floorf(uMin * 1000);
      S.ptx0 = uMin - randShift * dx; S.pty0 = vMin - randShift * dy;
#pragma unroll
      for (int idx = 0; idx < 8; idx++) {
        S.rot[idx][0] = T.KRKi[0] * kPatternP[idx][0] + T.KRKi[1] * kPatternP[idx][1];
        S.rot[idx][1] = T.KRKi[3] * kPatternP[idx][0] + T.KRKi[4] * kPatternP[idx][1];
      }
      if (!isfinite(dx) || !isfinite(dy)) { p.lastTraceUV[0] = p.lastTraceUV[1] = -1; p.lastTracePixelInterval = 0; S.status = p.lastTraceStatus = IPS_OOB; break; }
      if (numSteps >= 100) numSteps = 99;
      S.numSteps = numSteps; S.dx = dx; S.dy = dy; S.errorInPixel = errorInPixel;
      S.pr[0] = pr[0]; S.pr[1] = pr[1]; S.pr[2] = pr[2];
    } while (false);
  }
  __syncthreads();
  if (S.status >= 0) return;
  // ---- STEP3: discrete search, one step per thread (:659-691) ----
  const int numSteps = S.numSteps;
  if (tid < numSteps) {
    float ptx = S.ptx0, pty = S.pty0;
    for (int k = 0; k < tid; k++) { ptx += S.dx; pty += S.dy; }  // the reference positions step i by i additions
    float hit[8];
#pragma unroll
    for (int idx = 0; idx < 8; idx++) hit[idx] = interp31(T.img, (float)(ptx + S.rot[idx][0]), (float)(pty + S.rot[idx][1]), T.w);
    float energy = 0;
#pragma unroll
    for (int idx = 0; idx < 8; idx++) {
      if (!isfinite(hit[idx])) { energy += 1e5f; continue; }
      const float residual = hit[idx] - (float)(T.aff[0] * p.color[idx] + T.aff[1]);
      const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
      energy += hw * residual * residual * (2 - hw);
    }
    S.errors[tid] = energy;
  }
  __syncthreads();
  if (tid != 0) return;
  const float dx = S.dx, dy = S.dy;
  float bestU = 0, bestV = 0, bestEnergy = 1e10f;
  int bestIdx = -1;
  {
    float ptx = S.ptx0, pty = S.pty0;
    for (int i = 0; i < numSteps; i++) {
      if (S.errors[i] < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = S.errors[i]; bestIdx = i; }
      ptx += dx; pty += dy;
    }
  }
  float secondBest = 1e10f;
  for (int i = 0; i < numSteps; i++)
    if ((i < bestIdx - T.minTraceTestRadius || i > bestIdx + T.minTraceTestRadius) && S.errors[i] < secondBest) secondBest = S.errors[i];
  const float newQuality = secondBest / bestEnergy;
  if (newQuality < p.quality || numSteps > 10) p.quality = newQuality;
  p.numSteps = numSteps; p.bestIdx = bestIdx;
  // ---- STEP4: refinement ----
  if (T.GNIterations > 0) bestEnergy = 1e5f;
  if (!STEREO) {  // traceOn (:707-779): damped (H = 1 + ..) GN with back-off
    float uBak = bestU, vBak = bestV, stepBack = 0;
    const float gnstepsize = 1;
    for (int it = 0; it < T.GNIterations; it++) {
      float H = 1, b = 0, energy = 0;
      float3 hit[8];
#pragma unroll
      for (int idx = 0; idx < 8; idx++) hit[idx] = interp33(T.tex, (float)(bestU + S.rot[idx][0]), (float)(bestV + S.rot[idx][1]), T.w);
#pragma unroll
      for (int idx = 0; idx < 8; idx++) {
        if (!isfinite(hit[idx].x)) { energy += 1e5f; continue; }
        const float residual = hit[idx].x - (T.aff[0] * p.color[idx] + T.aff[1]);
        const float dResdDist = dx * hit[idx].y + dy * hit[idx].z;
        const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
        H += hw * dResdDist * dResdDist;
        b += hw * residual * dResdDist;
        energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
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
      for (int idx = 0; idx < 8; idx++) {
        double err = 0, J = 0;
        if (!outside) {
          const float3 hit = interp33(T.tex, (float)(U + S.rot[idx][0]), (float)(V + S.rot[idx][1]), T.w);
          if (isfinite(hit.x)) {
            err = hit.x - ((double)T.aff[0] * (double)p.color[idx] + (double)T.aff[1]);
            J = dxd * hit.y + dyd * hit.z;
          }
        }
        const float residual = (float)err;
        const float hw = fabsf(residual) < T.huberTH ? 1 : T.huberTH / fabsf(residual);
        energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
        const double e2 = err * err, dlt = T.huberTH;
        const double rho1 = (e2 <= dlt * dlt) ? 1.0 : dlt / sqrt(e2);  // RobustKernelHuber, first derivative
        H += rho1 * J * J;
        b -= rho1 * J * err;
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
  // ---- energy-based outlier (:781-793) ----
  if (!(bestEnergy < p.energyTH * T.extraSlackOnTH)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
    p.lastTraceStatus = (p.lastTraceStatus == IPS_OUTLIER) ? IPS_OOB : IPS_OUTLIER;
    return;
  }
  // ---- STEP5: new interval (:795-827) ----
  const float e = S.errorInPixel;
  float id_min, id_max;
  if (dx * dx > dy * dy) {
    id_min = (S.pr[2] * (bestU - e * dx) - S.pr[0]) / (T.Kt[0] - T.Kt[2] * (bestU - e * dx));
    id_max = (S.pr[2] * (bestU + e * dx) - S.pr[0]) / (T.Kt[0] - T.Kt[2] * (bestU + e * dx));
  } else {
    id_min = (S.pr[2] * (bestV - e * dy) - S.pr[1]) / (T.Kt[1] - T.Kt[2] * (bestV - e * dy));
    id_max = (S.pr[2] * (bestV + e * dy) - S.pr[1]) / (T.Kt[1] - T.Kt[2] * (bestV + e * dy));
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
  if (STEREO) p.idepth_stereo = (p.u_stereo - bestU) / T.bf;
  p.lastTraceStatus = IPS_GOOD;
}

int trace_create(sdso_ctx* ctx) { ctx->trace = new TraceState(); return SDSO_OK; }
void trace_destroy(sdso_ctx* ctx) {
  if (!ctx->trace) return;
  if (ctx->trace->d_pts) cudaFree(ctx->trace->d_pts);
  if (ctx->trace->d_uv) cudaFree(ctx->trace->d_uv);
  if (ctx->trace->d_ok) cudaFree(ctx->trace->d_ok);
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

static int run_trace(sdso_ctx* ctx, int frame, TraceParams& T, bool stereo, int n, sdso_immature_point* pts, int* status) {
  if (frame < 0 || frame >= (int)ctx->frames.size() || !ctx->frames[frame].valid) return fail(ctx, SDSO_E_INVALID, "trace: invalid frame");
  if (n < 0 || (n > 0 && !pts)) return SDSO_E_INVALID;
  if (n == 0) return SDSO_OK;
  int rc = ensure_pts(ctx, n);
  if (rc) return rc;
  TraceState* t = ctx->trace;
  rc = ensure_intensity_plane(ctx, ctx->frames[frame]);
  if (rc) return rc;
  T.tex = ctx->frames[frame].tex[0]; T.img = ctx->frames[frame].image;
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_pts, pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyHostToDevice, ctx->stream));
  if (stereo) trace_kernel<true><<<n, 128, 0, ctx->stream>>>(T, t->d_pts, n);
  else trace_kernel<false><<<n, 128, 0, ctx->stream>>>(T, t->d_pts, n);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(pts, t->d_pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (status) for (int i = 0; i < n; i++) status[i] = pts[i].lastTraceStatus;
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_immature_init(sdso_ctx* ctx, int host_frame, int n, const float* uv, sdso_immature_point* out, int* ok) {
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
  if (!ctx || !KRKi || !Kt || !aff) return SDSO_E_INVALID;
  TraceParams T{};
  fill_settings(ctx, T);
  for (int i = 0; i < 9; i++) T.KRKi[i] = KRKi[i];
  for (int i = 0; i < 3; i++) T.Kt[i] = Kt[i];
  T.aff[0] = aff[0]; T.aff[1] = aff[1];
  T.bf = 0;
  return run_trace(ctx, frame, T, false, n, pts, status);
}

int sdso_trace_stereo(sdso_ctx* ctx, int frame, const float K[9], int mode_right, int n, sdso_immature_point* pts, int* status) {
  if (!ctx || !K) return SDSO_E_INVALID;
  TraceParams T{};
  fill_settings(ctx, T);
  const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++) T.KRKi[i] = I3[i];
  const float bl[3] = {mode_right ? -ctx->baseline : ctx->baseline, 0, 0};
  for (int r = 0; r < 3; r++) T.Kt[r] = K[r * 3] * bl[0] + K[r * 3 + 1] * bl[1] + K[r * 3 + 2] * bl[2];
  T.aff[0] = 1; T.aff[1] = 0;
  T.bf = -K[0] * bl[0];
  return run_trace(ctx, frame, T, true, n, pts, status);
}

}  // extern "C"
