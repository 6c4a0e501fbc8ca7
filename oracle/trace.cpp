// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// Epipolar search for immature points: ImmaturePoint constructor (ImmaturePoint.cpp:33-88), traceOn (:459-828),
// traceStereo (:94-451) with its g2o Gauss-Newton refinement over EdgeTracePointUVDSO / VertexUVDSO
// (dso_g2o_edge.cpp:571-619, dso_g2o_vertex.cpp:73-88) restated per SURVEY.md Appendix A.10 / C (g2o is not
// in the reference tree: "parity unpinned" for the refinement driver; the discrete search is plain reference code).
#include "oracle_trace.hpp"

namespace orc {

bool immatureInit(const GlobalCalib& G, const Settings& S, const Frame& host, float u, float v, ImmaturePoint& p) {
  p = ImmaturePoint();
  p.u = u; p.v = v;
  p.idepth_min = 0; p.idepth_max = NAN; p.lastTraceStatus = IPS_UNINITIALIZED;
  // what every caller sets right after construction (FullSystem.cpp:582-585, CoarseInitializer.cpp:899-902)
  p.u_stereo = u; p.v_stereo = v; p.idepth_min_stereo = 0; p.idepth_max_stereo = NAN;
  const float* dI = host.dIp[0].data();
  for (int i = 0; i < 4; i++) p.gradH[i] = 0;
  for (int idx = 0; idx < patternNum; idx++) {
    float ptc[3];
    getInterpolatedElement33BiLin(dI, u + patternP[idx][0], v + patternP[idx][1], G.w[0], ptc);
    p.color[idx] = ptc[0];
    if (!std::isfinite(p.color[idx])) { p.energyTH = NAN; return false; }
    p.gradH[0] += ptc[1] * ptc[1]; p.gradH[1] += ptc[1] * ptc[2]; p.gradH[2] += ptc[2] * ptc[1]; p.gradH[3] += ptc[2] * ptc[2];
    p.weights[idx] = sqrtf(S.outlierTHSumComponent / (S.outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
  }
  p.energyTH = patternNum * S.outlierTH;
  p.energyTH *= S.overallEnergyTHWeight * S.overallEnergyTHWeight;
  p.quality = 10000;
  return true;
}

namespace {
struct Search {  // what STEP1-3 of traceOn / traceStereo leave behind for the refinement
  float pr[3], dx, dy, errorInPixel, bestU, bestV, bestEnergy;
  float rot[8][2];
  int numSteps, bestIdx;
};
inline bool inside(float u, float v, const GlobalCalib& G) { return u > 4 && v > 4 && u < G.w[0] - 5 && v < G.h[0] - 5; }

// Shared STEP1-3 (ImmaturePoint.cpp:493-705 == :112-303 with the stereo members). Returns -1 to continue with the
// refinement, otherwise the status to return. `stereo` selects the early-return conventions of traceStereo.
int searchSegment(const GlobalCalib& G, const Settings& S, ImmaturePoint& p, const Frame& frame, const float KRKi[9], const float Kt[3],
                  const float aff[2], bool stereo, float u0, float v0, float id_min, float id_max, Search& s) {
  const float maxPixSearch = (G.w[0] + G.h[0]) * S.maxPixSearch;
  for (int k = 0; k < 3; k++) s.pr[k] = KRKi[k * 3] * u0 + KRKi[k * 3 + 1] * v0 + KRKi[k * 3 + 2] * 1.0f;
  float ptpMin[3];
  for (int k = 0; k < 3; k++) ptpMin[k] = s.pr[k] + Kt[k] * id_min;
  float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
  auto oob = [&]() { p.lastTraceUV[0] = p.lastTraceUV[1] = -1; p.lastTracePixelInterval = 0; return (int)(p.lastTraceStatus = IPS_OOB); };
  if (!inside(uMin, vMin, G)) return oob();
  float dist, uMax, vMax, ptpMax[3];
  if (std::isfinite(id_max)) {
    for (int k = 0; k < 3; k++) ptpMax[k] = s.pr[k] + Kt[k] * id_max;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    if (!inside(uMax, vMax, G)) return oob();
    dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
    dist = sqrtf(dist);
    if (dist < S.trace_slackInterval) {
      if (!stereo) { p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f; p.lastTracePixelInterval = dist; }
      return p.lastTraceStatus = IPS_SKIPPED;
    }
  } else {
    dist = maxPixSearch;
    for (int k = 0; k < 3; k++) ptpMax[k] = s.pr[k] + Kt[k] * 0.01f;
    uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
    float dx = uMax - uMin, dy = vMax - vMin;
    float d = 1.0f / sqrtf(dx * dx + dy * dy);
    uMax = uMin + dist * dx * d;
    vMax = vMin + dist * dy * d;
    if (!inside(uMax, vMax, G)) return oob();
  }
  // scale-change test; traceStereo tests the TEMPORAL idepth_min member here (ImmaturePoint.cpp:197), as written
  if (!(p.idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) return oob();
  float dx = S.trace_stepsize * (uMax - uMin), dy = S.trace_stepsize * (vMax - vMin);
  const float* g = p.gradH;
  // (v^T gradH) v, evaluated left to right as Eigen does for `v.transpose() * gradH * v`
  float a = (dx * g[0] + dy * g[2]) * dx + (dx * g[1] + dy * g[3]) * dy;
  float b = (dy * g[0] + (-dx) * g[2]) * dy + (dy * g[1] + (-dx) * g[3]) * (-dx);
  float errorInPixel = 0.2f + 0.2f * (a + b) / a;
  if (errorInPixel * S.trace_minImprovementFactor > dist && std::isfinite(id_max)) {
    if (!stereo) { p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f; p.lastTracePixelInterval = dist; }
    return p.lastTraceStatus = IPS_BADCONDITION;
  }
  if (errorInPixel > 10) errorInPixel = 10;
  dx /= dist; dy /= dist;
  if (dist > maxPixSearch) { uMax = uMin + maxPixSearch * dx; vMax = vMin + maxPixSearch * dy; dist = maxPixSearch; }
  int numSteps = 1.9999f + dist / S.trace_stepsize;
  float randShift = uMin * 1000 - floorf(uMin * 1000);
  float ptx = uMin - randShift * dx, pty = vMin - randShift * dy;
  for (int idx = 0; idx < patternNum; idx++) {
    s.rot[idx][0] = KRKi[0] * patternP[idx][0] + KRKi[1] * patternP[idx][1];
    s.rot[idx][1] = KRKi[3] * patternP[idx][0] + KRKi[4] * patternP[idx][1];
  }
  if (!std::isfinite(dx) || !std::isfinite(dy)) return oob();
  float errors[100];
  float bestU = 0, bestV = 0, bestEnergy = 1e10;
  int bestIdx = -1;
  if (numSteps >= 100) numSteps = 99;
  const float* dI = frame.dIp[0].data();
  for (int i = 0; i < numSteps; i++) {
    float energy = 0;
    for (int idx = 0; idx < patternNum; idx++) {
      float hitColor = getInterpolatedElement31(dI, (float)(ptx + s.rot[idx][0]), (float)(pty + s.rot[idx][1]), G.w[0]);
      if (!std::isfinite(hitColor)) { energy += 1e5; continue; }
      float residual = hitColor - (float)(aff[0] * p.color[idx] + aff[1]);
      float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
      energy += hw * residual * residual * (2 - hw);
    }
    errors[i] = energy;
    if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
    ptx += dx; pty += dy;
  }
  float secondBest = 1e10;
  for (int i = 0; i < numSteps; i++)
    if ((i < bestIdx - S.minTraceTestRadius || i > bestIdx + S.minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
  float newQuality = secondBest / bestEnergy;
  if (newQuality < p.quality || numSteps > 10) p.quality = newQuality;
  s.dx = dx; s.dy = dy; s.errorInPixel = errorInPixel; s.bestU = bestU; s.bestV = bestV; s.bestEnergy = bestEnergy;
  s.numSteps = numSteps; s.bestIdx = bestIdx;
  p.numSteps = numSteps; p.bestIdx = bestIdx;
  return -1;
}

// STEP5 (ImmaturePoint.cpp:795-827 == :421-450): new inverse-depth interval from bestU/V +- errorInPixel
int newInterval(ImmaturePoint& p, const Search& s, const float Kt[3], float bestU, float bestV, float& id_min, float& id_max) {
  const float dx = s.dx, dy = s.dy, e = s.errorInPixel;
  if (dx * dx > dy * dy) {
    id_min = (s.pr[2] * (bestU - e * dx) - s.pr[0]) / (Kt[0] - Kt[2] * (bestU - e * dx));
    id_max = (s.pr[2] * (bestU + e * dx) - s.pr[0]) / (Kt[0] - Kt[2] * (bestU + e * dx));
  } else {
    id_min = (s.pr[2] * (bestV - e * dy) - s.pr[1]) / (Kt[1] - Kt[2] * (bestV - e * dy));
    id_max = (s.pr[2] * (bestV + e * dy) - s.pr[1]) / (Kt[1] - Kt[2] * (bestV + e * dy));
  }
  if (id_min > id_max) std::swap(id_min, id_max);
  if (!std::isfinite(id_min) || !std::isfinite(id_max) || (id_max < 0)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
    return p.lastTraceStatus = IPS_OUTLIER;
  }
  p.lastTracePixelInterval = 2 * e;
  p.lastTraceUV[0] = bestU; p.lastTraceUV[1] = bestV;
  return p.lastTraceStatus = IPS_GOOD;
}
int energyOutlier(ImmaturePoint& p) {  // :781-793 == :412-419
  p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
  if (p.lastTraceStatus == IPS_OUTLIER) return p.lastTraceStatus = IPS_OOB;
  return p.lastTraceStatus = IPS_OUTLIER;
}
}  // namespace

int traceOn(const GlobalCalib& G, const Settings& S, ImmaturePoint& p, const Frame& frame, const float KRKi[9], const float Kt[3], const float aff[2]) {
  p.numSteps = 0; p.bestIdx = -1;
  if (p.lastTraceStatus == IPS_OOB) return p.lastTraceStatus;
  Search s;
  int st = searchSegment(G, S, p, frame, KRKi, Kt, aff, false, p.u, p.v, p.idepth_min, p.idepth_max, s);
  if (st >= 0) return st;
  const float dx = s.dx, dy = s.dy;
  float bestU = s.bestU, bestV = s.bestV, bestEnergy = s.bestEnergy;
  // STEP4: GN refinement along the line (:707-779)
  float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
  if (S.trace_GNIterations > 0) bestEnergy = 1e5;
  const float* dI = frame.dIp[0].data();
  for (int it = 0; it < S.trace_GNIterations; it++) {
    float H = 1, b = 0, energy = 0;
    for (int idx = 0; idx < patternNum; idx++) {
      float hit[3];
      getInterpolatedElement33(dI, (float)(bestU + s.rot[idx][0]), (float)(bestV + s.rot[idx][1]), G.w[0], hit);
      if (!std::isfinite((float)hit[0])) { energy += 1e5; continue; }
      float residual = hit[0] - (aff[0] * p.color[idx] + aff[1]);
      float dResdDist = dx * hit[1] + dy * hit[2];
      float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
      H += hw * dResdDist * dResdDist;
      b += hw * residual * dResdDist;
      energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
    }
    if (energy > bestEnergy) {
      stepBack *= 0.5;
      bestU = uBak + stepBack * dx;
      bestV = vBak + stepBack * dy;
    } else {
      float step = -gnstepsize * b / H;
      if (step < -0.5) step = -0.5;
      else if (step > 0.5) step = 0.5;
      if (!std::isfinite(step)) step = 0;
      uBak = bestU; vBak = bestV; stepBack = step;
      bestU += step * dx; bestV += step * dy;
      bestEnergy = energy;
    }
    if (fabsf(stepBack) < S.trace_GNThreshold) break;
  }
  if (!(bestEnergy < p.energyTH * S.trace_extraSlackOnTH)) return energyOutlier(p);
  return newInterval(p, s, Kt, bestU, bestV, p.idepth_min, p.idepth_max);
}

int traceStereo(const GlobalCalib& G, const Settings& S, ImmaturePoint& p, const Frame& frame, const float K[9], bool mode_right) {
  p.numSteps = 0; p.bestIdx = -1;
  const float KRKi[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float bl[3] = {mode_right ? -G.baseline : G.baseline, 0, 0};
  float Kt[3];
  for (int r = 0; r < 3; r++) Kt[r] = K[r * 3] * bl[0] + K[r * 3 + 1] * bl[1] + K[r * 3 + 2] * bl[2];
  const float aff[2] = {1, 0};
  const float bf = -K[0] * bl[0];
  Search s;
  int st = searchSegment(G, S, p, frame, KRKi, Kt, aff, true, p.u_stereo, p.v_stereo, p.idepth_min_stereo, p.idepth_max_stereo, s);
  if (st >= 0) return st;
  // g2o Gauss-Newton over EdgeTracePointUVDSO (:309-411): per iteration 8 more edges join the same optimizer; all edges see
  // the same vertex, so the duplicates scale H and b alike. H has no +1 damping, the step is always applied (clamped to
  // +-0.5 by VertexUVDSO::oplusImpl), bestEnergy is the minimum of the pre-step energies, no early break.
  const double dxd = s.dx, dyd = s.dy;  // SetDxDy(dx, dy): float -> double
  double U = s.bestU, V = s.bestV;      // setEstimate(Vec2(bestU, bestV))
  float bestEnergy = s.bestEnergy;
  if (S.trace_GNIterations > 0) bestEnergy = 1e5;
  const float* dI = frame.dIp[0].data();
  const int wl = G.w[0] - 3, hl = G.h[0] - 3;
  for (int it = 0; it < S.trace_GNIterations; it++) {
    float energy = 0;
    double H = 0, b = 0;
    for (int idx = 0; idx < patternNum; idx++) {
      double err = 0, J = 0;  // fresh edge: _error / _jacobianOplusXi taken as 0 where the edge leaves them untouched
      const bool outside = (U - 2) < 0 || (U + 3) > wl || (V - 2) < 0 || (V + 3) > hl;
      if (!outside) {
        float hit[3];
        getInterpolatedElement33(dI, (float)(U + s.rot[idx][0]), (float)(V + s.rot[idx][1]), G.w[0], hit);
        if (std::isfinite(hit[0])) {
          err = hit[0] - ((double)aff[0] * (double)p.color[idx] + (double)aff[1]);
          J = dxd * hit[1] + dyd * hit[2];
        }
      }
      float residual = (float)err;
      float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
      energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
      // g2o robustified quadratic form: rho' = 1 if e^2 <= delta^2 else delta / |e|
      const double e2 = err * err, dlt = S.huberTH;
      const double rho1 = (e2 <= dlt * dlt) ? 1.0 : dlt / std::sqrt(e2);
      H += rho1 * J * J;
      b -= rho1 * J * err;
    }
    if (H > 0 && std::isfinite(H)) {  // LinearSolverEigen (LLT) fails on a non-positive pivot: no update then
      double update = b / H;
      if (update < -0.5) update = -0.5;
      else if (update > 0.5) update = 0.5;
      else if (!std::isfinite(update)) update = 0;
      U += update * dxd; V += update * dyd;
    }
    if (!(energy > bestEnergy)) bestEnergy = energy;
  }
  const float bestU = (float)U, bestV = (float)V;
  if (!(bestEnergy < p.energyTH * S.trace_extraSlackOnTH)) return energyOutlier(p);
  int r = newInterval(p, s, Kt, bestU, bestV, p.idepth_min_stereo, p.idepth_max_stereo);
  if (r == IPS_GOOD) p.idepth_stereo = (p.u_stereo - bestU) / bf;
  return r;
}

}  // namespace orc
