// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// D4: point activation. ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:886-985) + FullSystem::optimizeImmaturePoint
// (FullSystemOptPoint.cpp:52-238), both variants:
//   variant 0 (SSE): the original DSO body kept in the comments (ImmaturePoint.cpp:942-970, FullSystemOptPoint.cpp:124-171):
//                    Hdd/bd accumulation, 3 LM iterations on the inverse depth, lambda 0.1 x0.5 / x5, Hdd >= minIdepthH_act.
//   variant 1 (g2o): the live code: EdgePointActivationIdepthDSO (dso_g2o_edge.cpp:504-567) projects ONCE at construction, so the
//                    g2o LM sees a constant chi2, every trial has gain 0 and is rejected: the activated
//                    inverse depth is the initial 0.5 (idepth_min + idepth_max); IN / OUTLIER comes from that one evaluation
//                    with slack 1000; an out-of-image pixel contributes error 0.
#include "oracle_ba.hpp"
#include "oracle_trace.hpp"

namespace orc {

namespace {
struct TmpRes { int state_state = RS_IN, state_NewState = RS_OUTLIER; double state_energy = 0, state_NewEnergy = 0; int target = 0; };

// SSE variant of linearizeResidual
double linResSSE(const BAWindow& W, const ImmaturePoint& p, int host, float slack, TmpRes& tr, float& Hdd, float& bd, float idepth) {
  if (tr.state_state == RS_OOB) { tr.state_NewState = RS_OOB; return tr.state_energy; }
  const FrameFramePrecalc& pc = W.precalc[(size_t)host * W.n() + tr.target];
  const float* dIl = W.frames[tr.target].img->dIp[0].data();
  const CalibHessian& H = W.HCalib;
  float energyLeft = 0;
  for (int idx = 0; idx < patternNum; idx++) {
    const int dx = patternP[idx][0], dy = patternP[idx][1];
    const float K0 = (p.u + dx - H.cxl) * H.fxli, K1 = (p.v + dy - H.cyl) * H.fyli;
    float ptp[3];
    for (int k = 0; k < 3; k++) ptp[k] = (pc.PRE_RTll[k * 3] * K0 + pc.PRE_RTll[k * 3 + 1] * K1 + pc.PRE_RTll[k * 3 + 2] * 1.0f) + pc.PRE_tTll[k] * idepth;
    const float drescale = 1.0f / ptp[2];
    bool ok = drescale > 0;
    float u = 0, v = 0, Ku = 0, Kv = 0;
    if (ok) { u = ptp[0] * drescale; v = ptp[1] * drescale; Ku = u * H.fxl + H.cxl; Kv = v * H.fyl + H.cyl; ok = Ku > 1.1f && Kv > 1.1f && Ku < W.G->wM3G && Kv < W.G->hM3G; }
    if (!ok) { tr.state_NewState = RS_OOB; return tr.state_energy; }
    float hit[3];
    getInterpolatedElement33(dIl, Ku, Kv, W.G->w[0], hit);
    if (!std::isfinite((float)hit[0])) { tr.state_NewState = RS_OOB; return tr.state_energy; }
    const float residual = hit[0] - (pc.PRE_aff_mode[0] * p.color[idx] + pc.PRE_aff_mode[1]);
    float hw = fabsf(residual) < W.S.huberTH ? 1 : W.S.huberTH / fabsf(residual);
    energyLeft += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
    const float dxInterp = hit[1] * H.fxl, dyInterp = hit[2] * H.fyl;
    const float d_idepth = (dxInterp * drescale * (pc.PRE_tTll[0] - pc.PRE_tTll[2] * u) + dyInterp * drescale * (pc.PRE_tTll[1] - pc.PRE_tTll[2] * v)) * SCALE_IDEPTH;
    hw *= p.weights[idx] * p.weights[idx];
    Hdd += (hw * d_idepth) * d_idepth;
    bd += (hw * residual) * d_idepth;
  }
  if (energyLeft > p.energyTH * slack) { energyLeft = p.energyTH * slack; tr.state_NewState = RS_OUTLIER; }
  else tr.state_NewState = RS_IN;
  tr.state_NewEnergy = energyLeft;
  return energyLeft;
}

// live variant: one evaluation with the projection of the edge constructor
double linResG2O(const BAWindow& W, const ImmaturePoint& p, int host, float slack, TmpRes& tr, double idepth) {
  if (tr.state_state == RS_OOB) { tr.state_NewState = RS_OOB; return tr.state_energy; }
  const FrameFramePrecalc& pc = W.precalc[(size_t)host * W.n() + tr.target];
  const float* dIl = W.frames[tr.target].img->dIp[0].data();
  const CalibHessian& H = W.HCalib;
  const int wl = W.G->w[0] - 3, hl = W.G->h[0] - 3;
  float energyLeft = 0;
  for (int idx = 0; idx < patternNum; idx++) {
    const float up = p.u + patternP[idx][0], vp = p.v + patternP[idx][1];
    const float K0 = (up - H.cxl) * H.fxli, K1 = (vp - H.cyl) * H.fyli;
    const float idf = (float)idepth;
    float ptp[3];
    for (int k = 0; k < 3; k++) ptp[k] = (pc.PRE_RTll[k * 3] * K0 + pc.PRE_RTll[k * 3 + 1] * K1 + pc.PRE_RTll[k * 3 + 2] * 1.0f) + pc.PRE_tTll[k] * idf;
    const float drescale = 1.0f / ptp[2];
    double err = 0;
    if (drescale > 0) {  // (drescale <= 0 leaves Ku_/Kv_ unset in the reference; treated as out of the image)
      const float u = ptp[0] * drescale, v = ptp[1] * drescale;
      const float Ku = u * H.fxl + H.cxl, Kv = v * H.fyl + H.cyl;
      const bool outside = ((double)Ku - 2) < 0 || ((double)Ku + 3) > wl || ((double)Kv - 2) < 0 || ((double)Kv + 3) > hl;
      if (!outside) {
        float hit[3];
        getInterpolatedElement33(dIl, Ku, Kv, W.G->w[0], hit);
        if (std::isfinite((float)hit[0])) err = (double)hit[0] - ((double)pc.PRE_aff_mode[0] * (double)p.color[idx] + (double)pc.PRE_aff_mode[1]);  // _measurement is a double
      }
    }
    const float residual = (float)err;
    const float hw = fabsf(residual) < W.S.huberTH ? 1 : W.S.huberTH / fabsf(residual);
    energyLeft += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
  }
  if (energyLeft > p.energyTH * slack) { energyLeft = p.energyTH * slack; tr.state_NewState = RS_OUTLIER; }
  else tr.state_NewState = RS_IN;
  tr.state_NewEnergy = energyLeft;
  return energyLeft;
}
}  // namespace

// returns 1 = activated, 0 = not well constrained (skip), -1 = outlier / invalid. states[n] (RS_* per frame, -1 for the host)
int activatePoint(const BAWindow& W, int host, const ImmaturePoint& p, int variant, int minObs, float* idepth_out, int* states, float* energy_out) {
  const int nf = W.n();
  std::vector<TmpRes> res;
  for (int f = 0; f < nf; f++) { states[f] = -1; if (f != host) { TmpRes t; t.target = f; res.push_back(t); } }
  const int nres = (int)res.size();
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (p.idepth_max + p.idepth_min) * 0.5f;
  *idepth_out = currentIdepth; *energy_out = 0;
  if (variant == 1) {
    for (int i = 0; i < nres; i++) { lastEnergy += (float)linResG2O(W, p, host, 1000, res[i], (double)currentIdepth); res[i].state_state = res[i].state_NewState; res[i].state_energy = res[i].state_NewEnergy; }
  } else {
    for (int i = 0; i < nres; i++) { lastEnergy += (float)linResSSE(W, p, host, 1000, res[i], lastHdd, lastbd, currentIdepth); res[i].state_state = res[i].state_NewState; res[i].state_energy = res[i].state_NewEnergy; }
    if (!std::isfinite(lastEnergy) || lastHdd < W.S.minIdepthH_act) { *energy_out = lastEnergy; return 0; }
    float lambda = 0.1f;
    for (int iteration = 0; iteration < W.S.GNItsOnPointActivation; iteration++) {
      float Hh = lastHdd;
      Hh *= 1 + lambda;
      const float step = (1.0 / Hh) * lastbd;
      const float newIdepth = currentIdepth - step;
      float newHdd = 0, newbd = 0, newEnergy = 0;
      for (int i = 0; i < nres; i++) newEnergy += (float)linResSSE(W, p, host, 1, res[i], newHdd, newbd, newIdepth);
      if (!std::isfinite(lastEnergy) || newHdd < W.S.minIdepthH_act) { *energy_out = lastEnergy; *idepth_out = currentIdepth; return 0; }
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        for (int i = 0; i < nres; i++) { res[i].state_state = res[i].state_NewState; res[i].state_energy = res[i].state_NewEnergy; }
        lambda *= 0.5;
      } else lambda *= 5;
      if (fabsf(step) < 0.0001 * currentIdepth) break;
    }
  }
  *idepth_out = currentIdepth; *energy_out = lastEnergy;
  for (int i = 0; i < nres; i++) states[res[i].target] = res[i].state_state;
  if (!std::isfinite(currentIdepth)) return -1;
  int numGood = 0;
  for (int i = 0; i < nres; i++) if (res[i].state_state == RS_IN) numGood++;
  if (numGood < minObs) return -1;
  if (!std::isfinite(p.energyTH)) return -1;
  return 1;
}

}  // namespace orc
