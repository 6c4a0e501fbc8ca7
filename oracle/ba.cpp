// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// Windowed BA, SSE path. File:line citations into /root/reference/src.
#include "oracle_ba.hpp"
#include <cassert>

namespace orc {

static inline void mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
    C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}

// ---- FrameHessian state handling ----------------------------------------------------------------
void BAFrame::setState(const double s[10]) {  // HessianBlocks.h:177-199
  for (int i = 0; i < 10; i++) state[i] = s[i];
  for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * state[i];
  for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * state[i];
  state_scaled[6] = SCALE_A * state[6]; state_scaled[7] = SCALE_B * state[7];
  state_scaled[8] = SCALE_A * state[8]; state_scaled[9] = SCALE_B * state[9];
  PRE_worldToCam = SE3::exp(state_scaled) * worldToCam_evalPT;
  PRE_camToWorld = PRE_worldToCam.inverse();
}
void BAFrame::setStateScaled(const double s[10]) {  // :201-215
  for (int i = 0; i < 10; i++) state_scaled[i] = s[i];
  for (int i = 0; i < 3; i++) state[i] = (1.0f / SCALE_XI_TRANS) * state_scaled[i];
  for (int i = 3; i < 6; i++) state[i] = (1.0f / SCALE_XI_ROT) * state_scaled[i];
  state[6] = (1.0f / SCALE_A) * state_scaled[6]; state[7] = (1.0f / SCALE_B) * state_scaled[7];
  state[8] = (1.0f / SCALE_A) * state_scaled[8]; state[9] = (1.0f / SCALE_B) * state_scaled[9];
  PRE_worldToCam = SE3::exp(state_scaled) * worldToCam_evalPT;
  PRE_camToWorld = PRE_worldToCam.inverse();
}
void BAFrame::setStateZero(const double s[10]) {  // HessianBlocks.cpp:78-123
  for (int i = 0; i < 10; i++) state_zero[i] = s[i];
  SE3 inv0 = worldToCam_evalPT.inverse();
  for (int i = 0; i < 6; i++) {
    double eps[6] = {0, 0, 0, 0, 0, 0};
    eps[i] = 1e-3;
    SE3 EepsP = SE3::exp(eps);
    eps[i] = -1e-3;
    SE3 EepsM = SE3::exp(eps);
    SE3 P = (worldToCam_evalPT * EepsP) * inv0;
    SE3 M = (worldToCam_evalPT * EepsM) * inv0;
    double lp[6], lm[6];
    P.log(lp); M.log(lm);
    for (int r = 0; r < 6; r++) nullspaces_pose[r * 6 + i] = (lp[r] - lm[r]) / (2e-3);
  }
  SE3 P = worldToCam_evalPT;
  for (int k = 0; k < 3; k++) P.t[k] *= 1.00001;
  P = P * inv0;
  SE3 M = worldToCam_evalPT;
  for (int k = 0; k < 3; k++) M.t[k] /= 1.00001;
  M = M * inv0;
  double lp[6], lm[6];
  P.log(lp); M.log(lm);
  for (int r = 0; r < 6; r++) nullspaces_scale[r] = (lp[r] - lm[r]) / (2e-3);
  for (int i = 0; i < 8; i++) nullspaces_affine[i] = 0;
  nullspaces_affine[0 * 2 + 0] = 1; nullspaces_affine[1 * 2 + 0] = 0;
  double ab0[2]; aff_g2l_0(ab0);
  nullspaces_affine[0 * 2 + 1] = 0;
  nullspaces_affine[1 * 2 + 1] = expf((float)ab0[0]) * img->ab_exposure;
}
void BAFrame::setEvalPT_scaled(const SE3& w2c, double a, double b) {  // HessianBlocks.h:223-231
  double init[10] = {0, 0, 0, 0, 0, 0, a, b, 0, 0};
  worldToCam_evalPT = w2c;
  setStateScaled(init);
  double st[10]; for (int i = 0; i < 10; i++) st[i] = state[i];
  setStateZero(st);
}

// ---- FrameFramePrecalc::set (HessianBlocks.cpp:206-242) ----------------------------------------
void BAWindow::setPrecalcValues() {
  const int nf = n();
  precalc.assign((size_t)nf * nf, FrameFramePrecalc());
  float K[9] = {HCalib.fxl, 0, HCalib.cxl, 0, HCalib.fyl, HCalib.cyl, 0, 0, 1};
  float Kinv[9]; inverse3f(K, Kinv);
  for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) {
    FrameFramePrecalc& p = precalc[(size_t)h * nf + t];
    const BAFrame& host = frames[h]; const BAFrame& target = frames[t];
    SE3 l0 = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
    double R0[9]; l0.rotationMatrix(R0);
    for (int i = 0; i < 9; i++) p.PRE_RTll_0[i] = (float)R0[i];
    for (int i = 0; i < 3; i++) p.PRE_tTll_0[i] = (float)l0.t[i];
    SE3 l = target.PRE_worldToCam * host.PRE_camToWorld;
    double R[9]; l.rotationMatrix(R);
    for (int i = 0; i < 9; i++) p.PRE_RTll[i] = (float)R[i];
    for (int i = 0; i < 3; i++) p.PRE_tTll[i] = (float)l.t[i];
    p.distanceLL = (float)std::sqrt(l.t[0] * l.t[0] + l.t[1] * l.t[1] + l.t[2] * l.t[2]);
    float KR[9]; mat33f_mul(K, p.PRE_RTll, KR);
    mat33f_mul(KR, Kinv, p.PRE_KRKiTll);
    mat33f_mul(p.PRE_RTll, Kinv, p.PRE_RKiTll);
    for (int r = 0; r < 3; r++) p.PRE_KtTll[r] = K[r * 3] * p.PRE_tTll[0] + K[r * 3 + 1] * p.PRE_tTll[1] + K[r * 3 + 2] * p.PRE_tTll[2];
    double ah[2], at[2]; host.aff_g2l(ah); target.aff_g2l(at);
    double ab[2]; affFromToVecExposure(host.img->ab_exposure, target.img->ab_exposure, ah[0], ah[1], at[0], at[1], ab);
    p.PRE_aff_mode[0] = (float)ab[0]; p.PRE_aff_mode[1] = (float)ab[1];
    double a0[2]; host.aff_g2l_0(a0);
    p.PRE_b0_mode = (float)a0[1];
  }
}

// ---- EnergyFunctional::setAdjointsF (EnergyFunctional.cpp:41-119) -------------------------------
void BAWindow::setAdjointsF() {
  const int nf = n();
  adHost.assign((size_t)nf * nf * 64, 0.0); adTarget.assign((size_t)nf * nf * 64, 0.0);
  adHostF.assign((size_t)nf * nf * 64, 0.f); adTargetF.assign((size_t)nf * nf * 64, 0.f);
  for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) {
    const BAFrame& host = frames[h]; const BAFrame& target = frames[t];
    SE3 hostToTarget = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
    double Adj[36]; hostToTarget.Adj(Adj);
    double AH[64] = {0}, AT[64] = {0};
    for (int i = 0; i < 8; i++) { AH[i * 8 + i] = 1; AT[i * 8 + i] = 1; }
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) AH[r * 8 + c] = -Adj[c * 6 + r];
    double a0h[2], a0t[2]; host.aff_g2l_0(a0h); target.aff_g2l_0(a0t);
    double ab[2]; affFromToVecExposure(host.img->ab_exposure, target.img->ab_exposure, a0h[0], a0h[1], a0t[0], a0t[1], ab);
    float affLL0 = (float)ab[0];
    AT[6 * 8 + 6] = -affLL0; AT[7 * 8 + 7] = -1;
    AH[6 * 8 + 6] = affLL0; AH[7 * 8 + 7] = affLL0;
    const double rs[8] = {SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_A, SCALE_B};
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { AH[r * 8 + c] *= rs[r]; AT[r * 8 + c] *= rs[r]; }
    size_t o = ((size_t)h + (size_t)t * nf) * 64;
    for (int i = 0; i < 64; i++) { adHost[o + i] = AH[i]; adTarget[o + i] = AT[i]; adHostF[o + i] = (float)AH[i]; adTargetF[o + i] = (float)AT[i]; }
  }
  for (int i = 0; i < 4; i++) { cPrior[i] = S.initialCalibHessian; cPriorF[i] = (float)cPrior[i]; }
}

// ---- EnergyFunctional::setDeltaF (:173-207) -----------------------------------------------------
void BAWindow::setDeltaF() {
  const int nf = n();
  adHTdeltaF.assign((size_t)nf * nf * 8, 0.f);
  for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) {
    size_t idx = (size_t)h + (size_t)t * nf;
    float dh[8], dt[8];
    for (int i = 0; i < 8; i++) { dh[i] = (float)(frames[h].state[i] - frames[h].state_zero[i]); dt[i] = (float)(frames[t].state[i] - frames[t].state_zero[i]); }
    for (int j = 0; j < 8; j++) {
      float a = 0, b = 0;
      for (int i = 0; i < 8; i++) a += dh[i] * adHostF[idx * 64 + i * 8 + j];
      for (int i = 0; i < 8; i++) b += dt[i] * adTargetF[idx * 64 + i * 8 + j];
      adHTdeltaF[idx * 8 + j] = a + b;
    }
  }
  for (int i = 0; i < 4; i++) cDeltaF[i] = (float)HCalib.value_minus_value_zero[i];
  for (auto& f : frames) for (int i = 0; i < 8; i++) { f.delta[i] = f.state[i] - f.state_zero[i]; f.delta_prior[i] = f.state[i]; }
  for (auto& p : points) p.deltaF = p.idepth - p.idepth_zero;
}

// ---- FullSystem::getNullspaces (FullSystemOptimize.cpp:1087-1147) -------------------------------
void BAWindow::getNullspaces() {
  const int nf = n(), d = dim();
  lastNullspaces_pose.clear(); lastNullspaces_scale.clear();
  for (int i = 0; i < 6; i++) {
    std::vector<double> v(d, 0.0);
    for (int f = 0; f < nf; f++) {
      for (int r = 0; r < 6; r++) v[CPARS + f * 8 + r] = frames[f].nullspaces_pose[r * 6 + i];
      for (int r = 0; r < 3; r++) v[CPARS + f * 8 + r] *= (1.0f / SCALE_XI_TRANS);
      for (int r = 3; r < 6; r++) v[CPARS + f * 8 + r] *= (1.0f / SCALE_XI_ROT);
    }
    lastNullspaces_pose.push_back(v);
  }
  std::vector<double> v(d, 0.0);
  for (int f = 0; f < nf; f++) {
    for (int r = 0; r < 6; r++) v[CPARS + f * 8 + r] = frames[f].nullspaces_scale[r];
    for (int r = 0; r < 3; r++) v[CPARS + f * 8 + r] *= (1.0f / SCALE_XI_TRANS);
    for (int r = 3; r < 6; r++) v[CPARS + f * 8 + r] *= (1.0f / SCALE_XI_ROT);
  }
  lastNullspaces_scale.push_back(v);
}

// ---- PointFrameResidual::linearize (Residuals.cpp:83-336) ---------------------------------------
double BAWindow::linearize(BARes& r) {
  r.state_NewEnergyWithOutlier = -1;
  if (r.state_state == RS_OOB) { r.state_NewState = RS_OOB; return r.state_energy; }
  const BAPoint& point = points[r.point];
  const BAFrame& host = frames[r.host]; const BAFrame& target = frames[r.target];
  const FrameFramePrecalc& pc = precalc[(size_t)r.host * n() + r.target];
  float energyLeft = 0;
  const float* dIl = target.img->dIp[0].data();
  const float* color = point.color; const float* weights = point.weights;
  const float affLL[2] = {pc.PRE_aff_mode[0], pc.PRE_aff_mode[1]};
  const float b0 = pc.PRE_b0_mode;
  const float fxl = HCalib.fxl, fyl = HCalib.fyl, cxl = HCalib.cxl, cyl = HCalib.cyl, fxli = HCalib.fxli, fyli = HCalib.fyli;
  const int w0 = G->w[0];
  float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
  {
    // projectPoint (ResidualProjections.h:64-96) at the FEJ point
    const float* R = pc.PRE_RTll_0; const float* t = pc.PRE_tTll_0;
    float KliP[3] = {(point.u + 0 - cxl) * fxli, (point.v + 0 - cyl) * fyli, 1};
    float ptp[3];
    for (int k = 0; k < 3; k++) ptp[k] = (R[k * 3] * KliP[0] + R[k * 3 + 1] * KliP[1] + R[k * 3 + 2] * KliP[2]) + t[k] * point.idepth_zero_scaled;
    float drescale = 1.0f / ptp[2];
    float new_idepth = point.idepth_zero_scaled * drescale;
    bool ok = (drescale > 0);
    float u = 0, v = 0, Ku = 0, Kv = 0;
    if (ok) {
      u = ptp[0] * drescale; v = ptp[1] * drescale;
      Ku = u * fxl + cxl; Kv = v * fyl + cyl;
      ok = Ku > 1.1f && Kv > 1.1f && Ku < G->wM3G && Kv < G->hM3G;
    }
    if (!ok) { r.state_NewState = RS_OOB; return r.state_energy; }
    r.centerProjectedTo[0] = Ku; r.centerProjectedTo[1] = Kv; r.centerProjectedTo[2] = new_idepth;
    d_d_x = drescale * (t[0] - t[2] * u) * SCALE_IDEPTH * fxl;
    d_d_y = drescale * (t[1] - t[2] * v) * SCALE_IDEPTH * fyl;
    d_C_x[2] = drescale * (R[6] * u - R[0]);
    d_C_x[3] = fxl * drescale * (R[7] * u - R[1]) * fyli;
    d_C_x[0] = KliP[0] * d_C_x[2];
    d_C_x[1] = KliP[1] * d_C_x[3];
    d_C_y[2] = fyl * drescale * (R[6] * v - R[3]) * fxli;
    d_C_y[3] = drescale * (R[7] * v - R[4]);
    d_C_y[0] = KliP[0] * d_C_y[2];
    d_C_y[1] = KliP[1] * d_C_y[3];
    d_C_x[0] = (d_C_x[0] + u) * SCALE_F;
    d_C_x[1] *= SCALE_F;
    d_C_x[2] = (d_C_x[2] + 1) * SCALE_C;
    d_C_x[3] *= SCALE_C;
    d_C_y[0] *= SCALE_F;
    d_C_y[1] = (d_C_y[1] + v) * SCALE_F;
    d_C_y[2] *= SCALE_C;
    d_C_y[3] = (d_C_y[3] + 1) * SCALE_C;
    d_xi_x[0] = new_idepth * fxl; d_xi_x[1] = 0; d_xi_x[2] = -new_idepth * u * fxl;
    d_xi_x[3] = -u * v * fxl; d_xi_x[4] = (1 + u * u) * fxl; d_xi_x[5] = -v * fxl;
    d_xi_y[0] = 0; d_xi_y[1] = new_idepth * fyl; d_xi_y[2] = -new_idepth * v * fyl;
    d_xi_y[3] = -(1 + v * v) * fyl; d_xi_y[4] = u * v * fyl; d_xi_y[5] = u * fyl;
  }
  RawResidualJacobian& J = r.J;
  for (int i = 0; i < 6; i++) { J.Jpdxi[0][i] = d_xi_x[i]; J.Jpdxi[1][i] = d_xi_y[i]; }
  for (int i = 0; i < 4; i++) { J.Jpdc[0][i] = d_C_x[i]; J.Jpdc[1][i] = d_C_y[i]; }
  J.Jpdd[0] = d_d_x; J.Jpdd[1] = d_d_y;
  float JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
  float JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
  float JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
  float wJI2_sum = 0;
  const float* KRKi = pc.PRE_KRKiTll; const float* Kt = pc.PRE_KtTll;
  for (int idx = 0; idx < patternNum; idx++) {
    // projectPoint (ResidualProjections.h:45-58) at the current state
    float up = point.u + patternP[idx][0], vp = point.v + patternP[idx][1];
    float ptp[3];
    for (int k = 0; k < 3; k++) ptp[k] = (KRKi[k * 3] * up + KRKi[k * 3 + 1] * vp + KRKi[k * 3 + 2] * 1.0f) + Kt[k] * point.idepth_scaled;
    float Ku = ptp[0] / ptp[2], Kv = ptp[1] / ptp[2];
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < G->wM3G && Kv < G->hM3G)) { r.state_NewState = RS_OOB; return r.state_energy; }
    r.projectedTo[idx][0] = Ku; r.projectedTo[idx][1] = Kv;
    float hitColor[3];
    getInterpolatedElement33(dIl, Ku, Kv, w0, hitColor);
    float residual = hitColor[0] - (float)(affLL[0] * color[idx] + affLL[1]);
    float drdA = (color[idx] - b0);
    if (!std::isfinite((float)hitColor[0])) { r.state_NewState = RS_OOB; return r.state_energy; }
    float w = sqrtf(S.outlierTHSumComponent / (S.outlierTHSumComponent + (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2])));
    w = 0.5f * (w + weights[idx]);
    float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
    energyLeft += w * w * hw * residual * residual * (2 - hw);
    {
      if (hw < 1) hw = sqrtf(hw);
      hw = hw * w;
      hitColor[1] *= hw; hitColor[2] *= hw;
      J.resF[idx] = residual * hw;
      J.JIdx[0][idx] = hitColor[1]; J.JIdx[1][idx] = hitColor[2];
      J.JabF[0][idx] = drdA * hw; J.JabF[1][idx] = hw;
      JIdxJIdx_00 += hitColor[1] * hitColor[1];
      JIdxJIdx_11 += hitColor[2] * hitColor[2];
      JIdxJIdx_10 += hitColor[1] * hitColor[2];
      JabJIdx_00 += drdA * hw * hitColor[1];
      JabJIdx_01 += drdA * hw * hitColor[2];
      JabJIdx_10 += hw * hitColor[1];
      JabJIdx_11 += hw * hitColor[2];
      JabJab_00 += drdA * drdA * hw * hw;
      JabJab_01 += drdA * hw * hw;
      JabJab_11 += hw * hw;
      wJI2_sum += hw * hw * (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2]);
      if (S.affineOptModeA < 0) J.JabF[0][idx] = 0;
      if (S.affineOptModeB < 0) J.JabF[1][idx] = 0;
    }
  }
  J.JIdx2[0] = JIdxJIdx_00; J.JIdx2[1] = JIdxJIdx_10; J.JIdx2[2] = JIdxJIdx_10; J.JIdx2[3] = JIdxJIdx_11;
  J.JabJIdx[0] = JabJIdx_00; J.JabJIdx[1] = JabJIdx_01; J.JabJIdx[2] = JabJIdx_10; J.JabJIdx[3] = JabJIdx_11;
  J.Jab2[0] = JabJab_00; J.Jab2[1] = JabJab_01; J.Jab2[2] = JabJab_01; J.Jab2[3] = JabJab_11;
  r.state_NewEnergyWithOutlier = energyLeft;
  float th = std::max<float>(host.frameEnergyTH, target.frameEnergyTH);
  if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; r.state_NewState = RS_OUTLIER; }
  else r.state_NewState = RS_IN;
  r.state_NewEnergy = energyLeft;
  return energyLeft;
}

// Residuals.cpp:367-385 + EnergyFunctionalStructs.cpp:37-51
void BAWindow::applyRes(BARes& r, bool copyJacobians) {
  if (copyJacobians) {
    if (r.state_state == RS_OOB) return;
    if (r.state_NewState == RS_IN) {
      r.isActiveAndIsGoodNEW = true;
      std::swap(r.J, r.efJ);  // takeDataF
      const RawResidualJacobian& J = r.efJ;
      float v0 = J.JIdx2[0] * J.Jpdd[0] + J.JIdx2[1] * J.Jpdd[1];
      float v1 = J.JIdx2[2] * J.Jpdd[0] + J.JIdx2[3] * J.Jpdd[1];
      for (int i = 0; i < 6; i++) r.JpJdF[i] = J.Jpdxi[0][i] * v0 + J.Jpdxi[1][i] * v1;
      r.JpJdF[6] = J.JabJIdx[0] * J.Jpdd[0] + J.JabJIdx[1] * J.Jpdd[1];
      r.JpJdF[7] = J.JabJIdx[2] * J.Jpdd[0] + J.JabJIdx[3] * J.Jpdd[1];
    } else {
      r.isActiveAndIsGoodNEW = false;
    }
  }
  r.state_state = r.state_NewState;
  r.state_energy = r.state_NewEnergy;
}

// EnergyFunctionalStructs.cpp:96-123
void BAWindow::fixLinearizationF(BARes& r) {
  const float* dp = &adHTdeltaF[((size_t)r.host + (size_t)n() * r.target) * 8];
  const RawResidualJacobian& J = r.efJ;
  const float pdelta = points[r.point].deltaF;
  float Jp_delta_x = (J.Jpdxi[0][0] * dp[0] + J.Jpdxi[0][1] * dp[1] + J.Jpdxi[0][2] * dp[2] + J.Jpdxi[0][3] * dp[3] + J.Jpdxi[0][4] * dp[4] + J.Jpdxi[0][5] * dp[5]) +
                     (J.Jpdc[0][0] * cDeltaF[0] + J.Jpdc[0][1] * cDeltaF[1] + J.Jpdc[0][2] * cDeltaF[2] + J.Jpdc[0][3] * cDeltaF[3]) + J.Jpdd[0] * pdelta;
  float Jp_delta_y = (J.Jpdxi[1][0] * dp[0] + J.Jpdxi[1][1] * dp[1] + J.Jpdxi[1][2] * dp[2] + J.Jpdxi[1][3] * dp[3] + J.Jpdxi[1][4] * dp[4] + J.Jpdxi[1][5] * dp[5]) +
                     (J.Jpdc[1][0] * cDeltaF[0] + J.Jpdc[1][1] * cDeltaF[1] + J.Jpdc[1][2] * cDeltaF[2] + J.Jpdc[1][3] * cDeltaF[3]) + J.Jpdd[1] * pdelta;
  for (int i = 0; i < patternNum; i++) {
    float rtz = J.resF[i];
    rtz = rtz - J.JIdx[0][i] * Jp_delta_x;
    rtz = rtz - J.JIdx[1][i] * Jp_delta_y;
    rtz = rtz - J.JabF[0][i] * dp[6];
    rtz = rtz - J.JabF[1][i] * dp[7];
    r.res_toZeroF[i] = rtz;
  }
  r.isLinearized = true;
}

// FullSystemOptimize.cpp:142-203 (single-threaded restatement; returns the summed energy)
double BAWindow::linearizeAll(bool fixLinearization) {
  double lastEnergyP = 0;
  for (auto& r : res) {
    if (r.isLinearized) continue;  // activeResiduals = residuals that are not linearised (:900-902)
    lastEnergyP += linearize(r);
    if (fixLinearization) applyRes(r, true);
  }
  return lastEnergyP;
}

// ---- accumulators (OptimizationBackend/MatrixAccumulators.h) -------------------------------------
namespace {
struct AccumulatorApprox {  // :564-904
  float Data[60], Data1k[60], Data1m[60];
  float TopRight_Data[32], TopRight_Data1k[32], TopRight_Data1m[32];
  float BotRight_Data[8], BotRight_Data1k[8], BotRight_Data1m[8];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  float H[169];
  void initialize() {
    memset(this, 0, sizeof(*this));
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1k[i] = Data[i] + Data1k[i];
      for (int i = 0; i < 32; i++) TopRight_Data1k[i] = TopRight_Data[i] + TopRight_Data1k[i];
      for (int i = 0; i < 8; i++) BotRight_Data1k[i] = BotRight_Data[i] + BotRight_Data1k[i];
      numIn1k += numIn1; numIn1 = 0;
      memset(Data, 0, sizeof(Data)); memset(TopRight_Data, 0, sizeof(TopRight_Data)); memset(BotRight_Data, 0, sizeof(BotRight_Data));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1m[i] = Data1k[i] + Data1m[i];
      for (int i = 0; i < 32; i++) TopRight_Data1m[i] = TopRight_Data1k[i] + TopRight_Data1m[i];
      for (int i = 0; i < 8; i++) BotRight_Data1m[i] = BotRight_Data1k[i] + BotRight_Data1m[i];
      numIn1m += numIn1k; numIn1k = 0;
      memset(Data1k, 0, sizeof(Data1k)); memset(TopRight_Data1k, 0, sizeof(TopRight_Data1k)); memset(BotRight_Data1k, 0, sizeof(BotRight_Data1k));
    }
  }
  // :714-784 update(x4,x6,y4,y6,a,b,c)
  void update(const float* x4, const float* x6, const float* y4, const float* y6, float a, float b, float c) {
    float x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    int idx = 0;
    for (int r = 0; r < 10; r++) for (int cc = r; cc < 10; cc++) {
      Data[idx] += a * x[cc] * x[r] + c * y[cc] * y[r] + b * (x[cc] * y[r] + y[cc] * x[r]);
      idx++;
    }
    num++; numIn1++;
    shiftUp(false);
  }
  // :786-836
  void updateTopRight(const float* x4, const float* x6, const float* y4, const float* y6, float TR00, float TR10, float TR01, float TR11, float TR02, float TR12) {
    float x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    for (int r = 0; r < 10; r++) {
      TopRight_Data[3 * r + 0] += x[r] * TR00 + y[r] * TR10;
      TopRight_Data[3 * r + 1] += x[r] * TR01 + y[r] * TR11;
      TopRight_Data[3 * r + 2] += x[r] * TR02 + y[r] * TR12;
    }
  }
  // :838-852
  void updateBotRight(float a00, float a01, float a02, float a11, float a12, float a22) {
    BotRight_Data[0] += a00; BotRight_Data[1] += a01; BotRight_Data[2] += a02;
    BotRight_Data[3] += a11; BotRight_Data[4] += a12; BotRight_Data[5] += a22;
  }
  // :586-613
  void finish() {
    memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 10; r++) for (int c = r; c < 10; c++) { H[r * 13 + c] = H[c * 13 + r] = Data1m[idx]; idx++; }
    idx = 0;
    for (int r = 0; r < 10; r++) for (int c = 0; c < 3; c++) { H[r * 13 + c + 10] = H[(c + 10) * 13 + r] = TopRight_Data1m[idx]; idx++; }
    H[10 * 13 + 10] = BotRight_Data1m[0];
    H[10 * 13 + 11] = H[11 * 13 + 10] = BotRight_Data1m[1];
    H[10 * 13 + 12] = H[12 * 13 + 10] = BotRight_Data1m[2];
    H[11 * 13 + 11] = BotRight_Data1m[3];
    H[11 * 13 + 12] = H[12 * 13 + 11] = BotRight_Data1m[4];
    H[12 * 13 + 12] = BotRight_Data1m[5];
    num = (size_t)(numIn1 + numIn1k + numIn1m);
  }
};

template <int I, int Jn>
struct AccumulatorXX {  // :31-85
  float A[I * Jn], A1k[I * Jn], A1m[I * Jn];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  void initialize() { memset(this, 0, sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < I * Jn; i++) { A1k[i] += A[i]; A[i] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int i = 0; i < I * Jn; i++) { A1m[i] += A1k[i]; A1k[i] = 0; } numIn1m += numIn1k; numIn1k = 0; }
  }
  void update(const float* L, const float* R, float w) {
    for (int i = 0; i < I; i++) for (int j = 0; j < Jn; j++) A[i * Jn + j] += (w * L[i]) * R[j];
    numIn1++;
    shiftUp(false);
  }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
};
template <int I>
struct AccumulatorX {  // :157-215
  float A[I], A1k[I], A1m[I];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  void initialize() { memset(this, 0, sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < I; i++) { A1k[i] += A[i]; A[i] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int i = 0; i < I; i++) { A1m[i] += A1k[i]; A1k[i] = 0; } numIn1m += numIn1k; numIn1k = 0; }
  }
  void update(const float* L, float w) { for (int i = 0; i < I; i++) A[i] += w * L[i]; numIn1++; shiftUp(false); }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
};

inline double& M(std::vector<double>& H, int d, int r, int c) { return H[(size_t)r * d + c]; }

// dst(8x8 block at r0,c0) += A(8x8) * X(8x8) * B^T(8x8)
void addABt(std::vector<double>& H, int d, int r0, int c0, const double* A, const double* X, const double* B) {
  double AX[64];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 8; k++) s += A[i * 8 + k] * X[k * 8 + j]; AX[i * 8 + j] = s; }
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 8; k++) s += AX[i * 8 + k] * B[j * 8 + k]; M(H, d, r0 + i, c0 + j) += s; }
}
}  // namespace

// AccumulatedTopHessianSSE::addPoint<mode> (AccumulatedTopHessian.cpp:36-193) over all points, then
// stitchDoubleInternal with tid=-1 (:265-337) and the symmetrisation of stitchDoubleMT (.h:134-147)
void BAWindow::accumulateTop(int mode, std::vector<double>& H, std::vector<double>& b, bool usePrior) {
  const int nf = n(), d = dim();
  const int NT = reduce_threads > 1 ? reduce_threads : 1;   // per-worker accumulator sets (acc[tid][aidx], AccumulatedTopHessian.h:60-75)
  const size_t nf2s = (size_t)nf * nf;
  std::vector<AccumulatorApprox> acc(nf2s * NT);
  for (auto& a : acc) a.initialize();
  int resInA_tmp = 0;
  for (size_t pi_ = 0; pi_ < points.size(); pi_++) {
    auto& p = points[pi_];
    const size_t tbase = nf2s * (size_t)reduce_tid((int)pi_);
    if (mode == 2 && p.stateFlag != 1) continue;  // marginalizePointsF feeds only PS_MARGINALIZE points (:680-696)
    const float* dc = cDeltaF;
    float dd = p.deltaF;
    float bd_acc = 0, Hdd_acc = 0, Hcd_acc[4] = {0, 0, 0, 0};
    for (int ri : p.residuals) {
      BARes& r = res[ri];
      if (mode == 0) { if (r.isLinearized || !r.isActive()) continue; }
      if (mode == 1) { if (!r.isLinearized || !r.isActive()) continue; }
      if (mode == 2) { if (!r.isActive()) continue; }
      const RawResidualJacobian& rJ = r.efJ;
      int htIDX = r.host + r.target * nf;
      const float* dp = &adHTdeltaF[(size_t)htIDX * 8];
      float resApprox[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // (every mode fills it; the initialiser only quiets -Wmaybe-uninitialized)
      if (mode == 0) for (int i = 0; i < 8; i++) resApprox[i] = rJ.resF[i];
      if (mode == 1) {
        float Jp_delta_x = (rJ.Jpdxi[0][0] * dp[0] + rJ.Jpdxi[0][1] * dp[1] + rJ.Jpdxi[0][2] * dp[2] + rJ.Jpdxi[0][3] * dp[3] + rJ.Jpdxi[0][4] * dp[4] + rJ.Jpdxi[0][5] * dp[5]) +
                           (rJ.Jpdc[0][0] * dc[0] + rJ.Jpdc[0][1] * dc[1] + rJ.Jpdc[0][2] * dc[2] + rJ.Jpdc[0][3] * dc[3]) + rJ.Jpdd[0] * dd;
        float Jp_delta_y = (rJ.Jpdxi[1][0] * dp[0] + rJ.Jpdxi[1][1] * dp[1] + rJ.Jpdxi[1][2] * dp[2] + rJ.Jpdxi[1][3] * dp[3] + rJ.Jpdxi[1][4] * dp[4] + rJ.Jpdxi[1][5] * dp[5]) +
                           (rJ.Jpdc[1][0] * dc[0] + rJ.Jpdc[1][1] * dc[1] + rJ.Jpdc[1][2] * dc[2] + rJ.Jpdc[1][3] * dc[3]) + rJ.Jpdd[1] * dd;
        for (int i = 0; i < 8; i++) {
          float rtz = r.res_toZeroF[i];
          rtz = rtz + rJ.JIdx[0][i] * Jp_delta_x;
          rtz = rtz + rJ.JIdx[1][i] * Jp_delta_y;
          rtz = rtz + rJ.JabF[0][i] * dp[6];
          rtz = rtz + rJ.JabF[1][i] * dp[7];
          resApprox[i] = rtz;
        }
      }
      if (mode == 2) for (int i = 0; i < 8; i++) resApprox[i] = r.res_toZeroF[i];
      float JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
      for (int i = 0; i < patternNum; i++) {
        JI_r[0] += resApprox[i] * rJ.JIdx[0][i];
        JI_r[1] += resApprox[i] * rJ.JIdx[1][i];
        Jab_r[0] += resApprox[i] * rJ.JabF[0][i];
        Jab_r[1] += resApprox[i] * rJ.JabF[1][i];
        rr += resApprox[i] * resApprox[i];
      }
      AccumulatorApprox& a = acc[tbase + htIDX];
      if (mode == 0) resInA_tmp++;
      a.update(rJ.Jpdc[0], rJ.Jpdxi[0], rJ.Jpdc[1], rJ.Jpdxi[1], rJ.JIdx2[0], rJ.JIdx2[1], rJ.JIdx2[3]);
      a.updateBotRight(rJ.Jab2[0], rJ.Jab2[1], Jab_r[0], rJ.Jab2[3], Jab_r[1], rr);
      a.updateTopRight(rJ.Jpdc[0], rJ.Jpdxi[0], rJ.Jpdc[1], rJ.Jpdxi[1], rJ.JabJIdx[0], rJ.JabJIdx[1], rJ.JabJIdx[2], rJ.JabJIdx[3], JI_r[0], JI_r[1]);
      float Ji2_Jpdd[2] = {rJ.JIdx2[0] * rJ.Jpdd[0] + rJ.JIdx2[1] * rJ.Jpdd[1], rJ.JIdx2[2] * rJ.Jpdd[0] + rJ.JIdx2[3] * rJ.Jpdd[1]};
      bd_acc += JI_r[0] * rJ.Jpdd[0] + JI_r[1] * rJ.Jpdd[1];
      Hdd_acc += Ji2_Jpdd[0] * rJ.Jpdd[0] + Ji2_Jpdd[1] * rJ.Jpdd[1];
      for (int k = 0; k < 4; k++) Hcd_acc[k] += rJ.Jpdc[0][k] * Ji2_Jpdd[0] + rJ.Jpdc[1][k] * Ji2_Jpdd[1];
    }
    if (mode == 0) { p.Hdd_accAF = Hdd_acc; p.bd_accAF = bd_acc; for (int k = 0; k < 4; k++) p.Hcd_accAF[k] = Hcd_acc[k]; }
    if (mode == 1 || mode == 2) { p.Hdd_accLF = Hdd_acc; p.bd_accLF = bd_acc; for (int k = 0; k < 4; k++) p.Hcd_accLF[k] = Hcd_acc[k]; }
    if (mode == 2) { for (int k = 0; k < 4; k++) p.Hcd_accAF[k] = 0; p.Hdd_accAF = 0; p.bd_accAF = 0; }
  }
  if (mode == 0) resInA = resInA_tmp;
  // stitch
  H.assign((size_t)d * d, 0.0); b.assign(d, 0.0);
  lastTopBlocks.assign((size_t)nf * nf * 169, 0.f);
  for (int k = 0; k < nf * nf; k++) {
    int h = k % nf, t = k / nf;
    int hIdx = CPARS + h * 8, tIdx = CPARS + t * 8;
    double accH[169];
    for (int i = 0; i < 169; i++) accH[i] = 0;
    size_t num_all = 0;
    for (int tid2 = 0; tid2 < NT; tid2++) {   // :299-308
      AccumulatorApprox& a = acc[nf2s * tid2 + k];
      a.finish();
      if (a.num == 0) continue;
      num_all += a.num;
      for (int i = 0; i < 169; i++) accH[i] += (double)a.H[i];
    }
    for (int i = 0; i < 169; i++) lastTopBlocks[(size_t)k * 169 + i] = (float)accH[i];   // (exact for one worker)
    if (num_all == 0) continue;
    double A88[64], A8C[32], ACC[16], b8[8], bC[4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) A88[i * 8 + j] = accH[(CPARS + i) * 13 + CPARS + j];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) A8C[i * 4 + j] = accH[(CPARS + i) * 13 + j];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) ACC[i * 4 + j] = accH[i * 13 + j];
    for (int i = 0; i < 8; i++) b8[i] = accH[(CPARS + i) * 13 + 8 + CPARS];
    for (int i = 0; i < 4; i++) bC[i] = accH[i * 13 + 8 + CPARS];
    const double* AH = &adHost[(size_t)k * 64]; const double* AT = &adTarget[(size_t)k * 64];
    addABt(H, d, hIdx, hIdx, AH, A88, AH);
    addABt(H, d, tIdx, tIdx, AT, A88, AT);
    addABt(H, d, hIdx, tIdx, AH, A88, AT);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) {
      double sh = 0, st = 0;
      for (int q = 0; q < 8; q++) { sh += AH[i * 8 + q] * A8C[q * 4 + j]; st += AT[i * 8 + q] * A8C[q * 4 + j]; }
      M(H, d, hIdx + i, j) += sh; M(H, d, tIdx + i, j) += st;
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M(H, d, i, j) += ACC[i * 4 + j];
    for (int i = 0; i < 8; i++) {
      double sh = 0, st = 0;
      for (int q = 0; q < 8; q++) { sh += AH[i * 8 + q] * b8[q]; st += AT[i * 8 + q] * b8[q]; }
      b[hIdx + i] += sh; b[tIdx + i] += st;
    }
    for (int i = 0; i < 4; i++) b[i] += bC[i];
  }
  if (usePrior) {  // :324-336
    for (int i = 0; i < 4; i++) { M(H, d, i, i) += cPrior[i]; b[i] += cPrior[i] * (double)cDeltaF[i]; }
    for (int h = 0; h < nf; h++) for (int i = 0; i < 8; i++) {
      M(H, d, CPARS + h * 8 + i, CPARS + h * 8 + i) += frames[h].prior[i];
      b[CPARS + h * 8 + i] += frames[h].prior[i] * frames[h].delta_prior[i];
    }
  }
  // make diagonal by copying over parts (.h:134-147)
  for (int h = 0; h < nf; h++) {
    int hIdx = CPARS + h * 8;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) M(H, d, i, hIdx + j) = M(H, d, hIdx + j, i);
    for (int t = h + 1; t < nf; t++) {
      int tIdx = CPARS + t * 8;
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) M(H, d, hIdx + i, tIdx + j) += M(H, d, tIdx + j, hIdx + i);
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) M(H, d, tIdx + i, hIdx + j) = M(H, d, hIdx + j, tIdx + i);
    }
  }
}

// AccumulatedSCHessianSSE::addPoint (AccumulatedSCHessian.cpp:34-103) + stitchDoubleInternal tid=-1 (:106-195)
void BAWindow::accumulateSC(bool shiftPriorToZero, std::vector<double>& H, std::vector<double>& b) {
  const int nf = n(), d = dim(), nf2 = nf * nf;
  const int NT = reduce_threads > 1 ? reduce_threads : 1;   // per-worker accumulator sets (AccumulatedSCHessian.h:52-72)
  std::vector<AccumulatorXX<8, 4>> accE_all((size_t)nf2 * NT);
  std::vector<AccumulatorX<8>> accEB_all((size_t)nf2 * NT);
  std::vector<AccumulatorXX<8, 8>> accD_all((size_t)nf2 * nf * NT);
  std::vector<AccumulatorXX<4, 4>> accHcc_all(NT); std::vector<AccumulatorX<4>> accbc_all(NT);
  for (auto& a : accE_all) a.initialize();
  for (auto& a : accEB_all) a.initialize();
  for (auto& a : accD_all) a.initialize();
  for (auto& a : accHcc_all) a.initialize();
  for (auto& a : accbc_all) a.initialize();
  for (size_t pi_ = 0; pi_ < points.size(); pi_++) {
    auto& p = points[pi_];
    const int tid = reduce_tid((int)pi_);
    AccumulatorXX<8, 4>* accE = &accE_all[(size_t)nf2 * tid];
    AccumulatorX<8>* accEB = &accEB_all[(size_t)nf2 * tid];
    AccumulatorXX<8, 8>* accD = &accD_all[(size_t)nf2 * nf * tid];
    AccumulatorXX<4, 4>& accHcc = accHcc_all[tid]; AccumulatorX<4>& accbc = accbc_all[tid];
    if (!shiftPriorToZero && p.stateFlag != 1) continue;  // marginalizePointsF path: only PS_MARGINALIZE points
    int ngoodres = 0;
    for (int ri : p.residuals) if (res[ri].isActive()) ngoodres++;
    if (ngoodres == 0) { p.HdiF = 0; p.bdSumF = 0; p.idepth_hessian = 0; continue; }
    float Hh = p.Hdd_accAF + p.Hdd_accLF + p.priorF;
    if (Hh < 1e-10) Hh = 1e-10;
    p.idepth_hessian = Hh;
    p.HdiF = 1.0 / Hh;
    p.bdSumF = p.bd_accAF + p.bd_accLF;
    if (shiftPriorToZero) p.bdSumF += p.priorF * p.deltaF;
    float Hcd[4];
    for (int k = 0; k < 4; k++) Hcd[k] = p.Hcd_accAF[k] + p.Hcd_accLF[k];
    accHcc.update(Hcd, Hcd, p.HdiF);
    accbc.update(Hcd, p.bdSumF * p.HdiF);
    for (int r1i : p.residuals) {
      const BARes& r1 = res[r1i];
      if (!r1.isActive()) continue;
      int r1ht = r1.host + r1.target * nf;
      for (int r2i : p.residuals) {
        const BARes& r2 = res[r2i];
        if (!r2.isActive()) continue;
        accD[(size_t)r1ht + (size_t)r2.target * nf2].update(r1.JpJdF, r2.JpJdF, p.HdiF);
      }
      accE[r1ht].update(r1.JpJdF, Hcd, p.HdiF);
      accEB[r1ht].update(r1.JpJdF, p.HdiF * p.bdSumF);
    }
  }
  H.assign((size_t)d * d, 0.0); b.assign(d, 0.0);
  for (int k = 0; k < nf2; k++) {
    int i = k % nf, j = k / nf;
    int iIdx = CPARS + i * 8, jIdx = CPARS + j * 8, ijIdx = i + nf * j;
    double Hpc[32], bp[8];
    for (int q = 0; q < 32; q++) Hpc[q] = 0;
    for (int q = 0; q < 8; q++) bp[q] = 0;
    for (int tid2 = 0; tid2 < NT; tid2++) {   // sum of all workers (:140-146)
      auto& aE = accE_all[(size_t)nf2 * tid2 + ijIdx]; auto& aEB = accEB_all[(size_t)nf2 * tid2 + ijIdx];
      aE.finish(); aEB.finish();
      for (int q = 0; q < 32; q++) Hpc[q] += (double)aE.A1m[q];
      for (int q = 0; q < 8; q++) bp[q] += (double)aEB.A1m[q];
    }
    const double* AH = &adHost[(size_t)ijIdx * 64]; const double* AT = &adTarget[(size_t)ijIdx * 64];
    for (int r = 0; r < 8; r++) {
      for (int c = 0; c < 4; c++) {
        double sh = 0, st = 0;
        for (int q = 0; q < 8; q++) { sh += AH[r * 8 + q] * Hpc[q * 4 + c]; st += AT[r * 8 + q] * Hpc[q * 4 + c]; }
        M(H, d, iIdx + r, c) += sh; M(H, d, jIdx + r, c) += st;
      }
      double sh = 0, st = 0;
      for (int q = 0; q < 8; q++) { sh += AH[r * 8 + q] * bp[q]; st += AT[r * 8 + q] * bp[q]; }
      b[iIdx + r] += sh; b[jIdx + r] += st;
    }
    for (int kk = 0; kk < nf; kk++) {
      int kIdx = CPARS + kk * 8, ijkIdx = ijIdx + kk * nf2, ikIdx = i + nf * kk;
      double Dm[64];
      for (int q = 0; q < 64; q++) Dm[q] = 0;
      size_t numD = 0;
      for (int tid2 = 0; tid2 < NT; tid2++) {   // (:163-168)
        auto& aD = accD_all[(size_t)nf2 * nf * tid2 + ijkIdx];
        aD.finish();
        if (aD.num == 0) continue;
        numD += aD.num;
        for (int q = 0; q < 64; q++) Dm[q] += (double)aD.A1m[q];
      }
      if (numD == 0) continue;
      const double* AHik = &adHost[(size_t)ikIdx * 64]; const double* ATik = &adTarget[(size_t)ikIdx * 64];
      addABt(H, d, iIdx, iIdx, AH, Dm, AHik);
      addABt(H, d, jIdx, kIdx, AT, Dm, ATik);
      addABt(H, d, jIdx, iIdx, AT, Dm, AHik);
      addABt(H, d, iIdx, kIdx, AH, Dm, ATik);
    }
  }
  for (int tid2 = 0; tid2 < NT; tid2++) {   // (:183-190)
    accHcc_all[tid2].finish(); accbc_all[tid2].finish();
    for (int i = 0; i < 4; i++) { for (int j = 0; j < 4; j++) M(H, d, i, j) += accHcc_all[tid2].A1m[i * 4 + j]; b[i] += accbc_all[tid2].A1m[i]; }
  }
  for (int h = 0; h < nf; h++) {  // .h:130-134
    int hIdx = CPARS + h * 8;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) M(H, d, i, hIdx + j) = M(H, d, hIdx + j, i);
  }
}

// symmetric Jacobi eigen-decomposition (n <= 16): A = V diag(w) V^T
static void jacobiEig(int n, std::vector<double> A, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (std::fabs(A[p * n + q]) < 1e-300) continue;
      double theta = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
      double c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

// EnergyFunctional::orthogonalize (:775-835): P = N (N^T N)^+ N^T with singular values below
// solverModeDelta * max cut; b -= P b ; H -= P H P. The SVD of N is obtained from the eigen-decomposition of N^T N.
void BAWindow::orthogonalize(std::vector<double>* b, std::vector<double>* H) {
  const int d = dim();
  std::vector<std::vector<double>> ns;
  for (auto& v : lastNullspaces_pose) ns.push_back(v);
  for (auto& v : lastNullspaces_scale) ns.push_back(v);
  const int m = (int)ns.size();
  std::vector<double> N((size_t)d * m);
  for (int i = 0; i < m; i++) {
    double nrm = 0; for (int r = 0; r < d; r++) nrm += ns[i][r] * ns[i][r];
    nrm = std::sqrt(nrm);
    for (int r = 0; r < d; r++) N[(size_t)r * m + i] = ns[i][r] / nrm;
  }
  std::vector<double> G((size_t)m * m, 0.0), w, V;
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) { double s = 0; for (int r = 0; r < d; r++) s += N[(size_t)r * m + i] * N[(size_t)r * m + j]; G[i * m + j] = s; }
  jacobiEig(m, G, w, V);
  double maxSv = 0;
  for (int i = 0; i < m; i++) { double sv = std::sqrt(std::max(w[i], 0.0)); if (sv > maxSv) maxSv = sv; }
  // U_i = N v_i / sigma_i for kept singular values; P = sum U_i U_i^T
  std::vector<double> P((size_t)d * d, 0.0);
  for (int i = 0; i < m; i++) {
    double sv = std::sqrt(std::max(w[i], 0.0));
    if (!(sv > S.solverModeDelta * maxSv)) continue;
    std::vector<double> U(d);
    for (int r = 0; r < d; r++) { double s = 0; for (int j = 0; j < m; j++) s += N[(size_t)r * m + j] * V[j * m + i]; U[r] = s / sv; }
    for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) P[(size_t)r * d + c] += U[r] * U[c];
  }
  for (int r = 0; r < d; r++) for (int c = r + 1; c < d; c++) { double s = 0.5 * (P[(size_t)r * d + c] + P[(size_t)c * d + r]); P[(size_t)r * d + c] = P[(size_t)c * d + r] = s; }
  if (b) {
    std::vector<double> Pb(d, 0.0);
    for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) Pb[r] += P[(size_t)r * d + c] * (*b)[c];
    for (int r = 0; r < d; r++) (*b)[r] -= Pb[r];
  }
  if (H) {
    std::vector<double> PH((size_t)d * d, 0.0), PHP((size_t)d * d, 0.0);
    for (int r = 0; r < d; r++) for (int k = 0; k < d; k++) { double p = P[(size_t)r * d + k]; if (p == 0) continue; for (int c = 0; c < d; c++) PH[(size_t)r * d + c] += p * (*H)[(size_t)k * d + c]; }
    for (int r = 0; r < d; r++) for (int k = 0; k < d; k++) { double p = PH[(size_t)r * d + k]; if (p == 0) continue; for (int c = 0; c < d; c++) PHP[(size_t)r * d + c] += p * P[(size_t)k * d + c]; }
    for (size_t i = 0; i < (size_t)d * d; i++) (*H)[i] -= PHP[i];
  }
}

// EnergyFunctional::solveSystemF (:838-995), setting_solverMode = FIX_LAMBDA | ORTHOGONALIZE_X_LATER (settings.cpp:51)
void BAWindow::solveSystemF(int iteration, double lambda, std::vector<double>& x, std::vector<double>* Hfinal, std::vector<double>* bfinal) {
  lambda = 1e-5;  // SOLVER_FIX_LAMBDA (:844-846)
  const int d = dim();
  std::vector<double> HL, bL, HA, bA, Hsc, bsc;
  accumulateTop(0, HA, bA, false);   // accumulateAF_MT (:857)
  accumulateTop(1, HL, bL, true);    // accumulateLF_MT (:863)
  accumulateSC(true, Hsc, bsc);      // accumulateSCF_MT (:866)
  if (HM.empty()) { HM.assign((size_t)d * d, 0.0); bM.assign(d, 0.0); }
  std::vector<double> delta(d);
  for (int i = 0; i < 4; i++) delta[i] = (double)cDeltaF[i];
  for (int h = 0; h < n(); h++) for (int i = 0; i < 8; i++) delta[CPARS + 8 * h + i] = frames[h].delta[i];
  std::vector<double> bM_top(d);
  for (int r = 0; r < d; r++) { double s = 0; for (int c = 0; c < d; c++) s += HM[(size_t)r * d + c] * delta[c]; bM_top[r] = bM[r] + s; }
  std::vector<double> HF((size_t)d * d), bF(d);
  for (size_t i = 0; i < (size_t)d * d; i++) HF[i] = HL[i] + HM[i] + HA[i];
  for (int i = 0; i < d; i++) bF[i] = bL[i] + bM_top[i] + bA[i] - bsc[i];
  for (int i = 0; i < d; i++) HF[(size_t)i * d + i] *= (1 + lambda);
  const double f = (1.0f / (1 + lambda));
  for (size_t i = 0; i < (size_t)d * d; i++) HF[i] -= Hsc[i] * f;
  if (Hfinal) *Hfinal = HF;
  if (bfinal) *bfinal = bF;
  std::vector<double> SVecI(d), Hs((size_t)d * d), bs(d);
  for (int i = 0; i < d; i++) SVecI[i] = 1.0 / std::sqrt(HF[(size_t)i * d + i] + 10);
  for (int r = 0; r < d; r++) for (int c = 0; c < d; c++) Hs[(size_t)r * d + c] = SVecI[r] * HF[(size_t)r * d + c] * SVecI[c];
  for (int i = 0; i < d; i++) bs[i] = SVecI[i] * bF[i];
  x.assign(d, 0.0);
  ldlt_solve(d, Hs.data(), bs.data(), x.data());
  for (int i = 0; i < d; i++) x[i] *= SVecI[i];
  if (iteration >= 2) orthogonalize(&x, nullptr);  // SOLVER_ORTHOGONALIZE_X_LATER (:980-984)
}

// EnergyFunctional::resubstituteF_MT / resubstituteFPt (:272-341)
void BAWindow::resubstituteF(const std::vector<double>& x, double* frame_steps, double calib_step[4]) {
  const int nf = n();
  std::vector<float> xF(x.size());
  for (size_t i = 0; i < x.size(); i++) xF[i] = (float)x[i];
  for (int i = 0; i < 4; i++) calib_step[i] = -x[i];
  std::vector<float> xAd((size_t)nf * nf * 8);
  float cstep[4] = {xF[0], xF[1], xF[2], xF[3]};
  for (int h = 0; h < nf; h++) {
    for (int i = 0; i < 8; i++) frame_steps[h * 10 + i] = -x[CPARS + 8 * h + i];
    frame_steps[h * 10 + 8] = frame_steps[h * 10 + 9] = 0;
    for (int t = 0; t < nf; t++) {
      size_t ad = ((size_t)h + (size_t)nf * t) * 64;
      for (int j = 0; j < 8; j++) {
        float a = 0, b2 = 0;
        for (int i = 0; i < 8; i++) a += xF[CPARS + 8 * h + i] * adHostF[ad + i * 8 + j];
        for (int i = 0; i < 8; i++) b2 += xF[CPARS + 8 * t + i] * adTargetF[ad + i * 8 + j];
        xAd[((size_t)nf * h + t) * 8 + j] = a + b2;
      }
    }
  }
  for (auto& p : points) {
    int ngoodres = 0;
    for (int ri : p.residuals) if (res[ri].isActive()) ngoodres++;
    if (ngoodres == 0) { p.step = 0; continue; }
    float b = p.bdSumF;
    float dot = 0;
    for (int k = 0; k < 4; k++) dot += cstep[k] * (p.Hcd_accAF[k] + p.Hcd_accLF[k]);
    b -= dot;
    for (int ri : p.residuals) {
      const BARes& r = res[ri];
      if (!r.isActive()) continue;
      const float* xa = &xAd[((size_t)r.host * nf + r.target) * 8];
      float s = 0;
      for (int i = 0; i < 8; i++) s += xa[i] * r.JpJdF[i];
      b -= s;
    }
    p.step = -b * p.HdiF;
  }
}

// EnergyFunctional::marginalizePointsF (:663-736), solverMode without the ORTHOGONALIZE_* point-marg bits
void BAWindow::marginalizePointsF() {
  const int d = dim();
  for (auto& p : points) if (p.stateFlag == 1) p.priorF *= S.idepthFixPriorMargFac;
  std::vector<double> Mm, Mb, Msc, Mbsc;
  accumulateTop(2, Mm, Mb, false);
  accumulateSC(false, Msc, Mbsc);
  if (HM.empty()) { HM.assign((size_t)d * d, 0.0); bM.assign(d, 0.0); }
  for (size_t i = 0; i < (size_t)d * d; i++) HM[i] += S.margWeightFac * (Mm[i] - Msc[i]);
  for (int i = 0; i < d; i++) bM[i] += S.margWeightFac * (Mb[i] - Mbsc[i]);
  // removePoint: drop the marginalised points and their residuals
  std::vector<BAPoint> keepP; std::vector<BARes> keepR;
  for (auto& p : points) {
    if (p.stateFlag == 1) continue;
    BAPoint q = p; q.residuals.clear();
    for (int ri : p.residuals) { BARes r = res[ri]; r.point = (int)keepP.size(); q.residuals.push_back((int)keepR.size()); keepR.push_back(r); }
    keepP.push_back(q);
  }
  points.swap(keepP); res.swap(keepR);
}

// EnergyFunctional::marginalizeFrame (:554-660)
void BAWindow::marginalizeFrame(int idx) {
  const int nf = n();
  const int ndim = nf * 8 + CPARS - 8, odim = nf * 8 + CPARS;
  if (HM.empty()) { HM.assign((size_t)odim * odim, 0.0); bM.assign(odim, 0.0); }
  // permutation that moves frame idx to the end
  std::vector<int> perm;
  for (int i = 0; i < odim; i++) if (i < CPARS + idx * 8 || i >= CPARS + idx * 8 + 8) perm.push_back(i);
  for (int i = 0; i < 8; i++) perm.push_back(CPARS + idx * 8 + i);
  std::vector<double> Hp((size_t)odim * odim), bp(odim);
  for (int r = 0; r < odim; r++) { bp[r] = bM[perm[r]]; for (int c = 0; c < odim; c++) Hp[(size_t)r * odim + c] = HM[(size_t)perm[r] * odim + perm[c]]; }
  for (int i = 0; i < 8; i++) {
    Hp[(size_t)(ndim + i) * odim + ndim + i] += frames[idx].prior[i];
    bp[ndim + i] += frames[idx].prior[i] * frames[idx].delta_prior[i];
  }
  std::vector<double> SVec(odim), SVecI(odim);
  for (int i = 0; i < odim; i++) { SVec[i] = std::sqrt(std::fabs(Hp[(size_t)i * odim + i]) + 10); SVecI[i] = 1.0 / SVec[i]; }
  std::vector<double> Hs((size_t)odim * odim), bs(odim);
  for (int r = 0; r < odim; r++) { bs[r] = SVecI[r] * bp[r]; for (int c = 0; c < odim; c++) Hs[(size_t)r * odim + c] = SVecI[r] * Hp[(size_t)r * odim + c] * SVecI[c]; }
  double hpi[64], hpiInv[64];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) hpi[i * 8 + j] = Hs[(size_t)(ndim + i) * odim + ndim + j];
  for (int i = 0; i < 64; i++) hpi[i] = 0.5f * (hpi[i] + hpi[i]);
  mat_inverse(8, hpi, hpiInv);
  for (int i = 0; i < 64; i++) hpiInv[i] = 0.5f * (hpiInv[i] + hpiInv[i]);
  std::vector<double> bli((size_t)ndim * 8);
  for (int r = 0; r < ndim; r++) for (int c = 0; c < 8; c++) { double s = 0; for (int k = 0; k < 8; k++) s += Hs[(size_t)(ndim + k) * odim + r] * hpiInv[k * 8 + c]; bli[(size_t)r * 8 + c] = s; }
  for (int r = 0; r < ndim; r++) {
    for (int c = 0; c < ndim; c++) { double s = 0; for (int k = 0; k < 8; k++) s += bli[(size_t)r * 8 + k] * Hs[(size_t)(ndim + k) * odim + c]; Hs[(size_t)r * odim + c] -= s; }
    double s = 0; for (int k = 0; k < 8; k++) s += bli[(size_t)r * 8 + k] * bs[ndim + k];
    bs[r] -= s;
  }
  std::vector<double> Hn((size_t)ndim * ndim), bn(ndim);
  for (int r = 0; r < ndim; r++) { bn[r] = SVec[r] * bs[r]; for (int c = 0; c < ndim; c++) Hn[(size_t)r * ndim + c] = SVec[r] * Hs[(size_t)r * odim + c] * SVec[c]; }
  HM.assign((size_t)ndim * ndim, 0.0);
  for (int r = 0; r < ndim; r++) for (int c = 0; c < ndim; c++) HM[(size_t)r * ndim + c] = 0.5 * (Hn[(size_t)r * ndim + c] + Hn[(size_t)c * ndim + r]);
  bM = bn;
  // remove the frame; re-index points / residuals
  frames.erase(frames.begin() + idx);
  for (auto& p : points) if (p.host > idx) p.host--;
  for (auto& r : res) { if (r.host > idx) r.host--; if (r.target > idx) r.target--; }
}

// EnergyFunctional::calcMEnergyF (:344-351)
double BAWindow::calcMEnergyF() {
  const int d = dim();
  if (HM.empty()) return 0;
  std::vector<double> delta(d);
  for (int i = 0; i < 4; i++) delta[i] = (double)cDeltaF[i];
  for (int h = 0; h < n(); h++) for (int i = 0; i < 8; i++) delta[CPARS + 8 * h + i] = frames[h].delta[i];
  double e = 0;
  for (int r = 0; r < d; r++) { double s = 0; for (int c = 0; c < d; c++) s += HM[(size_t)r * d + c] * delta[c]; e += delta[r] * (2 * bM[r] + s); }
  return e;
}

// EnergyFunctional::calcLEnergyF_MT + calcLEnergyPt (:354-442), single accumulator
double BAWindow::calcLEnergyF() {
  double E = 0;
  for (auto& f : frames) for (int i = 0; i < 8; i++) E += f.delta_prior[i] * f.prior[i] * f.delta_prior[i];
  for (int i = 0; i < 4; i++) E += cDeltaF[i] * cPriorF[i] * cDeltaF[i];
  float acc = 0;
  for (auto& p : points) {
    float dd = p.deltaF;
    for (int ri : p.residuals) {
      const BARes& r = res[ri];
      if (!r.isLinearized || !r.isActive()) continue;
      const float* dp = &adHTdeltaF[((size_t)r.host + (size_t)n() * r.target) * 8];
      const RawResidualJacobian& rJ = r.efJ;
      float Jpx = (rJ.Jpdxi[0][0] * dp[0] + rJ.Jpdxi[0][1] * dp[1] + rJ.Jpdxi[0][2] * dp[2] + rJ.Jpdxi[0][3] * dp[3] + rJ.Jpdxi[0][4] * dp[4] + rJ.Jpdxi[0][5] * dp[5]) +
                  (rJ.Jpdc[0][0] * cDeltaF[0] + rJ.Jpdc[0][1] * cDeltaF[1] + rJ.Jpdc[0][2] * cDeltaF[2] + rJ.Jpdc[0][3] * cDeltaF[3]) + rJ.Jpdd[0] * dd;
      float Jpy = (rJ.Jpdxi[1][0] * dp[0] + rJ.Jpdxi[1][1] * dp[1] + rJ.Jpdxi[1][2] * dp[2] + rJ.Jpdxi[1][3] * dp[3] + rJ.Jpdxi[1][4] * dp[4] + rJ.Jpdxi[1][5] * dp[5]) +
                  (rJ.Jpdc[1][0] * cDeltaF[0] + rJ.Jpdc[1][1] * cDeltaF[1] + rJ.Jpdc[1][2] * cDeltaF[2] + rJ.Jpdc[1][3] * cDeltaF[3]) + rJ.Jpdd[1] * dd;
      for (int i = 0; i < 8; i++) {
        float Jdelta = rJ.JIdx[0][i] * Jpx;
        Jdelta = Jdelta + rJ.JIdx[1][i] * Jpy;
        Jdelta = Jdelta + rJ.JabF[0][i] * dp[6];
        Jdelta = Jdelta + rJ.JabF[1][i] * dp[7];
        float r0 = r.res_toZeroF[i];
        r0 = r0 + r0; r0 = r0 + Jdelta;
        acc += Jdelta * r0;
      }
    }
    acc += p.deltaF * p.deltaF * p.priorF;
  }
  return E + acc;
}


// ---- LM driver of the SSE path ---------------------------------------------------------------------------------
void BAWindow::initCalibValue() {
  if (calib_init) return;
  // CalibHessian(): value_scaled = (fx,fy,cx,cy); value = SCALE_*_INVERSE * value_scaled; value_zero = value
  calib_value[0] = (1.0f / SCALE_F) * (double)HCalib.fxl; calib_value[1] = (1.0f / SCALE_F) * (double)HCalib.fyl;
  calib_value[2] = (1.0f / SCALE_C) * (double)HCalib.cxl; calib_value[3] = (1.0f / SCALE_C) * (double)HCalib.cyl;
  for (int i = 0; i < 4; i++) calib_value_zero[i] = calib_value[i] - HCalib.value_minus_value_zero[i];
  calib_init = true;
}
void BAWindow::setCalibValue(const double v[4]) {
  for (int i = 0; i < 4; i++) calib_value[i] = v[i];
  const double vs[4] = {SCALE_F * v[0], SCALE_F * v[1], SCALE_C * v[2], SCALE_C * v[3]};
  HCalib.fxl = (float)vs[0]; HCalib.fyl = (float)vs[1]; HCalib.cxl = (float)vs[2]; HCalib.cyl = (float)vs[3];
  HCalib.fxli = 1.0f / HCalib.fxl; HCalib.fyli = 1.0f / HCalib.fyl;
  HCalib.cxli = -HCalib.cxl / HCalib.fxl; HCalib.cyli = -HCalib.cyl / HCalib.fyl;
  for (int i = 0; i < 4; i++) HCalib.value_minus_value_zero[i] = calib_value[i] - calib_value_zero[i];
}
void BAWindow::backupState() {  // FullSystemOptimize.cpp:309-350, non-momentum branch
  for (int i = 0; i < 4; i++) calib_backup[i] = calib_value[i];
  for (auto& f : frames) for (int i = 0; i < 10; i++) f.state_backup[i] = f.state[i];
  for (auto& p : points) p.idepth_backup = p.idepth;
}
bool BAWindow::doStepFromBackup(float stepfacC, float stepfacT, float stepfacR, float stepfacA, float stepfacD) {  // :207-305
  float sumA = 0, sumB = 0, sumT = 0, sumR = 0, sumID = 0, numID = 0, sumNID = 0;
  double v[4];
  for (int i = 0; i < 4; i++) v[i] = calib_backup[i] + stepfacC * calib_step[i];
  setCalibValue(v);
  const double pf[10] = {stepfacT, stepfacT, stepfacT, stepfacR, stepfacR, stepfacR, stepfacA, stepfacA, stepfacA, stepfacA};
  for (size_t h = 0; h < frames.size(); h++) {
    BAFrame& f = frames[h];
    double st[10];
    for (int i = 0; i < 10; i++) st[i] = f.state_backup[i] + pf[i] * f.step[i];
    f.setState(st);
    sumA += f.step[6] * f.step[6];
    sumB += f.step[7] * f.step[7];
    sumT += f.step[0] * f.step[0] + f.step[1] * f.step[1] + f.step[2] * f.step[2];
    sumR += f.step[3] * f.step[3] + f.step[4] * f.step[4] + f.step[5] * f.step[5];
  }
  for (auto& p : points) {
    const float nv = p.idepth_backup + stepfacD * p.step;
    p.idepth = nv; p.idepth_scaled = SCALE_IDEPTH * nv;
    sumID += p.step * p.step;
    sumNID += fabsf(p.idepth_backup);
    numID++;
    p.idepth_zero = nv; p.idepth_zero_scaled = SCALE_IDEPTH * nv;
  }
  sumA /= frames.size(); sumB /= frames.size(); sumR /= frames.size(); sumT /= frames.size();
  sumID /= numID; sumNID /= numID;
  setPrecalcValues(); setDeltaF();
  const float th = S.thOptIterations;
  return sqrtf(sumA) < 0.0005 * th && sqrtf(sumB) < 0.00005 * th && sqrtf(sumR) < 0.00005 * th && sqrtf(sumT) * sumNID < 0.00005 * th;
}
float BAWindow::newFrameEnergyTH() {  // :98-139
  std::vector<float> all;
  const int newest = n() - 1;
  for (auto& r : res) if (!r.isLinearized && r.state_NewEnergyWithOutlier >= 0 && r.target == newest) all.push_back((float)r.state_NewEnergyWithOutlier);
  if (all.empty()) return 12 * 12 * patternNum;
  const int nthIdx = (int)(S.frameEnergyTHN * all.size());
  std::nth_element(all.begin(), all.begin() + nthIdx, all.end());
  const float nthElement = sqrtf(all[nthIdx]);
  float th = nthElement * S.frameEnergyTHFacMedian;
  th = 26.0f * S.frameEnergyTHConstWeight + th * (1 - S.frameEnergyTHConstWeight);
  th = th * th;
  th *= S.overallEnergyTHWeight * S.overallEnergyTHWeight;
  return th;
}
double BAWindow::optimize(int mnumOptIts, int* iterations_done) {  // :870-1042 with setting_forceAceptStep = true (settings.cpp:53)
  initCalibValue();
  const int nf = n();
  if (nf < 2) return 0;
  if (nf < 3) mnumOptIts = 20;
  if (nf < 4) mnumOptIts = 15;
  for (auto& r : res) if (!r.isLinearized) { r.state_NewEnergy = r.state_energy = 0; r.state_NewState = RS_OUTLIER; r.state_state = RS_IN; }  // resetOOB
  // linearizeAll ends with setNewFrameEnergyTH() (FullSystemOptimize.cpp:163): the newest frame's threshold follows every pass
  double lastEnergy = linearizeAll(false);
  frames.back().frameEnergyTH = newFrameEnergyTH();
  for (auto& r : res) if (!r.isLinearized) applyRes(r, true);
  const double lambda = 1e-1;
  int it = 0;
  for (; it < mnumOptIts; it++) {
    backupState();
    std::vector<double> x;
    solveSystemF(it, lambda, x, nullptr, nullptr);
    std::vector<double> fs((size_t)nf * 10);
    resubstituteF(x, fs.data(), calib_step);
    for (int h = 0; h < nf; h++) for (int i = 0; i < 10; i++) frames[h].step[i] = fs[(size_t)h * 10 + i];
    const bool canbreak = doStepFromBackup(1, 1, 1, 1, 1);
    lastEnergy = linearizeAll(false);
    frames.back().frameEnergyTH = newFrameEnergyTH();
    for (auto& r : res) if (!r.isLinearized) applyRes(r, true);
    if (canbreak && it >= S.minOptIterations) { it++; break; }
  }
  if (iterations_done) *iterations_done = it;
  // new evaluation point of the newest frame (:996-1005)
  BAFrame& nw = frames.back();
  double nz[10] = {0, 0, 0, 0, 0, 0, nw.state[6], nw.state[7], 0, 0};
  nw.worldToCam_evalPT = nw.PRE_worldToCam;
  nw.setState(nz);
  nw.setStateZero(nz);
  setAdjointsF();
  setPrecalcValues(); setDeltaF(); getNullspaces();
  lastEnergy = linearizeAll(true);
  frames.back().frameEnergyTH = newFrameEnergyTH();
  std::vector<double> Ht, bt;
  accumulateTop(0, Ht, bt, false);  // resInA as the last accumulateAF would report it
  return sqrtf((float)(lastEnergy / (patternNum * std::max(resInA, 1))));
}

}  // namespace orc
