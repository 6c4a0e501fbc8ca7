// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// CoarseTracker restated: template construction, SSE path (calcRes/calcGSSSE/LM), g2o path
// (EdgeSE3PosePhotoDSO + restated g2o Levenberg). File:line citations into /root/reference/src.
#include "oracle_core.hpp"
#include <limits>

namespace orc {

void CoarseTracker::init(const GlobalCalib* G_, const Settings& S_) {
  G = G_; S = S_;
  for (int l = 0; l < G->levels; l++) {
    size_t n = (size_t)G->w[l] * G->h[l];
    idepth[l].assign(n, 0); weightSums[l].assign(n, 0); weightSums_bak[l].assign(n, 0);
    pc_u[l].assign(n, 0); pc_v[l].assign(n, 0); pc_idepth[l].assign(n, 0); pc_color[l].assign(n, 0);
    pc_n[l] = 0;
  }
  size_t n0 = (size_t)G->w[0] * G->h[0] + 4;
  buf_warped_idepth.assign(n0, 0); buf_warped_u.assign(n0, 0); buf_warped_v.assign(n0, 0);
  buf_warped_dx.assign(n0, 0); buf_warped_dy.assign(n0, 0); buf_warped_residual.assign(n0, 0);
  buf_warped_weight.assign(n0, 0); buf_warped_refColor.assign(n0, 0);
}

// FullSystem/CoarseTracker.cpp:108-136
void CoarseTracker::makeK(const CalibHessian& HCalib) {
  w[0] = G->w[0]; h[0] = G->h[0];
  fx[0] = HCalib.fxl; fy[0] = HCalib.fyl; cx[0] = HCalib.cxl; cy[0] = HCalib.cyl;
  for (int level = 1; level < G->levels; ++level) {
    w[level] = w[0] >> level; h[level] = h[0] >> level;
    fx[level] = fx[level - 1] * 0.5;
    fy[level] = fy[level - 1] * 0.5;
    cx[level] = (cx[0] + 0.5) / ((int)1 << level) - 0.5;
    cy[level] = (cy[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  for (int level = 0; level < G->levels; ++level) {
    float Kl[9] = {fx[level], 0.0f, cx[level], 0.0f, fy[level], cy[level], 0.0f, 0.0f, 1.0f};
    for (int i = 0; i < 9; i++) K[level][i] = Kl[i];
    inverse3f(K[level], Ki[level]);
    fxi[level] = Ki[level][0]; fyi[level] = Ki[level][4]; cxi[level] = Ki[level][2]; cyi[level] = Ki[level][5];
  }
}

// FullSystem/CoarseTracker.cpp:275-534 with STEP1's per-point decision factored out (:350-354 is the splat)
void CoarseTracker::setRefFromSplats(const Frame* ref, const RefPoint* pts, int n, const double aff_g2l[2]) {
  lastRef = ref;
  lastRef_aff_g2l[0] = aff_g2l[0]; lastRef_aff_g2l[1] = aff_g2l[1];
  int L = G->levels;
  std::fill(idepth[0].begin(), idepth[0].end(), 0.0f);
  std::fill(weightSums[0].begin(), weightSums[0].end(), 0.0f);
  for (int i = 0; i < n; i++) {
    int u = pts[i].u + 0.5f;  // :302-303 rounded
    int v = pts[i].v + 0.5f;
    idepth[0][u + w[0] * v] += pts[i].idepth * pts[i].weight;  // :353
    weightSums[0][u + w[0] * v] += pts[i].weight;
  }
  // STEP2 :360-386 sum-pool
  for (int lvl = 1; lvl < L; lvl++) {
    int lvlm1 = lvl - 1;
    int wl = w[lvl], hl = h[lvl], wlm1 = w[lvlm1];
    float* idepth_l = idepth[lvl].data(); float* weightSums_l = weightSums[lvl].data();
    const float* idepth_lm = idepth[lvlm1].data(); const float* weightSums_lm = weightSums[lvlm1].data();
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        int bidx = 2 * x + 2 * y * wlm1;
        idepth_l[x + y * wl] = idepth_lm[bidx] + idepth_lm[bidx + 1] + idepth_lm[bidx + wlm1] + idepth_lm[bidx + wlm1 + 1];
        weightSums_l[x + y * wl] = weightSums_lm[bidx] + weightSums_lm[bidx + 1] + weightSums_lm[bidx + wlm1] + weightSums_lm[bidx + wlm1 + 1];
      }
  }
  // STEP3 :390-442 diagonal dilation on levels 0,1 ; STEP4 :446-488 axis dilation on levels >= 2
  for (int lvl = 0; lvl < L; lvl++) {
    int wh = w[lvl] * h[lvl] - w[lvl];
    int wl = w[lvl];
    float* weightSumsl = weightSums[lvl].data();
    float* weightSumsl_bak = weightSums_bak[lvl].data();
    memcpy(weightSumsl_bak, weightSumsl, (size_t)w[lvl] * h[lvl] * sizeof(float));
    float* idepthl = idepth[lvl].data();
    int offs[4];
    if (lvl < 2) { offs[0] = 1 + wl; offs[1] = -1 - wl; offs[2] = wl - 1; offs[3] = -wl + 1; }
    else { offs[0] = 1; offs[1] = -1; offs[2] = wl; offs[3] = -wl; }
    for (int i = w[lvl]; i < wh; i++) {
      if (weightSumsl_bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        for (int k = 0; k < 4; k++) {
          int j = i + offs[k];
          if (j < 0 || j >= w[lvl] * h[lvl]) continue;  // the reference reads out of bounds here (UB); treated as no depth
          if (weightSumsl_bak[j] > 0) { sum += idepthl[j]; num += weightSumsl_bak[j]; numn++; }
        }
        if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
      }
    }
  }
  // STEP5 :492-533 normalise + raster-order compaction
  for (int lvl = 0; lvl < L; lvl++) {
    float* weightSumsl = weightSums[lvl].data();
    float* idepthl = idepth[lvl].data();
    const float* dIRefl = lastRef->dIp[lvl].data();
    int wl = w[lvl], hl = h[lvl];
    int lpc_n = 0;
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        int i = x + y * wl;
        if (weightSumsl[i] > 0) {
          idepthl[i] /= weightSumsl[i];
          pc_u[lvl][lpc_n] = x; pc_v[lvl][lpc_n] = y;
          pc_idepth[lvl][lpc_n] = idepthl[i];
          pc_color[lvl][lpc_n] = dIRefl[3 * i];
          if (!std::isfinite(pc_color[lvl][lpc_n]) || !(idepthl[i] > 0)) { idepthl[i] = -1; continue; }
          lpc_n++;
        } else
          idepthl[i] = -1;
        weightSumsl[i] = 1;
      }
    pc_n[lvl] = lpc_n;
  }
}

static inline void mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
    C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}

// ================================ SSE path =====================================================
// FullSystem/CoarseTracker.cpp:600-792 with the commented SSE body (:699-775) live.
void CoarseTracker::calcResSSE(int lvl, const SE3& refToNew, const double aff_g2l[2], float cutoffTH, double rs[6]) {
  float E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  int wl = w[lvl], hl = h[lvl];
  const float* dINewl = newFrame->dIp[lvl].data();
  float fxl = fx[lvl], fyl = fy[lvl], cxl = cx[lvl], cyl = cy[lvl];
  double Rd[9]; refToNew.rotationMatrix(Rd);
  float Rf[9]; for (int i = 0; i < 9; i++) Rf[i] = (float)Rd[i];
  float RKi[9]; mat33f_mul(Rf, Ki[lvl], RKi);
  float t[3] = {(float)refToNew.t[0], (float)refToNew.t[1], (float)refToNew.t[2]};
  double affLLd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], aff_g2l[0], aff_g2l[1], affLLd);
  float affLL[2] = {(float)affLLd[0], (float)affLLd[1]};
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  float maxEnergy = 2 * S.huberTH * cutoffTH - S.huberTH * S.huberTH;
  int nl = pc_n[lvl];
  const float* lpc_u = pc_u[lvl].data(); const float* lpc_v = pc_v[lvl].data();
  const float* lpc_idepth = pc_idepth[lvl].data(); const float* lpc_color = pc_color[lvl].data();
  const float* Kil = Ki[lvl];
  for (int i = 0; i < nl; i++) {
    float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    float pt[3];
    for (int r = 0; r < 3; r++) pt[r] = (RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2] * 1.0f) + t[r] * id;
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    float new_idepth = id / pt[2];
    evals++;
    if (lvl == 0 && i % 32 == 0) {
      float ptT[3], ptT2[3], pt3[3];
      for (int r = 0; r < 3; r++) {
        float kp = Kil[r * 3 + 0] * x + Kil[r * 3 + 1] * y + Kil[r * 3 + 2] * 1.0f;
        ptT[r] = kp + t[r] * id;
        ptT2[r] = kp - t[r] * id;
        pt3[r] = (RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2] * 1.0f) - t[r] * id;
      }
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    float refColor = lpc_color[i];
    float hitColor[3];
    getInterpolatedElement33(dINewl, Ku, Kv, wl, hitColor);
    if (!std::isfinite((float)hitColor[0])) continue;
    float residual = hitColor[0] - (float)(affLL[0] * refColor + affLL[1]);
    float hw = fabs(residual) < S.huberTH ? 1 : S.huberTH / fabs(residual);
    if (fabs(residual) > cutoffTH) {
      E += maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      buf_warped_idepth[numTermsInWarped] = new_idepth;
      buf_warped_u[numTermsInWarped] = u;
      buf_warped_v[numTermsInWarped] = v;
      buf_warped_dx[numTermsInWarped] = hitColor[1];
      buf_warped_dy[numTermsInWarped] = hitColor[2];
      buf_warped_residual[numTermsInWarped] = residual;
      buf_warped_weight[numTermsInWarped] = hw;
      buf_warped_refColor[numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    buf_warped_idepth[numTermsInWarped] = 0; buf_warped_u[numTermsInWarped] = 0; buf_warped_v[numTermsInWarped] = 0;
    buf_warped_dx[numTermsInWarped] = 0; buf_warped_dy[numTermsInWarped] = 0; buf_warped_residual[numTermsInWarped] = 0;
    buf_warped_weight[numTermsInWarped] = 0; buf_warped_refColor[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  buf_warped_n = numTermsInWarped;
  rs[0] = E; rs[1] = numTermsInE;
  rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
  rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
  rs[5] = numSaturated / (float)numTermsInE;
}

namespace {
// OptimizationBackend/MatrixAccumulators.h:907-1278 — 45 upper-tri entries x 4 SSE lanes, 3-tier shift-up
struct Accumulator9 {
  float SSEData[4 * 45], SSEData1k[4 * 45], SSEData1m[4 * 45];
  float numIn1, numIn1k, numIn1m;
  float H[81];
  void initialize() {
    memset(SSEData, 0, sizeof(SSEData)); memset(SSEData1k, 0, sizeof(SSEData1k)); memset(SSEData1m, 0, sizeof(SSEData1m));
    numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 180; i++) SSEData1k[i] = SSEData[i] + SSEData1k[i];
      numIn1k += numIn1; numIn1 = 0; memset(SSEData, 0, sizeof(SSEData));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 180; i++) SSEData1m[i] = SSEData1k[i] + SSEData1m[i];
      numIn1m += numIn1k; numIn1k = 0; memset(SSEData1k, 0, sizeof(SSEData1k));
    }
  }
  // :1025-1100 updateSSE_eighted on 4 lanes
  void updateSSE_eighted(const float J[9][4], const float w[4]) {
    float* pt = SSEData;
    for (int r = 0; r < 9; r++) {
      float Jw[4]; for (int l = 0; l < 4; l++) Jw[l] = J[r][l] * w[l];
      for (int c = r; c < 9; c++) { for (int l = 0; l < 4; l++) pt[l] = pt[l] + Jw[l] * J[c][l]; pt += 4; }
    }
    numIn1++;
    shiftUp(false);
  }
  void finish() {
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 9; r++) for (int c = r; c < 9; c++) {
      float d = SSEData1m[idx + 0] + SSEData1m[idx + 1] + SSEData1m[idx + 2] + SSEData1m[idx + 3];
      H[r * 9 + c] = H[c * 9 + r] = d; idx += 4;
    }
  }
};
}  // namespace

// FullSystem/CoarseTracker.cpp:537-596
void CoarseTracker::calcGSSSE(int lvl, double H_out[64], double b_out[8], const SE3& /*refToNew*/, const double aff_g2l[2]) {
  Accumulator9 acc; acc.initialize();
  float fxl = fx[lvl], fyl = fy[lvl];
  float b0 = (float)lastRef_aff_g2l[1];
  double affd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], aff_g2l[0], aff_g2l[1], affd);
  float a = (float)affd[0];
  int n = buf_warped_n;
  for (int i = 0; i < n; i += 4) {
    float J[9][4], wv[4];
    for (int l = 0; l < 4; l++) {
      float dx = buf_warped_dx[i + l] * fxl;
      float dy = buf_warped_dy[i + l] * fyl;
      float u = buf_warped_u[i + l], v = buf_warped_v[i + l], id = buf_warped_idepth[i + l];
      J[0][l] = id * dx;
      J[1][l] = id * dy;
      J[2][l] = 0.0f - id * (u * dx + v * dy);
      J[3][l] = 0.0f - ((u * v) * dx + dy * (1.0f + v * v));
      J[4][l] = (u * v) * dy + dx * (1.0f + u * u);
      J[5][l] = u * dy - v * dx;
      J[6][l] = a * (b0 - buf_warped_refColor[i + l]);
      J[7][l] = -1.0f;
      J[8][l] = buf_warped_residual[i + l];
      wv[l] = buf_warped_weight[i + l];
    }
    acc.updateSSE_eighted(J, wv);
  }
  acc.finish();
  float invn = 1.0f / n;
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] = (double)acc.H[r * 9 + c] * invn;
  for (int r = 0; r < 8; r++) b_out[r] = (double)acc.H[r * 9 + 8] * invn;
  // :584-595 (scale names swapped w.r.t. the J ordering; replicated as written)
  const double sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_A, SCALE_B};
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[c];
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[r];
  for (int r = 0; r < 8; r++) b_out[r] *= sc[r];
}

// FullSystem/CoarseTracker.cpp:827-1069, commented SSE control flow (:888-1023) live
bool CoarseTracker::trackNewestCoarseSSE(const Frame* fh, SE3& lastToNew_out, double aff_g2l_out[2], int coarsestLvl,
                                         const double minResForAbort[5], int* iterations_out) {
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
  newFrame = fh;
  int maxIterations[] = {10, 20, 50, 50, 50};
  float lambdaExtrapolationLimit = 0.001;
  SE3 refToNew_current = lastToNew_out;
  double aff_g2l_current[2] = {aff_g2l_out[0], aff_g2l_out[1]};
  bool haveRepeated = false;
  if (iterations_out) for (int i = 0; i < 5; i++) iterations_out[i] = 0;
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6];
    calcResSSE(lvl, refToNew_current, aff_g2l_current, S.coarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      calcResSSE(lvl, refToNew_current, aff_g2l_current, S.coarseCutoffTH * levelCutoffRepeat, resOld);
    }
    calcGSSSE(lvl, H, b, refToNew_current, aff_g2l_current);
    float lambda = 0.01;
    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      if (iterations_out) iterations_out[lvl]++;
      double Hl[64]; memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
      double nb[8]; for (int i = 0; i < 8; i++) nb[i] = -b[i];
      double inc[8];
      ldlt_solve(8, Hl, nb, inc);
      if (S.affineOptModeA < 0 && S.affineOptModeB < 0) {
        double H6[36], nb6[6], inc6[6];
        for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) H6[r * 6 + c] = Hl[r * 8 + c]; nb6[r] = -b[r]; }
        ldlt_solve(6, H6, nb6, inc6);
        for (int r = 0; r < 6; r++) inc[r] = inc6[r];
        inc[6] = inc[7] = 0;
      }
      if (!(S.affineOptModeA < 0) && S.affineOptModeB < 0) {
        double H7[49], nb7[7], inc7[7];
        for (int r = 0; r < 7; r++) { for (int c = 0; c < 7; c++) H7[r * 7 + c] = Hl[r * 8 + c]; nb7[r] = -b[r]; }
        ldlt_solve(7, H7, nb7, inc7);
        for (int r = 0; r < 7; r++) inc[r] = inc7[r];
        inc[7] = 0;
      }
      if (S.affineOptModeA < 0 && !(S.affineOptModeB < 0)) {
        double HlS[64], bS[8];
        memcpy(HlS, Hl, sizeof(HlS)); memcpy(bS, b, sizeof(bS));
        for (int r = 0; r < 8; r++) HlS[r * 8 + 6] = HlS[r * 8 + 7];
        for (int c = 0; c < 8; c++) HlS[6 * 8 + c] = HlS[7 * 8 + c];
        bS[6] = bS[7];
        double H7[49], nb7[7], inc7[7];
        for (int r = 0; r < 7; r++) { for (int c = 0; c < 7; c++) H7[r * 7 + c] = HlS[r * 8 + c]; nb7[r] = -bS[r]; }
        ldlt_solve(7, H7, nb7, inc7);
        for (int r = 0; r < 8; r++) inc[r] = 0;
        for (int r = 0; r < 6; r++) inc[r] = inc7[r];
        inc[6] = 0; inc[7] = inc7[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8]; memcpy(incScaled, inc, sizeof(inc));
      for (int i = 0; i < 3; i++) incScaled[i] *= SCALE_XI_ROT;
      for (int i = 3; i < 6; i++) incScaled[i] *= SCALE_XI_TRANS;
      incScaled[6] *= SCALE_A; incScaled[7] *= SCALE_B;
      double s = 0; for (int i = 0; i < 8; i++) s += incScaled[i];
      if (!std::isfinite(s)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
      SE3 refToNew_new = SE3::exp(incScaled) * refToNew_current;
      double aff_g2l_new[2] = {aff_g2l_current[0] + incScaled[6], aff_g2l_current[1] + incScaled[7]};
      double resNew[6];
      calcResSSE(lvl, refToNew_new, aff_g2l_new, S.coarseCutoffTH * levelCutoffRepeat, resNew);
      bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (accept) {
        calcGSSSE(lvl, H, b, refToNew_new, aff_g2l_new);
        memcpy(resOld, resNew, sizeof(resOld));
        aff_g2l_current[0] = aff_g2l_new[0]; aff_g2l_current[1] = aff_g2l_new[1];
        refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0; for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
      nrm = std::sqrt(nrm);
      if (!(nrm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = resOld[2 + i];
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return false;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  lastToNew_out = refToNew_current;
  aff_g2l_out[0] = aff_g2l_current[0]; aff_g2l_out[1] = aff_g2l_current[1];
  // :1050-1066
  if ((S.affineOptModeA != 0 && (fabsf((float)aff_g2l_out[0]) > 1.2)) || (S.affineOptModeB != 0 && (fabsf((float)aff_g2l_out[1]) > 200)))
    return false;
  double relAffd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], aff_g2l_out[0], aff_g2l_out[1], relAffd);
  float relAff[2] = {(float)relAffd[0], (float)relAffd[1]};
  if ((S.affineOptModeA == 0 && (fabsf(logf((float)relAff[0])) > 1.5)) || (S.affineOptModeB == 0 && (fabsf((float)relAff[1]) > 200)))
    return false;
  if (S.affineOptModeA < 0) aff_g2l_out[0] = 0;
  if (S.affineOptModeB < 0) aff_g2l_out[1] = 0;
  return true;
}

// ================================ g2o path =====================================================
// dso_util.hpp:25-45
static inline bool CheckBoundary(double u, double v, int wl, int hl) {
  return (u - 2) < 0 || (u + 3) > wl || (v - 2) < 0 || (v + 3) > hl;
}

// dso_g2o_edge.cpp:395-423. Uses the GLOBAL initial intrinsics KG[level] (dso_util.hpp:10-22).
void CoarseTracker::edgeComputeError(Edge& e, const SE3& pose, const double photo[2]) const {
  double Xr[3] = {e.Xref[0], e.Xref[1], e.Xref[2]}, Xc[3];
  pose.act(Xr, Xc);
  double fxg = G->K[e.level][0], fyg = G->K[e.level][4], cxg = G->K[e.level][2], cyg = G->K[e.level][5];
  double uu = fxg * (Xc[0] / Xc[2]) + cxg, vv = fyg * (Xc[1] / Xc[2]) + cyg;
  int wl = w[e.level], hl = h[e.level];
  if (CheckBoundary(uu, vv, wl, hl)) { e.error = 0.0; return; }
  double abd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], photo[0], photo[1], abd);
  float ab[2] = {(float)abd[0], (float)abd[1]};
  float hit[3];
  getInterpolatedElement33(newFrame->dIp[e.level].data(), (float)uu, (float)vv, wl, hit);
  if (!std::isfinite((float)hit[0])) return;  // error left stale
  e.error = hit[0] - (ab[0] * e.measurement + ab[1]);  // float*double promotes to double
}

// dso_g2o_edge.cpp:425-500 (VERSION2)
bool CoarseTracker::edgeLinearizeOplus(const Edge& e, const SE3& pose, const double photo[2], double Jp[6], double Ja[2]) const {
  double Xr[3] = {e.Xref[0], e.Xref[1], e.Xref[2]}, Xc[3];
  pose.act(Xr, Xc);
  double fxg = G->K[e.level][0], fyg = G->K[e.level][4], cxg = G->K[e.level][2], cyg = G->K[e.level][5];
  double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2];
  double uu = fxg * (Xc[0] / Xc[2]) + cxg, vv = fyg * (Xc[1] / Xc[2]) + cyg;
  int wl = w[e.level], hl = h[e.level];
  if (CheckBoundary(uu, vv, wl, hl)) { for (int i = 0; i < 6; i++) Jp[i] = 0; Ja[0] = Ja[1] = 0; return false; }
  float hit[3];
  getInterpolatedElement33(newFrame->dIp[e.level].data(), (float)uu, (float)vv, wl, hit);
  double u = x * invz, v = y * invz;
  double dx = hit[1] * fxg, dy = hit[2] * fyg;
  Jp[0] = invz * dx;
  Jp[1] = invz * dy;
  Jp[2] = -invz * (u * dx + v * dy);
  Jp[3] = -(u * v * dx + (1 + v * v) * dy);
  Jp[4] = u * v * dy + (1 + u * u) * dx;
  Jp[5] = u * dy - v * dx;
  double abd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], photo[0], photo[1], abd);
  float ab0 = (float)abd[0];
  Ja[0] = ab0 * (lastRef_aff_g2l[1] - e.measurement);
  Ja[1] = -1;
  return true;
}

// FullSystem/CoarseTracker.cpp:600-792 (live g2o body)
void CoarseTracker::calcResG2O(int lvl, const SE3& refToNew, float cutoffTH, const SE3& vtx_pose, const double vtx_photo[2], double rs[6]) {
  float E = 0;
  int numTermsInE = 0, numSaturated = 0;
  int wl = w[lvl], hl = h[lvl];
  float fxl = fx[lvl], fyl = fy[lvl], cxl = cx[lvl], cyl = cy[lvl];
  double Rd[9]; refToNew.rotationMatrix(Rd);
  float Rf[9]; for (int i = 0; i < 9; i++) Rf[i] = (float)Rd[i];
  float RKi[9]; mat33f_mul(Rf, Ki[lvl], RKi);
  float t[3] = {(float)refToNew.t[0], (float)refToNew.t[1], (float)refToNew.t[2]};
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  int nl = pc_n[lvl];
  const float* Kil = Ki[lvl];
  for (int i = 0; i < nl; i++) {
    float id = pc_idepth[lvl][i], x = pc_u[lvl][i], y = pc_v[lvl][i];
    float pt[3];
    for (int r = 0; r < 3; r++) pt[r] = (RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2] * 1.0f) + t[r] * id;
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    float new_idepth = id / pt[2];
    if (lvl == 0 && i % 32 == 0) {
      float ptT[3], ptT2[3], pt3[3];
      for (int r = 0; r < 3; r++) {
        float kp = Kil[r * 3 + 0] * x + Kil[r * 3 + 1] * y + Kil[r * 3 + 2] * 1.0f;
        ptT[r] = kp + t[r] * id; ptT2[r] = kp - t[r] * id;
        pt3[r] = (RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2] * 1.0f) - t[r] * id;
      }
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    Edge e;
    // :707  Xref = Ki[lvl] * Vec3f(x,y,1) / id
    for (int r = 0; r < 3; r++) e.Xref[r] = (Kil[r * 3 + 0] * x + Kil[r * 3 + 1] * y + Kil[r * 3 + 2] * 1.0f) / id;
    e.level = lvl; e.edge_level = lvl; e.measurement = pc_color[lvl][i]; e.error = 0.0;  // g2o zero-inits? see note
    edgeComputeError(e, vtx_pose, vtx_photo);
    evals++;
    if (e.error > cutoffTH * 10) { numSaturated++; continue; }  // :723 (signed)
    edges.push_back(e);
    numTermsInE++;
  }
  rs[0] = E; rs[1] = numTermsInE;
  rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
  rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
  rs[5] = numSaturated / (float)numTermsInE;
}

// g2o RobustKernelHuber::robustify (restated)
static inline void huberRobustify(double e, double delta, double rho[3]) {
  double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else { double sqrte = std::sqrt(e); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e; }
}

// Cholesky (LLT) solve as g2o::LinearSolverEigen does (SimplicialLLT); fails on non-PD.
static bool llt_solve(int n, const double* A, const double* b, double* x) {
  std::vector<double> L(n * n, 0.0);
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0) || !std::isfinite(s)) return false;
    L[j * n + j] = std::sqrt(s);
    for (int i = j + 1; i < n; i++) {
      double v = A[i * n + j];
      for (int k = 0; k < j; k++) v -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = v / L[j * n + j];
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * n + k] * y[k]; y[i] = v / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < n; k++) v -= L[k * n + i] * x[k]; x[i] = v / L[i * n + i]; }
  return true;
}

// FullSystem/CoarseTracker.cpp:827-1069 live body + restated g2o
bool CoarseTracker::trackNewestCoarseG2O(const Frame* fh, SE3& lastToNew_out, double aff_g2l_out[2], int coarsestLvl,
                                         const double minResForAbort[5], int* lm_iterations_out) {
  for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
  newFrame = fh;
  edges.clear();
  int maxIterations[] = {2, 2, 2, 2, 2};
  SE3 vtx_pose = lastToNew_out;                                  // VertexSE3PoseDSO estimate
  double vtx_photo[2] = {aff_g2l_out[0], aff_g2l_out[1]};        // VertexPhotometricDSO estimate
  SE3 refToNew_current = lastToNew_out;                          // never updated (:880, Appendix A.1)
  bool forceStop = false;                                        // SparseOptimizerTerminateAction's flag
  double lastChi = 0;
  const double delta = S.huberTH;
  if (lm_iterations_out) for (int i = 0; i < 5; i++) lm_iterations_out[i] = 0;
  g2o_trials = g2o_rejected = 0;

  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    double resOld[6];
    calcResG2O(lvl, refToNew_current, S.coarseCutoffTH * 1.0f, vtx_pose, vtx_photo, resOld);
    // initializeOptimization(lvl): active edges = edges with level()==lvl
    std::vector<int> act;
    for (size_t k = 0; k < edges.size(); k++) if (edges[k].edge_level == lvl) act.push_back((int)k);
    auto computeActiveErrors = [&]() { for (int k : act) { edgeComputeError(edges[k], vtx_pose, vtx_photo); evals++; } };
    auto activeRobustChi2 = [&]() { double chi = 0; for (int k : act) { double rho[3]; huberRobustify(edges[k].error * edges[k].error, delta, rho); chi += rho[0]; } return chi; };

    if (!act.empty()) {  // optimize(): returns -1 when there are no active vertices
      double lambda = 0, ni = 2;
      bool ok = true;
      for (int it = 0; it < maxIterations[lvl] && !forceStop && ok; it++) {
        if (lm_iterations_out) lm_iterations_out[lvl]++;
        // ---- OptimizationAlgorithmLevenberg::solve(it)
        computeActiveErrors();
        double currentChi = activeRobustChi2();
        double tempChi = currentChi;
        double H[64], b[8];
        memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
        for (int k : act) {  // buildSystem: linearizeOplus + constructQuadraticForm
          double J[8];
          edgeLinearizeOplus(edges[k], vtx_pose, vtx_photo, J, J + 6); evals++;
          double err = edges[k].error;
          double rho[3]; huberRobustify(err * err, delta, rho);
          double omega_r = -err * rho[1];
          for (int r = 0; r < 8; r++) { b[r] += J[r] * omega_r; for (int c = 0; c < 8; c++) H[r * 8 + c] += J[r] * rho[1] * J[c]; }
        }
        if (it == 0) { lambda = 0.01; ni = 2; }  // setUserLambdaInit(0.01) :840
        double rho = 0; int qmax = 0;
        do {
          SE3 pose_bak = vtx_pose; double photo_bak[2] = {vtx_photo[0], vtx_photo[1]};  // push()
          double Hl[64]; memcpy(Hl, H, sizeof(Hl));
          for (int i = 0; i < 8; i++) Hl[i * 8 + i] += lambda;  // additive damping
          double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          bool ok2 = llt_solve(8, Hl, b, x);
          // oplus: dso_g2o_vertex.cpp:15-18, 30-40
          vtx_pose = SE3::exp(x) * vtx_pose;
          vtx_photo[0] += x[6]; vtx_photo[1] += x[7];
          computeActiveErrors();
          tempChi = activeRobustChi2();
          if (!ok2) tempChi = std::numeric_limits<double>::max();
          rho = (currentChi - tempChi);
          double scale = 0; for (int j = 0; j < 8; j++) scale += x[j] * (lambda * x[j] + b[j]);
          scale += 1e-3;
          rho /= scale;
          if (rho > 0 && std::isfinite(tempChi)) {
            double alpha = 1. - std::pow((2 * rho - 1), 3);
            alpha = std::min(alpha, 2. / 3.);
            double scaleFactor = std::max(1. / 3., alpha);
            lambda *= scaleFactor; ni = 2; currentChi = tempChi;
          } else {
            lambda *= ni; ni *= 2;
            vtx_pose = pose_bak; vtx_photo[0] = photo_bak[0]; vtx_photo[1] = photo_bak[1];  // pop()
            g2o_rejected++;
          }
          qmax++; g2o_trials++;
        } while (rho < 0 && qmax < 10 && !forceStop);
        if (qmax == 10 || rho == 0) ok = false;  // SolverResult::Terminate
        // ---- postIteration(it): SparseOptimizerTerminateAction, gain threshold 1e-3 (:845-848)
        computeActiveErrors();
        if (it == 0) lastChi = activeRobustChi2();
        else {
          double chi = activeRobustChi2();
          double gain = (lastChi - chi) / chi;
          lastChi = chi;
          if (gain >= 0 && gain < 1e-3) forceStop = true;
        }
      }
    }
    // :1029 — divides by ALL edges in the graph (Appendix A.2)
    lastResiduals[lvl] = sqrtf((float)activeRobustChi2() / edges.size());
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = resOld[2 + i];
    if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return false;
  }
  lastToNew_out = vtx_pose;
  aff_g2l_out[0] = vtx_photo[0]; aff_g2l_out[1] = vtx_photo[1];
  if ((S.affineOptModeA != 0 && (fabsf((float)aff_g2l_out[0]) > 1.2)) || (S.affineOptModeB != 0 && (fabsf((float)aff_g2l_out[1]) > 200)))
    return false;
  double relAffd[2];
  affFromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l[0], lastRef_aff_g2l[1], aff_g2l_out[0], aff_g2l_out[1], relAffd);
  float relAff[2] = {(float)relAffd[0], (float)relAffd[1]};
  if ((S.affineOptModeA == 0 && (fabsf(logf((float)relAff[0])) > 1.5)) || (S.affineOptModeB == 0 && (fabsf((float)relAff[1]) > 200)))
    return false;
  if (S.affineOptModeA < 0) aff_g2l_out[0] = 0;
  if (S.affineOptModeB < 0) aff_g2l_out[1] = 0;
  return true;
}

}  // namespace orc
