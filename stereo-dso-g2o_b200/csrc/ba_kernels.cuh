// Kernels of the windowed bundle adjustment (B1-B9), SSE-path arithmetic of the reference:
//   ba_linearize_kernel   PointFrameResidual::linearize (Residuals.cpp:83-336) [+ applyRes/takeDataF when fix]
//   ba_apply_res_kernel   PointFrameResidual::applyRes (:367-385) + EFResidual::takeDataF (EnergyFunctionalStructs.cpp:37-51)
//   ba_fixlin_kernel      EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:96-123)
//   ba_top_kernel         AccumulatedTopHessianSSE::addPoint<mode> (AccumulatedTopHessian.cpp:36-193), residual part
//   ba_point_sums_kernel  ... its per-point tail (bd_acc, Hdd_acc, Hcd_acc -> EFPoint::*_acc{A,L}F, :160-192)
//   ba_sc_point_kernel    AccumulatedSCHessianSSE::addPoint, per-point part (AccumulatedSCHessian.cpp:34-75)
//   ba_sc_pair_kernel     ... its O(res^2) part: accE, accEB, accD (:77-102)
//   ba_*_finish_kernel    fixed-order sums of the per-CTA partials (AccumulatorApprox::finish etc.)
//   ba_stitch_top_kernel  AccumulatedTopHessianSSE::stitchDoubleInternal + the symmetrisation of stitchDoubleMT
//   ba_sc_uv_kernel + ba_stitch_sc_kernel   AccumulatedSCHessianSSE::stitchDoubleInternal (:106-195)
//   ba_solve_kernel       EnergyFunctional::solveSystemF (EnergyFunctional.cpp:838-995) + orthogonalize (:775-835)
//   ba_xad_kernel + ba_resub_kernel   resubstituteF_MT / resubstituteFPt (:272-341)
// All float sums are atomics-free and fixed-order (reproducible); decision paths (OOB tests, Huber switch,
// energy thresholds) run un-fused in the reference's operand order (-fmad=false), thread-sequential over
// the 8 pattern pixels, so ResState / energies are bit-identical to an un-fused CPU evaluation.
#pragma once
#include "ba_state.h"

namespace sdso {

struct BAView {  // plain pointers handed to the kernels
  int n, P, R, capP, capR;
  BACalib c;
  const float4* const* tex0; const float* frameTH; const PrecalcDev* precalc;
  const double* adHost; const double* adTarget; const float* adHostF; const float* adTargetF;
  const float* adHTdeltaF; const float* cDeltaF; const double* fprior;
  const int* p_host; const float* p_u; const float* p_v; const float* p_idepth; const float* p_idepth_zero;
  const float4* p_color; const float4* p_weights; const float* p_priorF; const float* p_deltaF;
  const int* p_res_begin; const int* p_res_list; const int* slot_of; float* p_acc; const unsigned char* p_flag;
  const int* s_point; const int* s_key;
  unsigned char* s_state; unsigned char* s_newstate; unsigned char* s_flags; unsigned char* s_sel;
  float* s_energy; float* J; float* s_rtz; float* s_JpJd; float* s_center; float* s_psum;
  const Chunk* chunks; int nchunks; const int* key_chunk_begin;
  float* tpart; float* dpart; float* pblockpart;
  double* G; float* Gf; double* D; double* E; double* Hcc; double* U; double* V;
  double* energy_part; double* scalars; unsigned int* counter;
};

__device__ __forceinline__ float* jplane(const BAView& B, int buf, int plane, int s) {
  return B.J + ((size_t)buf * kJ + plane) * B.capR + s;
}

// deterministic block sum of one double per thread; result valid in thread 0
__device__ __forceinline__ double block_sum_d(double v, double* smem /* >= 32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double s = 0;
  if (threadIdx.x == 0) for (int i = 0; i < nw; i++) s += smem[i];
  return s;
}

// takeDataF tail: JpJdF = [Jpdxi^T (JIdx2 Jpdd) ; JabJIdx Jpdd]
__device__ __forceinline__ void compute_JpJd(const BAView& B, int buf, int s) {
  const float Jpdd0 = *jplane(B, buf, J_PDD, s), Jpdd1 = *jplane(B, buf, J_PDD + 1, s);
  const float i00 = *jplane(B, buf, J_IDX2, s), i01 = *jplane(B, buf, J_IDX2 + 1, s), i10 = *jplane(B, buf, J_IDX2 + 2, s), i11 = *jplane(B, buf, J_IDX2 + 3, s);
  const float v0 = i00 * Jpdd0 + i01 * Jpdd1;
  const float v1 = i10 * Jpdd0 + i11 * Jpdd1;
#pragma unroll
  for (int i = 0; i < 6; i++) B.s_JpJd[(size_t)i * B.capR + s] = *jplane(B, buf, J_PDXI + i, s) * v0 + *jplane(B, buf, J_PDXI + 6 + i, s) * v1;
  const float a00 = *jplane(B, buf, J_ABIDX, s), a01 = *jplane(B, buf, J_ABIDX + 1, s), a10 = *jplane(B, buf, J_ABIDX + 2, s), a11 = *jplane(B, buf, J_ABIDX + 3, s);
  B.s_JpJd[(size_t)6 * B.capR + s] = a00 * Jpdd0 + a01 * Jpdd1;
  B.s_JpJd[(size_t)7 * B.capR + s] = a10 * Jpdd0 + a11 * Jpdd1;
}

// applyRes(copyJacobians) for one slot
__device__ __forceinline__ void apply_res_slot(const BAView& B, int s, bool copyJ) {
  const unsigned char st = B.s_state[s], ns = B.s_newstate[s];
  if (copyJ) {
    if (st == RS_OOB) return;  // (the reference asserts !efResidual->isActive() and returns)
    unsigned char fl = B.s_flags[s];
    if (ns == RS_IN) {
      fl |= RF_ACTIVE;
      const unsigned char sel = B.s_sel[s] ^ 1;  // takeDataF: std::swap(J, data->J)
      B.s_sel[s] = sel;
      compute_JpJd(B, sel, s);
    } else {
      fl &= ~RF_ACTIVE;
    }
    B.s_flags[s] = fl;
  }
  B.s_state[s] = ns;
  B.s_energy[s] = B.s_energy[(size_t)B.capR + s];
}

// ---- B1 ------------------------------------------------------------------------------------------
__device__ double linearize_slot(const BAView& B, int s, bool fix) {
  const size_t cR = B.capR;
  if (B.s_flags[s] & RF_LINEARIZED) return 0.0;  // activeResiduals = residuals that are not linearised (FullSystemOptimize.cpp:900-902)
  B.s_energy[2 * cR + s] = -1;
  const float state_energy = B.s_energy[s];
  if (B.s_state[s] == RS_OOB) { B.s_newstate[s] = RS_OOB; return state_energy; }
  const int pidx = B.s_point[s];
  const int key = B.s_key[s];
  const int h = key % B.n, t = key / B.n;
  const PrecalcDev& pc = B.precalc[h * B.n + t];
  const BACalib& c = B.c;
  const float pu = B.p_u[pidx], pv = B.p_v[pidx];
  const float idepth_scaled = SCALE_IDEPTH * B.p_idepth[pidx];
  const float idepth_zero_scaled = SCALE_IDEPTH * B.p_idepth_zero[pidx];
  const int buf = B.s_sel[s] ^ 1;  // candidate buffer (PointFrameResidual::J)
  float Jpdd0, Jpdd1;
  {
    // projectPoint (ResidualProjections.h:64-96) at the FEJ evaluation point
    const float* Rm = pc.PRE_RTll_0; const float* tt = pc.PRE_tTll_0;
    const float K0 = (pu + 0 - c.cxl) * c.fxli, K1 = (pv + 0 - c.cyl) * c.fyli;
    float ptp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) ptp[k] = (Rm[k * 3] * K0 + Rm[k * 3 + 1] * K1 + Rm[k * 3 + 2] * 1.0f) + tt[k] * idepth_zero_scaled;
    const float drescale = 1.0f / ptp[2];
    const float new_idepth = idepth_zero_scaled * drescale;
    bool ok = (drescale > 0);
    float u = 0, v = 0, Ku = 0, Kv = 0;
    if (ok) {
      u = ptp[0] * drescale; v = ptp[1] * drescale;
      Ku = u * c.fxl + c.cxl; Kv = v * c.fyl + c.cyl;
      ok = Ku > 1.1f && Kv > 1.1f && Ku < c.wM3G && Kv < c.hM3G;
    }
    if (!ok) { B.s_newstate[s] = RS_OOB; return state_energy; }
    B.s_center[s] = Ku; B.s_center[cR + s] = Kv; B.s_center[2 * cR + s] = new_idepth;
    Jpdd0 = drescale * (tt[0] - tt[2] * u) * SCALE_IDEPTH * c.fxl;
    Jpdd1 = drescale * (tt[1] - tt[2] * v) * SCALE_IDEPTH * c.fyl;
    float dCx[4], dCy[4];
    dCx[2] = drescale * (Rm[6] * u - Rm[0]);
    dCx[3] = c.fxl * drescale * (Rm[7] * u - Rm[1]) * c.fyli;
    dCx[0] = K0 * dCx[2];
    dCx[1] = K1 * dCx[3];
    dCy[2] = c.fyl * drescale * (Rm[6] * v - Rm[3]) * c.fxli;
    dCy[3] = drescale * (Rm[7] * v - Rm[4]);
    dCy[0] = K0 * dCy[2];
    dCy[1] = K1 * dCy[3];
    dCx[0] = (dCx[0] + u) * SCALE_F;
    dCx[1] *= SCALE_F;
    dCx[2] = (dCx[2] + 1) * SCALE_C;
    dCx[3] *= SCALE_C;
    dCy[0] *= SCALE_F;
    dCy[1] = (dCy[1] + v) * SCALE_F;
    dCy[2] *= SCALE_C;
    dCy[3] = (dCy[3] + 1) * SCALE_C;
    float dx[6], dy[6];
    dx[0] = new_idepth * c.fxl; dx[1] = 0; dx[2] = -new_idepth * u * c.fxl;
    dx[3] = -u * v * c.fxl; dx[4] = (1 + u * u) * c.fxl; dx[5] = -v * c.fxl;
    dy[0] = 0; dy[1] = new_idepth * c.fyl; dy[2] = -new_idepth * v * c.fyl;
    dy[3] = -(1 + v * v) * c.fyl; dy[4] = u * v * c.fyl; dy[5] = u * c.fyl;
#pragma unroll
    for (int i = 0; i < 6; i++) { *jplane(B, buf, J_PDXI + i, s) = dx[i]; *jplane(B, buf, J_PDXI + 6 + i, s) = dy[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { *jplane(B, buf, J_PDC + i, s) = dCx[i]; *jplane(B, buf, J_PDC + 4 + i, s) = dCy[i]; }
    *jplane(B, buf, J_PDD, s) = Jpdd0; *jplane(B, buf, J_PDD + 1, s) = Jpdd1;
  }
  // pattern pixels at the current state: all eight projections first (the early-outs of the reference's loop have
  // no side effect other than the OOB verdict), then the 32 independent 16-byte gathers, then the arithmetic in order
  const float* KRKi = pc.PRE_KRKiTll; const float* Kt = pc.PRE_KtTll;
  float Ku[8], Kv[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    const float up = pu + kPatternP[idx][0], vp = pv + kPatternP[idx][1];
    float ptp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) ptp[k] = (KRKi[k * 3] * up + KRKi[k * 3 + 1] * vp + KRKi[k * 3 + 2] * 1.0f) + Kt[k] * idepth_scaled;
    Ku[idx] = ptp[0] / ptp[2]; Kv[idx] = ptp[1] / ptp[2];
    if (!(Ku[idx] > 1.1f && Kv[idx] > 1.1f && Ku[idx] < c.wM3G && Kv[idx] < c.hM3G)) { B.s_newstate[s] = RS_OOB; return state_energy; }
  }
  const float4* tex = B.tex0[t];
  float3 hit[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) hit[idx] = interp33(tex, Ku[idx], Kv[idx], c.w0);
#pragma unroll
  for (int idx = 0; idx < 8; idx++) if (!isfinite(hit[idx].x)) { B.s_newstate[s] = RS_OOB; return state_energy; }
  const float4 col0 = B.p_color[2 * pidx], col1 = B.p_color[2 * pidx + 1];
  const float4 wt0 = B.p_weights[2 * pidx], wt1 = B.p_weights[2 * pidx + 1];
  const float color[8] = {col0.x, col0.y, col0.z, col0.w, col1.x, col1.y, col1.z, col1.w};
  const float weights[8] = {wt0.x, wt0.y, wt0.z, wt0.w, wt1.x, wt1.y, wt1.z, wt1.w};
  const float aff0 = pc.PRE_aff_mode[0], aff1 = pc.PRE_aff_mode[1], b0 = pc.PRE_b0_mode;
  float energyLeft = 0, wJI2_sum = 0;
  float II00 = 0, II11 = 0, II10 = 0, AI00 = 0, AI01 = 0, AI10 = 0, AI11 = 0, AA00 = 0, AA01 = 0, AA11 = 0;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    float h0 = hit[idx].x, h1 = hit[idx].y, h2 = hit[idx].z;
    const float residual = h0 - (float)(aff0 * color[idx] + aff1);
    const float drdA = (color[idx] - b0);
    float w = sqrtf(c.outlierTHSumComponent / (c.outlierTHSumComponent + (h1 * h1 + h2 * h2)));
    w = 0.5f * (w + weights[idx]);
    float hw = fabsf(residual) < c.huberTH ? 1 : c.huberTH / fabsf(residual);
    energyLeft += w * w * hw * residual * residual * (2 - hw);
    if (hw < 1) hw = sqrtf(hw);
    hw = hw * w;
    h1 *= hw; h2 *= hw;
    *jplane(B, buf, J_RESF + idx, s) = residual * hw;
    *jplane(B, buf, J_IDX + idx, s) = h1;
    *jplane(B, buf, J_IDX + 8 + idx, s) = h2;
    *jplane(B, buf, J_AB + idx, s) = (c.affineOptModeA < 0) ? 0.f : drdA * hw;
    *jplane(B, buf, J_AB + 8 + idx, s) = (c.affineOptModeB < 0) ? 0.f : hw;
    II00 += h1 * h1; II11 += h2 * h2; II10 += h1 * h2;
    AI00 += drdA * hw * h1; AI01 += drdA * hw * h2; AI10 += hw * h1; AI11 += hw * h2;
    AA00 += drdA * drdA * hw * hw; AA01 += drdA * hw * hw; AA11 += hw * hw;
    wJI2_sum += hw * hw * (h1 * h1 + h2 * h2);
  }
  *jplane(B, buf, J_IDX2, s) = II00; *jplane(B, buf, J_IDX2 + 1, s) = II10; *jplane(B, buf, J_IDX2 + 2, s) = II10; *jplane(B, buf, J_IDX2 + 3, s) = II11;
  *jplane(B, buf, J_ABIDX, s) = AI00; *jplane(B, buf, J_ABIDX + 1, s) = AI01; *jplane(B, buf, J_ABIDX + 2, s) = AI10; *jplane(B, buf, J_ABIDX + 3, s) = AI11;
  *jplane(B, buf, J_AB2, s) = AA00; *jplane(B, buf, J_AB2 + 1, s) = AA01; *jplane(B, buf, J_AB2 + 2, s) = AA01; *jplane(B, buf, J_AB2 + 3, s) = AA11;
  B.s_energy[2 * cR + s] = energyLeft;
  const float th = fmaxf(B.frameTH[h], B.frameTH[t]);
  unsigned char ns = RS_IN;
  if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; ns = RS_OUTLIER; }
  B.s_newstate[s] = ns;
  B.s_energy[cR + s] = energyLeft;
  return energyLeft;
}

__global__ void __launch_bounds__(128) ba_linearize_kernel(BAView B, int fix) {
  __shared__ double red[32];
  __shared__ bool last;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0;
  if (s < B.R) {
    const bool active_res = !(B.s_flags[s] & RF_LINEARIZED);
    e = linearize_slot(B, s, fix != 0);
    if (fix && active_res) apply_res_slot(B, s, true);
  }
  const double bs = block_sum_d(e, red);
  if (threadIdx.x == 0) {
    B.energy_part[blockIdx.x] = bs;
    __threadfence();
    last = (atomicAdd(B.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {  // fixed-order final sum by the last CTA to finish
    __threadfence();
    double tot = 0;
    for (unsigned i = 0; i < gridDim.x; i++) tot += ((volatile double*)B.energy_part)[i];
    B.scalars[0] = tot;
    *B.counter = 0;
  }
}

__global__ void ba_apply_res_kernel(BAView B, int copyJ) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < B.R) apply_res_slot(B, s, copyJ != 0);
}

// res_toZeroF = resF - J * delta  (EnergyFunctionalStructs.cpp:96-123); list == nullptr: every active residual
__global__ void ba_fixlin_kernel(BAView B, const int* list, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int s = list ? list[i] : i;
  if (!list && !(B.s_flags[s] & RF_ACTIVE)) return;
  const int buf = B.s_sel[s];
  const float* dp = B.adHTdeltaF + (size_t)B.s_key[s] * 8;
  const float pdelta = B.p_deltaF[B.s_point[s]];
  const float* cd = B.cDeltaF;
  float jx[6], jy[6], cx[4], cy[4];
#pragma unroll
  for (int k = 0; k < 6; k++) { jx[k] = *jplane(B, buf, J_PDXI + k, s); jy[k] = *jplane(B, buf, J_PDXI + 6 + k, s); }
#pragma unroll
  for (int k = 0; k < 4; k++) { cx[k] = *jplane(B, buf, J_PDC + k, s); cy[k] = *jplane(B, buf, J_PDC + 4 + k, s); }
  const float Jpx = (jx[0] * dp[0] + jx[1] * dp[1] + jx[2] * dp[2] + jx[3] * dp[3] + jx[4] * dp[4] + jx[5] * dp[5]) +
                    (cx[0] * cd[0] + cx[1] * cd[1] + cx[2] * cd[2] + cx[3] * cd[3]) + *jplane(B, buf, J_PDD, s) * pdelta;
  const float Jpy = (jy[0] * dp[0] + jy[1] * dp[1] + jy[2] * dp[2] + jy[3] * dp[3] + jy[4] * dp[4] + jy[5] * dp[5]) +
                    (cy[0] * cd[0] + cy[1] * cd[1] + cy[2] * cd[2] + cy[3] * cd[3]) + *jplane(B, buf, J_PDD + 1, s) * pdelta;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float rtz = *jplane(B, buf, J_RESF + k, s);
    rtz = rtz - *jplane(B, buf, J_IDX + k, s) * Jpx;
    rtz = rtz - *jplane(B, buf, J_IDX + 8 + k, s) * Jpy;
    rtz = rtz - *jplane(B, buf, J_AB + k, s) * dp[6];
    rtz = rtz - *jplane(B, buf, J_AB + 8 + k, s) * dp[7];
    B.s_rtz[(size_t)k * B.capR + s] = rtz;
  }
  B.s_flags[s] |= RF_LINEARIZED;
}

// ---- register-transposing warp reduction: 32 values x 32 lanes -> lane L holds the warp sum of value L in a[0]
#define SDSO_TSTAGE(O, N)                                          \
  _Pragma("unroll") for (int j = 0; j < N; j++) {                  \
    const bool up = (lane & O) != 0;                               \
    const float keep = up ? a[j + N] : a[j];                       \
    const float send = up ? a[j] : a[j + N];                       \
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, O);           \
  }
__device__ __forceinline__ float warp_reduce_transpose32(float (&a)[32], int lane) {
  SDSO_TSTAGE(16, 16)
  SDSO_TSTAGE(8, 8)
  SDSO_TSTAGE(4, 4)
  SDSO_TSTAGE(2, 2)
  SDSO_TSTAGE(1, 1)
  return a[0];
}
#undef SDSO_TSTAGE

// ---- B4: residual part of addPoint<mode>; one CTA per chunk of <= 256 slots of one (host,target) pair ----------
__global__ void __launch_bounds__(kChunk) ba_top_kernel(BAView B, int mode) {
  __shared__ float wsum[kChunk / 32][kTopVals];
  const Chunk ch = B.chunks[blockIdx.x];
  const int s = ch.begin + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t cR = B.capR;
  float g0[32], g1[32], g2[32];  // Data[55] | TopRight[30] | BotRight[6], in three register groups
#pragma unroll
  for (int i = 0; i < 32; i++) { g0[i] = 0; g1[i] = 0; g2[i] = 0; }
  bool use = s < ch.end;
  int pidx = 0;
  if (use) {
    const unsigned char fl = B.s_flags[s];
    const bool lin = fl & RF_LINEARIZED, act = fl & RF_ACTIVE;
    pidx = B.s_point[s];
    if (mode == 0) use = !lin && act;
    if (mode == 1) use = lin && act;
    if (mode == 2) use = act && B.p_flag[pidx] == PS_MARGINALIZE;
  }
  float ps[6] = {0, 0, 0, 0, 0, 0};
  if (use) {
    const int buf = B.s_sel[s];
    float x[10], y[10];  // x = [Jpdc[0] ; Jpdxi[0]], y = [Jpdc[1] ; Jpdxi[1]]
#pragma unroll
    for (int k = 0; k < 4; k++) { x[k] = *jplane(B, buf, J_PDC + k, s); y[k] = *jplane(B, buf, J_PDC + 4 + k, s); }
#pragma unroll
    for (int k = 0; k < 6; k++) { x[4 + k] = *jplane(B, buf, J_PDXI + k, s); y[4 + k] = *jplane(B, buf, J_PDXI + 6 + k, s); }
    const float Jpdd0 = *jplane(B, buf, J_PDD, s), Jpdd1 = *jplane(B, buf, J_PDD + 1, s);
    const float* dp = B.adHTdeltaF + (size_t)ch.key * 8;
    float Jpx = 0, Jpy = 0;
    if (mode == 1) {
      const float* cd = B.cDeltaF;
      const float dd = B.p_deltaF[pidx];
      Jpx = (x[4] * dp[0] + x[5] * dp[1] + x[6] * dp[2] + x[7] * dp[3] + x[8] * dp[4] + x[9] * dp[5]) +
            (x[0] * cd[0] + x[1] * cd[1] + x[2] * cd[2] + x[3] * cd[3]) + Jpdd0 * dd;
      Jpy = (y[4] * dp[0] + y[5] * dp[1] + y[6] * dp[2] + y[7] * dp[3] + y[8] * dp[4] + y[9] * dp[5]) +
            (y[0] * cd[0] + y[1] * cd[1] + y[2] * cd[2] + y[3] * cd[3]) + Jpdd1 * dd;
    }
    float JI_r0 = 0, JI_r1 = 0, Jab_r0 = 0, Jab_r1 = 0, rr = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float jx = *jplane(B, buf, J_IDX + k, s), jy = *jplane(B, buf, J_IDX + 8 + k, s);
      const float ja = *jplane(B, buf, J_AB + k, s), jb = *jplane(B, buf, J_AB + 8 + k, s);
      float ra;
      if (mode == 0) ra = *jplane(B, buf, J_RESF + k, s);
      else {
        ra = B.s_rtz[(size_t)k * cR + s];
        if (mode == 1) { ra = ra + jx * Jpx; ra = ra + jy * Jpy; ra = ra + ja * dp[6]; ra = ra + jb * dp[7]; }
      }
      JI_r0 += ra * jx; JI_r1 += ra * jy; Jab_r0 += ra * ja; Jab_r1 += ra * jb; rr += ra * ra;
    }
    const float a = *jplane(B, buf, J_IDX2, s), b = *jplane(B, buf, J_IDX2 + 1, s), cc = *jplane(B, buf, J_IDX2 + 3, s);
    {  // AccumulatorApprox::update (MatrixAccumulators.h:714-784): 10x10 upper triangle
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 10; r++)
#pragma unroll
        for (int q = r; q < 10; q++) {
          const float v = a * x[q] * x[r] + cc * y[q] * y[r] + b * (x[q] * y[r] + y[q] * x[r]);
          if (idx < 32) g0[idx] = v; else g1[idx - 32] = v;
          idx++;
        }
    }
    {  // updateTopRight (:786-836): 10x3
      const float TR00 = *jplane(B, buf, J_ABIDX, s), TR10 = *jplane(B, buf, J_ABIDX + 1, s);
      const float TR01 = *jplane(B, buf, J_ABIDX + 2, s), TR11 = *jplane(B, buf, J_ABIDX + 3, s);
#pragma unroll
      for (int r = 0; r < 10; r++) {
        const float t0 = x[r] * TR00 + y[r] * TR10, t1 = x[r] * TR01 + y[r] * TR11, t2 = x[r] * JI_r0 + y[r] * JI_r1;
        const int i0 = 55 + 3 * r;  // global value index
        if (i0 < 64) g1[i0 - 32] = t0; else g2[i0 - 64] = t0;
        if (i0 + 1 < 64) g1[i0 + 1 - 32] = t1; else g2[i0 + 1 - 64] = t1;
        if (i0 + 2 < 64) g1[i0 + 2 - 32] = t2; else g2[i0 + 2 - 64] = t2;
      }
    }
    // updateBotRight (:838-852): a00,a01,a02,a11,a12,a22 -> value indices 85..90
    g2[21] = *jplane(B, buf, J_AB2, s); g2[22] = *jplane(B, buf, J_AB2 + 1, s); g2[23] = Jab_r0;
    g2[24] = *jplane(B, buf, J_AB2 + 3, s); g2[25] = Jab_r1; g2[26] = rr;
    g2[27] = 1.f;  // AccumulatorApprox::num
    // per-point tail terms (AccumulatedTopHessian.cpp:160-176)
    const float i10 = *jplane(B, buf, J_IDX2 + 2, s);
    const float Ji2_0 = a * Jpdd0 + b * Jpdd1, Ji2_1 = i10 * Jpdd0 + cc * Jpdd1;
    ps[0] = JI_r0 * Jpdd0 + JI_r1 * Jpdd1;
    ps[1] = Ji2_0 * Jpdd0 + Ji2_1 * Jpdd1;
#pragma unroll
    for (int k = 0; k < 4; k++) ps[2 + k] = x[k] * Ji2_0 + y[k] * Ji2_1;
  }
  if (s < ch.end) {
#pragma unroll
    for (int k = 0; k < 6; k++) B.s_psum[(size_t)k * cR + s] = ps[k];
  }
  const float r0 = warp_reduce_transpose32(g0, lane), r1 = warp_reduce_transpose32(g1, lane), r2 = warp_reduce_transpose32(g2, lane);
  wsum[warp][lane] = r0; wsum[warp][32 + lane] = r1; wsum[warp][64 + lane] = r2;
  __syncthreads();
  if (threadIdx.x < kTopVals) {
    float t = 0;
#pragma unroll
    for (int w = 0; w < kChunk / 32; w++) t += wsum[w][threadIdx.x];
    B.tpart[(size_t)blockIdx.x * kTopVals + threadIdx.x] = t;
  }
}

// fixed-order sum of the chunk partials of one key, expanded to the 13x13 block AccumulatorApprox::finish builds
__global__ void ba_top_finish_kernel(BAView B) {
  __shared__ float v[kTopVals];
  const int key = blockIdx.x;
  if (threadIdx.x < kTopVals) {
    float t = 0;
    for (int c = B.key_chunk_begin[key]; c < B.key_chunk_begin[key + 1]; c++) t += B.tpart[(size_t)c * kTopVals + threadIdx.x];
    v[threadIdx.x] = t;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 169; e += blockDim.x) {
    int r = e / 13, c = e % 13;
    if (r > c) { const int q = r; r = c; c = q; }
    float val;
    if (c < 10) val = v[r * 10 - (r * (r - 1)) / 2 + (c - r)];
    else if (r < 10) val = v[55 + 3 * r + (c - 10)];
    else { const int rr = r - 10, cc = c - 10; val = v[85 + (rr == 0 ? cc : (rr == 1 ? 2 + cc : 5))]; }
    B.Gf[(size_t)key * 169 + e] = val;
    B.G[(size_t)key * 169 + e] = (double)val;
  }
  if (threadIdx.x == 0) B.G[(size_t)B.n * B.n * 169 + key] = (double)v[91];  // acc.num of this key
}

// per-point tail: sums of its residuals' terms in residualsAll order
__global__ void ba_point_sums_kernel(BAView B, int mode) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B.P) return;
  if (mode == 2 && B.p_flag[p] != PS_MARGINALIZE) return;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = B.p_res_begin[p]; i < B.p_res_begin[p + 1]; i++) {
    const int s = B.p_res_list[i];
#pragma unroll
    for (int k = 0; k < 6; k++) acc[k] += B.s_psum[(size_t)k * B.capR + s];
  }
  const size_t cP = B.capP;
  // p_acc planes: 0 Hdd_A, 1 bd_A, 2-5 Hcd_A, 6 Hdd_L, 7 bd_L, 8-11 Hcd_L
  const int base = (mode == 0) ? 0 : 6;
  B.p_acc[(base + 0) * cP + p] = acc[1];
  B.p_acc[(base + 1) * cP + p] = acc[0];
#pragma unroll
  for (int k = 0; k < 4; k++) B.p_acc[(base + 2 + k) * cP + p] = acc[2 + k];
  if (mode == 2) {
#pragma unroll
    for (int k = 0; k < 6; k++) B.p_acc[k * cP + p] = 0;
  }
}

// ---- B6, per-point part ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ba_sc_point_kernel(BAView B, int shiftPriorToZero) {
  __shared__ float red[4][20];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t cP = B.capP;
  float v[20];
#pragma unroll
  for (int i = 0; i < 20; i++) v[i] = 0;
  if (p < B.P && (shiftPriorToZero || B.p_flag[p] == PS_MARGINALIZE)) {
    int ngood = 0;
    for (int i = B.p_res_begin[p]; i < B.p_res_begin[p + 1]; i++) if (B.s_flags[B.p_res_list[i]] & RF_ACTIVE) ngood++;
    if (ngood == 0) {
      B.p_acc[12 * cP + p] = 0; B.p_acc[13 * cP + p] = 0; B.p_acc[15 * cP + p] = 0;
    } else {
      const float priorF = B.p_priorF[p];
      float H = B.p_acc[0 * cP + p] + B.p_acc[6 * cP + p] + priorF;
      if (H < 1e-10) H = 1e-10;
      const float HdiF = (float)(1.0 / H);
      float bdSum = B.p_acc[1 * cP + p] + B.p_acc[7 * cP + p];
      if (shiftPriorToZero) bdSum += priorF * B.p_deltaF[p];
      B.p_acc[15 * cP + p] = H; B.p_acc[12 * cP + p] = HdiF; B.p_acc[13 * cP + p] = bdSum;
      float Hcd[4];
#pragma unroll
      for (int k = 0; k < 4; k++) Hcd[k] = B.p_acc[(2 + k) * cP + p] + B.p_acc[(8 + k) * cP + p];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) v[i * 4 + j] = (HdiF * Hcd[i]) * Hcd[j];
      const float w = bdSum * HdiF;
#pragma unroll
      for (int i = 0; i < 4; i++) v[16 + i] = w * Hcd[i];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 20; i++) { const float t = warp_sum(v[i]); if (lane == 0) red[warp][i] = t; }
  __syncthreads();
  if (threadIdx.x < 20) B.pblockpart[(size_t)blockIdx.x * 32 + threadIdx.x] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// ---- B6, O(res^2) part: grid = (chunks, n+1); blockIdx.y < n: accD towards target y; == n: accE + accEB -------
__global__ void __launch_bounds__(kChunk) ba_sc_pair_kernel(BAView B, int shiftPriorToZero) {
  __shared__ float wsum[kChunk / 32][64];
  const Chunk ch = B.chunks[blockIdx.x];
  const int t2 = blockIdx.y;
  const int s1 = ch.begin + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t cR = B.capR, cP = B.capP;
  float g0[32], g1[32];
#pragma unroll
  for (int i = 0; i < 32; i++) { g0[i] = 0; g1[i] = 0; }
  bool use = s1 < ch.end && (B.s_flags[s1] & RF_ACTIVE);
  int p = 0;
  if (use) { p = B.s_point[s1]; use = shiftPriorToZero || B.p_flag[p] == PS_MARGINALIZE; }
  int s2 = -1;
  if (use && t2 < B.n) { s2 = B.slot_of[(size_t)p * B.n + t2]; use = s2 >= 0 && (B.s_flags[s2] & RF_ACTIVE); }
  if (use) {
    const float HdiF = B.p_acc[12 * cP + p];
    float L[8];
#pragma unroll
    for (int i = 0; i < 8; i++) L[i] = HdiF * B.s_JpJd[(size_t)i * cR + s1];
    if (t2 < B.n) {
      float Rv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) Rv[i] = B.s_JpJd[(size_t)i * cR + s2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) { g0[i * 8 + j] = L[i] * Rv[j]; g1[i * 8 + j] = L[4 + i] * Rv[j]; }
    } else {
      float Hcd[4];
#pragma unroll
      for (int k = 0; k < 4; k++) Hcd[k] = B.p_acc[(2 + k) * cP + p] + B.p_acc[(8 + k) * cP + p];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) g0[i * 4 + j] = L[i] * Hcd[j];
      const float w = HdiF * B.p_acc[13 * cP + p];
#pragma unroll
      for (int i = 0; i < 8; i++) g1[i] = w * B.s_JpJd[(size_t)i * cR + s1];
      g1[8] = 1.f;
    }
    if (t2 < B.n) { /* count of updates, for AccumulatorXX::num */ }
  }
  float cnt = use ? 1.f : 0.f;
  cnt = warp_sum(cnt);
  const float r0 = warp_reduce_transpose32(g0, lane), r1 = warp_reduce_transpose32(g1, lane);
  wsum[warp][lane] = r0; wsum[warp][32 + lane] = r1;
  __shared__ float wcnt[kChunk / 32];
  if (lane == 0) wcnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0;
#pragma unroll
    for (int w = 0; w < kChunk / 32; w++) t += wsum[w][threadIdx.x];
    B.dpart[((size_t)blockIdx.x * (B.n + 1) + t2) * 65 + threadIdx.x] = t;
  }
  if (threadIdx.x == 64) {
    float t = 0;
#pragma unroll
    for (int w = 0; w < kChunk / 32; w++) t += wcnt[w];
    B.dpart[((size_t)blockIdx.x * (B.n + 1) + t2) * 65 + 64] = t;
  }
}

// grid = (n*n keys, n+1): D[key][t2] (64 + count), E/EB[key]; block (0,0) also sums accHcc / accbc
__global__ void ba_sc_finish_kernel(BAView B, int pblocks) {
  const int key = blockIdx.x, t2 = blockIdx.y, e = threadIdx.x;
  const int n = B.n;
  if (e < 65) {
    float t = 0;
    for (int c = B.key_chunk_begin[key]; c < B.key_chunk_begin[key + 1]; c++) t += B.dpart[((size_t)c * (n + 1) + t2) * 65 + e];
    if (t2 < n) B.D[((size_t)key * n + t2) * 65 + e] = (double)t;
    else if (e < 40) B.E[(size_t)key * 40 + e] = (double)t;
  }
  if (key == 0 && t2 == 0 && e < 20) {
    float t = 0;
    for (int b = 0; b < pblocks; b++) t += B.pblockpart[(size_t)b * 32 + e];
    B.Hcc[e] = (double)t;
  }
}

// ---- B5: stitch. One thread per element of the (4+8n)^2 matrix and of b -------------------------------------
__device__ __forceinline__ double quad88(const double* A, int i, const double* G, int goff_r, int goff_c, const double* Bm, int j) {
  // sum_pq A[i,p] * G[(goff_r+p)*13 + goff_c+q] * Bm[j,q]
  double s = 0;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    double in = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) in += G[(goff_r + p) * 13 + goff_c + q] * Bm[j * 8 + q];
    s += A[i * 8 + p] * in;
  }
  return s;
}

__global__ void ba_stitch_top_kernel(BAView B, double* H, double* bvec, int usePrior, const double* cPrior) {
  const int n = B.n, d = kCPARS + 8 * n;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d + d) return;
  const double* G = B.G;
  if (e >= d * d) {  // b
    const int r = e - d * d;
    double s = 0;
    if (r < 4) {
      for (int k = 0; k < n * n; k++) s += G[(size_t)k * 169 + r * 13 + 12];
      if (usePrior) s += cPrior[r] * (double)B.cDeltaF[r];
    } else {
      const int a = (r - 4) / 8, i = (r - 4) % 8;
      for (int o = 0; o < n; o++) {
        const int kh = a + o * n, kt = o + a * n;  // keys with host a / with target a
        const double* AH = B.adHost + (size_t)kh * 64; const double* AT = B.adTarget + (size_t)kt * 64;
        for (int q = 0; q < 8; q++) {
          s += AH[i * 8 + q] * G[(size_t)kh * 169 + (4 + q) * 13 + 12];
          s += AT[i * 8 + q] * G[(size_t)kt * 169 + (4 + q) * 13 + 12];
        }
      }
      if (usePrior) s += B.fprior[a * 24 + i] * B.fprior[a * 24 + 8 + i];
    }
    bvec[r] = s;
    return;
  }
  int r = e / d, c = e % d;
  double s = 0;
  if (r < 4 && c < 4) {
    for (int k = 0; k < n * n; k++) s += G[(size_t)k * 169 + r * 13 + c];
    if (usePrior && r == c) s += cPrior[r];
  } else if (r < 4 || c < 4) {
    if (r < 4) { const int q = r; r = c; c = q; }  // H[0:4, hIdx] = H[hIdx, 0:4]^T
    const int a = (r - 4) / 8, i = (r - 4) % 8;
    for (int o = 0; o < n; o++) {
      const int kh = a + o * n, kt = o + a * n;
      const double* AH = B.adHost + (size_t)kh * 64; const double* AT = B.adTarget + (size_t)kt * 64;
      for (int q = 0; q < 8; q++) {
        s += AH[i * 8 + q] * G[(size_t)kh * 169 + (4 + q) * 13 + c];
        s += AT[i * 8 + q] * G[(size_t)kt * 169 + (4 + q) * 13 + c];
      }
    }
  } else {
    const int a = (r - 4) / 8, i = (r - 4) % 8, b = (c - 4) / 8, j = (c - 4) % 8;
    if (a != b) {
      const int k1 = a + b * n, k2 = b + a * n;
      s = quad88(B.adHost + (size_t)k1 * 64, i, G + (size_t)k1 * 169, 4, 4, B.adTarget + (size_t)k1 * 64, j) +
          quad88(B.adHost + (size_t)k2 * 64, j, G + (size_t)k2 * 169, 4, 4, B.adTarget + (size_t)k2 * 64, i);
    } else {
      for (int o = 0; o < n; o++) {
        const int kh = a + o * n, kt = o + a * n;
        s += quad88(B.adHost + (size_t)kh * 64, i, G + (size_t)kh * 169, 4, 4, B.adHost + (size_t)kh * 64, j);
        s += quad88(B.adTarget + (size_t)kt * 64, i, G + (size_t)kt * 169, 4, 4, B.adTarget + (size_t)kt * 64, j);
      }
      const int kd = a + a * n;
      s += quad88(B.adHost + (size_t)kd * 64, i, G + (size_t)kd * 169, 4, 4, B.adTarget + (size_t)kd * 64, j);
      if (usePrior && i == j) s += B.fprior[a * 24 + i];
    }
  }
  H[e] = s;
}

// ---- B7: U = adHost[key] * D[key][k], V = adTarget[key] * D[key][k]; grid = n*n*n, 64 threads ---------------
__global__ void ba_sc_uv_kernel(BAView B) {
  const int blk = blockIdx.x;  // key * n + k
  const int key = blk / B.n;
  const int i = threadIdx.x / 8, j = threadIdx.x % 8;
  const double* Dm = B.D + (size_t)blk * 65;
  const double* AH = B.adHost + (size_t)key * 64; const double* AT = B.adTarget + (size_t)key * 64;
  double u = 0, v = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) { u += AH[i * 8 + q] * Dm[q * 8 + j]; v += AT[i * 8 + q] * Dm[q * 8 + j]; }
  B.U[(size_t)blk * 64 + threadIdx.x] = u;
  B.V[(size_t)blk * 64 + threadIdx.x] = v;
}

__device__ __forceinline__ double rowdot8(const double* X, int i, const double* Bm, int j) {
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += X[i * 8 + q] * Bm[j * 8 + q];
  return s;
}

__global__ void ba_stitch_sc_kernel(BAView B, double* H, double* bvec) {
  const int n = B.n, d = kCPARS + 8 * n, n2 = n * n;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d + d) return;
  if (e >= d * d) {
    const int r = e - d * d;
    double s = 0;
    if (r < 4) s = B.Hcc[16 + r];
    else {
      const int a = (r - 4) / 8, i = (r - 4) % 8;
      for (int o = 0; o < n; o++) {
        const int kh = a + o * n, kt = o + a * n;
        const double* AH = B.adHost + (size_t)kh * 64; const double* AT = B.adTarget + (size_t)kt * 64;
        for (int q = 0; q < 8; q++) { s += AH[i * 8 + q] * B.E[(size_t)kh * 40 + 32 + q]; s += AT[i * 8 + q] * B.E[(size_t)kt * 40 + 32 + q]; }
      }
    }
    bvec[r] = s;
    return;
  }
  int r = e / d, c = e % d;
  double s = 0;
  if (r < 4 && c < 4) s = B.Hcc[r * 4 + c];
  else if (r < 4 || c < 4) {
    if (r < 4) { const int q = r; r = c; c = q; }
    const int a = (r - 4) / 8, i = (r - 4) % 8;
    for (int o = 0; o < n; o++) {
      const int kh = a + o * n, kt = o + a * n;
      const double* AH = B.adHost + (size_t)kh * 64; const double* AT = B.adTarget + (size_t)kt * 64;
      for (int q = 0; q < 8; q++) { s += AH[i * 8 + q] * B.E[(size_t)kh * 40 + q * 4 + c]; s += AT[i * 8 + q] * B.E[(size_t)kt * 40 + q * 4 + c]; }
    }
  } else {
    const int a = (r - 4) / 8, i = (r - 4) % 8, b = (c - 4) / 8, j = (c - 4) % 8;
    // H[jIdx,kIdx] += AT_ij D_ijk AT_ik^T  : (j=a, k=b), all hosts o
    for (int o = 0; o < n; o++) s += rowdot8(B.V + ((size_t)(o + n * a) * n + b) * 64, i, B.adTarget + (size_t)(o + n * b) * 64, j);
    // H[jIdx,iIdx] += AT_ij D_ijk AH_ik^T  : (j=a, host=b), all k
    for (int k = 0; k < n; k++) s += rowdot8(B.V + ((size_t)(b + n * a) * n + k) * 64, i, B.adHost + (size_t)(b + n * k) * 64, j);
    // H[iIdx,kIdx] += AH_ij D_ijk AT_ik^T  : (host=a, k=b), all j
    for (int o = 0; o < n; o++) s += rowdot8(B.U + ((size_t)(a + n * o) * n + b) * 64, i, B.adTarget + (size_t)(a + n * b) * 64, j);
    if (a == b)  // H[iIdx,iIdx] += AH_ij D_ijk AH_ik^T : all j, k
      for (int o = 0; o < n; o++)
        for (int k = 0; k < n; k++) s += rowdot8(B.U + ((size_t)(a + n * o) * n + k) * 64, i, B.adHost + (size_t)(a + n * k) * 64, j);
    (void)n2;
  }
  H[e] = s;
}

// ---- B8: assemble (elementwise) then scaled diagonal-pivoted LDLT (+ orthogonalisation) in ONE CTA ----------------
struct SolveParams {
  int n, d, iteration, have_M;
  double lambda, solverModeDelta;
  const double* HA; const double* bA; const double* HL; const double* bL; const double* Hsc; const double* bsc;
  const double* HM; const double* bM;
  const double* fprior; const float* cDeltaF;
  const double* N;   // d x 7 nullspace columns (row-major) or null
  double* HF; double* bF; double* x;
};

// HFinal = HL + HM + HA, diag *= (1 + lambda), -= H_sc / (1 + lambda); bFinal = bL + (bM + HM delta) + bA - b_sc
// (EnergyFunctional.cpp:869-918). Linear in the per-point contributions, so a rank that holds a shard of the points
// produces a PARTIAL (HFinal, bFinal) here (have_M and the priors only on one rank) and the shards are summed by one
// allreduce before ba_solve_kernel.
__global__ void ba_assemble_kernel(SolveParams S) {
  const int d = S.d;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d + d) return;
  const double lam = S.lambda;
  if (e >= d * d) {
    const int i = e - d * d;
    double bm = 0;
    if (S.have_M) {
      double s = 0;
      for (int c = 0; c < d; c++) {
        const double dl = c < 4 ? (double)S.cDeltaF[c] : S.fprior[((c - 4) / 8) * 24 + 16 + (c - 4) % 8];
        s += S.HM[(size_t)i * d + c] * dl;
      }
      bm = S.bM[i] + s;
    }
    S.bF[i] = S.bL[i] + bm + S.bA[i] - S.bsc[i];
    return;
  }
  const int r = e / d, c = e % d;
  const double f = (double)(1.0f) / (1 + lam);
  double v = S.HL[e] + (S.have_M ? S.HM[e] : 0.0) + S.HA[e];
  if (r == c) v *= (1 + lam);
  v -= S.Hsc[e] * f;
  S.HF[e] = v;
}

// x -= N (N^T N)^+ N^T x  (orthogonalize(&x, 0), EnergyFunctional.cpp:775-835), run by one CTA; scratch in shared memory
__device__ void ortho_vec(const double* __restrict__ Nraw, int d, int m, double delta, double* x, double* sN /* d*m */, double* sw /* >= 3*m*m+4*m */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double* G = sw; double* Vv = sw + m * m; double* wv = Vv + m * m; double* coef = wv + m; double* nrm = coef + m;
  if (tid < m) { double s = 0; for (int r = 0; r < d; r++) s += Nraw[r * m + tid] * Nraw[r * m + tid]; nrm[tid] = sqrt(s); }
  __syncthreads();
  for (int e = tid; e < d * m; e += nt) sN[e] = Nraw[e] / nrm[e % m];
  __syncthreads();
  if (tid < m * m) { const int i = tid / m, j = tid % m; double s = 0; for (int r = 0; r < d; r++) s += sN[r * m + i] * sN[r * m + j]; G[tid] = s; Vv[tid] = (i == j) ? 1.0 : 0.0; }
  __syncthreads();
  if (tid == 0) {  // cyclic Jacobi on the m x m Gram matrix (m = 7)
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0;
      for (int i = 0; i < m; i++) for (int j = i + 1; j < m; j++) off += G[i * m + j] * G[i * m + j];
      if (off < 1e-300) break;
      for (int p = 0; p < m; p++) for (int q = p + 1; q < m; q++) {
        if (fabs(G[p * m + q]) < 1e-300) continue;
        const double theta = (G[q * m + q] - G[p * m + p]) / (2 * G[p * m + q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < m; k++) { const double a = G[k * m + p], b = G[k * m + q]; G[k * m + p] = c * a - s * b; G[k * m + q] = s * a + c * b; }
        for (int k = 0; k < m; k++) { const double a = G[p * m + k], b = G[q * m + k]; G[p * m + k] = c * a - s * b; G[q * m + k] = s * a + c * b; }
        for (int k = 0; k < m; k++) { const double a = Vv[k * m + p], b = Vv[k * m + q]; Vv[k * m + p] = c * a - s * b; Vv[k * m + q] = s * a + c * b; }
      }
    }
    double mx = 0;
    for (int i = 0; i < m; i++) { wv[i] = sqrt(fmax(G[i * m + i], 0.0)); mx = fmax(mx, wv[i]); }
    for (int i = 0; i < m; i++) if (!(wv[i] > delta * mx)) wv[i] = 0;  // dropped singular values
  }
  __syncthreads();
  // coef_i = (U_i . x) / sigma_i with U_i = N v_i / sigma_i  ->  x -= sum_i (N v_i) * (v_i^T N^T x) / sigma_i^2
  if (tid < m) {
    double s = 0;
    if (wv[tid] > 0) {
      for (int r = 0; r < d; r++) { double nv = 0; for (int j = 0; j < m; j++) nv += sN[r * m + j] * Vv[j * m + tid]; s += nv * x[r]; }
      s /= (wv[tid] * wv[tid]);
    }
    coef[tid] = s;
  }
  __syncthreads();
  for (int r = tid; r < d; r += nt) {
    double sub = 0;
    for (int i = 0; i < m; i++) { if (coef[i] == 0) continue; double nv = 0; for (int j = 0; j < m; j++) nv += sN[r * m + j] * Vv[j * m + i]; sub += nv * coef[i]; }
    x[r] -= sub;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) ba_solve_kernel(SolveParams S) {
  extern __shared__ double sm[];
  const int d = S.d, tid = threadIdx.x, nt = blockDim.x;
  double* M = sm;                 // d*d
  double* bs = M + d * d;         // d
  double* sv = bs + d;            // d  SVecI
  double* dg = sv + d;            // d  D of LDLT
  double* y = dg + d;             // d
  double* delta = y + d;          // d
  double* scr = delta + d;        // d*7 + 256
  __shared__ int perm[kCPARS + 8 * kMaxFrames];
  __shared__ int piv;
  for (int i = tid; i < d; i += nt) bs[i] = S.bF[i];
  for (int e = tid; e < d * d; e += nt) M[e] = S.HF[e];
  __syncthreads();
  for (int i = tid; i < d; i += nt) { sv[i] = 1.0 / sqrt(M[i * d + i] + 10); perm[i] = i; }
  __syncthreads();
  for (int e = tid; e < d * d; e += nt) M[e] = sv[e / d] * M[e] * sv[e % d];
  for (int i = tid; i < d; i += nt) bs[i] = sv[i] * bs[i];
  __syncthreads();
  // diagonal-pivoted LDLT (the strategy of Eigen::LDLT, which EnergyFunctional.cpp:976 calls)
  for (int k = 0; k < d; k++) {
    if (tid == 0) {
      int p = k; double best = fabs(M[k * d + k]);
      for (int i = k + 1; i < d; i++) { const double v = fabs(M[i * d + i]); if (v > best) { best = v; p = i; } }
      piv = p;
      if (p != k) { const int q = perm[k]; perm[k] = perm[p]; perm[p] = q; }
    }
    __syncthreads();
    const int p = piv;
    if (p != k) {
      for (int j = tid; j < d; j += nt) { const double a = M[k * d + j]; M[k * d + j] = M[p * d + j]; M[p * d + j] = a; }
      __syncthreads();
      for (int j = tid; j < d; j += nt) { const double a = M[j * d + k]; M[j * d + k] = M[j * d + p]; M[j * d + p] = a; }
      __syncthreads();
    }
    const double dk = M[k * d + k];
    __syncthreads();
    if (tid == 0) dg[k] = dk;
    if (dk == 0.0 || !isfinite(dk)) { for (int i = k + 1 + tid; i < d; i += nt) M[i * d + k] = 0; __syncthreads(); continue; }
    for (int i = k + 1 + tid; i < d; i += nt) M[i * d + k] /= dk;
    __syncthreads();
    const int rem = d - k - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) M[i * d + j] -= M[i * d + k] * dk * M[j * d + k];
    }
    __syncthreads();
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j > i) M[i * d + j] = M[j * d + i];
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i < d; i++) y[i] = bs[perm[i]];
    for (int i = 0; i < d; i++) { double v = y[i]; for (int j = 0; j < i; j++) v -= M[i * d + j] * y[j]; y[i] = v; }
    for (int i = 0; i < d; i++) y[i] = (dg[i] != 0.0) ? y[i] / dg[i] : 0.0;
    for (int i = d - 1; i >= 0; i--) { double v = y[i]; for (int j = i + 1; j < d; j++) v -= M[j * d + i] * y[j]; y[i] = v; }
    for (int i = 0; i < d; i++) bs[perm[i]] = y[i];
  }
  __syncthreads();
  for (int i = tid; i < d; i += nt) bs[i] *= sv[i];
  __syncthreads();
  if (S.iteration >= 2 && S.N) ortho_vec(S.N, d, 7, S.solverModeDelta, bs, M, scr);  // SOLVER_ORTHOGONALIZE_X_LATER (:980-984)
  for (int i = tid; i < d; i += nt) S.x[i] = bs[i];
}

// ---- B9 -------------------------------------------------------------------------------------------------------
// xAd[h*n + t] = x_h^T adHostF[h + t*n] + x_t^T adTargetF[h + t*n]   (EnergyFunctional.cpp:283-293)
__global__ void ba_xad_kernel(BAView B, const double* x, float* xAd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = B.n;
  if (e >= n * n * 8) return;
  const int j = e % 8, ht = e / 8, h = ht / n, t = ht % n;
  const size_t ad = ((size_t)h + (size_t)n * t) * 64;
  float a = 0, b = 0;
  for (int i = 0; i < 8; i++) a += (float)x[kCPARS + 8 * h + i] * B.adHostF[ad + i * 8 + j];
  for (int i = 0; i < 8; i++) b += (float)x[kCPARS + 8 * t + i] * B.adTargetF[ad + i * 8 + j];
  xAd[e] = a + b;
}

__global__ void ba_resub_kernel(BAView B, const double* x, const float* xAd) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B.P) return;
  const size_t cP = B.capP, cR = B.capR;
  int ngood = 0;
  for (int i = B.p_res_begin[p]; i < B.p_res_begin[p + 1]; i++) if (B.s_flags[B.p_res_list[i]] & RF_ACTIVE) ngood++;
  if (ngood == 0) { B.p_acc[14 * cP + p] = 0; return; }
  float b = B.p_acc[13 * cP + p];
  float dot = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) dot += (float)x[k] * (B.p_acc[(2 + k) * cP + p] + B.p_acc[(8 + k) * cP + p]);
  b -= dot;
  const int n = B.n;
  for (int i = B.p_res_begin[p]; i < B.p_res_begin[p + 1]; i++) {
    const int s = B.p_res_list[i];
    if (!(B.s_flags[s] & RF_ACTIVE)) continue;
    const int key = B.s_key[s];
    const float* xa = xAd + ((size_t)(key % n) * n + key / n) * 8;
    float sacc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sacc += xa[k] * B.s_JpJd[(size_t)k * cR + s];
    b -= sacc;
  }
  B.p_acc[14 * cP + p] = -b * B.p_acc[12 * cP + p];
}

}  // namespace sdso
