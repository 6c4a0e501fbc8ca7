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
  const BACalib* cp;   // device-resident calibration (moves inside the LM loop)
  const int* done;     // OptDev::done: kernels of the LM iteration chain exit at once when set (nullptr: never)
  const float4* const* tex0; const float* frameTH; const PrecalcDev* precalc;
  const double* adHost; const double* adTarget; const float* adHostF; const float* adTargetF;
  const float* adHTdeltaF; const float* cDeltaF; const double* fprior;
  const int* p_host; const float* p_u; const float* p_v; float* p_idepth; float* p_idepth_zero; float* p_idepth_backup;
  const float4* p_color; const float4* p_weights; const float* p_priorF; float* p_deltaF;
  const int* p_res_begin; const int* p_res_list; const int* slot_of; float* p_acc; const unsigned char* p_flag;
  const int* s_point; const int* s_key;
  unsigned char* s_state; unsigned char* s_newstate; unsigned char* s_flags; unsigned char* s_sel;
  float* s_energy; float* J; float* s_rtz; float* s_JpJd; float* s_center; float* s_psum;
  const Chunk* chunks; int nchunks; const int* key_chunk_begin;
  float* tpart; float* dpart; float* pblockpart;
  double* G; float* Gf; double* D; double* E; double* Hcc; double* U; double* V;
  double* energy_part; double* scalars; unsigned int* counter;
};
#define BA_EXIT_IF_DONE(B) do { if ((B).done && *(B).done) return; } while (0)

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
  if (B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED)) return 0.0;  // activeResiduals = residuals of the graph that are not linearised (FullSystemOptimize.cpp:900-902)
  B.s_energy[2 * cR + s] = -1;
  const float state_energy = B.s_energy[s];
  if (B.s_state[s] == RS_OOB) { B.s_newstate[s] = RS_OOB; return state_energy; }
  const int pidx = B.s_point[s];
  const int key = B.s_key[s];
  const int h = key % B.n, t = key / B.n;
  const PrecalcDev& pc = B.precalc[h * B.n + t];
  const BACalib& c = *B.cp;
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
  BA_EXIT_IF_DONE(B);
  __shared__ double red[32];
  __shared__ bool last;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0;
  if (s < B.R) {
    const bool active_res = !(B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED));
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
  BA_EXIT_IF_DONE(B);  // applyRes_Reductor over activeResiduals
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < B.R && !(B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED))) apply_res_slot(B, s, copyJ != 0);
}

// ---- B12: pieces of FullSystem::optimize (FullSystemOptimize.cpp:870-1042) that touch per-point / per-residual state ------
// resetOOB for the active residuals (:888-893, Residuals.h:107-115)
__global__ void ba_reset_oob_kernel(BAView B) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.R || (B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED))) return;
  B.s_state[s] = RS_IN; B.s_newstate[s] = RS_OUTLIER;
  B.s_energy[s] = 0; B.s_energy[(size_t)B.capR + s] = 0;
}
// backupState (:309-350): idepth_backup = idepth
__global__ void ba_backup_points_kernel(BAView B) {
  BA_EXIT_IF_DONE(B);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < B.P) B.p_idepth_backup[p] = B.p_idepth[p];
}
// doStepFromBackup (:268-279) / loadSateBackup (:355-362): idepth = idepth_zero = idepth_backup + stepfacD * step (points carry no FEJ
// point), plus the sums the convergence test needs: out[0] += step^2, out[1] += |idepth_backup| (fixed-order double sums)
__global__ void __launch_bounds__(128) ba_step_points_kernel(BAView B, float stepfacD, double* part /* [blocks][2] */) {
  BA_EXIT_IF_DONE(B);
  __shared__ double red[32];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double s2 = 0, sn = 0;
  if (p < B.P) {
    const float bak = B.p_idepth_backup[p], st = B.p_acc[14 * (size_t)B.capP + p];
    const float nv = bak + stepfacD * st;
    B.p_idepth[p] = nv; B.p_idepth_zero[p] = nv; B.p_deltaF[p] = nv - nv;
    s2 = (double)(st * st); sn = (double)fabsf(bak);
  }
  const double a = block_sum_d(s2, red);
  const double b = block_sum_d(sn, red);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
}
__global__ void ba_sum_pairs_kernel(const double* part, int nblocks, double* out, const int* done = nullptr) {
  if (done && *done) return;
  if (threadIdx.x < 2) { double s = 0; for (int i = 0; i < nblocks; i++) s += part[2 * i + threadIdx.x]; out[threadIdx.x] = s; }
}
// number of residuals accumulateAF would count (ef->resInA): active, not linearised
__global__ void ba_count_active_kernel(BAView B, unsigned int* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool a = s < B.R && (B.s_flags[s] & RF_ACTIVE) && !(B.s_flags[s] & RF_LINEARIZED);
  const unsigned m = __ballot_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned)__popc(m));
}
// residuals that the final linearizeAll(true) found inactive leave the graph (toRemove, :170-200)
__global__ void ba_drop_inactive_kernel(BAView B) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.R) return;
  const unsigned char fl = B.s_flags[s];
  if (!(fl & (RF_LINEARIZED | RF_DROPPED)) && !(fl & RF_ACTIVE)) B.s_flags[s] = fl | RF_DROPPED;
}

// setNewFrameEnergyTH (:98-139) on ONE CTA: exact selection of the floor(0.7 N)-th smallest state_NewEnergyWithOutlier among the
// active residuals that target the newest frame — a 4-pass radix select on the (non-negative) float bit patterns, so the result
// is the very element std::nth_element would deliver — then the threshold arithmetic in float as written.
__global__ void __launch_bounds__(1024) ba_energy_th_kernel(BAView B, int newest, float thN, float facMedian, float constWeight, float overallW,
                                                            float* frameTH, float* out) {
  BA_EXIT_IF_DONE(B);
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_mask, s_k, s_n;
  const int tid = threadIdx.x, n = B.n;
  const float* NE = B.s_energy + 2 * (size_t)B.capR;
  if (tid == 0) { s_n = 0; s_prefix = 0; s_mask = 0; }
  __syncthreads();
  unsigned cnt = 0;
  for (int s = tid; s < B.R; s += blockDim.x) {
    const bool sel = !(B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED)) && (B.s_key[s] / n) == newest && NE[s] >= 0;
    cnt += sel ? 1u : 0u;
  }
  atomicAdd(&s_n, cnt);
  __syncthreads();
  const unsigned N = s_n;
  if (N == 0) { if (tid == 0) { const float th = 12 * 12 * 8; frameTH[newest] = th; out[0] = th; } return; }
  if (tid == 0) s_k = (unsigned)(int)(thN * N);
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, mask = s_mask;
    for (int s = tid; s < B.R; s += blockDim.x) {
      const bool sel = !(B.s_flags[s] & (RF_LINEARIZED | RF_DROPPED)) && (B.s_key[s] / n) == newest && NE[s] >= 0;
      if (!sel) continue;
      const unsigned bits = __float_as_uint(NE[s] + 0.0f);  // +0.0f folds -0 into +0
      if ((bits & mask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned k = s_k, d = 0;
      for (; d < 256; d++) { if (k < hist[d]) break; k -= hist[d]; }
      s_k = k; s_prefix = prefix | (d << shift); s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float nthElement = sqrtf(__uint_as_float(s_prefix));
    float th = nthElement * facMedian;
    th = 26.0f * constWeight + th * (1 - constWeight);
    th = th * th;
    th *= overallW * overallW;
    frameTH[newest] = th; out[0] = th;
  }
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
  BA_EXIT_IF_DONE(B);
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

// Fixed-order sum of the chunk partials of one key, expanded to the 13x13 block AccumulatorApprox::finish builds, then the
// first half of the adjoint stitch for this key: W = adHost*A88, Z = adTarget*A88 (8x8) and Wc = adHost*[A8C | b8],
// Zc = adTarget*[A8C | b8] (8x5), all in double (AccumulatedTopHessian.cpp:299-322). grid = n*n keys, 256 threads.
__global__ void __launch_bounds__(256) ba_top_finish_kernel(BAView B, double* W, double* Z, double* Wc, double* Zc) {
  BA_EXIT_IF_DONE(B);
  __shared__ float v[kTopVals];
  __shared__ double Gs[169];
  const int key = blockIdx.x, tid = threadIdx.x;
  if (tid < kTopVals) {
    float t = 0;
    for (int c = B.key_chunk_begin[key]; c < B.key_chunk_begin[key + 1]; c++) t += B.tpart[(size_t)c * kTopVals + tid];
    v[tid] = t;
  }
  __syncthreads();
  if (tid < 169) {
    int r = tid / 13, c = tid % 13;
    if (r > c) { const int q = r; r = c; c = q; }
    float val;
    if (c < 10) val = v[r * 10 - (r * (r - 1)) / 2 + (c - r)];
    else if (r < 10) val = v[55 + 3 * r + (c - 10)];
    else { const int rr = r - 10, cc = c - 10; val = v[85 + (rr == 0 ? cc : (rr == 1 ? 2 + cc : 5))]; }
    B.Gf[(size_t)key * 169 + tid] = val;
    B.G[(size_t)key * 169 + tid] = (double)val;
    Gs[tid] = (double)val;
  }
  if (tid == 0) B.G[(size_t)B.n * B.n * 169 + key] = (double)v[91];  // acc.num of this key
  __syncthreads();
  const double* AH = B.adHost + (size_t)key * 64; const double* AT = B.adTarget + (size_t)key * 64;
  if (tid < 128) {
    const int e = tid & 63, i = e >> 3, j = e & 7;
    const double* A = (tid < 64) ? AH : AT;
    double sacc = 0;
#pragma unroll
    for (int p = 0; p < 8; p++) sacc += A[i * 8 + p] * Gs[(4 + p) * 13 + 4 + j];
    ((tid < 64) ? W : Z)[(size_t)key * 64 + e] = sacc;
  } else if (tid < 128 + 80) {
    const int q = tid - 128, e = q % 40, i = e / 5, c = e % 5;
    const double* A = (q < 40) ? AH : AT;
    const int col = (c < 4) ? c : 12;
    double sacc = 0;
#pragma unroll
    for (int p = 0; p < 8; p++) sacc += A[i * 8 + p] * Gs[(4 + p) * 13 + col];
    ((q < 40) ? Wc : Zc)[(size_t)key * 40 + e] = sacc;
  }
}

// per-point tail: sums of its residuals' terms in residualsAll order
__global__ void ba_point_sums_kernel(BAView B, int mode) {
  BA_EXIT_IF_DONE(B);
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
  BA_EXIT_IF_DONE(B);
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
  BA_EXIT_IF_DONE(B);
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

// grid = (n*n keys, n+1), 96 threads. Fixed-order sums of the chunk partials: D[key][t2] (64 + count), E/EB[key]; then the
// first half of the adjoint stitch: U = adHost[key]*D, V = adTarget[key]*D (t2 < n) or Uc/Vc = ad*[E | EB] (t2 == n).
// Block (0,0) also sums accHcc / accbc (AccumulatedSCHessian.cpp:106-195).
__global__ void __launch_bounds__(96) ba_sc_finish_kernel(BAView B, int pblocks, double* Uc, double* Vc) {
  BA_EXIT_IF_DONE(B);
  __shared__ double Ds[65];
  const int key = blockIdx.x, t2 = blockIdx.y, e = threadIdx.x;
  const int n = B.n;
  if (e < 65) {
    float t = 0;
    for (int c = B.key_chunk_begin[key]; c < B.key_chunk_begin[key + 1]; c++) t += B.dpart[((size_t)c * (n + 1) + t2) * 65 + e];
    Ds[e] = (double)t;
    if (t2 < n) B.D[((size_t)key * n + t2) * 65 + e] = (double)t;
    else if (e < 40) B.E[(size_t)key * 40 + e] = (double)t;
  }
  if (key == 0 && t2 == 0 && e >= 65 && e < 85) {
    float t = 0;
    for (int b = 0; b < pblocks; b++) t += B.pblockpart[(size_t)b * 32 + (e - 65)];
    B.Hcc[e - 65] = (double)t;
  }
  __syncthreads();
  const double* AH = B.adHost + (size_t)key * 64; const double* AT = B.adTarget + (size_t)key * 64;
  if (t2 < n) {
    if (e < 64) {
      const int i = e >> 3, j = e & 7;
      double u = 0, v = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { u += AH[i * 8 + q] * Ds[q * 8 + j]; v += AT[i * 8 + q] * Ds[q * 8 + j]; }
      B.U[((size_t)key * n + t2) * 64 + e] = u;
      B.V[((size_t)key * n + t2) * 64 + e] = v;
    }
  } else if (e < 40) {
    const int i = e / 5, c = e % 5;  // c < 4: E column, c == 4: EB
    double u = 0, v = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { const double x = (c < 4) ? Ds[q * 4 + c] : Ds[32 + q]; u += AH[i * 8 + q] * x; v += AT[i * 8 + q] * x; }
    Uc[(size_t)key * 40 + e] = u;
    Vc[(size_t)key * 40 + e] = v;
  }
}

__device__ __forceinline__ double rowdot8(const double* __restrict__ X, int i, const double* __restrict__ Bm, int j) {
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += X[i * 8 + q] * Bm[j * 8 + q];
  return s;
}

// ---- B5: second half of the top stitch + symmetrisation. grid = n*n 8x8 frame blocks + 1 (calibration rows/cols and b),
// 256 threads = 64 elements x 4 term groups; the groups are summed in fixed order ----------------------------------------
__global__ void __launch_bounds__(256) ba_stitch_top_kernel(BAView B, const double* W, const double* Z, const double* Wc, const double* Zc,
                                                            double* H, double* bvec, int usePrior, const double* cPrior) {
  BA_EXIT_IF_DONE(B);
  __shared__ double red[4][64];
  const int n = B.n, d = kCPARS + 8 * n, tid = threadIdx.x;
  if ((int)blockIdx.x < n * n) {
    const int a = blockIdx.x % n, b = blockIdx.x / n;
    const int e = tid & 63, g = tid >> 6, i = e >> 3, j = e & 7;
    double s = 0;
    if (a != b) {  // M[a,b] + M[b,a]^T with M[h,t] = AH A88 AT^T of key (h,t)
      const int k1 = a + b * n, k2 = b + a * n;
      if (g == 0) s = rowdot8(W + (size_t)k1 * 64, i, B.adTarget + (size_t)k1 * 64, j);
      if (g == 1) s = rowdot8(W + (size_t)k2 * 64, j, B.adTarget + (size_t)k2 * 64, i);
    } else {
      for (int t = g; t < 2 * n + 1; t += 4) {
        if (t < n) { const int kh = a + t * n; s += rowdot8(W + (size_t)kh * 64, i, B.adHost + (size_t)kh * 64, j); }
        else if (t < 2 * n) { const int kt = (t - n) + a * n; s += rowdot8(Z + (size_t)kt * 64, i, B.adTarget + (size_t)kt * 64, j); }
        else { const int kd = a + a * n; s += rowdot8(W + (size_t)kd * 64, i, B.adTarget + (size_t)kd * 64, j); }
      }
    }
    red[g][e] = s;
    __syncthreads();
    if (g == 0) {
      double t = ((red[0][e] + red[1][e]) + red[2][e]) + red[3][e];
      if (usePrior && a == b && i == j) t += B.fprior[a * 24 + i];
      H[(size_t)(4 + 8 * a + i) * d + 4 + 8 * b + j] = t;
    }
    return;
  }
  // calibration block, calibration rows/columns, and b
  for (int idx = tid; idx < 8 * n * 5; idx += blockDim.x) {
    const int r = idx / 5, c = idx % 5, a = r / 8, i = r % 8;
    double s = 0;
    for (int o = 0; o < n; o++) s += Wc[(size_t)(a + o * n) * 40 + i * 5 + c] + Zc[(size_t)(o + a * n) * 40 + i * 5 + c];
    if (c < 4) { H[(size_t)(4 + r) * d + c] = s; H[(size_t)c * d + 4 + r] = s; }
    else { if (usePrior) s += B.fprior[a * 24 + i] * B.fprior[a * 24 + 8 + i]; bvec[4 + r] = s; }
  }
  for (int idx = tid; idx < 20; idx += blockDim.x) {
    const int r = idx / 5, c = idx % 5;
    double s = 0;
    for (int k = 0; k < n * n; k++) s += B.G[(size_t)k * 169 + r * 13 + (c < 4 ? c : 12)];
    if (c < 4) { if (usePrior && r == c) s += cPrior[r]; H[(size_t)r * d + c] = s; }
    else { if (usePrior) s += cPrior[r] * (double)B.cDeltaF[r]; bvec[r] = s; }
  }
}

// accumulateLF_MT when NO residual of the window is linearised — the normal case inside FullSystem::optimize: residuals are only
// linearised (fixLinearizationF) on their way into the marginalisation prior. The linearised top system is then just the priors
// (AccumulatedTopHessian.cpp:324-336) and every per-point L accumulator is zero; one small kernel instead of four.
__global__ void ba_prior_system_kernel(BAView B, double* H, double* bvec, int usePrior, const double* cPrior) {
  BA_EXIT_IF_DONE(B);
  const int n = B.n, d = kCPARS + 8 * n;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < d * d) {
    const int r = e / d, c = e % d;
    double v = 0;
    if (usePrior && r == c) v = r < 4 ? cPrior[r] : B.fprior[((r - 4) / 8) * 24 + (r - 4) % 8];
    H[e] = v;
  } else if (e < d * d + d) {
    const int r = e - d * d;
    double v = 0;
    if (usePrior) v = r < 4 ? cPrior[r] * (double)B.cDeltaF[r] : B.fprior[((r - 4) / 8) * 24 + (r - 4) % 8] * B.fprior[((r - 4) / 8) * 24 + 8 + (r - 4) % 8];
    bvec[r] = v;
  }
  // p_acc planes 6-11: Hdd_L, bd_L, Hcd_L
  for (int p = e; p < B.P; p += gridDim.x * blockDim.x)
#pragma unroll
    for (int k = 6; k < 12; k++) B.p_acc[(size_t)k * B.capP + p] = 0.f;
}

// ---- B7: second half of the Schur stitch, same decomposition --------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_stitch_sc_kernel(BAView B, const double* Uc, const double* Vc, double* H, double* bvec) {
  BA_EXIT_IF_DONE(B);
  __shared__ double red[4][64];
  const int n = B.n, d = kCPARS + 8 * n, tid = threadIdx.x;
  if ((int)blockIdx.x < n * n) {
    const int a = blockIdx.x % n, b = blockIdx.x / n;
    const int e = tid & 63, g = tid >> 6, i = e >> 3, j = e & 7;
    const int nterms = 3 * n + (a == b ? n * n : 0);
    double s = 0;
    for (int t = g; t < nterms; t += 4) {
      if (t < n) {             // H[jIdx,kIdx] += AT_ij D_ijk AT_ik^T : (j=a, k=b), host o
        const int o = t;
        s += rowdot8(B.V + ((size_t)(o + n * a) * n + b) * 64, i, B.adTarget + (size_t)(o + n * b) * 64, j);
      } else if (t < 2 * n) {  // H[jIdx,iIdx] += AT_ij D_ijk AH_ik^T : (j=a, host=b), target k
        const int k = t - n;
        s += rowdot8(B.V + ((size_t)(b + n * a) * n + k) * 64, i, B.adHost + (size_t)(b + n * k) * 64, j);
      } else if (t < 3 * n) {  // H[iIdx,kIdx] += AH_ij D_ijk AT_ik^T : (host=a, k=b), target o
        const int o = t - 2 * n;
        s += rowdot8(B.U + ((size_t)(a + n * o) * n + b) * 64, i, B.adTarget + (size_t)(a + n * b) * 64, j);
      } else {                 // H[iIdx,iIdx] += AH_ij D_ijk AH_ik^T : all (j,k)
        const int q = t - 3 * n, o = q / n, k = q % n;
        s += rowdot8(B.U + ((size_t)(a + n * o) * n + k) * 64, i, B.adHost + (size_t)(a + n * k) * 64, j);
      }
    }
    red[g][e] = s;
    __syncthreads();
    if (g == 0) H[(size_t)(4 + 8 * a + i) * d + 4 + 8 * b + j] = ((red[0][e] + red[1][e]) + red[2][e]) + red[3][e];
    return;
  }
  for (int idx = tid; idx < 8 * n * 5; idx += blockDim.x) {
    const int r = idx / 5, c = idx % 5, a = r / 8, i = r % 8;
    double s = 0;
    for (int o = 0; o < n; o++) s += Uc[(size_t)(a + o * n) * 40 + i * 5 + c] + Vc[(size_t)(o + a * n) * 40 + i * 5 + c];
    if (c < 4) { H[(size_t)(4 + r) * d + c] = s; H[(size_t)c * d + 4 + r] = s; }
    else bvec[4 + r] = s;
  }
  for (int idx = tid; idx < 20; idx += blockDim.x) {
    if (idx < 16) H[(size_t)(idx / 4) * d + idx % 4] = B.Hcc[idx];
    else bvec[idx - 16] = B.Hcc[idx];
  }
}

// ---- B8: assemble (elementwise) then scaled diagonal-pivoted LDLT (+ orthogonalisation) in ONE CTA ----------------
struct SolveParams {
  int n, d, iteration, have_M;
  double lambda, solverModeDelta;
  const double* HA; const double* bA; const double* HL; const double* bL; const double* Hsc; const double* bsc;
  const double* HM; const double* bM;
  const double* fprior; const float* cDeltaF;
  const double* N;   // d x 7 (row-major): orthonormal basis of the gauge nullspace in the first nrank columns, or null
  int nrank;
  int plain;         // 1: solve HF x = bF as is (g2o LinearSolver semantics): no (diag+10)^-1/2 scaling, no orthogonalisation
  double* HF; double* bF; double* x;
  const int* done;   // OptDev::done (nullptr: never)
  const int* it_ptr; // OptDev::it: iteration index read on the device (nullptr: `iteration` above)
};

// HFinal = HL + HM + HA, diag *= (1 + lambda), -= H_sc / (1 + lambda); bFinal = bL + (bM + HM delta) + bA - b_sc
// (EnergyFunctional.cpp:869-918). Linear in the per-point contributions, so a rank that holds a shard of the points
// produces a PARTIAL (HFinal, bFinal) here (have_M and the priors only on one rank) and the shards are summed by one
// allreduce before ba_solve_kernel.
__global__ void ba_assemble_kernel(SolveParams S) {
  if (S.done && *S.done) return;
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

// x -= Q Q^T x with Q (d x rank, orthonormal columns) = the left singular vectors of the 7 normalised nullspace vectors whose
// singular value exceeds solverModeDelta * max — i.e. orthogonalize(&x, 0) (EnergyFunctional.cpp:775-835). The basis depends
// only on the evaluation points, so it is computed once per window on the host (sdso_ba_prepare) instead of per solve.
__device__ void ortho_vec(const double* __restrict__ Q, int d, int rank, double* x, double* coef /* >= 8 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < rank) {
    double s = 0;
    for (int r = lane; r < d; r += 32) s += Q[r * 7 + warp] * x[r];
    s = warp_sum(s);
    if (lane == 0) coef[warp] = s;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < d; r += blockDim.x) {
    double sub = 0;
    for (int i = 0; i < rank; i++) sub += Q[r * 7 + i] * coef[i];
    x[r] -= sub;
  }
  __syncthreads();
}

// LDL^T without pivoting of the scaled system with the matrix in REGISTERS: element (i, j), i >= j, of the (d+1) x d augmented
// matrix [S H S ; (S b)^T] lives in thread (i % 16, j % 16) at local index (i / 16, j / 16) — T x T doubles per thread, T = 4 for
// d <= 63, T = 6 for d <= 95. Step k: the owners of column k publish it (all rows, the right-hand-side row included) in one of two
// shared buffers, ONE block barrier, every thread forms its <= T(T+1)/2 rank-1 updates from registers. The right-hand side rides
// along as row d, so the forward substitution L w = b costs nothing. Returns false (uniformly) on a non-positive pivot. On
// success: M (shared, leading dimension ld) holds the factor in the layout the substitutions below expect (column k unscaled,
// dg[k] = 1 / d_k) and y[] holds w.
template <int T>
__device__ bool ldlt_registers(int d, int ld, const double* __restrict__ HF, const double* __restrict__ sv, const double* __restrict__ bs,
                               double* M, double* dg, double* y, double* colbuf /* 2 x (16 T) */) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  constexpr int CB = 16 * T;
  double m[T][T];
#pragma unroll
  for (int a = 0; a < T; a++)
#pragma unroll
    for (int b = 0; b < T; b++) {
      const int i = a * 16 + ty, j = b * 16 + tx;
      double v = 0.0;
      if (j < d && i >= j) {
        if (i < d) v = sv[i] * HF[(size_t)i * d + j] * sv[j];
        else if (i == d) v = bs[j];   // (already scaled)
      }
      m[a][b] = v;
    }
  bool ok = true;
  for (int k = 0; k < d; k++) {
    double* cb = colbuf + (k & 1) * CB;
    const int kb = k >> 4, kx = k & 15;
    if (tx == kx) {   // owners of column k: rows a * 16 + ty
#pragma unroll
      for (int a = 0; a < T; a++)
#pragma unroll
        for (int b = 0; b < T; b++) if (b == kb) cb[a * 16 + ty] = m[a][b];
    }
    __syncthreads();
    const double mkk = cb[k];
    if (!(mkk > 0.0) || !isfinite(mkk)) { ok = false; break; }   // uniform
    const double rk = 1.0 / mkk;
    double r[T], c[T];
#pragma unroll
    for (int a = 0; a < T; a++) { r[a] = cb[a * 16 + ty] * rk; c[a] = cb[a * 16 + tx]; }
#pragma unroll
    for (int a = 0; a < T; a++)
#pragma unroll
      for (int b = 0; b <= a; b++) {
        const int i = a * 16 + ty, j = b * 16 + tx;
        if (j > k && i >= j && i <= d) m[a][b] -= r[a] * c[b];
      }
    if (tid == 0) dg[k] = rk;
  }
  if (!ok) return false;
  __syncthreads();
  // hand the factor to the substitutions: lower triangle into shared memory, w = row d
#pragma unroll
  for (int a = 0; a < T; a++)
#pragma unroll
    for (int b = 0; b < T; b++) {
      const int i = a * 16 + ty, j = b * 16 + tx;
      if (j < d && i >= j) { if (i < d) M[i * ld + j] = m[a][b]; else if (i == d) y[j] = m[a][b]; }
    }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(256) ba_solve_kernel(SolveParams S) {
  if (S.done && *S.done) return;
  extern __shared__ double sm[];
  const int d = S.d, tid = threadIdx.x, nt = blockDim.x;
  const int ld = d | 1;           // odd leading dimension: column walks hit distinct banks
  double* M = sm;                 // d*ld
  double* bs = M + d * ld;        // d
  double* sv = bs + d;            // d  SVecI
  double* dg = sv + d;            // d  D of LDLT
  double* y = dg + d;             // d
  double* lcol = y + d;           // d  current L column
  double* scr = lcol + d;         // d*7 + 256
  __shared__ int perm[kCPARS + 8 * kMaxFrames];
  __shared__ int piv;
  __shared__ double pivval;
  const int lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 tiling of the trailing update
  for (int i = tid; i < d; i += nt) { bs[i] = S.bF[i]; sv[i] = S.plain ? 1.0 : 1.0 / sqrt(S.HF[(size_t)i * d + i] + 10); perm[i] = i; }
  __syncthreads();
  if (d + 1 > 96)   // (the register-tiled factorisation below loads its elements straight from global memory)
    for (int r = ty; r < d; r += 16)
      for (int c = tx; c < d; c += 16) M[r * ld + c] = sv[r] * S.HF[(size_t)r * d + c] * sv[c];
  for (int i = tid; i < d; i += nt) bs[i] = sv[i] * bs[i];
  __syncthreads();
  // ---- fast path: LDL^T WITHOUT pivoting, one block barrier per column. The scaled system S H S is symmetric positive definite
  // (EnergyFunctional.cpp:967-976: damped top Hessian minus the Schur complement, unit-ish diagonal after the (diag+10)^-1/2
  // scaling), so no pivot search, no row swaps and no column scaling phase are needed: column k stays UNSCALED in place
  // (L_ik = M_ik / d_k is formed on the fly, here and in the substitutions), every thread reads the pivot itself, and the only
  // synchronisation per column is the barrier behind the trailing update. A non-positive or non-finite pivot (the matrix was
  // not positive definite after all) falls through to the pivoted factorisation below, which reloads the system.
  bool spd = true;
  bool have_w = false;   // the forward substitution was folded into the factorisation (y holds w)
  if (d + 1 <= 64) { spd = ldlt_registers<4>(d, ld, S.HF, sv, bs, M, dg, y, scr); have_w = spd; }
  else if (d + 1 <= 96) { spd = ldlt_registers<6>(d, ld, S.HF, sv, bs, M, dg, y, scr); have_w = spd; }
  else {   // larger windows: the same factorisation on the shared-memory copy
    for (int k = 0; k < d; k++) {
      const double mkk = M[k * ld + k];
      if (!(mkk > 0.0) || !isfinite(mkk)) { spd = false; break; }   // uniform: every thread reads the same value
      const double rk = 1.0 / mkk;
      for (int i = k + 1 + ty; i < d; i += 16) {
        const double lik = M[i * ld + k] * rk;
        for (int j = k + 1 + tx; j <= i; j += 16) M[i * ld + j] -= lik * M[j * ld + k];
      }
      if (tid == 0) dg[k] = rk;
      __syncthreads();
    }
  }
  if (spd) {
    if (warp == 0) {  // L y = b, z = D^-1 y, L^T x = z with L_ik = M_ik * dg[k]; column oriented, one warp
      if (!have_w) {
        for (int i = lane; i < d; i += 32) y[i] = bs[i];
        __syncwarp();
        for (int i = 0; i < d; i++) {
          const double yi = y[i] * dg[i];
          for (int j = i + 1 + lane; j < d; j += 32) y[j] -= M[j * ld + i] * yi;
          __syncwarp();
        }
      }
      for (int i = lane; i < d; i += 32) y[i] *= dg[i];
      __syncwarp();
      for (int i = d - 1; i >= 0; i--) {
        const double yi = y[i];
        for (int j = lane; j < i; j += 32) y[j] -= M[i * ld + j] * dg[j] * yi;
        __syncwarp();
      }
      for (int i = lane; i < d; i += 32) bs[i] = y[i];
    }
    __syncthreads();
    for (int i = tid; i < d; i += nt) bs[i] *= sv[i];
    __syncthreads();
    const int iteration_f = S.it_ptr ? *S.it_ptr : S.iteration;
    if (!S.plain && iteration_f >= 2 && S.N) ortho_vec(S.N, d, S.nrank, bs, scr);  // SOLVER_ORTHOGONALIZE_X_LATER (:980-984)
    for (int i = tid; i < d; i += nt) S.x[i] = bs[i];
    return;
  }
  // ---- fallback: reload, then the pivoted factorisation
  __syncthreads();
  for (int r = ty; r < d; r += 16)
    for (int c = tx; c < d; c += 16) M[r * ld + c] = sv[r] * S.HF[(size_t)r * d + c] * sv[c];
  __syncthreads();
  // diagonal-pivoted LDLT (the strategy of Eigen::LDLT, which EnergyFunctional.cpp:976 calls). Two block barriers per pivot:
  // after the swap + scale of column k, warp 0 updates the DIAGONAL of the trailing matrix and picks the next pivot from it
  // while warps 1..7 update the strict lower triangle — the pivot search is off the critical path.
  auto pick_pivot = [&](int k0) {  // warp 0: arg max |M_ii|, i >= k0, lowest index on ties; publishes piv / pivval and swaps perm
    double best = -1.0; int p = k0;
    for (int i = k0 + lane; i < d; i += 32) { const double v = fabs(M[i * ld + i]); if (v > best) { best = v; p = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, p, o);
      if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
    }
    if (lane == 0) {
      piv = p; pivval = M[p * ld + p];
      if (p != k0) { const int q = perm[k0]; perm[k0] = perm[p]; perm[p] = q; }
    }
  };
  if (warp == 0 && d > 0) pick_pivot(0);
  __syncthreads();
  for (int k = 0; k < d; k++) {
    const int p = piv;
    // Symmetric swap of k and p fused with the scaling of column k, one phase: thread j owns (k,j),(p,j),(j,k),(j,p); the thread
    // with j == p owns the 2x2 corner. Only the lower triangle (and the diagonal) is kept current from here on.
    const double dk = pivval;
    const bool singular = (dk == 0.0 || !isfinite(dk));
    for (int j = tid; j < d; j += nt) {
      if (j == k) continue;
      if (j == p) {  // (p != k here) corner: new (k,k) = old (p,p), new (p,p) = old (k,k), new (p,k) = old (k,p) = old (p,k)
        const double akk = M[k * ld + k], apk = M[p * ld + k];
        M[k * ld + k] = dk; M[p * ld + p] = akk;
        const double l = singular ? 0.0 : apk / dk;
        M[p * ld + k] = l; lcol[p] = l;
        continue;
      }
      // lower-triangle accessors: L(a,b) lives at M[max][min]
      if (p != k) {
        double* ek = (j < k) ? &M[k * ld + j] : &M[j * ld + k];
        double* ep = (j < p) ? &M[p * ld + j] : &M[j * ld + p];
        const double t = *ek; *ek = *ep; *ep = t;
      }
      if (j > k) { const double l = singular ? 0.0 : M[j * ld + k] / dk; M[j * ld + k] = l; lcol[j] = l; }
    }
    if (tid == 0) dg[k] = dk;
    __syncthreads();   // piv / pivval of step k are consumed, column k is final
    if (warp == 0) {
      if (!singular) {
        for (int j = k + 1 + lane; j < d; j += 32) M[j * ld + j] -= lcol[j] * dk * lcol[j];
      }
      __syncwarp();
      if (k + 1 < d) pick_pivot(k + 1);
    } else if (!singular) {
      const int t2 = tid - 32, tx2 = t2 & 15, ty2 = t2 >> 4;   // 16 x 14 tiling over the 224 threads of warps 1..7
      for (int i = k + 2 + ty2; i < d; i += 14)
        for (int j = k + 1 + tx2; j < i; j += 16) M[i * ld + j] -= lcol[i] * dk * lcol[j];
    }
    __syncthreads();
  }
  if (warp == 0) {  // triangular solves, column oriented, one warp
    for (int i = lane; i < d; i += 32) y[i] = bs[perm[i]];
    __syncwarp();
    for (int i = 0; i < d; i++) {
      const double yi = y[i];
      for (int j = i + 1 + lane; j < d; j += 32) y[j] -= M[j * ld + i] * yi;
      __syncwarp();
    }
    for (int i = lane; i < d; i += 32) y[i] = (dg[i] != 0.0) ? y[i] / dg[i] : 0.0;
    __syncwarp();
    for (int i = d - 1; i >= 0; i--) {
      const double yi = y[i];
      for (int j = lane; j < i; j += 32) y[j] -= M[i * ld + j] * yi;
      __syncwarp();
    }
    for (int i = lane; i < d; i += 32) bs[perm[i]] = y[i];
  }
  __syncthreads();
  for (int i = tid; i < d; i += nt) bs[i] *= sv[i];
  __syncthreads();
  const int iteration = S.it_ptr ? *S.it_ptr : S.iteration;
  if (!S.plain && iteration >= 2 && S.N) ortho_vec(S.N, d, S.nrank, bs, scr);  // SOLVER_ORTHOGONALIZE_X_LATER (:980-984)
  for (int i = tid; i < d; i += nt) S.x[i] = bs[i];
}

// ---- B11: marginalisation ------------------------------------------------------------------------------------------
// marginalizePointsF (EnergyFunctional.cpp:663-736): priorF *= setting_idepthFixPriorMargFac for the flagged points
__global__ void ba_marg_prior_kernel(int P, const unsigned char* flag, float* priorF, float fac) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P && flag[p] == PS_MARGINALIZE) priorF[p] *= fac;
}
// HM += w (M - Msc), bM += w (Mb - Mbsc)  (:712-718); src buffers are (H,b) pairs of (d*d + d) doubles
__global__ void ba_marg_add_kernel(int count, double w, const double* M, const double* Msc, double* HM, int init) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < count) HM[e] = (init ? 0.0 : HM[e]) + w * (M[e] - Msc[e]);
}
// removePoint for the marginalised points (:720-735): the points and their residuals leave the graph
__global__ void ba_marg_remove_kernel(BAView B, unsigned char* p_flag) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < B.R && p_flag[B.s_point[s]] == PS_MARGINALIZE) B.s_flags[s] = (B.s_flags[s] & ~RF_ACTIVE) | RF_DROPPED;
}
__global__ void ba_marg_flag_kernel(int P, unsigned char* p_flag) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P && p_flag[p] == PS_MARGINALIZE) p_flag[p] = PS_DROP;
}

// marginalizeFrame (EnergyFunctional.cpp:554-660) on ONE CTA: move the frame's 8 rows/cols to the end, add its prior, scale by
// (|diag|+10)^-1/2, invert the 8x8 corner, Schur complement, unscale, symmetrise. in: (HM,bM) of dimension odim; out: (H,b) of
// dimension odim-8, b directly behind H.
__global__ void __launch_bounds__(256) ba_marg_frame_kernel(int odim, int idx, const double* HMin, const double* bMin, const double* fprior,
                                                            double* Hout, double* bout) {
  extern __shared__ double sm[];
  const int ndim = odim - 8, tid = threadIdx.x, nt = blockDim.x;
  double* Hs = sm;                    // odim*odim, permuted + scaled
  double* bs = Hs + odim * odim;      // odim
  double* SV = bs + odim;             // odim
  double* hpi = SV + odim;            // 64
  double* bli = hpi + 64;             // ndim*8
  __shared__ int perm[kCPARS + 8 * kMaxFrames];
  for (int i = tid; i < odim; i += nt) {
    const int io = kCPARS + idx * 8;
    perm[i] = (i < io) ? i : (i < ndim ? i + 8 : io + (i - ndim));
  }
  __syncthreads();
  for (int e = tid; e < odim * odim; e += nt) {
    const int r = e / odim, c = e % odim;
    double v = HMin[(size_t)perm[r] * odim + perm[c]];
    if (r == c && r >= ndim) v += fprior[idx * 24 + (r - ndim)];
    Hs[e] = v;
  }
  for (int i = tid; i < odim; i += nt) {
    double v = bMin[perm[i]];
    if (i >= ndim) v += fprior[idx * 24 + (i - ndim)] * fprior[idx * 24 + 8 + (i - ndim)];
    bs[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < odim; i += nt) SV[i] = sqrt(fabs(Hs[i * odim + i]) + 10);
  __syncthreads();
  for (int e = tid; e < odim * odim; e += nt) Hs[e] = (1.0 / SV[e / odim]) * Hs[e] * (1.0 / SV[e % odim]);
  for (int i = tid; i < odim; i += nt) bs[i] = (1.0 / SV[i]) * bs[i];
  __syncthreads();
  if (tid == 0) {  // 8x8 inverse: Gauss-Jordan with partial pivoting
    double A[8][16];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { const double h = Hs[(ndim + i) * odim + ndim + j]; A[i][j] = 0.5f * (h + h); A[i][8 + j] = (i == j) ? 1.0 : 0.0; }
    for (int k = 0; k < 8; k++) {
      int p = k;
      for (int i = k + 1; i < 8; i++) if (fabs(A[i][k]) > fabs(A[p][k])) p = i;
      if (p != k) for (int j = 0; j < 16; j++) { const double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      const double inv = 1.0 / A[k][k];
      for (int j = 0; j < 16; j++) A[k][j] *= inv;
      for (int i = 0; i < 8; i++) if (i != k) { const double f = A[i][k]; if (f != 0) for (int j = 0; j < 16; j++) A[i][j] -= f * A[k][j]; }
    }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { const double h = A[i][8 + j]; hpi[i * 8 + j] = 0.5f * (h + h); }
  }
  __syncthreads();
  for (int e = tid; e < ndim * 8; e += nt) {
    const int r = e / 8, c = e % 8;
    double sacc = 0;
    for (int k = 0; k < 8; k++) sacc += Hs[(ndim + k) * odim + r] * hpi[k * 8 + c];
    bli[e] = sacc;
  }
  __syncthreads();
  for (int e = tid; e < ndim * ndim; e += nt) {
    const int r = e / ndim, c = e % ndim;
    double sacc = 0;
    for (int k = 0; k < 8; k++) sacc += bli[r * 8 + k] * Hs[(ndim + k) * odim + c];
    Hs[r * odim + c] -= sacc;
  }
  for (int r = tid; r < ndim; r += nt) {
    double sacc = 0;
    for (int k = 0; k < 8; k++) sacc += bli[r * 8 + k] * bs[ndim + k];
    bs[r] -= sacc;
  }
  __syncthreads();
  for (int e = tid; e < ndim * ndim; e += nt) {
    const int r = e / ndim, c = e % ndim;
    const double a = SV[r] * Hs[r * odim + c] * SV[c], bb = SV[c] * Hs[c * odim + r] * SV[r];
    Hout[e] = 0.5 * (a + bb);
  }
  for (int r = tid; r < ndim; r += nt) bout[r] = SV[r] * bs[r];
}

// calcMEnergyF (:344-351): delta^T (2 bM + HM delta) ; calcLEnergyF (:354-442): priors + linearised residuals + point priors
__global__ void __launch_bounds__(256) ba_menergy_kernel(int d, const double* HM, const double* bM, const double* fprior, const float* cDeltaF, double* out) {
  __shared__ double red[32];
  double acc = 0;
  for (int r = threadIdx.x; r < d; r += blockDim.x) {
    const double dr = r < 4 ? (double)cDeltaF[r] : fprior[((r - 4) / 8) * 24 + 16 + (r - 4) % 8];
    double sacc = 0;
    for (int c = 0; c < d; c++) { const double dc = c < 4 ? (double)cDeltaF[c] : fprior[((c - 4) / 8) * 24 + 16 + (c - 4) % 8]; sacc += HM[(size_t)r * d + c] * dc; }
    acc += dr * (2 * bM[r] + sacc);
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) out[0] = t;
}
__global__ void __launch_bounds__(128) ba_lenergy_kernel(BAView B, double* part) {
  __shared__ double red[32];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0;
  if (p < B.P && B.p_flag[p] != PS_DROP) {
    const float dd = B.p_deltaF[p];
    const float* cd = B.cDeltaF;
    for (int i = B.p_res_begin[p]; i < B.p_res_begin[p + 1]; i++) {
      const int s = B.p_res_list[i];
      const unsigned char fl = B.s_flags[s];
      if (!(fl & RF_LINEARIZED) || !(fl & RF_ACTIVE)) continue;
      const int buf = B.s_sel[s];
      const float* dp = B.adHTdeltaF + (size_t)B.s_key[s] * 8;
      float jx[6], jy[6], cx[4], cy[4];
#pragma unroll
      for (int k = 0; k < 6; k++) { jx[k] = *jplane(B, buf, J_PDXI + k, s); jy[k] = *jplane(B, buf, J_PDXI + 6 + k, s); }
#pragma unroll
      for (int k = 0; k < 4; k++) { cx[k] = *jplane(B, buf, J_PDC + k, s); cy[k] = *jplane(B, buf, J_PDC + 4 + k, s); }
      const float Jpx = (jx[0] * dp[0] + jx[1] * dp[1] + jx[2] * dp[2] + jx[3] * dp[3] + jx[4] * dp[4] + jx[5] * dp[5]) +
                        (cx[0] * cd[0] + cx[1] * cd[1] + cx[2] * cd[2] + cx[3] * cd[3]) + *jplane(B, buf, J_PDD, s) * dd;
      const float Jpy = (jy[0] * dp[0] + jy[1] * dp[1] + jy[2] * dp[2] + jy[3] * dp[3] + jy[4] * dp[4] + jy[5] * dp[5]) +
                        (cy[0] * cd[0] + cy[1] * cd[1] + cy[2] * cd[2] + cy[3] * cd[3]) + *jplane(B, buf, J_PDD + 1, s) * dd;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float Jdelta = *jplane(B, buf, J_IDX + k, s) * Jpx;
        Jdelta = Jdelta + *jplane(B, buf, J_IDX + 8 + k, s) * Jpy;
        Jdelta = Jdelta + *jplane(B, buf, J_AB + k, s) * dp[6];
        Jdelta = Jdelta + *jplane(B, buf, J_AB + 8 + k, s) * dp[7];
        float r0 = B.s_rtz[(size_t)k * B.capR + s];
        r0 = r0 + r0; r0 = r0 + Jdelta;
        acc += Jdelta * r0;
      }
    }
    acc += dd * dd * B.p_priorF[p];
  }
  const double t = block_sum_d((double)acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// ---- E2: EdgeLBASE3PosePhotoIdepthCamDSO::computeError + linearizeOplus (dso_g2o_edge.cpp:5-282), one thread per residual ----
struct LBAEdgeParams {
  const double* T_wh;    // [n][12] vertex poses (per host frame)
  const double* T_tw;    // [n][12] PRE_worldToCam of the frames (targets)
  const double* photo;   // [n][2]
  const double* idepth;  // [R] (caller order mapped through slot2rid)
  const double* b0;      // [n]
  const double* target_aff;  // [n][2] aff_g2l of the frames
  const float* exposure;     // [n]
  double cam[4];
  const double* cam_dev;     // driver: the camera vertex lives on the device (null: cam[] above)
  const double* run_if;      // driver: skip the launch's work when *run_if == 0 (the damping trial's x was not finite)
  const int* slot2rid;   // operator mode: outputs in the caller's residual order; null (driver mode): slot order
  int driver;            // 1: g2o driver semantics — skip inactive edges, keep stale Jacobians / energies where the edge returns early,
                         //    sticky level, idepth / outputs indexed by slot
  const unsigned char* active;   // driver mode: level-0 edges of initializeOptimization() (null while the graph is being built)
  int linearize;         // driver mode: 0 = computeError only, 1 = computeError + linearizeOplus
  double* error8; double* Jxi; double* Jphoto; double* Jid; double* JC;
  int* newState; double* newEnergy; double* newEnergyWO; float* center3; float* idepth_hessian; int* level;
};

__global__ void __launch_bounds__(128) ba_lba_edge_kernel(BAView B, LBAEdgeParams E) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.R) return;
  const int rid = E.slot2rid ? E.slot2rid[s] : s;
  if (E.driver && E.active && !E.active[s]) return;
  if (E.run_if && *E.run_if == 0.0) return;
  const int key = B.s_key[s], h = key % B.n, t = key / B.n, pidx = B.s_point[s];
  double* err = E.error8 + 8 * (size_t)rid;
  double* Jxi = E.Jxi + 48 * (size_t)rid; double* Jph = E.Jphoto + 16 * (size_t)rid; double* Jid = E.Jid + 8 * (size_t)rid; double* JC = E.JC + 32 * (size_t)rid;
  if (!E.driver) {
    for (int i = 0; i < 8; i++) { err[i] = 0; Jid[i] = 0; }
    for (int i = 0; i < 48; i++) Jxi[i] = 0;
    for (int i = 0; i < 16; i++) Jph[i] = 0;
    for (int i = 0; i < 32; i++) JC[i] = 0;
    E.newEnergy[rid] = -1; E.newEnergyWO[rid] = -1; E.level[rid] = 0; E.idepth_hessian[rid] = 0;
    E.center3[3 * rid] = E.center3[3 * rid + 1] = E.center3[3 * rid + 2] = 0;
  }
  int newState = E.driver ? E.newState[rid] : B.s_newstate[s];
  const double* camv = E.cam_dev ? E.cam_dev : E.cam;
  const double fx = camv[0], fy = camv[1], cx = camv[2], cy = camv[3];
  // Tth = Ttw * Twh in double, then cast to float (:23-29)
  const double* A = E.T_tw + 12 * t; const double* Bm = E.T_wh + 12 * h;
  float R[9], tt[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) R[r * 3 + c] = (float)(A[r * 4] * Bm[c] + A[r * 4 + 1] * Bm[4 + c] + A[r * 4 + 2] * Bm[8 + c]);
    tt[r] = (float)(A[r * 4] * Bm[3] + A[r * 4 + 1] * Bm[7] + A[r * 4 + 2] * Bm[11] + A[r * 4 + 3]);
  }
  float eF = E.exposure[h], eT = E.exposure[t];
  if (eF == 0 || eT == 0) { eT = eF = 1; }
  const double aa = exp(E.target_aff[2 * t] - E.photo[2 * h]) * eT / eF;
  const float ab0 = (float)aa, ab1 = (float)(E.target_aff[2 * t + 1] - aa * E.photo[2 * h + 1]);
  const double idepth = E.idepth[rid], b0 = E.b0[h];
  const float pu = B.p_u[pidx], pv = B.p_v[pidx];
  const float4 col0 = B.p_color[2 * pidx], col1 = B.p_color[2 * pidx + 1], wt0 = B.p_weights[2 * pidx], wt1 = B.p_weights[2 * pidx + 1];
  const float color[8] = {col0.x, col0.y, col0.z, col0.w, col1.x, col1.y, col1.z, col1.w};
  const float weights[8] = {wt0.x, wt0.y, wt0.z, wt0.w, wt1.x, wt1.y, wt1.z, wt1.w};
  const int wl = B.cp->w0 - 3, hl = B.cp->h0 - 3;
  const float4* tex = B.tex0[t];
  float energyLeft = 0, wJI2_sum = 0;
  double drs[8], us[8], vs[8], nids[8]; float3 hits[8]; float K0s[8], K1s[8];
  bool finite_all = true;
  for (int idx = 0; idx < 8; idx++) {
    const double u_host = pu + kPatternP[idx][0], v_host = pv + kPatternP[idx][1];
    const float K0 = (float)((u_host - cx) / fx), K1 = (float)((v_host - cy) / fy);
    const float idf = (float)idepth;
    float ptp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) ptp[k] = (R[k * 3] * K0 + R[k * 3 + 1] * K1 + R[k * 3 + 2] * 1.0f) + tt[k] * idf;
    const double drescale = 1.0f / ptp[2];
    if (drescale <= 0) { E.newState[rid] = RS_OOB; for (int i = 0; i < 8; i++) err[i] = 0; return; }
    const double new_idepth = idepth * drescale;
    const double _u = ptp[0] * drescale, _v = ptp[1] * drescale;
    const double _Ku = _u * fx + cx, _Kv = _v * fy + cy;
    if ((_Ku - 2) < 0 || (_Ku + 3) > wl || (_Kv - 2) < 0 || (_Kv + 3) > hl) {
      E.newState[rid] = RS_OOB; for (int i = 0; i < 8; i++) err[i] = 0; E.level[rid] = 1; return;
    }
    if (kPatternP[idx][0] == 0 && kPatternP[idx][1] == 0) { E.center3[3 * rid] = (float)_Ku; E.center3[3 * rid + 1] = (float)_Kv; E.center3[3 * rid + 2] = (float)new_idepth; }
    const float3 hit = interp33(tex, (float)_Ku, (float)_Kv, B.cp->w0);
    drs[idx] = drescale; us[idx] = _u; vs[idx] = _v; nids[idx] = new_idepth; hits[idx] = hit; K0s[idx] = K0; K1s[idx] = K1;
    if (!isfinite(hit.x)) { newState = RS_OOB; err[idx] = 0; finite_all = false; continue; }
    const double e = hit.x - (ab0 * color[idx] + ab1);
    err[idx] = e;
    float w = sqrtf(B.cp->outlierTHSumComponent / (B.cp->outlierTHSumComponent + (hit.y * hit.y + hit.z * hit.z)));
    w = 0.5f * (w + weights[idx]);
    const float hw = fabsf((float)e) < B.cp->huberTH ? 1 : B.cp->huberTH / fabsf((float)e);
    energyLeft += w * w * hw * e * e * (2 - hw);
    wJI2_sum += hw * hw * (hit.y * hit.y + hit.z * hit.z);
  }
  E.newEnergyWO[rid] = energyLeft;
  const float th = fmaxf(B.frameTH[h], B.frameTH[t]);
  if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; newState = RS_OUTLIER; }
  else newState = RS_IN;
  E.newEnergy[rid] = energyLeft;
  if (!finite_all) { E.newState[rid] = RS_OOB; return; }  // linearizeOplus bails out at the first non-finite pixel (:205-208)
  E.newState[rid] = newState;
  if (E.driver && (!E.linearize || E.level[rid] == 1)) return;  // computeError only / `if(level() == 1) return;` (:131)
  float Hii = 0;
  for (int idx = 0; idx < 8; idx++) {
    const double drescale = drs[idx], _u = us[idx], _v = vs[idx], new_idepth = nids[idx];
    const float3 hit = hits[idx];
    const double fxi = 1 / fx, fyi = 1 / fy;
    double dC[2][4];
    dC[0][2] = drescale * (R[6] * _u - R[0]);
    dC[0][3] = fx * fyi * drescale * (R[7] * _u - R[1]);
    dC[0][0] = K0s[idx] * dC[0][2];
    dC[0][1] = K1s[idx] * dC[0][3];
    dC[1][2] = fy * fxi * drescale * (R[6] * _v - R[3]);
    dC[1][3] = drescale * (R[7] * _v - R[4]);
    dC[1][0] = K0s[idx] * dC[1][2];
    dC[1][1] = K1s[idx] * dC[1][3];
    for (int k = 0; k < 4; k++) JC[idx * 4 + k] = (double)hit.y * dC[0][k] + (double)hit.z * dC[1][k];
    const double dx = hit.y * fx, dy = hit.z * fy;
    Jxi[idx * 6 + 0] = new_idepth * dx;
    Jxi[idx * 6 + 1] = new_idepth * dy;
    Jxi[idx * 6 + 2] = -new_idepth * (_u * dx + _v * dy);
    Jxi[idx * 6 + 3] = -(_u * _v * dx + (1 + _v * _v) * dy);
    Jxi[idx * 6 + 4] = _u * _v * dy + (1 + _u * _u) * dx;
    Jxi[idx * 6 + 5] = _u * dy - _v * dx;
    Jph[idx * 2 + 0] = ab0 * (b0 - color[idx]);
    Jph[idx * 2 + 1] = -1;
    const double jd = dx * drescale * (tt[0] - tt[2] * _u) + dy * drescale * (tt[1] - tt[2] * _v);
    Jid[idx] = jd;
    Hii += jd * jd;
  }
  if (Hii < 1e-10) Hii = 1e-10;
  E.idepth_hessian[rid] = Hii;
}

// ---- D4: point activation (ImmaturePoint::linearizeResidual, ImmaturePoint.cpp:886-985; FullSystem::optimizeImmaturePoint,
// FullSystemOptPoint.cpp:52-238). One thread per candidate: the sums over pixels and targets stay in the reference's order, so
// the IN / OUTLIER states, the Hdd >= minIdepthH_act test and the accept / reject decisions are exact. ------------------------
struct ActParams {
  const sdso_immature_point* pts; const int* host; int n;
  int variant, minObs, GNIts;
  float minIdepthH_act;
  int* result; float* idepth; int* states; float* energy;
};

struct ActRes { int state_state, state_NewState; float state_energy, state_NewEnergy; };

template <bool G2O>
__device__ __forceinline__ float act_lin_res(const BAView& B, const sdso_immature_point& p, int host, int target, float slack, ActRes& tr,
                                             float& Hdd, float& bd, float idepth) {
  if (tr.state_state == RS_OOB) { tr.state_NewState = RS_OOB; return tr.state_energy; }
  const PrecalcDev& pc = B.precalc[host * B.n + target];
  const float4* tex = B.tex0[target];
  const BACalib& c = *B.cp;
  float Ku[8], Kv[8], uu[8], vv[8], dr[8];
  bool inside[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    const float K0 = (p.u + kPatternP[idx][0] - c.cxl) * c.fxli, K1 = (p.v + kPatternP[idx][1] - c.cyl) * c.fyli;
    float ptp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) ptp[k] = (pc.PRE_RTll[k * 3] * K0 + pc.PRE_RTll[k * 3 + 1] * K1 + pc.PRE_RTll[k * 3 + 2] * 1.0f) + pc.PRE_tTll[k] * idepth;
    const float drescale = 1.0f / ptp[2];
    dr[idx] = drescale;
    bool ok = drescale > 0;
    uu[idx] = vv[idx] = Ku[idx] = Kv[idx] = 0;
    if (ok) {
      uu[idx] = ptp[0] * drescale; vv[idx] = ptp[1] * drescale;
      Ku[idx] = uu[idx] * c.fxl + c.cxl; Kv[idx] = vv[idx] * c.fyl + c.cyl;
      if (G2O) ok = !(((double)Ku[idx] - 2) < 0 || ((double)Ku[idx] + 3) > c.w0 - 3 || ((double)Kv[idx] - 2) < 0 || ((double)Kv[idx] + 3) > c.h0 - 3);
      else ok = Ku[idx] > 1.1f && Kv[idx] > 1.1f && Ku[idx] < c.wM3G && Kv[idx] < c.hM3G;
    }
    inside[idx] = ok;
  }
  float3 hit[8];
#pragma unroll
  for (int idx = 0; idx < 8; idx++) hit[idx] = inside[idx] ? interp33(tex, Ku[idx], Kv[idx], c.w0) : make_float3(0.f, 0.f, 0.f);
  float energyLeft = 0;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    float residual = 0;
    if (G2O) {
      if (inside[idx] && isfinite(hit[idx].x)) residual = (float)((double)hit[idx].x - ((double)pc.PRE_aff_mode[0] * (double)p.color[idx] + (double)pc.PRE_aff_mode[1]));  // _measurement is a double
    } else {
      // the reference returns at the FIRST failing pixel, after the pixels before it have already been added to Hdd / bd
      if (!inside[idx] || !isfinite(hit[idx].x)) { tr.state_NewState = RS_OOB; return tr.state_energy; }
      residual = hit[idx].x - (pc.PRE_aff_mode[0] * p.color[idx] + pc.PRE_aff_mode[1]);
    }
    float hw = fabsf(residual) < c.huberTH ? 1 : c.huberTH / fabsf(residual);
    energyLeft += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
    if (!G2O) {
      const float dxInterp = hit[idx].y * c.fxl, dyInterp = hit[idx].z * c.fyl;
      const float d_idepth = (dxInterp * dr[idx] * (pc.PRE_tTll[0] - pc.PRE_tTll[2] * uu[idx]) + dyInterp * dr[idx] * (pc.PRE_tTll[1] - pc.PRE_tTll[2] * vv[idx])) * SCALE_IDEPTH;
      hw *= p.weights[idx] * p.weights[idx];
      Hdd += (hw * d_idepth) * d_idepth;
      bd += (hw * residual) * d_idepth;
    }
  }
  if (energyLeft > p.energyTH * slack) { energyLeft = p.energyTH * slack; tr.state_NewState = RS_OUTLIER; }
  else tr.state_NewState = RS_IN;
  tr.state_NewEnergy = energyLeft;
  return energyLeft;
}

__global__ void __launch_bounds__(64) ba_activate_kernel(BAView B, ActParams A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const sdso_immature_point p = A.pts[i];
  const int host = A.host[i], nf = B.n;
  ActRes res[kMaxFrames];
  for (int f = 0; f < nf; f++) { res[f].state_state = RS_IN; res[f].state_NewState = RS_OUTLIER; res[f].state_energy = 0; res[f].state_NewEnergy = 0; A.states[(size_t)i * nf + f] = -1; }
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (p.idepth_max + p.idepth_min) * 0.5f;
  int result = 1;
  bool done = false;
  if (A.variant == 1) {
    for (int f = 0; f < nf; f++) {
      if (f == host) continue;
      lastEnergy += act_lin_res<true>(B, p, host, f, 1000, res[f], lastHdd, lastbd, (float)(double)currentIdepth);
      res[f].state_state = res[f].state_NewState; res[f].state_energy = res[f].state_NewEnergy;
    }
  } else {
    for (int f = 0; f < nf; f++) {
      if (f == host) continue;
      lastEnergy += act_lin_res<false>(B, p, host, f, 1000, res[f], lastHdd, lastbd, currentIdepth);
      res[f].state_state = res[f].state_NewState; res[f].state_energy = res[f].state_NewEnergy;
    }
    if (!isfinite(lastEnergy) || lastHdd < A.minIdepthH_act) { result = 0; done = true; }
    float lambda = 0.1f;
    for (int it = 0; it < A.GNIts && !done; it++) {
      float Hh = lastHdd;
      Hh *= 1 + lambda;
      const float step = (float)((1.0 / Hh) * lastbd);
      const float newIdepth = currentIdepth - step;
      float newHdd = 0, newbd = 0, newEnergy = 0;
      for (int f = 0; f < nf; f++) { if (f == host) continue; newEnergy += act_lin_res<false>(B, p, host, f, 1, res[f], newHdd, newbd, newIdepth); }
      if (!isfinite(lastEnergy) || newHdd < A.minIdepthH_act) { result = 0; done = true; break; }
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        for (int f = 0; f < nf; f++) { res[f].state_state = res[f].state_NewState; res[f].state_energy = res[f].state_NewEnergy; }
        lambda *= 0.5f;
      } else lambda *= 5;
      if (fabsf(step) < 0.0001 * currentIdepth) break;
    }
  }
  A.idepth[i] = currentIdepth; A.energy[i] = lastEnergy;
  if (!done) {
    int numGood = 0;
    for (int f = 0; f < nf; f++) { if (f == host) continue; A.states[(size_t)i * nf + f] = res[f].state_state; if (res[f].state_state == RS_IN) numGood++; }
    if (!isfinite(currentIdepth)) result = -1;
    else if (numGood < A.minObs) result = -1;
    else if (!isfinite(p.energyTH)) result = -1;
  }
  A.result[i] = result;
}

// ---- g2o LBA driver (FullSystem::optimize, g2o body, FullSystemOptimize.cpp:404-868; restated g2o LM + Schur) ----
struct LBAGraph {   // per-edge arrays in SLOT order
  int R;
  const unsigned char* active;
  double* idepth; double* idepth_bak;
  const double* err;      // [R][8]
  const double* Jxi; const double* Jph; const double* Jid; const double* JC;   // [R][48], [R][16], [R][8], [R][32]
  double* hll; double* bl; double* hpl;   // [R], [R], [R][12]
  double delta;
};

__device__ __forceinline__ double huber_rho_d(double e2, double delta, double& rho1) {
  if (e2 <= delta * delta) { rho1 = 1.0; return e2; }
  const double sq = sqrt(e2);
  rho1 = delta / sq;
  return 2 * sq * delta - delta * delta;
}
// column a (0..11 = pose 6 | photo 2 | cam 4, 12 = idepth) of the 8x13 edge Jacobian, row k
__device__ __forceinline__ double lba_J(const LBAGraph& G, int s, int k, int a) {
  if (a < 6) return G.Jxi[(size_t)s * 48 + k * 6 + a];
  if (a < 8) return G.Jph[(size_t)s * 16 + k * 2 + (a - 6)];
  if (a < 12) return G.JC[(size_t)s * 32 + k * 4 + (a - 8)];
  return G.Jid[(size_t)s * 8 + k];
}

// activeRobustChi2: sum over the active edges of Huber(e^T e); fixed-order block partials
__global__ void __launch_bounds__(128) lba_chi2_kernel(LBAGraph G, double* part) {
  __shared__ double red[32];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  if (s < G.R && G.active[s]) {
    double e2 = 0, r1;
    for (int k = 0; k < 8; k++) { const double e = G.err[(size_t)s * 8 + k]; e2 += e * e; }
    v = huber_rho_d(e2, G.delta, r1);
  }
  const double t = block_sum_d(v, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void lba_sum_kernel(const double* part, int n, int stride, int nvals, double* out) {
  const int v = threadIdx.x;
  if (v < nvals) { double s = 0; for (int i = 0; i < n; i++) s += part[(size_t)i * stride + v]; out[v] = s; }
}

// One CTA per chunk (<= 256 edges of one (host,target) pair). what == 0: buildSystem — J^T rho' J (12x12 upper triangle, 78 values)
// and -J^T rho' e (12) of every active edge summed over the chunk, and the edge's own idepth terms hll, bl, hpl stored per edge.
// what == 1: the Schur complement terms hpl hpl^T / (hll + lambda) (78) and hpl bl / (hll + lambda) (12) for the current lambda.
// Sums in double, fixed order, through shared memory (64-bit shuffles are slow on this part).
__global__ void __launch_bounds__(kChunk) lba_build_kernel(BAView B, LBAGraph G, int what, double lambda, double* part /* [nchunks][96] */) {
  __shared__ double wbuf[kChunk / 32][32][14];   // per warp: 32 lanes x up to 13 values (+1 pad)
  __shared__ double wtot[kChunk / 32][14];
  const Chunk ch = B.chunks[blockIdx.x];
  const int s = ch.begin + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool use = s < ch.end && G.active[s];
  double rho1 = 0, inv = 0, blv = 0;
  double hp[12];
#pragma unroll
  for (int a = 0; a < 12; a++) hp[a] = 0;
  if (use) {
    if (what == 0) {
      double e2 = 0;
      for (int k = 0; k < 8; k++) { const double e = G.err[(size_t)s * 8 + k]; e2 += e * e; }
      huber_rho_d(e2, G.delta, rho1);
      double hll = 0, bl = 0;
      for (int k = 0; k < 8; k++) { const double jd = G.Jid[(size_t)s * 8 + k]; hll += jd * rho1 * jd; bl -= jd * rho1 * G.err[(size_t)s * 8 + k]; }
      G.hll[s] = hll; G.bl[s] = bl;
    } else {
      inv = 1.0 / (G.hll[s] + lambda); blv = G.bl[s];
#pragma unroll
      for (int a = 0; a < 12; a++) hp[a] = G.hpl[(size_t)s * 12 + a];
    }
  }
  int out = 0;  // running index into the 90 values: row a -> (12 - a) matrix entries, then all 12 b entries
  for (int a = 0; a < 13; a++) {
    const int nv = (a < 12) ? 12 - a : 12;
    double v[12];
#pragma unroll
    for (int q = 0; q < 12; q++) v[q] = 0;
    if (use) {
      if (what == 0) {
        if (a < 12) {
          double ja[8];
          for (int k = 0; k < 8; k++) ja[k] = lba_J(G, s, k, a) * rho1;
          for (int q = 0; q < nv; q++) { double t = 0; for (int k = 0; k < 8; k++) t += ja[k] * lba_J(G, s, k, a + q); v[q] = t; }
          double t = 0;
          for (int k = 0; k < 8; k++) t += ja[k] * G.Jid[(size_t)s * 8 + k];
          G.hpl[(size_t)s * 12 + a] = t;
        } else {
          for (int q = 0; q < 12; q++) { double t = 0; for (int k = 0; k < 8; k++) t += lba_J(G, s, k, q) * rho1 * G.err[(size_t)s * 8 + k]; v[q] = -t; }
        }
      } else {
        if (a < 12) { for (int q = 0; q < nv; q++) v[q] = hp[a] * inv * hp[a + q]; }
        else { for (int q = 0; q < 12; q++) v[q] = hp[q] * inv * blv; }
      }
    }
    for (int q = 0; q < nv; q++) wbuf[warp][lane][q] = v[q];
    __syncwarp();
    if (lane < nv) { double t = 0; for (int l = 0; l < 32; l++) t += wbuf[warp][l][lane]; wtot[warp][lane] = t; }
    __syncthreads();
    if (threadIdx.x < nv) { double t = 0; for (int w = 0; w < kChunk / 32; w++) t += wtot[w][threadIdx.x]; part[(size_t)blockIdx.x * 96 + out + threadIdx.x] = t; }
    __syncthreads();
    out += nv;
  }
}

// per host: fixed-order sum of the chunk partials whose key has that host. grid = n hosts, 96 threads
__global__ void lba_host_sum_kernel(BAView B, const double* part, double* hostsum /* [n][96] */) {
  const int h = blockIdx.x, v = threadIdx.x, n = B.n;
  double s = 0;
  for (int t = 0; t < n; t++) {
    const int key = h + t * n;
    for (int c = B.key_chunk_begin[key]; c < B.key_chunk_begin[key + 1]; c++) s += part[(size_t)c * 96 + v];
  }
  hostsum[(size_t)h * 96 + v] = s;
}

// (Hpp + lambda I - Schur(lambda)) and (bp - schur_b) as a dense (4+8n) system in the layout [cam 4 | host h: pose 6, photo 2]
__device__ __forceinline__ int lba_tri(int a, int c) { if (a > c) { const int q = a; a = c; c = q; } return a * 12 - (a * (a - 1)) / 2 + (c - a); }
__global__ void lba_assemble_kernel(int n, const double* A /* [n][96] */, const double* Sc /* [n][96] */, const int* used, double lambda, double* H, double* bvec) {
  const int d = kCPARS + 8 * n;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d + d) return;
  if (e >= d * d) {
    const int r = e - d * d;
    double s = 0;
    if (r < 4) { for (int h = 0; h < n; h++) if (used[h]) s += A[(size_t)h * 96 + 78 + 8 + r] - Sc[(size_t)h * 96 + 78 + 8 + r]; }
    else { const int h = (r - 4) / 8, l = (r - 4) % 8; s = used[h] ? A[(size_t)h * 96 + 78 + l] - Sc[(size_t)h * 96 + 78 + l] : 0.0; }
    bvec[r] = s;
    return;
  }
  const int r = e / d, c = e % d;
  double s = 0;
  if (r < 4 && c < 4) {
    for (int h = 0; h < n; h++) if (used[h]) { const int i = lba_tri(8 + r, 8 + c); s += A[(size_t)h * 96 + i] - Sc[(size_t)h * 96 + i]; }
    if (r == c) s += lambda;
  } else if (r < 4 || c < 4) {
    const int cam = r < 4 ? r : c, o = r < 4 ? c : r;
    const int h = (o - 4) / 8, l = (o - 4) % 8;
    if (used[h]) { const int i = lba_tri(l, 8 + cam); s = A[(size_t)h * 96 + i] - Sc[(size_t)h * 96 + i]; }
  } else {
    const int h1 = (r - 4) / 8, l1 = (r - 4) % 8, h2 = (c - 4) / 8, l2 = (c - 4) % 8;
    if (h1 == h2) {
      if (used[h1]) { const int i = lba_tri(l1, l2); s = A[(size_t)h1 * 96 + i] - Sc[(size_t)h1 * 96 + i]; if (r == c) s += lambda; }
      else s = (r == c) ? 1.0 : 0.0;
    }
  }
  H[e] = s;
}

// idepth_r += (bl - hpl^T dx) / (hll + lambda); partial sums of dl (lambda dl + bl) for computeScale
__global__ void __launch_bounds__(128) lba_update_kernel(BAView B, LBAGraph G, const double* x, double lambda, double* part, const double* run_if) {
  __shared__ double red[32];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  if (s < G.R && G.active[s] && (!run_if || *run_if != 0.0)) {
    const int hb = kCPARS + 8 * (B.s_key[s] % B.n);
    double t = G.bl[s];
    for (int a = 0; a < 12; a++) t -= G.hpl[(size_t)s * 12 + a] * x[a < 8 ? hb + a : a - 8];
    const double dl = t / (G.hll[s] + lambda);
    G.idepth[s] += dl;
    v = dl * (lambda * dl + G.bl[s]);
  }
  const double t = block_sum_d(v, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void lba_copy_kernel(int n, const double* src, double* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// initializeOptimization(): the level-0 edges among the graph's edges
__global__ void lba_activate_kernel(int R, const unsigned char* in_graph, const int* level, unsigned char* active) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < R) active[s] = (in_graph[s] && level[s] == 0) ? 1 : 0;
}

// ---- B9 -------------------------------------------------------------------------------------------------------
// xAd[h*n + t] = x_h^T adHostF[h + t*n] + x_t^T adTargetF[h + t*n]   (EnergyFunctional.cpp:283-293)
__global__ void ba_xad_kernel(BAView B, const double* x, float* xAd) {
  BA_EXIT_IF_DONE(B);
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
  BA_EXIT_IF_DONE(B);
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


// ---- device-resident LM loop of FullSystem::optimize ---------------------------------------------------------------------------
// doStepFromBackup(1,1,1,1,1) for the frames and the calibration (FullSystemOptimize.cpp:207-305, the point part is
// ba_step_points_kernel), FrameHessian::setState (HessianBlocks.h:177-199), FullSystem::setPrecalcValues ->
// FrameFramePrecalc::set at the current state (HessianBlocks.cpp:206-242), EnergyFunctional::setDeltaF (:173-207) and the
// convergence test (:298-301) — what used to be two host round trips per LM iteration — in ONE small CTA. SE3 arithmetic in
// double with the formulas of se3_host.cu (the reference computes these on the host in double as well).
__device__ inline void dev_se3_exp(const double a[6], double T[12]) {
  const double* w = a + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double A, Bc, C;
  if (th < 1e-5) { A = 1 - th2 / 6 + th2 * th2 / 120; Bc = 0.5 - th2 / 24 + th2 * th2 / 720; C = 1.0 / 6 - th2 / 120 + th2 * th2 / 5040; }
  else { A = sin(th) / th; Bc = (1 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th); }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) W2[r * 3 + c] = W[r * 3] * W[c] + W[r * 3 + 1] * W[3 + c] + W[r * 3 + 2] * W[6 + c];
  double R[9], V[9];
  for (int i = 0; i < 9; i++) { const double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + A * W[i] + Bc * W2[i]; V[i] = I + Bc * W[i] + C * W2[i]; }
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = V[r * 3] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  }
}
__device__ inline void dev_se3_mul(const double A[12], const double B[12], double C[12]) {
  double out[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
    out[r * 4 + 3] = A[r * 4] * B[3] + A[r * 4 + 1] * B[7] + A[r * 4 + 2] * B[11] + A[r * 4 + 3];
  }
  for (int i = 0; i < 12; i++) C[i] = out[i];
}
__device__ inline void dev_se3_inv(const double A[12], double B[12]) {
  double out[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[r * 4 + c] = A[c * 4 + r];
    out[r * 4 + 3] = -(A[0 * 4 + r] * A[3] + A[1 * 4 + r] * A[7] + A[2 * 4 + r] * A[11]);
  }
  for (int i = 0; i < 12; i++) B[i] = out[i];
}
__device__ inline void dev_mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}
__device__ inline void dev_inverse3f(const float m[9], float out[9]) {  // = inverse3f (ctx.cu): cofactors * (1/det)
  float cofm[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    cofm[i * 3 + j] = m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
  }
  const float det = (cofm[0] * m[0] + cofm[3] * m[3]) + cofm[6] * m[6];
  const float invdet = 1.0f / det;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[j * 3 + i] = cofm[i * 3 + j] * invdet;
}

struct FrameUpdateParams {
  int n, P;
  FrameDev* frames; OptDev* opt; BACalib* calib; float* cDeltaF; double* fprior; PrecalcDev* precalc;
  const float* adHostF; const float* adTargetF; float* adHTdeltaF;
  const double* x;          // the solved increment (SYS_X); steps are -x
  const double* step_sums;  // [2]: sum of squared point steps, sum of |idepth_backup| (ba_sum_pairs_kernel)
};

__global__ void __launch_bounds__(256) ba_frame_update_kernel(FrameUpdateParams U) {
  if (U.opt->done) return;
  __shared__ double s_step[kMaxFrames][8];
  const int tid = threadIdx.x, n = U.n;
  if (tid < n) {   // backupState + doStepFromBackup + setState of frame tid
    FrameDev& f = U.frames[tid];
    double sc[10];
    for (int i = 0; i < 10; i++) {
      const double st = i < 8 ? -U.x[kCPARS + 8 * tid + i] : 0.0;
      if (i < 8) s_step[tid][i] = st;
      f.state_backup[i] = f.state[i];
      f.state[i] = f.state_backup[i] + 1.0f * st;
    }
    for (int i = 0; i < 3; i++) sc[i] = SCALE_XI_TRANS * f.state[i];
    for (int i = 3; i < 6; i++) sc[i] = SCALE_XI_ROT * f.state[i];
    sc[6] = SCALE_A * f.state[6]; sc[7] = SCALE_B * f.state[7]; sc[8] = SCALE_A * f.state[8]; sc[9] = SCALE_B * f.state[9];
    double E[12];
    dev_se3_exp(sc, E);
    dev_se3_mul(E, f.T_eval, f.T_w2c);
    dev_se3_inv(f.T_w2c, f.T_c2w);
    for (int i = 0; i < 8; i++) { U.fprior[tid * 24 + 8 + i] = f.state[i]; U.fprior[tid * 24 + 16 + i] = f.state[i] - f.state_zero[i]; }
  } else if (tid == 32) {   // the calibration: CalibHessian::setValue (HessianBlocks.h:316-331)
    OptDev& o = *U.opt;
    BACalib& c = *U.calib;
    for (int i = 0; i < 4; i++) { o.calib_backup[i] = o.calib_value[i]; o.calib_value[i] = o.calib_backup[i] + 1.0f * (-U.x[i]); }
    const double vs[4] = {SCALE_F * o.calib_value[0], SCALE_F * o.calib_value[1], SCALE_C * o.calib_value[2], SCALE_C * o.calib_value[3]};
    c.fxl = (float)vs[0]; c.fyl = (float)vs[1]; c.cxl = (float)vs[2]; c.cyl = (float)vs[3];
    c.fxli = 1.0f / c.fxl; c.fyli = 1.0f / c.fyl; c.cxli = -c.cxl / c.fxl; c.cyli = -c.cyl / c.fyl;
    for (int i = 0; i < 4; i++) U.cDeltaF[i] = (float)(o.calib_value[i] - o.calib_zero[i]);
  }
  __syncthreads();
  if (tid == 0) {   // the convergence test (:281-301), float sums in frame order as written
    float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
    for (int h = 0; h < n; h++) {
      const double* st = s_step[h];
      sumA += st[6] * st[6];
      sumB += st[7] * st[7];
      sumT += st[0] * st[0] + st[1] * st[1] + st[2] * st[2];
      sumR += st[3] * st[3] + st[4] * st[4] + st[5] * st[5];
    }
    sumA /= n; sumB /= n; sumR /= n; sumT /= n;
    const float sumNID = U.P > 0 ? (float)(U.step_sums[1] / U.P) : 0.f;
    const float th = U.opt->th_opt;
    const bool canbreak = sqrtf(sumA) < 0.0005 * th && sqrtf(sumB) < 0.00005 * th && sqrtf(sumR) < 0.00005 * th && sqrtf(sumT) * sumNID < 0.00005 * th;
    U.opt->pending = (canbreak && U.opt->it >= U.opt->min_its) ? 1 : 0;
  }
  // FrameFramePrecalc::set at the current state + adHTdeltaF for every (host, target) pair
  for (int ht = tid; ht < n * n; ht += blockDim.x) {
    const int h = ht / n, t = ht % n;
    const FrameDev& host = U.frames[h]; const FrameDev& target = U.frames[t];
    PrecalcDev& p = U.precalc[(size_t)h * n + t];
    const BACalib& c = *U.calib;
    const float K[9] = {c.fxl, 0, c.cxl, 0, c.fyl, c.cyl, 0, 0, 1};
    float Kinv[9];
    dev_inverse3f(K, Kinv);
    double l[12];
    dev_se3_mul(target.T_w2c, host.T_c2w, l);
    for (int r = 0; r < 3; r++) {
      for (int q = 0; q < 3; q++) p.PRE_RTll[r * 3 + q] = (float)l[r * 4 + q];
      p.PRE_tTll[r] = (float)l[r * 4 + 3];
    }
    p.distanceLL = (float)sqrt(l[3] * l[3] + l[7] * l[7] + l[11] * l[11]);
    float KR[9];
    dev_mat33f_mul(K, p.PRE_RTll, KR);
    dev_mat33f_mul(KR, Kinv, p.PRE_KRKiTll);
    dev_mat33f_mul(p.PRE_RTll, Kinv, p.PRE_RKiTll);
    for (int r = 0; r < 3; r++) p.PRE_KtTll[r] = K[r * 3] * p.PRE_tTll[0] + K[r * 3 + 1] * p.PRE_tTll[1] + K[r * 3 + 2] * p.PRE_tTll[2];
    float eF = host.ab_exposure, eT = target.ab_exposure;
    if (eF == 0 || eT == 0) { eT = eF = 1; }
    const double a = exp(SCALE_A * target.state[6] - SCALE_A * host.state[6]) * eT / eF;   // AffLight::fromToVecExposure on state_scaled
    p.PRE_aff_mode[0] = (float)a; p.PRE_aff_mode[1] = (float)(SCALE_B * target.state[7] - a * (SCALE_B * host.state[7]));
    // setDeltaF (:181-192): adHTdeltaF[h + t n] = delta_h^T adHostF + delta_t^T adTargetF
    const size_t idx = (size_t)h + (size_t)t * n;
    float dh[8], dt[8];
    for (int i = 0; i < 8; i++) { dh[i] = (float)(host.state[i] - host.state_zero[i]); dt[i] = (float)(target.state[i] - target.state_zero[i]); }
    for (int j = 0; j < 8; j++) {
      float a2 = 0, b2 = 0;
      for (int i = 0; i < 8; i++) a2 += dh[i] * U.adHostF[idx * 64 + i * 8 + j];
      for (int i = 0; i < 8; i++) b2 += dt[i] * U.adTargetF[idx * 64 + i * 8 + j];
      U.adHTdeltaF[idx * 8 + j] = a2 + b2;
    }
  }
}
// end of one LM iteration: count it, and latch the convergence decision so that the iterations the host enqueued behind this one
// become no-ops
__global__ void ba_iter_end_kernel(OptDev* o) {
  if (o->done) return;
  o->it += 1;
  o->its_done = o->it;
  if (o->pending) o->done = 1;
}

// ---- g2o LBA driver, per-trial kernels (sdso_lba_g2o) ------------------------------------------------------------------------------
// slots of the driver's scalar block (doubles)
enum { LS_CUR = 0, LS_SL = 1, LS_TEMP = 2, LS_SCALE = 3, LS_OK = 4, LS_POST = 5, LS_NUM = 8 };

// Schur complement terms of one chunk for the current lambda: sum_e hpl_e hpl_e^T / (hll_e + lambda) (78 values, upper triangle by
// rows) and sum_e hpl_e bl_e / (hll_e + lambda) (12). The chunk's edges are staged once in shared memory; thread t < 90 then sums
// ITS output element over the edges in slot order (fixed order, no barriers inside the sum). Replaces the 13-phase
// lba_build_kernel(what = 1) on the per-trial path (one launch per damping trial).
__global__ void __launch_bounds__(kChunk) lba_schur_kernel(BAView B, LBAGraph G, double lambda, double* part /* [nchunks][96] */) {
  __shared__ double hs[kChunk][13];   // hpl (12) and bl; odd stride: a warp's staging stores hit distinct banks
  __shared__ double wv[kChunk];       // 1 / (hll + lambda), 0 for inactive edges
  const Chunk ch = B.chunks[blockIdx.x];
  const int cnt = ch.end - ch.begin;
  const int tid = threadIdx.x;
  if (tid < cnt) {
    const int s = ch.begin + tid;
    const bool use = G.active[s] != 0;
    wv[tid] = use ? 1.0 / (G.hll[s] + lambda) : 0.0;
#pragma unroll
    for (int a = 0; a < 12; a++) hs[tid][a] = use ? G.hpl[(size_t)s * 12 + a] : 0.0;
    hs[tid][12] = use ? G.bl[s] : 0.0;
  }
  __syncthreads();
  if (tid < 90) {
    int a = 0, c = 12;
    if (tid < 78) { int rem = tid; while (rem >= 12 - a) { rem -= 12 - a; a++; } c = a + rem; }
    else a = tid - 78;
    double t = 0;
    for (int e = 0; e < cnt; e++) t += hs[e][a] * wv[e] * hs[e][c];
    part[(size_t)blockIdx.x * 96 + tid] = t;
  }
}

// One damping trial's vertex update on the device (the host used to read x back for this): push() of the pose / photometric /
// camera vertices, x finite?, oplus of every used host (dso_g2o_vertex.cpp:15-18, 30-40, 100-106) and the vertex part of
// computeScale, sum_k x_k (lambda x_k + b_k) with b the un-reduced gradient of the non-marginalised block.
// est = [T_wh n*12 | T_tw n*12 | photo n*2 | b0 n | target aff n*2 | cam 4]; bak = [T_wh n*12 | photo n*2 | cam 4]
__global__ void lba_trial_update_kernel(int n, const int* used, const double* x, double lambda, const double* hostA /* [n][96] */, double* est,
                                        double* bak, double* sc) {
  __shared__ int ok_s;
  const int h = threadIdx.x, d = kCPARS + 8 * n;
  double* T_wh = est; double* photo = est + (size_t)n * 24; double* cam = est + (size_t)n * 29;
  if (h == 0) {
    bool ok = true;
    for (int k = 0; k < d; k++) ok = ok && isfinite(x[k]);
    ok_s = ok ? 1 : 0;
    sc[LS_OK] = ok ? 1.0 : 0.0;
    for (int k = 0; k < 4; k++) bak[(size_t)n * 14 + k] = cam[k];
    double scale = 0;
    if (ok) {
      // b of the cam block: sum over the used hosts; of a pose block: that host's own 8 entries
      for (int k = 0; k < 4; k++) {
        double bp = 0;
        for (int g = 0; g < n; g++) if (used[g]) bp += hostA[(size_t)g * 96 + 78 + 8 + k];
        scale += x[k] * (lambda * x[k] + bp);
      }
      for (int g = 0; g < n; g++)
        for (int l = 0; l < 8; l++) {
          const double xk = x[kCPARS + 8 * g + l];
          const double bp = used[g] ? hostA[(size_t)g * 96 + 78 + l] : 0.0;
          scale += xk * (lambda * xk + bp);
        }
      for (int k = 0; k < 4; k++) cam[k] += x[k];
    }
    sc[LS_SCALE] = scale;
  }
  __syncthreads();
  if (h < n) {
    for (int k = 0; k < 12; k++) bak[(size_t)h * 12 + k] = T_wh[(size_t)h * 12 + k];
    bak[(size_t)n * 12 + 2 * h] = photo[2 * h]; bak[(size_t)n * 12 + 2 * h + 1] = photo[2 * h + 1];
    if (ok_s && used[h]) {
      double Ex[12], Tn[12];
      dev_se3_exp(x + kCPARS + 8 * h, Ex);
      dev_se3_mul(Ex, T_wh + (size_t)h * 12, Tn);
      for (int k = 0; k < 12; k++) T_wh[(size_t)h * 12 + k] = Tn[k];
      photo[2 * h] += x[kCPARS + 8 * h + 6]; photo[2 * h + 1] += x[kCPARS + 8 * h + 7];
    }
  }
}
// pop(): vertices and inverse depths back to the pushed values
__global__ void lba_pop_kernel(int n, int R, const double* bak, double* est, const double* idbak, double* idepth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) idepth[i] = idbak[i];
  if (i < n * 12) est[i] = bak[i];
  if (i < n * 2) est[(size_t)n * 24 + i] = bak[(size_t)n * 12 + i];
  if (i < 4) est[(size_t)n * 29 + i] = bak[(size_t)n * 14 + i];
}

}  // namespace sdso
