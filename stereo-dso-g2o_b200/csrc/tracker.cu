// A3–A7, E1 — CoarseTracker on the device (FullSystem/CoarseTracker.cpp, dso_g2o_edge.cpp:395-500).
//
// B200 design: the WHOLE coarse-to-fine optimisation of one (reference, new frame, initial pose)
// problem — every calcRes / calcGSSSE pass, the 8x8 solves, the SE3 updates, the LM accept/reject
// logic for all pyramid levels — runs inside ONE persistent kernel launch. One thread-block cluster
// owns one problem; its CTAs split the points, reduce per-CTA partial sums in shared memory, exchange
// them through distributed shared memory (one cluster barrier per evaluation, double-buffered) and
// every CTA redundantly runs the (deterministic, fixed-order) final sum + solve, so no broadcast and
// no host round trip is needed. Independent problems (motion hypotheses, FullSystem.cpp:351-376, or
// independent sequences) are additional clusters of the same launch.
//
// Per point ("eval"): 16 B point record + 4 x 16 B texel gather, projection, residual, Huber weight,
// 9-vector Jacobian row, 45-term weighted outer product — calcRes (:600-792) and calcGSSSE (:537-596)
// fused, so the eight buf_warped_* arrays never exist in memory.
#include "ctx.h"
#include "tracker_state.h"
#include <cooperative_groups.h>
#include <cstring>
#include <cmath>

namespace cg = cooperative_groups;

namespace sdso {

// ------------------------------------------------------------------------------------------------
constexpr int kAcc = 52;      // accumulators per thread (see enum)
enum { A_H = 0, A_E = 45, A_NE = 46, A_NSAT = 47, A_ST = 48, A_SRT = 49, A_SN = 50, A_NW = 51 };

struct TrackLevel {
  const float4* pc;  // {u, v, idepth, color}  (CoarseTracker.h:117-121 pc_u/pc_v/pc_idepth/pc_color fused)
  int n;
  int w, h;
  float fx, fy, cx, cy;  // tracker's K (optimised HCalib), CoarseTracker.cpp:108-136
  float Ki[9];
  double gfx, gfy, gcx, gcy;  // global initial KG[lvl] used by EdgeSE3PosePhotoDSO (dso_util.hpp:10-22)
};

struct TrackProblem {  // one per cluster, in global memory
  // inputs
  double T[12];
  double aff[2];
  double minResForAbort[5];
  const float4* tex[kPyrLevels];  // new frame pyramid
  float exposure_new;
  int pad0;
  // outputs
  double T_out[12];
  double aff_out[2];
  double lastResiduals[5];
  double flow[3];
  int iterations[5];
  int ok;
  unsigned long long evals;
  // single-eval mode outputs
  double rs[6];
  double H[64];
  double b[8];
  int warped_n;
  int pad1;
};

struct TrackParams {
  TrackLevel L[kPyrLevels];
  int levels;
  int coarsest;
  int variant;
  int mode;  // 0 = full track, 1 = single fused calcRes+calcGS at level eval_lvl
  int eval_lvl;
  float eval_cutoff;
  float* dump;          // mode 1: per-point records [9][n] (valid,idepth,u,v,dx,dy,residual,weight,refColor) or null
  double* dump_d;       // mode 2: per-point records [10][n] (flag, err, J[8]) or null
  float ref_exposure;
  double ref_aff[2];    // lastRef_aff_g2l
  float huberTH, coarseCutoffTH;
  float affineOptModeA, affineOptModeB;
  int g2o_stop_persists;
  TrackProblem* problems;
  // g2o variant scratch: per level edge flags/errors for each problem
  unsigned char* edge_flag[kPyrLevels];  // [problem][n_l]
  double* edge_err[kPyrLevels];          // [problem][n_l]
  int edge_stride[kPyrLevels];
};

struct EvalConst {  // per-evaluation constants, computed by thread 0 of each CTA
  float RKi[9];
  float t[3];
  float affLL[2];
  float a;       // (float) fromToVecExposure(...)[0]      (calcGSSSE :544)
  float b0;      // lastRef_aff_g2l.b
  float cutoff, maxEnergy;
  // g2o variant: pose in double and photo ab
  double R[9], tt[3];
  float ab[2];
  double b0d;
};

struct LMState {  // lives in shared memory of every CTA (identical content everywhere)
  double R[9], t[3];      // refToNew_current
  double aff[2];
  double Rn[9], tn[3];    // trial
  double affn[2];
  double H[64], b[8];
  double inc[8];
  double total[kAcc];
  double totalNew[kAcc];
  float lambda;
  int flag;
};

// ---- small double helpers (device) -------------------------------------------------------------
__device__ void d_mat3_mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// exp(a) * T   with a = [upsilon; omega]   (thirdparty/Sophus/sophus/se3.hpp:407-428; left-multiplicative
// update as CoarseTracker.cpp:978 / dso_g2o_vertex.cpp:17)
__device__ void d_se3_exp_mul(const double a[6], const double* R, const double* t, double* Ro, double* to) {
  const double wx = a[3], wy = a[4], wz = a[5];
  const double th2 = wx * wx + wy * wy + wz * wz;
  const double th = sqrt(th2);
  double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  d_mat3_mul(O, O, O2);
  double A, B, C;  // R = I + A O + B O2 ; V = I + B O + C O2
  if (th < 1e-10) { A = 1.0; B = 0.5; C = 1.0 / 6.0; }
  else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th); }
  double Re[9], V[9];
  for (int i = 0; i < 9; i++) {
    double I = (i % 4 == 0) ? 1.0 : 0.0;
    Re[i] = I + A * O[i] + B * O2[i];
    V[i] = I + B * O[i] + C * O2[i];
  }
  double te[3];
  for (int r = 0; r < 3; r++) te[r] = V[r * 3] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  double Rt[9];
  d_mat3_mul(Re, R, Rt);
  double tn[3];
  for (int r = 0; r < 3; r++) tn[r] = Re[r * 3] * t[0] + Re[r * 3 + 1] * t[1] + Re[r * 3 + 2] * t[2] + te[r];
  for (int i = 0; i < 9; i++) Ro[i] = Rt[i];
  for (int i = 0; i < 3; i++) to[i] = tn[i];
}

// LDLT solve (no pivoting; the systems on this path are SPD after damping). Returns false on breakdown.
template <int NMAX>
__device__ bool d_ldlt_solve(int n, const double* A, int lda, const double* b, double* x) {
  double L[NMAX * NMAX], d[NMAX];
  for (int j = 0; j < n; j++) {
    double s = A[j * lda + j];
    for (int k = 0; k < j; k++) s -= L[j * NMAX + k] * L[j * NMAX + k] * d[k];
    d[j] = s;
    double inv = 1.0 / s;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * lda + j];
      for (int k = 0; k < j; k++) v -= L[i * NMAX + k] * L[j * NMAX + k] * d[k];
      L[i * NMAX + j] = v * inv;
    }
  }
  double y[NMAX];
  for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * NMAX + k] * y[k]; y[i] = v; }
  for (int i = 0; i < n; i++) y[i] /= d[i];
  for (int i = n - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < n; k++) v -= L[k * NMAX + i] * x[k]; x[i] = v; }
  bool ok = true;
  for (int i = 0; i < n; i++) ok = ok && isfinite(x[i]);
  return ok;
}

// AffLight::fromToVecExposure (util/NumType.h:159-170)
__device__ __forceinline__ void d_aff_from_to(float expF, float expT, double aF, double bF, double aT, double bT, double out[2]) {
  if (expF == 0 || expT == 0) { expT = expF = 1; }
  double a = exp(aT - aF) * expT / expF;
  out[0] = a;
  out[1] = bT - a * bF;
}

// ---- shared memory layout ------------------------------------------------------------------------
struct __align__(16) TrackSmem {
  LMState lm;
  EvalConst ec;
  float partial[2][kAcc];  // this CTA's block-reduced sums, double-buffered by evaluation parity
  double partial_d[2];     // one extra channel summed in double (robust chi2 of the g2o path)
  double warp_d[32];
  // followed by float red[kAcc * blockDim.x]
};

__device__ __forceinline__ float* smem_red(TrackSmem* sm) { return reinterpret_cast<float*>(sm + 1); }

// Block + cluster reduction of per-thread accumulators; result (identical in every CTA) -> out[kAcc]
__device__ void reduce_all(float (&acc)[kAcc], TrackSmem* sm, int& parity, cg::cluster_group& cluster, double* out,
                           double dacc = 0.0, double* dout = nullptr) {
  float* red = smem_red(sm);
  const int BT = blockDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nw = BT >> 5;
#pragma unroll
  for (int k = 0; k < kAcc; k++) red[k * BT + tid] = acc[k];
  if (dout) {
    double dv = warp_sum(dacc);
    if (lane == 0) sm->warp_d[warp] = dv;
  }
  __syncthreads();
  if (dout && tid == 0) {
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += sm->warp_d[w];
    sm->partial_d[parity] = s;
  }
  for (int k = warp; k < kAcc; k += nw) {
    float s = 0.f;
    for (int j = lane; j < BT; j += 32) s += red[k * BT + j];
    s = warp_sum(s);
    if (lane == 0) sm->partial[parity][k] = s;
  }
  cluster.sync();  // partials of all CTAs visible (release/acquire at cluster scope)
  const unsigned C = cluster.num_blocks();
  if (tid < kAcc) {
    double s = 0.0;
    for (unsigned r = 0; r < C; r++) {
      const TrackSmem* rs = cluster.map_shared_rank(sm, r);
      s += (double)rs->partial[parity][tid];
    }
    out[tid] = s;
  }
  if (dout && tid == kAcc) {
    double s = 0.0;
    for (unsigned r = 0; r < C; r++) s += cluster.map_shared_rank(sm, r)->partial_d[parity];
    *dout = s;
  }
  parity ^= 1;
  __syncthreads();
}

// ---- SSE-path evaluation: calcRes (:645-773) + calcGSSSE (:537-596) fused over this CTA's points ----
__device__ void eval_points_sse(const TrackParams& P, const TrackLevel& L, int lvl, const float4* __restrict__ tex,
                                const EvalConst& ec, float (&acc)[kAcc], unsigned& evals, int gtid, int gthreads,
                                float* dump) {
#pragma unroll
  for (int k = 0; k < kAcc; k++) acc[k] = 0.f;
  const float fxl = L.fx, fyl = L.fy, cxl = L.cx, cyl = L.cy;
  const int wl = L.w, hl = L.h;
  const float huberTH = P.huberTH;
  const float4* __restrict__ pc = L.pc;
  for (int i = gtid; i < L.n; i += gthreads) {
    const float4 p = __ldg(pc + i);
    const float x = p.x, y = p.y, id = p.z, refColor = p.w;
    float pt[3];
#pragma unroll
    for (int r = 0; r < 3; r++) pt[r] = (ec.RKi[r * 3 + 0] * x + ec.RKi[r * 3 + 1] * y + ec.RKi[r * 3 + 2]) + ec.t[r] * id;
    const float u = pt[0] / pt[2], v = pt[1] / pt[2];
    const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    const float new_idepth = id / pt[2];
    evals++;
    if (lvl == 0 && (i % 32) == 0) {  // flow indicators :662-693
      float ptT[3], ptT2[3], pt3[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        float kp = L.Ki[r * 3 + 0] * x + L.Ki[r * 3 + 1] * y + L.Ki[r * 3 + 2];
        ptT[r] = kp + ec.t[r] * id;
        ptT2[r] = kp - ec.t[r] * id;
        pt3[r] = (ec.RKi[r * 3 + 0] * x + ec.RKi[r * 3 + 1] * y + ec.RKi[r * 3 + 2]) - ec.t[r] * id;
      }
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      acc[A_ST] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      acc[A_ST] += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      acc[A_SRT] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      acc[A_SRT] += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      acc[A_SN] += 2.f;
    }
    bool valid = false;
    float hitx = 0.f, hity = 0.f, hitz = 0.f, residual = 0.f, hw = 0.f;
    if (Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0) {  // :696
      const float3 hit = interp33(tex, Ku, Kv, wl);
      if (isfinite(hit.x)) {
        hitx = hit.x; hity = hit.y; hitz = hit.z;
        residual = hit.x - (ec.affLL[0] * refColor + ec.affLL[1]);
        const float ar = fabsf(residual);
        hw = ar < huberTH ? 1.f : huberTH / ar;
        if (ar > ec.cutoff) {
          acc[A_E] += ec.maxEnergy; acc[A_NE] += 1.f; acc[A_NSAT] += 1.f;
        } else {
          acc[A_E] += hw * residual * residual * (2 - hw);
          acc[A_NE] += 1.f; acc[A_NW] += 1.f;
          valid = true;
        }
      }
    }
    if (valid) {
      // calcGSSSE :553-581
      const float dx = hity * fxl, dy = hitz * fyl;
      float J[9];
      J[0] = new_idepth * dx;
      J[1] = new_idepth * dy;
      J[2] = 0.f - new_idepth * (u * dx + v * dy);
      J[3] = 0.f - ((u * v) * dx + dy * (1.f + v * v));
      J[4] = (u * v) * dy + dx * (1.f + u * u);
      J[5] = u * dy - v * dx;
      J[6] = ec.a * (ec.b0 - refColor);
      J[7] = -1.f;
      J[8] = residual;
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const float Jw = J[r] * hw;
#pragma unroll
        for (int c = r; c < 9; c++) { acc[A_H + idx] = __fmaf_rn(Jw, J[c], acc[A_H + idx]); idx++; }
      }
    }
    if (dump) {
      const int n = L.n;
      dump[0 * n + i] = valid ? 1.f : 0.f;
      dump[1 * n + i] = new_idepth; dump[2 * n + i] = u; dump[3 * n + i] = v;
      dump[4 * n + i] = hity; dump[5 * n + i] = hitz; dump[6 * n + i] = residual;
      dump[7 * n + i] = hw; dump[8 * n + i] = refColor;
    }
  }
}

// thread 0: per-evaluation constants for the SSE path (calcRes :617-621, calcGSSSE :540-544)
__device__ void make_eval_const_sse(const TrackParams& P, const TrackLevel& L, const TrackProblem& prob, const double* R,
                                    const double* t, const double* aff, float cutoff, EvalConst& ec) {
  float Rf[9];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) ec.RKi[r * 3 + c] = Rf[r * 3 + 0] * L.Ki[0 * 3 + c] + Rf[r * 3 + 1] * L.Ki[1 * 3 + c] + Rf[r * 3 + 2] * L.Ki[2 * 3 + c];
  for (int i = 0; i < 3; i++) ec.t[i] = (float)t[i];
  double ab[2];
  d_aff_from_to(P.ref_exposure, prob.exposure_new, P.ref_aff[0], P.ref_aff[1], aff[0], aff[1], ab);
  ec.affLL[0] = (float)ab[0]; ec.affLL[1] = (float)ab[1];
  ec.a = (float)ab[0];
  ec.b0 = (float)P.ref_aff[1];
  ec.cutoff = cutoff;
  ec.maxEnergy = 2 * P.huberTH * cutoff - P.huberTH * P.huberTH;
}

// H (8x8), b from the 45 accumulated entries: calcGSSSE :582-595
__device__ void finish_gs(const double* total, double* H, double* b) {
  const int nw = (int)total[A_NW];
  const int n = (nw + 3) & ~3;  // buf_warped_n is padded to a multiple of 4 (:763-773)
  const float invn = 1.0f / n;
  double M[81];
  int idx = 0;
  for (int r = 0; r < 9; r++)
    for (int c = r; c < 9; c++) { float d = (float)total[A_H + idx]; M[r * 9 + c] = M[c * 9 + r] = (double)d; idx++; }
  const double sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_A, SCALE_B};
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H[r * 8 + c] = M[r * 9 + c] * invn * sc[c] * sc[r];
    b[r] = M[r * 9 + 8] * invn * sc[r];
  }
}

__device__ void rs_from_total(const double* total, double rs[6]) {
  // CoarseTracker.cpp:783-789 (float arithmetic as written)
  const float E = (float)total[A_E];
  const int numTermsInE = (int)total[A_NE];
  const int numSaturated = (int)total[A_NSAT];
  const float sT = (float)total[A_ST], sRT = (float)total[A_SRT], sN = (float)total[A_SN];
  rs[0] = E;
  rs[1] = numTermsInE;
  rs[2] = sT / (sN + 0.1);
  rs[3] = 0;
  rs[4] = sRT / (sN + 0.1);
  rs[5] = numSaturated / (float)numTermsInE;
}

// thread 0: one LM step of the SSE path (commented block CoarseTracker.cpp:929-979) -> trial state
__device__ void lm_step_sse(const TrackParams& P, LMState& lm) {
  double Hl[64];
  for (int i = 0; i < 64; i++) Hl[i] = lm.H[i];
  for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lm.lambda);
  double nb[8];
  for (int i = 0; i < 8; i++) nb[i] = -lm.b[i];
  double inc[8];
  const bool fixA = P.affineOptModeA < 0, fixB = P.affineOptModeB < 0;
  if (!fixA && !fixB) {
    d_ldlt_solve<8>(8, Hl, 8, nb, inc);
  } else if (fixA && fixB) {
    d_ldlt_solve<8>(6, Hl, 8, nb, inc);
    inc[6] = inc[7] = 0;
  } else if (!fixA && fixB) {
    d_ldlt_solve<8>(7, Hl, 8, nb, inc);
    inc[7] = 0;
  } else {  // fix a, optimise b: stitch row/col 7 into 6 (:947-964)
    double Hs[64], bs[8];
    for (int i = 0; i < 64; i++) Hs[i] = Hl[i];
    for (int i = 0; i < 8; i++) bs[i] = lm.b[i];
    for (int r = 0; r < 8; r++) Hs[r * 8 + 6] = Hs[r * 8 + 7];
    for (int c = 0; c < 8; c++) Hs[6 * 8 + c] = Hs[7 * 8 + c];
    bs[6] = bs[7];
    double nbs[8], is[8];
    for (int i = 0; i < 8; i++) nbs[i] = -bs[i];
    d_ldlt_solve<8>(7, Hs, 8, nbs, is);
    for (int i = 0; i < 6; i++) inc[i] = is[i];
    inc[6] = 0; inc[7] = is[6];
  }
  float extrapFac = 1;
  const float lambdaExtrapolationLimit = 0.001f;
  if (lm.lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lm.lambda));
  for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
  for (int i = 0; i < 8; i++) lm.inc[i] = inc[i];
  double s[8];
  for (int i = 0; i < 8; i++) s[i] = inc[i];
  for (int i = 0; i < 3; i++) s[i] *= SCALE_XI_ROT;
  for (int i = 3; i < 6; i++) s[i] *= SCALE_XI_TRANS;
  s[6] *= SCALE_A; s[7] *= SCALE_B;
  double sum = 0;
  for (int i = 0; i < 8; i++) sum += s[i];
  if (!isfinite(sum)) for (int i = 0; i < 8; i++) s[i] = 0;
  d_se3_exp_mul(s, lm.R, lm.t, lm.Rn, lm.tn);
  lm.affn[0] = lm.aff[0] + s[6];
  lm.affn[1] = lm.aff[1] + s[7];
}

// ================================================================================================
__global__ void __launch_bounds__(256, 1) track_kernel(TrackParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackSmem* sm = reinterpret_cast<TrackSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const int prob_id = blockIdx.x / C;
  TrackProblem& prob = P.problems[prob_id];
  const int tid = threadIdx.x;
  const int gtid = rank * blockDim.x + tid, gthreads = C * blockDim.x;
  LMState& lm = sm->lm;
  float acc[kAcc];
  unsigned evals = 0;
  int parity = 0;

  if (tid == 0) {
    for (int i = 0; i < 9; i++) lm.R[i] = prob.T[(i / 3) * 4 + (i % 3)];
    for (int i = 0; i < 3; i++) lm.t[i] = prob.T[i * 4 + 3];
    lm.aff[0] = prob.aff[0]; lm.aff[1] = prob.aff[1];
  }
  __syncthreads();

  if (P.mode == 1) {  // ---- single fused calcRes + calcGSSSE (operator-level entry) ----
    const int lvl = P.eval_lvl;
    const TrackLevel& L = P.L[lvl];
    if (tid == 0) make_eval_const_sse(P, L, prob, lm.R, lm.t, lm.aff, P.eval_cutoff, sm->ec);
    __syncthreads();
    eval_points_sse(P, L, lvl, prob.tex[lvl], sm->ec, acc, evals, gtid, gthreads, P.dump);
    reduce_all(acc, sm, parity, cluster, lm.total);
    if (rank == 0 && tid == 0) {
      rs_from_total(lm.total, prob.rs);
      finish_gs(lm.total, prob.H, prob.b);
      const int nw = (int)lm.total[A_NW];
      prob.warped_n = (nw + 3) & ~3;
    }
    cluster.sync();
    return;
  }

  // ---- full coarse-to-fine tracking, SSE path (CoarseTracker.cpp:827-1069, commented control flow) ----
  const int maxIterations[5] = {10, 20, 50, 50, 50};
  bool haveRepeated = false;
  double lastRes[5] = {NAN, NAN, NAN, NAN, NAN};
  double flow[3] = {1000, 1000, 1000};
  int iters[5] = {0, 0, 0, 0, 0};
  bool aborted = false;

  for (int lvl = P.coarsest; lvl >= 0 && !aborted; lvl--) {
    const TrackLevel& L = P.L[lvl];
    const float4* tex = prob.tex[lvl];
    float levelCutoffRepeat = 1;
    // resOld = calcRes(...)
    if (tid == 0) make_eval_const_sse(P, L, prob, lm.R, lm.t, lm.aff, P.coarseCutoffTH * levelCutoffRepeat, sm->ec);
    __syncthreads();
    eval_points_sse(P, L, lvl, tex, sm->ec, acc, evals, gtid, gthreads, nullptr);
    reduce_all(acc, sm, parity, cluster, lm.total);
    double rsOld[6];
    rs_from_total(lm.total, rsOld);
    while (rsOld[5] > 0.6 && levelCutoffRepeat < 50) {  // :897-904
      levelCutoffRepeat *= 2;
      __syncthreads();
      if (tid == 0) make_eval_const_sse(P, L, prob, lm.R, lm.t, lm.aff, P.coarseCutoffTH * levelCutoffRepeat, sm->ec);
      __syncthreads();
      eval_points_sse(P, L, lvl, tex, sm->ec, acc, evals, gtid, gthreads, nullptr);
      reduce_all(acc, sm, parity, cluster, lm.total);
      rs_from_total(lm.total, rsOld);
    }
    if (tid == 0) { finish_gs(lm.total, lm.H, lm.b); lm.lambda = 0.01f; }
    __syncthreads();

    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      iters[lvl]++;
      if (tid == 0) {
        lm_step_sse(P, lm);
        make_eval_const_sse(P, L, prob, lm.Rn, lm.tn, lm.affn, P.coarseCutoffTH * levelCutoffRepeat, sm->ec);
      }
      __syncthreads();
      eval_points_sse(P, L, lvl, tex, sm->ec, acc, evals, gtid, gthreads, nullptr);
      reduce_all(acc, sm, parity, cluster, lm.totalNew);
      double rsNew[6];
      rs_from_total(lm.totalNew, rsNew);
      const bool accept = (rsNew[0] / rsNew[1]) < (rsOld[0] / rsOld[1]);  // :989
      double nrm = 0;
      for (int i = 0; i < 8; i++) nrm += lm.inc[i] * lm.inc[i];
      nrm = sqrt(nrm);
      __syncthreads();  // everyone has read lm.inc / totals before thread 0 mutates the state
      if (accept) {
        for (int i = 0; i < 6; i++) rsOld[i] = rsNew[i];
        if (tid == 0) {
          finish_gs(lm.totalNew, lm.H, lm.b);
          for (int i = 0; i < 9; i++) lm.R[i] = lm.Rn[i];
          for (int i = 0; i < 3; i++) lm.t[i] = lm.tn[i];
          lm.aff[0] = lm.affn[0]; lm.aff[1] = lm.affn[1];
          lm.lambda *= 0.5f;
        }
      } else if (tid == 0) {
        lm.lambda *= 4;
        if (lm.lambda < 0.001f) lm.lambda = 0.001f;
      }
      __syncthreads();
      if (!(nrm > 1e-3)) break;  // :1019
    }
    lastRes[lvl] = sqrtf((float)(rsOld[0] / rsOld[1]));  // :1028
    flow[0] = rsOld[2]; flow[1] = rsOld[3]; flow[2] = rsOld[4];
    if (lastRes[lvl] > 1.5 * prob.minResForAbort[lvl]) { aborted = true; break; }
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }

  // outputs (:1044-1068)
  if (rank == 0 && tid == 0) {
    bool ok = !aborted;
    double aout[2] = {lm.aff[0], lm.aff[1]};
    if (ok) {
      if ((P.affineOptModeA != 0 && (fabsf((float)aout[0]) > 1.2)) || (P.affineOptModeB != 0 && (fabsf((float)aout[1]) > 200))) ok = false;
    }
    if (ok) {
      double rel[2];
      d_aff_from_to(P.ref_exposure, prob.exposure_new, P.ref_aff[0], P.ref_aff[1], aout[0], aout[1], rel);
      const float r0 = (float)rel[0], r1 = (float)rel[1];
      if ((P.affineOptModeA == 0 && (fabsf(logf(r0)) > 1.5)) || (P.affineOptModeB == 0 && (fabsf(r1) > 200))) ok = false;
    }
    if (ok) {
      if (P.affineOptModeA < 0) aout[0] = 0;
      if (P.affineOptModeB < 0) aout[1] = 0;
    }
    if (!aborted) {
      for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) prob.T_out[r * 4 + c] = lm.R[r * 3 + c]; prob.T_out[r * 4 + 3] = lm.t[r]; }
      prob.aff_out[0] = aout[0]; prob.aff_out[1] = aout[1];
    } else {  // the reference returns before writing lastToNew_out / aff_g2l_out (:1032-1033)
      for (int i = 0; i < 12; i++) prob.T_out[i] = prob.T[i];
      prob.aff_out[0] = prob.aff[0]; prob.aff_out[1] = prob.aff[1];
    }
    for (int i = 0; i < 5; i++) { prob.lastResiduals[i] = lastRes[i]; prob.iterations[i] = iters[i]; }
    for (int i = 0; i < 3; i++) prob.flow[i] = flow[i];
    prob.ok = ok ? 1 : 0;
  }
  // evals: integer sum, order-independent
  {
    unsigned e = evals;
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((tid & 31) == 0) atomicAdd(&prob.evals, (unsigned long long)e);
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

#include "tracker_g2o.cuh"

// ================================================================================================
// host side
int tracker_create(sdso_ctx* ctx) {
  TrackerState* t = new TrackerState();
  ctx->tracker = t;
  t->K = ctx->G;
  for (int l = 0; l < ctx->G.levels; l++) {
    size_t n = (size_t)ctx->G.w[l] * ctx->G.h[l];
    SDSO_CUDA(ctx, cudaMalloc(&t->pc[l], n * sizeof(float4)));
    t->pc_cap[l] = (int)n;
    SDSO_CUDA(ctx, cudaMalloc(&t->idepth[l], n * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&t->wsum[l], n * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&t->wsum_bak[l], n * sizeof(float)));
  }
  SDSO_CUDA(ctx, cudaMalloc(&t->scan_tmp, ((size_t)ctx->G.w[0] * ctx->G.h[0] + 1024) * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&t->d_counts, 64 * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&t->d_problems, t->max_problems * sizeof(TrackProblem)));
  SDSO_CUDA(ctx, cudaMallocHost(&t->h_problems, t->max_problems * sizeof(TrackProblem)));
  return SDSO_OK;
}

void tracker_destroy(sdso_ctx* ctx) {
  TrackerState* t = ctx->tracker;
  if (!t) return;
  for (int l = 0; l < kPyrLevels; l++) {
    if (t->pc[l]) cudaFree(t->pc[l]);
    if (t->idepth[l]) cudaFree(t->idepth[l]);
    if (t->wsum[l]) cudaFree(t->wsum[l]);
    if (t->wsum_bak[l]) cudaFree(t->wsum_bak[l]);
  }
  if (t->scan_tmp) cudaFree(t->scan_tmp);
  if (t->d_counts) cudaFree(t->d_counts);
  if (t->d_problems) cudaFree(t->d_problems);
  if (t->h_problems) cudaFreeHost(t->h_problems);
  if (t->d_dump) cudaFree(t->d_dump);
  for (int l = 0; l < kPyrLevels; l++) { if (t->edge_flag[l]) cudaFree(t->edge_flag[l]); if (t->edge_err[l]) cudaFree(t->edge_err[l]); }
  delete t;
  ctx->tracker = nullptr;
}

static void fill_params(sdso_ctx* ctx, TrackParams& P) {
  TrackerState* t = ctx->tracker;
  memset(&P, 0, sizeof(P));
  P.levels = ctx->G.levels;
  for (int l = 0; l < P.levels; l++) {
    TrackLevel& L = P.L[l];
    L.pc = t->pc[l]; L.n = t->pc_n[l];
    L.w = t->K.w[l]; L.h = t->K.h[l];
    L.fx = t->K.fx[l]; L.fy = t->K.fy[l]; L.cx = t->K.cx[l]; L.cy = t->K.cy[l];
    for (int i = 0; i < 9; i++) L.Ki[i] = t->K.Ki[l][i];
    L.gfx = ctx->G.K[l][0]; L.gfy = ctx->G.K[l][4]; L.gcx = ctx->G.K[l][2]; L.gcy = ctx->G.K[l][5];
  }
  P.ref_exposure = ctx->frames[t->ref_frame].ab_exposure;
  P.ref_aff[0] = t->ref_aff[0]; P.ref_aff[1] = t->ref_aff[1];
  P.huberTH = ctx->S.huberTH; P.coarseCutoffTH = ctx->S.coarseCutoffTH;
  P.affineOptModeA = ctx->S.affineOptModeA; P.affineOptModeB = ctx->S.affineOptModeB;
  P.g2o_stop_persists = ctx->S.g2o_stop_flag_persists;
  P.problems = t->d_problems;
}

static int launch_track(sdso_ctx* ctx, const TrackParams& P, int nb, bool g2o) {
  int C = ctx->S.cluster_size > 0 ? ctx->S.cluster_size : 8;
  int BT = ctx->S.block_threads > 0 ? ctx->S.block_threads : 256;
  if (BT > 256 || BT < 64 || (BT & 31)) return fail(ctx, SDSO_E_INVALID, "block_threads must be a multiple of 32 in [64,256]");
  if (C < 1 || C > 16) return fail(ctx, SDSO_E_INVALID, "cluster_size must be in [1,16]");
  size_t smem = sizeof(TrackSmem) + (size_t)kAcc * BT * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_g2o_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_g2o_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C * nb);
  cfg.blockDim = dim3(BT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (g2o) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_g2o_kernel, P));
  else SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_kernel, P));
  ctx->launches++;
  return SDSO_OK;
}

// per-problem, per-level edge flags / errors of the g2o path
static int ensure_edge_scratch(sdso_ctx* ctx, TrackParams& P, int nb) {
  TrackerState* t = ctx->tracker;
  for (int l = 0; l < ctx->G.levels; l++) {
    int stride = (t->pc_n[l] + 63) & ~63;
    if (stride < 64) stride = 64;
    size_t need = (size_t)stride * nb;
    if (need > t->edge_cap[l]) {
      if (t->edge_flag[l]) cudaFree(t->edge_flag[l]);
      if (t->edge_err[l]) cudaFree(t->edge_err[l]);
      t->edge_flag[l] = nullptr; t->edge_err[l] = nullptr; t->edge_cap[l] = 0;
      SDSO_CUDA(ctx, cudaMalloc(&t->edge_flag[l], need));
      SDSO_CUDA(ctx, cudaMalloc(&t->edge_err[l], need * sizeof(double)));
      t->edge_cap[l] = need;
    }
    P.edge_flag[l] = t->edge_flag[l]; P.edge_err[l] = t->edge_err[l]; P.edge_stride[l] = stride;
  }
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_tracker_make_k(sdso_ctx* ctx, const float K[4]) {
  if (!ctx || !K) return SDSO_E_INVALID;
  ctx->tracker->K.set(ctx->G.w[0], ctx->G.h[0], K[0], K[1], K[2], K[3], false);
  ctx->tracker->K.levels = ctx->G.levels;
  return SDSO_OK;
}

int sdso_tracker_set_pc(sdso_ctx* ctx, int ref_frame, int lvl, int n, const float* u, const float* v, const float* idepth,
                        const float* color, const double ref_aff[2]) {
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels || n < 0) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (ref_frame < 0 || ref_frame >= (int)ctx->frames.size() || !ctx->frames[ref_frame].in_use) return fail(ctx, SDSO_E_INVALID, "bad ref_frame");
  if (n > t->pc_cap[lvl]) return fail(ctx, SDSO_E_INVALID, "pc larger than level");
  std::vector<float4> tmp(n);
  for (int i = 0; i < n; i++) tmp[i] = make_float4(u[i], v[i], idepth[i], color[i]);
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->pc[lvl], tmp.data(), n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->pc_n[lvl] = n;
  t->ref_frame = ref_frame;
  t->ref_aff[0] = ref_aff[0]; t->ref_aff[1] = ref_aff[1];
  t->have_ref = true;
  return SDSO_OK;
}

int sdso_tracker_get_pc(sdso_ctx* ctx, int lvl, int* n, float* u, float* v, float* idepth, float* color) {
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels || !n) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  *n = t->pc_n[lvl];
  std::vector<float4> tmp(*n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(tmp.data(), t->pc[lvl], (size_t)*n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < *n; i++) {
    if (u) u[i] = tmp[i].x;
    if (v) v[i] = tmp[i].y;
    if (idepth) idepth[i] = tmp[i].z;
    if (color) color[i] = tmp[i].w;
  }
  return SDSO_OK;
}

static int check_frame(sdso_ctx* ctx, int f) {
  if (f < 0 || f >= (int)ctx->frames.size() || !ctx->frames[f].in_use || !ctx->frames[f].valid) return fail(ctx, SDSO_E_INVALID, "bad frame id (not created or makeImages not run)");
  return SDSO_OK;
}

int sdso_calc_res_gs(sdso_ctx* ctx, int new_frame, int lvl, const double refToNew[12], const double aff[2], float cutoffTH,
                     double rs[6], double H[64], double b[8], int* warped_n, float* warped) {
  if (!ctx || !refToNew || !aff) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!t->have_ref) return fail(ctx, SDSO_E_STATE, "calcRes before setCoarseTrackingRef");
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  int rc = check_frame(ctx, new_frame);
  if (rc) return rc;
  TrackParams P;
  fill_params(ctx, P);
  P.mode = 1; P.eval_lvl = lvl; P.eval_cutoff = cutoffTH; P.variant = SDSO_VARIANT_SSE;
  const int n = t->pc_n[lvl];
  if (warped) {
    size_t need = (size_t)9 * (n > 0 ? n : 1) * sizeof(float);
    if (need > t->dump_cap) {
      if (t->d_dump) cudaFree(t->d_dump);
      SDSO_CUDA(ctx, cudaMalloc(&t->d_dump, need));
      t->dump_cap = need;
    }
    P.dump = t->d_dump;
  }
  TrackProblem& hp = t->h_problems[0];
  memset(&hp, 0, sizeof(hp));
  memcpy(hp.T, refToNew, sizeof(hp.T));
  hp.aff[0] = aff[0]; hp.aff[1] = aff[1];
  for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frame].tex[l];
  hp.exposure_new = ctx->frames[new_frame].ab_exposure;
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_problems, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_track(ctx, P, 1, false);
  if (rc) return rc;
  SDSO_CUDA(ctx, cudaMemcpyAsync(&hp, t->d_problems, sizeof(hp), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<float> hd;
  if (warped) {
    hd.resize((size_t)9 * n);
    SDSO_CUDA(ctx, cudaMemcpyAsync(hd.data(), t->d_dump, hd.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (rs) memcpy(rs, hp.rs, sizeof(hp.rs));
  if (H) memcpy(H, hp.H, sizeof(hp.H));
  if (b) memcpy(b, hp.b, sizeof(hp.b));
  if (warped_n) *warped_n = hp.warped_n;
  if (warped) {
    // stable compaction in point order = the order calcRes fills buf_warped_* (:750-757), zero padded (:763-773)
    const int wn = hp.warped_n;
    int k = 0;
    for (int i = 0; i < n; i++) {
      if (hd[i] != 0.f) {
        for (int a = 0; a < 8; a++) warped[(size_t)a * wn + k] = hd[(size_t)(a + 1) * n + i];
        k++;
      }
    }
    for (; k < wn; k++) for (int a = 0; a < 8; a++) warped[(size_t)a * wn + k] = 0.f;
  }
  return SDSO_OK;
}

int sdso_edge_eval(sdso_ctx* ctx, int new_frame, int lvl, const double T_select[12], const double T_pose[12], const double photo[2],
                   int* n_out, double* err, double* J8) {
  if (!ctx || !T_select || !T_pose || !photo || !n_out) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!t->have_ref) return fail(ctx, SDSO_E_STATE, "edge evaluation before setCoarseTrackingRef");
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  int rc = check_frame(ctx, new_frame);
  if (rc) return rc;
  TrackParams P;
  fill_params(ctx, P);
  P.mode = 2; P.eval_lvl = lvl; P.eval_cutoff = 1e30f; P.variant = SDSO_VARIANT_G2O;
  rc = ensure_edge_scratch(ctx, P, 1);
  if (rc) return rc;
  const int n = t->pc_n[lvl];
  size_t need = (size_t)10 * (n > 0 ? n : 1) * sizeof(double);
  if (need > t->dump_cap) {
    if (t->d_dump) cudaFree(t->d_dump);
    t->d_dump = nullptr; t->dump_cap = 0;
    SDSO_CUDA(ctx, cudaMalloc(&t->d_dump, need));
    t->dump_cap = need;
  }
  P.dump_d = reinterpret_cast<double*>(t->d_dump);
  TrackProblem& hp = t->h_problems[0];
  memset(&hp, 0, sizeof(hp));
  memcpy(hp.T, T_select, sizeof(hp.T));
  memcpy(hp.T_out, T_pose, sizeof(hp.T_out));
  hp.aff_out[0] = photo[0]; hp.aff_out[1] = photo[1];
  for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frame].tex[l];
  hp.exposure_new = ctx->frames[new_frame].ab_exposure;
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_problems, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_track(ctx, P, 1, true);
  if (rc) return rc;
  std::vector<double> hd((size_t)10 * n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(hd.data(), t->d_dump, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int k = 0;
  for (int i = 0; i < n; i++) {
    if (hd[i] != 0.0) {
      if (err) err[k] = hd[(size_t)n + i];
      if (J8) for (int a = 0; a < 8; a++) J8[(size_t)8 * k + a] = hd[(size_t)(2 + a) * n + i];
      k++;
    }
  }
  *n_out = k;
  return SDSO_OK;
}

int sdso_track_enqueue(sdso_ctx* ctx, int nb, const int* new_frames, const double* T_in, const double* aff_in, int coarsest_lvl,
                       const double* minResForAbort, int variant) {
  if (!ctx || nb <= 0 || !new_frames || !T_in || !aff_in || !minResForAbort) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!t->have_ref) return fail(ctx, SDSO_E_STATE, "trackNewestCoarse before setCoarseTrackingRef");
  if (nb > t->max_problems) return fail(ctx, SDSO_E_INVALID, "too many problems in one batch");
  if (coarsest_lvl < 0 || coarsest_lvl >= ctx->G.levels || coarsest_lvl >= 5) return fail(ctx, SDSO_E_INVALID, "coarsest_lvl out of range");
  if (variant != SDSO_VARIANT_SSE && variant != SDSO_VARIANT_G2O) return SDSO_E_INVALID;
  TrackParams P;
  fill_params(ctx, P);
  P.mode = 0; P.coarsest = coarsest_lvl; P.variant = variant;
  if (variant == SDSO_VARIANT_G2O) { int rc = ensure_edge_scratch(ctx, P, nb); if (rc) return rc; }
  for (int k = 0; k < nb; k++) {
    int rc = check_frame(ctx, new_frames[k]);
    if (rc) return rc;
    TrackProblem& hp = t->h_problems[k];
    memset(&hp, 0, sizeof(hp));
    memcpy(hp.T, T_in + 12 * k, sizeof(hp.T));
    hp.aff[0] = aff_in[2 * k]; hp.aff[1] = aff_in[2 * k + 1];
    for (int i = 0; i < 5; i++) hp.minResForAbort[i] = minResForAbort[5 * k + i];
    for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frames[k]].tex[l];
    hp.exposure_new = ctx->frames[new_frames[k]].ab_exposure;
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_problems, t->h_problems, nb * sizeof(TrackProblem), cudaMemcpyHostToDevice, ctx->stream));
  prof_begin(ctx, 0);
  int rc = launch_track(ctx, P, nb, variant == SDSO_VARIANT_G2O);
  if (rc) return rc;
  prof_end(ctx, 0);
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->h_problems, t->d_problems, nb * sizeof(TrackProblem), cudaMemcpyDeviceToHost, ctx->stream));
  t->last_nb = nb;
  return SDSO_OK;
}

int sdso_track_collect(sdso_ctx* ctx, int nb, double* T_out, double* aff_out, double* lastResiduals, double* flowIndicators,
                       int* iterations, int* ok, uint64_t* evals) {
  if (!ctx) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (nb != t->last_nb) return fail(ctx, SDSO_E_STATE, "collect does not match the last enqueue");
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t ev = 0;
  for (int k = 0; k < nb; k++) {
    const TrackProblem& hp = t->h_problems[k];
    if (T_out) memcpy(T_out + 12 * k, hp.T_out, sizeof(hp.T_out));
    if (aff_out) { aff_out[2 * k] = hp.aff_out[0]; aff_out[2 * k + 1] = hp.aff_out[1]; }
    if (lastResiduals) memcpy(lastResiduals + 5 * k, hp.lastResiduals, sizeof(hp.lastResiduals));
    if (flowIndicators) memcpy(flowIndicators + 3 * k, hp.flow, sizeof(hp.flow));
    if (iterations) memcpy(iterations + 5 * k, hp.iterations, sizeof(hp.iterations));
    if (ok) ok[k] = hp.ok;
    ev += hp.evals;
  }
  if (evals) *evals = ev;
  return SDSO_OK;
}

int sdso_track_batch(sdso_ctx* ctx, int nb, const int* new_frames, double* T_io, double* aff_io, int coarsest_lvl,
                     const double* minResForAbort, int variant, double* lastResiduals, double* flowIndicators, int* iterations,
                     int* ok) {
  int rc = sdso_track_enqueue(ctx, nb, new_frames, T_io, aff_io, coarsest_lvl, minResForAbort, variant);
  if (rc) return rc;
  return sdso_track_collect(ctx, nb, T_io, aff_io, lastResiduals, flowIndicators, iterations, ok, nullptr);
}

int sdso_track(sdso_ctx* ctx, int new_frame, double T_io[12], double aff_io[2], int coarsest_lvl, const double minResForAbort[5],
               int variant, double lastResiduals[5], double flowIndicators[3], int iterations[5], int* ok) {
  return sdso_track_batch(ctx, 1, &new_frame, T_io, aff_io, coarsest_lvl, minResForAbort, variant, lastResiduals, flowIndicators,
                          iterations, ok);
}

}  // extern "C"
