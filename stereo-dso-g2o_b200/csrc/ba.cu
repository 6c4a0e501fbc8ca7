// Windowed bundle adjustment on the device — host side of B1-B12 (SSE path of the reference):
// window upload (SoA arenas, residuals stored sorted by (host,target)), the O(n^2) double-precision
// bookkeeping the reference also does on the host (FrameFramePrecalc::set, setAdjointsF, setDeltaF,
// nullspaces), and the C-ABI entry points that launch the kernels of ba_kernels.cuh.
#include "ba_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace sdso {

// ------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t grow(T*& p, size_t n) {
  if (p) cudaFree(p);
  p = nullptr;
  return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
}
#define BA_ALLOC(ptr, n) SDSO_CUDA(ctx, grow(ptr, n))

// captured launch chains (see capture_begin): dropped whenever something a captured kernel node holds BY VALUE changes — counts,
// array pointers, the shard, the presence of HM. Calibration, frame states and thresholds are read through device pointers.
static void invalidate_graphs(BAState* b) {
  if (b->graph_iter) { cudaGraphExecDestroy((cudaGraphExec_t)b->graph_iter); b->graph_iter = nullptr; }
  if (b->graph_assemble) { cudaGraphExecDestroy((cudaGraphExec_t)b->graph_assemble); b->graph_assemble = nullptr; }
}

static void free_all(BAState* b) {
  invalidate_graphs(b);
  if (b->cap_stream) { cudaStreamDestroy(b->cap_stream); b->cap_stream = nullptr; }
  if (b->cap_stream2) { cudaStreamDestroy(b->cap_stream2); b->cap_stream2 = nullptr; }
  if (b->ev_fork) { cudaEventDestroy(b->ev_fork); b->ev_fork = nullptr; }
  if (b->ev_join) { cudaEventDestroy(b->ev_join); b->ev_join = nullptr; }
  void* ptrs[] = {b->d_tex0, b->d_frameTH, b->d_precalc, b->d_adHost, b->d_adTarget, b->d_adHostF, b->d_adTargetF, b->d_adHTdeltaF, b->d_cDeltaF,
                  b->d_fprior, b->d_p_host, b->d_p_u, b->d_p_v, b->d_p_idepth, b->d_p_idepth_zero, b->d_p_color, b->d_p_weights, b->d_p_priorF,
                  b->d_p_deltaF, b->d_p_idepth_backup, b->d_p_res_begin, b->d_slot_of, b->d_p_acc, b->d_p_flag, b->d_p_res_list, b->d_s_point, b->d_s_key, b->d_s_state,
                  b->d_s_newstate, b->d_s_flags, b->d_s_sel, b->d_s_energy, b->d_J, b->d_s_rtz, b->d_s_JpJd, b->d_s_center, b->d_s_psum,
                  b->d_slot2rid, b->d_rid2slot, b->d_chunks, b->d_key_chunk_begin, b->d_tpart, b->d_dpart, b->d_pblockpart, b->d_G, b->d_Gf,
                  b->d_D, b->d_E, b->d_Hcc, b->d_U, b->d_V, b->d_sys, b->d_energy_part, b->d_scalars, b->d_counter, b->d_N, b->d_xAd, b->d_list, b->d_step_part,
                  b->d_frames, b->d_calib, b->d_opt, b->lba_scratch, b->d_W};
  for (void* p : ptrs) if (p) cudaFree(p);
}

int ba_create(sdso_ctx* ctx) {
  BAState* b = new BAState();
  ctx->ba = b;
  const int F = kMaxFrames, F2 = F * F, dmax = kCPARS + 8 * F;
  BA_ALLOC(b->d_tex0, F); BA_ALLOC(b->d_frameTH, F); BA_ALLOC(b->d_precalc, F2);
  BA_ALLOC(b->d_adHost, F2 * 64); BA_ALLOC(b->d_adTarget, F2 * 64); BA_ALLOC(b->d_adHostF, F2 * 64); BA_ALLOC(b->d_adTargetF, F2 * 64);
  BA_ALLOC(b->d_adHTdeltaF, F2 * 8); BA_ALLOC(b->d_cDeltaF, 4); BA_ALLOC(b->d_fprior, F * 24);
  BA_ALLOC(b->d_G, F2 * 169 + F2); BA_ALLOC(b->d_Gf, F2 * 169);
  BA_ALLOC(b->d_D, (size_t)F2 * F * 65); BA_ALLOC(b->d_E, F2 * 40 * 3); BA_ALLOC(b->d_Hcc, 20);
  BA_ALLOC(b->d_U, (size_t)F2 * F * 64); BA_ALLOC(b->d_V, (size_t)F2 * F * 64);
  b->sys_stride = (size_t)dmax * dmax + dmax + 8;
  BA_ALLOC(b->d_sys, b->sys_stride * SYS_NUM);
  SDSO_CUDA(ctx, cudaMemset(b->d_sys, 0, b->sys_stride * SYS_NUM * sizeof(double)));
  BA_ALLOC(b->d_scalars, 16); BA_ALLOC(b->d_counter, 4); BA_ALLOC(b->d_N, dmax * 7); BA_ALLOC(b->d_xAd, F2 * 8);
  SDSO_CUDA(ctx, cudaMemset(b->d_counter, 0, 4 * sizeof(unsigned)));
  BA_ALLOC(b->d_frames, F); BA_ALLOC(b->d_calib, 1); BA_ALLOC(b->d_opt, 1);
  SDSO_CUDA(ctx, cudaMemset(b->d_opt, 0, sizeof(OptDev)));
  SDSO_CUDA(ctx, cudaStreamCreateWithFlags(&b->cap_stream, cudaStreamNonBlocking));
  SDSO_CUDA(ctx, cudaStreamCreateWithFlags(&b->cap_stream2, cudaStreamNonBlocking));
  SDSO_CUDA(ctx, cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
  SDSO_CUDA(ctx, cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  BA_ALLOC(b->d_W, (size_t)F2 * (64 + 64 + 40 + 40));
  SDSO_CUDA(ctx, cudaFuncSetAttribute(ba_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return SDSO_OK;
}
void ba_destroy(sdso_ctx* ctx) {
  if (!ctx->ba) return;
  free_all(ctx->ba);
  delete ctx->ba;
  ctx->ba = nullptr;
}

static BAView view(BAState* b) {
  BAView v;
  v.n = b->n; v.P = b->P; v.R = b->R; v.capP = b->capP; v.capR = b->capR; v.cp = b->d_calib; v.done = &b->d_opt->done;
  v.tex0 = b->d_tex0; v.frameTH = b->d_frameTH; v.precalc = b->d_precalc;
  v.adHost = b->d_adHost; v.adTarget = b->d_adTarget; v.adHostF = b->d_adHostF; v.adTargetF = b->d_adTargetF;
  v.adHTdeltaF = b->d_adHTdeltaF; v.cDeltaF = b->d_cDeltaF; v.fprior = b->d_fprior;
  v.p_host = b->d_p_host; v.p_u = b->d_p_u; v.p_v = b->d_p_v; v.p_idepth = b->d_p_idepth; v.p_idepth_zero = b->d_p_idepth_zero; v.p_idepth_backup = b->d_p_idepth_backup;
  v.p_color = b->d_p_color; v.p_weights = b->d_p_weights; v.p_priorF = b->d_p_priorF; v.p_deltaF = b->d_p_deltaF;
  v.p_res_begin = b->d_p_res_begin; v.p_res_list = b->d_p_res_list; v.slot_of = b->d_slot_of; v.p_acc = b->d_p_acc; v.p_flag = b->d_p_flag;
  v.s_point = b->d_s_point; v.s_key = b->d_s_key; v.s_state = b->d_s_state; v.s_newstate = b->d_s_newstate; v.s_flags = b->d_s_flags; v.s_sel = b->d_s_sel;
  v.s_energy = b->d_s_energy; v.J = b->d_J; v.s_rtz = b->d_s_rtz; v.s_JpJd = b->d_s_JpJd; v.s_center = b->d_s_center; v.s_psum = b->d_s_psum;
  v.chunks = b->d_chunks; v.nchunks = b->nchunks; v.key_chunk_begin = b->d_key_chunk_begin;
  v.tpart = b->d_tpart; v.dpart = b->d_dpart; v.pblockpart = b->d_pblockpart;
  v.G = b->d_G; v.Gf = b->d_Gf; v.D = b->d_D; v.E = b->d_E; v.Hcc = b->d_Hcc; v.U = b->d_U; v.V = b->d_V;
  v.energy_part = b->d_energy_part; v.scalars = b->d_scalars; v.counter = b->d_counter;
  return v;
}

static inline double* sysH(BAState* b, int which) { return b->d_sys + b->sys_stride * which; }
// b follows H contiguously ((4+8n)^2 + (4+8n) doubles): one buffer per (H,b) pair, so a shard's partial system is ONE allreduce
static inline double* sysb(BAState* b, int which) { const int d = b->dim(); return b->d_sys + b->sys_stride * which + (size_t)d * d; }

// ---- FrameHessian state handling (HessianBlocks.h:177-231, HessianBlocks.cpp:78-123) ----------------
static void frame_update_pre(HostBAFrame& f) {
  double E[12];
  se3_exp(f.state_scaled, E);
  se3_mul(E, f.T_eval, f.T_w2c);
  se3_inv(f.T_w2c, f.T_c2w);
}
static void frame_set_state(HostBAFrame& f, const double s[10]) {
  for (int i = 0; i < 10; i++) f.state[i] = s[i];
  for (int i = 0; i < 3; i++) f.state_scaled[i] = SCALE_XI_TRANS * f.state[i];
  for (int i = 3; i < 6; i++) f.state_scaled[i] = SCALE_XI_ROT * f.state[i];
  f.state_scaled[6] = SCALE_A * f.state[6]; f.state_scaled[7] = SCALE_B * f.state[7];
  f.state_scaled[8] = SCALE_A * f.state[8]; f.state_scaled[9] = SCALE_B * f.state[9];
  frame_update_pre(f);
}
static void frame_set_state_scaled(HostBAFrame& f, const double s[10]) {
  for (int i = 0; i < 10; i++) f.state_scaled[i] = s[i];
  for (int i = 0; i < 3; i++) f.state[i] = (1.0f / SCALE_XI_TRANS) * f.state_scaled[i];
  for (int i = 3; i < 6; i++) f.state[i] = (1.0f / SCALE_XI_ROT) * f.state_scaled[i];
  f.state[6] = (1.0f / SCALE_A) * f.state_scaled[6]; f.state[7] = (1.0f / SCALE_B) * f.state_scaled[7];
  f.state[8] = (1.0f / SCALE_A) * f.state_scaled[8]; f.state[9] = (1.0f / SCALE_B) * f.state_scaled[9];
  frame_update_pre(f);
}
static void frame_set_state_zero(HostBAFrame& f) {
  for (int i = 0; i < 10; i++) f.state_zero[i] = f.state[i];
  double inv0[12];
  se3_inv(f.T_eval, inv0);
  for (int i = 0; i < 6; i++) {  // numeric derivative of log(T exp(eps) T^-1) (HessianBlocks.cpp:83-93)
    double eps[6] = {0, 0, 0, 0, 0, 0}, Ep[12], Em[12], Pp[12], Pm[12], lp[6], lm[6];
    eps[i] = 1e-3; se3_exp(eps, Ep);
    eps[i] = -1e-3; se3_exp(eps, Em);
    se3_mul(f.T_eval, Ep, Pp); se3_mul(Pp, inv0, Pp);
    se3_mul(f.T_eval, Em, Pm); se3_mul(Pm, inv0, Pm);
    se3_log(Pp, lp); se3_log(Pm, lm);
    for (int r = 0; r < 6; r++) f.ns_pose[r * 6 + i] = (lp[r] - lm[r]) / (2e-3);
  }
  double Pp[12], Pm[12], lp[6], lm[6];
  memcpy(Pp, f.T_eval, sizeof(Pp)); memcpy(Pm, f.T_eval, sizeof(Pm));
  for (int k = 0; k < 3; k++) { Pp[k * 4 + 3] *= 1.00001; Pm[k * 4 + 3] /= 1.00001; }
  se3_mul(Pp, inv0, Pp); se3_mul(Pm, inv0, Pm);
  se3_log(Pp, lp); se3_log(Pm, lm);
  for (int r = 0; r < 6; r++) f.ns_scale[r] = (lp[r] - lm[r]) / (2e-3);
}

static void aff_from_to(float eF, float eT, double aF, double bF, double aT, double bT, double out[2]) {  // NumType.h:159-170
  if (eF == 0 || eT == 0) { eT = eF = 1; }
  const double a = std::exp(aT - aF) * eT / eF;
  out[0] = a; out[1] = bT - a * bF;
}
static void mat33f_mul(const float A[9], const float B[9], float C[9]) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}

// FrameFramePrecalc::set for all pairs (HessianBlocks.cpp:206-242), setAdjointsF (EnergyFunctional.cpp:41-119),
// setDeltaF (:173-207), getNullspaces (FullSystemOptimize.cpp:1087-1147); uploads the results.
static int prepare_window(sdso_ctx* ctx) {
  BAState* b = ctx->ba;
  const int n = b->n;
  const sdso_settings& S = ctx->S;
  std::vector<PrecalcDev> pre((size_t)n * n);
  std::vector<double> adH((size_t)n * n * 64, 0.0), adT((size_t)n * n * 64, 0.0);
  std::vector<float> adHF((size_t)n * n * 64), adTF((size_t)n * n * 64), adHTd((size_t)n * n * 8);
  const BACalib& c = b->calib;
  const float K[9] = {c.fxl, 0, c.cxl, 0, c.fyl, c.cyl, 0, 0, 1};
  float Kinv[9];
  inverse3f(K, Kinv);
  for (int h = 0; h < n; h++) for (int t = 0; t < n; t++) {
    const HostBAFrame& host = b->frames[h]; const HostBAFrame& target = b->frames[t];
    PrecalcDev& p = pre[(size_t)h * n + t];
    memset(&p, 0, sizeof(p));
    double hinv[12], l0[12], l[12];
    se3_inv(host.T_eval, hinv);
    se3_mul(target.T_eval, hinv, l0);
    se3_mul(target.T_w2c, host.T_c2w, l);
    for (int r = 0; r < 3; r++) {
      for (int q = 0; q < 3; q++) { p.PRE_RTll_0[r * 3 + q] = (float)l0[r * 4 + q]; p.PRE_RTll[r * 3 + q] = (float)l[r * 4 + q]; }
      p.PRE_tTll_0[r] = (float)l0[r * 4 + 3]; p.PRE_tTll[r] = (float)l[r * 4 + 3];
    }
    p.distanceLL = (float)std::sqrt(l[3] * l[3] + l[7] * l[7] + l[11] * l[11]);
    float KR[9];
    mat33f_mul(K, p.PRE_RTll, KR);
    mat33f_mul(KR, Kinv, p.PRE_KRKiTll);
    mat33f_mul(p.PRE_RTll, Kinv, p.PRE_RKiTll);
    for (int r = 0; r < 3; r++) p.PRE_KtTll[r] = K[r * 3] * p.PRE_tTll[0] + K[r * 3 + 1] * p.PRE_tTll[1] + K[r * 3 + 2] * p.PRE_tTll[2];
    double ab[2];
    aff_from_to(host.ab_exposure, target.ab_exposure, host.state_scaled[6], host.state_scaled[7], target.state_scaled[6], target.state_scaled[7], ab);
    p.PRE_aff_mode[0] = (float)ab[0]; p.PRE_aff_mode[1] = (float)ab[1];
    p.PRE_b0_mode = (float)(host.state_zero[7] * SCALE_B);
    // adjoints at the evaluation point
    double Adj[36], AH[64] = {0}, AT[64] = {0};
    se3_adj(l0, Adj);
    for (int i = 0; i < 8; i++) { AH[i * 8 + i] = 1; AT[i * 8 + i] = 1; }
    for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++) AH[r * 8 + q] = -Adj[q * 6 + r];
    double ab0[2];
    aff_from_to(host.ab_exposure, target.ab_exposure, host.state_zero[6] * SCALE_A, host.state_zero[7] * SCALE_B, target.state_zero[6] * SCALE_A,
                target.state_zero[7] * SCALE_B, ab0);
    const float affLL0 = (float)ab0[0];
    AT[6 * 8 + 6] = -affLL0; AT[7 * 8 + 7] = -1;
    AH[6 * 8 + 6] = affLL0; AH[7 * 8 + 7] = affLL0;
    const double rs[8] = {SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_A, SCALE_B};
    for (int r = 0; r < 8; r++) for (int q = 0; q < 8; q++) { AH[r * 8 + q] *= rs[r]; AT[r * 8 + q] *= rs[r]; }
    const size_t o = ((size_t)h + (size_t)t * n) * 64;
    for (int i = 0; i < 64; i++) { adH[o + i] = AH[i]; adT[o + i] = AT[i]; adHF[o + i] = (float)AH[i]; adTF[o + i] = (float)AT[i]; }
  }
  for (int i = 0; i < 4; i++) b->cPrior[i] = S.initialCalibHessian;
  // setDeltaF
  std::vector<double> fpr((size_t)kMaxFrames * 24, 0.0);
  for (int h = 0; h < n; h++) {
    HostBAFrame& f = b->frames[h];
    for (int i = 0; i < 8; i++) { f.delta[i] = f.state[i] - f.state_zero[i]; f.delta_prior[i] = f.state[i]; }
    for (int i = 0; i < 8; i++) { fpr[h * 24 + i] = f.prior[i]; fpr[h * 24 + 8 + i] = f.delta_prior[i]; fpr[h * 24 + 16 + i] = f.delta[i]; }
  }
  for (int h = 0; h < n; h++) for (int t = 0; t < n; t++) {
    const size_t idx = (size_t)h + (size_t)t * n;
    float dh[8], dt[8];
    for (int i = 0; i < 8; i++) { dh[i] = (float)(b->frames[h].state[i] - b->frames[h].state_zero[i]); dt[i] = (float)(b->frames[t].state[i] - b->frames[t].state_zero[i]); }
    for (int j = 0; j < 8; j++) {
      float a = 0, bb = 0;
      for (int i = 0; i < 8; i++) a += dh[i] * adHF[idx * 64 + i * 8 + j];
      for (int i = 0; i < 8; i++) bb += dt[i] * adTF[idx * 64 + i * 8 + j];
      adHTd[idx * 8 + j] = a + bb;
    }
  }
  float cDeltaF[4];
  for (int i = 0; i < 4; i++) cDeltaF[i] = (float)b->calib_delta[i];
  // nullspaces: 6 pose + 1 scale vectors over the frames (calibration rows are zero)
  const int d = b->dim();
  std::vector<double> N((size_t)d * 7, 0.0);
  for (int f = 0; f < n; f++) for (int r = 0; r < 6; r++) {
    const double sc = (r < 3) ? (1.0f / SCALE_XI_TRANS) : (1.0f / SCALE_XI_ROT);
    for (int i = 0; i < 6; i++) N[(size_t)(kCPARS + f * 8 + r) * 7 + i] = b->frames[f].ns_pose[r * 6 + i] * sc;
    N[(size_t)(kCPARS + f * 8 + r) * 7 + 6] = b->frames[f].ns_scale[r] * sc;
  }
  b->h_N = N;
  // Orthonormal basis of span(N) restricted to singular values > solverModeDelta * max (EnergyFunctional::orthogonalize,
  // EnergyFunctional.cpp:775-835): columns are normalised, the SVD comes from the Jacobi eigen-decomposition of N^T N.
  std::vector<double> Q((size_t)d * 7, 0.0);
  int nrank = 0;
  {
    const int m = 7;
    std::vector<double> Nn((size_t)d * m);
    for (int i = 0; i < m; i++) {
      double nrm = 0;
      for (int r = 0; r < d; r++) nrm += N[(size_t)r * m + i] * N[(size_t)r * m + i];
      nrm = std::sqrt(nrm);
      for (int r = 0; r < d; r++) Nn[(size_t)r * m + i] = N[(size_t)r * m + i] / nrm;
    }
    double G[49], V[49];
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) {
      double sacc = 0;
      for (int r = 0; r < d; r++) sacc += Nn[(size_t)r * m + i] * Nn[(size_t)r * m + j];
      G[i * m + j] = sacc; V[i * m + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0;
      for (int i = 0; i < m; i++) for (int j = i + 1; j < m; j++) off += G[i * m + j] * G[i * m + j];
      if (off < 1e-300) break;
      for (int p = 0; p < m; p++) for (int q = p + 1; q < m; q++) {
        if (std::fabs(G[p * m + q]) < 1e-300) continue;
        const double theta = (G[q * m + q] - G[p * m + p]) / (2 * G[p * m + q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double cc = 1 / std::sqrt(t * t + 1), ss = t * cc;
        for (int k = 0; k < m; k++) { const double x = G[k * m + p], y = G[k * m + q]; G[k * m + p] = cc * x - ss * y; G[k * m + q] = ss * x + cc * y; }
        for (int k = 0; k < m; k++) { const double x = G[p * m + k], y = G[q * m + k]; G[p * m + k] = cc * x - ss * y; G[q * m + k] = ss * x + cc * y; }
        for (int k = 0; k < m; k++) { const double x = V[k * m + p], y = V[k * m + q]; V[k * m + p] = cc * x - ss * y; V[k * m + q] = ss * x + cc * y; }
      }
    }
    double mx = 0, sv[7];
    for (int i = 0; i < m; i++) { sv[i] = std::sqrt(std::max(G[i * m + i], 0.0)); mx = std::max(mx, sv[i]); }
    for (int i = 0; i < m; i++) {
      if (!(sv[i] > S.solverModeDelta * mx)) continue;
      for (int r = 0; r < d; r++) {
        double nv = 0;
        for (int j = 0; j < m; j++) nv += Nn[(size_t)r * m + j] * V[j * m + i];
        Q[(size_t)r * 7 + nrank] = nv / sv[i];
      }
      nrank++;
    }
  }
  b->nrank = nrank;
  // per-frame device pointers and thresholds
  std::vector<const float4*> tex(kMaxFrames, nullptr);
  std::vector<float> th(kMaxFrames, 0.f);
  for (int h = 0; h < n; h++) { tex[h] = ctx->frames[b->frames[h].frame_id].tex[0]; th[h] = b->frames[h].frameEnergyTH; }
  b->h_pre = pre; b->h_adH = adH; b->h_adT = adT; b->h_adHTd = adHTd;
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_precalc, pre.data(), pre.size() * sizeof(PrecalcDev), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_adHost, adH.data(), adH.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_adTarget, adT.data(), adT.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_adHostF, adHF.data(), adHF.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_adTargetF, adTF.data(), adTF.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_adHTdeltaF, adHTd.data(), adHTd.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_cDeltaF, cDeltaF, sizeof(cDeltaF), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_fprior, fpr.data(), fpr.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_N, Q.data(), Q.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_tex0, tex.data(), tex.size() * sizeof(float4*), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_frameTH, th.data(), th.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  // device mirrors of what the LM loop moves: calibration, frame states, the loop's control block
  std::vector<FrameDev> fd(kMaxFrames);
  memset(fd.data(), 0, fd.size() * sizeof(FrameDev));
  for (int h = 0; h < n; h++) {
    const HostBAFrame& f = b->frames[h];
    memcpy(fd[h].T_eval, f.T_eval, sizeof(f.T_eval)); memcpy(fd[h].state, f.state, sizeof(f.state)); memcpy(fd[h].state_zero, f.state_zero, sizeof(f.state_zero));
    memcpy(fd[h].state_backup, f.state_backup, sizeof(f.state_backup)); memcpy(fd[h].T_w2c, f.T_w2c, sizeof(f.T_w2c)); memcpy(fd[h].T_c2w, f.T_c2w, sizeof(f.T_c2w));
    fd[h].ab_exposure = f.ab_exposure;
  }
  OptDev od;
  memset(&od, 0, sizeof(od));
  for (int i = 0; i < 4; i++) { od.calib_value[i] = b->calib_value[i]; od.calib_zero[i] = b->calib_zero[i]; od.calib_backup[i] = b->calib_backup[i]; }
  od.th_opt = S.thOptIterations; od.min_its = S.minOptIterations;
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_frames, fd.data(), fd.size() * sizeof(FrameDev), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_calib, &b->calib, sizeof(BACalib), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_opt, &od, sizeof(OptDev), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_scalars + 8, b->cPrior, 4 * sizeof(double), cudaMemcpyHostToDevice, st));   // cPrior lives at scalars[8..11]
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));  // the host vectors above go out of scope
  // deltaF of the points follows idepth - idepth_zero (EFPoint::takeData / setDeltaF :196-204): device side
  b->prepared = true;
  return SDSO_OK;
}

__global__ void ba_point_delta_kernel(int P, const float* idepth, const float* idepth_zero, float* deltaF) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) deltaF[p] = idepth[p] - idepth_zero[p];
}

// ---- launches ---------------------------------------------------------------------------------------------
// AccumulatedTopHessianSSE in three pieces, so that a captured chain can run the frame-block half beside the per-point / Schur half
static int launch_top_accumulate(sdso_ctx* ctx, int mode) {   // addPoint<mode> over the residual chunks: 13x13 partials + per-residual point terms
  BAState* b = ctx->ba;
  BAView v = view(b);
  if (b->nchunks > 0) { ba_top_kernel<<<b->nchunks, kChunk, 0, ctx->stream>>>(v, mode); SDSO_CHECK_LAUNCH(ctx); }
  return SDSO_OK;
}
static int launch_top_frames(sdso_ctx* ctx, int which, bool usePrior) {   // fixed-order sums of the partials + stitchDouble
  BAState* b = ctx->ba;
  BAView v = view(b);
  const int n = b->n;
  double* Wm = b->d_W; double* Zm = Wm + (size_t)kMaxFrames * kMaxFrames * 64; double* Wc = Zm + (size_t)kMaxFrames * kMaxFrames * 64; double* Zc = Wc + (size_t)kMaxFrames * kMaxFrames * 40;
  ba_top_finish_kernel<<<n * n, 256, 0, ctx->stream>>>(v, Wm, Zm, Wc, Zc); SDSO_CHECK_LAUNCH(ctx);
  double* dc = b->d_scalars + 8;   // cPrior (uploaded by prepare_window)
  ba_stitch_top_kernel<<<n * n + 1, 256, 0, ctx->stream>>>(v, Wm, Zm, Wc, Zc, sysH(b, which), sysb(b, which), usePrior ? 1 : 0, dc);
  SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}
static int launch_top_points(sdso_ctx* ctx, int mode) {   // bd_acc / Hdd_acc / Hcd_acc per point (AccumulatedTopHessian.cpp:160-192)
  BAState* b = ctx->ba;
  BAView v = view(b);
  if (b->P > 0) { ba_point_sums_kernel<<<(b->P + 127) / 128, 128, 0, ctx->stream>>>(v, mode); SDSO_CHECK_LAUNCH(ctx); }
  return SDSO_OK;
}
static int launch_top(sdso_ctx* ctx, int mode, int which, bool usePrior) {
  int rc = launch_top_accumulate(ctx, mode);
  if (!rc) rc = launch_top_points(ctx, mode);
  if (!rc) rc = launch_top_frames(ctx, which, usePrior);
  return rc;
}

static int launch_sc(sdso_ctx* ctx, bool shift, int which) {
  BAState* b = ctx->ba;
  BAView v = view(b);
  const int d = b->dim(), n = b->n;
  ba_sc_point_kernel<<<std::max(b->pblocks, 1), 128, 0, ctx->stream>>>(v, shift ? 1 : 0); SDSO_CHECK_LAUNCH(ctx);
  if (b->nchunks > 0) { ba_sc_pair_kernel<<<dim3(b->nchunks, n + 1), kChunk, 0, ctx->stream>>>(v, shift ? 1 : 0); SDSO_CHECK_LAUNCH(ctx); }
  double* Uc = b->d_E + (size_t)kMaxFrames * kMaxFrames * 40; double* Vc = Uc + (size_t)kMaxFrames * kMaxFrames * 40;
  ba_sc_finish_kernel<<<dim3(n * n, n + 1), 96, 0, ctx->stream>>>(v, b->pblocks, Uc, Vc); SDSO_CHECK_LAUNCH(ctx);
  ba_stitch_sc_kernel<<<n * n + 1, 256, 0, ctx->stream>>>(v, Uc, Vc, sysH(b, which), sysb(b, which)); SDSO_CHECK_LAUNCH(ctx);
  (void)d;
  return SDSO_OK;
}

static int download_sys(sdso_ctx* ctx, int which, double* H, double* bv) {
  BAState* b = ctx->ba;
  const int d = b->dim();
  if (!H && !bv) return SDSO_OK;  // nothing to return: stay asynchronous
  if (H) SDSO_CUDA(ctx, cudaMemcpyAsync(H, sysH(b, which), (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (bv) SDSO_CUDA(ctx, cudaMemcpyAsync(bv, sysb(b, which), d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

static void fill_solve_params(sdso_ctx* ctx, SolveParams& S, int iteration) {
  BAState* b = ctx->ba;
  S.n = b->n; S.d = b->dim(); S.iteration = iteration; S.plain = 0;
  S.have_M = (b->have_M && b->shard_rank == 0) ? 1 : 0;  // HM / bM enter once (rank 0 of a sharded window)
  S.lambda = 1e-5;  // SOLVER_FIX_LAMBDA (EnergyFunctional.cpp:844-846): setting_solverMode fixes lambda regardless of the argument
  S.solverModeDelta = ctx->S.solverModeDelta;
  S.HA = sysH(b, SYS_A); S.bA = sysb(b, SYS_A); S.HL = sysH(b, SYS_L); S.bL = sysb(b, SYS_L); S.Hsc = sysH(b, SYS_SC); S.bsc = sysb(b, SYS_SC);
  S.HM = sysH(b, SYS_M); S.bM = sysb(b, SYS_M);
  S.fprior = b->d_fprior; S.cDeltaF = b->d_cDeltaF; S.N = b->d_N; S.nrank = b->nrank;
  S.HF = sysH(b, SYS_FINAL); S.bF = sysb(b, SYS_FINAL); S.x = sysb(b, SYS_X);
  S.done = &b->d_opt->done; S.it_ptr = iteration < 0 ? &b->d_opt->it : nullptr;
}

// ---- CUDA graphs for the launch chains ----------------------------------------------------------------------------------
// One LM iteration is a chain of ~17-25 small dependent kernels; launched one by one the chain is bound by launch latency
// (config 3: 0.2 ms for ~20 us of work). The chain is captured once per uploaded window on an internal stream (the context's
// stream may be the legacy default stream, which cannot capture) and replayed as one graph launch on the context's stream.
struct Capture { cudaStream_t user = nullptr; bool active = false; };
static int capture_begin(sdso_ctx* ctx, Capture& c) {
  BAState* b = ctx->ba;
  c.user = ctx->stream;
  SDSO_CUDA(ctx, cudaStreamBeginCapture(b->cap_stream, cudaStreamCaptureModeThreadLocal));
  ctx->stream = b->cap_stream;
  c.active = true;
  return SDSO_OK;
}
static int capture_end(sdso_ctx* ctx, Capture& c, void** exec_out, int rc_body) {
  BAState* b = ctx->ba;
  ctx->stream = c.user;
  c.active = false;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(b->cap_stream, &g);
  if (rc_body) { if (g) cudaGraphDestroy(g); return rc_body; }
  if (e != cudaSuccess) return fail(ctx, SDSO_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(ctx, SDSO_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  *exec_out = ex;
  return SDSO_OK;
}

static int launch_assemble_chain(sdso_ctx* ctx);
// accumulateAF_MT + accumulateLF_MT + accumulateSCF_MT (EnergyFunctional.cpp:857-866) over this rank's points, then the
// (partial) damped reduced system into SYS_FINAL — replayed as one graph
static int launch_assemble(sdso_ctx* ctx) {
  BAState* b = ctx->ba;
  if (!b->graph_assemble) {
    Capture c;
    int rc = capture_begin(ctx, c);
    if (rc) return rc;
    const uint64_t l0 = ctx->launches;
    rc = launch_assemble_chain(ctx);
    b->graph_assemble_nodes = (int)(ctx->launches - l0);
    ctx->launches = l0;
    rc = capture_end(ctx, c, &b->graph_assemble, rc);
    if (rc) return rc;
  }
  SDSO_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)b->graph_assemble, ctx->stream));
  ctx->launches += b->graph_assemble_nodes;   // kernels executed (one graph launch)
  return SDSO_OK;
}
static int launch_assemble_chain(sdso_ctx* ctx) {
  BAState* b = ctx->ba;
  const int d = b->dim();
  int rc = SDSO_OK;
  auto priors_only = [&]() -> int {   // nothing linearised: the L system is the priors alone
    BAView v = view(b);
    ba_prior_system_kernel<<<(d * d + d + 127) / 128, 128, 0, ctx->stream>>>(v, sysH(b, SYS_L), sysb(b, SYS_L), b->shard_rank == 0 ? 1 : 0, b->d_scalars + 8);
    SDSO_CHECK_LAUNCH(ctx);
    return SDSO_OK;
  };
  if (ctx->stream == b->cap_stream && !b->any_linearized) {
    // Inside a capture: after the accumulation kernel the chain forks — {fixed-order block sums, top stitch, priors} on a second
    // stream beside {per-point sums, Schur point / pair / finish kernels, Schur stitch} — and joins in front of the assembly. The
    // two halves share no output (the top stitch has its own W/Z scratch), so the graph runs them concurrently.
    cudaStream_t s1 = b->cap_stream, s2 = b->cap_stream2;
    if ((rc = launch_top_accumulate(ctx, 0))) return rc;
    SDSO_CUDA(ctx, cudaEventRecord(b->ev_fork, s1));
    SDSO_CUDA(ctx, cudaStreamWaitEvent(s2, b->ev_fork, 0));
    ctx->stream = s2;
    rc = launch_top_frames(ctx, SYS_A, false);
    if (!rc) rc = priors_only();
    ctx->stream = s1;
    if (rc) return rc;
    SDSO_CUDA(ctx, cudaEventRecord(b->ev_join, s2));
    if ((rc = launch_top_points(ctx, 0))) return rc;
    if ((rc = launch_sc(ctx, true, SYS_SC))) return rc;
    SDSO_CUDA(ctx, cudaStreamWaitEvent(s1, b->ev_join, 0));
  } else {
    rc = launch_top(ctx, 0, SYS_A, false);
    if (!rc) rc = b->any_linearized ? launch_top(ctx, 1, SYS_L, b->shard_rank == 0) /* priors enter once */ : priors_only();
    if (!rc) rc = launch_sc(ctx, true, SYS_SC);
    if (rc) return rc;
  }
  SolveParams S{};
  fill_solve_params(ctx, S, 0);
  ba_assemble_kernel<<<(d * d + d + 127) / 128, 128, 0, ctx->stream>>>(S);
  SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}

static int launch_factor_solve(sdso_ctx* ctx, int iteration) {
  BAState* b = ctx->ba;
  const int d = b->dim();
  SolveParams S{};
  fill_solve_params(ctx, S, iteration);
  const size_t smem = ((size_t)d * (d | 1) + 6 * (size_t)d + 7 * (size_t)d + 256) * sizeof(double);
  ba_solve_kernel<<<1, 256, smem, ctx->stream>>>(S);   // (dynamic shared-memory limit raised in ba_create)
  SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}

static int launch_solve(sdso_ctx* ctx, int iteration, double lambda) {
  (void)lambda;
  int rc = launch_assemble(ctx);
  if (!rc) rc = launch_factor_solve(ctx, iteration);
  return rc;
}

static int launch_resub(sdso_ctx* ctx, const double* d_x) {
  BAState* b = ctx->ba;
  BAView v = view(b);
  const int n = b->n;
  ba_xad_kernel<<<(n * n * 8 + 127) / 128, 128, 0, ctx->stream>>>(v, d_x, b->d_xAd); SDSO_CHECK_LAUNCH(ctx);
  if (b->P > 0) { ba_resub_kernel<<<(b->P + 127) / 128, 128, 0, ctx->stream>>>(v, d_x, b->d_xAd); SDSO_CHECK_LAUNCH(ctx); }
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

#define BA_CHECK(ctx)                                                  \
  if (!(ctx) || !(ctx)->ba) return SDSO_E_INVALID;                     \
  BAState* b = (ctx)->ba;                                              \
  (void)b;
#define BA_PREPARED(ctx)                                               \
  BA_CHECK(ctx)                                                        \
  if (!b->prepared) return fail(ctx, SDSO_E_STATE, "sdso_ba_prepare has not been called for this window");

extern "C" {

int sdso_ba_reset(sdso_ctx* ctx) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  b->n = b->P = b->R = 0;
  b->frames.clear();
  b->prepared = false;
  b->have_M = false;
  b->shard_rank = 0; b->shard_n = 1;
  for (int i = 0; i < 4; i++) b->calib_delta[i] = 0;
  // default calibration = the context's initial one
  const float K[4] = {ctx->G.fx[0], ctx->G.fy[0], ctx->G.cx[0], ctx->G.cy[0]};
  return sdso_ba_set_calib(ctx, K, nullptr);
}

int sdso_ba_set_calib(sdso_ctx* ctx, const float K[4], const double value_minus_value_zero[4]) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (!K) return SDSO_E_INVALID;
  BACalib& c = b->calib;
  c.fxl = K[0]; c.fyl = K[1]; c.cxl = K[2]; c.cyl = K[3];
  c.fxli = 1.0f / c.fxl; c.fyli = 1.0f / c.fyl; c.cxli = -c.cxl / c.fxl; c.cyli = -c.cyl / c.fyl;  // HessianBlocks.h:320-323
  c.w0 = ctx->G.w[0]; c.h0 = ctx->G.h[0];
  c.wM3G = (float)(c.w0 - 3); c.hM3G = (float)(c.h0 - 3);
  c.huberTH = ctx->S.huberTH; c.outlierTHSumComponent = ctx->S.outlierTHSumComponent;
  c.affineOptModeA = ctx->S.affineOptModeA; c.affineOptModeB = ctx->S.affineOptModeB;
  if (value_minus_value_zero) for (int i = 0; i < 4; i++) b->calib_delta[i] = value_minus_value_zero[i];
  // CalibHessian: value = SCALE_*_INVERSE * value_scaled, value_zero = value - (value - value_zero)
  b->calib_value[0] = (1.0f / SCALE_F) * (double)c.fxl; b->calib_value[1] = (1.0f / SCALE_F) * (double)c.fyl;
  b->calib_value[2] = (1.0f / SCALE_C) * (double)c.cxl; b->calib_value[3] = (1.0f / SCALE_C) * (double)c.cyl;
  for (int i = 0; i < 4; i++) b->calib_zero[i] = b->calib_value[i] - b->calib_delta[i];
  b->prepared = false;
  return SDSO_OK;
}

int sdso_ba_add_frame(sdso_ctx* ctx, int frame_id, const double T_w2c[12], double a, double bb, int frameID, int* idx_out) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  if (!T_w2c || frame_id < 0 || frame_id >= (int)ctx->frames.size() || !ctx->frames[frame_id].valid) return SDSO_E_INVALID;
  if (b->n >= kMaxFrames) return fail(ctx, SDSO_E_INVALID, "window is full (kMaxFrames)");
  HostBAFrame f;
  f.frame_id = frame_id; f.frameID = frameID; f.ab_exposure = ctx->frames[frame_id].ab_exposure;
  memcpy(f.T_eval, T_w2c, sizeof(f.T_eval));
  so3_normalize(f.T_eval);   // worldToCam_evalPT is an SE3 (unit quaternion) in the reference
  const double init[10] = {0, 0, 0, 0, 0, 0, a, bb, 0, 0};  // setEvalPT_scaled (HessianBlocks.h:223-231)
  frame_set_state_scaled(f, init);
  frame_set_state_zero(f);
  // EFFrame::takeData -> FrameHessian::getPrior (HessianBlocks.h:246-268)
  const sdso_settings& S = ctx->S;
  for (int i = 0; i < 8; i++) f.prior[i] = 0;
  if (frameID == 0) {
    for (int i = 0; i < 3; i++) f.prior[i] = S.initialTransPrior;
    for (int i = 3; i < 6; i++) f.prior[i] = S.initialRotPrior;
    f.prior[6] = S.initialAffAPrior; f.prior[7] = S.initialAffBPrior;
  } else {
    f.prior[6] = S.affineOptModeA < 0 ? S.initialAffAPrior : S.affineOptModeA;
    f.prior[7] = S.affineOptModeB < 0 ? S.initialAffBPrior : S.affineOptModeB;
  }
  b->frames.push_back(f);
  b->n = (int)b->frames.size();
  b->prepared = false;
  if (idx_out) *idx_out = b->n - 1;
  return SDSO_OK;
}

int sdso_ba_set_state(sdso_ctx* ctx, int idx, const double state[10]) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (idx < 0 || idx >= b->n || !state) return SDSO_E_INVALID;
  frame_set_state(b->frames[idx], state);
  b->prepared = false;
  return SDSO_OK;
}

int sdso_ba_set_energy_th(sdso_ctx* ctx, int idx, float th) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (idx < 0 || idx >= b->n) return SDSO_E_INVALID;
  b->frames[idx].frameEnergyTH = th;
  b->prepared = false;
  return SDSO_OK;
}

int sdso_ba_set_points(sdso_ctx* ctx, int P, const int* host, const float* u, const float* v, const float* idepth, const float* idepth_zero,
                       const float* color8, const float* weights8, const unsigned char* has_prior) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  if (P < 0 || (P > 0 && (!host || !u || !v || !idepth || !idepth_zero || !color8 || !weights8))) return SDSO_E_INVALID;
  for (int i = 0; i < P; i++) if (host[i] < 0 || host[i] >= b->n) return fail(ctx, SDSO_E_INVALID, "point host index out of range (add the frames first)");
  if (P > b->capP) {
    const int cap = std::max(P + P / 2, 8192);   // geometric growth: a growing window must not reallocate at every key frame
    BA_ALLOC(b->d_p_host, cap); BA_ALLOC(b->d_p_u, cap); BA_ALLOC(b->d_p_v, cap); BA_ALLOC(b->d_p_idepth, cap); BA_ALLOC(b->d_p_idepth_zero, cap);
    BA_ALLOC(b->d_p_color, 2 * (size_t)cap); BA_ALLOC(b->d_p_weights, 2 * (size_t)cap); BA_ALLOC(b->d_p_priorF, cap); BA_ALLOC(b->d_p_deltaF, cap); BA_ALLOC(b->d_p_idepth_backup, cap);
    BA_ALLOC(b->d_p_res_begin, cap + 1); BA_ALLOC(b->d_slot_of, (size_t)cap * kMaxFrames); BA_ALLOC(b->d_p_acc, 16 * (size_t)cap);
    BA_ALLOC(b->d_p_flag, cap);
    const int pb = (cap + 127) / 128;
    BA_ALLOC(b->d_pblockpart, (size_t)pb * 32);
    b->capP = cap;
  }
  b->P = P;
  b->pblocks = (P + 127) / 128;
  b->h_p_host.assign(host, host + P);
  std::vector<float> prior(P);
  for (int i = 0; i < P; i++) prior[i] = (has_prior && has_prior[i]) ? ctx->S.idepthFixPrior * SCALE_IDEPTH * SCALE_IDEPTH : 0.f;  // EFPoint::takeData
  cudaStream_t st = ctx->stream;
  const size_t fb = (size_t)P * sizeof(float);
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_host, host, P * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_u, u, fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_v, v, fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_idepth, idepth, fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_idepth_zero, idepth_zero, fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_color, color8, 8 * fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_weights, weights8, 8 * fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_priorF, prior.data(), fb, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_p_acc, 0, 16 * (size_t)b->capP * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_p_flag, 0, (size_t)b->capP, st));
  if (P > 0) { ba_point_delta_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, b->d_p_idepth, b->d_p_idepth_zero, b->d_p_deltaF); SDSO_CHECK_LAUNCH(ctx); }
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  b->R = 0;
  b->prepared = false;
  return SDSO_OK;
}

int sdso_ba_set_residuals(sdso_ctx* ctx, int R, const int* point, const int* target) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  b->any_linearized = false;
  if (R < 0 || (R > 0 && (!point || !target))) return SDSO_E_INVALID;
  const int n = b->n, P = b->P;
  for (int i = 0; i < R; i++) if (point[i] < 0 || point[i] >= P || target[i] < 0 || target[i] >= n) return fail(ctx, SDSO_E_INVALID, "residual index out of range");
  // stable sort by key = host + target*n  (slot order)
  // (a counting sort: there are at most kMaxFrames^2 keys; stable, so the residuals of a key keep the caller's order)
  std::vector<int> key(R), order(R), seg(n * n + 1, 0);
  for (int i = 0; i < R; i++) { key[i] = b->h_p_host[point[i]] + target[i] * n; seg[key[i] + 1]++; }
  for (int k = 0; k < n * n; k++) seg[k + 1] += seg[k];
  {
    std::vector<int> pos(seg.begin(), seg.end() - 1);
    for (int i = 0; i < R; i++) order[pos[key[i]]++] = i;
  }
  b->h_slot2rid = order;
  b->h_rid2slot.assign(R, 0);
  for (int s = 0; s < R; s++) b->h_rid2slot[order[s]] = s;
  b->h_r_point.assign(point, point + R); b->h_r_target.assign(target, target + R);
  std::vector<int> s_point(R), s_key(R);
  for (int s = 0; s < R; s++) { s_point[s] = point[order[s]]; s_key[s] = key[order[s]]; }
  b->h_seg_begin = seg;
  // chunks of <= kChunk slots, never crossing a key boundary
  b->h_chunks.clear();
  std::vector<int> kcb(n * n + 1, 0);
  for (int k = 0; k < n * n; k++) {
    kcb[k] = (int)b->h_chunks.size();
    for (int s0 = seg[k]; s0 < seg[k + 1]; s0 += kChunk) b->h_chunks.push_back(Chunk{k, s0, std::min(s0 + kChunk, seg[k + 1]), 0});
  }
  kcb[n * n] = (int)b->h_chunks.size();
  b->nchunks = (int)b->h_chunks.size();
  // CSR point -> slots in residualsAll (= caller) order, and the (point,target) -> slot table
  std::vector<int> begin(P + 1, 0), list(R), slot_of((size_t)std::max(P, 1) * n, -1);
  for (int i = 0; i < R; i++) begin[point[i] + 1]++;
  for (int p = 0; p < P; p++) begin[p + 1] += begin[p];
  std::vector<int> fill(begin.begin(), begin.end() - 1);
  for (int i = 0; i < R; i++) {
    list[fill[point[i]]++] = b->h_rid2slot[i];
    int& so = slot_of[(size_t)point[i] * n + target[i]];
    if (so >= 0) return fail(ctx, SDSO_E_INVALID, "two residuals of one point towards the same target");
    so = b->h_rid2slot[i];
  }
  if (R > b->capR) {
    const int cap = std::max(R + R / 2, 32768);
    BA_ALLOC(b->d_p_res_list, cap); BA_ALLOC(b->d_s_point, cap); BA_ALLOC(b->d_s_key, cap);
    BA_ALLOC(b->d_s_state, cap); BA_ALLOC(b->d_s_newstate, cap); BA_ALLOC(b->d_s_flags, cap); BA_ALLOC(b->d_s_sel, cap);
    BA_ALLOC(b->d_s_energy, 3 * (size_t)cap); BA_ALLOC(b->d_J, 2 * (size_t)kJ * cap); BA_ALLOC(b->d_s_rtz, 8 * (size_t)cap);
    BA_ALLOC(b->d_s_JpJd, 8 * (size_t)cap); BA_ALLOC(b->d_s_center, 3 * (size_t)cap); BA_ALLOC(b->d_s_psum, 6 * (size_t)cap);
    BA_ALLOC(b->d_slot2rid, cap); BA_ALLOC(b->d_rid2slot, cap); BA_ALLOC(b->d_list, cap);
    BA_ALLOC(b->d_energy_part, (size_t)(cap + 127) / 128);
    b->capR = cap;
  }
  if (b->nchunks > b->capChunks) {
    const int cap = std::max(b->nchunks + b->nchunks / 2, 512);
    BA_ALLOC(b->d_chunks, cap); BA_ALLOC(b->d_tpart, (size_t)cap * kTopVals); BA_ALLOC(b->d_dpart, (size_t)cap * (kMaxFrames + 1) * 65);
    b->capChunks = cap;
  }
  if (!b->d_key_chunk_begin) BA_ALLOC(b->d_key_chunk_begin, kMaxFrames * kMaxFrames + 1);
  b->R = R;
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_s_point, s_point.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_s_key, s_key.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_res_list, list.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_res_begin, begin.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_slot_of, slot_of.data(), (size_t)P * n * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_slot2rid, b->h_slot2rid.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_rid2slot, b->h_rid2slot.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_chunks, b->h_chunks.data(), b->nchunks * sizeof(Chunk), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_key_chunk_begin, kcb.data(), kcb.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  // resetOOB (Residuals.h:107-115): state IN, new state OUTLIER, energy 0; EFResidual: not linearised, not active
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_state, RS_IN, b->capR, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_newstate, RS_OUTLIER, b->capR, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_flags, 0, b->capR, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_sel, 0, b->capR, st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_energy, 0, 3 * (size_t)b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_J, 0, 2 * (size_t)kJ * b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_rtz, 0, 8 * (size_t)b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_JpJd, 0, 8 * (size_t)b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_center, 0, 3 * (size_t)b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaMemsetAsync(b->d_s_psum, 0, 6 * (size_t)b->capR * sizeof(float), st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  return SDSO_OK;
}

int sdso_ba_set_point_flags(sdso_ctx* ctx, const unsigned char* flags) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (!flags) return SDSO_E_INVALID;
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_p_flag, flags, b->P, cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

int sdso_ba_prepare(sdso_ctx* ctx) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (b->n < 1) return fail(ctx, SDSO_E_STATE, "no frames in the window");
  return prepare_window(ctx);
}

int sdso_ba_counts(sdso_ctx* ctx, int* n, int* P, int* R, int* dim) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (n) *n = b->n;
  if (P) *P = b->P;
  if (R) *R = b->R;
  if (dim) *dim = b->dim();
  return SDSO_OK;
}

int sdso_ba_get_precalc(sdso_ctx* ctx, int h, int t, float out[49]) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (h < 0 || t < 0 || h >= b->n || t >= b->n || !out) return SDSO_E_INVALID;
  const PrecalcDev& q = b->h_pre[(size_t)h * b->n + t];
  int k = 0;
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_KRKiTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RKiTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RTll_0[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_tTll[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_KtTll[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_tTll_0[i];
  out[k++] = q.PRE_aff_mode[0]; out[k++] = q.PRE_aff_mode[1]; out[k++] = q.PRE_b0_mode; out[k++] = q.distanceLL;
  return SDSO_OK;
}

int sdso_ba_get_adjoints(sdso_ctx* ctx, double* adHost, double* adTarget, float* adHTdeltaF) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (adHost) memcpy(adHost, b->h_adH.data(), b->h_adH.size() * sizeof(double));
  if (adTarget) memcpy(adTarget, b->h_adT.data(), b->h_adT.size() * sizeof(double));
  if (adHTdeltaF) memcpy(adHTdeltaF, b->h_adHTd.data(), b->h_adHTd.size() * sizeof(float));
  return SDSO_OK;
}

int sdso_ba_nullspaces(sdso_ctx* ctx, double* N) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (!N) return SDSO_E_INVALID;
  memcpy(N, b->h_N.data(), b->h_N.size() * sizeof(double));
  return SDSO_OK;
}

int sdso_ba_linearize_all(sdso_ctx* ctx, int fix, double* energy) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  double e = 0;
  if (b->R > 0) {
    BAView v = view(b);
    ba_linearize_kernel<<<(b->R + 127) / 128, 128, 0, ctx->stream>>>(v, fix);
    SDSO_CHECK_LAUNCH(ctx);
    if (energy) {
      SDSO_CUDA(ctx, cudaMemcpyAsync(&e, b->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  if (energy) *energy = e;
  return SDSO_OK;
}

int sdso_ba_apply_res(sdso_ctx* ctx, int copy_jacobians) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (b->R == 0) return SDSO_OK;
  BAView v = view(b);
  ba_apply_res_kernel<<<(b->R + 127) / 128, 128, 0, ctx->stream>>>(v, copy_jacobians);
  SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}

int sdso_ba_fix_linearization(sdso_ctx* ctx, int count, const int* rids) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (!b->any_linearized) { b->any_linearized = true; invalidate_graphs(b); }   // the captured chains skip the linearised accumulation until now
  BAView v = view(b);
  if (!rids) {
    if (b->R == 0) return SDSO_OK;
    ba_fixlin_kernel<<<(b->R + 127) / 128, 128, 0, ctx->stream>>>(v, nullptr, b->R);
    SDSO_CHECK_LAUNCH(ctx);
    return SDSO_OK;
  }
  if (count <= 0) return SDSO_OK;
  std::vector<int> slots(count);
  for (int i = 0; i < count; i++) {
    if (rids[i] < 0 || rids[i] >= b->R) return SDSO_E_INVALID;
    slots[i] = b->h_rid2slot[rids[i]];
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(b->d_list, slots.data(), count * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ba_fixlin_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(v, b->d_list, count);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

// Readback in the CALLER's residual order. which: 0 = candidate J (PointFrameResidual::J), 1 = EFResidual::J.
int sdso_ba_get_res(sdso_ctx* ctx, int which, int* newState, int* state, double* newEnergy, double* newEnergyWO, int* active, int* linearized,
                    float* J74, float* JpJdF8, float* center3, float* resToZero8) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  const int R = b->R;
  const size_t cR = b->capR;
  if (R == 0) return SDSO_OK;
  std::vector<unsigned char> st(R), ns(R), fl(R), sel(R);
  std::vector<float> en(3 * cR);
  cudaStream_t s = ctx->stream;
  SDSO_CUDA(ctx, cudaMemcpyAsync(st.data(), b->d_s_state, R, cudaMemcpyDeviceToHost, s));
  SDSO_CUDA(ctx, cudaMemcpyAsync(ns.data(), b->d_s_newstate, R, cudaMemcpyDeviceToHost, s));
  SDSO_CUDA(ctx, cudaMemcpyAsync(fl.data(), b->d_s_flags, R, cudaMemcpyDeviceToHost, s));
  SDSO_CUDA(ctx, cudaMemcpyAsync(sel.data(), b->d_s_sel, R, cudaMemcpyDeviceToHost, s));
  SDSO_CUDA(ctx, cudaMemcpyAsync(en.data(), b->d_s_energy, 3 * cR * sizeof(float), cudaMemcpyDeviceToHost, s));
  std::vector<float> J, jp, ce, rz;
  if (J74) { J.resize(2 * (size_t)kJ * cR); SDSO_CUDA(ctx, cudaMemcpyAsync(J.data(), b->d_J, J.size() * sizeof(float), cudaMemcpyDeviceToHost, s)); }
  if (JpJdF8) { jp.resize(8 * cR); SDSO_CUDA(ctx, cudaMemcpyAsync(jp.data(), b->d_s_JpJd, jp.size() * sizeof(float), cudaMemcpyDeviceToHost, s)); }
  if (center3) { ce.resize(3 * cR); SDSO_CUDA(ctx, cudaMemcpyAsync(ce.data(), b->d_s_center, ce.size() * sizeof(float), cudaMemcpyDeviceToHost, s)); }
  if (resToZero8) { rz.resize(8 * cR); SDSO_CUDA(ctx, cudaMemcpyAsync(rz.data(), b->d_s_rtz, rz.size() * sizeof(float), cudaMemcpyDeviceToHost, s)); }
  SDSO_CUDA(ctx, cudaStreamSynchronize(s));
  for (int rid = 0; rid < R; rid++) {
    const int sl = b->h_rid2slot[rid];
    if (newState) newState[rid] = ns[sl];
    if (state) state[rid] = st[sl];
    if (newEnergy) newEnergy[rid] = en[cR + sl];
    if (newEnergyWO) newEnergyWO[rid] = en[2 * cR + sl];
    if (active) active[rid] = (fl[sl] & RF_ACTIVE) ? 1 : 0;
    if (linearized) linearized[rid] = (fl[sl] & RF_LINEARIZED) ? 1 : 0;
    if (J74) { const int buf = which == 0 ? (sel[sl] ^ 1) : sel[sl]; for (int k = 0; k < kJ; k++) J74[(size_t)rid * kJ + k] = J[((size_t)buf * kJ + k) * cR + sl]; }
    if (JpJdF8) for (int k = 0; k < 8; k++) JpJdF8[(size_t)rid * 8 + k] = jp[k * cR + sl];
    if (center3) for (int k = 0; k < 3; k++) center3[(size_t)rid * 3 + k] = ce[k * cR + sl];
    if (resToZero8) for (int k = 0; k < 8; k++) resToZero8[(size_t)rid * 8 + k] = rz[k * cR + sl];
  }
  return SDSO_OK;
}

// per point: Hdd_accAF, bd_accAF, Hcd_accAF[4], Hdd_accLF, bd_accLF, Hcd_accLF[4], HdiF, bdSumF, step, priorF (16 floats)
int sdso_ba_get_points(sdso_ctx* ctx, float* out16) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (!out16) return SDSO_E_INVALID;
  const int P = b->P;
  const size_t cP = b->capP;
  if (P == 0) return SDSO_OK;
  std::vector<float> acc(16 * cP), pr(P);
  SDSO_CUDA(ctx, cudaMemcpyAsync(acc.data(), b->d_p_acc, acc.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaMemcpyAsync(pr.data(), b->d_p_priorF, P * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int p = 0; p < P; p++) {
    for (int k = 0; k < 15; k++) out16[(size_t)p * 16 + k] = acc[k * cP + p];
    out16[(size_t)p * 16 + 15] = pr[p];
  }
  return SDSO_OK;
}

int sdso_ba_accumulate_top(sdso_ctx* ctx, int mode, int use_prior, double* H, double* bv, float* blocks) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (mode < 0 || mode > 2) return SDSO_E_INVALID;
  int rc = launch_top(ctx, mode, SYS_TMP, use_prior != 0);
  if (rc) return rc;
  if (blocks) SDSO_CUDA(ctx, cudaMemcpyAsync(blocks, b->d_Gf, (size_t)b->n * b->n * 169 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  return download_sys(ctx, SYS_TMP, H, bv);
}

int sdso_ba_accumulate_sc(sdso_ctx* ctx, int shift_prior_to_zero, double* H, double* bv) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  int rc = launch_sc(ctx, shift_prior_to_zero != 0, SYS_TMP);
  if (rc) return rc;
  return download_sys(ctx, SYS_TMP, H, bv);
}

int sdso_ba_solve(sdso_ctx* ctx, int iteration, double lambda, double* x, double* Hfinal, double* bfinal) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  int rc = launch_solve(ctx, iteration, lambda);
  if (rc) return rc;
  rc = launch_resub(ctx, sysb(b, SYS_X));  // solveSystemF ends with resubstituteF_MT (:989)
  if (rc) return rc;
  const int d = b->dim();
  if (x) SDSO_CUDA(ctx, cudaMemcpyAsync(x, sysb(b, SYS_X), d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return download_sys(ctx, SYS_FINAL, Hfinal, bfinal);
}

int sdso_ba_resubstitute(sdso_ctx* ctx, const double* x, double* frame_steps, double* calib_step) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  const int d = b->dim(), n = b->n;
  std::vector<double> xv(d);
  if (x) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(sysb(b, SYS_X), x, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    memcpy(xv.data(), x, d * sizeof(double));
  } else {
    SDSO_CUDA(ctx, cudaMemcpyAsync(xv.data(), sysb(b, SYS_X), d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int rc = launch_resub(ctx, sysb(b, SYS_X));
  if (rc) return rc;
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (calib_step) for (int i = 0; i < 4; i++) calib_step[i] = -xv[i];  // HCalib->step = -x.head<CPARS>() (:280)
  if (frame_steps) for (int h = 0; h < n; h++) {
    for (int i = 0; i < 8; i++) frame_steps[h * 10 + i] = -xv[kCPARS + 8 * h + i];
    frame_steps[h * 10 + 8] = frame_steps[h * 10 + 9] = 0;
  }
  return SDSO_OK;
}

/* ---- point-sharded windows ---- */
int sdso_ba_set_shard(sdso_ctx* ctx, int rank, int nranks) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  if (nranks < 1 || rank < 0 || rank >= nranks) return SDSO_E_INVALID;
  b->shard_rank = rank; b->shard_n = nranks;
  return SDSO_OK;
}

int sdso_ba_assemble(sdso_ctx* ctx, void** device_system, int* count) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  int rc = launch_assemble(ctx);
  if (rc) return rc;
  const int d = b->dim();
  if (device_system) *device_system = sysH(b, SYS_FINAL);
  if (count) *count = d * d + d;
  return SDSO_OK;
}

// one NCCL allreduce of the partial system [(4+8n)^2 + (4+8n)] plus the linearisation energy (scalars[0]) appended to it
int sdso_ba_allreduce(sdso_ctx* ctx, double* energy_out) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  const int d = b->dim();
  double* buf = sysH(b, SYS_FINAL);
  SDSO_CUDA(ctx, cudaMemcpyAsync(buf + (size_t)d * d + d, b->d_scalars, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  int rc = sdso_allreduce_f64(ctx, buf, d * d + d + 1);
  if (rc) return rc;
  if (energy_out) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(energy_out, buf + (size_t)d * d + d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SDSO_OK;
}

int sdso_ba_solve_assembled(sdso_ctx* ctx, int iteration, double* x, double* Hfinal, double* bfinal) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  int rc = launch_factor_solve(ctx, iteration);
  if (!rc) rc = launch_resub(ctx, sysb(b, SYS_X));
  if (rc) return rc;
  const int d = b->dim();
  if (x) SDSO_CUDA(ctx, cudaMemcpyAsync(x, sysb(b, SYS_X), d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return download_sys(ctx, SYS_FINAL, Hfinal, bfinal);
}

// CalibHessian::setValue (HessianBlocks.h:316-331)
static void set_calib_value(sdso_ctx* ctx, const double v[4]) {
  BAState* b = ctx->ba;
  BACalib& c = b->calib;
  for (int i = 0; i < 4; i++) b->calib_value[i] = v[i];
  const double vs[4] = {SCALE_F * v[0], SCALE_F * v[1], SCALE_C * v[2], SCALE_C * v[3]};
  c.fxl = (float)vs[0]; c.fyl = (float)vs[1]; c.cxl = (float)vs[2]; c.cyl = (float)vs[3];
  c.fxli = 1.0f / c.fxl; c.fyli = 1.0f / c.fyl; c.cxli = -c.cxl / c.fxl; c.cyli = -c.cyl / c.fyl;
  for (int i = 0; i < 4; i++) b->calib_delta[i] = b->calib_value[i] - b->calib_zero[i];
}

static int launch_energy_th(sdso_ctx* ctx, float* th_host, bool readback = true) {
  BAState* b = ctx->ba;
  BAView v = view(b);
  const sdso_settings& S = ctx->S;
  float* d_out = reinterpret_cast<float*>(b->d_scalars + 4);
  ba_energy_th_kernel<<<1, 1024, 0, ctx->stream>>>(v, b->n - 1, S.frameEnergyTHN, S.frameEnergyTHFacMedian, S.frameEnergyTHConstWeight,
                                                    S.overallEnergyTHWeight, b->d_frameTH, d_out);
  SDSO_CHECK_LAUNCH(ctx);
  if (!readback) return SDSO_OK;   // inside the device-resident LM loop: the host mirror is refreshed once, after the loop
  float th = 0;
  SDSO_CUDA(ctx, cudaMemcpyAsync(&th, d_out, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->frames[b->n - 1].frameEnergyTH = th;  // host mirror (prepare re-uploads the thresholds)
  if (th_host) *th_host = th;
  return SDSO_OK;
}

int sdso_ba_new_frame_energy_th(sdso_ctx* ctx, float* th) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  return launch_energy_th(ctx, th);
}

int sdso_ba_get_energy_th(sdso_ctx* ctx, float* frameEnergyTH) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  if (!frameEnergyTH) return SDSO_E_INVALID;
  for (int h = 0; h < b->n; h++) frameEnergyTH[h] = b->frames[h].frameEnergyTH;
  return SDSO_OK;
}

int sdso_ba_get_state(sdso_ctx* ctx, double* states, double* T_w2c, float* idepth, double* calib) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  for (int h = 0; h < b->n; h++) {
    if (states) memcpy(states + 10 * h, b->frames[h].state, 10 * sizeof(double));
    if (T_w2c) memcpy(T_w2c + 12 * h, b->frames[h].T_w2c, 12 * sizeof(double));
  }
  if (calib) { calib[0] = b->calib.fxl; calib[1] = b->calib.fyl; calib[2] = b->calib.cxl; calib[3] = b->calib.cyl; }
  if (idepth && b->P > 0) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(idepth, b->d_p_idepth, b->P * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return SDSO_OK;
}

int sdso_ba_optimize(sdso_ctx* ctx, int mnumOptIts, double* rmse, int* iterations_done) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  const int n = b->n, d = b->dim(), R = b->R, P = b->P;
  if (rmse) *rmse = 0;
  if (iterations_done) *iterations_done = 0;
  if (n < 2) return SDSO_OK;
  if (n < 3) mnumOptIts = 20;
  if (n < 4) mnumOptIts = 15;
  cudaStream_t st = ctx->stream;
  const int rb = (R + 127) / 128, pb = (P + 127) / 128;
  double* d_part = b->d_energy_part;  // reused: [pb][2] doubles (pb <= capacity of the energy partials: one per 128 slots... sized below)
  int rc = SDSO_OK;
  auto lin = [&](int fix) -> int {  // linearizeAll + setNewFrameEnergyTH (FullSystemOptimize.cpp:142-163)
    if (R > 0) { BAView v = view(b); ba_linearize_kernel<<<rb, 128, 0, st>>>(v, fix); SDSO_CHECK_LAUNCH(ctx); }
    return launch_energy_th(ctx, nullptr);
  };
  auto apply = [&]() -> int {
    if (R > 0) { BAView v = view(b); ba_apply_res_kernel<<<rb, 128, 0, st>>>(v, 1); SDSO_CHECK_LAUNCH(ctx); }
    return SDSO_OK;
  };
  if (R > 0) { BAView v = view(b); ba_reset_oob_kernel<<<rb, 128, 0, st>>>(v); SDSO_CHECK_LAUNCH(ctx); }
  if ((rc = lin(0))) return rc;
  if ((rc = apply())) return rc;
  if (pb > b->step_part_cap) {
    if (b->d_step_part) cudaFree(b->d_step_part);
    b->d_step_part = nullptr;
    SDSO_CUDA(ctx, cudaMalloc(&b->d_step_part, (size_t)(pb + 1) * 2 * sizeof(double)));
    b->step_part_cap = pb;
  }
  (void)d_part; (void)d;
  // ---- the LM loop, device-resident: every iteration is ONE graph launch; the convergence decision is latched on the device
  // (OptDev::done) and turns the iterations enqueued behind it into no-ops, so the host enqueues all mnumOptIts blindly and
  // synchronises once, after the loop.
  if (!b->graph_iter) {
    Capture c;
    if ((rc = capture_begin(ctx, c))) return rc;
    const uint64_t l0 = ctx->launches;
    cudaStream_t cs = ctx->stream;
    auto body = [&]() -> int {
      int r = SDSO_OK;
      if (P > 0) { BAView v = view(b); ba_backup_points_kernel<<<pb, 128, 0, cs>>>(v); SDSO_CHECK_LAUNCH(ctx); }   // backupState (:309-350)
      // solveSystem(iteration, lambda) (:1045-1053): accumulate, stitch, solve, back-substitute
      if ((r = launch_assemble_chain(ctx))) return r;
      if ((r = launch_factor_solve(ctx, -1))) return r;   // iteration index from OptDev::it
      if ((r = launch_resub(ctx, sysb(b, SYS_X)))) return r;
      // doStepFromBackup(1,1,1,1,1) (:207-305): points, then frames + calibration + precalc + convergence test
      if (P > 0) {
        BAView v = view(b);
        ba_step_points_kernel<<<pb, 128, 0, cs>>>(v, 1.0f, b->d_step_part); SDSO_CHECK_LAUNCH(ctx);
        ba_sum_pairs_kernel<<<1, 32, 0, cs>>>(b->d_step_part, pb, b->d_step_part + 2 * (size_t)pb, &b->d_opt->done); SDSO_CHECK_LAUNCH(ctx);
      }
      FrameUpdateParams U;
      U.n = n; U.P = P; U.frames = b->d_frames; U.opt = b->d_opt; U.calib = b->d_calib; U.cDeltaF = b->d_cDeltaF; U.fprior = b->d_fprior; U.precalc = b->d_precalc;
      U.adHostF = b->d_adHostF; U.adTargetF = b->d_adTargetF; U.adHTdeltaF = b->d_adHTdeltaF; U.x = sysb(b, SYS_X); U.step_sums = b->d_step_part + 2 * (size_t)pb;
      ba_frame_update_kernel<<<1, 256, 0, cs>>>(U); SDSO_CHECK_LAUNCH(ctx);
      // linearizeAll(false) + setNewFrameEnergyTH + applyRes: setting_forceAceptStep, every step is accepted (:965-978)
      if (R > 0) { BAView v = view(b); ba_linearize_kernel<<<rb, 128, 0, cs>>>(v, 0); SDSO_CHECK_LAUNCH(ctx); }
      if ((r = launch_energy_th(ctx, nullptr, false))) return r;
      if (R > 0) { BAView v = view(b); ba_apply_res_kernel<<<rb, 128, 0, cs>>>(v, 1); SDSO_CHECK_LAUNCH(ctx); }
      ba_iter_end_kernel<<<1, 1, 0, cs>>>(b->d_opt); SDSO_CHECK_LAUNCH(ctx);
      return SDSO_OK;
    };
    rc = body();
    b->graph_iter_nodes = (int)(ctx->launches - l0);
    ctx->launches = l0;
    if ((rc = capture_end(ctx, c, &b->graph_iter, rc))) return rc;
  }
  {
    // reset the loop's control block (it, pending, done, its_done) — the calibration values in front of it stay
    SDSO_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(b->d_opt) + offsetof(OptDev, it), 0, 4 * sizeof(int), st));
    for (int k = 0; k < mnumOptIts; k++) {
      SDSO_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)b->graph_iter, st));
      ctx->launches += b->graph_iter_nodes;
    }
  }
  // host mirrors of what the loop moved (ONE synchronisation for the whole loop)
  int it = 0;
  {
    std::vector<FrameDev> fd(kMaxFrames);
    OptDev od;
    BACalib cal;
    std::vector<float> th(kMaxFrames);
    SDSO_CUDA(ctx, cudaMemcpyAsync(fd.data(), b->d_frames, fd.size() * sizeof(FrameDev), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(&od, b->d_opt, sizeof(OptDev), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(&cal, b->d_calib, sizeof(BACalib), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(th.data(), b->d_frameTH, th.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
    it = od.its_done;
    for (int h = 0; h < n; h++) {
      HostBAFrame& f = b->frames[h];
      memcpy(f.state_backup, fd[h].state_backup, sizeof(f.state_backup));
      for (int i = 0; i < 10; i++) f.step[i] = fd[h].state[i] - fd[h].state_backup[i];
      frame_set_state(f, fd[h].state);   // state, state_scaled, PRE_worldToCam / PRE_camToWorld (host arithmetic, as prepare_window uses them)
      f.frameEnergyTH = th[h];
    }
    for (int i = 0; i < 4; i++) { b->calib_value[i] = od.calib_value[i]; b->calib_backup[i] = od.calib_backup[i]; b->calib_step[i] = od.calib_value[i] - od.calib_backup[i];
                                  b->calib_delta[i] = b->calib_value[i] - b->calib_zero[i]; }
    b->calib = cal;
    // the loop is over: clear the latch so that the kernels of the tail below (and of later operator calls) run
    SDSO_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(b->d_opt) + offsetof(OptDev, it), 0, 4 * sizeof(int), st));
  }
  if (iterations_done) *iterations_done = it;
  // new evaluation point of the newest frame (:996-1005): setEvalPT(PRE_worldToCam, [0.., a, b, 0, 0])
  {
    HostBAFrame& nw = b->frames[n - 1];
    const double nz[10] = {0, 0, 0, 0, 0, 0, nw.state[6], nw.state[7], 0, 0};
    memcpy(nw.T_eval, nw.T_w2c, sizeof(nw.T_eval));
    frame_set_state(nw, nz);
    frame_set_state_zero(nw);
  }
  if ((rc = prepare_window(ctx))) return rc;  // setAdjointsF + setPrecalcValues
  if ((rc = lin(1))) return rc;
  double E = 0; unsigned resInA = 0;
  if (R > 0) {
    BAView v = view(b);
    SDSO_CUDA(ctx, cudaMemsetAsync(b->d_counter + 1, 0, sizeof(unsigned), st));
    ba_count_active_kernel<<<rb, 128, 0, st>>>(v, b->d_counter + 1); SDSO_CHECK_LAUNCH(ctx);
    ba_drop_inactive_kernel<<<rb, 128, 0, st>>>(v); SDSO_CHECK_LAUNCH(ctx);
    SDSO_CUDA(ctx, cudaMemcpyAsync(&E, b->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(&resInA, b->d_counter + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if (rmse) *rmse = sqrtf((float)(E / (8 * (double)(resInA > 0 ? resInA : 1))));
  return SDSO_OK;
}

int sdso_ba_set_marg_prior(sdso_ctx* ctx, const double* HM, const double* bM) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  invalidate_graphs(b);
  if (!HM || !bM) return SDSO_E_INVALID;
  const int d = b->dim();
  SDSO_CUDA(ctx, cudaMemcpyAsync(sysH(b, SYS_M), HM, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaMemcpyAsync(sysb(b, SYS_M), bM, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->have_M = true;
  return SDSO_OK;
}

// D4: FullSystem::optimizeImmaturePoint for n candidates hosted in frames of the uploaded window (variant: SDSO_VARIANT_SSE =
// the original 3-iteration LM on the inverse depth, SDSO_VARIANT_G2O = the live frozen-projection semantics)
int sdso_activate_points(sdso_ctx* ctx, int n, const int* host, const sdso_immature_point* pts, int variant, int min_obs, int* result,
                         float* idepth, int* states, float* energy) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (n < 0 || (n > 0 && (!host || !pts || !result || !idepth || !states || !energy))) return SDSO_E_INVALID;
  if (variant != SDSO_VARIANT_SSE && variant != SDSO_VARIANT_G2O) return SDSO_E_INVALID;
  if (n == 0) return SDSO_OK;
  const int nf = b->n;
  for (int i = 0; i < n; i++) if (host[i] < 0 || host[i] >= nf) return SDSO_E_INVALID;
  cudaStream_t st = ctx->stream;
  const size_t bytes = (size_t)n * (sizeof(sdso_immature_point) + sizeof(int) * 2 + sizeof(float) * 2 + sizeof(int) * nf);
  unsigned char* d = nullptr;
  SDSO_CUDA(ctx, cudaMalloc(&d, bytes + 64));
  ActParams A{};
  sdso_immature_point* d_pts = reinterpret_cast<sdso_immature_point*>(d);
  int* d_host = reinterpret_cast<int*>(d_pts + n);
  int* d_res = d_host + n; int* d_states = d_res + n;
  float* d_id = reinterpret_cast<float*>(d_states + (size_t)n * nf); float* d_en = d_id + n;
  cudaMemcpyAsync(d_pts, pts, (size_t)n * sizeof(sdso_immature_point), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_host, host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st);
  A.pts = d_pts; A.host = d_host; A.n = n; A.variant = variant; A.minObs = min_obs; A.GNIts = 3; A.minIdepthH_act = 100;  // settings.cpp:114, :56
  A.result = d_res; A.idepth = d_id; A.states = d_states; A.energy = d_en;
  BAView v = view(b);
  ba_activate_kernel<<<(n + 63) / 64, 64, 0, st>>>(v, A);
  ctx->launches++;
  cudaError_t le = cudaGetLastError();
  cudaMemcpyAsync(result, d_res, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(states, d_states, (size_t)n * nf * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(idepth, d_id, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(energy, d_en, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(d);
  if (le != cudaSuccess) return fail(ctx, SDSO_E_CUDA, cudaGetErrorString(le));
  if (se != cudaSuccess) return fail(ctx, SDSO_E_CUDA, cudaGetErrorString(se));
  return SDSO_OK;
}

// E2 at operator level: every residual of the uploaded window evaluated as one EdgeLBASE3PosePhotoIdepthCamDSO
// (dso_g2o_edge.cpp:5-282) with the given vertex estimates. Outputs are in the caller's residual order.
int sdso_lba_edge_eval(sdso_ctx* ctx, const double* T_wh, const double* photo, const double* idepth, const double cam[4], const double* b0,
                       double* error8, double* J_xi, double* J_photo, double* J_idepth, double* J_C, int* newState, double* newEnergy,
                       double* newEnergyWithOutlier, float* center3, float* idepth_hessian, int* level) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (!T_wh || !photo || !idepth || !cam || !b0 || !error8 || !J_xi || !J_photo || !J_idepth || !J_C || !newState || !newEnergy ||
      !newEnergyWithOutlier || !center3 || !idepth_hessian || !level) return SDSO_E_INVALID;
  const int n = b->n, R = b->R;
  if (R == 0) return SDSO_OK;
  cudaStream_t st = ctx->stream;
  // one scratch allocation: inputs then outputs
  const size_t in_d = (size_t)n * 12 * 2 + (size_t)n * 2 + R + n + (size_t)n * 2;   // doubles
  const size_t out_d = (size_t)R * (8 + 48 + 16 + 8 + 32 + 2);                      // doubles
  const size_t bytes = (in_d + out_d) * sizeof(double) + (size_t)R * (3 + 1) * sizeof(float) + (size_t)R * 2 * sizeof(int) + (size_t)n * sizeof(float) + 64;
  unsigned char* d = nullptr;
  SDSO_CUDA(ctx, cudaMalloc(&d, bytes));
  double* pd = reinterpret_cast<double*>(d);
  LBAEdgeParams E{};
  std::vector<double> host_in(in_d);
  double* hp = host_in.data();
  double* h_Twh = hp; hp += n * 12;
  double* h_Ttw = hp; hp += n * 12;
  double* h_photo = hp; hp += n * 2;
  double* h_id = hp; hp += R;
  double* h_b0 = hp; hp += n;
  double* h_taff = hp; hp += n * 2;
  memcpy(h_Twh, T_wh, sizeof(double) * n * 12); memcpy(h_photo, photo, sizeof(double) * n * 2); memcpy(h_id, idepth, sizeof(double) * R); memcpy(h_b0, b0, sizeof(double) * n);
  std::vector<float> h_exp(n);
  for (int i = 0; i < n; i++) {
    memcpy(h_Ttw + 12 * i, b->frames[i].T_w2c, sizeof(double) * 12);
    h_taff[2 * i] = b->frames[i].state_scaled[6]; h_taff[2 * i + 1] = b->frames[i].state_scaled[7];
    h_exp[i] = b->frames[i].ab_exposure;
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(pd, host_in.data(), in_d * sizeof(double), cudaMemcpyHostToDevice, st));
  E.T_wh = pd; E.T_tw = pd + n * 12; E.photo = pd + n * 24; E.idepth = pd + n * 26; E.b0 = pd + n * 26 + R; E.target_aff = pd + n * 27 + R;
  double* po = pd + in_d;
  E.error8 = po; po += (size_t)R * 8; E.Jxi = po; po += (size_t)R * 48; E.Jphoto = po; po += (size_t)R * 16; E.Jid = po; po += (size_t)R * 8;
  E.JC = po; po += (size_t)R * 32; E.newEnergy = po; po += R; E.newEnergyWO = po; po += R;
  float* pf = reinterpret_cast<float*>(po);
  E.center3 = pf; pf += (size_t)R * 3; E.idepth_hessian = pf; pf += R;
  float* d_exp = pf; pf += n;
  int* pi = reinterpret_cast<int*>(pf);
  E.newState = pi; pi += R; E.level = pi;
  E.exposure = d_exp;
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_exp, h_exp.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
  for (int i = 0; i < 4; i++) E.cam[i] = cam[i];
  E.slot2rid = b->d_slot2rid; E.driver = 0; E.active = nullptr; E.linearize = 1;
  BAView v = view(b);
  ba_lba_edge_kernel<<<(R + 127) / 128, 128, 0, st>>>(v, E);
  ctx->launches++;
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) { cudaFree(d); return fail(ctx, SDSO_E_CUDA, cudaGetErrorString(le)); }
  cudaMemcpyAsync(error8, E.error8, sizeof(double) * R * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(J_xi, E.Jxi, sizeof(double) * R * 48, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(J_photo, E.Jphoto, sizeof(double) * R * 16, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(J_idepth, E.Jid, sizeof(double) * R * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(J_C, E.JC, sizeof(double) * R * 32, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(newEnergy, E.newEnergy, sizeof(double) * R, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(newEnergyWithOutlier, E.newEnergyWO, sizeof(double) * R, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(center3, E.center3, sizeof(float) * R * 3, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(idepth_hessian, E.idepth_hessian, sizeof(float) * R, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(newState, E.newState, sizeof(int) * R, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(level, E.level, sizeof(int) * R, cudaMemcpyDeviceToHost, st);
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(d);
  if (se != cudaSuccess) return fail(ctx, SDSO_E_CUDA, cudaGetErrorString(se));
  return SDSO_OK;
}

// FullSystem::optimize, g2o body (FullSystemOptimize.cpp:404-868): the graph of E2 edges over {cam, pose+photo per host frame, one
// marginalised idepth vertex per residual}, Huber(9) per edge, driven by the restated g2o Levenberg-Marquardt with Schur complement
// (g2o is not in the reference tree, so the driver is unpinned). All per-edge work runs on the device; the
// host loop moves the (4+8n)-vector increment and the n host poses per trial.
int sdso_lba_g2o(sdso_ctx* ctx, int mnumOptIts, double cam[4], double* T_wh, double* photo, double* idepth, int* used_host, double* chi2_out,
                 int* newState, float* center3, float* idepth_hessian, int* iterations_out, int* trials_out) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  if (!cam || !T_wh || !photo || !idepth || !used_host) return SDSO_E_INVALID;
  const int n = b->n, R = b->R, d = b->dim();
  if (iterations_out) *iterations_out = 0;
  if (trials_out) *trials_out = 0;
  if (n < 2 || R == 0) return SDSO_OK;
  if (n < 3) mnumOptIts = 10; else if (n < 4) mnumOptIts = 7; else mnumOptIts = 3;
  cudaStream_t st = ctx->stream;
  const int rb = (R + 127) / 128;
  // ---- one scratch allocation for the graph
  const size_t nd = (size_t)R * (1 + 1 + 8 + 48 + 16 + 8 + 32 + 1 + 1 + 12 + 2)   // idepth, bak, err, J blocks, hll, bl, hpl, energies
                    + (size_t)n * (12 + 12 + 2 + 1 + 2) + 4                        // T_wh, T_tw, photo, b0, target aff, cam
                    + (size_t)n * 14 + 4                                           // push() copy of T_wh, photo, cam
                    + (size_t)(b->nchunks + 1) * 96 * 2 + (size_t)n * 96 * 2 + rb + 64;
  const size_t bytes = nd * sizeof(double) + (size_t)R * (4 * sizeof(float) + 2 * sizeof(int) + 2) + (size_t)n * (sizeof(float) + sizeof(int)) + 256;
  if (bytes > b->lba_scratch_bytes) {   // kept across calls: cudaMalloc / cudaFree per call are device-wide synchronisations
    if (b->lba_scratch) cudaFree(b->lba_scratch);
    b->lba_scratch = nullptr; b->lba_scratch_bytes = 0;
    SDSO_CUDA(ctx, cudaMalloc(&b->lba_scratch, bytes));
    b->lba_scratch_bytes = bytes;
  }
  unsigned char* raw = reinterpret_cast<unsigned char*>(b->lba_scratch);
  SDSO_CUDA(ctx, cudaMemsetAsync(raw, 0, bytes, st));
  double* pd = reinterpret_cast<double*>(raw);
  auto takeD = [&](size_t k) { double* q = pd; pd += k; return q; };
  double* d_idepth = takeD(R); double* d_idbak = takeD(R); double* d_err = takeD((size_t)R * 8);
  double* d_Jxi = takeD((size_t)R * 48); double* d_Jph = takeD((size_t)R * 16); double* d_Jid = takeD((size_t)R * 8); double* d_JC = takeD((size_t)R * 32);
  double* d_hll = takeD(R); double* d_bl = takeD(R); double* d_hpl = takeD((size_t)R * 12); double* d_ne = takeD(R); double* d_newo = takeD(R);
  double* d_est = takeD((size_t)n * 29 + 4); double* d_estbak = takeD((size_t)n * 14 + 4);
  double* d_partA = takeD((size_t)(b->nchunks + 1) * 96); double* d_partS = takeD((size_t)(b->nchunks + 1) * 96);
  double* d_hostA = takeD((size_t)n * 96); double* d_hostS = takeD((size_t)n * 96);
  double* d_bpart = takeD(rb); double* d_sc = takeD(64);
  float* pf = reinterpret_cast<float*>(pd);
  float* d_center = pf; pf += (size_t)R * 3; float* d_ih = pf; pf += R; float* d_exp = pf; pf += n;
  int* pi = reinterpret_cast<int*>(pf);
  int* d_ns = pi; pi += R; int* d_level = pi; pi += R; int* d_used = pi; pi += n;
  unsigned char* d_active = reinterpret_cast<unsigned char*>(pi); unsigned char* d_ingraph = d_active + R;
  auto cleanup = [&](int rc) { cudaStreamSynchronize(st); return rc; };

  // ---- graph build on the host side: active residuals (not linearised, not dropped), used hosts, b0 (:438-542)
  std::vector<unsigned char> flags(R);
  SDSO_CUDA(ctx, cudaMemcpyAsync(flags.data(), b->d_s_flags, R, cudaMemcpyDeviceToHost, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  std::vector<unsigned char> ingraph(R, 0);
  std::vector<double> id_slot(R);
  std::vector<int> ns_init(R, RS_OUTLIER);
  for (int h = 0; h < n; h++) used_host[h] = 0;
  std::vector<double> b0(n, 0.0);
  for (int rid = 0; rid < R; rid++) {
    const int sl = b->h_rid2slot[rid];
    id_slot[sl] = idepth[rid];
    if (flags[sl] & (RF_LINEARIZED | RF_DROPPED)) continue;
    ingraph[sl] = 1;
    const int host = b->h_p_host[b->h_r_point[rid]];
    if (!used_host[host]) { used_host[host] = 1; b0[host] = photo[2 * host + 1]; }
  }
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_ingraph, ingraph.data(), R, cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_idepth, id_slot.data(), R * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_ns, ns_init.data(), R * sizeof(int), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_used, used_host, n * sizeof(int), cudaMemcpyHostToDevice, st));
  // the vertex estimates live on the device from here on (uploaded once; the trials update them there)
  std::vector<float> h_exp(n);
  std::vector<double> est((size_t)n * 29 + 4);
  {
    double* e = est.data();
    memcpy(e, T_wh, sizeof(double) * n * 12);
    for (int i = 0; i < n; i++) memcpy(e + n * 12 + 12 * i, b->frames[i].T_w2c, sizeof(double) * 12);
    memcpy(e + n * 24, photo, sizeof(double) * n * 2);
    memcpy(e + n * 26, b0.data(), sizeof(double) * n);
    for (int i = 0; i < n; i++) { e[n * 27 + 2 * i] = b->frames[i].state_scaled[6]; e[n * 27 + 2 * i + 1] = b->frames[i].state_scaled[7]; }
    for (int k = 0; k < 4; k++) e[n * 29 + k] = cam[k];
    SDSO_CUDA(ctx, cudaMemcpyAsync(d_est, e, est.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  for (int i = 0; i < n; i++) h_exp[i] = b->frames[i].ab_exposure;
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_exp, h_exp.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));

  BAView v = view(b);
  LBAEdgeParams E{};
  E.T_wh = d_est; E.T_tw = d_est + n * 12; E.photo = d_est + n * 24; E.b0 = d_est + n * 26; E.target_aff = d_est + n * 27; E.exposure = d_exp;
  E.cam_dev = d_est + n * 29;
  E.idepth = d_idepth; E.slot2rid = nullptr; E.driver = 1;
  E.error8 = d_err; E.Jxi = d_Jxi; E.Jphoto = d_Jph; E.Jid = d_Jid; E.JC = d_JC;
  E.newState = d_ns; E.newEnergy = d_ne; E.newEnergyWO = d_newo; E.center3 = d_center; E.idepth_hessian = d_ih; E.level = d_level;
  LBAGraph G;
  G.R = R; G.active = d_active; G.idepth = d_idepth; G.idepth_bak = d_idbak; G.err = d_err; G.Jxi = d_Jxi; G.Jph = d_Jph; G.Jid = d_Jid; G.JC = d_JC;
  G.hll = d_hll; G.bl = d_bl; G.hpl = d_hpl; G.delta = ctx->S.huberTH;
  auto eval = [&](const unsigned char* active, int linearize, const double* run_if) -> int {
    E.active = active; E.linearize = linearize; E.run_if = run_if;
    ba_lba_edge_kernel<<<rb, 128, 0, st>>>(v, E); SDSO_CHECK_LAUNCH(ctx);
    return SDSO_OK;
  };
  auto chi2_to = [&](int slot) -> int {   // activeRobustChi2 into the scalar block, no host round trip
    lba_chi2_kernel<<<rb, 128, 0, st>>>(G, d_bpart); SDSO_CHECK_LAUNCH(ctx);
    lba_sum_kernel<<<1, 32, 0, st>>>(d_bpart, rb, 1, 1, d_sc + slot); SDSO_CHECK_LAUNCH(ctx);
    return SDSO_OK;
  };
  double hsc[LS_NUM];
  auto read_scalars = [&]() -> int {      // the ONE host round trip of a damping trial
    SDSO_CUDA(ctx, cudaMemcpyAsync(hsc, d_sc, sizeof(hsc), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
    return SDSO_OK;
  };
  int rc;
  // first computeError of every edge while the graph is built (:538), then initializeOptimization(): level-0 edges are active
  if ((rc = eval(d_ingraph, 0, nullptr))) return cleanup(rc);
  lba_activate_kernel<<<rb, 128, 0, st>>>(R, d_ingraph, d_level, d_active); ctx->launches++;

  SolveParams S{};
  fill_solve_params(ctx, S, 0);
  S.plain = 1; S.N = nullptr; S.have_M = 0;
  const size_t smem = ((size_t)d * (d | 1) + 6 * (size_t)d + 7 * (size_t)d + 256) * sizeof(double);
  double lambda = 0, ni = 2, lastChi = 0, currentChi = 0;
  int it = 0, trials = 0;
  bool stop = false, have_current = false;
  for (; it < mnumOptIts && !stop; it++) {
    // computeActiveErrors + activeRobustChi2. From the second iteration on the edges were last evaluated at exactly this estimate
    // (behind the accepted trial, or by the terminate action's pass after a rejected one) and currentChi is that pass's chi2.
    if (!have_current) {
      if ((rc = eval(d_active, 0, nullptr))) return cleanup(rc);
      if ((rc = chi2_to(LS_CUR))) return cleanup(rc);
    }
    if ((rc = eval(d_active, 1, nullptr))) return cleanup(rc);           // buildSystem: linearizeOplus ...
    if (b->nchunks > 0) { lba_build_kernel<<<b->nchunks, kChunk, 0, st>>>(v, G, 0, 0.0, d_partA); SDSO_CHECK_LAUNCH(ctx); }
    lba_host_sum_kernel<<<n, 96, 0, st>>>(v, d_partA, d_hostA); SDSO_CHECK_LAUNCH(ctx);
    if (it == 0) { lambda = 0.1; ni = 2; }                       // setUserLambdaInit(0.1) (:425)
    double rho = 0, tempChi = 0;
    int qmax = 0;
    bool accepted = false;
    do {
      // push() of the inverse depths (the vertices are pushed by lba_trial_update_kernel)
      lba_copy_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, d_idepth, d_idbak); SDSO_CHECK_LAUNCH(ctx);
      if (b->nchunks > 0) { lba_schur_kernel<<<b->nchunks, kChunk, 0, st>>>(v, G, lambda, d_partS); SDSO_CHECK_LAUNCH(ctx); }
      lba_host_sum_kernel<<<n, 96, 0, st>>>(v, d_partS, d_hostS); SDSO_CHECK_LAUNCH(ctx);
      lba_assemble_kernel<<<(d * d + d + 127) / 128, 128, 0, st>>>(n, d_hostA, d_hostS, d_used, lambda, S.HF, S.bF); SDSO_CHECK_LAUNCH(ctx);
      ba_solve_kernel<<<1, 256, smem, st>>>(S); SDSO_CHECK_LAUNCH(ctx);
      lba_trial_update_kernel<<<1, 32 * ((n + 31) / 32), 0, st>>>(n, d_used, S.x, lambda, d_hostA, d_est, d_estbak, d_sc); SDSO_CHECK_LAUNCH(ctx);
      lba_update_kernel<<<rb, 128, 0, st>>>(v, G, S.x, lambda, d_bpart, d_sc + LS_OK); SDSO_CHECK_LAUNCH(ctx);
      lba_sum_kernel<<<1, 32, 0, st>>>(d_bpart, rb, 1, 1, d_sc + LS_SL); SDSO_CHECK_LAUNCH(ctx);
      if ((rc = eval(d_active, 0, d_sc + LS_OK))) return cleanup(rc);
      if ((rc = chi2_to(LS_TEMP))) return cleanup(rc);
      if ((rc = read_scalars())) return cleanup(rc);
      if (!have_current) { currentChi = hsc[LS_CUR]; have_current = true; }
      const bool ok = hsc[LS_OK] != 0.0;
      tempChi = ok ? hsc[LS_TEMP] : std::numeric_limits<double>::max();
      const double scale = ok ? hsc[LS_SCALE] + hsc[LS_SL] : 0.0;
      rho = (currentChi - tempChi) / (scale + 1e-3);
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        lambda *= std::max(1. / 3., alpha);
        ni = 2; currentChi = tempChi;
        accepted = true;
      } else {
        lambda *= ni; ni *= 2;
        lba_pop_kernel<<<(std::max(R, 12 * n) + 255) / 256, 256, 0, st>>>(n, R, d_estbak, d_est, d_idbak, d_idepth); SDSO_CHECK_LAUNCH(ctx);   // pop()
        accepted = false;
        if (!std::isfinite(lambda)) break;
      }
      qmax++; trials++;
    } while (rho < 0 && qmax < 10);
    const bool terminate_lm = (qmax == 10 || rho == 0 || !std::isfinite(lambda));
    // SparseOptimizerTerminateAction: computeActiveErrors + chi2. Behind an accepted trial the edges already hold the errors of this
    // estimate and the chi2 is tempChi; behind a rejected one the estimate went back, so the pass runs.
    double chi = currentChi;
    if (!accepted) {
      if ((rc = eval(d_active, 0, nullptr))) return cleanup(rc);
      if ((rc = chi2_to(LS_POST))) return cleanup(rc);
      if ((rc = read_scalars())) return cleanup(rc);
      chi = hsc[LS_POST];
      currentChi = chi;     // what the next iteration's computeActiveErrors would find
    }
    if (it == 0) lastChi = chi;
    else { const double gain = (lastChi - chi) / chi; lastChi = chi; if (gain >= 0 && gain < 1e-3) stop = true; }
    if (terminate_lm) { it++; break; }
  }
  // the vertices back to the caller
  SDSO_CUDA(ctx, cudaMemcpyAsync(est.data(), d_est, est.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(T_wh, est.data(), sizeof(double) * n * 12);
  memcpy(photo, est.data() + n * 24, sizeof(double) * n * 2);
  for (int k = 0; k < 4; k++) cam[k] = est[(size_t)n * 29 + k];
  // ---- results in the caller's residual order
  std::vector<double> id_out(R);
  std::vector<int> ns_out(R);
  std::vector<float> ce_out((size_t)R * 3), ih_out(R);
  cudaMemcpyAsync(id_out.data(), d_idepth, R * sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(ns_out.data(), d_ns, R * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(ce_out.data(), d_center, (size_t)R * 3 * sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(ih_out.data(), d_ih, R * sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  for (int rid = 0; rid < R; rid++) {
    const int sl = b->h_rid2slot[rid];
    idepth[rid] = id_out[sl];
    if (newState) newState[rid] = ns_out[sl];
    if (center3) for (int k = 0; k < 3; k++) center3[3 * rid + k] = ce_out[(size_t)sl * 3 + k];
    if (idepth_hessian) idepth_hessian[rid] = ih_out[sl];
  }
  if (chi2_out) *chi2_out = lastChi;
  if (iterations_out) *iterations_out = it;
  if (trials_out) *trials_out = trials;
  return cleanup(SDSO_OK);
}

// EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:663-736) for the points flagged PS_MARGINALIZE
int sdso_ba_marginalize_points(sdso_ctx* ctx) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  invalidate_graphs(b);
  const int d = b->dim(), P = b->P, R = b->R;
  cudaStream_t st = ctx->stream;
  if (P == 0) return SDSO_OK;
  ba_marg_prior_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, b->d_p_flag, b->d_p_priorF, ctx->S.idepthFixPriorMargFac); SDSO_CHECK_LAUNCH(ctx);
  int rc = launch_top(ctx, 2, SYS_A, false);   // accumulateTop<2> (:680-696)
  if (!rc) rc = launch_sc(ctx, false, SYS_SC); // accumulateSC without the prior shift
  if (rc) return rc;
  const int cnt = d * d + d;
  ba_marg_add_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(cnt, (double)ctx->S.margWeightFac, sysH(b, SYS_A), sysH(b, SYS_SC), sysH(b, SYS_M), b->have_M ? 0 : 1);
  SDSO_CHECK_LAUNCH(ctx);
  b->have_M = true;
  BAView v = view(b);
  if (R > 0) { ba_marg_remove_kernel<<<(R + 127) / 128, 128, 0, st>>>(v, b->d_p_flag); SDSO_CHECK_LAUNCH(ctx); }
  ba_marg_flag_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, b->d_p_flag); SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}

// EnergyFunctional::marginalizeFrame (EnergyFunctional.cpp:554-660). The frame must own no points any more (as the reference
// asserts). HM / bM shrink to dimension d-8 on the device; the host window loses the frame, so points / residuals / states have
// to be uploaded again before the next operator call (the reference re-indexes with makeIDX at this point).
int sdso_ba_marginalize_frame(sdso_ctx* ctx, int idx) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  invalidate_graphs(b);
  if (idx < 0 || idx >= b->n) return SDSO_E_INVALID;
  const int odim = b->dim(), ndim = odim - 8;
  cudaStream_t st = ctx->stream;
  if (!b->have_M) { SDSO_CUDA(ctx, cudaMemsetAsync(sysH(b, SYS_M), 0, ((size_t)odim * odim + odim) * sizeof(double), st)); b->have_M = true; }
  const size_t smem = ((size_t)odim * odim + 2 * (size_t)odim + 64 + (size_t)ndim * 8) * sizeof(double);
  static bool attr_set_dev[64] = {false};   // function attributes are per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) { cudaFuncSetAttribute(ba_marg_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  double* out = sysH(b, SYS_TMP);
  ba_marg_frame_kernel<<<1, 256, smem, st>>>(odim, idx, sysH(b, SYS_M), sysb(b, SYS_M), b->d_fprior, out, out + (size_t)ndim * ndim);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(sysH(b, SYS_M), out, ((size_t)ndim * ndim + ndim) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  b->frames.erase(b->frames.begin() + idx);
  b->n = (int)b->frames.size();
  b->P = 0; b->R = 0; b->nchunks = 0;
  b->prepared = false;
  return SDSO_OK;
}

// EnergyFunctional::calcMEnergyF (:344-351) and calcLEnergyF_MT (:354-442)
int sdso_ba_energies(sdso_ctx* ctx, double* menergy, double* lenergy) {
  sdso::enter(ctx);
  BA_PREPARED(ctx)
  const int d = b->dim(), P = b->P;
  cudaStream_t st = ctx->stream;
  double M = 0, L = 0;
  if (menergy && b->have_M) {
    ba_menergy_kernel<<<1, 256, 0, st>>>(d, sysH(b, SYS_M), sysb(b, SYS_M), b->d_fprior, b->d_cDeltaF, b->d_scalars + 2); SDSO_CHECK_LAUNCH(ctx);
    SDSO_CUDA(ctx, cudaMemcpyAsync(&M, b->d_scalars + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  std::vector<double> part;
  if (lenergy && P > 0) {
    const int pb = (P + 127) / 128;
    if (pb > b->step_part_cap) {
      if (b->d_step_part) cudaFree(b->d_step_part);
      b->d_step_part = nullptr;
      SDSO_CUDA(ctx, cudaMalloc(&b->d_step_part, (size_t)(pb + 1) * 2 * sizeof(double)));
      b->step_part_cap = pb;
    }
    BAView v = view(b);
    ba_lenergy_kernel<<<pb, 128, 0, st>>>(v, b->d_step_part); SDSO_CHECK_LAUNCH(ctx);
    part.resize(pb);
    SDSO_CUDA(ctx, cudaMemcpyAsync(part.data(), b->d_step_part, pb * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  if (lenergy) {
    for (double v : part) L += v;
    for (auto& f : b->frames) for (int i = 0; i < 8; i++) L += f.delta_prior[i] * f.prior[i] * f.delta_prior[i];   // :429-431
    for (int i = 0; i < 4; i++) { const float cd = (float)b->calib_delta[i], cp = (float)b->cPrior[i]; L += cd * cp * cd; }
    *lenergy = L;
  }
  if (menergy) *menergy = M;
  return SDSO_OK;
}

int sdso_ba_get_marg_prior(sdso_ctx* ctx, double* HM, double* bM) {
  sdso::enter(ctx);
  BA_CHECK(ctx)
  return download_sys(ctx, SYS_M, HM, bM);
}

}  // extern "C"
