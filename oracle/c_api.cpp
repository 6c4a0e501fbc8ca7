// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline / --impl reference).
#include "oracle_core.hpp"
#include "oracle_ba.hpp"
#include "oracle_trace.hpp"
#include "oracle_select.hpp"
#include "oracle_distmap.hpp"
#include "oracle_undistort.hpp"
#include <memory>

using namespace orc;

namespace {
struct Ctx {
  GlobalCalib G;
  Settings S;
  CalibHessian HCalib;
  std::vector<std::unique_ptr<Frame>> frames;
  CoarseTracker tracker;
  BAWindow ba;
};
}  // namespace

extern "C" {

void* orc_create(int w, int h, float fx, float fy, float cx, float cy, float baseline) {
  Ctx* c = new Ctx();
  float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  c->G.set(w, h, K);
  c->G.baseline = baseline;
  c->HCalib.setValueScaledf(c->G.fx[0], c->G.fy[0], c->G.cx[0], c->G.cy[0]);
  c->tracker.init(&c->G, c->S);
  c->tracker.makeK(c->HCalib);
  return c;
}
void orc_destroy(void* p) { delete (Ctx*)p; }
int orc_levels(void* p) { return ((Ctx*)p)->G.levels; }
void orc_level_size(void* p, int lvl, int* w, int* h) { *w = ((Ctx*)p)->G.w[lvl]; *h = ((Ctx*)p)->G.h[lvl]; }
void orc_level_K(void* p, int lvl, float K[9], float Ki[9]) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < 9; i++) { K[i] = c->G.K[lvl][i]; Ki[i] = c->G.Ki[lvl][i]; }
}
void orc_set_affine_opt_mode(void* p, float a, float b) {
  Ctx* c = (Ctx*)p; c->S.affineOptModeA = a; c->S.affineOptModeB = b; c->tracker.S = c->S;
}

int orc_frame_new(void* p) { Ctx* c = (Ctx*)p; c->frames.emplace_back(new Frame()); return (int)c->frames.size() - 1; }
void orc_make_images(void* p, int fid, const float* color, float exposure, int use_hcalib) {
  Ctx* c = (Ctx*)p;
  c->frames[fid]->ab_exposure = exposure;
  c->frames[fid]->makeImages(c->G, color, use_hcalib ? &c->HCalib : nullptr, c->S);
}
void orc_frame_get(void* p, int fid, int lvl, float* dI3, float* absgrad) {
  Ctx* c = (Ctx*)p;
  const Frame& f = *c->frames[fid];
  if (dI3) memcpy(dI3, f.dIp[lvl].data(), f.dIp[lvl].size() * sizeof(float));
  if (absgrad) memcpy(absgrad, f.absSquaredGrad[lvl].data(), f.absSquaredGrad[lvl].size() * sizeof(float));
}
void orc_interp33(void* p, int fid, int lvl, const float* xy, int n, float* out3) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) getInterpolatedElement33(c->frames[fid]->dIp[lvl].data(), xy[2 * i], xy[2 * i + 1], c->G.w[lvl], out3 + 3 * i);
}
void orc_interp33bilin(void* p, int fid, int lvl, const float* xy, int n, float* out3) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) getInterpolatedElement33BiLin(c->frames[fid]->dIp[lvl].data(), xy[2 * i], xy[2 * i + 1], c->G.w[lvl], out3 + 3 * i);
}

// ---- tracker ---------------------------------------------------------------------------------
void orc_tracker_makeK(void* p, float fx, float fy, float cx, float cy) {
  Ctx* c = (Ctx*)p; c->HCalib.setValueScaledf(fx, fy, cx, cy); c->tracker.makeK(c->HCalib);
}
void orc_tracker_set_ref(void* p, int fid, const float* uvidw, int n, const double aff[2]) {
  Ctx* c = (Ctx*)p;
  c->tracker.setRefFromSplats(c->frames[fid].get(), (const RefPoint*)uvidw, n, aff);
}
int orc_tracker_pc(void* p, int lvl, float* u, float* v, float* idepth, float* color) {
  Ctx* c = (Ctx*)p; int n = c->tracker.pc_n[lvl];
  if (u) memcpy(u, c->tracker.pc_u[lvl].data(), n * sizeof(float));
  if (v) memcpy(v, c->tracker.pc_v[lvl].data(), n * sizeof(float));
  if (idepth) memcpy(idepth, c->tracker.pc_idepth[lvl].data(), n * sizeof(float));
  if (color) memcpy(color, c->tracker.pc_color[lvl].data(), n * sizeof(float));
  return n;
}
void orc_tracker_set_pc(void* p, int fid, int lvl, int n, const float* u, const float* v, const float* idepth, const float* color, const double aff[2]) {
  Ctx* c = (Ctx*)p; CoarseTracker& t = c->tracker;
  t.lastRef = c->frames[fid].get(); t.lastRef_aff_g2l[0] = aff[0]; t.lastRef_aff_g2l[1] = aff[1];
  memcpy(t.pc_u[lvl].data(), u, n * sizeof(float)); memcpy(t.pc_v[lvl].data(), v, n * sizeof(float));
  memcpy(t.pc_idepth[lvl].data(), idepth, n * sizeof(float)); memcpy(t.pc_color[lvl].data(), color, n * sizeof(float));
  t.pc_n[lvl] = n;
}
void orc_calc_res_sse(void* p, int new_fid, int lvl, const double T[12], const double aff[2], float cutoff, double rs[6], int* warped_n) {
  Ctx* c = (Ctx*)p; c->tracker.newFrame = c->frames[new_fid].get();
  c->tracker.calcResSSE(lvl, SE3::fromMat34(T), aff, cutoff, rs);
  *warped_n = c->tracker.buf_warped_n;
}
// order: idepth,u,v,dx,dy,residual,weight,refColor, each warped_n floats
void orc_get_warped(void* p, float* out) {
  Ctx* c = (Ctx*)p; CoarseTracker& t = c->tracker; int n = t.buf_warped_n;
  const std::vector<float>* bufs[8] = {&t.buf_warped_idepth, &t.buf_warped_u, &t.buf_warped_v, &t.buf_warped_dx, &t.buf_warped_dy,
                                       &t.buf_warped_residual, &t.buf_warped_weight, &t.buf_warped_refColor};
  for (int k = 0; k < 8; k++) memcpy(out + (size_t)k * n, bufs[k]->data(), n * sizeof(float));
}
void orc_calc_gs_sse(void* p, int lvl, const double T[12], const double aff[2], double H[64], double b[8]) {
  Ctx* c = (Ctx*)p; c->tracker.calcGSSSE(lvl, H, b, SE3::fromMat34(T), aff);
}
int orc_track_sse(void* p, int new_fid, double T[12], double aff[2], int coarsest, const double minRes[5], double lastRes[5], double flow[3], int iters[5]) {
  Ctx* c = (Ctx*)p; SE3 s = SE3::fromMat34(T);
  bool ok = c->tracker.trackNewestCoarseSSE(c->frames[new_fid].get(), s, aff, coarsest, minRes, iters);
  s.toMat34(T);
  for (int i = 0; i < 5; i++) lastRes[i] = c->tracker.lastResiduals[i];
  for (int i = 0; i < 3; i++) flow[i] = c->tracker.lastFlowIndicators[i];
  return ok ? 1 : 0;
}
int orc_track_g2o(void* p, int new_fid, double T[12], double aff[2], int coarsest, const double minRes[5], double lastRes[5], double flow[3], int iters[5]) {
  Ctx* c = (Ctx*)p; SE3 s = SE3::fromMat34(T);
  bool ok = c->tracker.trackNewestCoarseG2O(c->frames[new_fid].get(), s, aff, coarsest, minRes, iters);
  s.toMat34(T);
  for (int i = 0; i < 5; i++) lastRes[i] = c->tracker.lastResiduals[i];
  for (int i = 0; i < 3; i++) flow[i] = c->tracker.lastFlowIndicators[i];
  return ok ? 1 : 0;
}
// E1 operator-level: per pc point of a level, error + Jacobians at (pose, photo); returns count of in-border points
int orc_edge_eval(void* p, int new_fid, int lvl, const double Tsel[12], const double Tpose[12], const double photo[2],
                  double* err /*n*/, double* J8 /*n*8*/, int* idx /*n*/) {
  Ctx* c = (Ctx*)p; CoarseTracker& t = c->tracker;
  t.newFrame = c->frames[new_fid].get(); t.edges.clear();
  double rs[6]; SE3 pose = SE3::fromMat34(Tpose);
  t.calcResG2O(lvl, SE3::fromMat34(Tsel), 1e30f, pose, photo, rs);
  int n = (int)t.edges.size();
  for (int k = 0; k < n; k++) {
    err[k] = t.edges[k].error;
    t.edgeLinearizeOplus(t.edges[k], pose, photo, J8 + 8 * k, J8 + 8 * k + 6);
    (void)idx;
  }
  return n;
}
unsigned long long orc_evals(void* p) { return ((Ctx*)p)->tracker.evals; }
void orc_reset_evals(void* p) { ((Ctx*)p)->tracker.evals = 0; }
void orc_g2o_trial_counts(void* p, int out[2]) { out[0] = ((Ctx*)p)->tracker.g2o_trials; out[1] = ((Ctx*)p)->tracker.g2o_rejected; }

// ---- SE3 -------------------------------------------------------------------------------------
void orc_se3_exp(const double a[6], double T[12]) { SE3::exp(a).toMat34(T); }
void orc_se3_log(const double T[12], double a[6]) { SE3::fromMat34(T).log(a); }
void orc_se3_adj(const double T[12], double A[36]) { SE3::fromMat34(T).Adj(A); }
void orc_se3_mul(const double A[12], const double B[12], double C[12]) { (SE3::fromMat34(A) * SE3::fromMat34(B)).toMat34(C); }
// V1-V5 oplusImpl (dso_g2o_vertex.cpp:15-18, 30-40, 56-58, 73-88, 100-106); kinds numbered as in include/sdso_b200.h
void orc_vertex_oplus(int kind, int n, double* est, const double* upd, const double* aux) {
  for (int i = 0; i < n; i++) {
    switch (kind) {
      case 1: { SE3 T = SE3::exp(upd + 6 * i) * SE3::fromMat34(est + 12 * i); T.toMat34(est + 12 * i); break; }
      case 2: est[2 * i] += upd[2 * i]; est[2 * i + 1] += upd[2 * i + 1]; break;
      case 3: est[i] += upd[i]; break;
      case 4: {
        double update = upd[i];
        if (update < -0.5) update = -0.5;
        else if (update > 0.5) update = 0.5;
        else if (!std::isfinite(update)) update = 0;
        est[2 * i] += update * aux[2 * i]; est[2 * i + 1] += update * aux[2 * i + 1];
        break;
      }
      case 5: for (int k = 0; k < 4; k++) est[4 * i + k] += upd[4 * i + k]; break;
    }
  }
}
// E3: EdgeTracePointUVDSO::computeError + linearizeOplus (dso_g2o_edge.cpp:571-619); error / J keep their previous contents where
// the edge leaves its members untouched
void orc_edge_trace_uv(void* p, int fid, int n, const double* uv, const float* rot, const double* meas, const float aff[2], const double* dxdy,
                       double* err, double* J, int* flag) {
  Ctx* c = (Ctx*)p;
  const Frame& f = *c->frames[fid];
  const int wl = c->G.w[0] - 3, hl = c->G.h[0] - 3;
  for (int i = 0; i < n; i++) {
    const double U = uv[2 * i], V = uv[2 * i + 1];
    if ((U - 2) < 0 || (U + 3) > wl || (V - 2) < 0 || (V + 3) > hl) { err[i] = 0.0; if (flag) flag[i] = 0; continue; }
    float hit[3];
    getInterpolatedElement33(f.dIp[0].data(), (float)(U + rot[2 * i]), (float)(V + rot[2 * i + 1]), c->G.w[0], hit);
    if (!std::isfinite(hit[0])) { if (flag) flag[i] = 2; continue; }
    err[i] = hit[0] - (aff[0] * meas[i] + aff[1]);
    J[i] = dxdy[2 * i] * hit[1] + dxdy[2 * i + 1] * hit[2];
    if (flag) flag[i] = 1;
  }
}
void orc_se3_inv(const double A[12], double B[12]) { SE3::fromMat34(A).inverse().toMat34(B); }
void orc_ldlt_solve(int n, const double* A, const double* b, double* x) { ldlt_solve(n, A, b, x); }

}  // extern "C"

// ---- D1: ImmaturePoint constructor (ImmaturePoint.cpp:33-88): colour, weights, gradH of the 8-pixel pattern ----
extern "C" int orc_immature_init(void* p, int fid, float u, float v, float* color8, float* weights8, float* gradH4, float* energyTH) {
  Ctx* c = (Ctx*)p;
  const float* dI = c->frames[fid]->dIp[0].data();
  float gradH[4] = {0, 0, 0, 0};
  for (int idx = 0; idx < patternNum; idx++) {
    int dx = patternP[idx][0], dy = patternP[idx][1];
    float ptc[3];
    getInterpolatedElement33BiLin(dI, u + dx, v + dy, c->G.w[0], ptc);
    color8[idx] = ptc[0];
    if (!std::isfinite(color8[idx])) { *energyTH = NAN; return 0; }
    gradH[0] += ptc[1] * ptc[1]; gradH[1] += ptc[1] * ptc[2]; gradH[2] += ptc[2] * ptc[1]; gradH[3] += ptc[2] * ptc[2];
    weights8[idx] = sqrtf(c->S.outlierTHSumComponent / (c->S.outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
  }
  for (int i = 0; i < 4; i++) gradH4[i] = gradH[i];
  float e = patternNum * c->S.outlierTH;
  e *= c->S.overallEnergyTHWeight * c->S.overallEnergyTHWeight;
  *energyTH = e;
  return 1;
}

// ---- windowed BA (SSE path) -------------------------------------------------------------------------
extern "C" {
void orc_ba_reset(void* p) {
  Ctx* c = (Ctx*)p;
  c->ba = BAWindow();
  c->ba.G = &c->G; c->ba.S = c->S; c->ba.HCalib = c->HCalib;
}
void orc_ba_set_calib_delta(void* p, const double d[4]) { Ctx* c = (Ctx*)p; for (int i = 0; i < 4; i++) c->ba.HCalib.value_minus_value_zero[i] = d[i]; }
int orc_ba_add_frame(void* p, int fid, const double T_w2c[12], double a, double b, int frameID) {
  Ctx* c = (Ctx*)p;
  BAFrame f; f.img = c->frames[fid].get(); f.frameID = frameID;
  f.setEvalPT_scaled(SE3::fromMat34(T_w2c), a, b);
  // EFFrame::takeData -> FrameHessian::getPrior (HessianBlocks.h:246-268)
  for (int i = 0; i < 8; i++) f.prior[i] = 0;
  if (frameID == 0) {
    for (int i = 0; i < 3; i++) f.prior[i] = c->S.initialTransPrior;
    for (int i = 3; i < 6; i++) f.prior[i] = c->S.initialRotPrior;
    f.prior[6] = c->S.initialAffAPrior; f.prior[7] = c->S.initialAffBPrior;
  } else {
    f.prior[6] = c->S.affineOptModeA < 0 ? c->S.initialAffAPrior : c->S.affineOptModeA;
    f.prior[7] = c->S.affineOptModeB < 0 ? c->S.initialAffBPrior : c->S.affineOptModeB;
  }
  c->ba.frames.push_back(f);
  return (int)c->ba.frames.size() - 1;
}
void orc_ba_set_state(void* p, int idx, const double state[10]) { ((Ctx*)p)->ba.frames[idx].setState(state); }
void orc_ba_set_energy_th(void* p, int idx, float th) { ((Ctx*)p)->ba.frames[idx].frameEnergyTH = th; }
int orc_ba_add_point(void* p, int host, float u, float v, float idepth, float idepth_zero, const float* color8, const float* weights8, int hasDepthPrior) {
  Ctx* c = (Ctx*)p;
  BAPoint q; q.host = host; q.u = u; q.v = v;
  q.idepth = idepth; q.idepth_scaled = SCALE_IDEPTH * idepth;
  q.idepth_zero = idepth_zero; q.idepth_zero_scaled = SCALE_IDEPTH * idepth_zero;
  for (int i = 0; i < 8; i++) { q.color[i] = color8[i]; q.weights[i] = weights8[i]; }
  q.hasDepthPrior = hasDepthPrior != 0;
  q.priorF = q.hasDepthPrior ? c->S.idepthFixPrior * SCALE_IDEPTH * SCALE_IDEPTH : 0;  // EFPoint::takeData
  q.deltaF = q.idepth - q.idepth_zero;
  c->ba.points.push_back(q);
  return (int)c->ba.points.size() - 1;
}
int orc_ba_add_residual(void* p, int pidx, int target);
int orc_ba_add_point(void* p, int host, float u, float v, float idepth, float idepth_zero, const float* color8, const float* weights8, int hasDepthPrior);
// batched forms (the per-point calls cost more in the Python harness than the operators they feed)
void orc_ba_add_points(void* p, int n, const int* host, const float* u, const float* v, const float* idepth, const float* idepth_zero, const float* color8,
                       const float* weights8, const unsigned char* prior) {
  for (int i = 0; i < n; i++) orc_ba_add_point(p, host[i], u[i], v[i], idepth[i], idepth_zero[i], color8 + 8 * i, weights8 + 8 * i, prior[i]);
}
void orc_ba_add_residuals(void* p, int n, const int* point, const int* target) {
  for (int i = 0; i < n; i++) orc_ba_add_residual(p, point[i], target[i]);
}
int orc_ba_add_residual(void* p, int pidx, int target) {
  Ctx* c = (Ctx*)p;
  BARes r; r.point = pidx; r.host = c->ba.points[pidx].host; r.target = target;
  r.state_state = RS_IN; r.state_NewState = RS_OUTLIER; r.state_energy = 0;  // resetOOB (Residuals.h:107-115)
  memset(&r.J, 0, sizeof(r.J)); memset(&r.efJ, 0, sizeof(r.efJ));
  c->ba.res.push_back(r);
  c->ba.points[pidx].residuals.push_back((int)c->ba.res.size() - 1);
  return (int)c->ba.res.size() - 1;
}
void orc_ba_set_point_flag(void* p, int pidx, int flag) { ((Ctx*)p)->ba.points[pidx].stateFlag = flag; }
void orc_ba_prepare(void* p) {
  Ctx* c = (Ctx*)p;
  c->ba.setPrecalcValues(); c->ba.setAdjointsF(); c->ba.setDeltaF(); c->ba.getNullspaces();
}
int orc_ba_counts(void* p, int* nframes, int* npoints, int* nres) {
  Ctx* c = (Ctx*)p; *nframes = c->ba.n(); *npoints = (int)c->ba.points.size(); *nres = (int)c->ba.res.size(); return c->ba.dim();
}
void orc_ba_precalc(void* p, int h, int t, float* out /*50*/) {
  Ctx* c = (Ctx*)p; const FrameFramePrecalc& q = c->ba.precalc[(size_t)h * c->ba.n() + t];
  int k = 0;
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_KRKiTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RKiTll[i];
  for (int i = 0; i < 9; i++) out[k++] = q.PRE_RTll_0[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_tTll[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_KtTll[i];
  for (int i = 0; i < 3; i++) out[k++] = q.PRE_tTll_0[i];
  out[k++] = q.PRE_aff_mode[0]; out[k++] = q.PRE_aff_mode[1]; out[k++] = q.PRE_b0_mode; out[k++] = q.distanceLL;
}
void orc_ba_adjoints(void* p, double* adHost, double* adTarget, float* adHTdeltaF) {
  Ctx* c = (Ctx*)p;
  if (adHost) memcpy(adHost, c->ba.adHost.data(), c->ba.adHost.size() * sizeof(double));
  if (adTarget) memcpy(adTarget, c->ba.adTarget.data(), c->ba.adTarget.size() * sizeof(double));
  if (adHTdeltaF) memcpy(adHTdeltaF, c->ba.adHTdeltaF.data(), c->ba.adHTdeltaF.size() * sizeof(float));
}
double orc_ba_linearize_all(void* p, int fix) { return ((Ctx*)p)->ba.linearizeAll(fix != 0); }
void orc_ba_apply_res(void* p, int copy) { Ctx* c = (Ctx*)p; for (auto& r : c->ba.res) if (!r.isLinearized) c->ba.applyRes(r, copy != 0); }  // activeResiduals only
void orc_ba_fix_linearization(void* p, int ridx) { Ctx* c = (Ctx*)p; c->ba.fixLinearizationF(c->ba.res[ridx]); }
// per residual: state_NewState, state_state, NewEnergy, NewEnergyWithOutlier, isActive, J (74 floats: candidate J if which==0, EF J if which==1), JpJdF, centerProjectedTo
void orc_ba_get_res(void* p, int which, int* newState, int* state, double* newEnergy, double* newEnergyWO, int* active, float* J74, float* JpJdF8, float* center3, float* resToZero8) {
  Ctx* c = (Ctx*)p;
  for (size_t i = 0; i < c->ba.res.size(); i++) {
    const BARes& r = c->ba.res[i];
    if (newState) newState[i] = r.state_NewState;
    if (state) state[i] = r.state_state;
    if (newEnergy) newEnergy[i] = r.state_NewEnergy;
    if (newEnergyWO) newEnergyWO[i] = r.state_NewEnergyWithOutlier;
    if (active) active[i] = r.isActive() ? 1 : 0;
    if (J74) memcpy(J74 + 74 * i, which == 0 ? &r.J : &r.efJ, 74 * sizeof(float));
    if (JpJdF8) memcpy(JpJdF8 + 8 * i, r.JpJdF, 8 * sizeof(float));
    if (center3) memcpy(center3 + 3 * i, r.centerProjectedTo, 3 * sizeof(float));
    if (resToZero8) memcpy(resToZero8 + 8 * i, r.res_toZeroF, 8 * sizeof(float));
  }
}
// per point: Hdd_accAF, bd_accAF, Hcd_accAF[4], Hdd_accLF, bd_accLF, Hcd_accLF[4], HdiF, bdSumF, step, priorF, deltaF (16 floats)
void orc_ba_get_points(void* p, float* out16) {
  Ctx* c = (Ctx*)p;
  for (size_t i = 0; i < c->ba.points.size(); i++) {
    const BAPoint& q = c->ba.points[i]; float* o = out16 + 16 * i;
    o[0] = q.Hdd_accAF; o[1] = q.bd_accAF; for (int k = 0; k < 4; k++) o[2 + k] = q.Hcd_accAF[k];
    o[6] = q.Hdd_accLF; o[7] = q.bd_accLF; for (int k = 0; k < 4; k++) o[8 + k] = q.Hcd_accLF[k];
    o[12] = q.HdiF; o[13] = q.bdSumF; o[14] = q.step; o[15] = q.priorF;
  }
}
void orc_ba_accumulate_top(void* p, int mode, int usePrior, double* H, double* b, float* blocks) {
  Ctx* c = (Ctx*)p; std::vector<double> Hv, bv;
  c->ba.accumulateTop(mode, Hv, bv, usePrior != 0);
  memcpy(H, Hv.data(), Hv.size() * sizeof(double)); memcpy(b, bv.data(), bv.size() * sizeof(double));
  if (blocks) memcpy(blocks, c->ba.lastTopBlocks.data(), c->ba.lastTopBlocks.size() * sizeof(float));
}
void orc_ba_accumulate_sc(void* p, int shift, double* H, double* b) {
  Ctx* c = (Ctx*)p; std::vector<double> Hv, bv;
  c->ba.accumulateSC(shift != 0, Hv, bv);
  memcpy(H, Hv.data(), Hv.size() * sizeof(double)); memcpy(b, bv.data(), bv.size() * sizeof(double));
}
void orc_ba_solve(void* p, int iteration, double lambda, double* x, double* Hfinal, double* bfinal) {
  Ctx* c = (Ctx*)p; std::vector<double> xv, Hf, bf;
  c->ba.solveSystemF(iteration, lambda, xv, &Hf, &bf);
  memcpy(x, xv.data(), xv.size() * sizeof(double));
  if (Hfinal) memcpy(Hfinal, Hf.data(), Hf.size() * sizeof(double));
  if (bfinal) memcpy(bfinal, bf.data(), bf.size() * sizeof(double));
}
void orc_ba_resubstitute(void* p, const double* x, double* frame_steps, double* calib_step) {
  Ctx* c = (Ctx*)p; std::vector<double> xv(x, x + c->ba.dim());
  c->ba.resubstituteF(xv, frame_steps, calib_step);
}
void orc_ba_set_marg_prior(void* p, const double* HM, const double* bM) {
  Ctx* c = (Ctx*)p; int d = c->ba.dim();
  c->ba.HM.assign(HM, HM + (size_t)d * d); c->ba.bM.assign(bM, bM + d);
}
void orc_ba_get_marg_prior(void* p, double* HM, double* bM) {
  Ctx* c = (Ctx*)p;
  memcpy(HM, c->ba.HM.data(), c->ba.HM.size() * sizeof(double)); memcpy(bM, c->ba.bM.data(), c->ba.bM.size() * sizeof(double));
}
void orc_ba_marginalize_points(void* p) { ((Ctx*)p)->ba.marginalizePointsF(); }
void orc_ba_marginalize_frame(void* p, int idx) { ((Ctx*)p)->ba.marginalizeFrame(idx); }
void orc_ba_orthogonalize(void* p, double* b, double* H) {
  Ctx* c = (Ctx*)p; int d = c->ba.dim();
  std::vector<double> bv, Hv;
  if (b) bv.assign(b, b + d);
  if (H) Hv.assign(H, H + (size_t)d * d);
  c->ba.orthogonalize(b ? &bv : nullptr, H ? &Hv : nullptr);
  if (b) memcpy(b, bv.data(), d * sizeof(double));
  if (H) memcpy(H, Hv.data(), (size_t)d * d * sizeof(double));
}
double orc_ba_energies(void* p, double* lenergy) { Ctx* c = (Ctx*)p; if (lenergy) *lenergy = c->ba.calcLEnergyF(); return c->ba.calcMEnergyF(); }
void orc_ba_nullspaces(void* p, double* N /* dim x 7, row-major */) {
  Ctx* c = (Ctx*)p; int d = c->ba.dim();
  for (int i = 0; i < 6; i++) for (int r = 0; r < d; r++) N[(size_t)r * 7 + i] = c->ba.lastNullspaces_pose[i][r];
  for (int r = 0; r < d; r++) N[(size_t)r * 7 + 6] = c->ba.lastNullspaces_scale[0][r];
}
}  // extern "C"

// ---- D1-D3 on arrays of ImmaturePoint records (oracle/oracle_trace.hpp) ---------------------------------------
extern "C" {
int orc_immature_record_size() { return (int)sizeof(ImmaturePoint); }
void orc_immature_init_batch(void* p, int fid, int n, const float* uv, ImmaturePoint* out, int* ok) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) ok[i] = immatureInit(c->G, c->S, *c->frames[fid], uv[2 * i], uv[2 * i + 1], out[i]) ? 1 : 0;
}
void orc_trace_on(void* p, int fid, const float KRKi[9], const float Kt[3], const float aff[2], int n, ImmaturePoint* pts, int* status) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) status[i] = traceOn(c->G, c->S, pts[i], *c->frames[fid], KRKi, Kt, aff);
}
void orc_trace_stereo(void* p, int fid, const float K[9], int mode_right, int n, ImmaturePoint* pts, int* status) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) status[i] = traceStereo(c->G, c->S, pts[i], *c->frames[fid], K, mode_right != 0);
}
}  // extern "C"

extern "C" {
// worker partition of the accumulators (oracle_ba.hpp: reduce_threads / reduce_seed); 1, 0 = the single-threaded path
void orc_ba_set_reduce(void* p, int threads, unsigned seed) { Ctx* c = (Ctx*)p; c->ba.reduce_threads = threads; c->ba.reduce_seed = seed; }
double orc_ba_optimize(void* p, int iters, int* done) { return ((Ctx*)p)->ba.optimize(iters, done); }
void orc_ba_get_energy_th(void* p, float* th) { Ctx* c = (Ctx*)p; for (size_t i = 0; i < c->ba.frames.size(); i++) th[i] = c->ba.frames[i].frameEnergyTH; }
float orc_ba_new_frame_energy_th(void* p) { return ((Ctx*)p)->ba.newFrameEnergyTH(); }
// frame states [n][10], world-to-camera [n][12], point idepths [P], calibration value_scaled [4]
void orc_ba_get_state(void* p, double* states, double* T_w2c, float* idepth, double* calib) {
  Ctx* c = (Ctx*)p;
  for (size_t h = 0; h < c->ba.frames.size(); h++) {
    if (states) for (int i = 0; i < 10; i++) states[h * 10 + i] = c->ba.frames[h].state[i];
    if (T_w2c) c->ba.frames[h].PRE_worldToCam.toMat34(T_w2c + 12 * h);
  }
  if (idepth) for (size_t i = 0; i < c->ba.points.size(); i++) idepth[i] = c->ba.points[i].idepth;
  if (calib) { calib[0] = c->ba.HCalib.fxl; calib[1] = c->ba.HCalib.fyl; calib[2] = c->ba.HCalib.cxl; calib[3] = c->ba.HCalib.cyl; }
}
}  // extern "C"

// ---- E2 operator level: every residual of the window as one edge (oracle/lba_edge.cpp) ----------------------------------
extern "C" void orc_lba_edge_eval(void* p, const double* T_wh /*[n][12]*/, const double* photo /*[n][2]*/, const double* idepth /*[R]*/,
                                  const double cam[4], const double* b0 /*[n]*/, double* error8, double* Jxi, double* Jphoto, double* Jid,
                                  double* JC, int* newState, double* newEnergy, double* newEnergyWO, float* center3, float* idepth_hessian,
                                  int* level) {
  Ctx* c = (Ctx*)p;
  for (size_t i = 0; i < c->ba.res.size(); i++) {
    const BARes& r = c->ba.res[i];
    LBAEdgeOut o;
    lbaEdgeEval(c->ba, r, SE3::fromMat34(T_wh + 12 * r.host), photo + 2 * r.host, idepth[i], cam, b0[r.host], o);
    memcpy(error8 + 8 * i, o.error, sizeof(o.error)); memcpy(Jxi + 48 * i, o.J_xi, sizeof(o.J_xi)); memcpy(Jphoto + 16 * i, o.J_photo, sizeof(o.J_photo));
    memcpy(Jid + 8 * i, o.J_idepth, sizeof(o.J_idepth)); memcpy(JC + 32 * i, o.J_C, sizeof(o.J_C));
    newState[i] = o.newState; newEnergy[i] = o.newEnergy; newEnergyWO[i] = o.newEnergyWithOutlier;
    memcpy(center3 + 3 * i, o.centerProjectedTo, sizeof(o.centerProjectedTo)); idepth_hessian[i] = o.idepth_hessian; level[i] = o.level;
  }
}

// ---- D4: activation of n candidates hosted in the window's frames (oracle/activate.cpp) ------------------------------------
extern "C" void orc_activate_points(void* p, int n, const int* host, const ImmaturePoint* pts, int variant, int minObs, int* result, float* idepth,
                                    int* states /*[n][nframes]*/, float* energy) {
  Ctx* c = (Ctx*)p;
  const int nf = c->ba.n();
  for (int i = 0; i < n; i++) result[i] = activatePoint(c->ba, host[i], pts[i], variant, minObs, idepth + i, states + (size_t)i * nf, energy + i);
}

extern "C" int orc_lba_g2o(void* p, int iters, double cam[4], double* T_wh, double* photo, double* idepth, int* used_host, double* chi2, int* newState,
                           float* center3, float* idepth_hessian, int* trials) {
  Ctx* c = (Ctx*)p;
  return lbaG2O(c->ba, iters, cam, T_wh, photo, idepth, used_host, chi2, newState, center3, idepth_hessian, trials);
}

// ---- pixel selector (PixelSelector2.cpp) ----
extern "C" {
void* orc_sel_create(int w, int h) { auto* s = new orc::PixelSelector(); s->init(w, h); return s; }
void orc_sel_destroy(void* s) { delete (orc::PixelSelector*)s; }
void orc_sel_random_pattern(void* s, unsigned char* out) { auto* p = (orc::PixelSelector*)s; memcpy(out, p->randomPattern.data(), p->randomPattern.size()); }
int orc_sel_potential(void* s, int set) { auto* p = (orc::PixelSelector*)s; if (set > 0) p->currentPotential = set; return p->currentPotential; }
void orc_sel_forget_hist(void* s) { ((orc::PixelSelector*)s)->gradHistFrame = nullptr; }
void orc_sel_ths(void* s, float* ths, float* smoothed) {
  auto* p = (orc::PixelSelector*)s; int k = (p->w / 32) * (p->h / 32);
  memcpy(ths, p->ths.data(), k * sizeof(float)); memcpy(smoothed, p->thsSmoothed.data(), k * sizeof(float));
}
void orc_sel_make_hists(void* s, void* ctx, int fid) { ((orc::PixelSelector*)s)->makeHists(*((Ctx*)ctx)->frames[fid]); }
void orc_sel_select(void* s, void* ctx, int fid, float* map_out, int pot, float thFactor, int n[3]) {
  Ctx* c = (Ctx*)ctx;
  if (c->G.levels < 3) { n[0] = n[1] = n[2] = -1; return; }  // select reads absSquaredGrad[0..2]
  ((orc::PixelSelector*)s)->select(*c->frames[fid], c->G, map_out, pot, thFactor, n);
}
int orc_sel_make_maps(void* s, void* ctx, int fid, float* map_out, float density, int recursionsLeft, float thFactor) {
  Ctx* c = (Ctx*)ctx;
  if (c->G.levels < 3) return -1;
  return ((orc::PixelSelector*)s)->makeMaps(*c->frames[fid], c->G, map_out, density, recursionsLeft, thFactor);
}
}

// ---- coarse distance map + activation candidate filter ----
extern "C" {
void* orc_dm_create(void* ctx) { Ctx* c = (Ctx*)ctx; auto* m = new orc::CoarseDistanceMap(); m->init(c->G.w[1], c->G.h[1]); return m; }
void orc_dm_destroy(void* m) { delete (orc::CoarseDistanceMap*)m; }
// makeDistanceMap: hosts in frameHessians order (the newest frame is simply not listed), points grouped by host
void orc_dm_make(void* mp, int n_hosts, const float* KRKi, const float* Kt, const int* host_count, const float* uvid) {
  auto* m = (orc::CoarseDistanceMap*)mp;
  std::fill(m->fwdWarpedIDDistFinal.begin(), m->fwdWarpedIDDistFinal.end(), 1000.f);
  int numItems = 0, off = 0;
  for (int h = 0; h < n_hosts; h++) { numItems = m->seed(KRKi + 9 * h, Kt + 3 * h, host_count[h], uvid + 3 * off, numItems); off += host_count[h]; }
  m->growDistBFS(numItems);
}
void orc_dm_add(void* mp, int n, const int* uv) { auto* m = (orc::CoarseDistanceMap*)mp; for (int i = 0; i < n; i++) m->addIntoDistFinal(uv[2 * i], uv[2 * i + 1]); }
void orc_dm_get(void* mp, float* out) { auto* m = (orc::CoarseDistanceMap*)mp; memcpy(out, m->fwdWarpedIDDistFinal.data(), m->fwdWarpedIDDistFinal.size() * sizeof(float)); }
void orc_dm_filter(void* mp, void* ctx, int n_hosts, const float* KRKi, const float* Kt, const unsigned char* flagged, int n, const int* cand_host,
                   const void* pts, const float* my_type, float currentMinActDist, int* verdict) {
  Ctx* c = (Ctx*)ctx;
  orc::activationFilter(*(orc::CoarseDistanceMap*)mp, n_hosts, KRKi, Kt, flagged, n, cand_host, (const orc::ImmaturePoint*)pts, my_type, currentMinActDist,
                        c->S.minTraceQuality, verdict);
}
}

// ---- undistortion + trajectory rows ----
extern "C" {
float orc_undistort(int wOrg, int hOrg, int w, int h, const float* remapX, const float* remapY, const float* G, const float* vignetteInv, int photoCalib,
                    int useExposure, const unsigned char* raw, float exposure, float factor, float* out) {
  orc::Undistorter u;
  u.wOrg = wOrg; u.hOrg = hOrg; u.w = w; u.h = h;
  u.remapX.assign(remapX, remapX + (size_t)w * h); u.remapY.assign(remapY, remapY + (size_t)w * h);
  if (G) u.G.assign(G, G + 256);
  if (vignetteInv) u.vignetteMapInv.assign(vignetteInv, vignetteInv + (size_t)wOrg * hOrg);
  u.photometricCalibration = photoCalib; u.useExposure = useExposure != 0;
  return u.undistort(raw, exposure, factor, out);
}
int orc_trajectory_row(const double T[12], char* buf, int n) {
  const std::string s = orc::trajectoryRow(T);
  if ((int)s.size() + 1 > n) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
}
