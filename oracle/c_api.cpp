// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline / --impl reference).
#include "oracle_core.hpp"
#include <memory>

using namespace orc;

namespace {
struct Ctx {
  GlobalCalib G;
  Settings S;
  CalibHessian HCalib;
  std::vector<std::unique_ptr<Frame>> frames;
  CoarseTracker tracker;
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

// ---- SE3 -------------------------------------------------------------------------------------
void orc_se3_exp(const double a[6], double T[12]) { SE3::exp(a).toMat34(T); }
void orc_se3_log(const double T[12], double a[6]) { SE3::fromMat34(T).log(a); }
void orc_se3_adj(const double T[12], double A[36]) { SE3::fromMat34(T).Adj(A); }
void orc_se3_mul(const double A[12], const double B[12], double C[12]) { (SE3::fromMat34(A) * SE3::fromMat34(B)).toMat34(C); }
void orc_se3_inv(const double A[12], double B[12]) { SE3::fromMat34(A).inverse().toMat34(B); }
void orc_ldlt_solve(int n, const double* A, const double* b, double* x) { ldlt_solve(n, A, b, x); }

}  // extern "C"
