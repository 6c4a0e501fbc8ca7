// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// The live windowed BA of the fork: FullSystem::optimize, g2o body (FullSystemOptimize.cpp:404-868): graph of
// EdgeLBASE3PosePhotoIdepthCamDSO edges (E2, oracle/lba_edge.cpp) over {VertexCamDSO, VertexSE3PoseDSO + VertexPhotometricDSO per
// host frame, one marginalised VertexInverseDepthDSO PER RESIDUAL}, Huber(9) on each edge's 8-vector chi2, driven by g2o's
// Levenberg-Marquardt with Schur complement. g2o is NOT part of the reference tree and no version is pinned (CMakeLists.txt:47-60):
// the driver below is RESTATED from upstream g2o behaviour — iteration-level parity is "device vs this
// restatement" and is unpinned against the real library.
#include "oracle_ba.hpp"

namespace orc {

namespace {
struct EdgeState {
  int res = 0, host = 0;
  bool active = false;          // level 0 at initializeOptimization()
  int level = 0;
  double idepth = 0, idepth_backup = 0;
  double err[8] = {0};
  double J[8][13];              // [pose 6 | photo 2 | cam 4 | idepth 1]; kept across calls (stale where linearizeOplus returns early)
  // quantities of the last buildSystem
  double hll = 0, bl = 0, hpl[12] = {0};
};
inline double huber_rho(double e2, double delta, double& rho1) {
  if (e2 <= delta * delta) { rho1 = 1.0; return e2; }
  const double sq = std::sqrt(e2);
  rho1 = delta / sq;
  return 2 * sq * delta - delta * delta;
}
}  // namespace

// estimates: cam[4], T_wh[n][12] (hosts), photo[n][2]; idepth per residual in/out. Returns the number of LM iterations run.
int lbaG2O(BAWindow& W, int mnumOptIts, double cam[4], double* T_wh, double* photo, double* idepth_io, int* used_host, double* chi2_out,
           int* newState_out, float* center_out, float* idepth_hessian_out, int* trials_out) {
  const int nf = W.n(), R = (int)W.res.size(), d = CPARS + 8 * nf;
  if (nf < 2) return 0;
  if (nf < 3) mnumOptIts = 10; else if (nf < 4) mnumOptIts = 7; else mnumOptIts = 3;
  std::vector<EdgeState> E(R);
  std::vector<double> b0(nf, 0.0);
  std::vector<SE3> pose(nf);
  for (int h = 0; h < nf; h++) { pose[h] = SE3::fromMat34(T_wh + 12 * h); used_host[h] = 0; }
  // graph build: resetOOB, vertices, edges, first computeError (:438-542)
  for (int i = 0; i < R; i++) {
    BARes& r = W.res[i];
    E[i].res = i; E[i].host = r.host; E[i].idepth = idepth_io[i];
    memset(E[i].J, 0, sizeof(E[i].J));
    if (r.isLinearized) continue;  // not in activeResiduals
    r.state_NewEnergy = r.state_energy = 0; r.state_NewState = RS_OUTLIER; r.state_state = RS_IN;
    if (!used_host[r.host]) { used_host[r.host] = 1; b0[r.host] = photo[2 * r.host + 1]; }  // SetB(a0b0.b) of the host's photo vertex
    E[i].active = true;
  }
  auto evalEdge = [&](int i, bool linearize) {
    EdgeState& e = E[i];
    BARes& r = W.res[i];
    LBAEdgeOut o;
    lbaEdgeEval(W, r, pose[e.host], photo + 2 * e.host, e.idepth, cam, b0[e.host], o);
    for (int k = 0; k < 8; k++) e.err[k] = o.error[k];
    r.state_NewState = o.newState;
    if (o.newEnergy >= 0) { r.state_NewEnergy = o.newEnergy; r.state_NewEnergyWithOutlier = o.newEnergyWithOutlier; }
    if (o.level == 1) e.level = 1;
    if (o.center_set) for (int k = 0; k < 3; k++) r.centerProjectedTo[k] = o.centerProjectedTo[k];  // SetCenterProjectedTo at the centre pixel (:66-69)
    if (linearize && e.level != 1 && o.newState != RS_OOB && o.idepth_hessian > 0) {
      for (int k = 0; k < 8; k++) {
        for (int c = 0; c < 6; c++) e.J[k][c] = o.J_xi[k][c];
        e.J[k][6] = o.J_photo[k][0]; e.J[k][7] = o.J_photo[k][1];
        for (int c = 0; c < 4; c++) e.J[k][8 + c] = o.J_C[k][c];
        e.J[k][12] = o.J_idepth[k];
      }
      idepth_hessian_out[i] = o.idepth_hessian;
    }
  };
  for (int i = 0; i < R; i++) if (E[i].active) { evalEdge(i, false); }
  // applyRes_Reductor(true) (:548-552): isActive follows the state of the first computeError
  for (int i = 0; i < R; i++) if (E[i].active) W.applyRes(W.res[i], true);
  // initializeOptimization(): the active set is fixed here (level 0)
  for (int i = 0; i < R; i++) if (E[i].active && E[i].level != 0) E[i].active = false;
  auto computeActiveErrors = [&]() { for (int i = 0; i < R; i++) if (E[i].active) evalEdge(i, false); };
  auto activeRobustChi2 = [&]() { double s = 0; for (int i = 0; i < R; i++) if (E[i].active) { double e2 = 0, r1; for (int k = 0; k < 8; k++) e2 += E[i].err[k] * E[i].err[k]; s += huber_rho(e2, W.S.huberTH, r1); } return s; };
  std::vector<double> Hpp((size_t)d * d), bp(d), x(d);
  double lambda = 0, ni = 2;
  double lastChi = 0;
  int it = 0, total_trials = 0;
  bool stop = false;
  for (; it < mnumOptIts && !stop; it++) {
    computeActiveErrors();
    double currentChi = activeRobustChi2();
    // buildSystem: linearizeOplus + robustified quadratic form of every active edge
    std::fill(Hpp.begin(), Hpp.end(), 0.0); std::fill(bp.begin(), bp.end(), 0.0);
    for (int i = 0; i < R; i++) {
      EdgeState& e = E[i];
      if (!e.active) continue;
      evalEdge(i, true);
      double e2 = 0, rho1;
      for (int k = 0; k < 8; k++) e2 += e.err[k] * e.err[k];
      huber_rho(e2, W.S.huberTH, rho1);
      const int hb = CPARS + 8 * e.host;
      auto col = [&](int c) { return c < 8 ? hb + c : c - 8; };  // local [pose6 photo2 cam4] -> global index
      e.hll = 0; e.bl = 0;
      for (int a = 0; a < 12; a++) {
        double ba = 0, hpl = 0;
        for (int k = 0; k < 8; k++) { ba += e.J[k][a] * rho1 * e.err[k]; hpl += e.J[k][a] * rho1 * e.J[k][12]; }
        bp[col(a)] -= ba; e.hpl[a] = hpl;
        for (int c = 0; c < 12; c++) { double s = 0; for (int k = 0; k < 8; k++) s += e.J[k][a] * rho1 * e.J[k][c]; Hpp[(size_t)col(a) * d + col(c)] += s; }
      }
      for (int k = 0; k < 8; k++) { e.hll += e.J[k][12] * rho1 * e.J[k][12]; e.bl -= e.J[k][12] * rho1 * e.err[k]; }
    }
    if (it == 0) { lambda = 0.1; ni = 2; }  // setUserLambdaInit(0.1) (:425)
    double rho = 0;
    int qmax = 0;
    do {
      // push(); H += lambda I on every vertex block; Schur over the marginalised idepth vertices; solve
      for (int i = 0; i < R; i++) E[i].idepth_backup = E[i].idepth;
      std::vector<SE3> pose_b = pose; std::vector<double> photo_b(photo, photo + 2 * nf); double cam_b[4] = {cam[0], cam[1], cam[2], cam[3]};
      std::vector<double> Hs = Hpp, bs = bp;
      for (int k = 0; k < d; k++) Hs[(size_t)k * d + k] += lambda;
      for (int i = 0; i < R; i++) {
        const EdgeState& e = E[i];
        if (!e.active) continue;
        const double inv = 1.0 / (e.hll + lambda);
        const int hb = CPARS + 8 * e.host;
        auto col = [&](int c) { return c < 8 ? hb + c : c - 8; };
        for (int a = 0; a < 12; a++) {
          bs[col(a)] -= e.hpl[a] * inv * e.bl;
          for (int c = 0; c < 12; c++) Hs[(size_t)col(a) * d + col(c)] -= e.hpl[a] * inv * e.hpl[c];
        }
      }
      for (int h = 0; h < nf; h++) if (!used_host[h]) for (int k = 0; k < 8; k++) { const int q = CPARS + 8 * h + k; Hs[(size_t)q * d + q] = 1.0; bs[q] = 0.0; }
      bool ok = ldlt_solve(d, Hs.data(), bs.data(), x.data());
      double tempChi = std::numeric_limits<double>::max();
      double scale = 0;
      if (ok) {
        // update: cam += dx, pose = exp(dx) * pose, photo += dx, idepth_r += (bl - hpl^T dx) / (hll + lambda)
        for (int k = 0; k < 4; k++) cam[k] += x[k];
        for (int h = 0; h < nf; h++) if (used_host[h]) {
          pose[h] = SE3::exp(&x[CPARS + 8 * h]) * pose[h];
          photo[2 * h] += x[CPARS + 8 * h + 6]; photo[2 * h + 1] += x[CPARS + 8 * h + 7];
        }
        for (int k = 0; k < d; k++) scale += x[k] * (lambda * x[k] + bp[k]);
        for (int i = 0; i < R; i++) {
          EdgeState& e = E[i];
          if (!e.active) continue;
          const int hb = CPARS + 8 * e.host;
          double s = e.bl;
          for (int a = 0; a < 12; a++) s -= e.hpl[a] * x[a < 8 ? hb + a : a - 8];
          const double dl = s / (e.hll + lambda);
          e.idepth += dl;
          scale += dl * (lambda * dl + e.bl);
        }
        computeActiveErrors();
        tempChi = activeRobustChi2();
      }
      rho = (currentChi - tempChi) / (scale + 1e-3);
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        lambda *= std::max(1. / 3., alpha);
        ni = 2; currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2;
        for (int i = 0; i < R; i++) E[i].idepth = E[i].idepth_backup;   // pop()
        pose = pose_b; for (int k = 0; k < 2 * nf; k++) photo[k] = photo_b[k]; for (int k = 0; k < 4; k++) cam[k] = cam_b[k];
        if (!std::isfinite(lambda)) break;
      }
      qmax++; total_trials++;
    } while (rho < 0 && qmax < 10);
    const bool terminate_lm = (qmax == 10 || rho == 0 || !std::isfinite(lambda));
    // SparseOptimizerTerminateAction (gain threshold 1e-3) runs after every iteration
    computeActiveErrors();
    const double chi = activeRobustChi2();
    if (it == 0) lastChi = chi;
    else { const double gain = (lastChi - chi) / chi; lastChi = chi; if (gain >= 0 && gain < 1e-3) stop = true; }
    if (terminate_lm) { it++; break; }
  }
  for (int h = 0; h < nf; h++) pose[h].toMat34(T_wh + 12 * h);
  for (int i = 0; i < R; i++) {
    idepth_io[i] = E[i].idepth;
    newState_out[i] = W.res[i].state_NewState;
    for (int k = 0; k < 3; k++) center_out[3 * i + k] = W.res[i].centerProjectedTo[k];
  }
  *chi2_out = lastChi;
  *trials_out = total_trials;
  return it;
}

}  // namespace orc
