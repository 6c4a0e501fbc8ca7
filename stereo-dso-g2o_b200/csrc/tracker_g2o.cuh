// g2o path of the tracker on the device: EdgeSE3PosePhotoDSO (dso_g2o_edge.cpp:395-500), the edge
// construction of calcRes (CoarseTracker.cpp:645-728) and the Levenberg driver g2o runs on the
// 8-dimensional problem {VertexSE3PoseDSO, VertexPhotometricDSO} (restated, SURVEY.md Appendix C;
// g2o itself is not in the reference tree: iteration-level parity is "restated-g2o vs GPU").
// Included by tracker.cu (shares its structs and reduction).
#pragma once
// (included inside namespace sdso)

// accumulator layout of the g2o passes
enum { G_H = 0, G_B = 36, G_NE = 44, G_NSAT = 45, G_ST = 46, G_SRT = 47, G_SN = 48 };

struct G2OConst {       // per pass, written by thread 0
  float RKi[9], t[3];   // selection pose (refToNew_current, float path of calcRes :617-618)
  double R[9], tt[3];   // vertex pose estimate (double)
  float ab[2];          // (float) fromToVecExposure(ref, new, a0b0, vertex photo)
  double b0;            // a0b0_.b
  float cutoff10;       // cutoffTH*10 (:723)
};

// dso_util.hpp:25-45
__device__ __forceinline__ bool d_check_boundary(double u, double v, int wl, int hl) {
  return (u - 2) < 0 || (u + 3) > wl || (v - 2) < 0 || (v + 3) > hl;
}

// g2o RobustKernelHuber::robustify -> rho[0], rho[1]
__device__ __forceinline__ void d_huber(double e2, double delta, double& rho0, double& rho1) {
  const double dsqr = delta * delta;
  if (e2 <= dsqr) { rho0 = e2; rho1 = 1.0; }
  else { const double sq = sqrt(e2); rho0 = 2 * sq * delta - dsqr; rho1 = delta / sq; }
}

// MODE 0: calcRes edge construction (+ first computeError)   MODE 1: computeActiveErrors
// MODE 2: computeActiveErrors + buildSystem (linearizeOplus + constructQuadraticForm)
// BUILD: also linearizeOplus + constructQuadraticForm of the edges that stay. <0, true> is the first pass of a level: the first LM
// iteration's computeActiveErrors + buildSystem run at the estimate the edges were built at, so their errors, Jacobians and sums are
// the ones of this pass (same gc, same points in the same order per thread) and are not recomputed.
template <int MODE, bool BUILD = (MODE == 2)>
__device__ void eval_points_g2o(const TrackParams& P, const TrackLevel& L, int lvl, const float4* __restrict__ tex, const G2OConst& gc,
                                unsigned char* __restrict__ flag, double* __restrict__ eerr, float (&acc)[kAccPad], double& chi,
                                unsigned& evals, int gtid, int gthreads, double* dump, const float4* __restrict__ pc, const int npc) {
#pragma unroll
  for (int k = 0; k < kAccPad; k++) acc[k] = 0.f;
  chi = 0.0;
  const int wl = L.w, hl = L.h;
  const double delta = P.huberTH;
  // flow indicators (:662-693) of every 32nd point, level 0, edge construction only — a separate dense pass: inside the main
  // loop exactly one lane of each warp would take the branch
  if (MODE == 0 && lvl == 0) {
    for (int i = 32 * gtid; i < npc; i += 32 * gthreads) {
      const float4 p = __ldg(pc + i);
      const float x = p.x, y = p.y, id = p.z;
      float pt[3], ptT[3], ptT2[3], pt3[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float kp = L.Ki[r * 3 + 0] * x + L.Ki[r * 3 + 1] * y + L.Ki[r * 3 + 2];
        const float rp = gc.RKi[r * 3 + 0] * x + gc.RKi[r * 3 + 1] * y + gc.RKi[r * 3 + 2];
        ptT[r] = kp + gc.t[r] * id;
        ptT2[r] = kp - gc.t[r] * id;
        pt[r] = rp + gc.t[r] * id;
        pt3[r] = rp - gc.t[r] * id;
      }
      const float u = pt[0] / pt[2], v = pt[1] / pt[2];
      const float Ku = L.fx * u + L.cx, Kv = L.fy * v + L.cy;
      const float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      const float KuT = L.fx * uT + L.cx, KvT = L.fy * vT + L.cy;
      const float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      const float KuT2 = L.fx * uT2 + L.cx, KvT2 = L.fy * vT2 + L.cy;
      const float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      const float Ku3 = L.fx * u3 + L.cx, Kv3 = L.fy * v3 + L.cy;
      acc[G_ST] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      acc[G_ST] += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      acc[G_SRT] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      acc[G_SRT] += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      acc[G_SN] += 2.f;
    }
  }
  for (int i = gtid; i < npc; i += gthreads) {
    const float4 p = __ldg(pc + i);
    const float x = p.x, y = p.y, id = p.z;
    if (MODE == 0) {
      // selection with the float path at refToNew_current (:651-696)
      float pt[3];
#pragma unroll
      for (int r = 0; r < 3; r++) pt[r] = (gc.RKi[r * 3 + 0] * x + gc.RKi[r * 3 + 1] * y + gc.RKi[r * 3 + 2]) + gc.t[r] * id;
      const float u = pt[0] / pt[2], v = pt[1] / pt[2];
      const float Ku = L.fx * u + L.cx, Kv = L.fy * v + L.cy;
      const float new_idepth = id / pt[2];
      if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) {
        flag[i] = 0;
        if (dump) dump[i] = 0.0;
        continue;
      }
    } else if (!flag[i]) {
      continue;
    }
    // ---- the edge: Xref = Ki*(x,y,1)/id (:707), computeError (dso_g2o_edge.cpp:395-423)
    float Xr[3];
#pragma unroll
    for (int r = 0; r < 3; r++) Xr[r] = (L.Ki[r * 3 + 0] * x + L.Ki[r * 3 + 1] * y + L.Ki[r * 3 + 2]) / id;
    const double X0 = Xr[0], X1 = Xr[1], X2 = Xr[2];
    const double Xc = gc.R[0] * X0 + gc.R[1] * X1 + gc.R[2] * X2 + gc.tt[0];
    const double Yc = gc.R[3] * X0 + gc.R[4] * X1 + gc.R[5] * X2 + gc.tt[1];
    const double Zc = gc.R[6] * X0 + gc.R[7] * X1 + gc.R[8] * X2 + gc.tt[2];
    const double uu = L.gfx * (Xc / Zc) + L.gcx, vv = L.gfy * (Yc / Zc) + L.gcy;
    evals++;
    double err;
    float hity = 0.f, hitz = 0.f;
    bool oob = d_check_boundary(uu, vv, wl, hl);
    if (oob) {
      err = 0.0;
    } else {
      const float3 hit = interp33(tex, (float)uu, (float)vv, wl);
      hity = hit.y; hitz = hit.z;
      if (!isfinite(hit.x)) err = (MODE == 0) ? 0.0 : eerr[i];  // error left stale (:413-415)
      else err = (double)hit.x - ((double)gc.ab[0] * (double)p.w + (double)gc.ab[1]);
    }
    if (MODE == 0) {
      if (err > (double)gc.cutoff10) {  // :723, signed comparison
        acc[G_NSAT] += 1.f;
        flag[i] = 0;
        if (dump) dump[i] = 0.0;
        continue;
      }
      flag[i] = 1;
      acc[G_NE] += 1.f;
    }
    eerr[i] = err;
    double rho0, rho1;
    d_huber(err * err, delta, rho0, rho1);
    chi += rho0;
    if (BUILD || dump) {
      // linearizeOplus (dso_g2o_edge.cpp:425-500, VERSION2)
      double J[8];
      if (oob) {
#pragma unroll
        for (int k = 0; k < 8; k++) J[k] = 0.0;
      } else {
        const double invz = 1.0 / Zc;
        const double u = Xc * invz, v = Yc * invz;
        const double dx = (double)hity * L.gfx, dy = (double)hitz * L.gfy;
        J[0] = invz * dx;
        J[1] = invz * dy;
        J[2] = -invz * (u * dx + v * dy);
        J[3] = -(u * v * dx + (1 + v * v) * dy);
        J[4] = u * v * dy + (1 + u * u) * dx;
        J[5] = u * dy - v * dx;
        J[6] = (double)gc.ab[0] * (gc.b0 - (double)p.w);
        J[7] = -1.0;
      }
      if (BUILD) {
        // constructQuadraticForm: b += J^T (-rho' e), H += J^T rho' J  (accumulated in float; summed in double)
        const float wr = (float)rho1, wre = (float)(-err * rho1);
        float Jf[8];
#pragma unroll
        for (int k = 0; k < 8; k++) Jf[k] = (float)J[k];
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const float Jw = Jf[r] * wr;
#pragma unroll
          for (int c = r; c < 8; c++) { acc[G_H + idx] = __fmaf_rn(Jw, Jf[c], acc[G_H + idx]); idx++; }
          acc[G_B + r] = __fmaf_rn(Jf[r], wre, acc[G_B + r]);
        }
      }
      if (dump) {
        const int n = npc;
        dump[i] = 1.0;
        dump[(size_t)n + i] = err;
#pragma unroll
        for (int k = 0; k < 8; k++) dump[(size_t)(2 + k) * n + i] = J[k];
      }
    }
  }
}

__device__ void make_g2o_const(const TrackParams& P, const TrackLevel& L, const TrackProblem& prob, const double* Rsel, const double* tsel,
                               const double* R, const double* t, const double* photo, float cutoffTH, G2OConst& gc) {
  float Rf[9];
  for (int i = 0; i < 9; i++) Rf[i] = (float)Rsel[i];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) gc.RKi[r * 3 + c] = Rf[r * 3 + 0] * L.Ki[0 * 3 + c] + Rf[r * 3 + 1] * L.Ki[1 * 3 + c] + Rf[r * 3 + 2] * L.Ki[2 * 3 + c];
  for (int i = 0; i < 3; i++) gc.t[i] = (float)tsel[i];
  for (int i = 0; i < 9; i++) gc.R[i] = R[i];
  for (int i = 0; i < 3; i++) gc.tt[i] = t[i];
  double ab[2];
  d_aff_from_to(prob.ref_exposure, prob.exposure_new, prob.ref_aff[0], prob.ref_aff[1], photo[0], photo[1], ab);
  gc.ab[0] = (float)ab[0]; gc.ab[1] = (float)ab[1];
  gc.b0 = prob.ref_aff[1];
  gc.cutoff10 = cutoffTH * 10;
}

struct G2OState {  // shared memory, thread-0 owned, read by all after barriers
  double R[9], t[3], photo[2];        // vertex estimates
  double Rb[9], tb[3], photob[2];     // push()/pop() backup
  double Rsel[9], tsel[3];            // refToNew_current (never updated, CoarseTracker.cpp:880)
  double H[64], b[8], x[8];
  double lambda, ni, currentChi, tempChi, lastChi, rho;
  double levelChi;                    // activeRobustChi2 of the current level's edges
  int ok2, qmax, forceStop, okIter, again, accepted, have_sys;
  unsigned long long totalEdges;
};

__global__ void __launch_bounds__(256, 2) track_g2o_kernel(TrackParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackSmem* sm = reinterpret_cast<TrackSmem*>(smem_raw);
  __shared__ G2OConst gc;
  __shared__ G2OState gs;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const int prob_id = blockIdx.x / C;
  TrackProblem& prob = P.problems[prob_id];
  const int tid = threadIdx.x;
  const int gtid = rank * blockDim.x + tid, gthreads = C * blockDim.x;
  float acc[kAccPad];
  double chi = 0.0, chiTot = 0.0;
  unsigned evals = 0;
  Exchange parity;
  double* tot = sm->lm.total;

  if (tid == 0) {
    for (int i = 0; i < 9; i++) gs.R[i] = gs.Rsel[i] = prob.T[(i / 3) * 4 + (i % 3)];
    for (int i = 0; i < 3; i++) gs.t[i] = gs.tsel[i] = prob.T[i * 4 + 3];
    gs.photo[0] = prob.aff[0]; gs.photo[1] = prob.aff[1];
    gs.forceStop = 0; gs.totalEdges = 0; gs.lastChi = 0;
    mbar_init(&sm->bar[0], 1); mbar_init(&sm->bar[1], 1);
    mbar_fence_init();
  }
  cluster.sync();

  if (P.mode == 2) {  // ---- operator-level E1 evaluation: edges selected at T (prob.T), evaluated at (T_out, aff_out) ----
    const int lvl = P.eval_lvl;
    const TrackLevel& L = P.L[lvl];
    unsigned char* flag = P.edge_flag[lvl] + (size_t)prob_id * P.edge_stride[lvl];
    double* eerr = P.edge_err[lvl] + (size_t)prob_id * P.edge_stride[lvl];
    if (tid == 0) {
      double Rp[9], tp[3];
      for (int i = 0; i < 9; i++) Rp[i] = prob.T_out[(i / 3) * 4 + (i % 3)];
      for (int i = 0; i < 3; i++) tp[i] = prob.T_out[i * 4 + 3];
      make_g2o_const(P, L, prob, gs.Rsel, gs.tsel, Rp, tp, prob.aff_out, P.eval_cutoff, gc);
    }
    __syncthreads();
    eval_points_g2o<0>(P, L, lvl, prob.tex[lvl], gc, flag, eerr, acc, chi, evals, gtid, gthreads, P.dump_d, prob.pc[lvl], prob.pc_n[lvl]);
    reduce_all(acc, sm, parity, cluster, tot, chi, &chiTot);
    if (rank == 0 && tid == 0) prob.warped_n = (int)tot[G_NE];
    cluster.sync();
    return;
  }

  double lastRes[5] = {NAN, NAN, NAN, NAN, NAN};
  double flow[3] = {1000, 1000, 1000};
  int iters[5] = {0, 0, 0, 0, 0};
  bool aborted = false;
  const int maxIterations = 2;  // CoarseTracker.cpp:863 {2,2,2,2,2}

  for (int lvl = P.coarsest; lvl >= 0 && !aborted; lvl--) {
    const TrackLevel& L = P.L[lvl];
    const float4* tex = prob.tex[lvl];
    unsigned char* flag = P.edge_flag[lvl] + (size_t)prob_id * P.edge_stride[lvl];
    double* eerr = P.edge_err[lvl] + (size_t)prob_id * P.edge_stride[lvl];
    // ---- calcRes: build this level's edges (:894). The pass also linearises the edges it keeps: the first iteration's
    // computeActiveErrors + buildSystem below run at this same estimate (see eval_points_g2o)
    if (tid == 0) make_g2o_const(P, L, prob, gs.Rsel, gs.tsel, gs.R, gs.t, gs.photo, P.coarseCutoffTH, gc);
    __syncthreads();
    eval_points_g2o<0, true>(P, L, lvl, tex, gc, flag, eerr, acc, chi, evals, gtid, gthreads, nullptr, prob.pc[lvl], prob.pc_n[lvl]);
    reduce_all(acc, sm, parity, cluster, tot, chi, &chiTot);
    const int nEdges = (int)tot[G_NE];
    {
      const float sT = (float)tot[G_ST], sRT = (float)tot[G_SRT], sN = (float)tot[G_SN];
      flow[0] = sT / (sN + 0.1); flow[1] = 0; flow[2] = sRT / (sN + 0.1);
    }
    auto take_system = [&]() {   // thread 0: H, b and chi2 of the pass just reduced
      int idx = 0;
      for (int r = 0; r < 8; r++) {
        for (int c = r; c < 8; c++) { gs.H[r * 8 + c] = gs.H[c * 8 + r] = tot[G_H + idx]; idx++; }
        gs.b[r] = tot[G_B + r];
      }
      gs.currentChi = chiTot; gs.tempChi = chiTot;
    };
    if (tid == 0) { gs.totalEdges += (unsigned long long)nEdges; gs.levelChi = chiTot; gs.okIter = 1; gs.have_sys = 0; take_system(); }
    __syncthreads();

    // ---- optimizer->initializeOptimization(lvl); optimize(maxIterations) (:923-926)
    if (nEdges > 0) {
      for (int it = 0; it < maxIterations; it++) {
        if (gs.forceStop || !gs.okIter) break;  // uniform: written before the last barrier
        iters[lvl]++;
        // LM.solve(it): computeActiveErrors + buildSystem at the current estimate (gc is current: it was last written for exactly
        // this estimate — by the level's first pass, or behind the last trial of the previous iteration). When the previous
        // iteration ended on an accepted trial, that trial's pass already linearised the edges at this estimate (have_sys).
        if (it > 0 && !gs.have_sys) {
          eval_points_g2o<2>(P, L, lvl, tex, gc, flag, eerr, acc, chi, evals, gtid, gthreads, nullptr, prob.pc[lvl], prob.pc_n[lvl]);
          reduce_all(acc, sm, parity, cluster, tot, chi, &chiTot);
          if (tid == 0) take_system();
        }
        const bool speculate = it + 1 < maxIterations;   // a later iteration of this level could reuse the trial's linearisation
        if (tid == 0) {
          if (it == 0) { gs.lambda = 0.01; gs.ni = 2; }  // setUserLambdaInit(0.01) (:840)
          gs.rho = 0; gs.qmax = 0;
        }
        __syncthreads();
        do {  // trials
          if (tid < 32) {
            // (H + lambda I) x = b by warp 0 (LinearSolverEigen's LLT succeeds exactly when every pivot is finite and positive)
            double row[9], xs[8];
            const int l = tid & 7;
#pragma unroll
            for (int j = 0; j < 8; j++) row[j] = gs.H[l * 8 + j] + (j == l ? gs.lambda : 0.0);
            row[8] = gs.b[l];
            bool pd;
            warp_solve8(row, tid, xs, pd);
            __syncwarp();
            if (tid == 0) {
              for (int i = 0; i < 9; i++) gs.Rb[i] = gs.R[i];
              for (int i = 0; i < 3; i++) gs.tb[i] = gs.t[i];
              gs.photob[0] = gs.photo[0]; gs.photob[1] = gs.photo[1];  // push()
              if (!pd) { for (int i = 0; i < 8; i++) xs[i] = 0; }     // the solver failed: x stays zero
              for (int i = 0; i < 8; i++) gs.x[i] = xs[i];
              gs.ok2 = pd ? 1 : 0;
              double Rn[9], tn[3];
              d_se3_exp_mul(xs, gs.R, gs.t, Rn, tn);  // oplus (dso_g2o_vertex.cpp:15-18)
              for (int i = 0; i < 9; i++) gs.R[i] = Rn[i];
              for (int i = 0; i < 3; i++) gs.t[i] = tn[i];
              gs.photo[0] += gs.x[6]; gs.photo[1] += gs.x[7];  // (:30-40)
              make_g2o_const(P, L, prob, gs.Rsel, gs.tsel, gs.R, gs.t, gs.photo, P.coarseCutoffTH, gc);
            }
          }
          __syncthreads();
          // computeActiveErrors at the trial estimate. With another iteration to come the pass also linearises (speculatively):
          // if the trial is accepted, the next iteration's buildSystem would evaluate the same edges at the same estimate
          if (speculate) eval_points_g2o<1, true>(P, L, lvl, tex, gc, flag, eerr, acc, chi, evals, gtid, gthreads, nullptr, prob.pc[lvl], prob.pc_n[lvl]);
          else eval_points_g2o<1, false>(P, L, lvl, tex, gc, flag, eerr, acc, chi, evals, gtid, gthreads, nullptr, prob.pc[lvl], prob.pc_n[lvl]);
          reduce_all(acc, sm, parity, cluster, tot, chi, &chiTot);
          if (tid == 0) {
            double tempChi = chiTot;
            if (!gs.ok2) tempChi = 1.7976931348623157e308;
            double rho = gs.currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 8; j++) scale += gs.x[j] * (gs.lambda * gs.x[j] + gs.b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
              double alpha = 1. - pow((2 * rho - 1), 3);
              alpha = fmin(alpha, 2. / 3.);
              const double scaleFactor = fmax(1. / 3., alpha);
              gs.lambda *= scaleFactor; gs.ni = 2;
              gs.accepted = 1;
              if (speculate) take_system();   // (scale above used the old b; x is not read again)
              gs.currentChi = tempChi;
            } else {
              gs.lambda *= gs.ni; gs.ni *= 2;
              for (int i = 0; i < 9; i++) gs.R[i] = gs.Rb[i];
              for (int i = 0; i < 3; i++) gs.t[i] = gs.tb[i];
              gs.photo[0] = gs.photob[0]; gs.photo[1] = gs.photob[1];  // pop()
              gs.accepted = 0;
            }
            gs.have_sys = (gs.accepted && speculate) ? 1 : 0;   // read by every thread behind the iteration's last barrier
            gs.rho = rho;
            gs.qmax++;
            gs.again = (rho < 0 && gs.qmax < 10) ? 1 : 0;
            if (!gs.again && (gs.qmax == 10 || rho == 0)) gs.okIter = 0;  // SolverResult::Terminate
            // after a rejected last trial gc has to follow the estimate back (an accepted one left it where it is)
            if (!gs.again && !gs.accepted) make_g2o_const(P, L, prob, gs.Rsel, gs.tsel, gs.R, gs.t, gs.photo, P.coarseCutoffTH, gc);
          }
          __syncthreads();
        } while (gs.again);
        // postIteration(it): SparseOptimizerTerminateAction -> computeActiveErrors, gain test (1e-3, :845-848). Behind an ACCEPTED
        // trial that pass re-evaluates the edges at the estimate the trial pass just evaluated them at (same gc, same order): its
        // errors and its chi2 are the trial's, bit for bit, so it is not run. Behind a rejected one the estimate went back and the
        // pass runs (the stale-error rule of computeError, :413-415, makes it differ from the build pass on non-finite pixels).
        if (!gs.accepted) {
          eval_points_g2o<1>(P, L, lvl, tex, gc, flag, eerr, acc, chi, evals, gtid, gthreads, nullptr, prob.pc[lvl], prob.pc_n[lvl]);
          reduce_all(acc, sm, parity, cluster, tot, chi, &chiTot);
        }
        if (tid == 0) {
          const double chiPost = gs.accepted ? gs.currentChi : chiTot;
          gs.levelChi = chiPost;
          if (it == 0) gs.lastChi = chiPost;
          else {
            const double gain = (gs.lastChi - chiPost) / chiPost;
            gs.lastChi = chiPost;
            if (gain >= 0 && gain < 1e-3) gs.forceStop = 1;
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    // :1029 — activeRobustChi2 over this level's edges / ALL edges added so far
    lastRes[lvl] = sqrtf((float)gs.levelChi / gs.totalEdges);
    if (!P.g2o_stop_persists) { __syncthreads(); if (tid == 0) gs.forceStop = 0; }
    if (lastRes[lvl] > 1.5 * prob.minResForAbort[lvl]) { aborted = true; break; }
  }
  __syncthreads();
  if (rank == 0 && tid == 0) {
    bool ok = !aborted;
    double aout[2] = {gs.photo[0], gs.photo[1]};
    if (ok) {
      if ((P.affineOptModeA != 0 && (fabsf((float)aout[0]) > 1.2)) || (P.affineOptModeB != 0 && (fabsf((float)aout[1]) > 200))) ok = false;
    }
    if (ok) {
      double rel[2];
      d_aff_from_to(prob.ref_exposure, prob.exposure_new, prob.ref_aff[0], prob.ref_aff[1], aout[0], aout[1], rel);
      const float r0 = (float)rel[0], r1 = (float)rel[1];
      if ((P.affineOptModeA == 0 && (fabsf(logf(r0)) > 1.5)) || (P.affineOptModeB == 0 && (fabsf(r1) > 200))) ok = false;
    }
    if (ok) {
      if (P.affineOptModeA < 0) aout[0] = 0;
      if (P.affineOptModeB < 0) aout[1] = 0;
    }
    if (!aborted) {
      for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) prob.T_out[r * 4 + c] = gs.R[r * 3 + c]; prob.T_out[r * 4 + 3] = gs.t[r]; }
      prob.aff_out[0] = aout[0]; prob.aff_out[1] = aout[1];
    } else {
      for (int i = 0; i < 12; i++) prob.T_out[i] = prob.T[i];
      prob.aff_out[0] = prob.aff[0]; prob.aff_out[1] = prob.aff[1];
    }
    for (int i = 0; i < 5; i++) { prob.lastResiduals[i] = lastRes[i]; prob.iterations[i] = iters[i]; }
    for (int i = 0; i < 3; i++) prob.flow[i] = flow[i];
    prob.ok = ok ? 1 : 0;
  }
  {
    unsigned e = evals;
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((tid & 31) == 0) atomicAdd(&prob.evals, (unsigned long long)e);
  }
  cluster.sync();
}

