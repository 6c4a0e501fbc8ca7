// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// E2: EdgeLBASE3PosePhotoIdepthCamDSO::computeError + linearizeOplus (dso_g2o_edge.cpp:5-128, 130-282), operator level.
// Mixed float/double arithmetic placed as the reference places it (Mat33f R, Vec3f t, Vec3f Klip / ptp; double drescale,
// _u, _v, _Ku, _Kv and Jacobians; float interpolation, weights and energies).
#include "oracle_ba.hpp"

namespace orc {

static inline bool CheckBoundaryD(double u, double v, int wl, int hl) { return (u - 2) < 0 || (u + 3) > wl || (v - 2) < 0 || (v + 3) > hl; }

void lbaEdgeEval(const BAWindow& W, const BARes& r, const SE3& T_wh, const double photo[2], double idepth, const double cam[4], double b0,
                 LBAEdgeOut& o) {
  memset(&o, 0, sizeof(o));
  o.newEnergy = -1; o.newEnergyWithOutlier = -1; o.newState = r.state_NewState; o.level = 0;
  const BAPoint& point = W.points[r.point];
  const BAFrame& host = W.frames[r.host]; const BAFrame& target = W.frames[r.target];
  const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  const SE3 Tth = target.PRE_worldToCam * T_wh;
  double Rd[9]; Tth.rotationMatrix(Rd);
  float R[9], t[3];
  for (int i = 0; i < 9; i++) R[i] = (float)Rd[i];
  for (int i = 0; i < 3; i++) t[i] = (float)Tth.t[i];
  const float* dIl = target.img->dIp[0].data();
  const int w0 = W.G->w[0], wl = W.G->w[0] - 3, hl = W.G->h[0] - 3;
  double at[2]; target.aff_g2l(at);
  double abd[2]; affFromToVecExposure(host.img->ab_exposure, target.img->ab_exposure, photo[0], photo[1], at[0], at[1], abd);
  const float ab[2] = {(float)abd[0], (float)abd[1]};
  float energyLeft = 0, wJI2_sum = 0;
  // per-pixel values shared by computeError and linearizeOplus
  double drescale_[8], u_[8], v_[8], nid_[8]; float hit_[8][3]; float Klip_[8][2];
  for (int idx = 0; idx < patternNum; idx++) {
    const double u_host = point.u + patternP[idx][0], v_host = point.v + patternP[idx][1];
    const float Klip[3] = {(float)((u_host - cx) / fx), (float)((v_host - cy) / fy), 1.0f};
    float ptp[3];
    const float idf = (float)idepth;
    for (int k = 0; k < 3; k++) ptp[k] = (R[k * 3] * Klip[0] + R[k * 3 + 1] * Klip[1] + R[k * 3 + 2] * Klip[2]) + t[k] * idf;
    const double drescale = 1.0f / ptp[2];
    if (drescale <= 0) { o.newState = RS_OOB; for (int i = 0; i < 8; i++) o.error[i] = 0; return; }
    const double new_idepth = idepth * drescale;
    const double _u = ptp[0] * drescale, _v = ptp[1] * drescale;
    const double _Ku = _u * fx + cx, _Kv = _v * fy + cy;
    if (CheckBoundaryD(_Ku, _Kv, wl, hl)) { o.newState = RS_OOB; for (int i = 0; i < 8; i++) o.error[i] = 0; o.level = 1; return; }
    if (patternP[idx][0] == 0 && patternP[idx][1] == 0) { o.centerProjectedTo[0] = (float)_Ku; o.centerProjectedTo[1] = (float)_Kv; o.centerProjectedTo[2] = (float)new_idepth; o.center_set = 1; }
    float hit[3];
    getInterpolatedElement33(dIl, (float)_Ku, (float)_Kv, w0, hit);
    drescale_[idx] = drescale; u_[idx] = _u; v_[idx] = _v; nid_[idx] = new_idepth;
    hit_[idx][0] = hit[0]; hit_[idx][1] = hit[1]; hit_[idx][2] = hit[2]; Klip_[idx][0] = Klip[0]; Klip_[idx][1] = Klip[1];
    if (!std::isfinite((float)hit[0])) { o.newState = RS_OOB; o.error[idx] = 0; continue; }
    o.error[idx] = hit[0] - (ab[0] * point.color[idx] + ab[1]);
    float w = sqrtf(W.S.outlierTHSumComponent / (W.S.outlierTHSumComponent + (hit[1] * hit[1] + hit[2] * hit[2])));
    w = 0.5f * (w + point.weights[idx]);
    const float hw = fabsf((float)o.error[idx]) < W.S.huberTH ? 1 : W.S.huberTH / fabsf((float)o.error[idx]);
    energyLeft += w * w * hw * o.error[idx] * o.error[idx] * (2 - hw);
    wJI2_sum += hw * hw * (hit[1] * hit[1] + hit[2] * hit[2]);
  }
  o.newEnergyWithOutlier = energyLeft;
  const float th = std::max<float>(host.frameEnergyTH, target.frameEnergyTH);
  if (energyLeft > th || wJI2_sum < 2) { energyLeft = th; o.newState = RS_OUTLIER; }
  else o.newState = RS_IN;
  o.newEnergy = energyLeft;
  // ---- linearizeOplus (:130-282); level()==1 / OOB returned above
  float H_idepth_idepth = 0;
  for (int idx = 0; idx < patternNum; idx++) {
    if (!std::isfinite(hit_[idx][0])) {  // (:205-208) returns before _jacobianOplus is assigned: reported as zero here
      o.newState = RS_OOB;
      memset(o.J_xi, 0, sizeof(o.J_xi)); memset(o.J_photo, 0, sizeof(o.J_photo)); memset(o.J_idepth, 0, sizeof(o.J_idepth)); memset(o.J_C, 0, sizeof(o.J_C));
      return;
    }
    const double drescale = drescale_[idx], _u = u_[idx], _v = v_[idx], new_idepth = nid_[idx];
    const float* hit = hit_[idx];
    const double fxi = 1 / fx, fyi = 1 / fy;
    double dC[2][4];
    dC[0][2] = drescale * (R[6] * _u - R[0]);
    dC[0][3] = fx * fyi * drescale * (R[7] * _u - R[1]);
    dC[0][0] = Klip_[idx][0] * dC[0][2];
    dC[0][1] = Klip_[idx][1] * dC[0][3];
    dC[1][2] = fy * fxi * drescale * (R[6] * _v - R[3]);
    dC[1][3] = drescale * (R[7] * _v - R[4]);
    dC[1][0] = Klip_[idx][0] * dC[1][2];
    dC[1][1] = Klip_[idx][1] * dC[1][3];
    for (int k = 0; k < 4; k++) o.J_C[idx][k] = (double)hit[1] * dC[0][k] + (double)hit[2] * dC[1][k];
    const double dx = hit[1] * fx, dy = hit[2] * fy;
    o.J_xi[idx][0] = new_idepth * dx;
    o.J_xi[idx][1] = new_idepth * dy;
    o.J_xi[idx][2] = -new_idepth * (_u * dx + _v * dy);
    o.J_xi[idx][3] = -(_u * _v * dx + (1 + _v * _v) * dy);
    o.J_xi[idx][4] = _u * _v * dy + (1 + _u * _u) * dx;
    o.J_xi[idx][5] = _u * dy - _v * dx;
    o.J_photo[idx][0] = ab[0] * (b0 - point.color[idx]);
    o.J_photo[idx][1] = -1;
    o.J_idepth[idx] = dx * drescale * (t[0] - t[2] * _u) + dy * drescale * (t[1] - t[2] * _v);
    H_idepth_idepth += o.J_idepth[idx] * o.J_idepth[idx];
  }
  if (H_idepth_idepth < 1e-10) H_idepth_idepth = 1e-10;
  o.idepth_hessian = H_idepth_idepth;
}

}  // namespace orc
