// Device-side data model of the sliding window (SoA arenas indexed by integer ids).
#pragma once
#include "ctx.h"

namespace sdso {

constexpr int kMaxFrames = 16;        // window size supported on the device (reference: setting_maxFrames 7; SURVEY config 4 uses 10)
constexpr int kJ = 74;                // floats of RawResidualJacobian (OptimizationBackend/RawResidualJacobian.h:32-65)
// offsets inside the 74-float record
enum { J_RESF = 0, J_PDXI = 8, J_PDC = 20, J_PDD = 28, J_IDX = 30, J_AB = 46, J_IDX2 = 62, J_ABIDX = 66, J_AB2 = 70 };

enum { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };  // Residuals.h:49
enum { RF_LINEARIZED = 1, RF_ACTIVE = 2 };       // EFResidual::isLinearized / isActiveAndIsGoodNEW

struct BAFrameDev {
  const float4* tex0;     // level-0 texels of the frame (target->dI)
  float ab_exposure;
  float frameEnergyTH;
  int frameID;
  int pad;
  double R0[9], t0[3];    // worldToCam_evalPT
  double state[10], state_zero[10], state_scaled[10];
  double Rw[9], tw[3];    // PRE_worldToCam
  double Rc[9], tc[3];    // PRE_camToWorld
  double prior[8], delta_prior[8], delta[8];
};

struct PrecalcDev {  // FrameFramePrecalc (HessianBlocks.h:72-97), index host*n + target
  float PRE_RTll[9], PRE_KRKiTll[9], PRE_RKiTll[9], PRE_RTll_0[9];
  float PRE_tTll[3], PRE_KtTll[3], PRE_tTll_0[3];
  float PRE_aff_mode[2];
  float PRE_b0_mode, distanceLL;
  float pad[3];
};

struct BACalib {  // CalibHessian value_scaledf / value_scaledi + wM3G, hM3G
  float fxl, fyl, cxl, cyl, fxli, fyli, cxli, cyli;
  float wM3G, hM3G;
  int w0, h0;
  float huberTH, outlierTHSumComponent, affineOptModeA, affineOptModeB;
};

struct BAState {
  int n = 0, P = 0, R = 0;
  bool prepared = false;
  BACalib calib;
  double calib_delta[4] = {0, 0, 0, 0};
  double cPrior[4];
  // frames
  BAFrameDev* d_frames = nullptr;
  PrecalcDev* d_precalc = nullptr;
  double* d_adHost = nullptr; double* d_adTarget = nullptr;   // [n*n][64], index h + t*n
  float* d_adHostF = nullptr; float* d_adTargetF = nullptr;
  float* d_adHTdeltaF = nullptr;                               // [n*n][8]
  float* d_cDeltaF = nullptr;                                  // [4]
  // points
  int cap_points = 0;
  int* d_p_host = nullptr; float* d_p_u = nullptr; float* d_p_v = nullptr;
  float* d_p_idepth = nullptr; float* d_p_idepth_zero = nullptr;
  float* d_p_color = nullptr; float* d_p_weights = nullptr;   // [P][8]
  float* d_p_priorF = nullptr; float* d_p_deltaF = nullptr;
  int* d_p_res_begin = nullptr;                                // CSR [P+1]
  float* d_p_acc = nullptr;                                    // [P][16]: Hdd_A, bd_A, Hcd_A[4], Hdd_L, bd_L, Hcd_L[4], HdiF, bdSumF, step, ngood
  unsigned char* d_p_flag = nullptr;                           // EFPointStatus
  // residuals
  int cap_res = 0;
  int* d_r_point = nullptr; int* d_r_target = nullptr; int* d_r_host = nullptr;
  unsigned char* d_r_state = nullptr; unsigned char* d_r_newstate = nullptr; unsigned char* d_r_flags = nullptr;
  double* d_r_energy = nullptr;                                // [R][3]: state_energy, state_NewEnergy, state_NewEnergyWithOutlier
  float* d_r_J = nullptr; float* d_r_efJ = nullptr;            // [R][74]
  float* d_r_res_toZero = nullptr; float* d_r_JpJdF = nullptr; // [R][8]
  float* d_r_center = nullptr;                                 // [R][3]
  float* d_r_rsum = nullptr;                                   // [R][8]: JI_r[2], Jab_r[2], rr, (mode-tag as float), pad
  int* d_ht_order = nullptr; int* d_ht_begin = nullptr;        // residual ids stably sorted by (h + t*n); [n*n+1]
  std::vector<int> h_ht_begin;
  // accumulators / system
  float* d_blocks = nullptr;       // per (h,t): 13x13 top block (169) + E (32) + EB (8) + D (n*64), x segments
  int block_stride = 0, segments = 1;
  double* d_sys = nullptr;         // H_top, b_top (A and L), H_sc, b_sc, HFinal, bFinal, x, HM, bM, N(7) ...
  size_t sys_doubles = 0;
  double* d_energy = nullptr;      // scalar outputs
  int dim() const { return kCPARS + 8 * n; }
};

}  // namespace sdso
