// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// Windowed bundle adjustment, SSE path of the reference: PointFrameResidual::linearize (Residuals.cpp:83-336),
// RawResidualJacobian, EFResidual::takeDataF / fixLinearizationF, AccumulatedTopHessianSSE,
// AccumulatedSCHessianSSE, EnergyFunctional::{setAdjointsF,setDeltaF,solveSystemF,resubstituteF,
// marginalizePointsF,marginalizeFrame,orthogonalize}, FrameFramePrecalc::set. Index-based restatement of
// the reference's pointer graph.
#pragma once
#include "oracle_core.hpp"

namespace orc {

enum ResState { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };  // Residuals.h:49

// OptimizationBackend/RawResidualJacobian.h:32-65 (74 floats)
struct RawResidualJacobian {
  float resF[8];
  float Jpdxi[2][6];
  float Jpdc[2][4];
  float Jpdd[2];
  float JIdx[2][8];
  float JabF[2][8];
  float JIdx2[4];    // 2x2 row-major
  float JabJIdx[4];  // 2x2 row-major
  float Jab2[4];     // 2x2 row-major
};

// FullSystem/HessianBlocks.h:72-97
struct FrameFramePrecalc {
  float PRE_RTll[9], PRE_KRKiTll[9], PRE_RKiTll[9], PRE_RTll_0[9];
  float PRE_aff_mode[2];
  float PRE_b0_mode;
  float PRE_tTll[3], PRE_KtTll[3], PRE_tTll_0[3];
  float distanceLL;
};

struct BAFrame {  // FrameHessian + EFFrame (only what the backend reads)
  const Frame* img = nullptr;
  int frameID = 0;
  SE3 worldToCam_evalPT;
  double state[10] = {0}, state_zero[10] = {0}, state_scaled[10] = {0};
  SE3 PRE_worldToCam, PRE_camToWorld;
  double nullspaces_pose[36];  // 6x6 row-major, column i = i-th nullspace
  double nullspaces_affine[8]; // 4x2
  double nullspaces_scale[6];
  double prior[8] = {0}, delta_prior[8] = {0}, delta[8] = {0};  // EFFrame
  double step[10] = {0}, state_backup[10] = {0};
  float frameEnergyTH = 8 * 8 * patternNum;
  void setState(const double s[10]);                               // HessianBlocks.h:177-199
  void setStateScaled(const double s[10]);                         // :201-215
  void setStateZero(const double s[10]);                           // HessianBlocks.cpp:78-123
  void setEvalPT_scaled(const SE3& w2c, double a, double b);       // HessianBlocks.h:223-231
  void aff_g2l(double ab[2]) const { ab[0] = state_scaled[6]; ab[1] = state_scaled[7]; }
  void aff_g2l_0(double ab[2]) const { ab[0] = state_zero[6] * SCALE_A; ab[1] = state_zero[7] * SCALE_B; }
};

struct BAPoint {  // PointHessian + EFPoint
  int host = 0;
  float u = 0, v = 0;
  float idepth = 0, idepth_scaled = 0, idepth_zero = 0, idepth_zero_scaled = 0;
  float color[8], weights[8];
  bool hasDepthPrior = false;
  float priorF = 0, deltaF = 0;
  float bdSumF = 0, HdiF = 0, Hdd_accLF = 0, bd_accLF = 0, Hdd_accAF = 0, bd_accAF = 0;
  float Hcd_accLF[4] = {0, 0, 0, 0}, Hcd_accAF[4] = {0, 0, 0, 0};
  float idepth_hessian = 0, step = 0, idepth_backup = 0;
  int stateFlag = 0;  // EFPointStatus: 0 GOOD, 1 MARGINALIZE, 2 DROP
  std::vector<int> residuals;  // indices into BAWindow::res (residualsAll order)
};

struct BARes {  // PointFrameResidual + EFResidual
  int point = 0, host = 0, target = 0;
  int state_state = RS_IN, state_NewState = RS_OUTLIER;
  double state_energy = 0, state_NewEnergy = 0, state_NewEnergyWithOutlier = 0;
  RawResidualJacobian J;    // PointFrameResidual::J  (candidate linearisation)
  RawResidualJacobian efJ;  // EFResidual::J          (accepted linearisation, after takeDataF)
  float res_toZeroF[8] = {0};
  float JpJdF[8] = {0};
  bool isLinearized = false, isActiveAndIsGoodNEW = false;
  float centerProjectedTo[3] = {0, 0, 0};
  float projectedTo[8][2];
  bool isActive() const { return isActiveAndIsGoodNEW; }
};

struct BAWindow {
  const GlobalCalib* G = nullptr;
  Settings S;
  CalibHessian HCalib;
  std::vector<BAFrame> frames;
  std::vector<BAPoint> points;   // "allPoints" order: by host frame, then insertion (EnergyFunctional.cpp:1003-1016)
  std::vector<BARes> res;
  std::vector<FrameFramePrecalc> precalc;  // [host * n + target]
  // EnergyFunctional state
  std::vector<double> adHost, adTarget;    // n*n 8x8 row-major, index h + t*n
  std::vector<float> adHostF, adTargetF;
  std::vector<float> adHTdeltaF;           // n*n x 8
  double cPrior[4];
  float cPriorF[4], cDeltaF[4];
  std::vector<double> HM, bM;              // (4+8n)^2, 4+8n
  std::vector<std::vector<double>> lastNullspaces_pose, lastNullspaces_scale;
  // Model of the reference's multi-threaded accumulation (IndexThreadReduce.h:69-123 driven by EnergyFunctional.cpp:214-257:
  // chunks of 50 points of allPoints go to whichever of the NUM_THREADS=6 workers asks next, each worker owns a float accumulator set,
  // the stitch sums the sets in double, AccumulatedTopHessian.cpp:299-308 / AccumulatedSCHessian.cpp:140-146). reduce_threads = 1 is
  // the reference's single-threaded path (tid = -1); reduce_seed picks the chunk -> worker assignment (0: round robin) — the
  // reference's own assignment is a race, so ANY seed is a result the reference can produce. Used by the tests that measure the
  // oracle's own run-to-run spread (tests/test_oracle_spread.py).
  int reduce_threads = 1;
  unsigned reduce_seed = 0;
  int reduce_tid(int point_index) const {
    if (reduce_threads <= 1) return 0;
    unsigned c = (unsigned)(point_index / 50);
    if (reduce_seed == 0) return (int)(c % (unsigned)reduce_threads);
    unsigned x = c * 2654435761u ^ (reduce_seed * 0x9E3779B9u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return (int)(x % (unsigned)reduce_threads);
  }
  int n() const { return (int)frames.size(); }
  int dim() const { return CPARS + 8 * n(); }

  void setPrecalcValues();                               // FullSystem.cpp:1633-1644 -> FrameFramePrecalc::set
  void setAdjointsF();                                   // EnergyFunctional.cpp:41-119
  void setDeltaF();                                      // :173-207
  void getNullspaces();                                  // FullSystemOptimize.cpp:1087-1147
  double linearize(BARes& r);                            // Residuals.cpp:83-336
  void applyRes(BARes& r, bool copyJacobians);           // Residuals.cpp:367-385 (+ takeDataF)
  void fixLinearizationF(BARes& r);                      // EnergyFunctionalStructs.cpp:96-123
  double linearizeAll(bool fixLinearization);            // FullSystemOptimize.cpp:142-203 (energy sum; applyRes when fix)
  // accumulate + stitch (single accumulator set, i.e. the reference's non-MT path)
  void accumulateTop(int mode, std::vector<double>& H, std::vector<double>& b, bool usePrior);  // B4 + B5
  void accumulateSC(bool shiftPriorToZero, std::vector<double>& H, std::vector<double>& b);      // B6 + B7
  // per (h,t) 13x13 blocks of the last accumulateTop call (row-major, index h + t*n), for operator-level parity
  std::vector<float> lastTopBlocks;
  void solveSystemF(int iteration, double lambda, std::vector<double>& x, std::vector<double>* Hfinal, std::vector<double>* bfinal);  // B8
  void resubstituteF(const std::vector<double>& x, double frame_steps[/*n*10*/], double calib_step[4]);  // B9
  void orthogonalize(std::vector<double>* b, std::vector<double>* H);                            // :775-835
  void marginalizePointsF();                                                                   // :663-736
  void marginalizeFrame(int idx);                                                              // :554-660
  double calcMEnergyF();                                                                       // :344-351
  double calcLEnergyF();                                                                       // :354-442
  // FullSystem LM driver of the SSE path (FullSystemOptimize.cpp:870-1042, 207-370, 98-139)
  double calib_value[4] = {0, 0, 0, 0}, calib_value_zero[4] = {0, 0, 0, 0}, calib_step[4] = {0, 0, 0, 0}, calib_backup[4] = {0, 0, 0, 0};
  bool calib_init = false;
  int resInA = 0;
  void initCalibValue();
  void setCalibValue(const double v[4]);                 // CalibHessian::setValue (HessianBlocks.h:316-331)
  void backupState();
  bool doStepFromBackup(float stepfacC, float stepfacT, float stepfacR, float stepfacA, float stepfacD);
  float newFrameEnergyTH();                              // setNewFrameEnergyTH value for the newest frame (:98-139)
  double optimize(int mnumOptIts, int* iterations_done);
};

// g2o LBA edge (dso_g2o_edge.cpp:5-282), operator level
struct LBAEdgeOut {
  double error[8];
  double J_xi[8][6], J_photo[8][2], J_idepth[8], J_C[8][4];
  int newState; double newEnergy, newEnergyWithOutlier; float centerProjectedTo[3]; float idepth_hessian; int level; int center_set;
};
void lbaEdgeEval(const BAWindow& W, const BARes& r, const SE3& T_wh /*vertex pose*/, const double photo[2], double idepth,
                 const double cam[4], double b0, LBAEdgeOut& out);

// FullSystem::optimize, g2o body (FullSystemOptimize.cpp:404-868) with the restated g2o LM (oracle/lba_g2o.cpp)
int lbaG2O(BAWindow& W, int mnumOptIts, double cam[4], double* T_wh, double* photo, double* idepth_io, int* used_host, double* chi2_out,
           int* newState_out, float* center_out, float* idepth_hessian_out, int* trials_out);

}  // namespace orc
