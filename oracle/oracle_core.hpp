// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// State + operators of the photometric hot path, restated on the CPU.
#pragma once
#include "oracle_common.hpp"

namespace orc {

// 3x3 float inverse the way Eigen computes Matrix3f::inverse() (cofactors * 1/det).
void inverse3f(const float m[9], float out[9]);

// util/globalCalib.cpp:48-108 (globals wG,hG,fxG..,KG,KiG,wM3G,hM3G,baseline) — initial calibration
struct GlobalCalib {
  int levels = 0;
  int w[PYR_LEVELS], h[PYR_LEVELS];
  float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
  float fxi[PYR_LEVELS], fyi[PYR_LEVELS], cxi[PYR_LEVELS], cyi[PYR_LEVELS];
  float K[PYR_LEVELS][9], Ki[PYR_LEVELS][9];
  float wM3G, hM3G;
  float baseline = 0;
  void set(int w0, int h0, const float K0[9]);
};

// FullSystem/HessianBlocks.h:272-371 (only what the hot path reads)
struct CalibHessian {
  float fxl, fyl, cxl, cyl;      // value_scaledf
  float fxli, fyli, cxli, cyli;  // value_scaledi
  double value_minus_value_zero[4] = {0, 0, 0, 0};
  float B[256];
  void setValueScaledf(float fx, float fy, float cx, float cy);
  float getBGradOnly(float color) const;  // HessianBlocks.h:356-362
};

// FullSystem/HessianBlocks.h:99-270 (image part)
struct Frame {
  std::vector<float> dIp[PYR_LEVELS];         // AoS {I,dx,dy} per pixel
  std::vector<float> absSquaredGrad[PYR_LEVELS];
  float ab_exposure = 1.0f;
  float frameEnergyTH = 8 * 8 * patternNum;
  void makeImages(const GlobalCalib& G, const float* color, const CalibHessian* HCalib, const Settings& S);
};

// util/globalFuncs.h:73-86, 122-135, 160-184
void getInterpolatedElement33(const float* mat3, float x, float y, int width, float out[3]);
float getInterpolatedElement31(const float* mat3, float x, float y, int width);
void getInterpolatedElement33BiLin(const float* mat3, float x, float y, int width, float out[3]);

// ---- coarse tracker (FullSystem/CoarseTracker.{h,cpp}) -------------------------------------
struct RefPoint { float u, v, idepth, weight; };  // one splat of makeCoarseDepthL0 STEP1

struct CoarseTracker {
  const GlobalCalib* G = nullptr;
  Settings S;
  int w[PYR_LEVELS], h[PYR_LEVELS];
  float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
  float fxi[PYR_LEVELS], fyi[PYR_LEVELS], cxi[PYR_LEVELS], cyi[PYR_LEVELS];
  float K[PYR_LEVELS][9], Ki[PYR_LEVELS][9];
  std::vector<float> idepth[PYR_LEVELS], weightSums[PYR_LEVELS], weightSums_bak[PYR_LEVELS];
  std::vector<float> pc_u[PYR_LEVELS], pc_v[PYR_LEVELS], pc_idepth[PYR_LEVELS], pc_color[PYR_LEVELS];
  int pc_n[PYR_LEVELS];
  std::vector<float> buf_warped_idepth, buf_warped_u, buf_warped_v, buf_warped_dx, buf_warped_dy,
      buf_warped_residual, buf_warped_weight, buf_warped_refColor;
  int buf_warped_n = 0;
  const Frame* lastRef = nullptr;
  const Frame* newFrame = nullptr;
  double lastRef_aff_g2l[2] = {0, 0};
  double lastResiduals[5];
  double lastFlowIndicators[3];

  void init(const GlobalCalib* G_, const Settings& S_);
  void makeK(const CalibHessian& HCalib);                              // CoarseTracker.cpp:108-136
  // makeCoarseDepthL0 STEP1..5 with the per-point (u,v,idepth,weight) already decided
  // (CoarseTracker.cpp:350-533; the stereo re-check of :314-347 is the caller's business)
  void setRefFromSplats(const Frame* ref, const RefPoint* pts, int n, const double aff_g2l[2]);
  // ---- SSE path (original DSO; CoarseTracker.cpp:537-596 + commented blocks :699-775,:888-1023)
  void calcResSSE(int lvl, const SE3& refToNew, const double aff_g2l[2], float cutoffTH, double rs[6]);
  void calcGSSSE(int lvl, double H[64], double b[8], const SE3& refToNew, const double aff_g2l[2]);
  bool trackNewestCoarseSSE(const Frame* fh, SE3& lastToNew_out, double aff_g2l_out[2], int coarsestLvl,
                            const double minResForAbort[5], int* iterations_out /*[5] nullable*/);
  // ---- g2o path (live code; CoarseTracker.cpp:600-792, 827-1069 + dso_g2o_edge.cpp:395-500)
  struct Edge {  // EdgeSE3PosePhotoDSO
    float Xref[3]; int level; double measurement; double error; int edge_level;
  };
  std::vector<Edge> edges;  // optimizer->edges(), all levels
  void calcResG2O(int lvl, const SE3& refToNew, float cutoffTH, const SE3& vtx_pose, const double vtx_photo[2],
                  double rs[6]);
  bool trackNewestCoarseG2O(const Frame* fh, SE3& lastToNew_out, double aff_g2l_out[2], int coarsestLvl,
                            const double minResForAbort[5], int* lm_iterations_out /*[5] nullable*/);
  // E1 operator-level entry points (dso_g2o_edge.cpp:395-423, 425-500)
  void edgeComputeError(Edge& e, const SE3& pose, const double photo[2]) const;
  bool edgeLinearizeOplus(const Edge& e, const SE3& pose, const double photo[2], double Jpose[6], double Jphoto[2]) const;
  uint64_t evals = 0;  // number of project+gather+residual evaluations performed (for the metric)
  int g2o_trials = 0, g2o_rejected = 0;  // damping trials of the last trackNewestCoarseG2O and how many of them were popped (test coverage)
};

}  // namespace orc
