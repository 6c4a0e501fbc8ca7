// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the photometric hot path of gyubeomim/stereo-dso-g2o.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (stereo-dso-g2o_b200/) never links or calls it.
//
// PARITY STATUS: the reference ships no tests, golden vectors or fixtures for this path
//  and cannot be compiled here (Eigen, g2o, Boost, OpenCV absent), so the
// restatement is pinned only by (a) the Sophus test point sets for SE3 exp/log/Adj
// (thirdparty/Sophus/sophus/test_se3.cpp), (b) analytic identities between the SSE and g2o
// variants and finite differences, (c) committed fixtures under tests/golden/.
// The g2o LM/GN driver is "parity unpinned" (g2o is not vendored, no version pinned).
//
// Shared constants, tiny fixed-size linear algebra, SE3.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace orc {

// ---- constants mirrored from the reference ------------------------------------------------
// util/settings.h:46,177-179 ; util/settings.cpp:216 (pattern 8 "for SSE efficiency")
constexpr int PYR_LEVELS = 6;
constexpr int patternNum = 8;
constexpr int patternPadding = 2;
static const int patternP[8][2] = {{0,-2},{-1,-1},{1,-1},{-2,0},{0,0},{2,0},{-1,1},{0,2}};
// FullSystem/HessianBlocks.h:54-61
constexpr float SCALE_IDEPTH = 1.0f, SCALE_XI_ROT = 1.0f, SCALE_XI_TRANS = 0.5f, SCALE_F = 50.0f,
                SCALE_C = 50.0f, SCALE_W = 1.0f, SCALE_A = 10.0f, SCALE_B = 1000.0f;
constexpr int CPARS = 4;             // util/NumType.h:47
constexpr int NUM_THREADS = 6;       // util/NumType.h:38
constexpr int MAX_RES_PER_POINT = 8; // util/NumType.h:37

// util/settings.cpp defaults + main_dso_pangolin.cpp preset=0 mode=1
struct Settings {
  float huberTH = 9;                   // settings.cpp:95
  float coarseCutoffTH = 20;           // :102
  float outlierTH = 12 * 12;           // :72
  float outlierTHSumComponent = 50*50; // :73
  float overallEnergyTHWeight = 1;     // :101
  float maxPixSearch = 0.027f;         // :111
  int minTraceTestRadius = 2;          // :113
  float trace_stepsize = 1.0f;         // :115
  int trace_GNIterations = 3;          // :116
  float trace_GNThreshold = 0.1f;      // :117
  float trace_extraSlackOnTH = 1.2f;   // :118
  float trace_slackInterval = 1.5f;    // :119
  float trace_minImprovementFactor = 2;// :120
  float affineOptModeA = 0;            // main_dso_pangolin.cpp:326 (mode=1)
  float affineOptModeB = 0;            // :327
  int gammaWeightsPixelSelect = 1;     // settings.cpp:93
  float idepthFixPrior = 50 * 50;      // :42
  float idepthFixPriorMargFac = 600*600; // :43
  float initialRotPrior = 1e11f, initialTransPrior = 1e10f, initialAffBPrior = 1e14f,
        initialAffAPrior = 1e14f, initialCalibHessian = 5e9f; // :44-48
  float margWeightFac = 0.5f * 0.5f;   // :76
  double solverModeDelta = 0.00001;    // :52
  int maxOptIterations = 6, minOptIterations = 1;  // :67-68
  float thOptIterations = 1.2f;        // :69
  float minIdepthH_act = 100;          // :56
  int GNItsOnPointActivation = 3;      // :114
  float minTraceQuality = 3;           // :112
  float frameEnergyTHConstWeight = 0.5f, frameEnergyTHN = 0.7f, frameEnergyTHFacMedian = 1.5f;  // :98-100
};

// ---- tiny dense helpers (row-major, runtime n) ---------------------------------------------
// Used for Hl.ldlt().solve (CoarseTracker.cpp:934) and HFinalScaled.ldlt().solve
// (EnergyFunctional.cpp:976).
inline bool ldlt_solve(int n, const double* A, const double* b, double* x) {
  // symmetric (diagonal-)pivoted LDLT, the strategy of Eigen::LDLT: P A P^T = L D L^T
  std::vector<double> M(A, A + n * n), d(n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k; double best = std::fabs(M[k * n + k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(M[i * n + i]) > best) { best = std::fabs(M[i * n + i]); p = i; }
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(M[k * n + j], M[p * n + j]);
      for (int j = 0; j < n; j++) std::swap(M[j * n + k], M[j * n + p]);
      std::swap(perm[k], perm[p]);
    }
    double dk = M[k * n + k];
    d[k] = dk;
    if (dk == 0.0 || !std::isfinite(dk)) { for (int i = k + 1; i < n; i++) M[i * n + k] = 0; continue; }
    for (int i = k + 1; i < n; i++) M[i * n + k] /= dk;          // L(i,k)
    for (int i = k + 1; i < n; i++) {
      double lik = M[i * n + k];
      for (int j = k + 1; j <= i; j++) M[i * n + j] -= lik * dk * M[j * n + k];
    }
    for (int i = k + 1; i < n; i++) for (int j = i + 1; j < n; j++) M[i * n + j] = M[j * n + i];
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) y[i] -= M[i * n + j] * y[j];
  for (int i = 0; i < n; i++) y[i] = (d[i] != 0.0) ? y[i] / d[i] : 0.0;
  for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) y[i] -= M[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
  bool ok = true;
  for (int i = 0; i < n; i++) ok = ok && std::isfinite(x[i]);
  return ok;
}

// Gauss-Jordan inverse with partial pivoting (8x8 hpi.inverse(), EnergyFunctional.cpp:614).
inline bool mat_inverse(int n, const double* A, double* Ainv) {
  std::vector<double> M(n * 2 * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { M[i * 2 * n + j] = A[i * n + j]; M[i * 2 * n + n + j] = (i == j); }
  for (int k = 0; k < n; k++) {
    int p = k; for (int i = k + 1; i < n; i++) if (std::fabs(M[i * 2 * n + k]) > std::fabs(M[p * 2 * n + k])) p = i;
    if (M[p * 2 * n + k] == 0.0) return false;
    if (p != k) for (int j = 0; j < 2 * n; j++) std::swap(M[k * 2 * n + j], M[p * 2 * n + j]);
    double inv = 1.0 / M[k * 2 * n + k];
    for (int j = 0; j < 2 * n; j++) M[k * 2 * n + j] *= inv;
    for (int i = 0; i < n; i++) if (i != k) { double f = M[i * 2 * n + k]; if (f != 0) for (int j = 0; j < 2 * n; j++) M[i * 2 * n + j] -= f * M[k * 2 * n + j]; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ainv[i * n + j] = M[i * 2 * n + n + j];
  return true;
}

// ---- SE3, following thirdparty/Sophus (quaternion + translation; tangent = [upsilon; omega]) -
struct Quat { double w = 1, x = 0, y = 0, z = 0; };
struct SE3 {
  Quat q; double t[3] = {0, 0, 0};
  void rotationMatrix(double R[9]) const;               // Eigen Quaternion::toRotationMatrix
  void toMat34(double M[12]) const;                     // row-major [R|t]
  static SE3 fromMat34(const double M[12]);             // Eigen Quaternion(Matrix3) (Shepperd)
  static SE3 exp(const double a[6]);                    // se3.hpp:407-428 / so3.hpp:343-369
  void log(double out[6]) const;                        // se3.hpp:560-586 / so3.hpp:491-531
  SE3 inverse() const;                                  // se3.hpp (q.conj, -(R^T t))
  SE3 operator*(const SE3& o) const;                    // se3.hpp operator*= (+ so3 renormalise)
  void Adj(double A[36]) const;                         // se3.hpp:131-139
  void act(const double p[3], double out[3]) const;     // R p + t
};

// util/NumType.h:159-170
inline void affFromToVecExposure(float exposureF, float exposureT, double aF, double bF, double aT, double bT,
                                 double out[2]) {
  if (exposureF == 0 || exposureT == 0) { exposureT = exposureF = 1; }
  double a = std::exp(aT - aF) * exposureT / exposureF;
  double b = bT - a * bF;
  out[0] = a; out[1] = b;
}

}  // namespace orc
