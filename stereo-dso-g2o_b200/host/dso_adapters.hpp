// dso_adapters.hpp — the reference's own class names and member signatures on top of the C ABI (layer 2).
//
// A FullSystem-style caller keeps writing
//     fh->makeImages(color, &HCalib);
//     coarseTracker->makeK(&HCalib);  coarseTracker->setCoarseTrackingRef(frameHessians, fh_right, HCalib);
//     coarseTracker->trackNewestCoarse(fh, lastToNew, aff_g2l, pyrLevelsUsed - 1, minResForAbort);
//     ph->traceOn(frame, KRKi, Kt, aff, &HCalib);  ph->traceStereo(frame, K, mode_right);
//     ef->insertFrame / insertPoint / insertResidual / makeIDX / solveSystemF / marginalizePointsF / marginalizeFrame ...
// and every body below forwards to include/sdso_b200.h (hand-written sm_100a kernels behind it; no CPU fallback).
//
// The reference's linear-algebra types (Sophus::SE3d, AffLight, Eigen Vec5 / Mat33f / Vec3f / Vec2f, util/NumType.h:28-175) are
// template parameters, bundled in a Types struct, so this header compiles without Eigen or Sophus: PlainTypes below are
// std::array-based stand-ins (used by tests/native/adapter_test.cpp); INTEGRATION.md shows the EigenTypes a maintainer of the
// reference passes instead. Everything is header-only and inline; link with -lsdso_b200.
//
// Reference declarations mirrored here:
//   FrameHessian::makeImages                       FullSystem/HessianBlocks.h:237
//   CoarseTracker::{makeK, setCoarseTrackingRef, trackNewestCoarse, lastResiduals, lastFlowIndicators}   FullSystem/CoarseTracker.h:59-95
//   ImmaturePoint::{ImmaturePoint, traceOn, traceStereo}                                                FullSystem/ImmaturePoint.h:59-114
//   PointFrameResidual::{linearize, applyRes, resetOOB}                                                 FullSystem/Residuals.h:49-130
//   EnergyFunctional public API                                                                         OptimizationBackend/EnergyFunctional.h:65-110
//   AccumulatedTopHessianSSE / AccumulatedSCHessianSSE::{setZero, addPoint, stitchDouble[MT]}            OptimizationBackend/Accumulated{Top,SC}Hessian.h
//   VertexSE3PoseDSO ... VertexCamDSO::oplusImpl, EdgeTracePointUVDSO::{computeError, linearizeOplus}   FullSystem/dso_g2o_{vertex,edge}.h
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/sdso_b200.h"

namespace dso_b200 {

// ---- stand-in linear-algebra types -------------------------------------------------------------------------------------
struct PlainSE3 { double m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}; };   // row-major 3x4 [R|t]
struct PlainAffLight { double a = 0, b = 0; };
struct PlainTypes {
  using SE3 = PlainSE3;
  using AffLight = PlainAffLight;
  using Vec5 = std::array<double, 5>;
  using Vec3 = std::array<double, 3>;
  using Mat33f = std::array<float, 9>;   // row-major
  using Vec3f = std::array<float, 3>;
  using Vec2f = std::array<float, 2>;
  static void to_m34(const SE3& T, double out[12]) { std::memcpy(out, T.m, sizeof(T.m)); }
  static SE3 from_m34(const double in[12]) { SE3 T; std::memcpy(T.m, in, sizeof(T.m)); return T; }
  static void to_ab(const AffLight& g, double ab[2]) { ab[0] = g.a; ab[1] = g.b; }
  static AffLight from_ab(const double ab[2]) { AffLight g; g.a = ab[0]; g.b = ab[1]; return g; }
  static double at(const Vec5& v, int i) { return v[i]; }
  static void mat33f(const Mat33f& M, float out[9]) { for (int i = 0; i < 9; i++) out[i] = M[i]; }
  static void vec3f(const Vec3f& v, float out[3]) { for (int i = 0; i < 3; i++) out[i] = v[i]; }
  static void vec2f(const Vec2f& v, float out[2]) { out[0] = v[0]; out[1] = v[1]; }
};

struct Error : std::runtime_error { int code; Error(int c, const std::string& w) : std::runtime_error(w), code(c) {} };

// ---- the globals of util/globalCalib.cpp + util/settings.cpp become one object -------------------------------------------
class Context {
 public:
  Context(int w, int h, float fx, float fy, float cx, float cy, float baseline, const sdso_settings* S = nullptr, int device = 0) {
    const float K[4] = {fx, fy, cx, cy};
    const int rc = sdso_ctx_create(&ctx_, device, w, h, K, baseline, S);
    if (rc != SDSO_OK) { ctx_ = nullptr; throw Error(rc, "sdso_ctx_create failed (no CUDA device? there is no CPU fallback)"); }
  }
  ~Context() { if (ctx_) sdso_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  sdso_ctx* get() const { return ctx_; }
  int pyrLevelsUsed() const { return sdso_pyr_levels(ctx_); }
  void check(int rc, const char* what) const { if (rc != SDSO_OK) throw Error(rc, std::string(what) + ": " + sdso_last_error(ctx_)); }
 private:
  sdso_ctx* ctx_ = nullptr;
};

enum class ResState { IN = 0, OOB, OUTLIER };                                                       // Residuals.h:49
enum ImmaturePointStatus { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };  // ImmaturePoint.h:50-56
enum EFPointStatus { PS_GOOD = 0, PS_MARGINALIZE, PS_DROP };                                         // EnergyFunctionalStructs.h:97

// CalibHessian (HessianBlocks.h:272-371): the four scaled intrinsics the path reads + value_minus_value_zero
struct CalibHessian {
  float fxl_ = 0, fyl_ = 0, cxl_ = 0, cyl_ = 0;
  double value_minus_value_zero[4] = {0, 0, 0, 0};
  float fxl() const { return fxl_; }
  float fyl() const { return fyl_; }
  float cxl() const { return cxl_; }
  float cyl() const { return cyl_; }
};

struct PointFrameResidualBase;
struct PointHessianBase;

// FrameHessian (HessianBlocks.h:99-270): the image part lives on the device (dIp[] / absSquaredGrad[] fused into float4 texels)
template <class T = PlainTypes>
struct FrameHessian {
  Context* gpu = nullptr;
  int gpu_id = -1;                 // device pyramid slot
  float ab_exposure = 1.0f;
  int frameID = -1, idx = -1;      // idx = EFFrame::idx after makeIDX
  typename T::SE3 worldToCam_evalPT;         // get_worldToCam_evalPT()
  typename T::AffLight aff_g2l_;
  double state[10] = {0};          // get_state()
  float frameEnergyTH = 8 * 8 * SDSO_PATTERN_NUM;
  bool flaggedForMarginalization = false;
  std::vector<PointHessianBase*> pointHessians;
  explicit FrameHessian(Context* g) : gpu(g) { gpu->check(sdso_frame_create(gpu->get(), &gpu_id), "sdso_frame_create"); }
  ~FrameHessian() { if (gpu && gpu_id >= 0) sdso_frame_release(gpu->get(), gpu_id); }
  FrameHessian(const FrameHessian&) = delete;
  const typename T::AffLight& aff_g2l() const { return aff_g2l_; }
  // HessianBlocks.cpp:141-203
  void makeImages(float* color, CalibHessian* HCalib) { gpu->check(sdso_make_images(gpu->get(), gpu_id, color, ab_exposure, HCalib != nullptr), "sdso_make_images"); }
};

// PointFrameResidual (Residuals.h:56-130): the fields the path reads / writes; J lives on the device (74 SoA planes)
struct PointFrameResidualBase {
  PointHessianBase* point = nullptr;
  int host = -1, target = -1;      // frame idx
  int slot = -1;                   // residual index in the uploaded window
  ResState state_state = ResState::IN, state_NewState = ResState::OUTLIER;
  double state_energy = 0, state_NewEnergy = 0, state_NewEnergyWithOutlier = 0;
  float centerProjectedTo[3] = {0, 0, 0};
  bool isActiveAndIsGoodNEW = false, isLinearized = false;
  void resetOOB() { state_NewEnergy = state_energy = 0; state_NewState = ResState::OUTLIER; state_state = ResState::IN; }   // Residuals.h:106-111
};

// PointHessian + EFPoint (HessianBlocks.h:378-460, EnergyFunctionalStructs.h:99-150)
struct PointHessianBase {
  int host = -1;
  float u = 0, v = 0, idepth_scaled = 0, idepth_zero_scaled = 0;
  float color[8] = {0}, weights[8] = {0};
  bool hasDepthPrior = false;
  float HdiF = 0, step = 0, idepth_hessian = 0;
  EFPointStatus stateFlag = PS_GOOD;
  std::vector<PointFrameResidualBase*> residuals;
  std::pair<PointFrameResidualBase*, ResState> lastResiduals[2] = {{nullptr, ResState::OOB}, {nullptr, ResState::OOB}};
  int slot = -1;
};

// ---- ImmaturePoint (ImmaturePoint.h:59-114) ----------------------------------------------------------------------------
template <class T = PlainTypes>
class ImmaturePoint {
 public:
  sdso_immature_point rec;         // same field names as the reference's members (u, v, idepth_min, ..., lastTraceUV, quality)
  FrameHessian<T>* host;
  float my_type = 1;
  ImmaturePointStatus lastTraceStatus = IPS_UNINITIALIZED;
  bool ok = false;                 // false: the constructor bailed out on a non-finite colour (energyTH = NaN, ImmaturePoint.cpp:49)
  // ImmaturePoint(int u_, int v_, FrameHessian* host_, float type, CalibHessian* HCalib)  (ImmaturePoint.cpp:33-62)
  ImmaturePoint(int u_, int v_, FrameHessian<T>* host_, float type, CalibHessian* HCalib) : host(host_), my_type(type) { init((float)u_, (float)v_); (void)HCalib; }
  // ImmaturePoint(float u_, float v_, FrameHessian* host_, CalibHessian* HCalib)           (:64-88)
  ImmaturePoint(float u_, float v_, FrameHessian<T>* host_, CalibHessian* HCalib) : host(host_) { init(u_, v_); (void)HCalib; }
  // :94-451
  ImmaturePointStatus traceStereo(FrameHessian<T>* frame, typename T::Mat33f K, bool mode_right) {
    float Kf[9]; T::mat33f(K, Kf);
    int st = 0;
    host->gpu->check(sdso_trace_stereo(host->gpu->get(), frame->gpu_id, Kf, mode_right ? 1 : 0, 1, &rec, &st), "sdso_trace_stereo");
    return lastTraceStatus = (ImmaturePointStatus)st;
  }
  // :459-828
  ImmaturePointStatus traceOn(FrameHessian<T>* frame, typename T::Mat33f hostToFrame_KRKi, typename T::Vec3f hostToFrame_Kt,
                              typename T::Vec2f hostToFrame_affine, CalibHessian* HCalib, bool debugPrint = false) {
    (void)HCalib; (void)debugPrint;
    float Kf[9], tf[3], af[2]; T::mat33f(hostToFrame_KRKi, Kf); T::vec3f(hostToFrame_Kt, tf); T::vec2f(hostToFrame_affine, af);
    int st = 0;
    host->gpu->check(sdso_trace_on(host->gpu->get(), frame->gpu_id, Kf, tf, af, 1, &rec, &st), "sdso_trace_on");
    return lastTraceStatus = (ImmaturePointStatus)st;
  }
  // The loops of FullSystem::traceNewCoarseKey / traceNewCoarseNonKey (FullSystem.cpp:632-781) over one host frame: ONE launch
  static void traceOnAll(std::vector<ImmaturePoint*>& pts, FrameHessian<T>* frame, typename T::Mat33f KRKi, typename T::Vec3f Kt, typename T::Vec2f aff) {
    if (pts.empty()) return;
    float Kf[9], tf[3], af[2]; T::mat33f(KRKi, Kf); T::vec3f(Kt, tf); T::vec2f(aff, af);
    std::vector<sdso_immature_point> recs(pts.size()); std::vector<int> st(pts.size());
    for (size_t i = 0; i < pts.size(); i++) recs[i] = pts[i]->rec;
    Context* g = pts[0]->host->gpu;
    g->check(sdso_trace_on(g->get(), frame->gpu_id, Kf, tf, af, (int)pts.size(), recs.data(), st.data()), "sdso_trace_on");
    for (size_t i = 0; i < pts.size(); i++) { pts[i]->rec = recs[i]; pts[i]->lastTraceStatus = (ImmaturePointStatus)st[i]; }
  }
 private:
  void init(float u, float v) {
    const float uv[2] = {u, v};
    int okf = 0;
    host->gpu->check(sdso_immature_init(host->gpu->get(), host->gpu_id, 1, uv, &rec, &okf), "sdso_immature_init");
    ok = okf != 0;
  }
};

// ---- CoarseTracker (CoarseTracker.h:55-140) -------------------------------------------------------------------------------
template <class T = PlainTypes>
class CoarseTracker {
 public:
  typename T::Vec5 lastResiduals{};
  typename T::Vec3 lastFlowIndicators{};
  int refFrameID = -1;
  FrameHessian<T>* lastRef = nullptr;
  typename T::AffLight lastRef_aff_g2l;
  int variant = SDSO_VARIANT_G2O;  // the fork's live code path; SDSO_VARIANT_SSE = the original DSO arithmetic (its commented-out blocks)
  int lastIterations[5] = {0, 0, 0, 0, 0};

  CoarseTracker(Context* gpu, int w, int h) : gpu_(gpu) { (void)w; (void)h; }   // the reference's (w, h) come from the context

  // CoarseTracker.cpp:108-136
  void makeK(CalibHessian* HCalib) {
    const float K[4] = {HCalib->fxl(), HCalib->fyl(), HCalib->cxl(), HCalib->cyl()};
    gpu_->check(sdso_tracker_make_k(gpu_->get(), K), "sdso_tracker_make_k");
  }

  // CoarseTracker.cpp:809-825 -> makeCoarseDepthL0 (:275-534). STEP1's per-point loop (:290-356) runs here as three batched device
  // calls instead of 2-3 heap ImmaturePoints per point: constructor at the rounded projection, traceStereo into the right image,
  // constructor + traceStereo back; then the splat / pool / dilate / compact steps (one call).
  void setCoarseTrackingRef(std::vector<FrameHessian<T>*> frameHessians, FrameHessian<T>* fh_right, CalibHessian Hcalib) {
    if (frameHessians.empty()) throw Error(SDSO_E_INVALID, "setCoarseTrackingRef: no frames");
    lastRef = frameHessians.back();
    sdso_ctx* c = gpu_->get();
    std::vector<float> uv, id0, wgt;   // rounded projection, centerProjectedTo[2], inverse-covariance weight
    for (FrameHessian<T>* fh : frameHessians)
      for (PointHessianBase* ph : fh->pointHessians) {
        if (ph->lastResiduals[0].first == nullptr || ph->lastResiduals[0].second != ResState::IN) continue;   // :295
        const PointFrameResidualBase* r = ph->lastResiduals[0].first;
        const int u = (int)(r->centerProjectedTo[0] + 0.5f), v = (int)(r->centerProjectedTo[1] + 0.5f);        // :302-303
        uv.push_back((float)u); uv.push_back((float)v);
        id0.push_back(r->centerProjectedTo[2]);
        wgt.push_back(sqrtf(1e-3 / (ph->HdiF + 1e-12)));                                                       // :350
      }
    const int n = (int)id0.size();
    std::vector<float> splat((size_t)n * 4);
    for (int i = 0; i < n; i++) { splat[4 * i] = uv[2 * i]; splat[4 * i + 1] = uv[2 * i + 1]; splat[4 * i + 2] = id0[i]; splat[4 * i + 3] = wgt[i]; }
    if (fh_right != nullptr && n > 0) {   // two-way static stereo re-check (:305-347)
      const float K1[9] = {Hcalib.fxl(), 0, Hcalib.cxl(), 0, Hcalib.fyl(), Hcalib.cyl(), 0, 0, 1};
      std::vector<sdso_immature_point> fwd(n);
      std::vector<int> st(n);
      gpu_->check(sdso_immature_init(c, lastRef->gpu_id, n, uv.data(), fwd.data(), nullptr), "sdso_immature_init");
      for (int i = 0; i < n; i++) { fwd[i].idepth_min_stereo = id0[i] * 0.1f; fwd[i].idepth_max_stereo = id0[i] * 1.9f; }   // :311-312
      gpu_->check(sdso_trace_stereo(c, fh_right->gpu_id, K1, 1, n, fwd.data(), st.data()), "sdso_trace_stereo");
      std::vector<int> good;
      std::vector<float> uvb;
      for (int i = 0; i < n; i++) if (st[i] == IPS_GOOD) { good.push_back(i); uvb.push_back(fwd[i].lastTraceUV[0]); uvb.push_back(fwd[i].lastTraceUV[1]); }
      if (!good.empty()) {
        const int m = (int)good.size();
        std::vector<sdso_immature_point> back(m);
        std::vector<int> stb(m);
        gpu_->check(sdso_immature_init(c, fh_right->gpu_id, m, uvb.data(), back.data(), nullptr), "sdso_immature_init");
        for (int k = 0; k < m; k++) { back[k].idepth_min_stereo = id0[good[k]] * 0.1f; back[k].idepth_max_stereo = id0[good[k]] * 1.9f; }   // :324-325
        gpu_->check(sdso_trace_stereo(c, lastRef->gpu_id, K1, 0, m, back.data(), stb.data()), "sdso_trace_stereo");
        for (int k = 0; k < m; k++) {
          const sdso_immature_point& p = fwd[good[k]];
          const float depth = 1.0f / p.idepth_stereo;
          const float u_delta = std::fabs(p.u - back[k].lastTraceUV[0]);
          if (u_delta < 1 && depth > 0 && depth < 50) splat[4 * good[k] + 2] = p.idepth_stereo;   // :332-335
        }
      }
    }
    double ab[2]; T::to_ab(lastRef->aff_g2l(), ab);
    gpu_->check(sdso_tracker_set_ref(c, lastRef->gpu_id, splat.data(), n, ab), "sdso_tracker_set_ref");
    refFrameID = lastRef->frameID;
    lastRef_aff_g2l = lastRef->aff_g2l();
  }

  // CoarseTracker.cpp:827-1069
  bool trackNewestCoarse(FrameHessian<T>* newFrameHessian, typename T::SE3& lastToNew_out, typename T::AffLight& aff_g2l_out, int coarsestLvl,
                         typename T::Vec5 minResForAbort, void* wrap = nullptr) {
    (void)wrap;
    double Tm[12], ab[2], mr[5], res[5], flow[3];
    T::to_m34(lastToNew_out, Tm); T::to_ab(aff_g2l_out, ab);
    for (int i = 0; i < 5; i++) mr[i] = T::at(minResForAbort, i);
    int ok = 0;
    gpu_->check(sdso_track(gpu_->get(), newFrameHessian->gpu_id, Tm, ab, coarsestLvl, mr, variant, res, flow, lastIterations, &ok), "sdso_track");
    for (int i = 0; i < 5; i++) lastResiduals[i] = res[i];
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = flow[i];
    lastToNew_out = T::from_m34(Tm);
    aff_g2l_out = T::from_ab(ab);
    return ok != 0;
  }
 private:
  Context* gpu_;
};

// ---- EnergyFunctional + the accumulators + PointFrameResidual::linearize (window level) -----------------------------------------
// The reference's pointer graph (EFFrame / EFPoint / EFResidual, EnergyFunctional.cpp:443-552) is mirrored as index arrays and
// uploaded by makeIDX(); the operators then run on the device for the whole window. PointFrameResidual::linearize is a per-object
// call in the reference (Residuals.h:103) but FullSystem only ever calls it from linearizeAll over all active residuals
// (FullSystemOptimize.cpp:142-203), so the adapter exposes linearizeAll() and writes each residual's state_New* / centerProjectedTo
// back into its object, where the reference's callers read them.
template <class T = PlainTypes>
class EnergyFunctional {
 public:
  std::vector<FrameHessian<T>*> frames;
  int nPoints = 0, nFrames = 0, nResiduals = 0;
  explicit EnergyFunctional(Context* gpu) : gpu_(gpu) {}

  void insertFrame(FrameHessian<T>* fh, CalibHessian* Hcalib) { fh->idx = (int)frames.size(); frames.push_back(fh); calib_ = *Hcalib; dirty_ = true; }   // :478-518
  void insertPoint(PointHessianBase* ph) { (void)ph; dirty_ = true; }                                     // :521-531 (points are read from fh->pointHessians)
  void insertResidual(PointFrameResidualBase* r) { (void)r; dirty_ = true; }                              // :443-476 (residuals are read from ph->residuals)
  // :998-1025 — (re)index frames, allPoints order = by frame, then pointHessians order; upload the window
  void makeIDX() {
    sdso_ctx* c = gpu_->get();
    gpu_->check(sdso_ba_reset(c), "sdso_ba_reset");
    const float K[4] = {calib_.fxl(), calib_.fyl(), calib_.cxl(), calib_.cyl()};
    gpu_->check(sdso_ba_set_calib(c, K, calib_.value_minus_value_zero), "sdso_ba_set_calib");
    nFrames = (int)frames.size();
    for (int i = 0; i < nFrames; i++) {
      FrameHessian<T>* fh = frames[i];
      double Tm[12], ab[2]; T::to_m34(fh->worldToCam_evalPT, Tm); T::to_ab(fh->aff_g2l(), ab);
      int idx = -1;
      gpu_->check(sdso_ba_add_frame(c, fh->gpu_id, Tm, ab[0], ab[1], fh->frameID, &idx), "sdso_ba_add_frame");
      fh->idx = idx;
      gpu_->check(sdso_ba_set_state(c, idx, fh->state), "sdso_ba_set_state");
      gpu_->check(sdso_ba_set_energy_th(c, idx, fh->frameEnergyTH), "sdso_ba_set_energy_th");
    }
    points_.clear(); residuals_.clear();
    std::vector<int> host, rp, rt; std::vector<float> u, v, id, idz, col, wts; std::vector<unsigned char> prior, flags;
    for (FrameHessian<T>* fh : frames)
      for (PointHessianBase* ph : fh->pointHessians) {
        ph->slot = (int)points_.size(); ph->host = fh->idx; points_.push_back(ph);
        host.push_back(fh->idx); u.push_back(ph->u); v.push_back(ph->v); id.push_back(ph->idepth_scaled); idz.push_back(ph->idepth_zero_scaled);
        col.insert(col.end(), ph->color, ph->color + 8); wts.insert(wts.end(), ph->weights, ph->weights + 8);
        prior.push_back(ph->hasDepthPrior ? 1 : 0); flags.push_back((unsigned char)ph->stateFlag);
        for (PointFrameResidualBase* r : ph->residuals) { r->slot = (int)residuals_.size(); r->host = fh->idx; residuals_.push_back(r); rp.push_back(ph->slot); rt.push_back(r->target); }
      }
    nPoints = (int)points_.size(); nResiduals = (int)residuals_.size();
    gpu_->check(sdso_ba_set_points(c, nPoints, host.data(), u.data(), v.data(), id.data(), idz.data(), col.data(), wts.data(), prior.data()), "sdso_ba_set_points");
    gpu_->check(sdso_ba_set_residuals(c, nResiduals, rp.data(), rt.data()), "sdso_ba_set_residuals");
    gpu_->check(sdso_ba_set_point_flags(c, flags.data()), "sdso_ba_set_point_flags");
    gpu_->check(sdso_ba_prepare(c), "sdso_ba_prepare");   // setPrecalcValues + setAdjointsF + setDeltaF
    dirty_ = false;
  }
  void setAdjointsF(CalibHessian* Hcalib) { (void)Hcalib; sync(); }   // :41-119 (part of sdso_ba_prepare)
  void setDeltaF(CalibHessian* HCalib) { (void)HCalib; sync(); }      // :173-207

  // FullSystem::linearizeAll(fixLinearization) (FullSystemOptimize.cpp:142-203) = PointFrameResidual::linearize over the active set
  double linearizeAll(bool fixLinearization) {
    sync();
    double e = 0;
    gpu_->check(sdso_ba_linearize_all(gpu_->get(), fixLinearization ? 1 : 0, &e), "sdso_ba_linearize_all");
    readResiduals(fixLinearization ? 1 : 0);
    return e;
  }
  // AccumulatedTopHessianSSE::{setZero, addPoint<mode>, stitchDoubleMT} over the whole window (AccumulatedTopHessian.cpp:36-337)
  void accumulateTop(int mode, bool usePrior, std::vector<double>& H, std::vector<double>& b) {
    sync(); const int d = dim(); H.assign((size_t)d * d, 0.0); b.assign(d, 0.0);
    gpu_->check(sdso_ba_accumulate_top(gpu_->get(), mode, usePrior ? 1 : 0, H.data(), b.data(), nullptr), "sdso_ba_accumulate_top");
  }
  // AccumulatedSCHessianSSE::{setZero, addPoint, stitchDoubleMT} (AccumulatedSCHessian.cpp:34-256)
  void accumulateSC(bool shiftPriorToZero, std::vector<double>& H, std::vector<double>& b) {
    sync(); const int d = dim(); H.assign((size_t)d * d, 0.0); b.assign(d, 0.0);
    gpu_->check(sdso_ba_accumulate_sc(gpu_->get(), shiftPriorToZero ? 1 : 0, H.data(), b.data()), "sdso_ba_accumulate_sc");
  }
  // :838-995 incl. resubstituteF_MT; lastX = the solved increment, frame / calib steps and point steps land in the objects
  void solveSystemF(int iteration, double lambda, CalibHessian* HCalib) {
    (void)HCalib; sync();
    lastX.assign(dim(), 0.0);
    gpu_->check(sdso_ba_solve(gpu_->get(), iteration, lambda, lastX.data(), nullptr, nullptr), "sdso_ba_solve");
    frameSteps.assign((size_t)nFrames * 10, 0.0);
    gpu_->check(sdso_ba_resubstitute(gpu_->get(), nullptr, frameSteps.data(), calibStep), "sdso_ba_resubstitute");
    std::vector<float> pts((size_t)nPoints * 16);
    gpu_->check(sdso_ba_get_points(gpu_->get(), pts.data()), "sdso_ba_get_points");
    for (int i = 0; i < nPoints; i++) { points_[i]->HdiF = pts[(size_t)i * 16 + 12]; points_[i]->step = pts[(size_t)i * 16 + 14]; }
  }
  void marginalizePointsF() {   // :663-736: points with stateFlag == PS_MARGINALIZE
    sync();
    std::vector<unsigned char> flags(nPoints);
    for (int i = 0; i < nPoints; i++) flags[i] = (unsigned char)points_[i]->stateFlag;
    gpu_->check(sdso_ba_set_point_flags(gpu_->get(), flags.data()), "sdso_ba_set_point_flags");
    gpu_->check(sdso_ba_marginalize_points(gpu_->get()), "sdso_ba_marginalize_points");
  }
  void marginalizeFrame(FrameHessian<T>* fh) {   // :554-660
    sync();
    gpu_->check(sdso_ba_marginalize_frame(gpu_->get(), fh->idx), "sdso_ba_marginalize_frame");
    frames.erase(frames.begin() + fh->idx);
    for (size_t i = 0; i < frames.size(); i++) frames[i]->idx = (int)i;
    dirty_ = true;   // the reference calls makeIDX() here as well
  }
  double calcMEnergyF() { sync(); double m = 0; gpu_->check(sdso_ba_energies(gpu_->get(), &m, nullptr), "sdso_ba_energies"); return m; }      // :344-351
  double calcLEnergyF_MT() { sync(); double l = 0; gpu_->check(sdso_ba_energies(gpu_->get(), nullptr, &l), "sdso_ba_energies"); return l; }   // :354-442
  // FullSystem::optimize, SSE body (FullSystemOptimize.cpp:870-1042)
  double optimize(int mnumOptIts, int* iterations_done = nullptr) {
    sync();
    double rmse = 0; int its = 0;
    gpu_->check(sdso_ba_optimize(gpu_->get(), mnumOptIts, &rmse, &its), "sdso_ba_optimize");
    if (iterations_done) *iterations_done = its;
    std::vector<double> st((size_t)nFrames * 10); std::vector<float> idp(nPoints);
    gpu_->check(sdso_ba_get_state(gpu_->get(), st.data(), nullptr, idp.data(), nullptr), "sdso_ba_get_state");
    for (int i = 0; i < nFrames; i++) std::memcpy(frames[i]->state, &st[(size_t)i * 10], sizeof(double) * 10);
    for (int i = 0; i < nPoints; i++) points_[i]->idepth_scaled = idp[i];
    readResiduals(1);
    return rmse;
  }
  int dim() const { return 4 + 8 * nFrames; }
  std::vector<double> lastX, frameSteps;
  double calibStep[4] = {0, 0, 0, 0};
 private:
  void sync() { if (dirty_) makeIDX(); }
  void readResiduals(int which) {
    const int R = nResiduals;
    if (R == 0) return;
    std::vector<int> ns(R), st(R), ac(R), li(R); std::vector<double> ne(R), nw(R); std::vector<float> ce((size_t)R * 3);
    gpu_->check(sdso_ba_get_res(gpu_->get(), which, ns.data(), st.data(), ne.data(), nw.data(), ac.data(), li.data(), nullptr, nullptr, ce.data(), nullptr), "sdso_ba_get_res");
    for (int i = 0; i < R; i++) {
      PointFrameResidualBase* r = residuals_[i];
      r->state_NewState = (ResState)ns[i]; r->state_state = (ResState)st[i]; r->state_NewEnergy = ne[i]; r->state_NewEnergyWithOutlier = nw[i];
      r->isActiveAndIsGoodNEW = ac[i] != 0; r->isLinearized = li[i] != 0;
      for (int k = 0; k < 3; k++) r->centerProjectedTo[k] = ce[(size_t)i * 3 + k];
    }
  }
  Context* gpu_;
  CalibHessian calib_;
  std::vector<PointHessianBase*> points_;
  std::vector<PointFrameResidualBase*> residuals_;
  bool dirty_ = true;
};

// ---- g2o vertices and the trace edge (dso_g2o_vertex.h:24-115, dso_g2o_edge.h:173-205) ---------------------------------------
// estimate() / setEstimate() / oplusImpl(const double*) as g2o::BaseVertex exposes them; one object = a batch of one. For whole
// batches call sdso_vertex_oplus / sdso_edge_trace_uv_eval directly (SoA, one launch).
template <int D, int KIND, int W>
class VertexDSO {
 public:
  explicit VertexDSO(Context* gpu) : gpu_(gpu) { for (int i = 0; i < W; i++) est_[i] = 0; if (KIND == SDSO_VERTEX_SE3_POSE) est_[0] = est_[5] = est_[10] = 1; }
  const double* estimate() const { return est_; }
  void setEstimate(const double* e) { for (int i = 0; i < W; i++) est_[i] = e[i]; }
  void setToOriginImpl() { *this = VertexDSO(gpu_); }
  void SetDxDy(double dx, double dy) { aux_[0] = dx; aux_[1] = dy; }   // VertexUVDSO only
  void oplusImpl(const double* update_) { gpu_->check(sdso_vertex_oplus(gpu_->get(), KIND, 1, est_, update_, aux_), "sdso_vertex_oplus"); }
  static constexpr int Dimension = D;
 private:
  Context* gpu_;
  double est_[W];
  double aux_[2] = {0, 0};
};
using VertexSE3PoseDSO = VertexDSO<6, SDSO_VERTEX_SE3_POSE, 12>;
using VertexPhotometricDSO = VertexDSO<2, SDSO_VERTEX_PHOTOMETRIC, 2>;
using VertexInverseDepthDSO = VertexDSO<1, SDSO_VERTEX_INVERSE_DEPTH, 1>;
using VertexUVDSO = VertexDSO<1, SDSO_VERTEX_UV, 2>;
using VertexCamDSO = VertexDSO<4, SDSO_VERTEX_CAM, 4>;

template <class T = PlainTypes>
class EdgeTracePointUVDSO {
 public:
  EdgeTracePointUVDSO(Context* gpu, FrameHessian<T>* frame, typename T::Vec2f affLL, double dx, double dy, typename T::Vec2f rotatePattern)
      : gpu_(gpu), frame_(frame) { T::vec2f(affLL, aff_); T::vec2f(rotatePattern, rot_); dxdy_[0] = dx; dxdy_[1] = dy; }
  void setMeasurement(double m) { meas_ = m; }
  void setVertex(const VertexUVDSO* v) { v_ = v; }
  double error() const { return err_; }
  double jacobianOplusXi() const { return J_; }
  void computeError() { eval(); }     // (the device call evaluates both; the members not touched by the reference stay as they were)
  void linearizeOplus() { eval(); }
 private:
  void eval() { gpu_->check(sdso_edge_trace_uv_eval(gpu_->get(), frame_->gpu_id, 1, v_->estimate(), rot_, &meas_, aff_, dxdy_, &err_, &J_, nullptr), "sdso_edge_trace_uv_eval"); }
  Context* gpu_; FrameHessian<T>* frame_; const VertexUVDSO* v_ = nullptr;
  float aff_[2], rot_[2]; double dxdy_[2], meas_ = 0, err_ = 0, J_ = 0;
};

}  // namespace dso_b200
