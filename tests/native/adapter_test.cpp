// C++ caller written against the ADAPTER header (the reference's class names and member signatures) and linked with
// libsdso_b200.so: FrameHessian::makeImages, CoarseTracker::{makeK, setCoarseTrackingRef (incl. the two-way static-stereo
// re-check of makeCoarseDepthL0), trackNewestCoarse}, ImmaturePoint::{ctor, traceStereo}, VertexSE3PoseDSO::oplusImpl,
// EnergyFunctional::{insertFrame, makeIDX, linearizeAll, solveSystemF}. Compiles without Eigen (PlainTypes).
//
// usage: adapter_test <input.bin> <variant> <use_right>
//   input.bin: int32 w, h, n; float K[4], baseline; float img0[w*h], img1[w*h], imgR[w*h]; float pts[n*4] = {u, v, idepth, HdiF};
//              double T0[12]
//   prints "T ...", "aff ...", "res ...", "ok ...", "stereo <u v status idepth_stereo>", "vertex <12>", "ba <energy> <|x|>" (%.17g)
//   or "nodevice" (exit 3) when no CUDA device is usable.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "../../stereo-dso-g2o_b200/host/dso_adapters.hpp"

using namespace dso_b200;
using Types = PlainTypes;

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s input.bin variant use_right\n", argv[0]); return 2; }
  const int variant = atoi(argv[2]), use_right = atoi(argv[3]);
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int32_t hdr[3]; float K[4], baseline;
  if (fread(hdr, 4, 3, f) != 3 || fread(K, 4, 4, f) != 4 || fread(&baseline, 4, 1, f) != 1) return 2;
  const size_t npx = (size_t)hdr[0] * hdr[1];
  const int n = hdr[2];
  std::vector<float> img0(npx), img1(npx), imgR(npx), pts((size_t)n * 4);
  double T0[12];
  if (fread(img0.data(), 4, npx, f) != npx || fread(img1.data(), 4, npx, f) != npx || fread(imgR.data(), 4, npx, f) != npx ||
      fread(pts.data(), 4, (size_t)n * 4, f) != (size_t)n * 4 || fread(T0, 8, 12, f) != 12) return 2;
  fclose(f);
  std::unique_ptr<Context> gpu;
  try {
    gpu.reset(new Context(hdr[0], hdr[1], K[0], K[1], K[2], K[3], baseline));
  } catch (const Error& e) {
    if (e.code == SDSO_E_NODEVICE || e.code == SDSO_E_CUDA) { printf("nodevice\n"); return 3; }
    fprintf(stderr, "%s\n", e.what()); return 1;
  }
  try {
    CalibHessian HCalib; HCalib.fxl_ = K[0]; HCalib.fyl_ = K[1]; HCalib.cxl_ = K[2]; HCalib.cyl_ = K[3];
    FrameHessian<Types> fh0(gpu.get()), fh1(gpu.get()), fhR(gpu.get());
    fh0.frameID = 0; fh1.frameID = 1;
    fh0.makeImages(img0.data(), &HCalib); fh1.makeImages(img1.data(), &HCalib); fhR.makeImages(imgR.data(), &HCalib);
    // the window's points as the reference holds them: PointHessian with its last residual IN, projected into the newest key frame
    std::vector<PointHessianBase> ph(n); std::vector<PointFrameResidualBase> res(n);
    for (int i = 0; i < n; i++) {
      res[i].centerProjectedTo[0] = pts[4 * i]; res[i].centerProjectedTo[1] = pts[4 * i + 1]; res[i].centerProjectedTo[2] = pts[4 * i + 2];
      ph[i].HdiF = pts[4 * i + 3];
      ph[i].lastResiduals[0] = {&res[i], ResState::IN};
      fh0.pointHessians.push_back(&ph[i]);
    }
    CoarseTracker<Types> tracker(gpu.get(), hdr[0], hdr[1]);
    tracker.variant = variant;
    tracker.makeK(&HCalib);
    tracker.setCoarseTrackingRef({&fh0}, use_right ? &fhR : nullptr, HCalib);
    Types::SE3 lastToNew = Types::from_m34(T0);
    Types::AffLight aff;
    Types::Vec5 minRes; for (auto& v : minRes) v = NAN;
    const bool ok = tracker.trackNewestCoarse(&fh1, lastToNew, aff, gpu->pyrLevelsUsed() - 1, minRes);
    printf("T"); for (int i = 0; i < 12; i++) printf(" %.17g", lastToNew.m[i]);
    printf("\naff %.17g %.17g\nres", aff.a, aff.b);
    for (int i = 0; i < 5; i++) printf(" %.17g", tracker.lastResiduals[i]);
    printf("\nok %d\n", ok ? 1 : 0);
    // one ImmaturePoint through its constructor + traceStereo, as makeCoarseDepthL0 does per point
    ImmaturePoint<Types> ip(pts[0], pts[1], &fh0, &HCalib);
    ip.rec.idepth_min_stereo = pts[2] * 0.1f; ip.rec.idepth_max_stereo = pts[2] * 1.9f;
    const ImmaturePointStatus st = ip.traceStereo(&fhR, Types::Mat33f{K[0], 0, K[2], 0, K[1], K[3], 0, 0, 1}, true);
    printf("stereo %.9g %.9g %d %.9g\n", ip.rec.lastTraceUV[0], ip.rec.lastTraceUV[1], (int)st, ip.rec.idepth_stereo);
    // VertexSE3PoseDSO::oplusImpl
    VertexSE3PoseDSO vp(gpu.get());
    vp.setEstimate(T0);
    const double upd[6] = {0.01, -0.02, 0.03, 0.004, -0.005, 0.006};
    vp.oplusImpl(upd);
    printf("vertex"); for (int i = 0; i < 12; i++) printf(" %.17g", vp.estimate()[i]);
    printf("\n");
    // a two-frame window through EnergyFunctional: points hosted in fh0 observed in fh1
    EnergyFunctional<Types> ef(gpu.get());
    double Tw1[12]; for (int i = 0; i < 12; i++) Tw1[i] = lastToNew.m[i];
    fh0.worldToCam_evalPT = Types::SE3(); fh1.worldToCam_evalPT = Types::from_m34(Tw1);
    fh0.pointHessians.clear();
    const int nb = n < 400 ? n : 400;
    std::vector<PointHessianBase> bp(nb); std::vector<PointFrameResidualBase> br(nb);
    for (int i = 0; i < nb; i++) {
      ImmaturePoint<Types> q(pts[4 * i], pts[4 * i + 1], &fh0, &HCalib);
      bp[i].u = pts[4 * i]; bp[i].v = pts[4 * i + 1]; bp[i].idepth_scaled = bp[i].idepth_zero_scaled = pts[4 * i + 2];
      for (int k = 0; k < 8; k++) { bp[i].color[k] = q.rec.color[k]; bp[i].weights[k] = q.rec.weights[k]; }
      br[i].point = &bp[i]; br[i].target = 1; br[i].resetOOB();
      bp[i].residuals.push_back(&br[i]);
      fh0.pointHessians.push_back(&bp[i]);
    }
    ef.insertFrame(&fh0, &HCalib); ef.insertFrame(&fh1, &HCalib);
    ef.makeIDX();
    const double energy = ef.linearizeAll(true);
    ef.solveSystemF(0, 1e-5, &HCalib);
    double xn = 0; for (double v : ef.lastX) xn += v * v;
    int in = 0; for (int i = 0; i < nb; i++) in += (br[i].state_state == ResState::IN);
    printf("ba %.17g %.17g %d\n", energy, std::sqrt(xn), in);
  } catch (const Error& e) {
    fprintf(stderr, "adapter_test: %s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
