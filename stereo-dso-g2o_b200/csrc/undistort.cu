// Input preparation on the device: PhotometricUndistorter::processFrame (util/Undistort.cpp:222-260) fused with the bilinear
// remap of Undistort::undistort (:398-489, benchmark noise off as in the reference's defaults) — one thread per output pixel
// reads its four raw 8-bit taps, applies the response LUT / vignette (or the plain factor) to each tap and interpolates in
// the reference's term order (un-fused multiplies and adds), so the rectified float image is bit-identical. The remap
// tables are inputs (the caller's Undistort object computes them once from the calibration file, :556-714).
// Also the row format of FullSystem::printResult (FullSystem.cpp:236-285), host only.
#include "ctx.h"
#include <cstdio>
#include <cstring>
#include <string>

namespace sdso {

struct UndistortState {
  int wOrg = 0, hOrg = 0;
  float *d_remapX = nullptr, *d_remapY = nullptr, *d_G = nullptr, *d_vig = nullptr, *d_out = nullptr;
  unsigned char* d_raw = nullptr;
  int photometricCalibration = 2, useExposure = 1;
  bool has_G = false, has_vig = false, ready = false;
};

// mode 0: factor * raw; 1: G[raw]; 2: G[raw] * vignetteMapInv
template <int MODE>
__device__ __forceinline__ float photo_tap(const unsigned char* __restrict__ raw, const float* __restrict__ G, const float* __restrict__ vig, float factor, int i) {
  const unsigned char r = raw[i];
  if (MODE == 0) return factor * r;
  if (MODE == 1) return G[r];
  return G[r] * vig[i];
}

template <int MODE>
__global__ void __launch_bounds__(256) undistort_kernel(const unsigned char* __restrict__ raw, const float* __restrict__ remapX,
                                                        const float* __restrict__ remapY, const float* __restrict__ G, const float* __restrict__ vig,
                                                        float factor, int wOrg, int n, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float xx = remapX[idx], yy = remapY[idx];
  float v = 0.f;
  if (!(xx < 0)) {
    const int xxi = (int)xx, yyi = (int)yy;
    xx -= xxi; yy -= yyi;
    const float xxyy = xx * yy;
    const int b = xxi + yyi * wOrg;
    const float s00 = photo_tap<MODE>(raw, G, vig, factor, b), s10 = photo_tap<MODE>(raw, G, vig, factor, b + 1);
    const float s01 = photo_tap<MODE>(raw, G, vig, factor, b + wOrg), s11 = photo_tap<MODE>(raw, G, vig, factor, b + 1 + wOrg);
    v = xxyy * s11 + (yy - xxyy) * s01 + (xx - xxyy) * s10 + (1 - xx - yy + xxyy) * s00;
  }
  out[idx] = v;
}

int undistort_create(sdso_ctx* ctx) { ctx->undistort = new UndistortState(); return SDSO_OK; }
void undistort_destroy(sdso_ctx* ctx) {
  UndistortState* s = ctx->undistort;
  if (!s) return;
  void* ptrs[] = {s->d_remapX, s->d_remapY, s->d_G, s->d_vig, s->d_out, s->d_raw};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete s;
  ctx->undistort = nullptr;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_undistort_setup(sdso_ctx* ctx, int wOrg, int hOrg, const float* remapX, const float* remapY, const float* G, const float* vignetteMapInv,
                         int photometricCalibration, int useExposure) {
  sdso::enter(ctx);
  if (!ctx || !ctx->undistort || wOrg < 2 || hOrg < 2 || !remapX || !remapY) return SDSO_E_INVALID;
  if (photometricCalibration < 0 || photometricCalibration > 2) return SDSO_E_INVALID;
  if (photometricCalibration == 2 && G && !vignetteMapInv) return fail(ctx, SDSO_E_INVALID, "undistort: photometricCalibration 2 needs the inverse vignette");
  UndistortState* s = ctx->undistort;
  const size_t n = (size_t)ctx->G.w[0] * ctx->G.h[0], nOrg = (size_t)wOrg * hOrg;
  // every non-negative remap entry must leave room for the 2x2 taps, as the reference's tables do (Undistort.cpp:690-712)
  for (size_t i = 0; i < n; i++)
    if (!(remapX[i] < 0) && !(remapX[i] >= 0 && remapY[i] >= 0 && (int)remapX[i] + 1 < wOrg && (int)remapY[i] + 1 < hOrg))
      return fail(ctx, SDSO_E_INVALID, "undistort: remap entry outside the raw image");
  void* old[] = {s->d_remapX, s->d_remapY, s->d_G, s->d_vig, s->d_out, s->d_raw};
  for (void* p : old) if (p) cudaFree(p);
  s->d_remapX = s->d_remapY = s->d_G = s->d_vig = s->d_out = nullptr; s->d_raw = nullptr; s->ready = false;
  SDSO_CUDA(ctx, cudaMalloc(&s->d_remapX, n * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_remapY, n * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_out, n * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_raw, nOrg));
  SDSO_CUDA(ctx, cudaMemcpy(s->d_remapX, remapX, n * sizeof(float), cudaMemcpyHostToDevice));
  SDSO_CUDA(ctx, cudaMemcpy(s->d_remapY, remapY, n * sizeof(float), cudaMemcpyHostToDevice));
  s->has_G = G != nullptr; s->has_vig = vignetteMapInv != nullptr;
  if (G) { SDSO_CUDA(ctx, cudaMalloc(&s->d_G, 256 * sizeof(float))); SDSO_CUDA(ctx, cudaMemcpy(s->d_G, G, 256 * sizeof(float), cudaMemcpyHostToDevice)); }
  if (vignetteMapInv) {
    SDSO_CUDA(ctx, cudaMalloc(&s->d_vig, nOrg * sizeof(float)));
    SDSO_CUDA(ctx, cudaMemcpy(s->d_vig, vignetteMapInv, nOrg * sizeof(float), cudaMemcpyHostToDevice));
  }
  s->wOrg = wOrg; s->hOrg = hOrg; s->photometricCalibration = photometricCalibration; s->useExposure = useExposure;
  s->ready = true;
  return SDSO_OK;
}

int sdso_undistort(sdso_ctx* ctx, const unsigned char* raw, float exposure, float factor, float* out_image, int frame, int use_hcalib, float* exposure_out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->undistort || !raw) return SDSO_E_INVALID;
  UndistortState* s = ctx->undistort;
  if (!s->ready) return fail(ctx, SDSO_E_STATE, "undistort before undistort_setup");
  if (frame >= 0 && (frame >= (int)ctx->frames.size() || !ctx->frames[frame].in_use)) return SDSO_E_INVALID;
  const int n = ctx->G.w[0] * ctx->G.h[0];
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->d_raw, raw, (size_t)s->wOrg * s->hOrg, cudaMemcpyHostToDevice, st));
  // processFrame's branch (:231-252): no response calibration, exposure <= 0 or setting_photometricCalibration == 0 -> factor * raw
  const int mode = (!s->has_G || exposure <= 0 || s->photometricCalibration == 0) ? 0 : (s->photometricCalibration == 2 ? 2 : 1);
  const int grid = (n + 255) / 256;
  if (mode == 0) undistort_kernel<0><<<grid, 256, 0, st>>>(s->d_raw, s->d_remapX, s->d_remapY, s->d_G, s->d_vig, factor, s->wOrg, n, s->d_out);
  else if (mode == 1) undistort_kernel<1><<<grid, 256, 0, st>>>(s->d_raw, s->d_remapX, s->d_remapY, s->d_G, s->d_vig, factor, s->wOrg, n, s->d_out);
  else undistort_kernel<2><<<grid, 256, 0, st>>>(s->d_raw, s->d_remapX, s->d_remapY, s->d_G, s->d_vig, factor, s->wOrg, n, s->d_out);
  SDSO_CHECK_LAUNCH(ctx);
  const float e = s->useExposure ? exposure : 1.f;
  if (exposure_out) *exposure_out = e;
  if (frame >= 0) {
    int rc = sdso_make_images_device(ctx, frame, s->d_out, e, use_hcalib);
    if (rc) return rc;
  }
  if (out_image) {
    SDSO_CUDA(ctx, cudaMemcpyAsync(out_image, s->d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return SDSO_OK;
}

int sdso_trajectory_row(const double camToWorld[12], char* buf, int n) {
  if (!camToWorld || !buf || n <= 0) return SDSO_E_INVALID;
  std::string s;
  char tmp[64];
  for (int i = 0; i < 12; i++) {
    snprintf(tmp, sizeof(tmp), "%.15g", camToWorld[i]);   // operator<< with setprecision(15) and the default float field
    s += tmp;
    s += (i == 11 ? "\n" : " ");
  }
  if ((int)s.size() + 1 > n) return SDSO_E_INVALID;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // extern "C"
