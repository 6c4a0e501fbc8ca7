// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// Calibration pyramid, image pyramid + gradients, bilinear gathers.
#include "oracle_core.hpp"

namespace orc {

void inverse3f(const float m[9], float out[9]) {
  // Eigen compute_inverse_size3: cofactor(i,j) with cyclic indices, result(j,i) = cof(i,j) * (1/det),
  // det = sum_i cof(i,0) * m(i,0)
  auto cof = [&](int i, int j) -> float {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
  };
  float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  float det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  float invdet = 1.0f / det;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[j * 3 + i] = cof(i, j) * invdet;
}

// util/globalCalib.cpp:48-108
void GlobalCalib::set(int w0, int h0, const float K0[9]) {
  int wlvl = w0, hlvl = h0;
  levels = 1;
  while (wlvl % 2 == 0 && hlvl % 2 == 0 && wlvl * hlvl > 5000 && levels < PYR_LEVELS) { wlvl /= 2; hlvl /= 2; levels++; }
  wM3G = w0 - 3; hM3G = h0 - 3;
  w[0] = w0; h[0] = h0;
  for (int i = 0; i < 9; i++) K[0][i] = K0[i];
  fx[0] = K0[0]; fy[0] = K0[4]; cx[0] = K0[2]; cy[0] = K0[5];
  inverse3f(K[0], Ki[0]);
  fxi[0] = Ki[0][0]; fyi[0] = Ki[0][4]; cxi[0] = Ki[0][2]; cyi[0] = Ki[0][5];
  for (int l = 1; l < levels; l++) {
    w[l] = w0 >> l; h[l] = h0 >> l;
    fx[l] = fx[l - 1] * 0.5;  // float * double -> double -> float (exact: power of two)
    fy[l] = fy[l - 1] * 0.5;
    cx[l] = (cx[0] + 0.5) / ((int)1 << l) - 0.5;
    cy[l] = (cy[0] + 0.5) / ((int)1 << l) - 0.5;
    float Kl[9] = {fx[l], 0, cx[l], 0, fy[l], cy[l], 0, 0, 1};
    for (int i = 0; i < 9; i++) K[l][i] = Kl[i];
    inverse3f(K[l], Ki[l]);
    fxi[l] = Ki[l][0]; fyi[l] = Ki[l][4]; cxi[l] = Ki[l][2]; cyi[l] = Ki[l][5];
  }
}

// FullSystem/HessianBlocks.h:332-347 (setValueScaled, float part)
void CalibHessian::setValueScaledf(float fx, float fy, float cx, float cy) {
  fxl = fx; fyl = fy; cxl = cx; cyl = cy;
  fxli = 1.0f / fxl; fyli = 1.0f / fyl; cxli = -cxl / fxl; cyli = -cyl / fyl;
  for (int i = 0; i < 256; i++) B[i] = i;
}
float CalibHessian::getBGradOnly(float color) const {
  int c = color + 0.5f;
  if (c < 5) c = 5;
  if (c > 250) c = 250;
  return B[c + 1] - B[c];
}

// FullSystem/HessianBlocks.cpp:141-203
void Frame::makeImages(const GlobalCalib& G, const float* color, const CalibHessian* HCalib, const Settings& S) {
  for (int i = 0; i < G.levels; i++) {
    // the reference leaves border rows of dx/dy/absSquaredGrad uninitialised (new[] without init);
    // the oracle zero-fills so comparisons are defined. Tests compare rows 1..h-2 only.
    dIp[i].assign((size_t)G.w[i] * G.h[i] * 3, 0.0f);
    absSquaredGrad[i].assign((size_t)G.w[i] * G.h[i], 0.0f);
  }
  int w = G.w[0], h = G.h[0];
  float* dI = dIp[0].data();
  for (int i = 0; i < w * h; i++) dI[3 * i] = color[i];
  for (int lvl = 0; lvl < G.levels; lvl++) {
    int wl = G.w[lvl], hl = G.h[lvl];
    float* dI_l = dIp[lvl].data();
    float* dabs_l = absSquaredGrad[lvl].data();
    if (lvl > 0) {
      int wlm1 = G.w[lvl - 1];
      const float* dI_lm = dIp[lvl - 1].data();
      for (int y = 0; y < hl; y++)
        for (int x = 0; x < wl; x++) {
          dI_l[3 * (x + y * wl)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                            dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] +
                                            dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
        }
    }
    for (int idx = wl; idx < wl * (hl - 1); idx++) {
      float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
      float dy = 0.5f * (dI_l[3 * (idx + wl)] - dI_l[3 * (idx - wl)]);
      if (!std::isfinite(dx)) dx = 0;
      if (!std::isfinite(dy)) dy = 0;
      dI_l[3 * idx + 1] = dx;
      dI_l[3 * idx + 2] = dy;
      dabs_l[idx] = dx * dx + dy * dy;
      if (S.gammaWeightsPixelSelect == 1 && HCalib != 0) {
        float gw = HCalib->getBGradOnly((float)(dI_l[3 * idx]));
        dabs_l[idx] *= gw * gw;
      }
    }
  }
}

// util/globalFuncs.h:73-86
void getInterpolatedElement33(const float* mat, float x, float y, int width, float out[3]) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int c = 0; c < 3; c++)
    out[c] = w11 * bp[3 * (1 + width) + c] + w01 * bp[3 * width + c] + w10 * bp[3 + c] + w00 * bp[c];
}
// util/globalFuncs.h:122-135
float getInterpolatedElement31(const float* mat, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}
// util/globalFuncs.h:160-184
void getInterpolatedElement33BiLin(const float* mat, float x, float y, int width, float out[3]) {
  if (x == -1 || y == -1) { out[0] = out[1] = out[2] = 0; return; }  // reference returns an uninitialised vector
  int ix = (int)x, iy = (int)y;
  const float* bp = mat + 3 * (ix + iy * width);
  float tl = bp[0], tr = bp[3], bl = bp[3 * width], br = bp[3 * (width + 1)];
  float dx = x - ix, dy = y - iy;
  float topInt = dx * tr + (1 - dx) * tl;
  float botInt = dx * br + (1 - dx) * bl;
  float leftInt = dy * bl + (1 - dy) * tl;
  float rightInt = dy * br + (1 - dy) * tr;
  out[0] = dx * rightInt + (1 - dx) * leftInt;
  out[1] = rightInt - leftInt;
  out[2] = botInt - topInt;
}

}  // namespace orc
