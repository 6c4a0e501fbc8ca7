// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// PixelSelector (FullSystem/PixelSelector2.cpp): makeHists (:84-178), select (:340-536), makeMaps (:192-327). Sequential
// restatement: the direction index of every pot-cell is randomPattern[n2] & 0xF with n2 the RUNNING count of level-0 selections,
// and the sub-sampling walks randomPattern[rn] with rn the running count of selected pixels. randomPattern is glibc's
// srand(3141592); rand() & 0xFF (:43-45) — the oracle calls glibc itself, which also pins the product's own generator.
#pragma once
#include "oracle_core.hpp"
#include <cstdlib>
#include <cstring>

namespace orc {

struct PixelSelector {
  int w = 0, h = 0, currentPotential = 3, thsStep = 0;
  std::vector<unsigned char> randomPattern;
  std::vector<float> ths, thsSmoothed;
  const Frame* gradHistFrame = nullptr;
  float minGradHistCut = 0.5f, minGradHistAdd = 7, gradDownweightPerLevel = 0.75f;  // settings.cpp:105-107
  void init(int w_, int h_) {
    w = w_; h = h_;
    randomPattern.resize((size_t)w * h);
    std::srand(3141592);
    for (int i = 0; i < w * h; i++) randomPattern[i] = rand() & 0xFF;
    currentPotential = 3;
    ths.assign((w / 32) * (h / 32) + 100, 0.f); thsSmoothed.assign((w / 32) * (h / 32) + 100, 0.f);
    gradHistFrame = nullptr;
  }
  static int computeHistQuantil(const int* hist, float below) {
    int th = hist[0] * below + 0.5f;
    for (int i = 0; i < 90; i++) { th -= hist[i + 1]; if (th < 0) return i; }
    return 90;
  }
  void makeHists(const Frame& fh) {
    gradHistFrame = &fh;
    const float* mapmax0 = fh.absSquaredGrad[0].data();
    const int w32 = w / 32, h32 = h / 32;
    thsStep = w32;
    int hist0[100];
    for (int y = 0; y < h32; y++) for (int x = 0; x < w32; x++) {
      const float* map0 = mapmax0 + 32 * x + 32 * y * w;
      // (the reference clears 50 ints but the quantile walks 91 entries of a 100-int scratch that is otherwise never written)
      memset(hist0, 0, sizeof(hist0));
      for (int j = 0; j < 32; j++) for (int i = 0; i < 32; i++) {
        const int it = i + 32 * x, jt = j + 32 * y;
        if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
        int g = sqrtf(map0[i + j * w]);
        if (g > 48) g = 48;
        hist0[g + 1]++; hist0[0]++;
      }
      ths[x + y * w32] = computeHistQuantil(hist0, minGradHistCut) + minGradHistAdd;
    }
    for (int y = 0; y < h32; y++) for (int x = 0; x < w32; x++) {
      float sum = 0, num = 0;
      if (x > 0) {
        if (y > 0) { num++; sum += ths[x - 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += ths[x - 1 + (y + 1) * w32]; }
        num++; sum += ths[x - 1 + y * w32];
      }
      if (x < w32 - 1) {
        if (y > 0) { num++; sum += ths[x + 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += ths[x + 1 + (y + 1) * w32]; }
        num++; sum += ths[x + 1 + y * w32];
      }
      if (y > 0) { num++; sum += ths[x + (y - 1) * w32]; }
      if (y < h32 - 1) { num++; sum += ths[x + (y + 1) * w32]; }
      num++; sum += ths[x + y * w32];
      thsSmoothed[x + y * w32] = (sum / num) * (sum / num);
    }
  }
  void select(const Frame& fh, const GlobalCalib& G, float* map_out, int pot, float thFactor, int n[3]) {
    static const float directions[16][2] = {{0, 1.0000f}, {0.3827f, 0.9239f}, {0.1951f, 0.9808f}, {0.9239f, 0.3827f}, {0.7071f, 0.7071f}, {0.3827f, -0.9239f},
                                            {0.8315f, 0.5556f}, {0.8315f, -0.5556f}, {0.5556f, -0.8315f}, {0.9808f, 0.1951f}, {0.9239f, -0.3827f},
                                            {0.7071f, -0.7071f}, {0.5556f, 0.8315f}, {0.9808f, -0.1951f}, {1.0000f, 0.0000f}, {0.1951f, -0.9808f}};
    const float* map0 = fh.dIp[0].data();
    const float* mapmax0 = fh.absSquaredGrad[0].data(); const float* mapmax1 = fh.absSquaredGrad[1].data(); const float* mapmax2 = fh.absSquaredGrad[2].data();
    const int w1 = G.w[1], w2 = G.w[2];
    memset(map_out, 0, (size_t)w * h * sizeof(float));
    const float dw1 = gradDownweightPerLevel, dw2 = dw1 * dw1;
    int n3 = 0, n2 = 0, n4 = 0;
    for (int y4 = 0; y4 < h; y4 += (4 * pot)) for (int x4 = 0; x4 < w; x4 += (4 * pot)) {
      const int my3 = std::min((4 * pot), h - y4), mx3 = std::min((4 * pot), w - x4);
      int bestIdx4 = -1; float bestVal4 = 0;
      const float* dir4 = directions[randomPattern[n2] & 0xF];
      for (int y3 = 0; y3 < my3; y3 += (2 * pot)) for (int x3 = 0; x3 < mx3; x3 += (2 * pot)) {
        const int x34 = x3 + x4, y34 = y3 + y4;
        const int my2 = std::min((2 * pot), h - y34), mx2 = std::min((2 * pot), w - x34);
        int bestIdx3 = -1; float bestVal3 = 0;
        const float* dir3 = directions[randomPattern[n2] & 0xF];
        for (int y2 = 0; y2 < my2; y2 += pot) for (int x2 = 0; x2 < mx2; x2 += pot) {
          const int x234 = x2 + x34, y234 = y2 + y34;
          const int my1 = std::min(pot, h - y234), mx1 = std::min(pot, w - x234);
          int bestIdx2 = -1; float bestVal2 = 0;
          const float* dir2 = directions[randomPattern[n2] & 0xF];
          for (int y1 = 0; y1 < my1; y1 += 1) for (int x1 = 0; x1 < mx1; x1 += 1) {
            const int idx = x1 + x234 + w * (y1 + y234);
            const int xf = x1 + x234, yf = y1 + y234;
            if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
            const float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * thsStep];
            const float pixelTH1 = pixelTH0 * dw1;
            const float pixelTH2 = pixelTH1 * dw2;
            const float ag0 = mapmax0[idx];
            if (ag0 > pixelTH0 * thFactor) {
              const float dirNorm = fabsf((float)(map0[3 * idx + 1] * dir2[0] + map0[3 * idx + 2] * dir2[1]));
              if (dirNorm > bestVal2) { bestVal2 = dirNorm; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2; }
            }
            if (bestIdx3 == -2) continue;
            const float ag1 = mapmax1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1];
            if (ag1 > pixelTH1 * thFactor) {
              const float dirNorm = fabsf((float)(map0[3 * idx + 1] * dir3[0] + map0[3 * idx + 2] * dir3[1]));
              if (dirNorm > bestVal3) { bestVal3 = dirNorm; bestIdx3 = idx; bestIdx4 = -2; }
            }
            if (bestIdx4 == -2) continue;
            const float ag2 = mapmax2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2];
            if (ag2 > pixelTH2 * thFactor) {
              const float dirNorm = fabsf((float)(map0[3 * idx + 1] * dir4[0] + map0[3 * idx + 2] * dir4[1]));
              if (dirNorm > bestVal4) { bestVal4 = dirNorm; bestIdx4 = idx; }
            }
          }
          if (bestIdx2 > 0) { map_out[bestIdx2] = 1; bestVal3 = 1e10; n2++; }
        }
        if (bestIdx3 > 0) { map_out[bestIdx3] = 2; bestVal4 = 1e10; n3++; }
      }
      if (bestIdx4 > 0) { map_out[bestIdx4] = 4; n4++; }
    }
    n[0] = n2; n[1] = n3; n[2] = n4;
  }
  int makeMaps(const Frame& fh, const GlobalCalib& G, float* map_out, float density, int recursionsLeft, float thFactor) {
    float numHave = 0, numWant = density, quotia;
    int idealPotential = currentPotential;
    if (&fh != gradHistFrame) makeHists(fh);
    int n[3];
    select(fh, G, map_out, currentPotential, thFactor, n);
    numHave = n[0] + n[1] + n[2];
    quotia = numWant / numHave;
    const float K = numHave * (currentPotential + 1) * (currentPotential + 1);
    idealPotential = sqrtf(K / numWant) - 1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft > 0 && quotia > 1.25 && currentPotential > 1) {
      if (idealPotential >= currentPotential) idealPotential = currentPotential - 1;
      currentPotential = idealPotential;
      return makeMaps(fh, G, map_out, density, recursionsLeft - 1, thFactor);
    } else if (recursionsLeft > 0 && quotia < 0.25) {
      if (idealPotential <= currentPotential) idealPotential = currentPotential + 1;
      currentPotential = idealPotential;
      return makeMaps(fh, G, map_out, density, recursionsLeft - 1, thFactor);
    }
    int numHaveSub = numHave;
    if (quotia < 0.95) {
      const int wh = w * h;
      int rn = 0;
      const unsigned char charTH = 255 * quotia;
      for (int i = 0; i < wh; i++) if (map_out[i] != 0) { if (randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; } rn++; }
    }
    currentPotential = idealPotential;
    return numHaveSub;
  }
};

}  // namespace orc

