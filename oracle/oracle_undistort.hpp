// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// Input preparation and trajectory output, the wire formats either side of the path:
// PhotometricUndistorter::processFrame (util/Undistort.cpp:222-260), the remap loop of Undistort::undistort (:398-489,
// benchmark noise off as in the reference's defaults) and the row format of FullSystem::printResult (FullSystem.cpp:236-285).
#pragma once
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>

namespace orc {

struct Undistorter {
  int wOrg = 0, hOrg = 0, w = 0, h = 0;
  std::vector<float> remapX, remapY, G, vignetteMapInv;  // G: 256 entries when a response calibration is loaded
  int photometricCalibration = 2;  // setting_photometricCalibration (settings.cpp:35)
  bool useExposure = true;         // setting_useExposure
  bool valid() const { return !G.empty(); }
  // returns output->exposure_time
  float processFrame(const unsigned char* image_in, float exposure_time, float factor, float* data) const {
    const int wh = wOrg * hOrg;
    if (!valid() || exposure_time <= 0 || photometricCalibration == 0) {
      for (int i = 0; i < wh; i++) data[i] = factor * image_in[i];
    } else {
      for (int i = 0; i < wh; i++) data[i] = G[image_in[i]];
      if (photometricCalibration == 2)
        for (int i = 0; i < wh; i++) data[i] *= vignetteMapInv[i];
    }
    return useExposure ? exposure_time : 1.f;
  }
  float undistort(const unsigned char* raw, float exposure, float factor, float* out_data) const {
    std::vector<float> in((size_t)wOrg * hOrg);
    const float e = processFrame(raw, exposure, factor, in.data());
    const float* in_data = in.data();
    for (int idx = w * h - 1; idx >= 0; idx--) {
      float xx = remapX[idx], yy = remapY[idx];
      if (xx < 0) out_data[idx] = 0;
      else {
        const int xxi = xx, yyi = yy;
        xx -= xxi; yy -= yyi;
        const float xxyy = xx * yy;
        const float* src = in_data + xxi + yyi * wOrg;
        out_data[idx] = xxyy * src[1 + wOrg] + (yy - xxyy) * src[wOrg] + (xx - xxyy) * src[1] + (1 - xx - yy + xxyy) * src[0];
      }
    }
    return e;
  }
};

// one row of the KITTI-format trajectory file: camToWorld as R(0,:) t0 R(1,:) t1 R(2,:) t2. The reference streams doubles with
// setprecision(15) and the default float field, which the C++ standard defines as printf's %.15g.
inline std::string trajectoryRow(const double T[12]) {
  std::string s;
  char tmp[64];
  for (int i = 0; i < 12; i++) { snprintf(tmp, sizeof(tmp), "%.15g", T[i]); s += tmp; s += (i == 11 ? "\n" : " "); }
  return s;
}

}  // namespace orc
