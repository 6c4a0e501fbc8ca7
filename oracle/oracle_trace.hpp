// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// ImmaturePoint (FullSystem/ImmaturePoint.h:59-114): fields read or written by the constructor, traceOn, traceStereo.
#pragma once
#include "oracle_core.hpp"

namespace orc {

enum ImmaturePointStatus { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };  // ImmaturePoint.h:50-56

// Same field order as sdso_immature_point in include/sdso_b200.h so tests can share one numpy record type.
struct ImmaturePoint {
  float u = 0, v = 0;
  float idepth_min = 0, idepth_max = NAN;
  float quality = 10000, energyTH = 0;
  float color[8] = {0}, weights[8] = {0}, gradH[4] = {0};
  float u_stereo = 0, v_stereo = 0, idepth_min_stereo = 0, idepth_max_stereo = NAN, idepth_stereo = 0;
  float lastTraceUV[2] = {0, 0}, lastTracePixelInterval = 0;
  int32_t lastTraceStatus = IPS_UNINITIALIZED;
  int32_t bestIdx = -1, numSteps = 0;  // diagnostics of the last discrete search
};

bool immatureInit(const GlobalCalib& G, const Settings& S, const Frame& host, float u, float v, ImmaturePoint& p);
int traceOn(const GlobalCalib& G, const Settings& S, ImmaturePoint& p, const Frame& frame, const float KRKi[9], const float Kt[3], const float aff[2]);
int traceStereo(const GlobalCalib& G, const Settings& S, ImmaturePoint& p, const Frame& frame, const float K[9], bool mode_right);

struct BAWindow;
// D4 (oracle/activate.cpp)
int activatePoint(const BAWindow& W, int host, const ImmaturePoint& p, int variant, int minObs, float* idepth_out, int* states, float* energy_out);

}  // namespace orc
