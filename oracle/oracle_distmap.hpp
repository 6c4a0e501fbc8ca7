// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.hpp).
// CoarseDistanceMap::{makeDistanceMap, growDistBFS, addIntoDistFinal} (FullSystem/CoarseTracker.cpp:1216-1366) and the
// candidate filter of FullSystem::activatePointsMT, STEP2 (FullSystem.cpp:838-901). Sequential restatement: the BFS lists,
// the alternating 8- / 4-neighbourhood steps, border cells that never expand, and the candidate loop that inserts every
// accepted candidate into the distance field before the next one is judged.
#pragma once
#include "oracle_trace.hpp"
#include <vector>

namespace orc {

struct CoarseDistanceMap {
  int w1 = 0, h1 = 0;
  std::vector<float> fwdWarpedIDDistFinal;
  std::vector<int> bfsList1, bfsList2;  // (x, y) pairs
  void init(int w1_, int h1_) { w1 = w1_; h1 = h1_; fwdWarpedIDDistFinal.assign((size_t)w1 * h1, 1000.f); bfsList1.assign((size_t)w1 * h1 * 2 * 9 + 16, 0); bfsList2 = bfsList1; }
  // the point loop of makeDistanceMap for one host frame (:1231-1250)
  int seed(const float KRKi[9], const float Kt[3], int n, const float* uvid, int numItems) {
    for (int i = 0; i < n; i++) {
      const float pu = uvid[3 * i], pv = uvid[3 * i + 1], id = uvid[3 * i + 2];
      const float p0 = (KRKi[0] * pu + KRKi[1] * pv + KRKi[2] * 1.f) + Kt[0] * id;
      const float p1 = (KRKi[3] * pu + KRKi[4] * pv + KRKi[5] * 1.f) + Kt[1] * id;
      const float p2 = (KRKi[6] * pu + KRKi[7] * pv + KRKi[8] * 1.f) + Kt[2] * id;
      const float fu = p0 / p2 + 0.5f, fv = p1 / p2 + 0.5f;
      if (!(fu > -2e9f && fu < 2e9f && fv > -2e9f && fv < 2e9f)) continue;  // int conversion of NaN / inf is undefined; such points fail the test below
      const int u = (int)fu, v = (int)fv;
      if (!(u > 0 && v > 0 && u < w1 && v < h1)) continue;
      fwdWarpedIDDistFinal[u + w1 * v] = 0;
      bfsList1[2 * numItems] = u; bfsList1[2 * numItems + 1] = v;
      numItems++;
    }
    return numItems;
  }
  void growDistBFS(int bfsNum) {
    static const int d4[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    static const int d8[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
    for (int k = 1; k < 40; k++) {
      const int bfsNum2 = bfsNum;
      std::swap(bfsList1, bfsList2);
      bfsNum = 0;
      const int nn = (k % 2 == 0) ? 4 : 8;
      const int(*dd)[2] = (k % 2 == 0) ? d4 : d8;
      for (int i = 0; i < bfsNum2; i++) {
        const int x = bfsList2[2 * i], y = bfsList2[2 * i + 1];
        if (x == 0 || y == 0 || x == w1 - 1 || y == h1 - 1) continue;
        for (int j = 0; j < nn; j++) {
          const int xx = x + dd[j][0], yy = y + dd[j][1];
          if (fwdWarpedIDDistFinal[xx + yy * w1] > k) {
            fwdWarpedIDDistFinal[xx + yy * w1] = k;
            bfsList1[2 * bfsNum] = xx; bfsList1[2 * bfsNum + 1] = yy; bfsNum++;
          }
        }
      }
    }
  }
  void addIntoDistFinal(int u, int v) {
    bfsList1[0] = u; bfsList1[1] = v;
    fwdWarpedIDDistFinal[u + w1 * v] = 0;
    growDistBFS(1);
  }
};

// verdicts of the candidate loop: 0 stays immature, 1 goes to toOptimize, 2 deleted
inline void activationFilter(CoarseDistanceMap& M, int n_hosts, const float* KRKi, const float* Kt, const unsigned char* flaggedForMarg, int n,
                             const int* cand_host, const ImmaturePoint* pts, const float* my_type, float currentMinActDist, float minTraceQuality,
                             int* verdict) {
  for (int i = 0; i < n; i++) {
    const ImmaturePoint& ph = pts[i];
    const int hst = cand_host[i];
    if (!std::isfinite(ph.idepth_max) || ph.lastTraceStatus == IPS_OUTLIER) { verdict[i] = 2; continue; }
    const bool canActivate = (ph.lastTraceStatus == IPS_GOOD || ph.lastTraceStatus == IPS_SKIPPED || ph.lastTraceStatus == IPS_BADCONDITION ||
                              ph.lastTraceStatus == IPS_OOB) &&
                             ph.lastTracePixelInterval < 8 && ph.quality > minTraceQuality && (ph.idepth_max + ph.idepth_min) > 0;
    if (!canActivate) { verdict[i] = (flaggedForMarg[hst] || ph.lastTraceStatus == IPS_OOB) ? 2 : 0; continue; }
    const float* Kr = KRKi + 9 * hst; const float* kt = Kt + 3 * hst;
    const float id = 0.5f * (ph.idepth_max + ph.idepth_min);
    const float p0 = (Kr[0] * ph.u + Kr[1] * ph.v + Kr[2] * 1.f) + kt[0] * id;
    const float p1 = (Kr[3] * ph.u + Kr[4] * ph.v + Kr[5] * 1.f) + kt[1] * id;
    const float p2 = (Kr[6] * ph.u + Kr[7] * ph.v + Kr[8] * 1.f) + kt[2] * id;
    const float fu = p0 / p2 + 0.5f, fv = p1 / p2 + 0.5f;
    const bool conv = fu > -2e9f && fu < 2e9f && fv > -2e9f && fv < 2e9f;
    const int u = conv ? (int)fu : -1, v = conv ? (int)fv : -1;
    if (u > 0 && v > 0 && u < M.w1 && v < M.h1) {
      const float dist = M.fwdWarpedIDDistFinal[u + M.w1 * v] + (p0 - floorf((float)(p0)));
      if (dist >= currentMinActDist * my_type[i]) { M.addIntoDistFinal(u, v); verdict[i] = 1; }
      else verdict[i] = 0;
    } else verdict[i] = 2;
  }
}

}  // namespace orc
