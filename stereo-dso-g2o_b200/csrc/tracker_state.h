// Host-side state of the device CoarseTracker (shared by tracker.cu and tracker_ref.cu).
#pragma once
#include "ctx.h"

namespace sdso {

struct TrackProblem;

// One reference keyframe's tracking template (the members of the same names below hold the CURRENT slot)
struct RefSlot {
  float4* pc[kPyrLevels] = {nullptr};
  int pc_n[kPyrLevels] = {0};
  int pc_cap[kPyrLevels] = {0};
  int ref_frame = -1;
  float ref_exposure = 1.0f;  // snapshot at setCoarseTrackingRef time: the frame slot may be released / reused afterwards
  double ref_aff[2] = {0, 0};
  bool have_ref = false;
};

struct TrackerState {
  HostCalib K;  // tracker's own pyramid of intrinsics (makeK from the optimised HCalib)
  float4* pc[kPyrLevels] = {nullptr};
  int pc_n[kPyrLevels] = {0};
  int pc_cap[kPyrLevels] = {0};
  int ref_frame = -1;
  float ref_exposure = 1.0f;
  double ref_aff[2] = {0, 0};
  bool have_ref = false;
  TrackProblem* d_problems = nullptr;
  TrackProblem* h_problems = nullptr;  // pinned
  int max_problems = 2048;
  cudaEvent_t results_ready = nullptr;     // recorded behind the D2H copy of the results: collect waits for it, not for the whole stream
  unsigned int* d_work_counter = nullptr;  // dynamic problem scheduling of the throughput configuration
  std::vector<RefSlot> saved;  // parked reference slots (independent sequences tracked by one launch); saved[cur_slot] is stale
  int cur_slot = 0;
  float* d_dump = nullptr;
  size_t dump_cap = 0;
  int last_nb = 0;
  long long last_cyc[16] = {0};
  unsigned char* edge_flag[kPyrLevels] = {nullptr};
  double* edge_err[kPyrLevels] = {nullptr};
  size_t edge_cap[kPyrLevels] = {0};
  // A4 scratch
  float* idepth[kPyrLevels] = {nullptr};
  float* wsum[kPyrLevels] = {nullptr};
  float* wsum_bak[kPyrLevels] = {nullptr};
  int* scan_tmp = nullptr;
  int* d_counts = nullptr;
};


}  // namespace sdso
