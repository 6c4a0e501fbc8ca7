// Device-side data model of the sliding window: SoA arenas indexed by integer ids
// instead of the reference's pointer graph (FrameHessian / PointHessian / PointFrameResidual / EF* mirrors).
//
// Residuals are STORED sorted by their (host,target) key = host + target*n (the reference's htIDX,
// AccumulatedTopHessian.cpp:84). A "slot" is a position in that order. Consequences:
//   - every warp of the linearise kernel works on one (host,target) pair: the FrameFramePrecalc is a
//     broadcast load, all gathers of the warp hit one target image, point records are read in ascending order;
//   - the per-(host,target) 13x13 accumulators become segmented block reductions without atomics, so the
//     summation order is fixed and results are reproducible run to run.
// The caller's residual ids (order of sdso_ba_set_residuals) are mapped through rid2slot / slot2rid.
#pragma once
#include "ctx.h"
#include <vector>

namespace sdso {

constexpr int kMaxFrames = 16;  // window size supported on device (reference: setting_maxFrames 7; SURVEY config 4 uses 10)
constexpr int kJ = 74;          // floats of RawResidualJacobian (OptimizationBackend/RawResidualJacobian.h:32-65)
// plane offsets inside the 74-plane SoA record (plane p of slot s lives at J[p * Rcap + s])
enum { J_RESF = 0, J_PDXI = 8, J_PDC = 20, J_PDD = 28, J_IDX = 30, J_AB = 46, J_IDX2 = 62, J_ABIDX = 66, J_AB2 = 70 };

enum { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };  // Residuals.h:49
enum { RF_LINEARIZED = 1, RF_ACTIVE = 2, RF_DROPPED = 4 };  // EFResidual::isLinearized / isActiveAndIsGoodNEW / removed from the graph (dropResidual)
enum { PS_GOOD = 0, PS_MARGINALIZE = 1, PS_DROP = 2 };  // EnergyFunctionalStructs.h:97

constexpr int kTopVals = 96;   // 55 (10x10 upper triangle) + 30 (10x3) + 6 (3x3 upper triangle), padded to 3x32
constexpr int kChunk = 256;    // slots per CTA of the segmented reductions

struct PrecalcDev {  // FrameFramePrecalc (HessianBlocks.h:72-97), index host*n + target
  float PRE_RTll[9], PRE_KRKiTll[9], PRE_RKiTll[9], PRE_RTll_0[9];
  float PRE_tTll[3], PRE_KtTll[3], PRE_tTll_0[3];
  float PRE_aff_mode[2];
  float PRE_b0_mode, distanceLL;
  float pad[3];
};

struct BACalib {  // CalibHessian value_scaledf / value_scaledi (HessianBlocks.h:300-340) + wM3G, hM3G
  float fxl, fyl, cxl, cyl, fxli, fyli, cxli, cyli;
  float wM3G, hM3G;
  int w0, h0;
  float huberTH, outlierTHSumComponent, affineOptModeA, affineOptModeB;
};

struct HostBAFrame {
  int frame_id = -1, frameID = 0;
  float ab_exposure = 1;
  float frameEnergyTH = 8 * 8 * 8;
  double T_eval[12];                       // worldToCam_evalPT
  double state[10] = {0}, state_zero[10] = {0}, state_scaled[10] = {0};
  double T_w2c[12], T_c2w[12];             // PRE_worldToCam / PRE_camToWorld
  double ns_pose[36], ns_scale[6];         // nullspaces (HessianBlocks.cpp:78-123)
  double prior[8] = {0}, delta_prior[8] = {0}, delta[8] = {0};
  double step[10] = {0}, state_backup[10] = {0};
};

struct Chunk { int key, begin, end, pad; };

// Device mirror of the per-frame state the LM loop of FullSystem::optimize moves (FrameHessian state / state_zero / state_backup,
// worldToCam_evalPT, PRE_worldToCam / PRE_camToWorld; HessianBlocks.h:121-231): with it the loop runs without host round trips.
struct FrameDev {
  double T_eval[12];
  double state[10], state_zero[10], state_backup[10];
  double T_w2c[12], T_c2w[12];
  float ab_exposure, pad;
};
// Control block of the device-resident LM loop (sdso_ba_optimize): iteration counter, convergence flags, calibration value
struct OptDev {
  double calib_value[4], calib_zero[4], calib_backup[4];   // CalibHessian value / value_zero / value_backup (unscaled)
  int it;          // LM iterations started so far (the solve orthogonalises x from iteration 2 on)
  int pending;     // the convergence test of this iteration's step passed: the loop ends after its linearizeAll / applyRes
  int done;        // every kernel of the iteration chain exits at once when set (the host enqueues all mnumOptIts iterations blindly)
  int its_done;    // what FullSystem::optimize's loop counter would be
  float th_opt; int min_its; int pad[2];
};

struct BAState {
  int n = 0, P = 0, R = 0;
  bool prepared = false;
  BACalib calib;
  double calib_delta[4] = {0, 0, 0, 0};   // HCalib.value_minus_value_zero
  double cPrior[4] = {0, 0, 0, 0};
  double calib_value[4] = {0, 0, 0, 0}, calib_zero[4] = {0, 0, 0, 0}, calib_step[4] = {0, 0, 0, 0}, calib_backup[4] = {0, 0, 0, 0};  // CalibHessian value / value_zero / step / value_backup
  std::vector<HostBAFrame> frames;
  // host copies of the graph
  std::vector<int> h_p_host, h_r_point, h_r_target, h_rid2slot, h_slot2rid, h_seg_begin;
  std::vector<float> h_p_idepth, h_p_idepth_zero;
  std::vector<Chunk> h_chunks;
  // ---- device: frames
  const float4** d_tex0 = nullptr;        // [kMaxFrames] level-0 texels (target->dI)
  float* d_frameTH = nullptr;             // [kMaxFrames]
  PrecalcDev* d_precalc = nullptr;        // [kMaxFrames^2]
  double* d_adHost = nullptr; double* d_adTarget = nullptr;  // [n*n][64], index h + t*n
  float* d_adHostF = nullptr; float* d_adTargetF = nullptr;
  float* d_adHTdeltaF = nullptr;          // [n*n][8]
  float* d_cDeltaF = nullptr;             // [4]
  double* d_fprior = nullptr;             // [kMaxFrames][24]: prior[8], delta_prior[8], delta[8]
  // ---- device: points
  int capP = 0;
  int* d_p_host = nullptr; float* d_p_u = nullptr; float* d_p_v = nullptr;
  float* d_p_idepth = nullptr; float* d_p_idepth_zero = nullptr;
  float4* d_p_color = nullptr; float4* d_p_weights = nullptr;  // [P][2]
  float* d_p_priorF = nullptr; float* d_p_deltaF = nullptr; float* d_p_idepth_backup = nullptr;
  int* d_p_res_begin = nullptr;           // CSR [P+1] into d_p_res_list (slots, residualsAll order)
  int* d_slot_of = nullptr;               // [P*n]: slot of the point's residual towards target t, or -1
  float* d_p_acc = nullptr;               // [16][capP]: Hdd_A, bd_A, Hcd_A[4], Hdd_L, bd_L, Hcd_L[4], HdiF, bdSumF, step, idepth_hessian
  unsigned char* d_p_flag = nullptr;      // EFPointStatus
  // ---- device: residual slots
  int capR = 0;
  int* d_p_res_list = nullptr;
  int* d_s_point = nullptr; int* d_s_key = nullptr;
  unsigned char* d_s_state = nullptr; unsigned char* d_s_newstate = nullptr; unsigned char* d_s_flags = nullptr; unsigned char* d_s_sel = nullptr;
  float* d_s_energy = nullptr;            // [3][capR]: state_energy, state_NewEnergy, state_NewEnergyWithOutlier
  float* d_J = nullptr;                   // [2][74][capR]; efJ of slot s = buffer d_s_sel[s], candidate J = the other
  float* d_s_rtz = nullptr; float* d_s_JpJd = nullptr;   // [8][capR]
  float* d_s_center = nullptr;            // [3][capR]
  float* d_s_psum = nullptr;              // [6][capR]: per-residual terms of bd_acc, Hdd_acc, Hcd_acc[4]
  int* d_slot2rid = nullptr; int* d_rid2slot = nullptr;
  Chunk* d_chunks = nullptr; int nchunks = 0; int capChunks = 0;
  int* d_key_chunk_begin = nullptr;       // [n*n+1] first chunk of each key
  // ---- device: accumulators and the reduced system
  float* d_tpart = nullptr;               // [nchunks][96] top partials
  float* d_dpart = nullptr;               // [nchunks][n+1][64] SC partials (t2 < n: D blocks; t2 == n: E (32) + EB (8))
  float* d_pblockpart = nullptr;          // [pblocks][32] per-block partials of accHcc (16) + accbc (4)
  int pblocks = 0;
  double* d_G = nullptr;                  // [n*n][169] stitched-input 13x13 blocks (double), rebuilt per accumulate_top
  float* d_Gf = nullptr;                  // [n*n][169] same in float (operator-level readback)
  double* d_D = nullptr;                  // [n*n][n][64]
  double* d_E = nullptr;                  // [n*n][40]: E 8x4, EB 8
  double* d_Hcc = nullptr;                // [20]
  double* d_U = nullptr; double* d_V = nullptr;  // [n*n*n][64]: adHost*D, adTarget*D
  double* d_sys = nullptr;                // see SYS_* offsets
  size_t sys_stride = 0;                  // dmax*dmax + dmax doubles per (H,b) pair
  double* d_energy_part = nullptr;        // per-block energy partials
  double* d_scalars = nullptr;            // [16]
  unsigned int* d_counter = nullptr;
  double* d_N = nullptr;                  // [d][7] orthonormal basis of the gauge nullspace (first nrank columns)
  int nrank = 0;
  float* d_xAd = nullptr;                 // [n*n][8]
  int* d_list = nullptr;                  // scratch slot list
  double* d_step_part = nullptr; int step_part_cap = 0;  // per-block partial sums of doStepFromBackup
  FrameDev* d_frames = nullptr;           // [kMaxFrames]
  BACalib* d_calib = nullptr;             // the kernels read the calibration from here (it moves inside the LM loop)
  OptDev* d_opt = nullptr;
  void* graph_iter = nullptr;             // cudaGraphExec_t of one LM iteration of sdso_ba_optimize (captured per prepared window)
  void* graph_assemble = nullptr;         // cudaGraphExec_t of the accumulate + stitch + assemble chain (sdso_ba_assemble / sdso_ba_solve)
  int graph_iter_nodes = 0, graph_assemble_nodes = 0;   // kernels per replay (launch accounting)
  cudaStream_t cap_stream = nullptr;      // capture stream (the context's stream may be the legacy default stream, which cannot capture)
  cudaStream_t cap_stream2 = nullptr;     // second capture stream: the forked half of the accumulate / stitch chain
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double* d_W = nullptr;                  // [F^2][64 + 64 + 40 + 40]: adHost*A, adTarget*A and their calibration columns (top stitch scratch)
  void* lba_scratch = nullptr; size_t lba_scratch_bytes = 0;   // sdso_lba_g2o's graph arrays
  bool any_linearized = false;            // some residual has been through fixLinearizationF since the window was uploaded
  int shard_rank = 0, shard_n = 1;        // point-sharded window: priors and HM enter on rank 0 only
  bool have_M = false;                    // HM/bM (marginalisation prior) present in SYS_M
  std::vector<double> h_N, h_adH, h_adT;
  std::vector<float> h_adHTd;
  std::vector<PrecalcDev> h_pre;
  int dim() const { return kCPARS + 8 * n; }
};

// (H,b) pairs inside d_sys
enum { SYS_A = 0, SYS_L = 1, SYS_SC = 2, SYS_M = 3, SYS_FINAL = 4, SYS_TMP = 5, SYS_X = 6, SYS_NUM = 7 };

}  // namespace sdso
