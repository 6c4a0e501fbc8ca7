/*
 * sdso_b200.h — C ABI of the B200-native photometric hot path for Stereo-DSO-g2o.
 *
 * This is the drop-in boundary: plain C, opaque handles, caller-owned host
 * buffers, int status returns, no exceptions and no Eigen/Sophus/torch types in any signature.
 * The reference has no FFI layer of its own — its operators are C++ classes linked into libdso.a
 * (CMakeLists.txt:159-163) — so every entry point below names the reference member function it
 * replaces (file:line into the reference's src/). The C++ adapter classes that carry the
 * reference's own names/signatures on top of this ABI live in
 * stereo-dso-g2o_b200/host/dso_adapters.hpp; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - SE3 crosses the boundary as double[12], row-major 3x4 [R|t].
 *   - AffLight crosses as double[2] = {a, b}  (util/NumType.h:152-175).
 *   - Images are float32, row-major, w*h, values as produced by the reference's undistorter
 *     (raw 0..255 in mode=1).
 *   - Every function returns SDSO_OK (0) or a negative SDSO_E_* code; sdso_last_error() gives text.
 *   - All work is enqueued on the context's CUDA stream (sdso_set_stream); calls that return data
 *     to host buffers synchronise that stream before returning.
 *   - There is NO CPU fallback: if no CUDA device is usable, sdso_ctx_create fails.
 */
#ifndef SDSO_B200_H_
#define SDSO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDSO_OK 0
#define SDSO_E_INVALID -1   /* bad argument */
#define SDSO_E_CUDA -2      /* CUDA runtime error (see sdso_last_error) */
#define SDSO_E_NODEVICE -3  /* no usable CUDA device */
#define SDSO_E_STATE -4     /* call order violated (e.g. track before set_ref) */
#define SDSO_E_NOMEM -5

#define SDSO_PYR_LEVELS 6       /* util/settings.h:46 */
#define SDSO_PATTERN_NUM 8      /* util/settings.h:177 */
#define SDSO_MAX_RES_PER_POINT 8

/* which of the reference's two algorithmic variants an operator follows */
#define SDSO_VARIANT_SSE 0 /* original DSO arithmetic (calcRes/calcGSSSE, LM with multiplicative lambda) */
#define SDSO_VARIANT_G2O 1 /* live fork code: g2o edges + restated g2o Levenberg */

typedef struct sdso_ctx sdso_ctx;

/* Tunables the hot path reads. Defaults = util/settings.cpp with main_dso_pangolin.cpp preset=0 mode=1. */
typedef struct sdso_settings {
  float huberTH;                    /* settings.cpp:95  */
  float coarseCutoffTH;             /* :102 */
  float outlierTH;                  /* :72  */
  float outlierTHSumComponent;      /* :73  */
  float overallEnergyTHWeight;      /* :101 */
  float maxPixSearch;               /* :111 */
  int32_t minTraceTestRadius;       /* :113 */
  float trace_stepsize;             /* :115 */
  int32_t trace_GNIterations;       /* :116 */
  float trace_GNThreshold;          /* :117 */
  float trace_extraSlackOnTH;       /* :118 */
  float trace_slackInterval;        /* :119 */
  float trace_minImprovementFactor; /* :120 */
  float affineOptModeA;             /* main_dso_pangolin.cpp:326 */
  float affineOptModeB;             /* :327 */
  int32_t gammaWeightsPixelSelect;  /* settings.cpp:93 */
  int32_t g2o_stop_flag_persists;   /* SURVEY.md Appendix C open point (1); default 1 */
  int32_t cluster_size;             /* thread-block cluster size of the persistent kernels (0 = default 8) */
  int32_t block_threads;            /* threads per CTA of the persistent kernels (0 = default 256) */
  int32_t gather_batch;             /* points gathered per thread before the arithmetic: 1, 2 or 4 (0 = default) */
  /* windowed BA (util/settings.cpp:42-52,76) */
  float idepthFixPrior;             /* :42 */
  float idepthFixPriorMargFac;      /* :43 */
  float initialRotPrior;            /* :44 */
  float initialTransPrior;          /* :45 */
  float initialAffBPrior;           /* :46 */
  float initialAffAPrior;           /* :47 */
  float initialCalibHessian;        /* :48 */
  float margWeightFac;              /* :76 */
  double solverModeDelta;           /* :52 */
  int32_t minOptIterations;         /* :68 */
  float thOptIterations;            /* :69 */
  float frameEnergyTHConstWeight;   /* :98 */
  float frameEnergyTHN;             /* :99 */
  float frameEnergyTHFacMedian;     /* :100 */
  float minGradHistCut;             /* :105  pixel selector */
  float minGradHistAdd;             /* :106 */
  float gradDownweightPerLevel;     /* :107 */
  float desiredImmatureDensity;     /* :59 */
  float minTraceQuality;            /* :112 activation candidate filter */
  int32_t track_cache;              /* device tuning: 1 = per-CTA shared-memory cache of texel patches / point records across LM iterations (default) */
} sdso_settings;

void sdso_default_settings(sdso_settings* s);

/* ---- context ---------------------------------------------------------------------------------
 * Replaces the globals set by setGlobalCalib(w,h,K) (util/globalCalib.cpp:48-108): wG,hG,fxG..,KG,
 * KiG,wM3G,hM3G,baseline,pyrLevelsUsed. K = {fx,fy,cx,cy} of the (already cropped) working image. */
int sdso_ctx_create(sdso_ctx** out, int device, int w, int h, const float K[4], float baseline,
                    const sdso_settings* settings /* nullable */);
void sdso_ctx_destroy(sdso_ctx* ctx);
const char* sdso_last_error(const sdso_ctx* ctx);
int sdso_set_stream(sdso_ctx* ctx, void* cuda_stream /* cudaStream_t, NULL = default */);
int sdso_synchronize(sdso_ctx* ctx);
int sdso_pyr_levels(const sdso_ctx* ctx);
int sdso_level_size(const sdso_ctx* ctx, int lvl, int* w, int* h);
/* per-level initial intrinsics KG[lvl], KiG[lvl] (row-major 3x3 float) */
int sdso_level_K(const sdso_ctx* ctx, int lvl, float K[9], float Ki[9]);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t sdso_launch_count(const sdso_ctx* ctx);
/* CUDA-event timing of the hot launches on the context's stream: enable, run, then read the summed
 * durations (ms) and launch counts of the track kernel and of the makeImages kernel pair since the
 * last read. Used by bench.py for the roofline figure; off by default. */
int sdso_profile_enable(sdso_ctx* ctx, int on);
int sdso_profile_read(sdso_ctx* ctx, double* track_ms, int* track_launches, double* images_ms, int* images_launches);
/* with profiling enabled: SM-clock cycles spent per phase inside the last collected track launch
 * ([0] serial LM step, [1] point loop, [2] block reduction, [3] cluster barrier + final sum, [4] bookkeeping) */
int sdso_track_phase_cycles(sdso_ctx* ctx, long long cyc[16]);
/* CalibHessian::B (FullSystem/HessianBlocks.h:352): photometric response table used by getBGradOnly
 * in makeImages; identity (B[i]=i) unless set. */
int sdso_set_gamma(sdso_ctx* ctx, const float B[256]);

/* ---- A1: FrameHessian::makeImages (FullSystem/HessianBlocks.cpp:141-203) ------------------------
 * A frame is a device-resident pyramid: per level float4{I,dx,dy,absSquaredGrad} (the reference's
 * Vector3f dIp[lvl] + float absSquaredGrad[lvl], HessianBlocks.h:107-109, fused into one 16-byte
 * texel). Border rows 0 and h-1 of dx/dy/absSquaredGrad, which the reference leaves uninitialised,
 * are written as 0. */
int sdso_frame_create(sdso_ctx* ctx, int* frame_id);
int sdso_frame_release(sdso_ctx* ctx, int frame_id);
/* host image -> H2D copy -> pyramid kernels. use_hcalib != 0 mirrors passing a non-null CalibHessian*
 * (gamma-gradient factor on absSquaredGrad; B is the identity in mode=1). */
int sdso_make_images(sdso_ctx* ctx, int frame_id, const float* host_image, float ab_exposure, int use_hcalib);
/* same, image already on the device (device pointer, w*h floats) */
int sdso_make_images_device(sdso_ctx* ctx, int frame_id, const float* device_image, float ab_exposure, int use_hcalib);
/* Batched forms (one pyramid launch + one gradient launch per 32 frames). src_u8 != 0: the sources are
 * 8-bit grey images; the widening to float — PhotometricUndistorter::processFrame in mode=1 (util/Undistort.cpp:222-260,
 * no response / vignette calibration) — is fused into the pyramid kernel, which cuts the H2D bytes by 4.
 *   sdso_upload_images_async : H2D of the sources on the context's copy stream (returns at once; overlaps running kernels)
 *   sdso_make_images_uploaded: the compute stream waits for those copies, then builds the nb pyramids
 *   sdso_make_images_batch_device: sources already in device memory */
int sdso_upload_images_async(sdso_ctx* ctx, int nb, const int* frame_ids, const void* const* host_images, int src_u8);
int sdso_make_images_uploaded(sdso_ctx* ctx, int nb, const int* frame_ids, const float* ab_exposure /* nullable: 1 */, int use_hcalib);
int sdso_make_images_batch_device(sdso_ctx* ctx, int nb, const int* frame_ids, const void* const* device_images, int src_u8,
                                  const float* ab_exposure /* nullable: 1 */, int use_hcalib);
/* read back level lvl: dI3 = w_l*h_l*3 floats AoS {I,dx,dy} (the reference's layout), absgrad = w_l*h_l */
int sdso_frame_download(sdso_ctx* ctx, int frame_id, int lvl, float* dI3 /* nullable */, float* absgrad /* nullable */);
/* getInterpolatedElement33 / 33BiLin (util/globalFuncs.h:73-86, 160-184) at n points of level lvl */
int sdso_interp33(sdso_ctx* ctx, int frame_id, int lvl, const float* xy, int n, float* out3, int bilin_variant);

/* ---- A3/A4: CoarseTracker::makeK, setCoarseTrackingRef / makeCoarseDepthL0 ----------------------
 * (FullSystem/CoarseTracker.cpp:108-136, 275-534, 809-825). */
int sdso_tracker_make_k(sdso_ctx* ctx, const float K[4] /* optimised HCalib fxl,fyl,cxl,cyl */);
/* STEP1 splat .. STEP5 compaction of makeCoarseDepthL0 from n splats {u,v,idepth,weight}
 * (CoarseTracker.cpp:350-533); ref_aff = lastRef->aff_g2l(). */
int sdso_tracker_set_ref(sdso_ctx* ctx, int ref_frame, const float* uvidw /* n*4 */, int n, const double ref_aff[2]);
/* direct upload of pc_u/pc_v/pc_idepth/pc_color[lvl] (CoarseTracker.h:117-121) */
int sdso_tracker_set_pc(sdso_ctx* ctx, int ref_frame, int lvl, int n, const float* u, const float* v,
                        const float* idepth, const float* color, const double ref_aff[2]);
int sdso_tracker_get_pc(sdso_ctx* ctx, int lvl, int* n, float* u, float* v, float* idepth, float* color /* nullable, capacity w_l*h_l */);

/* ---- A5/A6: CoarseTracker::calcRes + calcGSSSE (CoarseTracker.cpp:600-792, 537-596) -------------
 * One fused pass at (refToNew, aff): rs[6] as calcRes returns it, H/b as calcGSSSE returns them
 * (8x8 row-major, scaled by SCALE_* as written at :584-595), warped_n = buf_warped_n (padded to x4).
 * warped (nullable, capacity 8*(N_l+4)) receives the eight buf_warped_* arrays, each warped_n long, in
 * the order idepth,u,v,dx,dy,residual,weight,refColor. */
int sdso_calc_res_gs(sdso_ctx* ctx, int new_frame, int lvl, const double refToNew[12], const double aff[2],
                     float cutoffTH, double rs[6], double H[64], double b[8], int* warped_n, float* warped);

/* ---- E1: EdgeSE3PosePhotoDSO::computeError / linearizeOplus (dso_g2o_edge.cpp:395-500) ----------
 * For every pc point of level lvl that passes calcRes's border test at T_select (CoarseTracker.cpp:696)
 * evaluate the edge at (T_pose, photo): err[n_out], J[n_out*8] = {J_pose(6), J_photo(2)}. */
int sdso_edge_eval(sdso_ctx* ctx, int new_frame, int lvl, const double T_select[12], const double T_pose[12],
                   const double photo[2], int* n_out, double* err, double* J8);

/* ---- A7: CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:827-1069) --------------------------
 * Whole coarse-to-fine optimisation in ONE persistent cluster kernel (no host round trips).
 * T_io / aff_io: in = initial guess (lastToNew_out, aff_g2l_out), out = result.
 * Returns (through *ok) what the reference returns. iterations[5] (nullable) = LM iterations per level. */
int sdso_track(sdso_ctx* ctx, int new_frame, double T_io[12], double aff_io[2], int coarsest_lvl,
               const double minResForAbort[5], int variant, double lastResiduals[5], double flowIndicators[3],
               int iterations[5], int* ok);
/* Batched form: nb independent (initial pose, aff) hypotheses against the same reference and the same
 * or different new frames, one cluster each, one launch (FullSystem.cpp:351-376,441-501 tries them
 * sequentially). Arrays are nb-strided versions of sdso_track's. */
int sdso_track_batch(sdso_ctx* ctx, int nb, const int* new_frames, double* T_io, double* aff_io, int coarsest_lvl,
                     const double* minResForAbort, int variant, double* lastResiduals, double* flowIndicators,
                     int* iterations, int* ok);
/* Several reference keyframes at once (independent sequences sharing one GPU): sdso_tracker_select_ref switches the slot that
 * sdso_tracker_set_ref / set_pc / get_pc / calc_res_gs operate on; sdso_track_enqueue_multi tracks problem k against the
 * template of slot ref_slots[k] (NULL: the current slot for all), one cluster per problem, ONE launch. */
int sdso_tracker_select_ref(sdso_ctx* ctx, int slot);
int sdso_track_enqueue_multi(sdso_ctx* ctx, int nb, const int* ref_slots, const int* new_frames, const double* T_in, const double* aff_in,
                             int coarsest_lvl, const double* minResForAbort, int variant);
/* async pair used by bench.py so that CUDA events bracket the kernels without host syncs */
int sdso_track_enqueue(sdso_ctx* ctx, int nb, const int* new_frames, const double* T_in, const double* aff_in,
                       int coarsest_lvl, const double* minResForAbort, int variant);
int sdso_track_collect(sdso_ctx* ctx, int nb, double* T_out, double* aff_out, double* lastResiduals,
                       double* flowIndicators, int* iterations, int* ok, uint64_t* evals /* nullable, total */);

/* ---- B1-B12: windowed bundle adjustment, SSE path (Residuals.cpp, OptimizationBackend/) ----------------
 * The window is uploaded once as SoA arenas (frames, points, residuals); every operator below runs on the
 * device. Indices: frames 0..n-1 in insertion order (= EFFrame::idx), points 0..P-1 in the reference's
 * allPoints order (EnergyFunctional.cpp:1003-1016), residuals 0..R-1 in the order given to
 * sdso_ba_set_residuals (per point = PointHessian::residuals order). Matrices are dense row-major doubles of
 * dimension dim = 4 + 8 n (CPARS + 8 per frame), as EnergyFunctional assembles them. */
int sdso_ba_reset(sdso_ctx* ctx);
/* CalibHessian::value_scaledf {fxl,fyl,cxl,cyl} (HessianBlocks.h:300-340) and value_minus_value_zero (nullable = 0) */
int sdso_ba_set_calib(sdso_ctx* ctx, const float K[4], const double value_minus_value_zero[4]);
/* FrameHessian::setEvalPT_scaled(worldToCam, aff_g2l) + EFFrame::takeData (HessianBlocks.h:223-268); frame_id = a
 * pyramid built by sdso_make_images; frameID == 0 receives the gauge priors */
int sdso_ba_add_frame(sdso_ctx* ctx, int frame_id, const double T_w2c[12], double a, double b, int frameID, int* idx_out);
int sdso_ba_set_state(sdso_ctx* ctx, int idx, const double state[10]);     /* FrameHessian::setState (HessianBlocks.h:177-199) */
int sdso_ba_set_energy_th(sdso_ctx* ctx, int idx, float frameEnergyTH);    /* FrameHessian::frameEnergyTH */
/* PointHessian / EFPoint fields: host frame idx, u, v, idepth, idepth_zero, color[8], weights[8], hasDepthPrior */
int sdso_ba_set_points(sdso_ctx* ctx, int P, const int* host, const float* u, const float* v, const float* idepth,
                       const float* idepth_zero, const float* color8, const float* weights8, const unsigned char* has_prior);
int sdso_ba_set_residuals(sdso_ctx* ctx, int R, const int* point, const int* target);  /* PointFrameResidual(point, host, target) + resetOOB */
int sdso_ba_set_point_flags(sdso_ctx* ctx, const unsigned char* flags);    /* EFPointStatus per point (0 GOOD, 1 MARGINALIZE, 2 DROP) */
/* FullSystem::setPrecalcValues (FullSystem.cpp:1633-1644): FrameFramePrecalc::set for all pairs (HessianBlocks.cpp:206-242),
 * EnergyFunctional::setAdjointsF (EnergyFunctional.cpp:41-119), setDeltaF (:173-207), getNullspaces (FullSystemOptimize.cpp:1087-1147) */
int sdso_ba_prepare(sdso_ctx* ctx);
int sdso_ba_counts(sdso_ctx* ctx, int* n, int* P, int* R, int* dim);
int sdso_ba_get_precalc(sdso_ctx* ctx, int h, int t, float out[49]);
int sdso_ba_get_adjoints(sdso_ctx* ctx, double* adHost, double* adTarget, float* adHTdeltaF);  /* [n*n][64], [n*n][64], [n*n][8]; index h + t*n */
int sdso_ba_nullspaces(sdso_ctx* ctx, double* N /* dim x 7 row-major */);
/* FullSystem::linearizeAll(fixLinearization) (FullSystemOptimize.cpp:142-203): PointFrameResidual::linearize
 * (Residuals.cpp:83-336) over the active residuals, summed energy; fix != 0 also applies applyRes(true) */
int sdso_ba_linearize_all(sdso_ctx* ctx, int fix, double* energy);
int sdso_ba_apply_res(sdso_ctx* ctx, int copy_jacobians);                 /* PointFrameResidual::applyRes (Residuals.cpp:367-385) */
/* EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:96-123) for the listed residuals (rids == NULL: all active) */
int sdso_ba_fix_linearization(sdso_ctx* ctx, int count, const int* rids);
int sdso_ba_get_res(sdso_ctx* ctx, int which /* 0 candidate J, 1 EFResidual::J */, int* newState, int* state, double* newEnergy,
                    double* newEnergyWithOutlier, int* active, int* linearized, float* J74, float* JpJdF8, float* center3, float* resToZero8);
int sdso_ba_get_points(sdso_ctx* ctx, float* out16);
/* AccumulatedTopHessianSSE::addPoint<mode> over all points + stitchDouble (AccumulatedTopHessian.cpp:36-193, 265-337);
 * blocks (nullable) = the n*n 13x13 float accumulators, index h + t*n */
int sdso_ba_accumulate_top(sdso_ctx* ctx, int mode, int use_prior, double* H, double* b, float* blocks);
/* AccumulatedSCHessianSSE::addPoint over all points + stitchDouble (AccumulatedSCHessian.cpp:34-195) */
int sdso_ba_accumulate_sc(sdso_ctx* ctx, int shift_prior_to_zero, double* H, double* b);
/* EnergyFunctional::solveSystemF(iteration, lambda, HCalib) (EnergyFunctional.cpp:838-995) incl. resubstituteF_MT;
 * x = the solved increment (before the sign flip of resubstitute), Hfinal/bfinal (nullable) = the damped reduced system */
int sdso_ba_solve(sdso_ctx* ctx, int iteration, double lambda, double* x, double* Hfinal, double* bfinal);
/* EnergyFunctional::resubstituteF_MT (:272-341) for a given x (NULL: the last solve's); frame_steps [n][10], calib_step [4];
 * point steps are read with sdso_ba_get_points */
int sdso_ba_resubstitute(sdso_ctx* ctx, const double* x, double* frame_steps, double* calib_step);
/* FullSystem::setNewFrameEnergyTH (FullSystemOptimize.cpp:98-139): exact 70 % quantile of state_NewEnergyWithOutlier over the
 * active residuals that target the newest frame -> that frame's frameEnergyTH (applied, and returned) */
int sdso_ba_new_frame_energy_th(sdso_ctx* ctx, float* th);
/* FrameHessian::frameEnergyTH of every window frame [n] as the last linearizeAll / optimize left them (setNewFrameEnergyTH moves the
 * newest frame's) */
int sdso_ba_get_energy_th(sdso_ctx* ctx, float* frameEnergyTH);
/* FullSystem::optimize, SSE body (FullSystemOptimize.cpp:870-1042) with backupState / solveSystem / doStepFromBackup /
 * linearizeAll / applyRes per iteration (setting_forceAceptStep = true, settings.cpp:53), then the new evaluation point of the
 * newest frame and the final linearizeAll(true). Returns what the reference returns: sqrt(energy / (patternNum * resInA)). */
int sdso_ba_optimize(sdso_ctx* ctx, int mnumOptIts, double* rmse, int* iterations_done);
/* current frame states [n][10], PRE_worldToCam [n][12], point idepths [P], calibration value_scaled [4] (each nullable) */
int sdso_ba_get_state(sdso_ctx* ctx, double* states, double* T_w2c, float* idepth, double* calib);
int sdso_ba_set_marg_prior(sdso_ctx* ctx, const double* HM, const double* bM);   /* EnergyFunctional::HM, bM */
int sdso_ba_get_marg_prior(sdso_ctx* ctx, double* HM, double* bM);
/* EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:663-736): points flagged PS_MARGINALIZE go into HM / bM and leave the graph */
int sdso_ba_marginalize_points(sdso_ctx* ctx);
/* EnergyFunctional::marginalizeFrame (:554-660): Schur-eliminates frame idx from HM / bM (now of dimension dim-8) and removes the
 * frame from the window; points, residuals and states must be uploaded again afterwards (the reference calls makeIDX here) */
int sdso_ba_marginalize_frame(sdso_ctx* ctx, int idx);
/* EnergyFunctional::calcMEnergyF (:344-351) and calcLEnergyF_MT (:354-442); each output nullable */
int sdso_ba_energies(sdso_ctx* ctx, double* menergy, double* lenergy);

/* ---- E2: EdgeLBASE3PosePhotoIdepthCamDSO::computeError + linearizeOplus (dso_g2o_edge.cpp:5-128, 130-282) -----------------
 * Every residual of the uploaded window is evaluated as one LBA edge with the given vertex estimates: VertexSE3PoseDSO T_wh
 * [n][12] and VertexPhotometricDSO [n][2] of the HOST frames, VertexInverseDepthDSO idepth[R] (one per residual, as
 * FullSystemOptimize.cpp:493-512 builds them), VertexCamDSO cam[4], and b0[n] (SetB). Targets use the window's PRE_worldToCam
 * and aff_g2l. Outputs (caller's residual order): error[R][8]; the four Jacobian blocks J_xi[R][8][6], J_photo[R][8][2],
 * J_idepth[R][8], J_C[R][8][4]; the side effects state_NewState / state_NewEnergy / state_NewEnergyWithOutlier,
 * CenterProjectedTo, point->idepth_hessian, and the edge level (1 = setLevel(1) on an out-of-image pixel). */
int sdso_lba_edge_eval(sdso_ctx* ctx, const double* T_wh, const double* photo, const double* idepth, const double cam[4], const double* b0,
                       double* error8, double* J_xi, double* J_photo, double* J_idepth, double* J_C, int* newState, double* newEnergy,
                       double* newEnergyWithOutlier, float* center3, float* idepth_hessian, int* level);

/* FullSystem::optimize, g2o body (FullSystemOptimize.cpp:404-868): LM over the graph of E2 edges of the uploaded window's active
 * residuals with one marginalised inverse-depth vertex PER RESIDUAL (:493-512), the restated g2o Levenberg-Marquardt (lambda_init
 * 0.1, additive damping, gain-ratio accept / reject, <= 10 trials, gain-threshold terminate action) and the Schur complement
 * over the idepth vertices. In: cam[4] (VertexCamDSO), T_wh[n][12] (host PRE_camToWorld), photo[n][2] (host aff_g2l), idepth[R].
 * Out: the same arrays hold the optimised estimates (the write-back of :676-727 is the caller's: a point takes the idepth of its
 * LAST active residual); used_host[n]; final robust chi2; per residual state_NewState / CenterProjectedTo / idepth_hessian.
 * mnumOptIts is overridden as the reference does (10 / 7 / 3 for 2 / 3 / >= 4 frames). */
int sdso_lba_g2o(sdso_ctx* ctx, int mnumOptIts, double cam[4], double* T_wh, double* photo, double* idepth, int* used_host, double* chi2_out,
                 int* newState, float* center3, float* idepth_hessian, int* iterations_out, int* trials_out);

/* ---- V1-V5, E3: the g2o vertices and the trace edge as operators over SoA batches (FullSystem/dso_g2o_vertex.cpp, dso_g2o_edge.cpp) ----
 * oplusImpl of n vertices of one kind in one launch; estimate is updated in place (caller-owned host arrays):
 *   SDSO_VERTEX_SE3_POSE      VertexSE3PoseDSO::oplusImpl      (:15-18)    estimate[n][12] row-major 3x4, update[n][6] = [upsilon; omega]: exp(update) * T
 *   SDSO_VERTEX_PHOTOMETRIC   VertexPhotometricDSO::oplusImpl  (:30-40)    estimate[n][2] = {a, b} += update[n][2]
 *   SDSO_VERTEX_INVERSE_DEPTH VertexInverseDepthDSO::oplusImpl (:56-58)    estimate[n] += update[n]
 *   SDSO_VERTEX_UV            VertexUVDSO::oplusImpl           (:73-88)    estimate[n][2] += clamp(update[n], +-0.5; non-finite -> 0) * aux[n][2] (SetDxDy)
 *   SDSO_VERTEX_CAM           VertexCamDSO::oplusImpl          (:100-106)  estimate[n][4] = {fx, fy, cx, cy} += update[n][4] */
#define SDSO_VERTEX_SE3_POSE 1
#define SDSO_VERTEX_PHOTOMETRIC 2
#define SDSO_VERTEX_INVERSE_DEPTH 3
#define SDSO_VERTEX_UV 4
#define SDSO_VERTEX_CAM 5
int sdso_vertex_oplus(sdso_ctx* ctx, int kind, int n, double* estimate, const double* update, const double* aux /* VertexUVDSO only */);
/* EdgeTracePointUVDSO::computeError + linearizeOplus (dso_g2o_edge.cpp:571-619) for n edges on level 0 of `frame`: VertexUVDSO
 * estimates uv[n][2], rotatePattern[n][2], _measurement[n], affLL[2], (dx_, dy_)[n][2]. error[n] / J[n] are read AND written: an edge
 * whose uv fails util::CheckBoundary gets error 0 and keeps its Jacobian, a non-finite intensity keeps both (g2o leaves the members
 * untouched). flag[n] (nullable): 1 evaluated, 0 outside, 2 non-finite. */
int sdso_edge_trace_uv_eval(sdso_ctx* ctx, int frame, int n, const double* uv, const float* rotatePattern, const double* measurement,
                            const float affLL[2], const double* dxdy, double* error, double* J, int* flag);

/* ---- point-sharded windowed BA over 2/4/8 GPUs ------------------------------------------------------
 * Every rank holds all keyframe pyramids and a contiguous block of the allPoints order with its residuals. Per LM
 * iteration: sdso_ba_linearize_all (local) -> sdso_ba_assemble (local partial damped system; priors and HM on rank 0)
 * -> sdso_ba_allreduce (ONE NCCL allreduce of (4+8n)^2+(4+8n)+1 doubles over NVLink) -> sdso_ba_solve_assembled (every rank
 * solves redundantly and back-substitutes its own points). The reference's only reduce is the per-thread accumulator sum
 * of stitchDoubleInternal (AccumulatedTopHessian.cpp:299-308); this is its multi-GPU analogue. */
int sdso_shard_range(int npoints, int rank, int nranks, int* begin, int* end);  /* contiguous block of allPoints owned by rank */
int sdso_ba_set_shard(sdso_ctx* ctx, int rank, int nranks);
int sdso_ba_assemble(sdso_ctx* ctx, void** device_system /* nullable */, int* count /* nullable */);
int sdso_ba_allreduce(sdso_ctx* ctx, double* energy_out /* nullable: summed linearisation energy (synchronises) */);
int sdso_ba_solve_assembled(sdso_ctx* ctx, int iteration, double* x, double* Hfinal, double* bfinal);
int sdso_nccl_unique_id(unsigned char id[128]);                 /* ncclGetUniqueId, to be broadcast by the host's rendezvous */
int sdso_nccl_init(sdso_ctx* ctx, int rank, int nranks, const unsigned char id[128]);
int sdso_nccl_destroy(sdso_ctx* ctx);
int sdso_allreduce_f64(sdso_ctx* ctx, void* device_buffer, int count);  /* in-place sum on the context's stream */
/* The same exchange as ONE hand-written kernel over NVLink peer memory instead of NCCL (57 KB per iteration: NCCL's launch and
 * protocol latency is the whole cost — 99 us on 8 GPUs): every rank allocates an exchange block (sdso_peer_alloc returns its
 * 64-byte CUDA IPC handle), the host's rendezvous gathers the handles, sdso_peer_connect maps the peers' blocks; from then on
 * sdso_ba_allreduce / sdso_allreduce_f64 push the local part into every peer's block (flag-carrying words) and reduce in one launch (csrc/collective.cu). One process per
 * GPU on one node; all ranks must issue the same sequence of exchanges. */
int sdso_peer_alloc(sdso_ctx* ctx, int nranks, int max_doubles, unsigned char handle_out[64]);
int sdso_peer_connect(sdso_ctx* ctx, int rank, int nranks, const unsigned char* handles /* nranks x 64 bytes in rank order */);
int sdso_peer_select(sdso_ctx* ctx, int which /* 0 = NCCL, 1 = peer-memory kernel (default once connected) */);
int sdso_peer_status(sdso_ctx* ctx, int* timed_out /* 1: a wait on a peer gave up; synchronises */);

/* ---- D1-D3, E3: immature points — constructor, temporal and static-stereo epipolar search ------------------------
 * One record per ImmaturePoint (FullSystem/ImmaturePoint.h:59-114): the fields the constructor, traceOn and
 * traceStereo read or write. Records are caller-owned host memory, updated in place. */
typedef struct sdso_immature_point {
  float u, v;
  float idepth_min, idepth_max;
  float quality, energyTH;
  float color[8], weights[8], gradH[4];                  /* gradH row-major 2x2 */
  float u_stereo, v_stereo, idepth_min_stereo, idepth_max_stereo, idepth_stereo;
  float lastTraceUV[2], lastTracePixelInterval;
  int32_t lastTraceStatus;                               /* ImmaturePointStatus (ImmaturePoint.h:50-56): 0 GOOD 1 OOB 2 OUTLIER 3 SKIPPED 4 BADCONDITION 5 UNINITIALIZED */
  int32_t bestIdx, numSteps;                             /* diagnostics of the last discrete search (-1 / 0 if it did not run) */
} sdso_immature_point;
/* ImmaturePoint::ImmaturePoint(u, v, host, ...) (ImmaturePoint.cpp:33-88) for n pixel positions uv[n][2]; ok[i] = 0 where the
 * constructor bailed out on a non-finite colour (energyTH = NaN). Also sets u_stereo=u, v_stereo=v, idepth_*_stereo = (0, NaN)
 * as every caller does (FullSystem.cpp:582-585). */
int sdso_immature_init(sdso_ctx* ctx, int host_frame, int n, const float* uv, sdso_immature_point* out, int* ok /* nullable */);
/* ImmaturePoint::traceOn(frame, hostToFrame_KRKi, hostToFrame_Kt, hostToFrame_affine, HCalib) (ImmaturePoint.cpp:459-828) */
int sdso_trace_on(sdso_ctx* ctx, int frame, const float KRKi[9], const float Kt[3], const float aff[2], int n,
                  sdso_immature_point* pts, int* status /* nullable */);
/* ImmaturePoint::traceStereo(frame, K, mode_right) (ImmaturePoint.cpp:94-451); baseline = the context's */
int sdso_trace_stereo(sdso_ctx* ctx, int frame, const float K[9], int mode_right, int n, sdso_immature_point* pts, int* status /* nullable */);
/* The loop of FullSystem::traceNewCoarse over ALL key frames of the window (FullSystem.cpp:745-781) in ONE launch: n points, point i
 * hosted in key frame host_of_point[i] < n_hosts, whose hostToFrame_KRKi / Kt / affine are KRKi[h][9], Kt[h][3], aff[h][2].
 * pts == NULL: the device-resident records (sdso_immature_upload) are traced in place and nothing but status[] (nullable) comes back —
 * the records of a window's immature points then cross the host link once per key frame instead of twice per tracked frame. */
int sdso_trace_on_hosts(sdso_ctx* ctx, int frame, int n_hosts, const float* KRKi, const float* Kt, const float* aff, int n, const int* host_of_point,
                        sdso_immature_point* pts /* nullable */, int* status /* nullable */);
/* traceStereo of the first n device-resident records in place */
int sdso_trace_stereo_resident(sdso_ctx* ctx, int frame, const float K[9], int mode_right, int n, int* status /* nullable */);
/* device-resident immature-point records: upload replaces the pool, download reads records [first, first + n) back */
int sdso_immature_upload(sdso_ctx* ctx, int n, const sdso_immature_point* pts);
int sdso_immature_download(sdso_ctx* ctx, int first, int n, sdso_immature_point* pts);
/* D4: FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:52-238) + ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:886-985)
 * for n candidates, host[i] = index of the candidate's host frame in the window uploaded with sdso_ba_* (its FrameFramePrecalc
 * and calibration are used). variant SDSO_VARIANT_SSE: the original body (Hdd/bd, 3 LM iterations, Hdd >= setting_minIdepthH_act);
 * SDSO_VARIANT_G2O: the live code, whose activation edge projects once (the inverse depth stays 0.5 (idepth_min + idepth_max)).
 * result[i]: 1 activated, 0 not well constrained (returns 0 in the reference), -1 outlier / invalid ((PointHessian*)-1);
 * idepth[i]; states[i][nframes] = ResState per target (-1 for the host); energy[i] = lastEnergy. */
int sdso_activate_points(sdso_ctx* ctx, int n, const int* host, const sdso_immature_point* pts, int variant, int min_obs, int* result,
                         float* idepth, int* states, float* energy);

/* ---- candidate pixel selection (FullSystem/PixelSelector2.{h,cpp}) ----------------------------------------------------
 * One PixelSelector per context (FullSystem holds one, FullSystem.cpp:168). State as in the reference: currentPotential
 * (starts at 3), the frame the block thresholds were last made for (gradHistFrame), and randomPattern — glibc's
 * srand(3141592); rand() & 0xFF sequence (PixelSelector2.cpp:42-44), generated by the library without touching the
 * process-wide rand() state. Every result is integer / index work and bit-exact. */
/* new PixelSelector(w, h): currentPotential = 3, gradHistFrame = 0 */
int sdso_selector_reset(sdso_ctx* ctx);
/* The generator behind randomPattern: out[i] = i-th value of srand(seed); rand() & 0xFF. Host only, needs no context. */
void sdso_selector_pattern_host(unsigned seed, unsigned char* out, size_t n);
/* randomPattern[0..w*h) (for checks against the platform's rand()) */
int sdso_selector_random_pattern(sdso_ctx* ctx, unsigned char* out);
/* currentPotential; set > 0 overwrites it first */
int sdso_selector_potential(sdso_ctx* ctx, int set, int* potential);
/* PixelSelector::makeHists(fh) (PixelSelector2.cpp:84-178): 32x32-block gradient histograms -> ths, thsSmoothed ((w/32)*(h/32) floats each) */
int sdso_selector_make_hists(sdso_ctx* ctx, int frame, float* ths /* nullable */, float* ths_smoothed /* nullable */);
/* PixelSelector::select(fh, map_out, pot, thFactor) (PixelSelector2.cpp:340-536); thresholds = those of the last makeHists.
 * map_out: w*h floats in {0,1,2,4} (nullable: the map stays on the device); n = {n2, n3, n4}. */
int sdso_selector_select(sdso_ctx* ctx, int frame, int pot, float thFactor, float* map_out, int n[3]);
/* PixelSelector::makeMaps(fh, map_out, density, recursionsLeft, plot=false, thFactor) (PixelSelector2.cpp:192-327): select at
 * currentPotential, re-select once per recursion when the count is off by > 25 % / < 1/4, random sub-sampling to the wanted
 * density, currentPotential update. Returns numHaveSub in *num_selected. */
int sdso_make_maps(sdso_ctx* ctx, int frame, float density, int recursionsLeft, float thFactor, float* map_out /* nullable */, int* num_selected);
/* The selected pixels of the last select / makeMaps in raster order, as FullSystem::makeNewTraces walks selectionMap
 * (FullSystem.cpp:1609-1621): uv[2i] = x, uv[2i+1] = y, type[i] = map value (1, 2 or 4). *n = count (<= max_n else SDSO_E_INVALID). */
int sdso_selector_points(sdso_ctx* ctx, int max_n, float* uv, float* type, int* n);

/* ---- coarse distance map and the activation candidate filter --------------------------------------------------------
 * CoarseDistanceMap (FullSystem/CoarseTracker.cpp:1186-1420) lives on the level-1 grid (w >> 1, h >> 1). Hosts are the
 * keyframes other than the newest one, in frameHessians order; per host the caller passes
 * KRKi = K[1] * R(newest <- host) * Ki[0] and Kt = K[1] * t (floats, CoarseTracker.cpp:1233-1235 / FullSystem.cpp:845-847). */
/* makeDistanceMap(frameHessians, newest) (:1216-1253) + growDistBFS (:1258-1355): pt_uvid[i] = {u, v, idepth_scaled} of the
 * active points grouped by host (pt_host[i] = host index). map_out: (w >> 1) * (h >> 1) floats (0..39, 1000), nullable. */
int sdso_distmap_make(sdso_ctx* ctx, int n_hosts, const float* KRKi, const float* Kt, int n_pts, const int* pt_host, const float* pt_uvid, float* map_out);
/* addIntoDistFinal(u, v) (:1358-1366) for n cells; the field of the interior cells does not depend on the insertion order. */
int sdso_distmap_add(sdso_ctx* ctx, int n, const int* uv, float* map_out);
/* FullSystem::activatePointsMT STEP2 (FullSystem.cpp:838-901) over the candidates in the reference's order (host by host,
 * immaturePoints order): cand_host[i], pts[i], my_type[i] (1, 2 or 4), host_flagged[h] = flaggedForMarginalization.
 * verdict[i]: 0 stays immature, 1 goes to toOptimize (and was inserted into the distance field), 2 deleted.
 * Needs the field of sdso_distmap_make; leaves the field with the accepted candidates inserted. *rounds (nullable) =
 * dependency rounds the device needed. */
int sdso_activation_filter(sdso_ctx* ctx, int n_hosts, const float* KRKi, const float* Kt, const unsigned char* host_flagged, int n, const int* cand_host,
                           const sdso_immature_point* pts, const float* my_type, float currentMinActDist, int* verdict, int* rounds, float* map_out);

/* ---- input preparation and trajectory rows (the wire formats either side of the path) ---------------------------------
 * Undistort::undistort<unsigned char> (util/Undistort.cpp:398-489) = PhotometricUndistorter::processFrame (:222-260) + the
 * bilinear remap, with the benchmark noise switched off as in the reference's defaults. The remap tables (w*h floats in
 * raw-image pixels, negative x = outside) and the photometric calibration (G: 256 floats or NULL = no response calibration;
 * vignetteMapInv: wOrg*hOrg floats or NULL) come from the caller's Undistort object, which computes them once per run. */
int sdso_undistort_setup(sdso_ctx* ctx, int wOrg, int hOrg, const float* remapX, const float* remapY, const float* G, const float* vignetteMapInv,
                         int photometricCalibration /* setting_photometricCalibration 0..2 */, int useExposure /* setting_useExposure */);
/* raw: wOrg*hOrg bytes (host). out_image: w*h floats (nullable). frame >= 0: FrameHessian::makeImages runs on the rectified
 * image without leaving the device (ab_exposure = the returned exposure). *exposure_out = ImageAndExposure::exposure_time. */
int sdso_undistort(sdso_ctx* ctx, const unsigned char* raw, float exposure, float factor, float* out_image, int frame, int use_hcalib, float* exposure_out);
/* One row of FullSystem::printResult (FullSystem.cpp:236-285): camToWorld as "R00 R01 R02 t0 R10 ... t2\n" with 15 significant
 * digits. Host only. Returns the length written (without the terminator) or SDSO_E_INVALID if buf is too small. */
int sdso_trajectory_row(const double camToWorld[12], char* buf, int n);

#ifdef __cplusplus
}
#endif
#endif /* SDSO_B200_H_ */
