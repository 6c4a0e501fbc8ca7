// A3–A7, E1 — CoarseTracker on the device (FullSystem/CoarseTracker.cpp, dso_g2o_edge.cpp:395-500).
//
// B200 design: the WHOLE coarse-to-fine optimisation of one (reference, new frame, initial pose)
// problem — every calcRes / calcGSSSE pass, the 8x8 solves, the SE3 updates, the LM accept/reject
// logic for all pyramid levels — runs inside ONE persistent kernel launch. One thread-block cluster
// owns one problem; its CTAs split the points, reduce per-CTA partial sums in shared memory, exchange
// them through distributed shared memory (one cluster barrier per evaluation, double-buffered) and
// every CTA redundantly runs the (deterministic, fixed-order) final sum + solve, so no broadcast and
// no host round trip is needed. Independent problems (motion hypotheses, FullSystem.cpp:351-376, or
// independent sequences) are additional clusters of the same launch.
//
// Per point ("eval"): 16 B point record + 4 x 16 B texel gather, projection, residual, Huber weight,
// 9-vector Jacobian row, 45-term weighted outer product — calcRes (:600-792) and calcGSSSE (:537-596)
// fused, so the eight buf_warped_* arrays never exist in memory.
#include "ctx.h"
#include "tracker_state.h"
#include <cooperative_groups.h>
#include <cstring>
#include <cmath>

namespace cg = cooperative_groups;

namespace sdso {

// ------------------------------------------------------------------------------------------------
constexpr int kAcc = 52;      // accumulators per thread (see enum)
enum { A_H = 0, A_E = 45, A_NE = 46, A_NSAT = 47, A_ST = 48, A_SRT = 49, A_SN = 50, A_NW = 51 };

struct TrackLevel {
  const float4* pc;  // {u, v, idepth, color}  (CoarseTracker.h:117-121 pc_u/pc_v/pc_idepth/pc_color fused)
  int n;
  int w, h;
  float fx, fy, cx, cy;  // tracker's K (optimised HCalib), CoarseTracker.cpp:108-136
  float Ki[9];
  double gfx, gfy, gcx, gcy;  // global initial KG[lvl] used by EdgeSE3PosePhotoDSO (dso_util.hpp:10-22)
};

struct TrackProblem {  // one per cluster, in global memory
  // inputs
  double T[12];
  double aff[2];
  double minResForAbort[5];
  const float4* tex[kPyrLevels];  // new frame pyramid
  float exposure_new;
  float ref_exposure;             // per-problem reference keyframe: independent sequences of one launch track against their own templates
  const float4* pc[kPyrLevels];   // {u, v, idepth, color} per level of that reference (CoarseTracker.h:117-121 fused)
  int pc_n[kPyrLevels];
  double ref_aff[2];              // lastRef_aff_g2l
  // outputs
  double T_out[12];
  double aff_out[2];
  double lastResiduals[5];
  double flow[3];
  int iterations[5];
  int ok;
  unsigned long long evals;
  // single-eval mode outputs
  double rs[6];
  double H[64];
  double b[8];
  int warped_n;
  int pad1;
  long long cyc[16];  // optional phase timing (clock64 of rank 0 / thread 0), see TrackParams::timing
};

struct TrackParams {
  TrackLevel L[kPyrLevels];
  int levels;
  int coarsest;
  int variant;
  int mode;  // 0 = full track, 1 = single fused calcRes+calcGS at level eval_lvl
  int eval_lvl;
  float eval_cutoff;
  float* dump;          // mode 1: per-point records [9][n] (valid,idepth,u,v,dx,dy,residual,weight,refColor) or null
  double* dump_d;       // mode 2: per-point records [10][n] (flag, err, J[8]) or null
  float ref_exposure;
  double ref_aff[2];    // lastRef_aff_g2l
  float huberTH, coarseCutoffTH;
  float affineOptModeA, affineOptModeB;
  int g2o_stop_persists;
  int timing;
  TrackProblem* problems;
  int nb;                  // problems in this launch
  int dynamic;             // 1: CTAs pull problems from work_counter (cluster size 1 only)
  int use_cache;           // 1: cache_bytes(blockDim.x) of dynamic shared memory follow TrackSmem (texel / point cache of the SSE evaluation)
  unsigned int* work_counter;
  // g2o variant scratch: per level edge flags/errors for each problem
  unsigned char* edge_flag[kPyrLevels];  // [problem][n_l]
  double* edge_err[kPyrLevels];          // [problem][n_l]
  int edge_stride[kPyrLevels];
};

struct EvalConst {  // per-evaluation constants, computed by thread 0 of each CTA
  float RKi[9];
  float t[3];
  float affLL[2];
  float a;       // (float) fromToVecExposure(...)[0]      (calcGSSSE :544)
  float b0;      // lastRef_aff_g2l.b
  float cutoff, maxEnergy;
  // g2o variant: pose in double and photo ab
  double R[9], tt[3];
  float ab[2];
  double b0d;
};

struct LMState {  // lives in shared memory of every CTA (identical content everywhere)
  double R[9], t[3];      // refToNew_current
  double aff[2];
  double Rn[9], tn[3];    // trial
  double affn[2];
  double H[64], b[8];
  double inc[8];
  double total[kAcc];     // totals in double (g2o path, operator-level outputs)
  float totf[kAcc];       // totals of the accepted state (SSE path; float: the block partials are float anyway)
  float lambda;
  int small;              // !(inc.norm() > 1e-3) of the last LM step (:1019)
};

// ---- small double helpers (device) -------------------------------------------------------------
__device__ __forceinline__ void d_mat3_mul(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// 1/x to ~1e-14 relative without the IEEE division sequence: float seed + one Newton step in double.
// Falls back to the division outside the float range.
__device__ __forceinline__ double fast_rcp(double x) {
  const double ax = fabs(x);
  if (!(ax > 1e-30 && ax < 1e30)) return 1.0 / x;
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)x));  // 1 ulp seed; an IEEE float division costs ~85 cycles here
  double r = (double)rf;
  r = fma(r, fma(-x, r, 1.0), r);  // 1e-7 -> 1e-14: far below the float precision of the accumulated H
  return r;
}

// closed-form coefficients for |theta| >= 0.25 (rare on this path): kept out of line to bound code size
__device__ __noinline__ void se3_coeffs_general(double th2, double* A, double* B, double* C) {
  const double th = sqrt(th2);
  double sn, cs;
  sincos(th, &sn, &cs);
  *A = sn / th; *B = (1.0 - cs) / th2; *C = (th - sn) / (th2 * th);
}

// exp(a) * T   with a = [upsilon; omega]   (thirdparty/Sophus/sophus/se3.hpp:407-428; left-multiplicative
// update as CoarseTracker.cpp:978 / dso_g2o_vertex.cpp:17). R = I + A W + B W^2, V = I + B W + C W^2 with
// A = sin(t)/t, B = (1-cos t)/t^2, C = (t-sin t)/t^3; for t < 0.25 the three are evaluated by their Taylor
// series in t^2 (truncation < 1e-20), which needs no sqrt / sincos / division. Fully unrolled: registers only.
__device__ __forceinline__ void d_se3_exp_mul(const double (&a)[8], const double* Rin, const double* tin, double (&Ro)[9], double (&to)[3]) {
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = Rin[i];
#pragma unroll
  for (int i = 0; i < 3; i++) t[i] = tin[i];
  const double wx = a[3], wy = a[4], wz = a[5];
  const double x = wx * wx + wy * wy + wz * wz;
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  d_mat3_mul(O, O, O2);
  double A, B, C;
  if (x < 0.0625) {
    // sum_k (-1)^k x^k / (2k+1)!, /(2k+2)!, /(2k+3)!   (Horner, k = 7..0)
    A = -1.0 / 1307674368000.0; B = -1.0 / 20922789888000.0; C = -1.0 / 355687428096000.0;
    A = fma(A, x, 1.0 / 6227020800.0);  B = fma(B, x, 1.0 / 87178291200.0);  C = fma(C, x, 1.0 / 1307674368000.0);
    A = fma(A, x, -1.0 / 39916800.0);   B = fma(B, x, -1.0 / 479001600.0);   C = fma(C, x, -1.0 / 6227020800.0);
    A = fma(A, x, 1.0 / 362880.0);      B = fma(B, x, 1.0 / 3628800.0);      C = fma(C, x, 1.0 / 39916800.0);
    A = fma(A, x, -1.0 / 5040.0);       B = fma(B, x, -1.0 / 40320.0);       C = fma(C, x, -1.0 / 362880.0);
    A = fma(A, x, 1.0 / 120.0);         B = fma(B, x, 1.0 / 720.0);          C = fma(C, x, 1.0 / 5040.0);
    A = fma(A, x, -1.0 / 6.0);          B = fma(B, x, -1.0 / 24.0);          C = fma(C, x, -1.0 / 120.0);
    A = fma(A, x, 1.0);                 B = fma(B, x, 0.5);                  C = fma(C, x, 1.0 / 6.0);
  } else {
    se3_coeffs_general(x, &A, &B, &C);
  }
  double Re[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    Re[i] = I + A * O[i] + B * O2[i];
    V[i] = I + B * O[i] + C * O2[i];
  }
  double te[3];
#pragma unroll
  for (int r = 0; r < 3; r++) te[r] = V[r * 3] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  d_mat3_mul(Re, R, Ro);
#pragma unroll
  for (int r = 0; r < 3; r++) to[r] = Re[r * 3] * t[0] + Re[r * 3 + 1] * t[1] + Re[r * 3 + 2] * t[2] + te[r];
}

// LDLT solve (no pivoting; the systems on this path are SPD after damping). Returns false on breakdown.
template <int NMAX>
__device__ bool d_ldlt_solve(int n, const double* A, int lda, const double* b, double* x) {
  double L[NMAX * NMAX], d[NMAX];
  for (int j = 0; j < n; j++) {
    double s = A[j * lda + j];
    for (int k = 0; k < j; k++) s -= L[j * NMAX + k] * L[j * NMAX + k] * d[k];
    d[j] = s;
    double inv = 1.0 / s;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * lda + j];
      for (int k = 0; k < j; k++) v -= L[i * NMAX + k] * L[j * NMAX + k] * d[k];
      L[i * NMAX + j] = v * inv;
    }
  }
  double y[NMAX];
  for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * NMAX + k] * y[k]; y[i] = v; }
  for (int i = 0; i < n; i++) y[i] /= d[i];
  for (int i = n - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < n; k++) v -= L[k * NMAX + i] * x[k]; x[i] = v; }
  bool ok = true;
  for (int i = 0; i < n; i++) ok = ok && isfinite(x[i]);
  return ok;
}

// AffLight::fromToVecExposure (util/NumType.h:159-170)
__device__ __forceinline__ void d_aff_from_to(float expF, float expT, double aF, double bF, double aT, double bT, double out[2]) {
  if (expF == 0 || expT == 0) { expT = expF = 1; }
  double a = exp(aT - aF) * expT / expF;
  out[0] = a;
  out[1] = bT - a * bF;
}

// ---- shared memory layout ------------------------------------------------------------------------
constexpr int kAccPad = 64;    // accumulator registers per thread (kAcc used, padded to a power of two)
constexpr int kMaxCluster = 16;
constexpr int kMaxWarps = 8;   // 256 threads per CTA

struct __align__(16) TrackSmem {
  LMState lm;
  EvalConst ec;
  float wpart[kMaxWarps][kAccPad];           // per-warp sums of this CTA
  float gather[2][kMaxCluster][kAccPad];     // [parity][source CTA][k]: every CTA PUSHES its sums into every peer
  double gather_d[2][kMaxCluster];           // same for the one channel that is summed in double
  double warp_d[kMaxWarps];
  double total_d;
  unsigned long long bar[2];                 // mbarriers of the two exchange buffers
};

// Per-CTA cache of what the LM iterations of one level keep re-reading (dynamic shared memory behind TrackSmem): the point
// records of the level and, per point, the 2x2 texel patch {I, dx, dy} of the last evaluation with its integer position as tag.
// Between two iterations of a level the pose moves by a fraction of a pixel, so most patches are the same texels again; without
// the cache every one of those re-reads is four scattered 32-byte sectors that, with ~300 sequences in flight, no longer fit L2.
// Slot (m, tid) belongs to thread tid alone (its m-th point), so no synchronisation is involved.
constexpr int kCacheRounds = 4;                  // points per thread that are cached (the rest gathers from global memory)
struct TexCache {
  float* tex;      // [12][slots]: t00.xyz, t10.xyz, t01.xyz, t11.xyz, component-major (conflict-free per warp)
  int* tag;        // [slots]: (lvl << 28) | (iy << 14) | ix, -1 = empty
  float4* pc;      // [slots] point records of level *pc_lvl
  int* pc_lvl;     // level whose point records are cached (-1 = none); set by thread 0 behind the evaluation's barrier
  int slots;       // kCacheRounds * kBT: the capacity the kernel instance was compiled for (>= kCacheRounds * blockDim.x)
};
__host__ __device__ constexpr size_t cache_bytes(int block_threads) {
  return (size_t)kCacheRounds * block_threads * (12 * sizeof(float) + sizeof(int) + sizeof(float4)) + 16;
}

struct PhaseTimer {  // phase breakdown of the persistent kernel, enabled by TrackParams::timing
  long long last = 0;
  long long* cyc = nullptr;
  __device__ __forceinline__ void start(long long* c) { cyc = c; last = clock64(); }
  __device__ __forceinline__ void tick(int k) {
    if (cyc) { long long now = clock64(); cyc[k] += now - last; last = now; }
  }
};

// Register-transposing warp reduction: 64 values x 32 lanes -> lane L ends with the warp sums of
// indices 2L and 2L+1 in a[0], a[1]. 62 shuffles instead of 64 x 5, no shared memory.
#define SDSO_RSTAGE(O, N)                                                  \
  _Pragma("unroll") for (int j = 0; j < N; j++) {                          \
    const bool up = (lane & O) != 0;                                       \
    const float keep = up ? a[j + N] : a[j];                               \
    const float send = up ? a[j] : a[j + N];                               \
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, O);                   \
  }
__device__ __forceinline__ void warp_reduce_transpose64(float (&a)[kAccPad], int lane) {
  SDSO_RSTAGE(16, 32)
  SDSO_RSTAGE(8, 16)
  SDSO_RSTAGE(4, 8)
  SDSO_RSTAGE(2, 4)
  SDSO_RSTAGE(1, 2)
}
#undef SDSO_RSTAGE

// ---- DSMEM exchange primitives (inline PTX; sm_90+ cluster features) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// asynchronous remote shared-memory store that signals the destination CTA's mbarrier with the byte count
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_b64(uint32_t raddr, unsigned long long v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(v), "r"(rbar) : "memory");
}

struct Exchange {  // per-thread bookkeeping of the double-buffered exchange
  int parity = 0;         // which buffer the next exchange uses
  unsigned phases = 0;    // bit p = phase parity the next wait on bar[p] expects
};

// Block reduction + all-to-all exchange inside the cluster WITHOUT a cluster barrier: every CTA pushes its 64
// block sums (+ one double) into every peer with st.async, each destination's mbarrier counts the bytes, and
// every thread waits on its own CTA's mbarrier only (acquire at CTA scope: L1 stays warm, no CCTL.IVALL).
// Returns the buffer index p; sm->gather[p][r][k], r < C, is then valid. Double-buffered: a CTA can start
// pushing exchange n+2 only after every peer has pushed n+1, i.e. after every peer finished reading n.
template <bool WITH_D>
__device__ __forceinline__ int reduce_exchange(float (&acc)[kAccPad], TrackSmem* sm, Exchange& ex, unsigned C, unsigned rank, double dacc,
                                               PhaseTimer* tm = nullptr) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int p = ex.parity;
  warp_reduce_transpose64(acc, lane);
  *reinterpret_cast<float2*>(&sm->wpart[warp][2 * lane]) = make_float2(acc[0], acc[1]);
  if (WITH_D) {  // double-precision ops are slow on this part: only the g2o path pays for the extra channel
    double dv = warp_sum(dacc);
    if (lane == 0) sm->warp_d[warp] = dv;
  }
  __syncthreads();
  if (tid < kAccPad) {
    float s = 0.f;
    for (int w = 0; w < nw; w++) s += sm->wpart[w][tid];
    const uint32_t dst = smem_u32(&sm->gather[p][rank][tid]), bar = smem_u32(&sm->bar[p]);
    for (unsigned r = 0; r < C; r++) st_async_b32(mapa_u32(dst, r), __float_as_uint(s), mapa_u32(bar, r));
  } else if (WITH_D && tid == kAccPad) {
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += sm->warp_d[w];
    const uint32_t dst = smem_u32(&sm->gather_d[p][rank]), bar = smem_u32(&sm->bar[p]);
    for (unsigned r = 0; r < C; r++) st_async_b64(mapa_u32(dst, r), (unsigned long long)__double_as_longlong(s), mapa_u32(bar, r));
  } else if (tid == kAccPad + 32) {
    mbar_arrive_expect_tx(&sm->bar[p], C * (kAccPad * 4 + (WITH_D ? 8 : 0)));
  }
  if (tm) tm->tick(2);
  mbar_wait(&sm->bar[p], (ex.phases >> p) & 1u);
  ex.phases ^= (1u << p);
  ex.parity ^= 1;
  if (tm) tm->tick(3);
  return p;
}

__device__ __forceinline__ double gather_sum(const TrackSmem* sm, int p, unsigned C, int k) {
  double s = 0.0;
#pragma unroll 1
  for (unsigned r = 0; r < C; r++) s += (double)sm->gather[p][r][k];  // fixed order
  return s;
}
__device__ __forceinline__ float gather_sumf(const TrackSmem* sm, int p, unsigned C, int k) {
  float s0 = 0.f, s1 = 0.f;  // two interleaved chains, fixed order: even ranks, odd ranks
#pragma unroll 1
  for (unsigned r = 0; r + 1 < C; r += 2) { s0 += sm->gather[p][r][k]; s1 += sm->gather[p][r + 1][k]; }
  if (C & 1) s0 += sm->gather[p][C - 1][k];
  return s0 + s1;
}
__device__ __forceinline__ double gather_sum_d(const TrackSmem* sm, int p, unsigned C) {
  double s = 0.0;
#pragma unroll 1
  for (unsigned r = 0; r < C; r++) s += sm->gather_d[p][r];
  return s;
}

// Convenience form: exchange, then the totals (bit-identical in every CTA: fixed summation order) -> out[kAcc]
// in shared memory, visible to the whole CTA on return.
__device__ void reduce_all(float (&acc)[kAccPad], TrackSmem* sm, Exchange& ex, cg::cluster_group& cluster, double* out,
                           double dacc = 0.0, double* dout = nullptr, PhaseTimer* tm = nullptr) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const int p = reduce_exchange<true>(acc, sm, ex, C, rank, dacc, tm);
  const int tid = threadIdx.x;
  if (tid < kAcc) out[tid] = gather_sum(sm, p, C, tid);
  if (dout) *dout = gather_sum_d(sm, p, C);
  __syncthreads();
}

// 8x8 SPD solve by one warp: lane l (mod 8) owns row l of the augmented matrix [A | rhs] in registers.
// Elimination without pivoting (the damped systems on this path are SPD); all lanes end with x[0..7].
// pd = every pivot was finite and > 0 (what a Cholesky factorisation needs to succeed).
__device__ __forceinline__ void warp_solve8(double (&row)[9], int lane, double (&x)[8], bool& pd) {
  const int l = lane & 7;
  double pinv[8];
  pd = true;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double pk[9];
#pragma unroll
    for (int j = k; j < 9; j++) pk[j] = __shfl_sync(0xffffffffu, row[j], k);
    pd = pd && (pk[k] > 0) && isfinite(pk[k]);
    const double inv = fast_rcp(pk[k]);
    pinv[k] = inv;
    if (l > k) {
      const double f = row[k] * inv;
#pragma unroll
      for (int j = k + 1; j < 9; j++) row[j] -= f * pk[j];
    }
  }
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const double xk = __shfl_sync(0xffffffffu, row[8] * pinv[k], k);
    x[k] = xk;
    if (l < k) row[8] -= row[k] * xk;
  }
}

// ---- SSE-path evaluation: calcRes (:645-773) + calcGSSSE (:537-596) fused over this CTA's points ----
// Points are processed in batches of U per thread: all point records first, then all 4xU texel gathers,
// then the arithmetic, so up to 4U independent 16-byte loads are in flight per thread.
template <int U, int CS /* slots of the texel cache = component stride: a constant, so the 12 offsets are immediates */>
__device__ void eval_points_sse(const TrackParams& P, const TrackLevel& L, int lvl, const float4* __restrict__ tex,
                                const EvalConst& ec, float (&acc)[kAccPad], unsigned& evals, int gtid, int gthreads,
                                float* dump, const float4* __restrict__ pc, const int n, const TexCache tc) {
#pragma unroll
  for (int k = 0; k < kAccPad; k++) acc[k] = 0.f;
  const bool use_cache = tc.tex != nullptr;
  const bool pc_cached = use_cache && *tc.pc_lvl == lvl;
  const float fxl = L.fx, fyl = L.fy, cxl = L.cx, cyl = L.cy;
  const int wl = L.w, hl = L.h;
  const float huberTH = P.huberTH;
  float RKi[9], tt[3];
#pragma unroll
  for (int k = 0; k < 9; k++) RKi[k] = ec.RKi[k];
#pragma unroll
  for (int k = 0; k < 3; k++) tt[k] = ec.t[k];
  const float affLL0 = ec.affLL[0], affLL1 = ec.affLL[1], cutoff = ec.cutoff, maxEnergy = ec.maxEnergy, ea = ec.a, eb0 = ec.b0;

  // flow indicators (:662-693) of every 32nd point, level 0 only. A separate dense pass: inside the main loop exactly one
  // lane of each warp would take this branch and the warp would pay its ~150 instructions for every point
  if (lvl == 0) {
    for (int i = 32 * gtid; i < n; i += 32 * gthreads) {
      const float4 pp = __ldg(pc + i);
      const float x = pp.x, y = pp.y, id = pp.z;
      float pt[3], ptT[3], ptT2[3], pt3[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float kp = L.Ki[r * 3 + 0] * x + L.Ki[r * 3 + 1] * y + L.Ki[r * 3 + 2];
        const float rp = RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2];
        ptT[r] = kp + tt[r] * id;
        ptT2[r] = kp - tt[r] * id;
        pt[r] = rp + tt[r] * id;
        pt3[r] = rp - tt[r] * id;
      }
      const float u = pt[0] / pt[2], v = pt[1] / pt[2];
      const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
      const float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      const float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      const float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      const float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      const float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      const float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      acc[A_ST] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      acc[A_ST] += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      acc[A_SRT] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      acc[A_SRT] += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      acc[A_SN] += 2.f;
    }
  }
  for (int base = gtid, m0 = 0; base < n; base += gthreads * U, m0 += U) {
    float4 p[U];
#pragma unroll
    for (int q = 0; q < U; q++) {
      const int i = base + q * gthreads;
      const int m = m0 + q;                                      // this thread's m-th point
      const int slot = (use_cache && m < kCacheRounds) ? m * (CS / kCacheRounds) + (int)threadIdx.x : -1;
      if (i >= n) p[q] = make_float4(0.f, 0.f, 1.f, 0.f);
      else if (slot >= 0 && pc_cached) p[q] = tc.pc[slot];
      else { p[q] = __ldg(pc + i); if (slot >= 0) tc.pc[slot] = p[q]; }
    }
    float uu[U], vv[U], Kuu[U], Kvv[U], nid[U];
    bool inb[U];
    float4 t00[U], t10[U], t01[U], t11[U];
#pragma unroll
    for (int q = 0; q < U; q++) {
      const int i = base + q * gthreads;
      const float x = p[q].x, y = p[q].y, id = p[q].z;
      float pt[3];
#pragma unroll
      for (int r = 0; r < 3; r++) pt[r] = (RKi[r * 3 + 0] * x + RKi[r * 3 + 1] * y + RKi[r * 3 + 2]) + tt[r] * id;
      const float u = pt[0] / pt[2], v = pt[1] / pt[2];
      const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
      const float new_idepth = id / pt[2];
      uu[q] = u; vv[q] = v; Kuu[q] = Ku; Kvv[q] = Kv; nid[q] = new_idepth;
      if (i < n) {
        evals++;
      }
      inb[q] = (i < n) && (Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0);  // :696
      if (inb[q]) {
        const int ix = (int)Ku, iy = (int)Kv;
        const int m = m0 + q;
        const int slot = (use_cache && m < kCacheRounds) ? m * (CS / kCacheRounds) + (int)threadIdx.x : -1;
        const int tag = (lvl << 28) | (iy << 14) | ix;
        if (slot >= 0 && tc.tag[slot] == tag) {   // same texels as in the last evaluation of this point
          const float* c = tc.tex + slot;
          constexpr int cs = CS;
          t00[q] = make_float4(c[0 * cs], c[1 * cs], c[2 * cs], 0.f);
          t10[q] = make_float4(c[3 * cs], c[4 * cs], c[5 * cs], 0.f);
          t01[q] = make_float4(c[6 * cs], c[7 * cs], c[8 * cs], 0.f);
          t11[q] = make_float4(c[9 * cs], c[10 * cs], c[11 * cs], 0.f);
        } else {
          const float4* bp = tex + ix + iy * wl;
          t00[q] = __ldg(bp); t10[q] = __ldg(bp + 1); t01[q] = __ldg(bp + wl); t11[q] = __ldg(bp + 1 + wl);
          if (slot >= 0) {
            float* c = tc.tex + slot;
            constexpr int cs = CS;
            c[0 * cs] = t00[q].x; c[1 * cs] = t00[q].y; c[2 * cs] = t00[q].z;
            c[3 * cs] = t10[q].x; c[4 * cs] = t10[q].y; c[5 * cs] = t10[q].z;
            c[6 * cs] = t01[q].x; c[7 * cs] = t01[q].y; c[8 * cs] = t01[q].z;
            c[9 * cs] = t11[q].x; c[10 * cs] = t11[q].y; c[11 * cs] = t11[q].z;
            tc.tag[slot] = tag;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < U; q++) {
      const int i = base + q * gthreads;
      const float refColor = p[q].w;
      bool valid = false;
      float hity = 0.f, hitz = 0.f, residual = 0.f, hw = 0.f;
      if (inb[q]) {
        // getInterpolatedElement33 (util/globalFuncs.h:73-86), weights and sum order as written there
        const float Ku = Kuu[q], Kv = Kvv[q];
        const int ix = (int)Ku, iy = (int)Kv;
        const float dx = Ku - ix, dy = Kv - iy;
        const float dxdy = dx * dy;
        const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
        const float hitx = w11 * t11[q].x + w01 * t01[q].x + w10 * t10[q].x + w00 * t00[q].x;
        hity = w11 * t11[q].y + w01 * t01[q].y + w10 * t10[q].y + w00 * t00[q].y;
        hitz = w11 * t11[q].z + w01 * t01[q].z + w10 * t10[q].z + w00 * t00[q].z;
        if (isfinite(hitx)) {
          residual = hitx - (affLL0 * refColor + affLL1);
          const float ar = fabsf(residual);
          hw = ar < huberTH ? 1.f : huberTH / ar;
          if (ar > cutoff) {
            acc[A_E] += maxEnergy; acc[A_NE] += 1.f; acc[A_NSAT] += 1.f;
          } else {
            acc[A_E] += hw * residual * residual * (2 - hw);
            acc[A_NE] += 1.f; acc[A_NW] += 1.f;
            valid = true;
          }
        }
      }
      if (valid) {
        // calcGSSSE :553-581
        const float u = uu[q], v = vv[q], id = nid[q];
        const float dx = hity * fxl, dy = hitz * fyl;
        float J[9];
        // (the Jacobian only enters the accumulated H / b, which are held to 1e-4, not the bit-exact buffers: fused multiply-adds)
        const float uv = u * v;
        J[0] = id * dx;
        J[1] = id * dy;
        J[2] = -(id * __fmaf_rn(u, dx, v * dy));
        J[3] = -__fmaf_rn(uv, dx, dy * __fmaf_rn(v, v, 1.f));
        J[4] = __fmaf_rn(uv, dy, dx * __fmaf_rn(u, u, 1.f));
        J[5] = __fmaf_rn(u, dy, -(v * dx));
        J[6] = ea * (eb0 - refColor);
        J[7] = -1.f;
        J[8] = residual;
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const float Jw = J[r] * hw;
#pragma unroll
          for (int c = r; c < 9; c++) { acc[A_H + idx] = __fmaf_rn(Jw, J[c], acc[A_H + idx]); idx++; }
        }
      }
      if (dump && i < n) {
        dump[0 * n + i] = valid ? 1.f : 0.f;
        dump[1 * n + i] = nid[q]; dump[2 * n + i] = uu[q]; dump[3 * n + i] = vv[q];
        dump[4 * n + i] = hity; dump[5 * n + i] = hitz; dump[6 * n + i] = residual;
        dump[7 * n + i] = hw; dump[8 * n + i] = refColor;
      }
    }
  }
}

// per-evaluation constants for the SSE path (calcRes :617-621, calcGSSSE :540-544); called by ONE thread
__device__ void make_eval_const_sse(const TrackParams& P, const TrackLevel& L, const TrackProblem& prob, const double* R,
                                    const double* t, const double* aff, float cutoff, EvalConst& ec) {
  float Rf[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) ec.RKi[r * 3 + c] = Rf[r * 3 + 0] * L.Ki[0 * 3 + c] + Rf[r * 3 + 1] * L.Ki[1 * 3 + c] + Rf[r * 3 + 2] * L.Ki[2 * 3 + c];
#pragma unroll
  for (int i = 0; i < 3; i++) ec.t[i] = (float)t[i];
  double ab[2];
  d_aff_from_to(prob.ref_exposure, prob.exposure_new, prob.ref_aff[0], prob.ref_aff[1], aff[0], aff[1], ab);
  ec.affLL[0] = (float)ab[0]; ec.affLL[1] = (float)ab[1];
  ec.a = (float)ab[0];
  ec.b0 = (float)prob.ref_aff[1];
  ec.cutoff = cutoff;
  ec.maxEnergy = 2 * P.huberTH * cutoff - P.huberTH * P.huberTH;
}

// H (8x8), b from the 45 accumulated entries: calcGSSSE :582-595 (used by the operator-level entry)
template <typename T>
__device__ void finish_gs(const T* total, double* H, double* b) {
  const int nw = (int)total[A_NW];
  const int n = (nw + 3) & ~3;  // buf_warped_n is padded to a multiple of 4 (:763-773)
  const float invn = 1.0f / n;
  double M[81];
  int idx = 0;
  for (int r = 0; r < 9; r++)
    for (int c = r; c < 9; c++) { float d = (float)total[A_H + idx]; M[r * 9 + c] = M[c * 9 + r] = (double)d; idx++; }
  const double sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_A, SCALE_B};
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H[r * 8 + c] = M[r * 9 + c] * invn * sc[c] * sc[r];
    b[r] = M[r * 9 + 8] * invn * sc[r];
  }
}

template <typename T>
__device__ void rs_from_total(const T* total, double rs[6]) {
  // CoarseTracker.cpp:783-789 (float arithmetic as written)
  const float E = (float)total[A_E];
  const int numTermsInE = (int)total[A_NE];
  const int numSaturated = (int)total[A_NSAT];
  const float sT = (float)total[A_ST], sRT = (float)total[A_SRT], sN = (float)total[A_SN];
  rs[0] = E;
  rs[1] = numTermsInE;
  rs[2] = sT / (sN + 0.1);
  rs[3] = 0;
  rs[4] = sRT / (sN + 0.1);
  rs[5] = numSaturated / (float)numTermsInE;
}

__device__ __forceinline__ int tri9(int r, int c) {  // index of (r,c), r<=c, in the row-major upper triangle of a 9x9
  return r * 9 - (r * (r - 1)) / 2 + (c - r);
}

// 8x8 SPD solve, Gauss-Jordan over one warp with the matrix spread over all 32 lanes: lane = 4*r + q holds
// a[r][q], a[r][q+4] and (a copy of) the right-hand side a[r][8]. Each pivot step is 5 shuffles, one
// reciprocal and 3 FMAs per lane and the dependency chain per step is ~100 cycles (a single warp issues
// dependent instructions every ~5 cycles, so instruction COUNT and chain length are what matters here).
// No pivoting: the damped systems on this path are SPD. All lanes end with x[0..7].
__device__ __forceinline__ void warp_solve8_gj(double e0, double e1, double e2, int lane, double (&x)[8], bool& pd) {
  const int r = lane >> 2, q = lane & 3;
  pd = true;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int kq = k & 3, src = k * 4 + q;
    const double pk0 = __shfl_sync(0xffffffffu, e0, src), pk1 = __shfl_sync(0xffffffffu, e1, src), pk2 = __shfl_sync(0xffffffffu, e2, src);
    const double mine = (k >> 2) ? e1 : e0;  // slot that holds column k (compile-time choice)
    const double piv = __shfl_sync(0xffffffffu, mine, k * 4 + kq);
    const double ark = __shfl_sync(0xffffffffu, mine, (lane & ~3) + kq);
    pd = pd && (piv > 0) && isfinite(piv);
    const double f = ark * fast_rcp(piv);
    if (r != k) { e0 = fma(-f, pk0, e0); e1 = fma(-f, pk1, e1); e2 = fma(-f, pk2, e2); }
  }
  const double dsel = (r >> 2) ? e1 : e0;
  const double d = __shfl_sync(0xffffffffu, dsel, r * 4 + (r & 3));
  const double xr = e2 * fast_rcp(d);
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = __shfl_sync(0xffffffffu, xr, i * 4);
}

// exp(x) in double for the affine brightness factor: |x| < 0.5 uses a 15-term Horner (truncation 2e-17)
// instead of the library routine (~160 cycles); both round to the same float except in measure-zero cases.
__device__ __forceinline__ double d_exp_small(double x) {
  if (!(fabs(x) < 0.5)) return exp(x);
  double r = 1.0 / 1307674368000.0;
  r = fma(r, x, 1.0 / 87178291200.0);
  r = fma(r, x, 1.0 / 6227020800.0);
  r = fma(r, x, 1.0 / 479001600.0);
  r = fma(r, x, 1.0 / 39916800.0);
  r = fma(r, x, 1.0 / 3628800.0);
  r = fma(r, x, 1.0 / 362880.0);
  r = fma(r, x, 1.0 / 40320.0);
  r = fma(r, x, 1.0 / 5040.0);
  r = fma(r, x, 1.0 / 720.0);
  r = fma(r, x, 1.0 / 120.0);
  r = fma(r, x, 1.0 / 24.0);
  r = fma(r, x, 1.0 / 6.0);
  r = fma(r, x, 0.5);
  r = fma(r, x, 1.0);
  r = fma(r, x, 1.0);
  return r;
}

// Prologue of every evaluation, executed by warps 0 and 1 (the others wait at the CTA barrier):
//   warp 0, ITER stage : one LM step of the SSE path (commented block CoarseTracker.cpp:929-979): rows of
//                        Hl = H (diag*(1+lambda)) | -b straight from the accepted-state sums (calcGSSSE :582-595
//                        folded in), warp solve, scaled increment -> shared memory, then the trial pose
//                        exp(inc)*T lane-parallel (lane i < 9 owns entry (i/3, i%3) of the 3x3 blocks);
//   warp 0, any stage  : R*Ki and t of the evaluated pose (calcRes :617-618);
//   warp 1             : after warp 0 published the increment (named barrier), the brightness transfer
//                        (calcRes :621, calcGSSSE :543-544) of the evaluated affine state.
__device__ __forceinline__ void eval_prologue(const TrackParams& P, const TrackLevel& L, const TrackProblem& prob, LMState& lm, EvalConst& ec,
                                              float cutoff, bool trial, int tid, long long* cyc) {
  const int lane = tid & 31;
  if (tid < 32) {
    long long c0 = cyc ? clock64() : 0;
    const int e = lane < 9 ? lane : 0, er = e / 3, ecn = e % 3;
    double Rn_e, tn_e;
    if (trial) {
      const int r = lane >> 2, q = lane & 3;
      const float* tot = lm.totf;
      const int nwv = (int)tot[A_NW];
      const int n = (nwv + 3) & ~3;
      const float invn = 1.0f / n;
      // H(r,c) = acc * (1/n) * sc[c] * sc[r], sc = {1,1,1,.5,.5,.5,10,1000} (:584-595, names swapped as written there)
      const double scr = (r < 3) ? (double)SCALE_XI_ROT : (r < 6 ? (double)SCALE_XI_TRANS : (r == 6 ? (double)SCALE_A : (double)SCALE_B));
      const double sc0 = (q < 3) ? (double)SCALE_XI_ROT : (double)SCALE_XI_TRANS;                                  // column q
      const double sc1 = (q < 2) ? (double)SCALE_XI_TRANS : (q == 2 ? (double)SCALE_A : (double)SCALE_B);          // column q+4
      const double f1 = (double)invn * scr;
      const int c0i = q, c1i = q + 4;
      double e0 = (double)tot[A_H + tri9(r < c0i ? r : c0i, r < c0i ? c0i : r)] * (f1 * sc0);
      double e1 = (double)tot[A_H + tri9(r < c1i ? r : c1i, r < c1i ? c1i : r)] * (f1 * sc1);
      double e2 = -((double)tot[A_H + tri9(r, 8)] * f1);
      const float onePlusLambda = 1 + lm.lambda;
      if (c0i == r) e0 *= onePlusLambda;
      if (c1i == r) e1 *= onePlusLambda;
      // fixed affine parameters (:937-964): decouple the fixed unknown(s); the reduced solves are identical
      if (P.affineOptModeA < 0) { if (r == 6) { e0 = 0; e1 = (c1i == 6) ? 1.0 : 0.0; e2 = 0; } else if (c1i == 6) e1 = 0; }
      if (P.affineOptModeB < 0) { if (r == 7) { e0 = 0; e1 = (c1i == 7) ? 1.0 : 0.0; e2 = 0; } else if (c1i == 7) e1 = 0; }
      double inc[8];
      bool pd;
      if (cyc) { long long c1 = clock64(); cyc[6] += c1 - c0; c0 = c1; }
      warp_solve8_gj(e0, e1, e2, lane, inc, pd);
      if (cyc) { long long c1 = clock64(); cyc[7] += c1 - c0; c0 = c1; }
      const float lambdaExtrapolationLimit = 0.001f;
      if (lm.lambda < lambdaExtrapolationLimit) {
        const float extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lm.lambda));
#pragma unroll
        for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      }
      // incScaled (:969-976); SCALE_XI_ROT == 1
      double s[8];
#pragma unroll
      for (int i = 0; i < 3; i++) s[i] = inc[i];
#pragma unroll
      for (int i = 3; i < 6; i++) s[i] = inc[i] * SCALE_XI_TRANS;
      s[6] = inc[6] * SCALE_A; s[7] = inc[7] * SCALE_B;
      bool fin = true;
#pragma unroll
      for (int i = 0; i < 8; i++) fin = fin && isfinite(s[i]);  // !isfinite(incScaled.sum()) -> setZero
      if (!fin) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = 0;
      }
      // publish what warp 1 and the break test need
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) if (i == (lane & 7)) { const float v = (float)inc[i]; sq = v * v; }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      if (lane == 0) { lm.affn[0] = lm.aff[0] + s[6]; lm.affn[1] = lm.aff[1] + s[7]; lm.small = !(sq > 1e-6f) ? 1 : 0; }
      asm volatile("bar.arrive 1, 64;" ::: "memory");  // warp 1 may start the brightness transfer
      if (cyc) { long long c1 = clock64(); cyc[8] += c1 - c0; c0 = c1; }
      // ---- trial pose = exp(s[0:6]) * T, lane-parallel ----
      const double wx = s[3], wy = s[4], wz = s[5];
      const double x = fma(wx, wx, fma(wy, wy, wz * wz));
      double A, B, Cc;
      if (x < 0.0625) {
        A = -1.0 / 1307674368000.0; B = -1.0 / 20922789888000.0; Cc = -1.0 / 355687428096000.0;
        A = fma(A, x, 1.0 / 6227020800.0);  B = fma(B, x, 1.0 / 87178291200.0);  Cc = fma(Cc, x, 1.0 / 1307674368000.0);
        A = fma(A, x, -1.0 / 39916800.0);   B = fma(B, x, -1.0 / 479001600.0);   Cc = fma(Cc, x, -1.0 / 6227020800.0);
        A = fma(A, x, 1.0 / 362880.0);      B = fma(B, x, 1.0 / 3628800.0);      Cc = fma(Cc, x, 1.0 / 39916800.0);
        A = fma(A, x, -1.0 / 5040.0);       B = fma(B, x, -1.0 / 40320.0);       Cc = fma(Cc, x, -1.0 / 362880.0);
        A = fma(A, x, 1.0 / 120.0);         B = fma(B, x, 1.0 / 720.0);          Cc = fma(Cc, x, 1.0 / 5040.0);
        A = fma(A, x, -1.0 / 6.0);          B = fma(B, x, -1.0 / 24.0);          Cc = fma(Cc, x, -1.0 / 120.0);
        A = fma(A, x, 1.0);                 B = fma(B, x, 0.5);                  Cc = fma(Cc, x, 1.0 / 6.0);
      } else {
        se3_coeffs_general(x, &A, &B, &Cc);
      }
      // entry (er, ecn) of W = hat(w) and of W^2 = w w^T - |w|^2 I
      const double wr = er == 0 ? wx : (er == 1 ? wy : wz), wc = ecn == 0 ? wx : (ecn == 1 ? wy : wz);
      double Oe = 0.0;
      if (e == 1) Oe = -wz; else if (e == 2) Oe = wy; else if (e == 3) Oe = wz; else if (e == 5) Oe = -wx; else if (e == 6) Oe = -wy; else if (e == 7) Oe = wx;
      const double ident = (er == ecn) ? 1.0 : 0.0;
      const double O2e = fma(wr, wc, -(ident * x));
      const double Re_e = fma(B, O2e, fma(A, Oe, ident));   // exp(W)
      const double V_e = fma(Cc, O2e, fma(B, Oe, ident));   // V
      // te = V * upsilon, to = Re * t + te, Ro = Re * R : row er of Re / V lives in lanes 3*er .. 3*er+2
      const double up = ecn == 0 ? s[0] : (ecn == 1 ? s[1] : s[2]);
      const double vte = fma(Re_e, lm.t[ecn], V_e * up);  // lane (er,k): Re[er][k]*t[k] + V[er][k]*u[k]
      const int base = 3 * er;
      tn_e = __shfl_sync(0xffffffffu, vte, base) + __shfl_sync(0xffffffffu, vte, base + 1) + __shfl_sync(0xffffffffu, vte, base + 2);
      const double Re0 = __shfl_sync(0xffffffffu, Re_e, base), Re1 = __shfl_sync(0xffffffffu, Re_e, base + 1), Re2 = __shfl_sync(0xffffffffu, Re_e, base + 2);
      Rn_e = fma(Re0, lm.R[ecn], fma(Re1, lm.R[3 + ecn], Re2 * lm.R[6 + ecn]));  // (Re*R)[er][ecn]
      if (lane < 9) lm.Rn[lane] = Rn_e;
      if (lane < 9 && ecn == 0) lm.tn[er] = tn_e;
      if (cyc) { long long c1 = clock64(); cyc[9] += c1 - c0; c0 = c1; }
    } else {
      asm volatile("bar.arrive 1, 64;" ::: "memory");
      Rn_e = lm.R[e];
      tn_e = lm.t[er];
    }
    // R*Ki and t in float (calcRes :617-618): lane (r,c) needs row r of R
    const float Rf = (float)Rn_e;
    const int base = 3 * er;
    const float R0 = __shfl_sync(0xffffffffu, Rf, base), R1 = __shfl_sync(0xffffffffu, Rf, base + 1), R2 = __shfl_sync(0xffffffffu, Rf, base + 2);
    if (lane < 9) {
      ec.RKi[lane] = R0 * L.Ki[ecn] + R1 * L.Ki[3 + ecn] + R2 * L.Ki[6 + ecn];
      if (ecn == 0) ec.t[er] = (float)tn_e;
    }
    if (cyc) { long long c1 = clock64(); cyc[10] += c1 - c0; c0 = c1; }
  } else if (tid < 64) {
    asm volatile("bar.sync 1, 64;" ::: "memory");  // warp 0 has published lm.affn (trial) / nothing to wait for otherwise
    if (lane == 0) {
      // AffLight::fromToVecExposure (util/NumType.h:159-170)
      const double a0 = trial ? lm.affn[0] : lm.aff[0], b0 = trial ? lm.affn[1] : lm.aff[1];
      float eF = prob.ref_exposure, eT = prob.exposure_new;
      if (eF == 0 || eT == 0) { eT = eF = 1; }
      double a = d_exp_small(a0 - prob.ref_aff[0]) * eT;
      if (eF != 1.0f) a = a / eF;
      const double b = b0 - a * prob.ref_aff[1];
      ec.affLL[0] = (float)a; ec.affLL[1] = (float)b;
      ec.a = (float)a;
      ec.b0 = (float)prob.ref_aff[1];
      ec.cutoff = cutoff;
      ec.maxEnergy = 2 * P.huberTH * cutoff - P.huberTH * P.huberTH;
    }
  }
}

// The kernel is written as ONE loop with a single eval + exchange site (stages below): the persistent
// kernel walks through its code once per evaluation with only 8 warps per SM, so instruction fetch is on
// the critical path and every duplicated inlined copy of the evaluation costs real time.
enum { ST_LEVEL_INIT = 0, ST_CUTOFF_REPEAT = 1, ST_ITER = 2 };

// kU = gather batch per thread. kU == 1 is compiled for two resident CTAs per SM (<= 128 registers): the throughput
// configuration (many independent sequences, one small cluster each); kU >= 2 keeps all 255 registers for one CTA per SM,
// the latency configuration (one sequence spread over an 8-CTA cluster).
template <int kU, int kBT = 256, int kMB = (kU == 1 ? 2 : 1)>
__global__ void __launch_bounds__(kBT, kMB) track_kernel(TrackParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackSmem* sm = reinterpret_cast<TrackSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const int tid = threadIdx.x;
  const int gtid = rank * blockDim.x + tid, gthreads = C * blockDim.x;
  LMState& lm = sm->lm;
  float acc[kAccPad];
  Exchange ex;
  __shared__ int s_prob;
  __shared__ int s_pc_lvl;
  TexCache tc{nullptr, nullptr, nullptr, &s_pc_lvl, kCacheRounds * kBT};
  if (P.use_cache) {
    unsigned char* cb = smem_raw + ((sizeof(TrackSmem) + 15) & ~(size_t)15);
    tc.pc = reinterpret_cast<float4*>(cb);
    tc.tex = reinterpret_cast<float*>(cb + (size_t)tc.slots * sizeof(float4));
    tc.tag = reinterpret_cast<int*>(tc.tex + 12 * tc.slots);
  }
  if (tid == 0) {
    mbar_init(&sm->bar[0], 1); mbar_init(&sm->bar[1], 1);
    mbar_fence_init();
  }
  cluster.sync();  // barriers initialised and every CTA of the cluster resident before any DSMEM push

  // Problems of this CTA: its own one (static), or — throughput configuration, cluster size 1 — pulled from a work counter
  // until the batch is empty, so a sequence whose LM needs many iterations does not leave the other SMs idle.
  for (;;) {
  int prob_id = blockIdx.x / C;
  if (P.dynamic) {
    __syncthreads();
    if (tid == 0) s_prob = (int)atomicAdd(P.work_counter, 1u);
    __syncthreads();
    prob_id = s_prob;
    if (prob_id >= P.nb) break;
  }
  TrackProblem& prob = P.problems[prob_id];
  unsigned evals = 0;
  if (tc.tex) for (int k = tid; k < tc.slots; k += blockDim.x) tc.tag[k] = -1;   // a new frame: nothing cached
  if (tid == 0) {
    s_pc_lvl = -1;
    for (int i = 0; i < 9; i++) lm.R[i] = prob.T[(i / 3) * 4 + (i % 3)];
    for (int i = 0; i < 3; i++) lm.t[i] = prob.T[i * 4 + 3];
    lm.aff[0] = prob.aff[0]; lm.aff[1] = prob.aff[1];
  }
  __syncthreads();

  // phase timers accumulate in SHARED memory (a global read-modify-write per tick would cost more than the phases)
  __shared__ long long s_cyc[16];
  if (tid < 16) s_cyc[tid] = 0;
  __syncthreads();
  PhaseTimer tm;
  if (P.timing && P.mode == 0 && rank == 0 && tid == 0) tm.start(s_cyc);
  const bool single = (P.mode == 1);  // operator-level entry: one fused calcRes + calcGSSSE at eval_lvl
  bool haveRepeated = false, aborted = false;
  int lvl = single ? P.eval_lvl : P.coarsest;
  int stage = ST_LEVEL_INIT, iteration = 0;
  float levelCutoffRepeat = 1;
  float meanOld = 0;

  // per-level results live in shared memory (dynamic indexing by level would otherwise go to local memory)
  __shared__ double s_lastRes[5];
  __shared__ double s_flow[3];
  __shared__ int s_iters[5];
  if (tid < 5) { s_lastRes[tid] = NAN; s_iters[tid] = 0; }
  if (tid < 3) s_flow[tid] = 1000;
  __syncthreads();

  while (true) {
    const TrackLevel& L = P.L[lvl];
    const float cutoff = single ? P.eval_cutoff : P.coarseCutoffTH * levelCutoffRepeat;
    // ---- prologue: constants of this evaluation (warp 0) ----
    if (tid < 64) {
      eval_prologue(P, L, prob, lm, sm->ec, cutoff, stage == ST_ITER, tid, tm.cyc);
      if (tid == 0 && stage == ST_LEVEL_INIT) lm.lambda = 0.01f;
      tm.tick(11);
    }
    __syncthreads();
    tm.tick(0);
    // ---- the evaluation + the exchange: the only instance of this code in the kernel ----
    eval_points_sse<kU, kCacheRounds * kBT>(P, L, lvl, prob.tex[lvl], sm->ec, acc, evals, gtid, gthreads, single ? P.dump : nullptr, prob.pc[lvl], prob.pc_n[lvl], tc);
    tm.tick(1);
    const int pb = reduce_exchange<false>(acc, sm, ex, C, rank, 0.0, &tm);
    // behind the block barriers of the reduction: every thread has read the flag for this evaluation and stored its point
    // records of this level; the next evaluation reads the flag behind the prologue's barrier
    if (tid == 0) s_pc_lvl = lvl;
    const float sumE = gather_sumf(sm, pb, C, A_E), sumNE = gather_sumf(sm, pb, C, A_NE);

    // ---- epilogue ----
    if (stage != ST_ITER) {
      if (tid < kAcc) lm.totf[tid] = gather_sumf(sm, pb, C, tid);
      if (single) {
        __syncthreads();
        if (rank == 0 && tid == 0) {
          rs_from_total(lm.totf, prob.rs);
          finish_gs(lm.totf, prob.H, prob.b);
          const int nw = (int)lm.totf[A_NW];
          prob.warped_n = (nw + 3) & ~3;
        }
        break;
      }
      // rs[5] = numSaturated / numTermsInE (:789); cutoff doubling while > 60 % saturated (:897-904)
      const float sat = (int)gather_sumf(sm, pb, C, A_NSAT) / (float)(int)sumNE;
      __syncthreads();
      if (sat > 0.6 && levelCutoffRepeat < 50) { levelCutoffRepeat *= 2; stage = ST_CUTOFF_REPEAT; continue; }
      meanOld = sumE / sumNE;  // resOld[0] / resOld[1]  (float: the sums carry float precision)
      stage = ST_ITER; iteration = 0;
      tm.tick(5);
      continue;
    }

    // ST_ITER: accept / reject (:989-1019)
    const int maxIt = lvl == 0 ? 10 : (lvl == 1 ? 20 : 50);  // {10,20,50,50,50} (:861)
    const float meanNew = sumE / sumNE;
    const bool accept = meanNew < meanOld;
    const bool small = lm.small != 0;  // !(inc.norm() > 1e-3)
    __syncthreads();  // everyone has read the LM state before it is mutated
    if (accept) {
      meanOld = meanNew;
      if (tid < kAcc) lm.totf[tid] = gather_sumf(sm, pb, C, tid);  // sums of the accepted state (calcGSSSE at :996)
      if (tid >= 64 && tid < 73) lm.R[tid - 64] = lm.Rn[tid - 64];
      if (tid >= 73 && tid < 76) lm.t[tid - 73] = lm.tn[tid - 73];
      if (tid >= 76 && tid < 78) lm.aff[tid - 76] = lm.affn[tid - 76];
      if (tid == 78) lm.lambda *= 0.5f;
    } else if (tid == 0) {
      lm.lambda *= 4;
      if (lm.lambda < 0.001f) lm.lambda = 0.001f;
    }
    iteration++;
    __syncthreads();
    tm.tick(4);
    if (!small && iteration < maxIt) continue;

    // ---- level finished (:1026-1041) ----
    double rsOld[6];
    rs_from_total(lm.totf, rsOld);  // resOld of the accepted state
    const double lr = sqrtf((float)(rsOld[0] / rsOld[1]));  // :1028
    if (tid == 0) {
      s_lastRes[lvl] = lr; s_iters[lvl] += iteration;
      s_flow[0] = rsOld[2]; s_flow[1] = rsOld[3]; s_flow[2] = rsOld[4];
    }
    if (lr > 1.5 * prob.minResForAbort[lvl]) { aborted = true; break; }
    if (levelCutoffRepeat > 1 && !haveRepeated) { haveRepeated = true; }  // repeat this level once (:1036-1040)
    else lvl--;
    if (lvl < 0) break;
    levelCutoffRepeat = 1;
    stage = ST_LEVEL_INIT;
  }
  __syncthreads();

  // outputs (:1044-1068)
  if (!single && rank == 0 && tid == 0) {
    bool ok = !aborted;
    double aout[2] = {lm.aff[0], lm.aff[1]};
    if (ok) {
      if ((P.affineOptModeA != 0 && (fabsf((float)aout[0]) > 1.2)) || (P.affineOptModeB != 0 && (fabsf((float)aout[1]) > 200))) ok = false;
    }
    if (ok) {
      double rel[2];
      d_aff_from_to(prob.ref_exposure, prob.exposure_new, prob.ref_aff[0], prob.ref_aff[1], aout[0], aout[1], rel);
      const float r0 = (float)rel[0], r1 = (float)rel[1];
      if ((P.affineOptModeA == 0 && (fabsf(logf(r0)) > 1.5)) || (P.affineOptModeB == 0 && (fabsf(r1) > 200))) ok = false;
    }
    if (ok) {
      if (P.affineOptModeA < 0) aout[0] = 0;
      if (P.affineOptModeB < 0) aout[1] = 0;
    }
    if (!aborted) {
      for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) prob.T_out[r * 4 + c] = lm.R[r * 3 + c]; prob.T_out[r * 4 + 3] = lm.t[r]; }
      prob.aff_out[0] = aout[0]; prob.aff_out[1] = aout[1];
    } else {  // the reference returns before writing lastToNew_out / aff_g2l_out (:1032-1033)
      for (int i = 0; i < 12; i++) prob.T_out[i] = prob.T[i];
      prob.aff_out[0] = prob.aff[0]; prob.aff_out[1] = prob.aff[1];
    }
    for (int i = 0; i < 5; i++) { prob.lastResiduals[i] = s_lastRes[i]; prob.iterations[i] = s_iters[i]; }
    for (int i = 0; i < 3; i++) prob.flow[i] = s_flow[i];
    prob.ok = ok ? 1 : 0;
    for (int i = 0; i < 16; i++) prob.cyc[i] = s_cyc[i];
  }
  // evals: integer sum, order-independent
  {
    unsigned e = evals;
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((tid & 31) == 0) atomicAdd(&prob.evals, (unsigned long long)e);
  }
  if (!P.dynamic) break;
  }  // problems
  cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

#include "tracker_g2o.cuh"

// ================================================================================================
// host side
int tracker_create(sdso_ctx* ctx) {
  TrackerState* t = new TrackerState();
  ctx->tracker = t;
  t->K = ctx->G;
  for (int l = 0; l < ctx->G.levels; l++) {
    size_t n = (size_t)ctx->G.w[l] * ctx->G.h[l];
    SDSO_CUDA(ctx, cudaMalloc(&t->pc[l], n * sizeof(float4)));
    t->pc_cap[l] = (int)n;
    SDSO_CUDA(ctx, cudaMalloc(&t->idepth[l], n * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&t->wsum[l], n * sizeof(float)));
    SDSO_CUDA(ctx, cudaMalloc(&t->wsum_bak[l], n * sizeof(float)));
  }
  SDSO_CUDA(ctx, cudaMalloc(&t->scan_tmp, ((size_t)ctx->G.w[0] * ctx->G.h[0] + 1024) * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&t->d_counts, 64 * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&t->d_problems, t->max_problems * sizeof(TrackProblem)));
  SDSO_CUDA(ctx, cudaMallocHost(&t->h_problems, t->max_problems * sizeof(TrackProblem)));
  SDSO_CUDA(ctx, cudaMalloc(&t->d_work_counter, 4 * sizeof(unsigned)));
  return SDSO_OK;
}

void tracker_destroy(sdso_ctx* ctx) {
  TrackerState* t = ctx->tracker;
  if (!t) return;
  for (int l = 0; l < kPyrLevels; l++) {
    if (t->pc[l]) cudaFree(t->pc[l]);
    if (t->idepth[l]) cudaFree(t->idepth[l]);
    if (t->wsum[l]) cudaFree(t->wsum[l]);
    if (t->wsum_bak[l]) cudaFree(t->wsum_bak[l]);
  }
  if (t->scan_tmp) cudaFree(t->scan_tmp);
  if (t->d_counts) cudaFree(t->d_counts);
  if (t->d_problems) cudaFree(t->d_problems);
  if (t->h_problems) cudaFreeHost(t->h_problems);
  if (t->d_dump) cudaFree(t->d_dump);
  if (t->d_work_counter) cudaFree(t->d_work_counter);
  if (t->results_ready) cudaEventDestroy(t->results_ready);
  for (size_t k = 0; k < t->saved.size(); k++)
    if ((int)k != t->cur_slot) for (int l = 0; l < kPyrLevels; l++) if (t->saved[k].pc[l]) cudaFree(t->saved[k].pc[l]);
  for (int l = 0; l < kPyrLevels; l++) { if (t->edge_flag[l]) cudaFree(t->edge_flag[l]); if (t->edge_err[l]) cudaFree(t->edge_err[l]); }
  delete t;
  ctx->tracker = nullptr;
}

static void fill_params(sdso_ctx* ctx, TrackParams& P) {
  TrackerState* t = ctx->tracker;
  memset(&P, 0, sizeof(P));
  P.levels = ctx->G.levels;
  for (int l = 0; l < P.levels; l++) {
    TrackLevel& L = P.L[l];
    L.pc = t->pc[l]; L.n = t->pc_n[l];
    L.w = t->K.w[l]; L.h = t->K.h[l];
    L.fx = t->K.fx[l]; L.fy = t->K.fy[l]; L.cx = t->K.cx[l]; L.cy = t->K.cy[l];
    for (int i = 0; i < 9; i++) L.Ki[i] = t->K.Ki[l][i];
    L.gfx = ctx->G.K[l][0]; L.gfy = ctx->G.K[l][4]; L.gcx = ctx->G.K[l][2]; L.gcy = ctx->G.K[l][5];
  }
  P.ref_exposure = t->ref_exposure;  // (the kernels read the per-problem copy; the current slot may be empty in a multi-reference batch)
  P.ref_aff[0] = t->ref_aff[0]; P.ref_aff[1] = t->ref_aff[1];
  P.huberTH = ctx->S.huberTH; P.coarseCutoffTH = ctx->S.coarseCutoffTH;
  P.affineOptModeA = ctx->S.affineOptModeA; P.affineOptModeB = ctx->S.affineOptModeB;
  P.g2o_stop_persists = ctx->S.g2o_stop_flag_persists;
  P.timing = ctx->profile ? 1 : 0;
  P.problems = t->d_problems;
}

// Problem records go to the device through a kernel that reads the pinned (UVA-mapped) host array directly, not through a
// cudaMemcpyAsync: an H2D copy on the compute stream would queue on the same copy engine behind the bulk image upload of the
// NEXT step (copy stream) and stall the tracker launch until that upload has finished.
__global__ void stage_problems_kernel(const uint4* __restrict__ host_mapped, uint4* __restrict__ dev, int n16) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dev[i] = host_mapped[i];
}
static int stage_problems(sdso_ctx* ctx, int nb) {
  TrackerState* t = ctx->tracker;
  static_assert(sizeof(TrackProblem) % 16 == 0, "TrackProblem must be a multiple of 16 bytes");
  const int n16 = (int)(nb * sizeof(TrackProblem) / 16);
  int blocks = (n16 + 255) / 256;
  if (blocks > 64) blocks = 64;
  stage_problems_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(t->h_problems), reinterpret_cast<uint4*>(t->d_problems), n16);
  SDSO_CHECK_LAUNCH(ctx);
  return SDSO_OK;
}

// reference slot `slot` as a RefSlot view (the current slot lives in the TrackerState members)
static RefSlot slot_view(const TrackerState* t, int slot) {
  if (slot == t->cur_slot || slot < 0) {
    RefSlot r;
    for (int l = 0; l < kPyrLevels; l++) { r.pc[l] = t->pc[l]; r.pc_n[l] = t->pc_n[l]; r.pc_cap[l] = t->pc_cap[l]; }
    r.ref_frame = t->ref_frame; r.ref_exposure = t->ref_exposure; r.ref_aff[0] = t->ref_aff[0]; r.ref_aff[1] = t->ref_aff[1]; r.have_ref = t->have_ref;
    return r;
  }
  if (slot >= (int)t->saved.size()) return RefSlot();
  return t->saved[slot];
}
static int fill_problem_ref(sdso_ctx* ctx, TrackProblem& hp, int slot) {
  const RefSlot r = slot_view(ctx->tracker, slot);
  if (!r.have_ref) return fail(ctx, SDSO_E_STATE, "trackNewestCoarse before setCoarseTrackingRef (reference slot is empty)");
  for (int l = 0; l < ctx->G.levels; l++) { hp.pc[l] = r.pc[l]; hp.pc_n[l] = r.pc_n[l]; }
  hp.ref_exposure = r.ref_exposure;
  hp.ref_aff[0] = r.ref_aff[0]; hp.ref_aff[1] = r.ref_aff[1];
  return SDSO_OK;
}

static int launch_track(sdso_ctx* ctx, const TrackParams& P, int nb, bool g2o) {
  int C = ctx->S.cluster_size > 0 ? ctx->S.cluster_size : 8;
  int BT = ctx->S.block_threads > 0 ? ctx->S.block_threads : 256;
  if (BT > 256 || BT < 128 || (BT & 31)) return fail(ctx, SDSO_E_INVALID, "block_threads must be a multiple of 32 in [128,256]");
  if (C < 1 || C > 16) return fail(ctx, SDSO_E_INVALID, "cluster_size must be in [1,16]");
  size_t smem = sizeof(TrackSmem);
  const bool use_cache = !g2o && ctx->S.track_cache != 0;
  // (the cache is sized for the thread count the kernel instance was compiled for: 256, or 192 for the 168-register build)
  if (use_cache) smem = ((sizeof(TrackSmem) + 15) & ~(size_t)15) + cache_bytes((ctx->S.gather_batch == 2 && BT <= 192 && C == 1) ? 192 : 256);
  static bool attr_set_dev[64] = {false};   // function attributes are per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    const int big = (int)(((sizeof(TrackSmem) + 15) & ~(size_t)15) + cache_bytes(256));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    SDSO_CUDA(ctx, cudaFuncSetAttribute((track_kernel<2, 192, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_kernel<4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    SDSO_CUDA(ctx, cudaFuncSetAttribute((track_kernel<2, 192, 2>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    SDSO_CUDA(ctx, cudaFuncSetAttribute(track_g2o_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  // Throughput configuration (cluster size 1, SSE path): a persistent grid of at most two CTAs per SM pulls the problems from a
  // counter (dynamic load balance across sequences); otherwise one cluster per problem.
  const bool dynamic = (!g2o && C == 1 && P.mode == 0);
  int grid = C * nb;
  TrackParams Pl = P;
  Pl.nb = nb; Pl.dynamic = dynamic ? 1 : 0; Pl.work_counter = ctx->tracker->d_work_counter;
  Pl.use_cache = use_cache ? 1 : 0;
  if (dynamic) {
    const int cap = 2 * ctx->num_sms;
    if (grid > cap) grid = cap;
    SDSO_CUDA(ctx, cudaMemsetAsync(ctx->tracker->d_work_counter, 0, sizeof(unsigned), ctx->stream));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(BT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (g2o) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_g2o_kernel, Pl));
  else {
    const int U = ctx->S.gather_batch > 0 ? ctx->S.gather_batch : 2;
    if (U == 1) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_kernel<1>, Pl));
    else if (U == 2 && BT <= 192 && C == 1) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, (track_kernel<2, 192, 2>), Pl));   // 170 registers, 2 CTAs / SM
    else if (U == 2) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_kernel<2>, Pl));
    else if (U == 4) SDSO_CUDA(ctx, cudaLaunchKernelEx(&cfg, track_kernel<4>, Pl));
    else return fail(ctx, SDSO_E_INVALID, "gather_batch must be 1, 2 or 4");
  }
  ctx->launches++;
  return SDSO_OK;
}

// per-problem, per-level edge flags / errors of the g2o path
// (the stride is the largest template of the batch: each problem of a multi-reference batch has its own pc_n)
static int ensure_edge_scratch(sdso_ctx* ctx, TrackParams& P, int nb, const int* ref_slots = nullptr) {
  TrackerState* t = ctx->tracker;
  for (int l = 0; l < ctx->G.levels; l++) {
    int nmax = t->pc_n[l];
    if (ref_slots) {
      nmax = 0;
      for (int k = 0; k < nb; k++) { const int n = slot_view(t, ref_slots[k]).pc_n[l]; if (n > nmax) nmax = n; }
    }
    int stride = (nmax + 63) & ~63;
    if (stride < 64) stride = 64;
    size_t need = (size_t)stride * nb;
    if (need > t->edge_cap[l]) {
      if (t->edge_flag[l]) cudaFree(t->edge_flag[l]);
      if (t->edge_err[l]) cudaFree(t->edge_err[l]);
      t->edge_flag[l] = nullptr; t->edge_err[l] = nullptr; t->edge_cap[l] = 0;
      SDSO_CUDA(ctx, cudaMalloc(&t->edge_flag[l], need));
      SDSO_CUDA(ctx, cudaMalloc(&t->edge_err[l], need * sizeof(double)));
      t->edge_cap[l] = need;
    }
    P.edge_flag[l] = t->edge_flag[l]; P.edge_err[l] = t->edge_err[l]; P.edge_stride[l] = stride;
  }
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_tracker_make_k(sdso_ctx* ctx, const float K[4]) {
  sdso::enter(ctx);
  if (!ctx || !K) return SDSO_E_INVALID;
  ctx->tracker->K.set(ctx->G.w[0], ctx->G.h[0], K[0], K[1], K[2], K[3], false);
  ctx->tracker->K.levels = ctx->G.levels;
  return SDSO_OK;
}

int sdso_tracker_set_pc(sdso_ctx* ctx, int ref_frame, int lvl, int n, const float* u, const float* v, const float* idepth,
                        const float* color, const double ref_aff[2]) {
  sdso::enter(ctx);
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels || n < 0) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (ref_frame < 0 || ref_frame >= (int)ctx->frames.size() || !ctx->frames[ref_frame].in_use) return fail(ctx, SDSO_E_INVALID, "bad ref_frame");
  if (n > t->pc_cap[lvl]) return fail(ctx, SDSO_E_INVALID, "pc larger than level");
  std::vector<float4> tmp(n);
  for (int i = 0; i < n; i++) tmp[i] = make_float4(u[i], v[i], idepth[i], color[i]);
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->pc[lvl], tmp.data(), n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->pc_n[lvl] = n;
  t->ref_frame = ref_frame;
  t->ref_exposure = ctx->frames[ref_frame].ab_exposure;
  t->ref_aff[0] = ref_aff[0]; t->ref_aff[1] = ref_aff[1];
  t->have_ref = true;
  return SDSO_OK;
}

int sdso_tracker_get_pc(sdso_ctx* ctx, int lvl, int* n, float* u, float* v, float* idepth, float* color) {
  sdso::enter(ctx);
  if (!ctx || lvl < 0 || lvl >= ctx->G.levels || !n) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  *n = t->pc_n[lvl];
  std::vector<float4> tmp(*n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(tmp.data(), t->pc[lvl], (size_t)*n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < *n; i++) {
    if (u) u[i] = tmp[i].x;
    if (v) v[i] = tmp[i].y;
    if (idepth) idepth[i] = tmp[i].z;
    if (color) color[i] = tmp[i].w;
  }
  return SDSO_OK;
}

static int check_frame(sdso_ctx* ctx, int f) {
  if (f < 0 || f >= (int)ctx->frames.size() || !ctx->frames[f].in_use || !ctx->frames[f].valid) return fail(ctx, SDSO_E_INVALID, "bad frame id (not created or makeImages not run)");
  return SDSO_OK;
}

int sdso_calc_res_gs(sdso_ctx* ctx, int new_frame, int lvl, const double refToNew[12], const double aff[2], float cutoffTH,
                     double rs[6], double H[64], double b[8], int* warped_n, float* warped) {
  sdso::enter(ctx);
  if (!ctx || !refToNew || !aff) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!t->have_ref) return fail(ctx, SDSO_E_STATE, "calcRes before setCoarseTrackingRef");
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  int rc = check_frame(ctx, new_frame);
  if (rc) return rc;
  TrackParams P{};
  fill_params(ctx, P);
  P.mode = 1; P.eval_lvl = lvl; P.eval_cutoff = cutoffTH; P.variant = SDSO_VARIANT_SSE;
  const int n = t->pc_n[lvl];
  if (warped) {
    size_t need = (size_t)9 * (n > 0 ? n : 1) * sizeof(float);
    if (need > t->dump_cap) {
      if (t->d_dump) cudaFree(t->d_dump);
      SDSO_CUDA(ctx, cudaMalloc(&t->d_dump, need));
      t->dump_cap = need;
    }
    P.dump = t->d_dump;
  }
  TrackProblem& hp = t->h_problems[0];
  memset(&hp, 0, sizeof(hp));
  memcpy(hp.T, refToNew, sizeof(hp.T));
  so3_normalize(hp.T);
  hp.aff[0] = aff[0]; hp.aff[1] = aff[1];
  for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frame].tex[l];
  hp.exposure_new = ctx->frames[new_frame].ab_exposure;
  rc = fill_problem_ref(ctx, hp, -1);
  if (rc) return rc;
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_problems, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_track(ctx, P, 1, false);
  if (rc) return rc;
  SDSO_CUDA(ctx, cudaMemcpyAsync(&hp, t->d_problems, sizeof(hp), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<float> hd;
  if (warped) {
    hd.resize((size_t)9 * n);
    SDSO_CUDA(ctx, cudaMemcpyAsync(hd.data(), t->d_dump, hd.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (rs) memcpy(rs, hp.rs, sizeof(hp.rs));
  if (H) memcpy(H, hp.H, sizeof(hp.H));
  if (b) memcpy(b, hp.b, sizeof(hp.b));
  if (warped_n) *warped_n = hp.warped_n;
  if (warped) {
    // stable compaction in point order = the order calcRes fills buf_warped_* (:750-757), zero padded (:763-773)
    const int wn = hp.warped_n;
    int k = 0;
    for (int i = 0; i < n; i++) {
      if (hd[i] != 0.f) {
        for (int a = 0; a < 8; a++) warped[(size_t)a * wn + k] = hd[(size_t)(a + 1) * n + i];
        k++;
      }
    }
    for (; k < wn; k++) for (int a = 0; a < 8; a++) warped[(size_t)a * wn + k] = 0.f;
  }
  return SDSO_OK;
}

int sdso_edge_eval(sdso_ctx* ctx, int new_frame, int lvl, const double T_select[12], const double T_pose[12], const double photo[2],
                   int* n_out, double* err, double* J8) {
  sdso::enter(ctx);
  if (!ctx || !T_select || !T_pose || !photo || !n_out) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!t->have_ref) return fail(ctx, SDSO_E_STATE, "edge evaluation before setCoarseTrackingRef");
  if (lvl < 0 || lvl >= ctx->G.levels) return SDSO_E_INVALID;
  int rc = check_frame(ctx, new_frame);
  if (rc) return rc;
  TrackParams P{};
  fill_params(ctx, P);
  P.mode = 2; P.eval_lvl = lvl; P.eval_cutoff = 1e30f; P.variant = SDSO_VARIANT_G2O;
  rc = ensure_edge_scratch(ctx, P, 1);
  if (rc) return rc;
  const int n = t->pc_n[lvl];
  size_t need = (size_t)10 * (n > 0 ? n : 1) * sizeof(double);
  if (need > t->dump_cap) {
    if (t->d_dump) cudaFree(t->d_dump);
    t->d_dump = nullptr; t->dump_cap = 0;
    SDSO_CUDA(ctx, cudaMalloc(&t->d_dump, need));
    t->dump_cap = need;
  }
  P.dump_d = reinterpret_cast<double*>(t->d_dump);
  TrackProblem& hp = t->h_problems[0];
  memset(&hp, 0, sizeof(hp));
  memcpy(hp.T, T_select, sizeof(hp.T));
  memcpy(hp.T_out, T_pose, sizeof(hp.T_out));
  so3_normalize(hp.T); so3_normalize(hp.T_out);
  hp.aff_out[0] = photo[0]; hp.aff_out[1] = photo[1];
  for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frame].tex[l];
  hp.exposure_new = ctx->frames[new_frame].ab_exposure;
  { int rcr = fill_problem_ref(ctx, hp, -1); if (rcr) return rcr; }
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->d_problems, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_track(ctx, P, 1, true);
  if (rc) return rc;
  std::vector<double> hd((size_t)10 * n);
  SDSO_CUDA(ctx, cudaMemcpyAsync(hd.data(), t->d_dump, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int k = 0;
  for (int i = 0; i < n; i++) {
    if (hd[i] != 0.0) {
      if (err) err[k] = hd[(size_t)n + i];
      if (J8) for (int a = 0; a < 8; a++) J8[(size_t)8 * k + a] = hd[(size_t)(2 + a) * n + i];
      k++;
    }
  }
  *n_out = k;
  return SDSO_OK;
}

int sdso_track_enqueue(sdso_ctx* ctx, int nb, const int* new_frames, const double* T_in, const double* aff_in, int coarsest_lvl,
                       const double* minResForAbort, int variant) {
  sdso::enter(ctx);
  return sdso_track_enqueue_multi(ctx, nb, nullptr, new_frames, T_in, aff_in, coarsest_lvl, minResForAbort, variant);
}

int sdso_tracker_select_ref(sdso_ctx* ctx, int slot) {
  sdso::enter(ctx);
  if (!ctx || slot < 0 || slot >= 1024) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (slot == t->cur_slot) return SDSO_OK;
  const size_t need = (size_t)(slot > t->cur_slot ? slot : t->cur_slot) + 1;
  if (t->saved.size() < need) t->saved.resize(need);
  t->saved[t->cur_slot] = slot_view(t, t->cur_slot);  // park the current template
  RefSlot& r = t->saved[slot];
  for (int l = 0; l < ctx->G.levels; l++) {
    if (!r.pc[l]) {
      const size_t n = (size_t)ctx->G.w[l] * ctx->G.h[l];
      SDSO_CUDA(ctx, cudaMalloc(&r.pc[l], n * sizeof(float4)));
      r.pc_cap[l] = (int)n;
    }
  }
  for (int l = 0; l < kPyrLevels; l++) { t->pc[l] = r.pc[l]; t->pc_n[l] = r.pc_n[l]; t->pc_cap[l] = r.pc_cap[l]; }
  t->ref_frame = r.ref_frame; t->ref_exposure = r.ref_exposure; t->ref_aff[0] = r.ref_aff[0]; t->ref_aff[1] = r.ref_aff[1]; t->have_ref = r.have_ref;
  t->cur_slot = slot;
  return SDSO_OK;
}

int sdso_track_enqueue_multi(sdso_ctx* ctx, int nb, const int* ref_slots, const int* new_frames, const double* T_in, const double* aff_in,
                             int coarsest_lvl, const double* minResForAbort, int variant) {
  sdso::enter(ctx);
  if (!ctx || nb <= 0 || !new_frames || !T_in || !aff_in || !minResForAbort) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (!ref_slots && !t->have_ref) return fail(ctx, SDSO_E_STATE, "trackNewestCoarse before setCoarseTrackingRef");
  if (nb > t->max_problems) return fail(ctx, SDSO_E_INVALID, "too many problems in one batch");
  if (coarsest_lvl < 0 || coarsest_lvl >= ctx->G.levels || coarsest_lvl >= 5) return fail(ctx, SDSO_E_INVALID, "coarsest_lvl out of range");
  if (variant != SDSO_VARIANT_SSE && variant != SDSO_VARIANT_G2O) return SDSO_E_INVALID;
  TrackParams P{};
  fill_params(ctx, P);
  P.mode = 0; P.coarsest = coarsest_lvl; P.variant = variant;
  if (variant == SDSO_VARIANT_G2O) { int rc = ensure_edge_scratch(ctx, P, nb, ref_slots); if (rc) return rc; }
  for (int k = 0; k < nb; k++) {
    int rc = check_frame(ctx, new_frames[k]);
    if (rc) return rc;
    TrackProblem& hp = t->h_problems[k];
    memset(&hp, 0, sizeof(hp));
    memcpy(hp.T, T_in + 12 * k, sizeof(hp.T));
    so3_normalize(hp.T);   // SE3& lastToNew_out is a unit quaternion in the reference
    hp.aff[0] = aff_in[2 * k]; hp.aff[1] = aff_in[2 * k + 1];
    for (int i = 0; i < 5; i++) hp.minResForAbort[i] = minResForAbort[5 * k + i];
    for (int l = 0; l < ctx->G.levels; l++) hp.tex[l] = ctx->frames[new_frames[k]].tex[l];
    hp.exposure_new = ctx->frames[new_frames[k]].ab_exposure;
    rc = fill_problem_ref(ctx, hp, ref_slots ? ref_slots[k] : -1);
    if (rc) return rc;
  }
  int rc = stage_problems(ctx, nb);
  if (rc) return rc;
  prof_begin(ctx, 0);
  rc = launch_track(ctx, P, nb, variant == SDSO_VARIANT_G2O);
  if (rc) return rc;
  prof_end(ctx, 0);
  SDSO_CUDA(ctx, cudaMemcpyAsync(t->h_problems, t->d_problems, nb * sizeof(TrackProblem), cudaMemcpyDeviceToHost, ctx->stream));
  if (!t->results_ready) SDSO_CUDA(ctx, cudaEventCreateWithFlags(&t->results_ready, cudaEventDisableTiming));
  SDSO_CUDA(ctx, cudaEventRecord(t->results_ready, ctx->stream));
  t->last_nb = nb;
  return SDSO_OK;
}

int sdso_track_collect(sdso_ctx* ctx, int nb, double* T_out, double* aff_out, double* lastResiduals, double* flowIndicators,
                       int* iterations, int* ok, uint64_t* evals) {
  sdso::enter(ctx);
  if (!ctx) return SDSO_E_INVALID;
  TrackerState* t = ctx->tracker;
  if (nb != t->last_nb) return fail(ctx, SDSO_E_STATE, "collect does not match the last enqueue");
  // wait for this enqueue's results only: work queued behind it (the next frames' makeImages) keeps running
  if (t->results_ready) SDSO_CUDA(ctx, cudaEventSynchronize(t->results_ready));
  else SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t ev = 0;
  for (int k = 0; k < nb; k++) {
    const TrackProblem& hp = t->h_problems[k];
    if (T_out) memcpy(T_out + 12 * k, hp.T_out, sizeof(hp.T_out));
    if (aff_out) { aff_out[2 * k] = hp.aff_out[0]; aff_out[2 * k + 1] = hp.aff_out[1]; }
    if (lastResiduals) memcpy(lastResiduals + 5 * k, hp.lastResiduals, sizeof(hp.lastResiduals));
    if (flowIndicators) memcpy(flowIndicators + 3 * k, hp.flow, sizeof(hp.flow));
    if (iterations) memcpy(iterations + 5 * k, hp.iterations, sizeof(hp.iterations));
    if (ok) ok[k] = hp.ok;
    ev += hp.evals;
    for (int i = 0; i < 16; i++) t->last_cyc[i] = (k == 0 ? 0 : t->last_cyc[i]) + hp.cyc[i];
  }
  if (evals) *evals = ev;
  return SDSO_OK;
}

int sdso_track_batch(sdso_ctx* ctx, int nb, const int* new_frames, double* T_io, double* aff_io, int coarsest_lvl,
                     const double* minResForAbort, int variant, double* lastResiduals, double* flowIndicators, int* iterations,
                     int* ok) {
  sdso::enter(ctx);
  int rc = sdso_track_enqueue(ctx, nb, new_frames, T_io, aff_io, coarsest_lvl, minResForAbort, variant);
  if (rc) return rc;
  return sdso_track_collect(ctx, nb, T_io, aff_io, lastResiduals, flowIndicators, iterations, ok, nullptr);
}

int sdso_track(sdso_ctx* ctx, int new_frame, double T_io[12], double aff_io[2], int coarsest_lvl, const double minResForAbort[5],
               int variant, double lastResiduals[5], double flowIndicators[3], int iterations[5], int* ok) {
  sdso::enter(ctx);
  return sdso_track_batch(ctx, 1, &new_frame, T_io, aff_io, coarsest_lvl, minResForAbort, variant, lastResiduals, flowIndicators,
                          iterations, ok);
}

}  // extern "C"

// phase cycles (clock64 on rank 0 / thread 0) of the last collected track launch, when profiling is enabled:
// [0] serial LM step, [1] point loop, [2] block reduction, [3] cluster barrier + final sum, [4] accept/reject bookkeeping
extern "C" int sdso_track_phase_cycles(sdso_ctx* ctx, long long cyc[16]) {
  sdso::enter(ctx);
  if (!ctx || !cyc) return SDSO_E_INVALID;
  for (int i = 0; i < 16; i++) cyc[i] = ctx->tracker->last_cyc[i];
  return SDSO_OK;
}
