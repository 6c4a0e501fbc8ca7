// A1 — FrameHessian::makeImages (FullSystem/HessianBlocks.cpp:141-203) on the device.
//
// One streaming (HBM-bound) pass per batch of up to 128 images:
//   pyr_fused_kernel : one CTA per 64x64 level-0 tile (+ halo): source (float or 8-bit) -> shared memory -> all pyramid
//                      levels of the tile in shared memory (2x2 box mean in the reference's order 0.25f*(((a+b)+c)+d),
//                      :172-178) -> central differences, non-finite -> 0, absSquaredGrad (+ gamma factor, :196-200) ->
//                      one coalesced 16-byte streaming store {I,dx,dy,absSquaredGrad} per pixel of every level.
//   pyr_fused_u8_kernel : the same for 8-bit sources (width and pointer multiples of 16, >= 5 levels) with the level-0 tile staged as bytes.
//   pyr_wrap_kernel  : the two image columns whose flat-index difference wraps to the neighbouring row (:182-184).
// Algorithmic bytes per image: read W*H (8-bit) or 4*W*H + write 16*sum_l(w_l*h_l) texels. No separate intensity
// plane is written; the epipolar search extracts its level-0 plane on first use (ensure_intensity_plane).
#include "ctx.h"

namespace sdso {

constexpr int kMaxBatch = 128;  // images per launch (blockIdx.z); 3 pointers per image in the 4 KB kernel-parameter space

struct PyrGeom {  // identical for every frame of a context
  int levels;
  int w[kPyrLevels], h[kPyrLevels];
  int px_offset[kPyrLevels + 1];  // prefix sum of w_l*h_l: level l of a frame lives at base + px_offset[l]
  int use_gamma;
  int src_u8;                     // 0: float sources; else 8-bit sources (PhotometricUndistorter::processFrame in mode 1 = plain widening,
                                  // Undistort.cpp:222-260) and the widest load every row start / region origin / pointer allows: 1, 4 or 16 bytes
};
struct PyrBatch {
  const void* src[kMaxBatch];     // level-0 source (float or uint8), device memory
  float* img[kMaxBatch];          // per-frame intensity planes, all levels (level 0 = the float image)
  float4* tex[kMaxBatch];         // per-frame texels, all levels
};

__device__ float g_Bgamma[256];  // CalibHessian::B (HessianBlocks.h:352), identity unless sdso_set_gamma

constexpr int kTile = 64;

// getBGradOnly factor + absSquaredGrad (HessianBlocks.cpp:190-200)
__device__ __forceinline__ float4 make_texel(float c, float dx, float dy, int use_gamma) {
  if (!isfinite(dx)) dx = 0.f;
  if (!isfinite(dy)) dy = 0.f;
  float ag = dx * dx + dy * dy;
  if (use_gamma) {
    int ci = (int)(c + 0.5f);  // CalibHessian::getBGradOnly (HessianBlocks.h:356-362)
    if (ci < 5) ci = 5;
    if (ci > 250) ci = 250;
    const float gw = g_Bgamma[ci + 1] - g_Bgamma[ci];
    ag *= gw * gw;
  }
  return make_float4(c, dx, dy, ag);
}

// ONE pass per image: a CTA owns a 64x64 level-0 tile plus a halo of 2^(levels-1) pixels, widens / loads it into shared
// memory once, builds every coarser level of the tile (with its halo) in shared memory, and writes the {I,dx,dy,|grad|^2}
// texels and the intensity planes of ALL levels from there. The source is read once (the halo re-reads hit L2), nothing is
// read back from HBM. Image columns 0 and w-1, whose horizontal difference wraps to the neighbouring row in the reference
// (flat idx +- 1, HessianBlocks.cpp:182-184), are finished by pyr_wrap_kernel.
// The level count is a template parameter: every region size, divisor and trip count is a constant, and the write phase
// maps threads to (column, row-stride) so a warp stores 32 consecutive texels (512 B) per instruction with no index
// division — the first version of this kernel spent 170 instructions per pixel and was issue-bound at 48 % of HBM.
template <int L, int l, bool FINITE, bool GAMMA, typename TS = float>
__device__ __forceinline__ void pyr_write_level(const PyrGeom& P, const TS* __restrict__ src, float* __restrict__ base, float4* __restrict__ tbase,
                                                int tid, const float* __restrict__ s_dB) {
  constexpr int H0 = 1 << (L - 1), R0 = kTile + 2 * H0;
  constexpr int Rs = R0 >> l, T = kTile >> l, Hl = H0 >> l;
  constexpr int TX = T < 32 ? T : (T > 64 ? 64 : T);   // threads along x
  constexpr int NB = 256 / TX;                           // row bands
  constexpr int BR = (T + NB - 1) / NB;                  // consecutive rows per band (0 threads idle when NB > T)
  const int wl = P.w[l], hl = P.h[l];
  const int tx0 = (blockIdx.x * kTile) >> l, ty0 = (blockIdx.y * kTile) >> l;
  const int lx = tid % TX, ly0 = (tid / TX) * BR;
  const int x = tx0 + lx;
  if (x >= wl || ly0 >= T) return;
  const int ylast = min(min(T, ly0 + BR), hl - ty0);     // end of this band inside the image
  const size_t o0 = (size_t)P.px_offset[l] + x + (size_t)(ty0 + ly0) * wl;
  float4* __restrict__ Tl = tbase + o0;
  const TS* q = src + (Hl + ly0) * Rs + (lx + Hl);
  // a thread walks down its column: the centre values of the rows above / below stay in registers
  float cu = (float)q[-Rs], c = (float)q[0];
#pragma unroll 4
  for (int ly = ly0; ly < ylast; ly++, q += Rs, Tl += wl) {
    const int y = ty0 + ly;
    const float cd = (float)q[Rs];
    // the halo holds the neighbours of every tile pixel, so the differences are formed unconditionally and dropped for the
    // first / last image row (flat idx outside [w, w(h-1)), :182); columns 0 and w-1 are redone by pyr_wrap_kernel
    float dx = 0.5f * ((float)q[1] - (float)q[-1]);
    float dy = 0.5f * (cd - cu);
    if (!FINITE) { if (!isfinite(dx)) dx = 0.f; if (!isfinite(dy)) dy = 0.f; }
    float ag = dx * dx + dy * dy;
    if (GAMMA) {
      int ci = (int)(c + 0.5f);  // CalibHessian::getBGradOnly (HessianBlocks.h:356-362)
      ci = min(max(ci, 5), 250);
      const float gw = s_dB[ci];
      ag *= gw * gw;
    }
    const bool inner = (y >= 1 && y < hl - 1);
    // streaming store: the texels of an image pass through L2 once and must not evict what the tracker keeps there.
    // (No separate intensity plane is written: the epipolar search extracts its level-0 plane on first use, trace.cu.)
    __stcs(Tl, make_float4(c, inner ? dx : 0.f, inner ? dy : 0.f, inner ? ag : 0.f));
    cu = c; c = cd;
  }
}

template <int L, int l, bool FINITE, bool GAMMA>
struct PyrWriteLevels {
  static __device__ __forceinline__ void run(const PyrGeom& P, const float* src, float* base, float4* tbase, int tid, const float* s_dB) {
    constexpr int H0 = 1 << (L - 1), R0 = kTile + 2 * H0, Rs = R0 >> l;
    pyr_write_level<L, l, FINITE, GAMMA>(P, src, base, tbase, tid, s_dB);
    PyrWriteLevels<L, l + 1, FINITE, GAMMA>::run(P, src + Rs * Rs, base, tbase, tid, s_dB);
  }
};
template <int L, bool FINITE, bool GAMMA>
struct PyrWriteLevels<L, L, FINITE, GAMMA> {
  static __device__ __forceinline__ void run(const PyrGeom&, const float*, float*, float4*, int, const float*) {}
};

template <int L>
__global__ void __launch_bounds__(256) pyr_fused_kernel(PyrGeom P, PyrBatch B) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  constexpr int H0 = 1 << (L - 1);          // level-0 halo
  constexpr int R0 = kTile + 2 * H0;        // level-0 region side
  const int ox = blockIdx.x * kTile - H0, oy = blockIdx.y * kTile - H0;  // level-0 origin of the region
  const int w0 = P.w[0], h0 = P.h[0];
  float* __restrict__ base = B.img[blockIdx.z];
  float4* __restrict__ tbase = B.tex[blockIdx.z];
  // ---- level 0 region into shared memory (0 outside the image)
  float* s0 = sm;
  __shared__ float s_dB[256];   // B[ci + 1] - B[ci], the factor getBGradOnly returns
  if (P.use_gamma) s_dB[tid] = tid < 255 ? g_Bgamma[tid + 1] - g_Bgamma[tid] : 0.f;
  if (P.src_u8 == 16 && H0 % 16 == 0) {
    // 16 pixels per load: ox (= 64 bx - H0, L >= 5), w0 and the source pointer are multiples of 16 (checked by the host, src_u8 == 16),
    // so a group is aligned and lies inside or outside the image as a whole
    const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
    constexpr int R16 = R0 >> 4, NG = R0 * R16, NIT = (NG + 255) / 256;
    uint4 qv[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {   // every load of the thread is in flight before the first conversion
      const int k = tid + it * 256;
      const int ly = k / R16, lx = (k - ly * R16) * 16;
      const int x = ox + lx, y = oy + ly;
      qv[it] = make_uint4(0u, 0u, 0u, 0u);
      if (k < NG && y >= 0 && y < h0 && x >= 0 && x < w0) qv[it] = __ldg(reinterpret_cast<const uint4*>(S8 + x + y * w0));
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int k = tid + it * 256;
      if (k >= NG) break;
      const int ly = k / R16, lx = (k - ly * R16) * 16;
      const uint4 q = qv[it];
      float4* d = reinterpret_cast<float4*>(s0 + ly * R0 + lx);
      // lanes are 64 B apart: rotating the order in which a lane stores its four float4 spreads every 8-lane wavefront over
      // all 32 banks (in lane order the 128-bit stores are 4-way conflicted)
      const int rot = (tid >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int jj = (j + rot) & 3;
        const unsigned wv = jj == 0 ? q.x : (jj == 1 ? q.y : (jj == 2 ? q.z : q.w));
        d[jj] = make_float4((float)(wv & 0xffu), (float)((wv >> 8) & 0xffu), (float)((wv >> 16) & 0xffu), (float)(wv >> 24));
      }
    }
  } else if (P.src_u8 >= 4 && H0 % 4 == 0) {   // L >= 3; w0 and the pointer are multiples of 4 (host check)
    const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
    constexpr int R4 = R0 >> 2;
#pragma unroll 3
    for (int k = tid; k < R0 * R4; k += 256) {
      const int ly = k / R4, lx = (k - ly * R4) * 4;
      const int x = ox + lx, y = oy + ly;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < h0 && x >= 0 && x < w0) {   // x and w0 are multiples of 4: the group is inside or outside as a whole
        const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(S8 + x + y * w0));
        v = make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
      }
      *reinterpret_cast<float4*>(s0 + ly * R0 + lx) = v;
    }
  } else if (P.src_u8) {
    const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
    for (int k = tid; k < R0 * R0; k += 256) {
      const int ly = k / R0, lx = k - ly * R0;
      const int x = ox + lx, y = oy + ly;
      s0[k] = (y >= 0 && y < h0 && x >= 0 && x < w0) ? (float)__ldg(S8 + x + (size_t)y * w0) : 0.f;
    }
  } else {
    const float* __restrict__ SF = (const float*)B.src[blockIdx.z];
    for (int k = tid; k < R0 * R0; k += 256) {
      const int ly = k / R0, lx = k - ly * R0;
      const int x = ox + lx, y = oy + ly;
      s0[k] = (y >= 0 && y < h0 && x >= 0 && x < w0) ? __ldg(SF + x + (size_t)y * w0) : 0.f;
    }
  }
  __syncthreads();
  // ---- coarser levels of the region: 2x2 mean in the reference's order 0.25f*(((a+b)+c)+d) (:172-178)
  {
    float* src = s0; int Rs = R0;
    float* dst = s0 + R0 * R0;
#pragma unroll
    for (int l = 1; l < L; l++) {
      const int Rd = Rs >> 1;
      if (Rd % 2 == 0) {      // two outputs per thread from two 16-byte rows
        const int Rh = Rd >> 1;
        for (int k = tid; k < Rd * Rh; k += 256) {
          const int ly = k / Rh, lx = (k - ly * Rh) * 2;
          const float4 a = *reinterpret_cast<const float4*>(src + (2 * ly) * Rs + 2 * lx);
          const float4 b = *reinterpret_cast<const float4*>(src + (2 * ly + 1) * Rs + 2 * lx);
          *reinterpret_cast<float2*>(dst + ly * Rd + lx) = make_float2(0.25f * (((a.x + a.y) + b.x) + b.y), 0.25f * (((a.z + a.w) + b.z) + b.w));
        }
      } else {
        for (int k = tid; k < Rd * Rd; k += 256) {
          const int ly = k / Rd, lx = k - ly * Rd;
          const float2 a = *reinterpret_cast<const float2*>(src + (2 * ly) * Rs + 2 * lx);
          const float2 b = *reinterpret_cast<const float2*>(src + (2 * ly + 1) * Rs + 2 * lx);
          dst[k] = 0.25f * (((a.x + a.y) + b.x) + b.y);
        }
      }
      __syncthreads();
      src = dst; Rs = Rd; dst = dst + Rd * Rd;
    }
  }
  // ---- texels + intensity planes of every level from shared memory
  if (P.src_u8) {   // 8-bit sources cannot produce non-finite differences
    if (P.use_gamma) PyrWriteLevels<L, 0, true, true>::run(P, s0, base, tbase, tid, s_dB);
    else PyrWriteLevels<L, 0, true, false>::run(P, s0, base, tbase, tid, s_dB);
  } else {
    if (P.use_gamma) PyrWriteLevels<L, 0, false, true>::run(P, s0, base, tbase, tid, s_dB);
    else PyrWriteLevels<L, 0, false, false>::run(P, s0, base, tbase, tid, s_dB);
  }
}

// 8-bit sources whose width and pointer are multiples of 16 and pyramids of >= 5 levels (region side and origin multiples of 16): the level-0 region is
// staged as BYTES (9 KB instead of 37 KB for 5 levels, 21 KB per CTA in all), so six CTAs are resident per SM instead of four, the
// load phase is one 16-byte shared store per 16 pixels, and the widening happens where the values are consumed. Same arithmetic
// on the same values as pyr_fused_kernel (every 8-bit value is exact in float).
template <int L>
__global__ void __launch_bounds__(256) pyr_fused_u8_kernel(PyrGeom P, PyrBatch B) {
  extern __shared__ __align__(16) unsigned char smb[];
  const int tid = threadIdx.x;
  constexpr int H0 = 1 << (L - 1), R0 = kTile + 2 * H0;
  static_assert(R0 % 16 == 0 && H0 % 16 == 0, "region side and origin must be multiples of 16");
  const int ox = blockIdx.x * kTile - H0, oy = blockIdx.y * kTile - H0;
  const int w0 = P.w[0], h0 = P.h[0];
  float* __restrict__ base = B.img[blockIdx.z];
  float4* __restrict__ tbase = B.tex[blockIdx.z];
  unsigned char* s0 = smb;
  float* s1 = reinterpret_cast<float*>(smb + R0 * R0);   // R0 * R0 is a multiple of 16
  __shared__ float s_dB[256];
  if (P.use_gamma) s_dB[tid] = tid < 255 ? g_Bgamma[tid + 1] - g_Bgamma[tid] : 0.f;
  {
    const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
    constexpr int R16 = R0 >> 4, NG = R0 * R16, NIT = (NG + 255) / 256;
    uint4 qv[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {   // every load of the thread is in flight before the first store
      const int k = tid + it * 256;
      const int ly = k / R16, lx = (k - ly * R16) * 16;
      const int x = ox + lx, y = oy + ly;
      qv[it] = make_uint4(0u, 0u, 0u, 0u);
      if (k < NG && y >= 0 && y < h0 && x >= 0 && x < w0) qv[it] = __ldg(reinterpret_cast<const uint4*>(S8 + x + y * w0));
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int k = tid + it * 256;
      if (k >= NG) break;
      const int ly = k / R16, lx = (k - ly * R16) * 16;
      *reinterpret_cast<uint4*>(s0 + ly * R0 + lx) = qv[it];
    }
  }
  __syncthreads();
  // level 1 from the bytes: 2x2 mean in the reference's order 0.25f*(((a+b)+c)+d) (:172-178), two outputs per thread
  {
    constexpr int Rd = R0 >> 1, Rh = Rd >> 1;
    for (int k = tid; k < Rd * Rh; k += 256) {
      const int ly = k / Rh, lx = (k - ly * Rh) * 2;
      const uchar4 a = *reinterpret_cast<const uchar4*>(s0 + (2 * ly) * R0 + 2 * lx);
      const uchar4 b = *reinterpret_cast<const uchar4*>(s0 + (2 * ly + 1) * R0 + 2 * lx);
      *reinterpret_cast<float2*>(s1 + ly * Rd + lx) =
          make_float2(0.25f * ((((float)a.x + (float)a.y) + (float)b.x) + (float)b.y), 0.25f * ((((float)a.z + (float)a.w) + (float)b.z) + (float)b.w));
    }
    __syncthreads();
    float* src = s1; int Rs = Rd;
    float* dst = s1 + Rd * Rd;
#pragma unroll
    for (int l = 2; l < L; l++) {
      const int Rn = Rs >> 1;
      if (Rn % 2 == 0) {
        const int Rq = Rn >> 1;
        for (int k = tid; k < Rn * Rq; k += 256) {
          const int ly = k / Rq, lx = (k - ly * Rq) * 2;
          const float4 a = *reinterpret_cast<const float4*>(src + (2 * ly) * Rs + 2 * lx);
          const float4 b = *reinterpret_cast<const float4*>(src + (2 * ly + 1) * Rs + 2 * lx);
          *reinterpret_cast<float2*>(dst + ly * Rn + lx) = make_float2(0.25f * (((a.x + a.y) + b.x) + b.y), 0.25f * (((a.z + a.w) + b.z) + b.w));
        }
      } else {
        for (int k = tid; k < Rn * Rn; k += 256) {
          const int ly = k / Rn, lx = k - ly * Rn;
          const float2 a = *reinterpret_cast<const float2*>(src + (2 * ly) * Rs + 2 * lx);
          const float2 b = *reinterpret_cast<const float2*>(src + (2 * ly + 1) * Rs + 2 * lx);
          dst[k] = 0.25f * (((a.x + a.y) + b.x) + b.y);
        }
      }
      __syncthreads();
      src = dst; Rs = Rn; dst = dst + Rn * Rn;
    }
  }
  if (P.use_gamma) {
    pyr_write_level<L, 0, true, true, unsigned char>(P, s0, base, tbase, tid, s_dB);
    PyrWriteLevels<L, 1, true, true>::run(P, s1, base, tbase, tid, s_dB);
  } else {
    pyr_write_level<L, 0, true, false, unsigned char>(P, s0, base, tbase, tid, s_dB);
    PyrWriteLevels<L, 1, true, false>::run(P, s1, base, tbase, tid, s_dB);
  }
}

// The reference differentiates on the flat index, so at x = 0 the left neighbour is (w-1, y-1) and at x = w-1 the right
// neighbour is (0, y+1) (HessianBlocks.cpp:182-184). One thread per (level, row, side); reads the intensities (.x) of the texels
// written above (a neighbour may be rewritten concurrently by another thread of this kernel, with the same .x).
__global__ void __launch_bounds__(128) pyr_wrap_kernel(PyrGeom P, PyrBatch B) {
  float4* tbase = B.tex[blockIdx.z];
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < P.levels; l++) {
    const int wl = P.w[l], hl = P.h[l];
    const int cnt = 2 * (hl - 2);
    if (k < cnt) {
      const int y = 1 + (k >> 1), side = k & 1;
      const float4* I = tbase + P.px_offset[l];
      const int x = side ? wl - 1 : 0;
      const int idx = x + y * wl;
      const float c = I[idx].x;
      const float dx = 0.5f * (I[idx + 1].x - I[idx - 1].x);
      const float dy = 0.5f * (I[idx + wl].x - I[idx - wl].x);
      tbase[P.px_offset[l] + idx] = make_texel(c, dx, dy, P.use_gamma);
      return;
    }
    k -= cnt;
  }
}

// nb frames, one launch pair. srcs[i]: device pointer to the level-0 source of frames[i] (float, or uint8 when src_u8).
int make_images_batch_launch(sdso_ctx* ctx, int nb, Frame* const* frames, const void* const* srcs, bool src_u8, bool use_hcalib) {
  if (nb <= 0) return SDSO_OK;
  if (nb > kMaxBatch) return fail(ctx, SDSO_E_INVALID, "make_images batch larger than 128");
  PyrGeom P;
  P.levels = ctx->G.levels;
  int off = 0;
  for (int l = 0; l < kPyrLevels; l++) {
    P.w[l] = l < P.levels ? ctx->G.w[l] : 0; P.h[l] = l < P.levels ? ctx->G.h[l] : 0;
    P.px_offset[l] = off;
    off += P.w[l] * P.h[l];
  }
  P.px_offset[kPyrLevels] = off;
  for (int l = P.levels; l <= kPyrLevels; l++) P.px_offset[l] = off;
  P.use_gamma = (use_hcalib && ctx->S.gammaWeightsPixelSelect == 1) ? 1 : 0;
  P.src_u8 = 0;
  if (src_u8) {   // widest aligned load: every row start (w0), every source pointer; the region origin is checked per kernel (H0)
    uintptr_t bits = (uintptr_t)P.w[0];
    for (int i = 0; i < nb; i++) bits |= (uintptr_t)srcs[i];
    P.src_u8 = (bits & 15) == 0 ? 16 : ((bits & 3) == 0 ? 4 : 1);
  }
  PyrBatch B;
  for (int i = 0; i < kMaxBatch; i++) { B.src[i] = nullptr; B.img[i] = nullptr; B.tex[i] = nullptr; }
  for (int i = 0; i < nb; i++) { B.src[i] = srcs[i]; B.img[i] = frames[i]->image; B.tex[i] = frames[i]->tex[0]; }
  prof_begin(ctx, 1);
  {
    const int H0 = 1 << (P.levels - 1);
    size_t smem = 0;
    for (int l = 0, R = kTile + 2 * H0; l < P.levels; l++, R >>= 1) smem += (size_t)R * R * sizeof(float);
    dim3 grid((P.w[0] + kTile - 1) / kTile, (P.h[0] + kTile - 1) / kTile, nb);
    static bool attr_set_dev[64] = {false};   // function attributes are per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
    if (!attr_set) {   // the 6-level region needs 87 KB of dynamic shared memory
      SDSO_CUDA(ctx, cudaFuncSetAttribute(pyr_fused_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      SDSO_CUDA(ctx, cudaFuncSetAttribute(pyr_fused_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    // byte-staged kernel: 16-byte loads at x = 64 bx - H0 + 16 j need H0 = 2^(L-1) to be a multiple of 16, i.e. L >= 5
    const bool fast8 = P.src_u8 == 16 && P.levels >= 5;
    if (fast8) {
      const int R0 = kTile + 2 * H0;
      size_t smem8 = (size_t)R0 * R0;
      for (int l = 1, R = R0 >> 1; l < P.levels; l++, R >>= 1) smem8 += (size_t)R * R * sizeof(float);
      switch (P.levels) {
        case 5: pyr_fused_u8_kernel<5><<<grid, 256, smem8, ctx->stream>>>(P, B); break;
        default: pyr_fused_u8_kernel<6><<<grid, 256, smem8, ctx->stream>>>(P, B); break;
      }
    } else
    switch (P.levels) {
      case 1: pyr_fused_kernel<1><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      case 2: pyr_fused_kernel<2><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      case 3: pyr_fused_kernel<3><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      case 4: pyr_fused_kernel<4><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      case 5: pyr_fused_kernel<5><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      case 6: pyr_fused_kernel<6><<<grid, 256, smem, ctx->stream>>>(P, B); break;
      default: return fail(ctx, SDSO_E_INVALID, "make_images: unsupported pyramid depth");
    }
    SDSO_CHECK_LAUNCH(ctx);
    int rows = 0;
    for (int l = 0; l < P.levels; l++) rows += 2 * (P.h[l] - 2);
    pyr_wrap_kernel<<<dim3((rows + 127) / 128, 1, nb), 128, 0, ctx->stream>>>(P, B);
    SDSO_CHECK_LAUNCH(ctx);
  }
  prof_end(ctx, 1);
  return SDSO_OK;
}

__global__ void plane_extract_kernel(const float4* __restrict__ tex0, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex0[i].x;
}

// The epipolar search reads 4-byte pixels; makeImages does not write a separate plane (19 % of its traffic), so the plane is
// extracted from the level-0 texels the first time a frame is searched.
int ensure_intensity_plane(sdso_ctx* ctx, Frame& f) {
  if (f.plane_valid) return SDSO_OK;
  const int n = ctx->G.w[0] * ctx->G.h[0];
  plane_extract_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f.tex[0], n, f.image);
  SDSO_CHECK_LAUNCH(ctx);
  f.plane_valid = true;
  return SDSO_OK;
}

int make_images_launch(sdso_ctx* ctx, Frame& f, const float* dev_image, bool use_hcalib) {
  Frame* fp = &f;
  const void* src = dev_image;
  return make_images_batch_launch(ctx, 1, &fp, &src, false, use_hcalib);
}

int set_gamma_table(sdso_ctx* ctx, const float B[256]) {
  SDSO_CUDA(ctx, cudaMemcpyToSymbolAsync(g_Bgamma, B, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

}  // namespace sdso
