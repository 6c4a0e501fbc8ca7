// A1 — FrameHessian::makeImages (FullSystem/HessianBlocks.cpp:141-203) on the device.
//
// One streaming (HBM-bound) pass per batch of up to 32 images:
//   pyr_fused_kernel : one CTA per 64x64 level-0 tile (+ halo): source (float or 8-bit) -> shared memory -> all pyramid
//                      levels of the tile in shared memory (2x2 box mean in the reference's order 0.25f*(((a+b)+c)+d),
//                      :172-178) -> central differences, non-finite -> 0, absSquaredGrad (+ gamma factor, :196-200) ->
//                      one coalesced 16-byte store {I,dx,dy,absSquaredGrad} per pixel of every level + the intensity planes.
//   pyr_wrap_kernel  : the two image columns whose flat-index difference wraps to the neighbouring row (:182-184).
// Algorithmic bytes per image: read W*H (8-bit) or 4*W*H + write 16*sum_l(w_l*h_l) texels + 4*sum_l planes.
#include "ctx.h"

namespace sdso {

constexpr int kMaxBatch = 32;  // images per launch (blockIdx.z)

struct PyrGeom {  // identical for every frame of a context
  int levels;
  int w[kPyrLevels], h[kPyrLevels];
  int px_offset[kPyrLevels + 1];  // prefix sum of w_l*h_l: level l of a frame lives at base + px_offset[l]
  int use_gamma;
  int src_u8;                     // source images are 8-bit (PhotometricUndistorter::processFrame in mode 1 = plain widening, Undistort.cpp:222-260)
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
__global__ void __launch_bounds__(256) pyr_fused_kernel(PyrGeom P, PyrBatch B) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  const int L = P.levels;
  const int H0 = 1 << (L - 1);              // level-0 halo
  const int R0 = kTile + 2 * H0;            // level-0 region side
  const int ox = blockIdx.x * kTile - H0, oy = blockIdx.y * kTile - H0;  // level-0 origin of the region
  const int w0 = P.w[0], h0 = P.h[0];
  float* __restrict__ base = B.img[blockIdx.z];
  float4* __restrict__ tbase = B.tex[blockIdx.z];
  // ---- level 0 region into shared memory (0 outside the image)
  float* s0 = sm;
  if (P.src_u8) {
    const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
    if ((w0 & 3) == 0) {
      const int R4 = R0 >> 2;
      for (int k = tid; k < R0 * R4; k += 256) {
        const int ly = k / R4, lx = (k - ly * R4) * 4;
        const int x = ox + lx, y = oy + ly;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < h0 && x >= 0 && x + 3 < w0) {
          const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(S8 + x + (size_t)y * w0));
          v = make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
        }
        *reinterpret_cast<float4*>(s0 + ly * R0 + lx) = v;
      }
    } else {
      for (int k = tid; k < R0 * R0; k += 256) {
        const int ly = k / R0, lx = k - ly * R0;
        const int x = ox + lx, y = oy + ly;
        s0[k] = (y >= 0 && y < h0 && x >= 0 && x < w0) ? (float)__ldg(S8 + x + (size_t)y * w0) : 0.f;
      }
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
    for (int l = 1; l < L; l++) {
      const int Rd = Rs >> 1;
      for (int k = tid; k < Rd * Rd; k += 256) {
        const int ly = k / Rd, lx = k - ly * Rd;
        const float* q = src + (2 * ly) * Rs + 2 * lx;
        dst[k] = 0.25f * (((q[0] + q[1]) + q[Rs]) + q[Rs + 1]);
      }
      __syncthreads();
      src = dst; Rs = Rd; dst = dst + Rd * Rd;
    }
  }
  // ---- texels + intensity planes of every level from shared memory
  {
    const float* src = s0; int Rs = R0;
    for (int l = 0; l < L; l++) {
      const int wl = P.w[l], hl = P.h[l];
      const int T = kTile >> l, Hl = H0 >> l;           // tile side and halo at this level
      const int tx0 = (blockIdx.x * kTile) >> l, ty0 = (blockIdx.y * kTile) >> l;
      float* __restrict__ Il = base + P.px_offset[l];
      float4* __restrict__ Tl = tbase + P.px_offset[l];
      for (int k = tid; k < T * T; k += 256) {
        const int ly = k / T, lx = k - ly * T;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= wl || y >= hl) continue;
        const float* q = src + (ly + Hl) * Rs + (lx + Hl);
        const float c = q[0];
        float dx = 0.f, dy = 0.f;
        const bool inner = (y >= 1 && y < hl - 1);        // flat idx in [w, w(h-1))
        if (inner) {
          dx = 0.5f * (q[1] - q[-1]);                   // columns 0 and w-1 are redone by pyr_wrap_kernel
          dy = 0.5f * (q[Rs] - q[-Rs]);
        }
        Il[x + (size_t)y * wl] = c;
        Tl[x + (size_t)y * wl] = inner ? make_texel(c, dx, dy, P.use_gamma) : make_float4(c, 0.f, 0.f, 0.f);
      }
      src += Rs * Rs; Rs >>= 1;
    }
  }
}

// The reference differentiates on the flat index, so at x = 0 the left neighbour is (w-1, y-1) and at x = w-1 the right
// neighbour is (0, y+1) (HessianBlocks.cpp:182-184). One thread per (level, row, side); reads the planes written above.
__global__ void __launch_bounds__(128) pyr_wrap_kernel(PyrGeom P, PyrBatch B) {
  const float* __restrict__ base = B.img[blockIdx.z];
  float4* __restrict__ tbase = B.tex[blockIdx.z];
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < P.levels; l++) {
    const int wl = P.w[l], hl = P.h[l];
    const int cnt = 2 * (hl - 2);
    if (k < cnt) {
      const int y = 1 + (k >> 1), side = k & 1;
      const float* __restrict__ I = base + P.px_offset[l];
      const int x = side ? wl - 1 : 0;
      const int idx = x + y * wl;
      const float c = I[idx];
      const float dx = 0.5f * (I[idx + 1] - I[idx - 1]);
      const float dy = 0.5f * (I[idx + wl] - I[idx - wl]);
      tbase[P.px_offset[l] + idx] = make_texel(c, dx, dy, P.use_gamma);
      return;
    }
    k -= cnt;
  }
}

// nb frames, one launch pair. srcs[i]: device pointer to the level-0 source of frames[i] (float, or uint8 when src_u8).
int make_images_batch_launch(sdso_ctx* ctx, int nb, Frame* const* frames, const void* const* srcs, bool src_u8, bool use_hcalib) {
  if (nb <= 0) return SDSO_OK;
  if (nb > kMaxBatch) return fail(ctx, SDSO_E_INVALID, "make_images batch larger than 32");
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
  P.src_u8 = src_u8 ? 1 : 0;
  PyrBatch B;
  for (int i = 0; i < kMaxBatch; i++) { B.src[i] = nullptr; B.img[i] = nullptr; B.tex[i] = nullptr; }
  for (int i = 0; i < nb; i++) { B.src[i] = srcs[i]; B.img[i] = frames[i]->image; B.tex[i] = frames[i]->tex[0]; }
  prof_begin(ctx, 1);
  {
    const int H0 = 1 << (P.levels - 1);
    size_t smem = 0;
    for (int l = 0, R = kTile + 2 * H0; l < P.levels; l++, R >>= 1) smem += (size_t)R * R * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) { SDSO_CUDA(ctx, cudaFuncSetAttribute(pyr_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
    dim3 grid((P.w[0] + kTile - 1) / kTile, (P.h[0] + kTile - 1) / kTile, nb);
    pyr_fused_kernel<<<grid, 256, smem, ctx->stream>>>(P, B);
    SDSO_CHECK_LAUNCH(ctx);
    int rows = 0;
    for (int l = 0; l < P.levels; l++) rows += 2 * (P.h[l] - 2);
    pyr_wrap_kernel<<<dim3((rows + 127) / 128, 1, nb), 128, 0, ctx->stream>>>(P, B);
    SDSO_CHECK_LAUNCH(ctx);
  }
  prof_end(ctx, 1);
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
