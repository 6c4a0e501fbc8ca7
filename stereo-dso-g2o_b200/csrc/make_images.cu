// A1 — FrameHessian::makeImages (FullSystem/HessianBlocks.cpp:141-203) on the device.
//
// Two launches per image, both pure streaming (HBM-bound):
//   pyr_down_kernel  : one CTA per 64x64 level-0 tile; the tile is reduced through ALL coarser levels
//                      in shared memory (2x2 box mean, summed in the reference's order
//                      0.25f*(((a+b)+c)+d), :172-178), so level l>0 never re-reads level l-1 from DRAM.
//   gradient_kernel  : one thread per pixel of EVERY level (flattened index space); central
//                      differences on the flat index with the reference's row wrap (:182-184),
//                      non-finite -> 0, absSquaredGrad (+ gamma factor, :196-200), one coalesced
//                      16-byte store {I,dx,dy,absSquaredGrad} per pixel.
// Algorithmic bytes per image: read 4*W*H + write 16*sum_l(w_l*h_l)  (SURVEY.md §8d).
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

__global__ void __launch_bounds__(256) pyr_down_kernel(PyrGeom P, PyrBatch B) {
  // level-1 tile 32x32, level-2 16x16, ... in shared memory
  __shared__ float s1[32][33];
  __shared__ float s2[16][17];
  __shared__ float s3[8][9];
  __shared__ float s4[4][5];
  __shared__ float s5[2][3];
  const int tx0 = blockIdx.x * kTile, ty0 = blockIdx.y * kTile;  // level-0 origin of this tile
  const int tid = threadIdx.x;
  const int w0 = P.w[0], h0 = P.h[0];
  float* __restrict__ base = B.img[blockIdx.z];
  const float* __restrict__ I0 = P.src_u8 ? base : (const float*)B.src[blockIdx.z];
  if (P.src_u8 || (const float*)B.src[blockIdx.z] != base) {
    // widen / copy the level-0 tile into the frame's own intensity plane (read again by the gradient kernel and the epipolar search)
    if (P.src_u8) {
      const unsigned char* __restrict__ S8 = (const unsigned char*)B.src[blockIdx.z];
      for (int k = tid; k < kTile * kTile / 4; k += 256) {
        const int lx = (k & 15) * 4, ly = k >> 4;
        const int x = tx0 + lx, y = ty0 + ly;
        if (y < h0 && x < w0) {
          const size_t o = x + (size_t)y * w0;
          if (x + 3 < w0 && (o & 3) == 0) {
            const uchar4 v = *reinterpret_cast<const uchar4*>(S8 + o);
            *reinterpret_cast<float4*>(base + o) = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
          } else {
            for (int q = 0; q < 4 && x + q < w0; q++) base[o + q] = (float)S8[o + q];
          }
        }
      }
    } else {
      const float* __restrict__ SF = (const float*)B.src[blockIdx.z];
      for (int k = tid; k < kTile * kTile; k += 256) {
        const int x = tx0 + (k & 63), y = ty0 + (k >> 6);
        if (y < h0 && x < w0) base[x + (size_t)y * w0] = SF[x + (size_t)y * w0];
      }
    }
    __syncthreads();  // level 1 below reads this CTA's own tile only
  }
  if (P.levels > 1) {
    const int w1 = P.w[1], h1 = P.h[1];
    float* __restrict__ O = base + P.px_offset[1];
    for (int k = tid; k < 32 * 32; k += 256) {
      int lx = k & 31, ly = k >> 5;
      int x = (tx0 >> 1) + lx, y = (ty0 >> 1) + ly;
      float v = 0.f;
      if (x < w1 && y < h1) {
        const float2 top = *reinterpret_cast<const float2*>(I0 + 2 * x + (size_t)(2 * y) * w0);
        const float2 bot = *reinterpret_cast<const float2*>(I0 + 2 * x + (size_t)(2 * y + 1) * w0);
        v = 0.25f * (((top.x + top.y) + bot.x) + bot.y);
        O[x + (size_t)y * w1] = v;
      }
      s1[ly][lx] = v;
    }
  }
  __syncthreads();
#define SDSO_DOWN(LVL, SRC, DST, N)                                                          \
  if (P.levels > LVL) {                                                                      \
    const int wl = P.w[LVL], hl = P.h[LVL];                                                  \
    float* __restrict__ Ol = base + P.px_offset[LVL];                                        \
    for (int k = tid; k < N * N; k += 256) {                                                 \
      int lx = k % N, ly = k / N;                                                            \
      int x = (tx0 >> LVL) + lx, y = (ty0 >> LVL) + ly;                                      \
      float v = 0.25f * (((SRC[2 * ly][2 * lx] + SRC[2 * ly][2 * lx + 1]) + SRC[2 * ly + 1][2 * lx]) + SRC[2 * ly + 1][2 * lx + 1]); \
      DST[ly][lx] = v;                                                                       \
      if (x < wl && y < hl) Ol[x + (size_t)y * wl] = v;                                      \
    }                                                                                        \
  }                                                                                          \
  __syncthreads();
  SDSO_DOWN(2, s1, s2, 16)
  SDSO_DOWN(3, s2, s3, 8)
  SDSO_DOWN(4, s3, s4, 4)
  SDSO_DOWN(5, s4, s5, 2)
#undef SDSO_DOWN
}

__global__ void __launch_bounds__(256) gradient_kernel(PyrGeom P, PyrBatch B) {
  const int total = P.px_offset[P.levels];
  const float* __restrict__ Ibase = B.img[blockIdx.z];
  float4* __restrict__ Tbase = B.tex[blockIdx.z];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < kPyrLevels; l++) if (l < P.levels && g >= P.px_offset[l]) lvl = l;
    const int idx = g - P.px_offset[lvl];
    const int wl = P.w[lvl], hl = P.h[lvl];
    const float* __restrict__ I = Ibase + P.px_offset[lvl];
    const float c = I[idx];
    float dx = 0.f, dy = 0.f, ag = 0.f;
    if (idx >= wl && idx < wl * (hl - 1)) {
      dx = 0.5f * (I[idx + 1] - I[idx - 1]);
      dy = 0.5f * (I[idx + wl] - I[idx - wl]);
      if (!isfinite(dx)) dx = 0.f;
      if (!isfinite(dy)) dy = 0.f;
      ag = dx * dx + dy * dy;
      if (P.use_gamma) {
        // CalibHessian::getBGradOnly (HessianBlocks.h:356-362)
        int ci = (int)(c + 0.5f);
        if (ci < 5) ci = 5;
        if (ci > 250) ci = 250;
        float gw = g_Bgamma[ci + 1] - g_Bgamma[ci];
        ag *= gw * gw;
      }
    }
    Tbase[g] = make_float4(c, dx, dy, ag);
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
    dim3 grid((P.w[0] + kTile - 1) / kTile, (P.h[0] + kTile - 1) / kTile, nb);
    pyr_down_kernel<<<grid, 256, 0, ctx->stream>>>(P, B);
    SDSO_CHECK_LAUNCH(ctx);
  }
  {
    int blocks = (off + 255) / 256;
    int cap = (ctx->num_sms * 8 + nb - 1) / nb;
    if (cap < 64) cap = 64;
    if (blocks > cap) blocks = cap;
    gradient_kernel<<<dim3(blocks, 1, nb), 256, 0, ctx->stream>>>(P, B);
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
