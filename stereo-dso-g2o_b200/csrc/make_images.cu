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

struct PyrParams {
  int levels;
  int w[kPyrLevels], h[kPyrLevels];
  float* I[kPyrLevels];     // intensity planes (I[0] = input image)
  float4* tex[kPyrLevels];  // output texels
  int px_offset[kPyrLevels + 1];  // prefix sum of w_l*h_l
  int use_gamma;
};

__device__ float g_Bgamma[256];  // CalibHessian::B (HessianBlocks.h:352), identity unless sdso_set_gamma

constexpr int kTile = 64;

__global__ void __launch_bounds__(256) pyr_down_kernel(PyrParams P) {
  // level-1 tile 32x32, level-2 16x16, ... in shared memory
  __shared__ float s1[32][33];
  __shared__ float s2[16][17];
  __shared__ float s3[8][9];
  __shared__ float s4[4][5];
  __shared__ float s5[2][3];
  const int tx0 = blockIdx.x * kTile, ty0 = blockIdx.y * kTile;  // level-0 origin of this tile
  const int tid = threadIdx.x;
  const float* __restrict__ I0 = P.I[0];
  const int w0 = P.w[0];
  if (P.levels > 1) {
    const int w1 = P.w[1], h1 = P.h[1];
    float* __restrict__ O = P.I[1];
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
    for (int k = tid; k < N * N; k += 256) {                                                 \
      int lx = k % N, ly = k / N;                                                            \
      int x = (tx0 >> LVL) + lx, y = (ty0 >> LVL) + ly;                                      \
      float v = 0.25f * (((SRC[2 * ly][2 * lx] + SRC[2 * ly][2 * lx + 1]) + SRC[2 * ly + 1][2 * lx]) + SRC[2 * ly + 1][2 * lx + 1]); \
      DST[ly][lx] = v;                                                                       \
      if (x < wl && y < hl) P.I[LVL][x + (size_t)y * wl] = v;                                \
    }                                                                                        \
  }                                                                                          \
  __syncthreads();
  SDSO_DOWN(2, s1, s2, 16)
  SDSO_DOWN(3, s2, s3, 8)
  SDSO_DOWN(4, s3, s4, 4)
  SDSO_DOWN(5, s4, s5, 2)
#undef SDSO_DOWN
}

__global__ void __launch_bounds__(256) gradient_kernel(PyrParams P) {
  const int total = P.px_offset[P.levels];
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < kPyrLevels; l++) if (l < P.levels && g >= P.px_offset[l]) lvl = l;
    const int idx = g - P.px_offset[lvl];
    const int wl = P.w[lvl], hl = P.h[lvl];
    const float* __restrict__ I = P.I[lvl];
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
    P.tex[lvl][idx] = make_float4(c, dx, dy, ag);
  }
}

int make_images_launch(sdso_ctx* ctx, Frame& f, const float* dev_image, bool use_hcalib) {
  PyrParams P;
  P.levels = ctx->G.levels;
  int off = 0;
  for (int l = 0; l < P.levels; l++) {
    P.w[l] = ctx->G.w[l]; P.h[l] = ctx->G.h[l];
    P.tex[l] = f.tex[l];
    P.px_offset[l] = off;
    off += P.w[l] * P.h[l];
  }
  P.px_offset[P.levels] = off;
  for (int l = P.levels; l < kPyrLevels; l++) { P.w[l] = P.h[l] = 0; P.tex[l] = nullptr; P.I[l] = nullptr; P.px_offset[l + 1] = off; }
  P.I[0] = const_cast<float*>(dev_image);
  // coarser intensity planes live behind the level-0 plane in the frame's image buffer
  {
    float* p = f.image + (size_t)P.w[0] * P.h[0];
    for (int l = 1; l < P.levels; l++) { P.I[l] = p; p += (size_t)P.w[l] * P.h[l]; }
  }
  P.use_gamma = (use_hcalib && ctx->S.gammaWeightsPixelSelect == 1) ? 1 : 0;
  prof_begin(ctx, 1);
  if (P.levels > 1) {
    dim3 grid((P.w[0] + kTile - 1) / kTile, (P.h[0] + kTile - 1) / kTile);
    pyr_down_kernel<<<grid, 256, 0, ctx->stream>>>(P);
    SDSO_CHECK_LAUNCH(ctx);
  }
  {
    int total = off;
    int blocks = (total + 255) / 256;
    int cap = ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    gradient_kernel<<<blocks, 256, 0, ctx->stream>>>(P);
    SDSO_CHECK_LAUNCH(ctx);
  }
  prof_end(ctx, 1);
  return SDSO_OK;
}

int set_gamma_table(sdso_ctx* ctx, const float B[256]) {
  SDSO_CUDA(ctx, cudaMemcpyToSymbolAsync(g_Bgamma, B, 256 * sizeof(float), 0, cudaMemcpyHostToDevice, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

}  // namespace sdso
