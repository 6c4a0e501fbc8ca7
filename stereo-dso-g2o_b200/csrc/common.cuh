// Shared device/host definitions of the B200 photometric hot path.
// Built with -fmad=false: every float expression on a decision path (border tests, Huber switch,
// saturation, status enums) is evaluated un-fused, in the reference's operand order, so integer /
// index / status results are bit-identical to an IEEE un-fused CPU evaluation. Where fusion is
// harmless (Hessian accumulation) the code asks for it explicitly with __fmaf_rn.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace sdso {

constexpr int kPyrLevels = 6;     // util/settings.h:46
constexpr int kPatternNum = 8;    // util/settings.h:177
constexpr int kCPARS = 4;         // util/NumType.h:47
// FullSystem/HessianBlocks.h:54-61
constexpr float SCALE_IDEPTH = 1.0f, SCALE_XI_ROT = 1.0f, SCALE_XI_TRANS = 0.5f, SCALE_F = 50.0f, SCALE_C = 50.0f,
                SCALE_A = 10.0f, SCALE_B = 1000.0f;

// util/settings.cpp:216 — the 8-pixel residual pattern (dx,dy)
__device__ __constant__ static const int kPatternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};
static const int kPatternP_host[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// One pyramid level of one frame: texel = {I, dx, dy, absSquaredGrad} (HessianBlocks.h:107-109 fused)
struct LevelView {
  const float4* tex;
  int w, h;
};

// util/globalFuncs.h:73-86 getInterpolatedElement33 on float4 texels. Four 16-byte loads through the
// read-only path; weights and summation order exactly as the reference writes them:
//   dxdy*bp[1+w] + (dy-dxdy)*bp[w] + (dx-dxdy)*bp[1] + (1-dx-dy+dxdy)*bp[0]
__device__ __forceinline__ float3 interp33(const float4* __restrict__ tex, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float4* bp = tex + ix + iy * width;
  float4 t00 = __ldg(bp), t10 = __ldg(bp + 1), t01 = __ldg(bp + width), t11 = __ldg(bp + 1 + width);
  float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  float3 r;
  r.x = w11 * t11.x + w01 * t01.x + w10 * t10.x + w00 * t00.x;
  r.y = w11 * t11.y + w01 * t01.y + w10 * t10.y + w00 * t00.y;
  r.z = w11 * t11.z + w01 * t01.z + w10 * t10.z + w00 * t00.z;
  return r;
}

// util/globalFuncs.h:122-135 getInterpolatedElement31 on the plain intensity plane (4 B/px)
__device__ __forceinline__ float interp31(const float* __restrict__ I, float x, float y, int width) {
  int ix = (int)x, iy = (int)y;
  float dx = x - ix, dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = I + ix + iy * width;
  return dxdy * __ldg(bp + 1 + width) + (dy - dxdy) * __ldg(bp + width) + (dx - dxdy) * __ldg(bp + 1) + (1 - dx - dy + dxdy) * __ldg(bp);
}

// util/globalFuncs.h:160-184 getInterpolatedElement33BiLin (value + interpolated finite differences)
__device__ __forceinline__ float3 interp33BiLin(const float4* __restrict__ tex, float x, float y, int width) {
  if (x == -1 || y == -1) return make_float3(0.f, 0.f, 0.f);
  int ix = (int)x, iy = (int)y;
  const float4* bp = tex + ix + iy * width;
  float tl = __ldg(bp).x, tr = __ldg(bp + 1).x, bl = __ldg(bp + width).x, br = __ldg(bp + width + 1).x;
  float dx = x - ix, dy = y - iy;
  float topInt = dx * tr + (1 - dx) * tl;
  float botInt = dx * br + (1 - dx) * bl;
  float leftInt = dy * bl + (1 - dy) * tl;
  float rightInt = dy * br + (1 - dy) * tr;
  return make_float3(dx * rightInt + (1 - dx) * leftInt, rightInt - leftInt, botInt - topInt);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace sdso
