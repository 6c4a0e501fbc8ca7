// Micro-benchmark: what HBM delivers for the tracker's access pattern on B200 — 2x2 patches of 16-byte texels at scattered
// positions of many different 1232x368 images (no arithmetic besides the sum). Sweeps resident warps per SM and the number of
// patches in flight per thread. Output: GB/s of algorithmic bytes (64 B per patch), the unit of bench.py's roofline.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_microbench tools/gather_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int W = 1232, H = 368, NIMG = 592;
// LAYOUT 0: row-major; 1: 4x4-texel tiles (256 B); 2: 8x8 tiles (1 KB); 3: 16x4 tiles (1 KB)
template <int LAYOUT>
__device__ __forceinline__ int tidx(int x, int y) {
  if (LAYOUT == 0) return x + y * W;
  if (LAYOUT == 1) return (((y >> 2) * (W >> 2) + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3);
  if (LAYOUT == 2) return (((y >> 3) * (W >> 3) + (x >> 3)) << 6) + ((y & 7) << 3) + (x & 7);
  return (((y >> 2) * (W >> 4) + (x >> 4)) << 6) + ((y & 3) << 4) + (x & 15);
}
template <int U, int LAYOUT>
__global__ void k(const float4* __restrict__ tex, float* out, int patches_per_thread, int spacing) {
  const int img = blockIdx.x % NIMG;
  const float4* base = tex + (size_t)img * W * H;
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  // a thread's patches walk through the image like projected template points: lane-neighbours ~`spacing` pixels apart
  for (int it = 0; it < patches_per_thread; it += U) {
    float4 t[U][4];
#pragma unroll
    for (int u = 0; u < U; u++) {
      s = s * 1664525u + 1013904223u;
      const int p = ((it + u) * blockDim.x + threadIdx.x) * spacing + (s >> 28);   // raster position with jitter
      const int x = 3 + p % (W - 8), y = 3 + (p / (W - 8)) % (H - 8);
      t[u][0] = __ldg(base + tidx<LAYOUT>(x, y)); t[u][1] = __ldg(base + tidx<LAYOUT>(x + 1, y));
      t[u][2] = __ldg(base + tidx<LAYOUT>(x, y + 1)); t[u][3] = __ldg(base + tidx<LAYOUT>(x + 1, y + 1));
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += t[u][0].x + t[u][1].y + t[u][2].z + t[u][3].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
static bool g_quiet = false;
static double g_last = 0;
template <int U, int LAYOUT>
void run(const float4* tex, float* out, int ctas_per_sm, int threads, int spacing) {
  int sms = 148, grid = sms * ctas_per_sm, ppt = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<U, LAYOUT><<<grid, threads>>>(tex, out, ppt, spacing);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<U, LAYOUT><<<grid, threads>>>(tex, out, ppt, spacing);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double bytes = (double)grid * threads * ppt * 64.0;
  g_last = bytes / ms / 1e6;
  if (!g_quiet) printf("layout=%d U=%d warps/SM=%2d spacing=%2d: %.3f ms  %.0f GB/s algorithmic\n", LAYOUT, U, ctas_per_sm * threads / 32, spacing, ms, bytes / ms / 1e6);
}
int main(int argc, char** argv) {
  float4* tex; float* out;
  size_t n = (size_t)NIMG * W * H;
  if (cudaMalloc(&tex, n * sizeof(float4)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(tex, 0, n * sizeof(float4)); cudaMalloc(&out, 4);
  if (argc > 1) {   // --json: the tracker's own configuration (row-major texels, one patch in flight, 16 warps / SM, ~15 px apart)
    g_quiet = true;
    run<1, 0>(tex, out, 2, 256, 15);
    printf("{\"gather_gbs\": %.1f, \"pattern\": \"2x2 patches of 16-B texels, lanes ~15 px apart, 592 images of 1232x368, 16 warps/SM, no arithmetic\"}\n", g_last);
    return 0;
  }
  for (int spacing : {15, 40}) {
    run<1, 0>(tex, out, 2, 256, spacing); run<1, 1>(tex, out, 2, 256, spacing); run<1, 2>(tex, out, 2, 256, spacing); run<1, 3>(tex, out, 2, 256, spacing);
    run<2, 0>(tex, out, 4, 256, spacing); run<2, 1>(tex, out, 4, 256, spacing); run<2, 2>(tex, out, 4, 256, spacing); run<2, 3>(tex, out, 4, 256, spacing);
  }
  return 0;
}
