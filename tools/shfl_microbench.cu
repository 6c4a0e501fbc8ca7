#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, float seed, int n) {
  __shared__ float sm[64];
  __shared__ double smd[64];
  float a = seed + threadIdx.x;
  long long t0, t1;
  t0 = clock64();
  for (int i = 0; i < n; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0f;
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) a = __shfl_sync(0xffffffffu, a, i & 7) + 1.0f;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  double d = a;
  t0 = clock64();
  for (int i = 0; i < n; i++) d = __shfl_sync(0xffffffffu, d, i & 7) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // smem broadcast round trip
  t0 = clock64();
  for (int i = 0; i < n; i++) { sm[threadIdx.x & 31] = a; __syncwarp(); a = sm[i & 7] + 1.0f; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) { smd[threadIdx.x & 31] = d; __syncwarp(); d = smd[i & 7] + 1.0; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) d = fma(d, 1.0000001, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) a = fmaf(a, 1.0000001f, 1e-9f);
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) d = sqrt(d + 2.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) { float q = (float)d; d = (double)(q + 1.0f); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) a = 1.0f / (a + 1.5f);
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
  out[threadIdx.x] = a + (float)d;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, 64 * sizeof(long long));
  int n = 512;
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, cyc, 1.0f, n); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 1.0f, n); cudaDeviceSynchronize();
    long long h[10]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[10] = {"shfl_xor32+add", "shfl_idx32+add", "shfl_idx64+add", "smem bcast f32", "smem bcast f64", "DFMA dep", "FFMA dep", "DSQRT dep", "cvt f64->f32->f64", "FDIV dep"};
    printf("threads=%d:", threads);
    for (int i = 0; i < 10; i++) printf(" %s=%.1f", names[i], (double)h[i] / n);
    printf("\n");
  }
  return 0;
}
