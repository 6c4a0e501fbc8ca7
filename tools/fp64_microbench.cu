// Micro-benchmark: FP64 latency / throughput of one warp on B200 (informs the LM-step design).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) a = fma(a, b, c);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 8 independent chains
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = a + j;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = fma(x[j], b, c);
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  for (int j = 0; j < 8; j++) a += x[j];
  // dependent divisions
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) a = 1.0 / (a + 1.5);
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // dependent float FMA chain
  float f = (float)a, g = 1.0000001f, h = 1e-9f;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) f = fmaf(f, g, h);
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // conversions f64<->f32 dependent
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { float q = (float)a; a = (double)q + 1e-3; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // sqrt
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; i++) a = sqrt(a + 2.0);
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // double shuffle chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) a = __shfl_xor_sync(0xffffffffu, a, 1);
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // exp
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; i++) a = exp(a * 1e-3);
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = a + f;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&cyc, 64 * sizeof(long long));
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d: DFMA dep %.1f cyc/op | DFMA 8-indep %.1f cyc/op | DDIV dep %.1f | FFMA dep %.1f | cvt pair+add %.1f | DSQRT %.1f | shfl64 %.1f | exp %.1f\n",
           threads, h[0] / 256.0, h[1] / 512.0, h[2] / 64.0, h[3] / 256.0, h[4] / 64.0, h[5] / 32.0, h[6] / 64.0, h[7] / 16.0);
  }
  return 0;
}
