// V1-V5 and E3 as operators over SoA batches (SURVEY.md §8b layer 2): the g2o vertices' oplusImpl
// (FullSystem/dso_g2o_vertex.cpp:15-18, 30-40, 56-58, 73-88, 100-106) and EdgeTracePointUVDSO::computeError / linearizeOplus
// (FullSystem/dso_g2o_edge.cpp:571-619), each as ONE launch over n vertices / edges. The fused kernels (tracker, traceStereo
// refinement, LBA driver) embed the same updates; these entry points are the operator-level surface a g2o-style caller binds.
#include "ctx.h"

namespace sdso {

// exp([upsilon; omega]) * T in double (thirdparty/Sophus/sophus/se3.hpp:407-428: quaternion exponential + V matrix, small-angle
// branch below 1e-10). T row-major 3x4.
__device__ void d_se3_exp_left(const double a[6], double T[12]) {
  const double wx = a[3], wy = a[4], wz = a[5];
  const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
  double A, B, C;   // sin t / t, (1 - cos t) / t^2, (t - sin t) / t^3
  if (th < 1e-10) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0; }
  else { double s, c; sincos(th, &s, &c); A = s / th; B = (1.0 - c) / th2; C = (th - s) / (th2 * th); }
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) O2[r * 3 + c] = O[r * 3] * O[c] + O[r * 3 + 1] * O[3 + c] + O[r * 3 + 2] * O[6 + c];
  double R[9], V[9];
  for (int i = 0; i < 9; i++) { const double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + A * O[i] + B * O2[i]; V[i] = I + B * O[i] + C * O2[i]; }
  double te[3];
  for (int r = 0; r < 3; r++) te[r] = V[r * 3] * a[0] + V[r * 3 + 1] * a[1] + V[r * 3 + 2] * a[2];
  double out[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[r * 4 + c] = R[r * 3] * T[c] + R[r * 3 + 1] * T[4 + c] + R[r * 3 + 2] * T[8 + c];
    out[r * 4 + 3] = R[r * 3] * T[3] + R[r * 3 + 1] * T[7] + R[r * 3 + 2] * T[11] + te[r];
  }
  for (int i = 0; i < 12; i++) T[i] = out[i];
}

__global__ void vertex_oplus_kernel(int kind, int n, double* __restrict__ est, const double* __restrict__ upd, const double* __restrict__ aux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (kind) {
    case SDSO_VERTEX_SE3_POSE: {  // _estimate = SE3::exp(update) * estimate()  (dso_g2o_vertex.cpp:15-18)
      double T[12], a[6];
      for (int k = 0; k < 12; k++) T[k] = est[(size_t)12 * i + k];
      for (int k = 0; k < 6; k++) a[k] = upd[(size_t)6 * i + k];
      d_se3_exp_left(a, T);
      for (int k = 0; k < 12; k++) est[(size_t)12 * i + k] = T[k];
      break;
    }
    case SDSO_VERTEX_PHOTOMETRIC:  // a += update(0); b += update(1)  (:30-40)
      est[2 * i] += upd[2 * i]; est[2 * i + 1] += upd[2 * i + 1];
      break;
    case SDSO_VERTEX_INVERSE_DEPTH:  // _estimate += *update_  (:56-58)
      est[i] += upd[i];
      break;
    case SDSO_VERTEX_UV: {  // clamp to +-0.5, non-finite -> 0, uv += update * (dx_, dy_)  (:73-88)
      double u = upd[i];
      if (u < -0.5) u = -0.5;
      else if (u > 0.5) u = 0.5;
      else if (!isfinite(u)) u = 0;
      est[2 * i] += u * aux[2 * i]; est[2 * i + 1] += u * aux[2 * i + 1];
      break;
    }
    case SDSO_VERTEX_CAM:  // fx, fy, cx, cy += update  (:100-106)
      for (int k = 0; k < 4; k++) est[4 * i + k] += upd[4 * i + k];
      break;
  }
}

// EdgeTracePointUVDSO (dso_g2o_edge.cpp:571-619). flag: 1 = error and Jacobian written; 0 = util::CheckBoundary failed (error set to
// 0, Jacobian left as it was); 2 = non-finite intensity (both left as they were).
__global__ void edge_trace_uv_kernel(const float4* __restrict__ tex0, int w0, int h0, int n, const double* __restrict__ uv,
                                     const float* __restrict__ rot, const double* __restrict__ meas, float a0, float a1,
                                     const double* __restrict__ dxdy, double* __restrict__ err, double* __restrict__ J, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double bu = uv[2 * i], bv = uv[2 * i + 1];
  // util::CheckBoundary(u, v, wG[0]-3, hG[0]-3) (dso_util.hpp:29): u-2 < 0 || u+3 > w-3 ... evaluated in double as written
  const double wM = (double)(w0 - 3), hM = (double)(h0 - 3);
  if (bu - 2 < 0 || bu + 3 > wM || bv - 2 < 0 || bv + 3 > hM) { err[i] = 0.0; if (flag) flag[i] = 0; return; }
  const float3 hit = interp33(tex0, (float)(bu + rot[2 * i]), (float)(bv + rot[2 * i + 1]), w0);
  if (!isfinite(hit.x)) { if (flag) flag[i] = 2; return; }
  err[i] = (double)hit.x - ((double)a0 * meas[i] + (double)a1);   // Vec2f * double promotes to double (:595)
  J[i] = dxdy[2 * i] * (double)hit.y + dxdy[2 * i + 1] * (double)hit.z;
  if (flag) flag[i] = 1;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

int sdso_vertex_oplus(sdso_ctx* ctx, int kind, int n, double* estimate, const double* update, const double* aux) {
  sdso::enter(ctx);
  if (!ctx || n < 0 || (n > 0 && (!estimate || !update))) return SDSO_E_INVALID;
  int es = 0, us = 0;
  switch (kind) {
    case SDSO_VERTEX_SE3_POSE: es = 12; us = 6; break;
    case SDSO_VERTEX_PHOTOMETRIC: es = 2; us = 2; break;
    case SDSO_VERTEX_INVERSE_DEPTH: es = 1; us = 1; break;
    case SDSO_VERTEX_UV: es = 2; us = 1; if (n > 0 && !aux) return fail(ctx, SDSO_E_INVALID, "VertexUVDSO needs aux = (dx_, dy_) per vertex"); break;
    case SDSO_VERTEX_CAM: es = 4; us = 4; break;
    default: return fail(ctx, SDSO_E_INVALID, "unknown vertex kind");
  }
  if (n == 0) return SDSO_OK;
  double *d_e = nullptr, *d_u = nullptr, *d_a = nullptr;
  cudaStream_t st = ctx->stream;
  SDSO_CUDA(ctx, sdso::alloc_async(ctx, &d_e, (size_t)n * es * sizeof(double), st));
  SDSO_CUDA(ctx, sdso::alloc_async(ctx, &d_u, (size_t)n * us * sizeof(double), st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_e, estimate, (size_t)n * es * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_u, update, (size_t)n * us * sizeof(double), cudaMemcpyHostToDevice, st));
  if (kind == SDSO_VERTEX_UV) {
    SDSO_CUDA(ctx, sdso::alloc_async(ctx, &d_a, (size_t)n * 2 * sizeof(double), st));
    SDSO_CUDA(ctx, cudaMemcpyAsync(d_a, aux, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  vertex_oplus_kernel<<<(n + 127) / 128, 128, 0, st>>>(kind, n, d_e, d_u, d_a);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(estimate, d_e, (size_t)n * es * sizeof(double), cudaMemcpyDeviceToHost, st));
  SDSO_CUDA(ctx, cudaFreeAsync(d_e, st));
  SDSO_CUDA(ctx, cudaFreeAsync(d_u, st));
  if (d_a) SDSO_CUDA(ctx, cudaFreeAsync(d_a, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  return SDSO_OK;
}

int sdso_edge_trace_uv_eval(sdso_ctx* ctx, int frame, int n, const double* uv, const float* rotatePattern, const double* measurement,
                            const float affLL[2], const double* dxdy, double* error, double* J, int* flag) {
  sdso::enter(ctx);
  if (!ctx || n < 0 || !affLL || (n > 0 && (!uv || !rotatePattern || !measurement || !dxdy || !error || !J))) return SDSO_E_INVALID;
  if (frame < 0 || frame >= (int)ctx->frames.size() || !ctx->frames[frame].in_use || !ctx->frames[frame].valid)
    return fail(ctx, SDSO_E_INVALID, "bad frame id (not created or makeImages not run)");
  if (n == 0) return SDSO_OK;
  cudaStream_t st = ctx->stream;
  // one staging block: uv (2n d) | meas (n d) | dxdy (2n d) | err (n d) | J (n d) | rot (2n f) | flag (n i)
  const size_t nd = (size_t)7 * n * sizeof(double), nf = (size_t)2 * n * sizeof(float), ni = (size_t)n * sizeof(int);
  char* d = nullptr;
  SDSO_CUDA(ctx, sdso::alloc_async(ctx, &d, nd + nf + ni, st));
  double* d_uv = reinterpret_cast<double*>(d); double* d_me = d_uv + 2 * (size_t)n; double* d_dx = d_me + n; double* d_er = d_dx + 2 * (size_t)n; double* d_J = d_er + n;
  float* d_rot = reinterpret_cast<float*>(d + nd); int* d_fl = reinterpret_cast<int*>(d + nd + nf);
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_uv, uv, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_me, measurement, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_dx, dxdy, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_er, error, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));   // stale values survive (see flag)
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_J, J, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(d_rot, rotatePattern, nf, cudaMemcpyHostToDevice, st));
  edge_trace_uv_kernel<<<(n + 127) / 128, 128, 0, st>>>(ctx->frames[frame].tex[0], ctx->G.w[0], ctx->G.h[0], n, d_uv, d_rot, d_me, affLL[0], affLL[1],
                                                        d_dx, d_er, d_J, d_fl);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(error, d_er, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  SDSO_CUDA(ctx, cudaMemcpyAsync(J, d_J, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (flag) SDSO_CUDA(ctx, cudaMemcpyAsync(flag, d_fl, ni, cudaMemcpyDeviceToHost, st));
  SDSO_CUDA(ctx, cudaFreeAsync(d, st));
  SDSO_CUDA(ctx, cudaStreamSynchronize(st));
  return SDSO_OK;
}

}  // extern "C"
