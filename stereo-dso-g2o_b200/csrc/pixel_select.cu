// Candidate pixel selection on the device (FullSystem/PixelSelector2.cpp: makeHists :84-178, select :340-536,
// makeMaps :192-327; call site FullSystem::makeNewTraces, FullSystem.cpp:1599-1625).
//
// The reference walks the image once, sequentially, in nested 4pot / 2pot / pot blocks; the direction a pot-cell
// projects its gradients on is randomPattern[n2] & 0xF with n2 the RUNNING count of cells that selected a level-0
// pixel before it, and a 2pot (4pot) block selects a level-1 (level-2) pixel only if nothing finer fired inside it.
// The device path removes the sequential walk without changing one result:
//   1. sel_cell_mask: one thread per pot-cell — 16-bit mask "this cell selects under direction k" (some pixel above the
//      level-0 threshold has a non-zero projection on direction k); stored in the reference's visiting order.
//   2. sel_scan (one CTA): cells whose mask is 0xFFFF (0) select (do not select) whatever n2 is, so n2 is an exclusive
//      scan of the certain cells plus a short serial walk over the few ambiguous cells (gradient exactly orthogonal to
//      some direction), which is the only part that really depends on the running count.
//   3. sel_level0: one thread per cell — arg-max of |g . dir2| in raster order (strict >), map = 1.
//   4. sel_level12: one warp per 4pot block — for 2pot blocks without a level-0 selection the level-1 arg-max
//      (map = 2); if nothing fired in the whole block, the level-2 arg-max (map = 4). Inside such blocks n2 does
//      not move, so dir2 == dir3 == dir4 and "fired" reduces to "a pixel passes the threshold with non-zero projection".
//   5. sel_rowcount / sel_subsample / sel_list (one warp per image row): raster-order rank of every selected pixel (ballots),
//      the random sub-sampling of makeMaps (randomPattern[rank] > 255 * quotia drops the pixel) and the compact
//      (x, y, type) list makeNewTraces walks.
// All of it is integer / comparison work on values makeImages produced, so maps and counts are bit-exact.
#include "ctx.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace sdso {

struct SelectorState {
  int w = 0, h = 0, w32 = 0, h32 = 0, ths_alloc = 0;
  int currentPotential = 3;
  int histFrame = -1; unsigned histGen = 0;  // gradHistFrame
  std::vector<unsigned char> h_rp;
  unsigned char* d_rp = nullptr;
  float *d_ths = nullptr, *d_thsSmoothed = nullptr, *d_map = nullptr;
  unsigned short* d_mask = nullptr;  // per cell, visiting order
  int *d_cpre = nullptr, *d_apre = nullptr, *d_ambsel = nullptr, *d_amb_cpre = nullptr, *d_n2cell = nullptr;
  unsigned short* d_amb_mask = nullptr;
  unsigned char* d_selc = nullptr;   // per cell, raster order: selected a level-0 pixel
  int *d_rowhave = nullptr, *d_rowkeep = nullptr;  // selected pixels per image row before / after the sub-sampling
  int* d_counts = nullptr;           // 0 n2, 1 n3, 2 n4, 3 namb, 4 numHave (compaction in), 5 numHaveSub (compaction out)
  float *d_list_uv = nullptr, *d_list_type = nullptr;
  int* h_counts = nullptr;           // pinned
  bool list_valid = false;
  int last_n[3] = {0, 0, 0};
};

// glibc random_r TYPE_3 (r[i] = r[i-3] + r[i-31], seeded by the 16807 Lehmer generator, first 310 outputs discarded):
// the sequence srand(seed); rand() produces on the reference's platform.
static void glibc_rand_bytes(unsigned seed, unsigned char* out, size_t n) {
  int32_t r[34];
  r[0] = seed ? (int32_t)seed : 1;
  for (int i = 1; i < 31; i++) {
    const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
    int64_t word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r[i] = (int32_t)word;
  }
  std::vector<uint32_t> s(344 + n);
  for (int i = 0; i < 31; i++) s[i] = (uint32_t)r[i];
  for (int i = 31; i < 34; i++) s[i] = s[i - 31];
  for (size_t i = 34; i < 344 + n; i++) s[i] = s[i - 31] + s[i - 3];
  for (size_t k = 0; k < n; k++) out[k] = (unsigned char)((s[k + 344] >> 1) & 0xFF);
}

__device__ __constant__ static const float kDirections[16][2] = {
    {0, 1.0000f}, {0.3827f, 0.9239f}, {0.1951f, 0.9808f}, {0.9239f, 0.3827f}, {0.7071f, 0.7071f}, {0.3827f, -0.9239f},
    {0.8315f, 0.5556f}, {0.8315f, -0.5556f}, {0.5556f, -0.8315f}, {0.9808f, 0.1951f}, {0.9239f, -0.3827f}, {0.7071f, -0.7071f},
    {0.5556f, 0.8315f}, {0.9808f, -0.1951f}, {1.0000f, 0.0000f}, {0.1951f, -0.9808f}};  // PixelSelector2.cpp:364-380

// ---- makeHists ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sel_hist_kernel(const float4* __restrict__ tex0, int w, int h, int w32, float below, float add, float* ths) {
  __shared__ int hist[100];
  const int bx = blockIdx.x, by = blockIdx.y, t = threadIdx.x;
  if (t < 100) hist[t] = 0;
  __syncthreads();
  for (int k = t; k < 1024; k += 256) {
    const int i = k & 31, j = k >> 5;
    const int it = i + 32 * bx, jt = j + 32 * by;
    if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
    int g = (int)sqrtf(tex0[it + jt * w].w);
    if (g > 48) g = 48;
    atomicAdd(&hist[g + 1], 1);
    atomicAdd(&hist[0], 1);
  }
  __syncthreads();
  if (t == 0) {
    int th = (int)(hist[0] * below + 0.5f);
    int q = 90;
    for (int i = 0; i < 90; i++) { th -= hist[i + 1]; if (th < 0) { q = i; break; } }
    ths[bx + by * w32] = q + add;
  }
}

__global__ void sel_smooth_kernel(const float* __restrict__ ths, float* thsSmoothed, int w32, int h32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w32 * h32) return;
  const int x = i % w32, y = i / w32;
  float sum = 0, num = 0;
  if (x > 0) {
    if (y > 0) { num++; sum += ths[x - 1 + (y - 1) * w32]; }
    if (y < h32 - 1) { num++; sum += ths[x - 1 + (y + 1) * w32]; }
    num++; sum += ths[x - 1 + y * w32];
  }
  if (x < w32 - 1) {
    if (y > 0) { num++; sum += ths[x + 1 + (y - 1) * w32]; }
    if (y < h32 - 1) { num++; sum += ths[x + 1 + (y + 1) * w32]; }
    num++; sum += ths[x + 1 + y * w32];
  }
  if (y > 0) { num++; sum += ths[x + (y - 1) * w32]; }
  if (y < h32 - 1) { num++; sum += ths[x + (y + 1) * w32]; }
  num++; sum += ths[x + y * w32];
  thsSmoothed[i] = (sum / num) * (sum / num);
}

// ---- select ---------------------------------------------------------------------------------------------------------
struct SelParams {
  const float4 *tex0, *tex1, *tex2;
  const float* thsSmoothed;
  const unsigned char* rp;
  int w, h, w1, w2, thsStep, ths_alloc;
  int pot, ncx, ncy;
  float thFactor, dw1, dw2;
};

// position of cell (cx, cy) in the reference's nested 4pot / 2pot / pot walk (PixelSelector2.cpp:392-445)
__device__ __forceinline__ int visit_index(int cx, int cy, int ncx, int ncy) {
  const int bx = cx >> 2, by = cy >> 2;
  const int cw = min(4, ncx - 4 * bx), ch = min(4, ncy - 4 * by);
  const int sx = (cx & 3) >> 1, sy = (cy & 3) >> 1;
  const int sw = min(2, cw - 2 * sx), sh = min(2, ch - 2 * sy);
  return by * 4 * ncx + bx * 4 * ch + sy * 2 * cw + sx * 2 * sh + (cy & 1) * sw + (cx & 1);
}

__device__ __forceinline__ bool sel_inb(int xf, int yf, int w, int h) { return !(xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4); }
__device__ __forceinline__ float sel_th0(const SelParams& P, int xf, int yf) {
  const int i = (xf >> 5) + (yf >> 5) * P.thsStep;
  return i < P.ths_alloc ? P.thsSmoothed[i] : 0.f;
}

__global__ void __launch_bounds__(128) sel_cell_mask_kernel(SelParams P, unsigned short* mask) {
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
  if (cx >= P.ncx) return;
  const int x0 = cx * P.pot, y0 = cy * P.pot;
  const int mx = min(P.pot, P.w - x0), my = min(P.pot, P.h - y0);
  unsigned m = 0;
  for (int y1 = 0; y1 < my; y1++)
    for (int x1 = 0; x1 < mx; x1++) {
      const int xf = x0 + x1, yf = y0 + y1;
      if (!sel_inb(xf, yf, P.w, P.h)) continue;
      const float4 t = P.tex0[xf + P.w * yf];
      if (!(t.w > sel_th0(P, xf, yf) * P.thFactor)) continue;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const float dn = fabsf(t.y * kDirections[k][0] + t.z * kDirections[k][1]);
        if (dn > 0.f) m |= 1u << k;
      }
    }
  mask[visit_index(cx, cy, P.ncx, P.ncy)] = (unsigned short)m;
}

// one CTA of 32 warps: n2 at the start of every cell (visiting order) = #certain cells before + #ambiguous cells before that
// selected. Each warp owns a contiguous slice of the cells and ranks them with ballots (coalesced); warp 0 then resolves the
// ambiguous cells 32 at a time: lane l precomputes, for every possible number j of selections among the lanes before it,
// whether its cell selects (32 bytes of randomPattern), and the true j is then threaded through the 32 words with shuffles.
__global__ void __launch_bounds__(1024) sel_scan_kernel(const unsigned short* __restrict__ mask, int ncells, const unsigned char* __restrict__ rp, int rp_n,
                                                        int* __restrict__ cpre, int* __restrict__ apre, unsigned short* __restrict__ amb_mask,
                                                        int* __restrict__ amb_cpre, int* __restrict__ ambsel, int* __restrict__ counts) {
  __shared__ int s_c[32], s_a[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int per = ((ncells + 31) / 32 + 31) & ~31;
  const int lo = min(ncells, warp * per), hi = min(ncells, lo + per);
  int c = 0, a = 0;
  for (int i = lo + lane; i - lane < hi; i += 32) {
    const unsigned m = i < hi ? mask[i] : 0u;
    c += __popc(__ballot_sync(0xffffffffu, m == 0xFFFFu));
    a += __popc(__ballot_sync(0xffffffffu, m != 0u && m != 0xFFFFu));
  }
  if (lane == 0) { s_c[warp] = c; s_a[warp] = a; }
  __syncthreads();
  int pc = 0, pa = 0, tc = 0, ta = 0;
  for (int k = 0; k < 32; k++) { if (k < warp) { pc += s_c[k]; pa += s_a[k]; } tc += s_c[k]; ta += s_a[k]; }
  for (int i = lo + lane; i - lane < hi; i += 32) {
    const unsigned m = i < hi ? mask[i] : 0u;
    const bool isc = m == 0xFFFFu, isa = m != 0u && m != 0xFFFFu;
    const unsigned bc = __ballot_sync(0xffffffffu, isc), ba = __ballot_sync(0xffffffffu, isa);
    const int myc = pc + __popc(bc & lt), mya = pa + __popc(ba & lt);
    if (i < hi) { cpre[i] = myc; apre[i] = mya; }
    if (isa) { amb_mask[mya] = (unsigned short)m; amb_cpre[mya] = myc; }
    pc += __popc(bc); pa += __popc(ba);
  }
  __syncthreads();
  if (warp != 0) return;
  int s = 0;
  for (int k0 = 0; k0 < ta; k0 += 32) {
    const int k = k0 + lane;
    const unsigned m = k < ta ? amb_mask[k] : 0u;
    const int base = (k < ta ? amb_cpre[k] : 0) + s;
    unsigned W = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {  // (only j <= lane can occur; the fixed trip count keeps the 32 byte loads independent)
      const int idx = min(base + j, rp_n - 1);
      W |= ((m >> (rp[idx] & 0xF)) & 1u) << j;
    }
    int j = 0; unsigned selbits = 0;
#pragma unroll
    for (int l = 0; l < 32; l++) {
      const unsigned Wl = __shfl_sync(0xffffffffu, W, l);
      const unsigned bsel = (Wl >> j) & 1u;
      selbits |= bsel << l; j += bsel;
    }
    if (k < ta) ambsel[k] = s + __popc(selbits & lt);
    s += j;
  }
  if (lane == 0) { ambsel[ta] = s; counts[0] = tc + s; counts[1] = 0; counts[2] = 0; counts[3] = ta; }
}

__global__ void __launch_bounds__(128) sel_level0_kernel(SelParams P, const unsigned short* __restrict__ mask, const int* __restrict__ cpre,
                                                         const int* __restrict__ apre, const int* __restrict__ ambsel, float* map, unsigned char* selc,
                                                         int* n2cell) {
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
  if (cx >= P.ncx) return;
  const int vi = visit_index(cx, cy, P.ncx, P.ncy);
  const int n2 = cpre[vi] + ambsel[apre[vi]];
  const int d = P.rp[n2] & 0xF;
  const bool sel = (mask[vi] >> d) & 1;
  selc[cx + cy * P.ncx] = sel;
  n2cell[cx + cy * P.ncx] = n2;
  if (!sel) return;
  const float d0 = kDirections[d][0], d1 = kDirections[d][1];
  const int x0 = cx * P.pot, y0 = cy * P.pot;
  const int mx = min(P.pot, P.w - x0), my = min(P.pot, P.h - y0);
  int best = -1; float bestVal = 0;
  for (int y1 = 0; y1 < my; y1++)
    for (int x1 = 0; x1 < mx; x1++) {
      const int xf = x0 + x1, yf = y0 + y1;
      if (!sel_inb(xf, yf, P.w, P.h)) continue;
      const float4 t = P.tex0[xf + P.w * yf];
      if (!(t.w > sel_th0(P, xf, yf) * P.thFactor)) continue;
      const float dn = fabsf(t.y * d0 + t.z * d1);
      if (dn > bestVal) { bestVal = dn; best = xf + P.w * yf; }
    }
  if (best > 0) map[best] = 1.f;
}

// One warp per 4pot block. Lanes stride over the block's pixels; every candidate carries (|g . dir| bits, ~visiting key)
// packed in 64 bits, so a warp max returns the largest projection and, among equals, the pixel the reference's walk meets
// first (its strict '>' keeps the first maximum).
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
  return v;
}

__global__ void __launch_bounds__(256) sel_level12_kernel(SelParams P, const unsigned char* __restrict__ selc, const int* __restrict__ n2cell, float* map,
                                                          int* counts) {
  const int lane = threadIdx.x & 31;
  const int nbx = (P.ncx + 3) >> 2, nby = (P.ncy + 3) >> 2;
  const int blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blk >= nbx * nby) return;
  const int bx = blk % nbx, by = blk / nbx;
  const int cw = min(4, P.ncx - 4 * bx), ch = min(4, P.ncy - 4 * by);
  // which cells of the block selected a level-0 pixel (lane = ly4 * 4 + lx4)
  const int lx4 = lane & 3, ly4 = (lane >> 2) & 3;
  const bool csel = lane < 16 && lx4 < cw && ly4 < ch && selc[4 * bx + lx4 + (4 * by + ly4) * P.ncx] != 0;
  const unsigned cb = __ballot_sync(0xffffffffu, csel);
  unsigned subsel = 0, subex = 0, subd = 0;   // per 2pot sub-block: holds a level-0 pixel / exists / direction index (4 bits each)
#pragma unroll
  for (int sb = 0; sb < 4; sb++) {
    const int sx = sb & 1, sy = sb >> 1;
    const unsigned bits = (0x33u << (2 * sx)) << (8 * sy);   // the 2x2 cells of sub-block (sx, sy) in the 4x4 ballot
    const bool ex = 2 * sx < cw && 2 * sy < ch;
    subex |= (unsigned)ex << sb;
    subsel |= (unsigned)((cb & bits) != 0) << sb;
    if (ex) subd |= (unsigned)(P.rp[n2cell[4 * bx + 2 * sx + (4 * by + 2 * sy) * P.ncx]] & 0xF) << (4 * sb);
  }
  if ((subex & ~subsel) == 0) return;   // every 2pot block already holds a level-0 pixel
  const bool need2 = cb == 0;    // level 2 only if no level-0 (and, checked below, no level-1) selection in the block
  const int x4 = 4 * bx * P.pot, y4 = 4 * by * P.pot;
  const int bw = min(4 * P.pot, P.w - x4), bh = min(4 * P.pot, P.h - y4);
  unsigned long long best1[4] = {0ull, 0ull, 0ull, 0ull}, best2 = 0ull;
  const float d40 = kDirections[subd & 15][0], d41 = kDirections[subd & 15][1];
  for (int q = lane; q < bw * bh; q += 32) {
    const int xl = q % bw, yl = q / bw;
    const int xf = x4 + xl, yf = y4 + yl;
    if (!sel_inb(xf, yf, P.w, P.h)) continue;
    const int cxl = xl / P.pot, cyl = yl / P.pot;
    const int sb = (cxl >> 1) + 2 * (cyl >> 1);
    const bool want1 = !((subsel >> sb) & 1);
    if (!want1 && !need2) continue;
    const unsigned key = (unsigned)(((sb * 4 + (cyl & 1) * 2 + (cxl & 1)) << 20) | ((yl - cyl * P.pot) << 10) | (xl - cxl * P.pot));
    const float th0 = sel_th0(P, xf, yf);
    const float pixelTH1 = th0 * P.dw1;
    const float4 t = P.tex0[xf + P.w * yf];
    if (want1) {
      const float ag1 = P.tex1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * P.w1].w;
      if (ag1 > pixelTH1 * P.thFactor) {
        const int d3 = (subd >> (4 * sb)) & 15;
        const float dn = fabsf(t.y * kDirections[d3][0] + t.z * kDirections[d3][1]);
        if (dn > 0.f) {
          const unsigned long long cand = ((unsigned long long)__float_as_uint(dn) << 32) | (unsigned)(~key);
#pragma unroll
          for (int k = 0; k < 4; k++) if (sb == k && cand > best1[k]) best1[k] = cand;
        }
      }
    }
    if (need2) {
      const float pixelTH2 = pixelTH1 * P.dw2;
      const float ag2 = P.tex2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * P.w2].w;
      if (ag2 > pixelTH2 * P.thFactor) {
        const float dn = fabsf(t.y * d40 + t.z * d41);
        if (dn > 0.f) {
          const unsigned long long cand = ((unsigned long long)__float_as_uint(dn) << 32) | (unsigned)(~key);
          if (cand > best2) best2 = cand;
        }
      }
    }
  }
  auto decode = [&](unsigned long long v) {
    const unsigned key = ~(unsigned)(v & 0xffffffffull);
    const int cell = key >> 20, y1 = (key >> 10) & 1023, x1 = key & 1023;
    const int sb = cell >> 2, cxl = 2 * (sb & 1) + (cell & 1), cyl = 2 * (sb >> 1) + ((cell >> 1) & 1);
    return (x4 + cxl * P.pot + x1) + P.w * (y4 + cyl * P.pot + y1);
  };
  bool fired1 = false;
  int n3 = 0;
#pragma unroll
  for (int sb = 0; sb < 4; sb++) {
    if (!((subex >> sb) & 1) || ((subsel >> sb) & 1)) continue;
    const unsigned long long v = warp_max_u64(best1[sb]);
    if (v != 0ull) { fired1 = true; n3++; if (lane == 0) map[decode(v)] = 2.f; }
  }
  if (lane == 0 && n3) atomicAdd(&counts[1], n3);
  if (!need2 || fired1) return;
  const unsigned long long v = warp_max_u64(best2);
  if (v != 0ull && lane == 0) { map[decode(v)] = 4.f; atomicAdd(&counts[2], 1); }
}

// Raster-order rank of the selected pixels, one warp per image row, in three small launches: per-row counts; the random
// sub-sampling of makeMaps (:300-316; randomPattern[rank] > 255 * quotia drops the pixel) with per-row survivor counts; the
// compact (x, y, type) list of the survivors.
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int rows_before(const int* __restrict__ cnt, int row, int lane) {
  int s = 0;
  for (int r = lane; r < row; r += 32) s += cnt[r];
  return warp_sum_i(s);
}

__global__ void __launch_bounds__(256) sel_rowcount_kernel(const float* __restrict__ map, int w, int h, int* rowhave) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= h) return;
  int c = 0;
  for (int x = lane; x < w; x += 32) c += map[row * w + x] != 0.f;
  c = warp_sum_i(c);
  if (lane == 0) rowhave[row] = c;
}

__global__ void __launch_bounds__(256) sel_subsample_kernel(float* map, int w, int h, const unsigned char* __restrict__ rp, unsigned charTH,
                                                            const int* __restrict__ rowhave, int* rowkeep) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= h) return;
  int rn = rows_before(rowhave, row, lane), keep = 0;
  for (int x0 = 0; x0 < w; x0 += 32) {
    const int x = x0 + lane;
    const bool nz = x < w && map[row * w + x] != 0.f;
    const unsigned b = __ballot_sync(0xffffffffu, nz);
    bool kept = nz;
    if (nz && rp[rn + __popc(b & ((1u << lane) - 1))] > charTH) { map[row * w + x] = 0.f; kept = false; }
    rn += __popc(b);
    keep += __popc(__ballot_sync(0xffffffffu, kept));
  }
  if (lane == 0) rowkeep[row] = keep;
}

__global__ void __launch_bounds__(256) sel_list_kernel(const float* __restrict__ map, int w, int h, const int* __restrict__ rowhave,
                                                       const int* __restrict__ rowkeep, float* list_uv, float* list_type, int* counts) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= h) return;
  int kn = rows_before(rowkeep, row, lane);
  for (int x0 = 0; x0 < w; x0 += 32) {
    const int x = x0 + lane;
    const float v = x < w ? map[row * w + x] : 0.f;
    const unsigned b = __ballot_sync(0xffffffffu, v != 0.f);
    if (v != 0.f) {
      const int r = kn + __popc(b & ((1u << lane) - 1));
      list_uv[2 * r] = (float)x; list_uv[2 * r + 1] = (float)row; list_type[r] = v;
    }
    kn += __popc(b);
  }
  if (row == h - 1) {
    const int have = rows_before(rowhave, h, lane);
    if (lane == 0) { counts[4] = have; counts[5] = kn; }
  }
}

int selector_create(sdso_ctx* ctx) {
  SelectorState* s = new SelectorState();
  ctx->selector = s;
  const int w = ctx->G.w[0], h = ctx->G.h[0];
  s->w = w; s->h = h; s->w32 = w / 32; s->h32 = h / 32; s->ths_alloc = s->w32 * s->h32 + 100;
  const size_t wh = (size_t)w * h;
  s->h_rp.resize(wh);
  glibc_rand_bytes(3141592u, s->h_rp.data(), wh);
  SDSO_CUDA(ctx, cudaMalloc(&s->d_rp, wh));
  SDSO_CUDA(ctx, cudaMemcpy(s->d_rp, s->h_rp.data(), wh, cudaMemcpyHostToDevice));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_ths, s->ths_alloc * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_thsSmoothed, s->ths_alloc * sizeof(float)));
  SDSO_CUDA(ctx, cudaMemset(s->d_ths, 0, s->ths_alloc * sizeof(float)));
  SDSO_CUDA(ctx, cudaMemset(s->d_thsSmoothed, 0, s->ths_alloc * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_map, wh * sizeof(float)));
  SDSO_CUDA(ctx, cudaMemset(s->d_map, 0, wh * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_mask, wh * sizeof(unsigned short)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_amb_mask, wh * sizeof(unsigned short)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_cpre, wh * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_apre, wh * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_amb_cpre, wh * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_ambsel, (wh + 1) * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_n2cell, wh * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_selc, wh));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_rowhave, h * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_rowkeep, h * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_counts, 8 * sizeof(int)));
  SDSO_CUDA(ctx, cudaMemset(s->d_counts, 0, 8 * sizeof(int)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_list_uv, wh * 2 * sizeof(float)));
  SDSO_CUDA(ctx, cudaMalloc(&s->d_list_type, wh * sizeof(float)));
  SDSO_CUDA(ctx, cudaMallocHost(&s->h_counts, 8 * sizeof(int)));
  return SDSO_OK;
}

void selector_destroy(sdso_ctx* ctx) {
  SelectorState* s = ctx->selector;
  if (!s) return;
  void* ptrs[] = {s->d_rp, s->d_ths, s->d_thsSmoothed, s->d_map, s->d_mask, s->d_amb_mask, s->d_cpre, s->d_apre, s->d_amb_cpre, s->d_ambsel,
                  s->d_n2cell, s->d_selc, s->d_counts, s->d_list_uv, s->d_list_type, s->d_rowhave, s->d_rowkeep};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (s->h_counts) cudaFreeHost(s->h_counts);
  delete s;
  ctx->selector = nullptr;
}

static int check_frame(sdso_ctx* ctx, int frame) {
  if (!ctx || !ctx->selector) return SDSO_E_INVALID;
  if (frame < 0 || frame >= (int)ctx->frames.size() || !ctx->frames[frame].valid) return fail(ctx, SDSO_E_INVALID, "selector: invalid frame");
  if (ctx->G.levels < 3) return fail(ctx, SDSO_E_INVALID, "selector: needs 3 pyramid levels");
  return SDSO_OK;
}

static int launch_hists(sdso_ctx* ctx, int frame) {
  SelectorState* s = ctx->selector;
  const Frame& f = ctx->frames[frame];
  s->histFrame = frame; s->histGen = f.gen;
  if (s->w32 > 0 && s->h32 > 0) {
    sel_hist_kernel<<<dim3(s->w32, s->h32), 256, 0, ctx->stream>>>(f.tex[0], s->w, s->h, s->w32, ctx->S.minGradHistCut, ctx->S.minGradHistAdd, s->d_ths);
    SDSO_CHECK_LAUNCH(ctx);
    sel_smooth_kernel<<<(s->w32 * s->h32 + 127) / 128, 128, 0, ctx->stream>>>(s->d_ths, s->d_thsSmoothed, s->w32, s->h32);
    SDSO_CHECK_LAUNCH(ctx);
  }
  return SDSO_OK;
}

// select() on the stream; the counts land in h_counts after the caller synchronises
static int launch_select(sdso_ctx* ctx, int frame, int pot, float thFactor) {
  SelectorState* s = ctx->selector;
  const Frame& f = ctx->frames[frame];
  if (pot < 1) return fail(ctx, SDSO_E_INVALID, "selector: pot < 1");
  SelParams P{};
  P.tex0 = f.tex[0]; P.tex1 = f.tex[1]; P.tex2 = f.tex[2];
  P.thsSmoothed = s->d_thsSmoothed; P.rp = s->d_rp;
  P.w = s->w; P.h = s->h; P.w1 = ctx->G.w[1]; P.w2 = ctx->G.w[2]; P.thsStep = s->w32; P.ths_alloc = s->ths_alloc;
  P.pot = pot; P.ncx = (s->w + pot - 1) / pot; P.ncy = (s->h + pot - 1) / pot;
  P.thFactor = thFactor; P.dw1 = ctx->S.gradDownweightPerLevel; P.dw2 = P.dw1 * P.dw1;
  const int ncells = P.ncx * P.ncy;
  SDSO_CUDA(ctx, cudaMemsetAsync(s->d_map, 0, (size_t)s->w * s->h * sizeof(float), ctx->stream));
  sel_cell_mask_kernel<<<dim3((P.ncx + 127) / 128, P.ncy), 128, 0, ctx->stream>>>(P, s->d_mask);
  SDSO_CHECK_LAUNCH(ctx);
  sel_scan_kernel<<<1, 1024, 0, ctx->stream>>>(s->d_mask, ncells, s->d_rp, s->w * s->h, s->d_cpre, s->d_apre, s->d_amb_mask, s->d_amb_cpre, s->d_ambsel, s->d_counts);
  SDSO_CHECK_LAUNCH(ctx);
  sel_level0_kernel<<<dim3((P.ncx + 127) / 128, P.ncy), 128, 0, ctx->stream>>>(P, s->d_mask, s->d_cpre, s->d_apre, s->d_ambsel, s->d_map, s->d_selc, s->d_n2cell);
  SDSO_CHECK_LAUNCH(ctx);
  const int nbx = (P.ncx + 3) / 4, nby = (P.ncy + 3) / 4;
  if (pot > 1023) return fail(ctx, SDSO_E_INVALID, "selector: pot > 1023");
  sel_level12_kernel<<<(nbx * nby + 7) / 8, 256, 0, ctx->stream>>>(P, s->d_selc, s->d_n2cell, s->d_map, s->d_counts);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->h_counts, s->d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  s->list_valid = false;
  return SDSO_OK;
}

static int launch_compact(sdso_ctx* ctx, int subsample, unsigned charTH) {
  SelectorState* s = ctx->selector;
  const int rows_per_cta = 8, grid = (s->h + rows_per_cta - 1) / rows_per_cta;
  sel_rowcount_kernel<<<grid, 32 * rows_per_cta, 0, ctx->stream>>>(s->d_map, s->w, s->h, s->d_rowhave);
  SDSO_CHECK_LAUNCH(ctx);
  if (subsample) {
    sel_subsample_kernel<<<grid, 32 * rows_per_cta, 0, ctx->stream>>>(s->d_map, s->w, s->h, s->d_rp, charTH, s->d_rowhave, s->d_rowkeep);
    SDSO_CHECK_LAUNCH(ctx);
  }
  sel_list_kernel<<<grid, 32 * rows_per_cta, 0, ctx->stream>>>(s->d_map, s->w, s->h, s->d_rowhave, subsample ? s->d_rowkeep : s->d_rowhave, s->d_list_uv,
                                                              s->d_list_type, s->d_counts);
  SDSO_CHECK_LAUNCH(ctx);
  SDSO_CUDA(ctx, cudaMemcpyAsync(s->h_counts + 4, s->d_counts + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  s->list_valid = true;
  return SDSO_OK;
}

}  // namespace sdso

using namespace sdso;

extern "C" {

void sdso_selector_pattern_host(unsigned seed, unsigned char* out, size_t n) { if (out) glibc_rand_bytes(seed, out, n); }

int sdso_selector_reset(sdso_ctx* ctx) {
  sdso::enter(ctx);
  if (!ctx || !ctx->selector) return SDSO_E_INVALID;
  ctx->selector->currentPotential = 3;
  ctx->selector->histFrame = -1;
  ctx->selector->list_valid = false;
  return SDSO_OK;
}

int sdso_selector_random_pattern(sdso_ctx* ctx, unsigned char* out) {
  sdso::enter(ctx);
  if (!ctx || !ctx->selector || !out) return SDSO_E_INVALID;
  memcpy(out, ctx->selector->h_rp.data(), ctx->selector->h_rp.size());
  return SDSO_OK;
}

int sdso_selector_potential(sdso_ctx* ctx, int set, int* potential) {
  sdso::enter(ctx);
  if (!ctx || !ctx->selector) return SDSO_E_INVALID;
  if (set > 0) ctx->selector->currentPotential = set;
  if (potential) *potential = ctx->selector->currentPotential;
  return SDSO_OK;
}

int sdso_selector_make_hists(sdso_ctx* ctx, int frame, float* ths, float* ths_smoothed) {
  sdso::enter(ctx);
  int rc = check_frame(ctx, frame);
  if (rc) return rc;
  rc = launch_hists(ctx, frame);
  if (rc) return rc;
  SelectorState* s = ctx->selector;
  const size_t nb = (size_t)s->w32 * s->h32 * sizeof(float);
  if (ths && nb) SDSO_CUDA(ctx, cudaMemcpyAsync(ths, s->d_ths, nb, cudaMemcpyDeviceToHost, ctx->stream));
  if (ths_smoothed && nb) SDSO_CUDA(ctx, cudaMemcpyAsync(ths_smoothed, s->d_thsSmoothed, nb, cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

int sdso_selector_select(sdso_ctx* ctx, int frame, int pot, float thFactor, float* map_out, int n[3]) {
  sdso::enter(ctx);
  int rc = check_frame(ctx, frame);
  if (rc) return rc;
  rc = launch_select(ctx, frame, pot, thFactor);
  if (rc) return rc;
  SelectorState* s = ctx->selector;
  if (map_out) SDSO_CUDA(ctx, cudaMemcpyAsync(map_out, s->d_map, (size_t)s->w * s->h * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 3; i++) s->last_n[i] = s->h_counts[i];
  if (n) for (int i = 0; i < 3; i++) n[i] = s->h_counts[i];
  return SDSO_OK;
}

int sdso_make_maps(sdso_ctx* ctx, int frame, float density, int recursionsLeft, float thFactor, float* map_out, int* num_selected) {
  sdso::enter(ctx);
  int rc = check_frame(ctx, frame);
  if (rc) return rc;
  SelectorState* s = ctx->selector;
  const Frame& f = ctx->frames[frame];
  for (;;) {  // the reference recurses; every level repeats the same body (PixelSelector2.cpp:192-327)
    float numHave = 0, numWant = density, quotia;
    int idealPotential = s->currentPotential;
    if (s->histFrame != frame || s->histGen != f.gen) { rc = launch_hists(ctx, frame); if (rc) return rc; }
    rc = launch_select(ctx, frame, s->currentPotential, thFactor);
    if (rc) return rc;
    SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; i++) s->last_n[i] = s->h_counts[i];
    numHave = s->h_counts[0] + s->h_counts[1] + s->h_counts[2];
    quotia = numWant / numHave;
    const float K = numHave * (s->currentPotential + 1) * (s->currentPotential + 1);
    idealPotential = sqrtf(K / numWant) - 1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft > 0 && quotia > 1.25 && s->currentPotential > 1) {
      if (idealPotential >= s->currentPotential) idealPotential = s->currentPotential - 1;
      s->currentPotential = idealPotential;
      recursionsLeft--;
      continue;
    } else if (recursionsLeft > 0 && quotia < 0.25) {
      if (idealPotential <= s->currentPotential) idealPotential = s->currentPotential + 1;
      s->currentPotential = idealPotential;
      recursionsLeft--;
      continue;
    }
    const int subsample = quotia < 0.95 ? 1 : 0;
    const unsigned char charTH = subsample ? (unsigned char)(255 * quotia) : 0;
    rc = launch_compact(ctx, subsample, charTH);
    if (rc) return rc;
    if (map_out) SDSO_CUDA(ctx, cudaMemcpyAsync(map_out, s->d_map, (size_t)s->w * s->h * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s->currentPotential = idealPotential;
    if (num_selected) *num_selected = s->h_counts[5];
    return SDSO_OK;
  }
}

int sdso_selector_points(sdso_ctx* ctx, int max_n, float* uv, float* type, int* n) {
  sdso::enter(ctx);
  if (!ctx || !ctx->selector || !n) return SDSO_E_INVALID;
  SelectorState* s = ctx->selector;
  if (!s->list_valid) {
    int rc = launch_compact(ctx, 0, 0);
    if (rc) return rc;
  }
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int cnt = s->h_counts[5];
  *n = cnt;
  if (cnt > max_n) return fail(ctx, SDSO_E_INVALID, "selector_points: max_n too small");
  if (cnt > 0 && uv) SDSO_CUDA(ctx, cudaMemcpyAsync(uv, s->d_list_uv, (size_t)cnt * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (cnt > 0 && type) SDSO_CUDA(ctx, cudaMemcpyAsync(type, s->d_list_type, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SDSO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDSO_OK;
}

}  // extern "C"
