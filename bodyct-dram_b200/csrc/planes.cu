// Elementwise kernels whose activation operands live as bf16 split planes (hi = bf16(x), lo = bf16(x - hi)), the operand
// format of the tcgen05 convolutions.  Between two tensor-core convolutions an activation is written ONCE, as planes, by
// the BatchNorm+ReLU(+pool) kernel / the upsample+concat kernel / the BatchNorm backward kernel and read by the next
// convolution's TMA loads: the fp32 copy and the separate split pass do not exist (4 B/element either way).
// hi + lo is exact in fp32 and carries 16-17 significant bits of the fp32 value it was split from.
// All kernels are HBM-bound streams: 8 channels (2 x float4 in, one 16-byte store per plane) per thread.
#include <stdlib.h>
#include "common.cuh"

namespace dram {

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z);
  split2(v[6], v[7], hi.w, lo.w);
}
// bf16 -> fp32 is a 16-bit shift
__device__ __forceinline__ void merge2(uint32_t hi, uint32_t lo, float& a, float& b) {
  a = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
  b = __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u);
}
__device__ __forceinline__ void merge8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  merge2(hi.x, lo.x, v[0], v[1]);
  merge2(hi.y, lo.y, v[2], v[3]);
  merge2(hi.z, lo.z, v[4], v[5]);
  merge2(hi.w, lo.w, v[6], v[7]);
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ------------------------------------------------------------------------------------------------ BN + ReLU (+ pool) -> planes
// y [rows][C] fp32 -> a planes [rows][Cpad]; channels [C, Cpad) are written as zeros.
// The launch guarantees (gridDim.x * blockDim.x) % groups == 0: a thread keeps its channel group.
__global__ void __launch_bounds__(256)
k_bn_relu_apply_planes(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                       uint4* __restrict__ hi, uint4* __restrict__ lo, long long rows, int C, int Cpad) {
  const int groups = Cpad / 8;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dr = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);
  const bool real = g * 8 < C;
  float sc[8], sh[8];
  if (real) { load8(scale + g * 8, sc); load8(shift + g * 8, sh); }
  for (long long r = tid / groups; r < rows; r += dr) {
    uint4 H = make_uint4(0, 0, 0, 0), L = make_uint4(0, 0, 0, 0);
    if (real) {
      float v[8];
      load8(y + r * C + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
      split8(v, H, L);
    }
    hi[r * groups + g] = H;
    if (lo) lo[r * groups + g] = L;
  }
}

// one thread per (2x2x2 pool cell, 8-channel group): writes the 8 activations of the cell and their max, all as planes
__global__ void __launch_bounds__(256)
k_bn_relu_pool_planes(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                      uint4* __restrict__ hi, uint4* __restrict__ lo, uint4* __restrict__ phi, uint4* __restrict__ plo,
                      int N, int D, int H, int W, int C, int Cpad) {
  const int groups = Cpad / 8;
  const int cd = (D + 1) / 2, ch = (H + 1) / 2, cw = (W + 1) / 2;   // cells incl. ragged tail
  const int pd = D / 2, ph = H / 2, pw = W / 2;                       // pooled size (floor)
  const long long cells = (long long)N * cd * ch * cw;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dcell = ((long long)gridDim.x * blockDim.x) / groups;   // launch: total threads % groups == 0
  const int g = (int)(tid % groups);
  const bool real = g * 8 < C;
  float sc[8], sh[8];
  if (real) { load8(scale + g * 8, sc); load8(shift + g * 8, sh); }
  for (long long cell = tid / groups; cell < cells; cell += dcell) {
    unsigned r = (unsigned)cell;                                        // cells < 2^31 (checked by the host)
    const int x = (int)(r % (unsigned)cw); r /= (unsigned)cw;
    const int yy = (int)(r % (unsigned)ch); r /= (unsigned)ch;
    const int z = (int)(r % (unsigned)cd);
    const int n = (int)(r / (unsigned)cd);
    float mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zz = 2 * z + (k >> 2), yv = 2 * yy + ((k >> 1) & 1), xv = 2 * x + (k & 1);
      if (zz < D && yv < H && xv < W) {
        const long long row = (((long long)n * D + zz) * H + yv) * W + xv;
        uint4 Hh = make_uint4(0, 0, 0, 0), Ll = make_uint4(0, 0, 0, 0);
        if (real) {
          float v[8];
          load8(y + row * C + g * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f); mx[j] = fmaxf(mx[j], v[j]); }
          split8(v, Hh, Ll);
        }
        hi[row * groups + g] = Hh;
        if (lo) lo[row * groups + g] = Ll;
      }
    }
    if (z < pd && yy < ph && x < pw) {
      const long long prow = (((long long)n * pd + z) * ph + yy) * pw + x;
      uint4 Hh = make_uint4(0, 0, 0, 0), Ll = make_uint4(0, 0, 0, 0);
      if (real) split8(mx, Hh, Ll);
      phi[prow * groups + g] = Hh;
      if (plo) plo[prow * groups + g] = Ll;
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN + ReLU backward -> dy planes
// RANK1: da[r][c] = g[r] * wtop[c] (fused RAM head), otherwise da rows are da_pitch floats apart.
// dy = k1*dz - k2 - (y - mean)*k3 with per-channel constants folded once per thread (the launch guarantees
// (gridDim.x * blockDim.x) % groups == 0, so a thread keeps its channel group).
template <bool RANK1>
__global__ void __launch_bounds__(256)
k_bn_relu_bwd_apply_planes(const float* __restrict__ da, long long da_pitch, const float* __restrict__ wtop,
                           const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const double* __restrict__ sums, double count, uint4* __restrict__ hi, uint4* __restrict__ lo,
                           long long rows, int C, int Cpad) {
  const int groups = Cpad / 8;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dr = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);
  const bool real = g * 8 < C;
  float sc[8], sh[8], mu[8], k1[8], k2[8], k3[8], wt[8];
  if (real) {
    const float inv = sums ? (float)(1.0 / count) : 0.f;
    load8(scale + g * 8, sc);
    load8(shift + g * 8, sh);
    if (RANK1) load8(wtop + g * 8, wt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      if (sums) {
        const float rs = rstd[c], gm = gamma ? gamma[c] : 1.f;
        mu[j] = mean[c];
        k1[j] = gm * rs;
        k2[j] = gm * rs * ((float)sums[c] * inv);
        k3[j] = gm * rs * rs * ((float)sums[C + c] * inv);
      } else {
        mu[j] = 0.f; k1[j] = sc[j]; k2[j] = 0.f; k3[j] = 0.f;
      }
    }
  }
  for (long long r = tid / groups; r < rows; r += dr) {
    uint4 H = make_uint4(0, 0, 0, 0), L = make_uint4(0, 0, 0, 0);
    if (real) {
      float yv[8], dv[8], o[8];
      load8(y + r * C + g * 8, yv);
      if (RANK1) {
        const float gr = __ldg(da + r);
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = wt[j] * gr;
      } else {
        load8(da + r * da_pitch + g * 8, dv);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dz = (fmaf(yv[j], sc[j], sh[j]) > 0.f) ? dv[j] : 0.f;
        o[j] = k1[j] * dz - k2[j] - (yv[j] - mu[j]) * k3[j];
      }
      split8(o, H, L);
    }
    hi[r * groups + g] = H;
    if (lo) lo[r * groups + g] = L;
  }
}

// ------------------------------------------------------------------------------------------------ pooled unit backward
// [conv -> BN -> ReLU -> (a, maxpool(a))]: the gradient of a is  ga (skip path, rows ga_pitch apart, nullable) plus gp
// routed to the FIRST maximum of each 2x2x2 window (ATen: val > max || isnan), recomputed here from y — neither `a` nor
// a dense `da` exist.  4 channels per thread, one thread per window (ragged tail cells have no pooled output).
struct PoolCell {
  int n, z, y, x;
};
__device__ __forceinline__ PoolCell decode_cell(long long cell64, int cd, int ch, int cw) {
  unsigned cell = (unsigned)cell64;                                     // cells < 2^31 (checked by the host)
  PoolCell c;
  c.x = (int)(cell % (unsigned)cw); cell /= (unsigned)cw;
  c.y = (int)(cell % (unsigned)ch); cell /= (unsigned)ch;
  c.z = (int)(cell % (unsigned)cd);
  c.n = (int)(cell / (unsigned)cd);
  return c;
}

// dz[k][v] of one window for a 4-channel group; returns the validity mask of the 8 voxels
__device__ __forceinline__ unsigned pool_window_dz(const float* __restrict__ y, const float* __restrict__ ga, long long ga_pitch,
                                                   const float* __restrict__ gp, const PoolCell& pc, int g, int D, int H,
                                                   int W, int C, const float (&sc)[4], const float (&sh)[4],
                                                   float (&yv)[8][4], float (&dz)[8][4]) {
  const int pd = D / 2, ph = H / 2, pw = W / 2;
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int arg[4] = {0, 0, 0, 0};
  unsigned valid = 0;
  // All 17 loads of the window are issued before the first one is consumed: coordinates are clamped into the volume and
  // the loads are unconditional (a ragged-tail voxel is masked afterwards), so no data-dependent branch or argmax chain
  // stands between them and a thread pays one memory latency per window instead of eight.
  float4 yt[8], qt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int zz = 2 * pc.z + (k >> 2), yy = 2 * pc.y + ((k >> 1) & 1), xx = 2 * pc.x + (k & 1);
    if (zz < D && yy < H && xx < W) valid |= 1u << k;
    const long long row = (((long long)pc.n * D + min(zz, D - 1)) * H + min(yy, H - 1)) * W + min(xx, W - 1);
    yt[k] = __ldg(reinterpret_cast<const float4*>(y + row * C) + g);
    qt[k] = ga ? __ldg(reinterpret_cast<const float4*>(ga + row * ga_pitch) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (valid & (1u << k)) {
      yv[k][0] = yt[k].x; yv[k][1] = yt[k].y; yv[k][2] = yt[k].z; yv[k][3] = yt[k].w;
      dz[k][0] = qt[k].x; dz[k][1] = qt[k].y; dz[k][2] = qt[k].z; dz[k][3] = qt[k].w;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float a = fmaxf(fmaf(yv[k][v], sc[v], sh[v]), 0.f);
        if (a > mx[v] || a != a) { mx[v] = a; arg[v] = k; }
      }
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v) { yv[k][v] = 0.f; dz[k][v] = 0.f; }
    }
  }
  if (gp && pc.z < pd && pc.y < ph && pc.x < pw) {
    const long long prow = (((long long)pc.n * pd + pc.z) * ph + pc.y) * pw + pc.x;
    const float4 t = __ldg(reinterpret_cast<const float4*>(gp + prow * C) + g);
    const float gpv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (arg[v] == k) dz[k][v] += gpv[v];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (!(fmaf(yv[k][v], sc[v], sh[v]) > 0.f)) dz[k][v] = 0.f;
  return valid;
}

constexpr int kPoolThreads = 256;
__global__ void __launch_bounds__(kPoolThreads, 2)
k_bn_pool_bwd_reduce(const float* __restrict__ ga, long long ga_pitch, const float* __restrict__ gp,
                     const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ rstd, double* __restrict__ sums, int N, int D,
                     int H, int W, int C) {
  __shared__ float s0[kPoolThreads * 4], s1[kPoolThreads * 4];
  const int groups = C / 4;
  const int gl = groups < kPoolThreads ? groups : kPoolThreads;
  const int lanes = kPoolThreads / gl;
  const int cd = (D + 1) / 2, ch = (H + 1) / 2, cw = (W + 1) / 2;
  const long long cells = (long long)N * cd * ch * cw;
  const long long per_block = (cells + gridDim.x - 1) / gridDim.x;
  const long long c_begin = (long long)blockIdx.x * per_block;
  const long long c_end = c_begin + per_block < cells ? c_begin + per_block : cells;
  for (int g0 = 0; g0 < groups; g0 += kPoolThreads) {
    const int g = g0 + (threadIdx.x % gl);
    const int lane = threadIdx.x / gl;
    const bool active = g < groups && lane < lanes;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      float sc[4], sh[4], mu[4], rs[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) { sc[v] = scale[g * 4 + v]; sh[v] = shift[g * 4 + v]; mu[v] = mean[g * 4 + v]; rs[v] = rstd[g * 4 + v]; }
      for (long long cell = c_begin + lane; cell < c_end; cell += lanes) {
        float yv[8][4], dz[8][4];
        const PoolCell pc = decode_cell(cell, cd, ch, cw);
        pool_window_dz(y, ga, ga_pitch, gp, pc, g, D, H, W, C, sc, sh, yv, dz);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            a0[v] += dz[k][v];
            a1[v] += dz[k][v] * (yv[k][v] - mu[v]) * rs[v];
          }
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) { s0[threadIdx.x * 4 + v] = a0[v]; s1[threadIdx.x * 4 + v] = a1[v]; }
    __syncthreads();
    if (active && lane == 0) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double t0 = 0.0, t1 = 0.0;
        for (int l = 0; l < lanes; ++l) {
          t0 += (double)s0[(l * gl + (threadIdx.x % gl)) * 4 + v];
          t1 += (double)s1[(l * gl + (threadIdx.x % gl)) * 4 + v];
        }
        atomicAdd(&sums[g * 4 + v], t0);
        atomicAdd(&sums[C + g * 4 + v], t1);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_bn_pool_bwd_apply_planes(const float* __restrict__ ga, long long ga_pitch, const float* __restrict__ gp,
                           const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const double* __restrict__ sums, double count, uint2* __restrict__ hi, uint2* __restrict__ lo,
                           int N, int D, int H, int W, int C, int Cpad) {
  const int groups = Cpad / 4;
  const int cd = (D + 1) / 2, ch = (H + 1) / 2, cw = (W + 1) / 2;
  const long long cells = (long long)N * cd * ch * cw;
  const float inv = sums ? (float)(1.0 / count) : 0.f;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dcell = ((long long)gridDim.x * blockDim.x) / groups;   // launch: total threads % groups == 0
  const int g = (int)(tid % groups);
  const bool real = g * 4 < C;
  float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  float k1[4], k2[4], k3[4], mu[4];          // dy = k1*dz - k2 - (y - mu)*k3
  if (real) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = g * 4 + v;
      sc[v] = scale[c]; sh[v] = shift[c];
      if (sums) {
        const float rs = rstd[c], gm = gamma ? gamma[c] : 1.f;
        mu[v] = mean[c];
        k1[v] = gm * rs;
        k2[v] = gm * rs * ((float)sums[c] * inv);
        k3[v] = gm * rs * rs * ((float)sums[C + c] * inv);
      } else {
        mu[v] = 0.f; k1[v] = sc[v]; k2[v] = 0.f; k3[v] = 0.f;
      }
    }
  }
  for (long long cell = tid / groups; cell < cells; cell += dcell) {
    const PoolCell pc = decode_cell(cell, cd, ch, cw);
    float yv[8][4], dz[8][4];
    if (real) pool_window_dz(y, ga, ga_pitch, gp, pc, g, D, H, W, C, sc, sh, yv, dz);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zz = 2 * pc.z + (k >> 2), yy = 2 * pc.y + ((k >> 1) & 1), xx = 2 * pc.x + (k & 1);
      if (zz < D && yy < H && xx < W) {
        const long long row = (((long long)pc.n * D + zz) * H + yy) * W + xx;
        uint2 Hh = make_uint2(0, 0), Ll = make_uint2(0, 0);
        if (real) {
          float o[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) o[v] = k1[v] * dz[k][v] - k2[v] - (yv[k][v] - mu[v]) * k3[v];
          split2(o[0], o[1], Hh.x, Ll.x);
          split2(o[2], o[3], Hh.y, Ll.y);
        }
        hi[row * groups + g] = Hh;
        if (lo) lo[row * groups + g] = Ll;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ upsample x2 + concat on planes
// cat[..., 0:C1] = trilinear x2 (align_corners=True) of x, cat[..., C1:C1+C2] = skip (centre crop, ceil offsets); all
// three tensors are planes (pitches P1, P2, Pc in channels).  One thread per (output voxel, 8-channel group).
__global__ void __launch_bounds__(256)
k_upsample2x_concat_planes(const uint4* __restrict__ xh, const uint4* __restrict__ xl, const uint4* __restrict__ sh_,
                           const uint4* __restrict__ sl, uint4* __restrict__ ch_, uint4* __restrict__ cl, int N, int d,
                           int h, int w, int C1, int P1, int Ds, int Hs, int Ws, int C2, int P2, int Pc, int oz, int oy,
                           int ox, float sz, float sy, float sx) {
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const int groups = Pc / 8, g1 = C1 / 8, g2 = (C1 + C2) / 8, pg1 = P1 / 8, pg2 = P2 / 8;
  const long long rows = (long long)N * D * H * W;
  for (RowGroupIter it(groups); it.r < rows; it.next()) {
    const int g = it.g;
    const long long i = it.r * groups + g;
    unsigned r = (unsigned)it.r;                                        // rows < 2^31 (checked by the host)
    const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
    const int Y = (int)(r % (unsigned)H); r /= (unsigned)H;
    const int Z = (int)(r % (unsigned)D);
    const int n = (int)(r / (unsigned)D);
    uint4 Hh = make_uint4(0, 0, 0, 0), Ll = make_uint4(0, 0, 0, 0);
    if (g < g1) {
      const Lerp lz = lerp_setup(Z, sz, d), ly = lerp_setup(Y, sy, h), lx = lerp_setup(X, sx, w);
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int zi = (k & 4) ? lz.i1 : lz.i0, yi = (k & 2) ? ly.i1 : ly.i0, xi = (k & 1) ? lx.i1 : lx.i0;
        const float wt = ((k & 4) ? lz.w1 : lz.w0) * ((k & 2) ? ly.w1 : ly.w0) * ((k & 1) ? lx.w1 : lx.w0);
        const long long row = (((long long)n * d + zi) * h + yi) * w + xi;
        const uint4 a = __ldg(xh + row * pg1 + g);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (xl) b = __ldg(xl + row * pg1 + g);
        float v[8];
        merge8(a, b, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wt * v[j];
      }
      split8(acc, Hh, Ll);
    } else if (g < g2) {
      const long long row = (((long long)n * Ds + Z + oz) * Hs + Y + oy) * Ws + X + ox;
      Hh = __ldg(sh_ + row * pg2 + (g - g1));
      if (sl) Ll = __ldg(sl + row * pg2 + (g - g1));
    }
    ch_[i] = Hh;
    if (cl) cl[i] = Ll;
  }
}

// Register-blocked variant: one thread produces the 2x2x2 block of output voxels (2k, 2k+1 per axis) of one 8-channel
// group.  With align_corners=True and an exact factor 2 those outputs read source indices k-1, k, k+1 only, so the block
// needs 27 source loads (3.4 per output instead of 8) and the axes are interpolated one after the other in registers.
// Weights come from the same lerp_setup as every other trilinear kernel of the library.
__global__ void __launch_bounds__(128)
k_up2x_planes_blocked(const uint4* __restrict__ xh, const uint4* __restrict__ xl, uint4* __restrict__ ch_,
                      uint4* __restrict__ cl, int N, int d, int h, int w, int C1, int P1, int Pc, float sz, float sy, float sx) {
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const int groups = C1 / 8, pg1 = P1 / 8, pgc = Pc / 8;
  const long long blocks = (long long)N * d * h * w;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long db = ((long long)gridDim.x * blockDim.x) / groups;     // launch: total threads % groups == 0
  const int g = (int)(tid % groups);
  for (long long b = tid / groups; b < blocks; b += db) {
    unsigned r = (unsigned)b;
    const int kx = (int)(r % (unsigned)w); r /= (unsigned)w;
    const int ky = (int)(r % (unsigned)h); r /= (unsigned)h;
    const int kz = (int)(r % (unsigned)d);
    const int n = (int)(r / (unsigned)d);
    float wz[2][3], wy[2][3], wx[2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const Lerp lz = lerp_setup(2 * kz + j, sz, d), ly = lerp_setup(2 * ky + j, sy, h), lx = lerp_setup(2 * kx + j, sx, w);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        wz[j][t] = (lz.i0 == kz - 1 + t ? lz.w0 : 0.f) + (lz.i1 == kz - 1 + t ? lz.w1 : 0.f);
        wy[j][t] = (ly.i0 == ky - 1 + t ? ly.w0 : 0.f) + (ly.i1 == ky - 1 + t ? ly.w1 : 0.f);
        wx[j][t] = (lx.i0 == kx - 1 + t ? lx.w0 : 0.f) + (lx.i1 == kx - 1 + t ? lx.w1 : 0.f);
      }
    }
    // Source indices are clamped into the volume and every load is unconditional (an out-of-range candidate has weight
    // 0), so the 9 x 2 loads of a z-plane are all in flight before the first one is consumed.
    int xi[3], yi[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      xi[t] = min(max(kx - 1 + t, 0), w - 1);
      yi[t] = min(max(ky - 1 + t, 0), h - 1);
    }
    float out[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) out[i][c] = 0.f;
#pragma unroll
    for (int tz = 0; tz < 3; ++tz) {
      const int zi = min(max(kz - 1 + tz, 0), d - 1);
      const long long zbase = ((long long)n * d + zi) * h;
      uint4 A[3][3], B[3][3];
#pragma unroll
      for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const long long row = (zbase + yi[ty]) * w + xi[tx];
          A[ty][tx] = __ldg(xh + row * pg1 + g);
          B[ty][tx] = xl ? __ldg(xl + row * pg1 + g) : make_uint4(0, 0, 0, 0);
        }
      float u[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) u[i][c] = 0.f;
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        float t0[8], t1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) t0[c] = t1[c] = 0.f;
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          float v[8];
          merge8(A[ty][tx], B[ty][tx], v);
#pragma unroll
          for (int c = 0; c < 8; ++c) { t0[c] = fmaf(wx[0][tx], v[c], t0[c]); t1[c] = fmaf(wx[1][tx], v[c], t1[c]); }
        }
#pragma unroll
        for (int jy = 0; jy < 2; ++jy)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            u[jy * 2 + 0][c] = fmaf(wy[jy][ty], t0[c], u[jy * 2 + 0][c]);
            u[jy * 2 + 1][c] = fmaf(wy[jy][ty], t1[c], u[jy * 2 + 1][c]);
          }
      }
#pragma unroll
      for (int jz = 0; jz < 2; ++jz)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) out[jz * 4 + i][c] = fmaf(wz[jz][tz], u[i][c], out[jz * 4 + i][c]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int Z = 2 * kz + (i >> 2), Y = 2 * ky + ((i >> 1) & 1), X = 2 * kx + (i & 1);
      const long long row = (((long long)n * D + Z) * H + Y) * W + X;
      uint4 Hh, Ll;
      split8(out[i], Hh, Ll);
      ch_[row * pgc + g] = Hh;
      if (cl) cl[row * pgc + g] = Ll;
    }
  }
}

// Sliding-window variant (the default): a thread owns one (n, ky, kx, 8-channel group) COLUMN of 2x2x2 output blocks and
// walks it along z.  The x/y-interpolated source planes kz-1, kz, kz+1 (4 outputs x 8 channels each) stay in registers, so
// every step loads and x/y-interpolates ONE new source plane (9 row loads instead of 27, 30 FMAs per channel instead of 90)
// and then blends the three planes along z.  Same weights (lerp_setup, 3-tap form with exact zeros) and the same per-axis
// order x -> y -> z as k_up2x_planes_blocked.  `zseg` blocks per thread: each segment start pays two extra planes.
__device__ __forceinline__ void up2x_plane(const uint4* __restrict__ xh, const uint4* __restrict__ xl, long long zbase, int w,
                                           const int (&yi)[3], const int (&xi)[3], int pg1, int g, const float (&wy)[2][3],
                                           const float (&wx)[2][3], float (&u)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) u[i][c] = 0.f;
#pragma unroll
  for (int ty = 0; ty < 3; ++ty) {
    uint4 A[3], B[3];
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      const long long row = (zbase + yi[ty]) * w + xi[tx];
      A[tx] = __ldg(xh + row * pg1 + g);
      B[tx] = xl ? __ldg(xl + row * pg1 + g) : make_uint4(0, 0, 0, 0);
    }
    float t0[8], t1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) t0[c] = t1[c] = 0.f;
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      float v[8];
      merge8(A[tx], B[tx], v);
#pragma unroll
      for (int c = 0; c < 8; ++c) { t0[c] = fmaf(wx[0][tx], v[c], t0[c]); t1[c] = fmaf(wx[1][tx], v[c], t1[c]); }
    }
#pragma unroll
    for (int jy = 0; jy < 2; ++jy)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        u[jy * 2 + 0][c] = fmaf(wy[jy][ty], t0[c], u[jy * 2 + 0][c]);
        u[jy * 2 + 1][c] = fmaf(wy[jy][ty], t1[c], u[jy * 2 + 1][c]);
      }
  }
}

__global__ void __launch_bounds__(128)
k_up2x_planes_zslide(const uint4* __restrict__ xh, const uint4* __restrict__ xl, uint4* __restrict__ ch_,
                     uint4* __restrict__ cl, int N, int d, int h, int w, int C1, int P1, int Pc, float sz, float sy, float sx,
                     int zseg, int nseg) {
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const int groups = C1 / 8, pg1 = P1 / 8, pgc = Pc / 8;
  const long long cols = (long long)N * nseg * h * w;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dcol = ((long long)gridDim.x * blockDim.x) / groups;     // launch: total threads % groups == 0
  const int g = (int)(tid % groups);
  for (long long b = tid / groups; b < cols; b += dcol) {
    unsigned r = (unsigned)b;
    const int kx = (int)(r % (unsigned)w); r /= (unsigned)w;
    const int ky = (int)(r % (unsigned)h); r /= (unsigned)h;
    const int seg = (int)(r % (unsigned)nseg);
    const int n = (int)(r / (unsigned)nseg);
    const int kz0 = seg * zseg, kz1 = min(kz0 + zseg, d);
    float wy[2][3], wx[2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const Lerp ly = lerp_setup(2 * ky + j, sy, h), lx = lerp_setup(2 * kx + j, sx, w);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        wy[j][t] = (ly.i0 == ky - 1 + t ? ly.w0 : 0.f) + (ly.i1 == ky - 1 + t ? ly.w1 : 0.f);
        wx[j][t] = (lx.i0 == kx - 1 + t ? lx.w0 : 0.f) + (lx.i1 == kx - 1 + t ? lx.w1 : 0.f);
      }
    }
    int xi[3], yi[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      xi[t] = min(max(kx - 1 + t, 0), w - 1);
      yi[t] = min(max(ky - 1 + t, 0), h - 1);
    }
    const long long nbase = (long long)n * d;
    float P0[4][8], P1_[4][8], P2[4][8];
    up2x_plane(xh, xl, (nbase + max(kz0 - 1, 0)) * h, w, yi, xi, pg1, g, wy, wx, P1_);     // becomes P0 in the first step
    up2x_plane(xh, xl, (nbase + kz0) * h, w, yi, xi, pg1, g, wy, wx, P2);                  // becomes P1
    for (int kz = kz0; kz < kz1; ++kz) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) { P0[i][c] = P1_[i][c]; P1_[i][c] = P2[i][c]; }
      up2x_plane(xh, xl, (nbase + min(kz + 1, d - 1)) * h, w, yi, xi, pg1, g, wy, wx, P2);
      float wz[2][3];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const Lerp lz = lerp_setup(2 * kz + j, sz, d);
#pragma unroll
        for (int t = 0; t < 3; ++t) wz[j][t] = (lz.i0 == kz - 1 + t ? lz.w0 : 0.f) + (lz.i1 == kz - 1 + t ? lz.w1 : 0.f);
      }
#pragma unroll
      for (int jz = 0; jz < 2; ++jz)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float o[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) o[c] = fmaf(wz[jz][2], P2[i][c], fmaf(wz[jz][1], P1_[i][c], wz[jz][0] * P0[i][c]));
          const int Z = 2 * kz + jz, Y = 2 * ky + (i >> 1), X = 2 * kx + (i & 1);
          const long long row = (((long long)n * D + Z) * H + Y) * W + X;
          uint4 Hh, Ll;
          split8(o, Hh, Ll);
          ch_[row * pgc + g] = Hh;
          if (cl) cl[row * pgc + g] = Ll;
        }
    }
  }
}

// cat[..., C1:C1+C2] = skip (centre crop), cat[..., C1+C2:Pc] = 0: 16-byte chunks, one thread per (voxel, chunk)
__global__ void __launch_bounds__(256)
k_skip_copy_planes(const uint4* __restrict__ sh_, const uint4* __restrict__ sl, uint4* __restrict__ ch_, uint4* __restrict__ cl,
                   int N, int D, int H, int W, int C1, int Ds, int Hs, int Ws, int C2, int P2, int Pc, int oz, int oy, int ox) {
  const int g1 = C1 / 8, g2 = (C1 + C2) / 8, pg2 = P2 / 8, pgc = Pc / 8;
  const int groups = pgc - g1;
  const long long rows = (long long)N * D * H * W;
  for (RowGroupIter it(groups); it.r < rows; it.next()) {
    const int g = g1 + it.g;
    uint4 Hh = make_uint4(0, 0, 0, 0), Ll = make_uint4(0, 0, 0, 0);
    if (g < g2) {
      long long srow = it.r;
      if (oz | oy | ox | (Ds - D) | (Hs - H) | (Ws - W)) {
        unsigned r = (unsigned)it.r;
        const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
        const int Y = (int)(r % (unsigned)H); r /= (unsigned)H;
        const int Z = (int)(r % (unsigned)D);
        const int n = (int)(r / (unsigned)D);
        srow = (((long long)n * Ds + Z + oz) * Hs + Y + oy) * Ws + X + ox;
      }
      Hh = __ldg(sh_ + srow * pg2 + (g - g1));
      if (sl) Ll = __ldg(sl + srow * pg2 + (g - g1));
    }
    ch_[it.r * pgc + g] = Hh;
    if (cl) cl[it.r * pgc + g] = Ll;
  }
}

// planes [rows][P] -> fp32 [rows][C]  (materialises an activation for a consumer that is not a tensor-core convolution)
__global__ void __launch_bounds__(256)
k_merge_planes(const uint4* __restrict__ hi, const uint4* __restrict__ lo, float* __restrict__ out, long long rows, int C,
               int P) {
  const int groups = C / 8, pg = P / 8;
  for (RowGroupIter it(groups); it.r < rows; it.next()) {
    const int g = it.g;
    const long long r = it.r;
    const uint4 a = __ldg(hi + r * pg + g);
    uint4 b = make_uint4(0, 0, 0, 0);
    if (lo) b = __ldg(lo + r * pg + g);
    float v[8];
    merge8(a, b, v);
    float4* q = reinterpret_cast<float4*>(out + r * C + g * 8);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ------------------------------------------------------------------------------------------------ 1x1x1 conv on planes
// The attention reshape heads (models.py:488-494, 564): Conv3d(C, 8, kernel_size=1) on a DETACHED decoder feature map that
// exists only as planes.  HBM-bound: 4 B/element in, 32 B/voxel out.  LPR = Cpad/8 lanes share a voxel row (a lane owns
// 8 channels: one 16-byte load per plane, the 8x8 weights of its channels in registers); a warp pass covers 32/LPR rows
// with fully coalesced loads.  The 8 partial outputs are combined by a reduce-scatter over the row's lanes (4+2+1
// shuffles: every lane ends with ONE output, lane sub holds output sub), so the store of a row is 32 contiguous bytes.
template <int LPR>
__global__ void __launch_bounds__(256)
k_pointwise8_planes_fwd(const uint4* __restrict__ hi, const uint4* __restrict__ lo, const float* __restrict__ w,
                        const float* __restrict__ bias, float* __restrict__ y, long long rows, int Cin) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
  float wr[8][8];                                               // wr[j][c] = w[j][8*sub + c]
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) wr[j][c] = (8 * sub + c < Cin) ? __ldg(w + j * Cin + 8 * sub + c) : 0.f;
  const float b = bias ? __ldg(bias + (sub & 7)) : 0.f;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  constexpr int U = 4;                                           // warp passes in flight: 2*U 16-byte loads per lane
  const long long rows_up = (rows + RPW - 1) / RPW * RPW;        // whole warps stay in the loop: the shuffles need them
  for (long long r0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW + rw; r0 < rows_up; r0 += warps * RPW * U) {
    uint4 H[U], L[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * warps * RPW;
      const bool live = r < rows;
      H[u] = live ? __ldg(hi + r * LPR + sub) : zero;
      L[u] = (live && lo) ? __ldg(lo + r * LPR + sub) : zero;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * warps * RPW;
      if (r - rw >= rows_up) break;                              // warp-uniform
      float v[8], p[8];
      merge8(H[u], L[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) a = fmaf(v[c], wr[j][c], a);
        p[j] = a;
      }
      if (LPR == 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] += __shfl_xor_sync(0xffffffffu, p[j], 8);
      }
      float p4[4], p2[2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float send = (sub & 4) ? p[j] : p[j + 4], keep = (sub & 4) ? p[j + 4] : p[j];
        p4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float send = (sub & 2) ? p4[j] : p4[j + 2], keep = (sub & 2) ? p4[j + 2] : p4[j];
        p2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const float send = (sub & 1) ? p2[0] : p2[1], keep = (sub & 1) ? p2[1] : p2[0];
      const float out = keep + __shfl_xor_sync(0xffffffffu, send, 1) + b;
      if (r < rows && sub < 8) y[r * 8 + sub] = out;
    }
  }
}

// weight / bias gradient of the same layer: dW[j][c] = sum_r dy[r][j] x[r][c], db[j] = sum_r dy[r][j].  Same lane layout;
// a lane accumulates the 8x8 block of its channels over its rows, the block combines its warps through shared memory
// and writes ONE partial [8*Cpad + 8] per block; k_pointwise8_reduce sums the partials in a fixed order (deterministic).
template <int LPR>
__global__ void __launch_bounds__(256, 2)
k_pointwise8_planes_wgrad(const uint4* __restrict__ hi, const uint4* __restrict__ lo, const float* __restrict__ dy,
                          float* __restrict__ partial, long long rows) {
  constexpr int RPW = 32 / LPR, CP = LPR * 8, NOUT = 8 * CP + 8;
  __shared__ float sm[8][NOUT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rw = lane / LPR;
  float acc[8][8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    db[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  }
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int U = 2;                                           // rows in flight per lane: 4*U 16-byte loads
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (long long r0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW + rw; r0 < rows; r0 += warps * RPW * U) {
    uint4 H[U], L[U];
    float4 G0[U], G1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * warps * RPW;
      const bool live = r < rows;                                // dead rows contribute x = 0, dy = 0
      H[u] = live ? __ldg(hi + r * LPR + sub) : zero;
      L[u] = (live && lo) ? __ldg(lo + r * LPR + sub) : zero;
      G0[u] = live ? __ldg(reinterpret_cast<const float4*>(dy + r * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
      G1[u] = live ? __ldg(reinterpret_cast<const float4*>(dy + r * 8) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[8];
      merge8(H[u], L[u], v);
      const float g[8] = {G0[u].x, G0[u].y, G0[u].z, G0[u].w, G1[u].x, G1[u].y, G1[u].z, G1[u].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        db[j] += g[j];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(g[j], v[c], acc[j][c]);
      }
    }
  }
  // rows of one warp pass -> lanes sub, sub + LPR, ...: fold them onto rw == 0
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      db[j] += __shfl_xor_sync(0xffffffffu, db[j], off);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] += __shfl_xor_sync(0xffffffffu, acc[j][c], off);
    }
  }
  if (rw == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) sm[warp][j * CP + 8 * sub + c] = acc[j][c];
    if (sub == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[warp][8 * CP + j] = db[j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NOUT; e += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sm[wv][e];
    partial[(long long)blockIdx.x * NOUT + e] = t;
  }
}

// dw[j][c] (c < Cin) and db[j] from the per-block partials, summed in block order in double
__global__ void k_pointwise8_reduce(const float* __restrict__ partial, int nblocks, int CP, int Cin, float* __restrict__ dw,
                                    float* __restrict__ db) {
  const int NOUT = 8 * CP + 8;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NOUT) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += (double)partial[(long long)b * NOUT + e];
  if (e < 8 * CP) {
    const int j = e / CP, c = e - j * CP;
    if (c < Cin) dw[j * Cin + c] = (float)t;
  } else if (db) {
    db[e - 8 * CP] = (float)t;
  }
}

}  // namespace dram

using namespace dram;

extern "C" {

int dram_bn_relu_apply_planes(const float* y, const float* scale, const float* shift, void* a_hi, void* a_lo, void* p_hi,
                              void* p_lo, int N, int D, int H, int W, int C, int Cpad, void* stream) {
  DRAM_REQUIRE(y && scale && shift && a_hi && N > 0 && D > 0 && H > 0 && W > 0, "bn_relu_apply_planes: bad arguments");
  DRAM_REQUIRE(C > 0 && C % 8 == 0 && Cpad >= C && Cpad % 8 == 0, "bn_relu_apply_planes: C=%d must be a multiple of 8 (Cpad=%d)", C, Cpad);
  DRAM_REQUIRE(!p_lo || p_hi, "bn_relu_apply_planes: p_lo without p_hi");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)N * D * H * W;
  if (!p_hi) {
    k_bn_relu_apply_planes<<<grid_fixed_group(rows, Cpad / 8, 256), 256, 0, st>>>(y, scale, shift, (uint4*)a_hi, (uint4*)a_lo, rows, C, Cpad);
  } else {
    DRAM_REQUIRE(D >= 2 && H >= 2 && W >= 2, "bn_relu_apply_planes: pooling needs every spatial size >= 2");
    DRAM_REQUIRE(rows < (1ll << 31), "bn_relu_apply_planes: volume too large");
    const long long cells = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
    k_bn_relu_pool_planes<<<grid_fixed_group(cells, Cpad / 8, 256), 256, 0, st>>>(y, scale, shift, (uint4*)a_hi, (uint4*)a_lo,
                                                                            (uint4*)p_hi, (uint4*)p_lo, N, D, H, W, C, Cpad);
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_relu_bwd_apply_planes(const float* da, long long da_pitch, const float* wtop, const float* y, const float* scale,
                                  const float* shift, const float* mean, const float* rstd, const float* gamma,
                                  const double* sums, double count, void* dy_hi, void* dy_lo, long long rows, int C, int Cpad,
                                  void* stream) {
  DRAM_REQUIRE(da && y && scale && shift && dy_hi && rows > 0, "bn_relu_bwd_apply_planes: bad arguments");
  DRAM_REQUIRE(C > 0 && C % 8 == 0 && Cpad >= C && Cpad % 8 == 0, "bn_relu_bwd_apply_planes: C=%d must be a multiple of 8 (Cpad=%d)", C, Cpad);
  DRAM_REQUIRE(!sums || (mean && rstd && count > 0), "bn_relu_bwd_apply_planes: training mode needs mean/rstd/count");
  if (da_pitch == 0) da_pitch = C;
  DRAM_REQUIRE(wtop || (da_pitch >= C && da_pitch % 4 == 0), "bn_relu_bwd_apply_planes: da_pitch %lld invalid for C %d", da_pitch, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_fixed_group(rows, Cpad / 8, 256);
  if (wtop)
    k_bn_relu_bwd_apply_planes<true><<<grid, 256, 0, st>>>(da, 0, wtop, y, scale, shift, mean, rstd, gamma, sums, count,
                                                           (uint4*)dy_hi, (uint4*)dy_lo, rows, C, Cpad);
  else
    k_bn_relu_bwd_apply_planes<false><<<grid, 256, 0, st>>>(da, da_pitch, nullptr, y, scale, shift, mean, rstd, gamma, sums,
                                                            count, (uint4*)dy_hi, (uint4*)dy_lo, rows, C, Cpad);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_pool_bwd_reduce(const float* ga, long long ga_pitch, const float* gp, const float* y, const float* scale,
                            const float* shift, const float* mean, const float* rstd, double* sums, int N, int D, int H,
                            int W, int C, void* stream) {
  DRAM_REQUIRE((ga || gp) && y && scale && shift && mean && rstd && sums && N > 0 && D >= 2 && H >= 2 && W >= 2,
               "bn_pool_bwd_reduce: bad arguments");
  DRAM_REQUIRE(C > 0 && C % 4 == 0, "bn_pool_bwd_reduce: C=%d must be a multiple of 4", C);
  if (ga_pitch == 0) ga_pitch = C;
  DRAM_REQUIRE(ga_pitch >= C && ga_pitch % 4 == 0, "bn_pool_bwd_reduce: ga_pitch %lld invalid for C %d", ga_pitch, C);
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  const long long cells = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
  long long want = (cells + 255) / 256;                     // ~256 windows (2048 voxels) per block, whole waves
  long long waves = (want + kNumSMs - 1) / kNumSMs;
  if (waves > 8) waves = 8;
  long long grid = waves * kNumSMs;
  if (grid > cells) grid = cells;
  k_bn_pool_bwd_reduce<<<(int)grid, kPoolThreads, 0, st>>>(ga, ga_pitch, gp, y, scale, shift, mean, rstd, sums, N, D, H, W, C);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_pool_bwd_apply_planes(const float* ga, long long ga_pitch, const float* gp, const float* y, const float* scale,
                                  const float* shift, const float* mean, const float* rstd, const float* gamma,
                                  const double* sums, double count, void* dy_hi, void* dy_lo, int N, int D, int H, int W,
                                  int C, int Cpad, void* stream) {
  DRAM_REQUIRE((ga || gp) && y && scale && shift && dy_hi && N > 0 && D >= 2 && H >= 2 && W >= 2, "bn_pool_bwd_apply_planes: bad arguments");
  DRAM_REQUIRE(C > 0 && C % 4 == 0 && Cpad >= C && Cpad % 8 == 0, "bn_pool_bwd_apply_planes: C=%d / Cpad=%d unsupported", C, Cpad);
  DRAM_REQUIRE(!sums || (mean && rstd && count > 0), "bn_pool_bwd_apply_planes: training mode needs mean/rstd/count");
  if (ga_pitch == 0) ga_pitch = C;
  DRAM_REQUIRE(ga_pitch >= C && ga_pitch % 4 == 0, "bn_pool_bwd_apply_planes: ga_pitch %lld invalid for C %d", ga_pitch, C);
  const long long cells = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
  k_bn_pool_bwd_apply_planes<<<grid_fixed_group(cells, Cpad / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      ga, ga_pitch, gp, y, scale, shift, mean, rstd, gamma, sums, count, (uint2*)dy_hi, (uint2*)dy_lo, N, D, H, W, C, Cpad);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static inline int ceil_half_(int a) { return (a + 1) / 2; }

int dram_upsample2x_concat_planes(const void* x_hi, const void* x_lo, const void* skip_hi, const void* skip_lo, void* cat_hi,
                                  void* cat_lo, int N, int d, int h, int w, int C1, int P1, int Ds, int Hs, int Ws, int C2,
                                  int P2, int Pc, void* stream) {
  DRAM_REQUIRE(x_hi && skip_hi && cat_hi && N > 0 && d > 0 && h > 0 && w > 0, "upsample2x_concat_planes: bad arguments");
  DRAM_REQUIRE((x_lo == nullptr) == (skip_lo == nullptr) && (x_lo == nullptr) == (cat_lo == nullptr),
               "upsample2x_concat_planes: lo planes must be all set or all NULL");
  DRAM_REQUIRE(C1 > 0 && C2 > 0 && C1 % 8 == 0 && C2 % 8 == 0 && P1 >= C1 && P2 >= C2 && Pc >= C1 + C2 && P1 % 8 == 0 &&
               P2 % 8 == 0 && Pc % 8 == 0, "upsample2x_concat_planes: channel counts must be multiples of 8 (C1=%d C2=%d)", C1, C2);
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  DRAM_REQUIRE(Ds >= D && Hs >= H && Ws >= W, "upsample2x_concat_planes: skip (%d,%d,%d) smaller than upsampled (%d,%d,%d)", Ds, Hs, Ws, D, H, W);
  const long long total = (long long)N * D * H * W * (Pc / 8);
  DRAM_REQUIRE((long long)N * D * H * W < (1ll << 31), "upsample2x_concat_planes: volume too large");
  static const bool blocked = getenv("DRAM_UP2X_SIMPLE") == nullptr;
  if (blocked) {
    cudaStream_t st = (cudaStream_t)stream;
    static const bool zslide = getenv("DRAM_UP2X_BLOCKED") == nullptr;
    if (zslide) {
      // z segments: enough columns for ~4 waves of 128-thread blocks, at least 8 blocks of z per thread
      const long long cols1 = (long long)N * h * w * (C1 / 8);
      int nseg = (int)((4ll * kNumSMs * 8 * 128 + cols1 - 1) / cols1);
      if (nseg > d / 8) nseg = d / 8;
      if (nseg < 1) nseg = 1;
      const int zseg = (d + nseg - 1) / nseg;
      nseg = (d + zseg - 1) / zseg;
      k_up2x_planes_zslide<<<grid_fixed_group((long long)N * nseg * h * w, C1 / 8, 128, 64), 128, 0, st>>>(
          (const uint4*)x_hi, (const uint4*)x_lo, (uint4*)cat_hi, (uint4*)cat_lo, N, d, h, w, C1, P1, Pc, ac_scale(d, D),
          ac_scale(h, H), ac_scale(w, W), zseg, nseg);
    } else {
      k_up2x_planes_blocked<<<grid_fixed_group((long long)N * d * h * w, C1 / 8, 128, 64), 128, 0, st>>>(
          (const uint4*)x_hi, (const uint4*)x_lo, (uint4*)cat_hi, (uint4*)cat_lo, N, d, h, w, C1, P1, Pc, ac_scale(d, D),
          ac_scale(h, H), ac_scale(w, W));
    }
    DRAM_LAUNCH_CHECK();
    k_skip_copy_planes<<<grid_for((long long)N * D * H * W * ((Pc - C1) / 8), 256), 256, 0, st>>>(
        (const uint4*)skip_hi, (const uint4*)skip_lo, (uint4*)cat_hi, (uint4*)cat_lo, N, D, H, W, C1, Ds, Hs, Ws, C2, P2, Pc,
        ceil_half_(Ds - D), ceil_half_(Hs - H), ceil_half_(Ws - W));
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  k_upsample2x_concat_planes<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x_hi, (const uint4*)x_lo, (const uint4*)skip_hi, (const uint4*)skip_lo, (uint4*)cat_hi, (uint4*)cat_lo, N, d,
      h, w, C1, P1, Ds, Hs, Ws, C2, P2, Pc, ceil_half_(Ds - D), ceil_half_(Hs - H), ceil_half_(Ws - W), ac_scale(d, D),
      ac_scale(h, H), ac_scale(w, W));
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_upsample2x_planes(const void* x_hi, const void* x_lo, void* out_hi, void* out_lo, int N, int d, int h, int w, int C1,
                           int P1, int Pout, void* stream) {
  DRAM_REQUIRE(x_hi && out_hi && N > 0 && d > 0 && h > 0 && w > 0, "upsample2x_planes: bad arguments");
  DRAM_REQUIRE((x_lo == nullptr) == (out_lo == nullptr), "upsample2x_planes: lo planes must be both set or both NULL");
  DRAM_REQUIRE(C1 > 0 && C1 % 8 == 0 && P1 >= C1 && P1 % 8 == 0 && Pout == C1,
               "upsample2x_planes: C1=%d must be a multiple of 8 and fill the output rows (Pout=%d)", C1, Pout);
  DRAM_REQUIRE((long long)N * 8 * d * h * w < (1ll << 31), "upsample2x_planes: volume too large");
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const long long cols1 = (long long)N * h * w * (C1 / 8);
  int nseg = (int)((4ll * kNumSMs * 8 * 128 + cols1 - 1) / cols1);
  if (nseg > d / 8) nseg = d / 8;
  if (nseg < 1) nseg = 1;
  const int zseg = (d + nseg - 1) / nseg;
  nseg = (d + zseg - 1) / zseg;
  k_up2x_planes_zslide<<<grid_fixed_group((long long)N * nseg * h * w, C1 / 8, 128, 64), 128, 0, (cudaStream_t)stream>>>(
      (const uint4*)x_hi, (const uint4*)x_lo, (uint4*)out_hi, (uint4*)out_lo, N, d, h, w, C1, P1, Pout, ac_scale(d, D),
      ac_scale(h, H), ac_scale(w, W), zseg, nseg);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_merge_planes(const void* hi, const void* lo, float* out, long long rows, int C, int Cpad, void* stream) {
  DRAM_REQUIRE(hi && out && rows > 0 && C > 0 && C % 8 == 0 && Cpad >= C && Cpad % 8 == 0, "merge_planes: bad arguments (C=%d Cpad=%d)", C, Cpad);
  k_merge_planes<<<grid_for(rows * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)hi, (const uint4*)lo, out, rows, C, Cpad);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static const int kPointwiseBlocks = kNumSMs * 2;

int dram_pointwise8_planes_supported(int Cin_pad, int Cout) { return (Cout == 8 && (Cin_pad == 64 || Cin_pad == 128)) ? 1 : 0; }

int dram_pointwise8_planes_fwd(const void* x_hi, const void* x_lo, const float* w, const float* bias, float* y, long long rows,
                               int Cin, int Cin_pad, void* stream) {
  DRAM_REQUIRE(x_hi && w && y && rows > 0, "pointwise8_planes_fwd: bad arguments");
  DRAM_REQUIRE((Cin_pad == 64 || Cin_pad == 128) && Cin > 0 && Cin <= Cin_pad, "pointwise8_planes_fwd: Cin_pad %d unsupported (64 | 128)", Cin_pad);
  const int lpr = Cin_pad / 8;
  const int grid = grid_for((rows + (32 / lpr) - 1) / (32 / lpr) * 32, 256, 8);
  if (lpr == 8)
    k_pointwise8_planes_fwd<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x_hi, (const uint4*)x_lo, w, bias, y, rows, Cin);
  else
    k_pointwise8_planes_fwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x_hi, (const uint4*)x_lo, w, bias, y, rows, Cin);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

size_t dram_pointwise8_planes_wgrad_workspace_bytes(int Cin_pad) { return sizeof(float) * (size_t)kPointwiseBlocks * (8 * (size_t)Cin_pad + 8); }

int dram_pointwise8_planes_wgrad(const void* x_hi, const void* x_lo, const float* dy, float* dw, float* dbias, void* workspace,
                                 long long rows, int Cin, int Cin_pad, void* stream) {
  DRAM_REQUIRE(x_hi && dy && dw && workspace && rows > 0, "pointwise8_planes_wgrad: bad arguments");
  DRAM_REQUIRE((Cin_pad == 64 || Cin_pad == 128) && Cin > 0 && Cin <= Cin_pad, "pointwise8_planes_wgrad: Cin_pad %d unsupported (64 | 128)", Cin_pad);
  const int lpr = Cin_pad / 8;
  float* partial = (float*)workspace;
  if (lpr == 8)
    k_pointwise8_planes_wgrad<8><<<kPointwiseBlocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x_hi, (const uint4*)x_lo, dy, partial, rows);
  else
    k_pointwise8_planes_wgrad<16><<<kPointwiseBlocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x_hi, (const uint4*)x_lo, dy, partial, rows);
  DRAM_LAUNCH_CHECK();
  const int nout = 8 * Cin_pad + 8;
  k_pointwise8_reduce<<<(nout + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, kPointwiseBlocks, Cin_pad, Cin, dw, dbias);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
