// BatchNorm / ReLU / MaxPool / trilinear upsample+concat / resize kernels on channels-last fp32 volumes.
// All of them are HBM-bound streaming kernels: float4 accesses along the contiguous channel dimension, grids sized
// as multiples of the 148 SMs with grid-stride loops, per-block reductions finished with a few double atomics.
#include <stdarg.h>
#include <mutex>
#include <stdlib.h>
#include "common.cuh"

namespace dram {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ layout
__global__ void k_nc2cl(const float* __restrict__ src, float* __restrict__ dst, int C, long long S, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    long long s = r % S, n = r / S;
    dst[i] = src[(n * C + c) * S + s];
  }
}
__global__ void k_cl2nc(const float* __restrict__ src, float* __restrict__ dst, int C, long long S, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long s = i % S;
    long long r = i / S;
    int c = (int)(r % C);
    long long n = r / C;
    dst[i] = src[(n * S + s) * C + c];
  }
}

// ------------------------------------------------------------------------------------------------ BN statistics
// y [rows][C].  Block = 256 threads arranged as (row lanes) x (channel groups of VEC).  The grid is a multiple of the 148
// SMs and every block owns one contiguous, equally sized range of rows; each thread keeps UNROLL independent loads in
// flight (the kernel is a pure HBM stream), accumulates in fp32 over <= a few hundred rows, the block reduces across
// row lanes in shared memory and adds its per-channel partial to the double accumulators.
constexpr int kStatThreads = 256;
constexpr int kStatUnroll = 4;

// MODE 0: forward statistics (sum y, sum y^2).  MODE 1: backward sums (sum dz, sum dz*xhat) with dz = da * relu'; `da` rows
// are `da_pitch` floats apart (a channel slice of a wider gradient tensor is read in place).  MODE 2: the same with the
// rank-1 gradient of the fused RAM head, da[r][c] = g[r] * wtop[c] (`da` = g, one float per row), plus the RAM head's own
// sums: sums[2C + c] = sum g * relu(y*scale+shift)  (d top_layer.weight) and sums[3C] = sum g (d top_layer.bias).
template <int VEC, int MODE>
__global__ void __launch_bounds__(kStatThreads)
k_bn_reduce(const float* __restrict__ y, const float* __restrict__ da, long long da_pitch, const float* __restrict__ wtop,
            const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
            const float* __restrict__ rstd, double* __restrict__ sums, long long rows, int C) {
  constexpr bool BWD = MODE != 0;
  extern __shared__ float sm[];  // [2 or 4][kStatThreads][VEC]
  const int groups = (C + VEC - 1) / VEC;             // channel groups per row
  const int gl = groups < kStatThreads ? groups : kStatThreads;
  const int lanes = kStatThreads / gl;                // row lanes per pass
  const long long per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r_begin = (long long)blockIdx.x * per_block;
  const long long r_end = r_begin + per_block < rows ? r_begin + per_block : rows;
  // when groups > kStatThreads each thread loops over several channel groups
  for (int g0 = 0; g0 < groups; g0 += kStatThreads) {
    const int g = g0 + (threadIdx.x % gl);
    const int lane = threadIdx.x / gl;
    float a0[VEC], a1[VEC], a2[VEC], a3 = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) a0[v] = a1[v] = a2[v] = 0.f;
    const bool active = (g < groups) && (lane < lanes);
    float sc[VEC], sh[VEC], mu[VEC], rs[VEC], wt[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int c = g * VEC + v;
      bool ok = BWD && active && c < C;
      sc[v] = ok ? scale[c] : 0.f; sh[v] = ok ? shift[c] : 0.f; mu[v] = ok ? mean[c] : 0.f; rs[v] = ok ? rstd[c] : 0.f;
      wt[v] = (MODE == 2 && ok) ? wtop[c] : 0.f;
    }
    if (active) {
      for (long long r0 = r_begin + lane; r0 < r_end; r0 += (long long)lanes * kStatUnroll) {
        float yv[kStatUnroll][VEC], dv[kStatUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kStatUnroll; ++u) {                 // issue all loads first
          const long long r = r0 + (long long)u * lanes;
          const bool rok = r < r_end;
          if (VEC == 4) {
            float4 t = rok ? __ldg(reinterpret_cast<const float4*>(y + r * C) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            yv[u][0] = t.x; yv[u][1] = t.y; yv[u][2] = t.z; yv[u][3 % VEC] = t.w;
            if (MODE == 1) {
              float4 q = rok ? __ldg(reinterpret_cast<const float4*>(da + r * da_pitch) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
              dv[u][0] = q.x; dv[u][1] = q.y; dv[u][2] = q.z; dv[u][3 % VEC] = q.w;
            }
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const bool ok = rok && (g * VEC + v < C);
              yv[u][v] = ok ? y[r * C + (long long)g * VEC + v] : 0.f;
              if (MODE == 1) dv[u][v] = ok ? da[r * da_pitch + (long long)g * VEC + v] : 0.f;
            }
          }
          if (MODE == 2) {
            const float gr = rok ? __ldg(da + r) : 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) dv[u][v] = gr;          // g[r]; multiplied by wtop[c] below
          }
        }
#pragma unroll
        for (int u = 0; u < kStatUnroll; ++u) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (!BWD) {
              a0[v] += yv[u][v];
              a1[v] += yv[u][v] * yv[u][v];
            } else {
              // out-of-range rows carry dv = 0, so they contribute nothing
              const float act = yv[u][v] * sc[v] + sh[v];
              float dz = (act > 0.f) ? dv[u][v] : 0.f;
              if (MODE == 2) {
                a2[v] += dz * act;                                   // g * relu(act)
                dz *= wt[v];
              }
              a0[v] += dz;
              a1[v] += dz * (yv[u][v] - mu[v]) * rs[v];
            }
          }
          if (MODE == 2 && g == 0) a3 += dv[u][0];
        }
      }
    }
    // cross-lane reduction in shared memory
    float* s0 = sm;
    float* s1 = sm + kStatThreads * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      s0[threadIdx.x * VEC + v] = a0[v];
      s1[threadIdx.x * VEC + v] = a1[v];
    }
    __syncthreads();
    if (active && lane == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        int c = g * VEC + v;
        if (c < C) {
          double t0 = 0.0, t1 = 0.0;
          for (int l = 0; l < lanes; ++l) {
            t0 += (double)s0[(l * gl + (threadIdx.x % gl)) * VEC + v];
            t1 += (double)s1[(l * gl + (threadIdx.x % gl)) * VEC + v];
          }
          atomicAdd(&sums[c], t0);
          atomicAdd(&sums[C + c], t1);
        }
      }
    }
    __syncthreads();
    if (MODE == 2) {                                                 // second round for the RAM head's sums
#pragma unroll
      for (int v = 0; v < VEC; ++v) s0[threadIdx.x * VEC + v] = a2[v];
      s1[threadIdx.x] = a3;
      __syncthreads();
      if (active && lane == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          int c = g * VEC + v;
          if (c < C) {
            double t0 = 0.0;
            for (int l = 0; l < lanes; ++l) t0 += (double)s0[(l * gl + (threadIdx.x % gl)) * VEC + v];
            atomicAdd(&sums[2 * C + c], t0);
          }
        }
        if (g == 0) {
          double t1 = 0.0;
          for (int l = 0; l < lanes; ++l) t1 += (double)s1[l * gl];
          atomicAdd(&sums[3 * C], t1);
        }
      }
      __syncthreads();
    }
  }
}

__global__ void k_bn_finalize(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                              float eps, int n_updates, float* mean, float* rstd, float* scale, float* shift, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sums[c] / count;
  double var = sums[C + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  float r = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = (float)m;
  rstd[c] = r;
  scale[c] = g * r;
  shift[c] = b - (float)m * g * r;
  if (running_mean) {
    float unbiased = count > 1.0 ? (float)(var * count / (count - 1.0)) : (float)var;
    float rm = running_mean[c], rv = running_var[c];
    for (int i = 0; i < n_updates; ++i) {
      rm = (1.f - momentum) * rm + momentum * (float)m;
      rv = (1.f - momentum) * rv + momentum * unbiased;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
}

__global__ void k_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                               float* scale, float* shift, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float r = 1.0f / sqrtf(rv[c] + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * r;
  shift[c] = b - rm[c] * g * r;
}

// ------------------------------------------------------------------------------------------------ BN apply + ReLU (+pool)
template <int VEC>
__global__ void k_bn_relu_apply(const float* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, float* __restrict__ a, long long rows, int C) {
  const int groups = C / VEC;
  const long long total = rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    if (VEC == 4) {
      float4 t = reinterpret_cast<const float4*>(y)[i];
      float4 sc = reinterpret_cast<const float4*>(scale)[g], sh = reinterpret_cast<const float4*>(shift)[g];
      t.x = fmaxf(t.x * sc.x + sh.x, 0.f); t.y = fmaxf(t.y * sc.y + sh.y, 0.f);
      t.z = fmaxf(t.z * sc.z + sh.z, 0.f); t.w = fmaxf(t.w * sc.w + sh.w, 0.f);
      reinterpret_cast<float4*>(a)[i] = t;
    } else {
      a[i] = fmaxf(y[i] * scale[g] + shift[g], 0.f);
    }
  }
}

// one thread per (pool cell, channel group): writes the 8 activations of the cell and their max
template <int VEC>
__global__ void k_bn_relu_pool(const float* __restrict__ y, const float* __restrict__ scale,
                               const float* __restrict__ shift, float* __restrict__ a, float* __restrict__ pooled,
                               int N, int D, int H, int W, int C) {
  const int groups = C / VEC;
  const int cd = (D + 1) / 2, ch = (H + 1) / 2, cw = (W + 1) / 2;   // cells incl. ragged tail
  const int pd = D / 2, ph = H / 2, pw = W / 2;                       // pooled size (floor)
  const long long total = (long long)N * cd * ch * cw * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    long long r = i / groups;
    int x = (int)(r % cw); r /= cw;
    int yy = (int)(r % ch); r /= ch;
    int z = (int)(r % cd);
    int n = (int)(r / cd);
    float sc[VEC], sh[VEC], mx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { sc[v] = scale[g * VEC + v]; sh[v] = shift[g * VEC + v]; mx[v] = -INFINITY; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int zz = 2 * z + (k >> 2), yv = 2 * yy + ((k >> 1) & 1), xv = 2 * x + (k & 1);
      if (zz < D && yv < H && xv < W) {
        long long off = ((((long long)n * D + zz) * H + yv) * W + xv) * C + (long long)g * VEC;
        float t[VEC];
        if (VEC == 4) {
          float4 q = *reinterpret_cast<const float4*>(y + off);
          t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3 % VEC] = q.w;
        } else {
          t[0] = y[off];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) { t[v] = fmaxf(t[v] * sc[v] + sh[v], 0.f); mx[v] = fmaxf(mx[v], t[v]); }
        if (VEC == 4) *reinterpret_cast<float4*>(a + off) = make_float4(t[0], t[1], t[2], t[3 % VEC]);
        else a[off] = t[0];
      }
    }
    if (z < pd && yy < ph && x < pw) {
      long long po = ((((long long)n * pd + z) * ph + yy) * pw + x) * C + (long long)g * VEC;
      if (VEC == 4) *reinterpret_cast<float4*>(pooled + po) = make_float4(mx[0], mx[1], mx[2], mx[3 % VEC]);
      else pooled[po] = mx[0];
    }
  }
}

template <int VEC>
__global__ void k_bn_relu_bwd_apply(const float* __restrict__ da, const float* __restrict__ y,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sums, double count,
                                    float* __restrict__ dy, long long rows, int C, long long da_pitch) {
  const int groups = C / VEC;
  const long long total = rows * groups;
  const float inv = sums ? (float)(1.0 / count) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    const long long r = i / groups;
    float yv[VEC], dv[VEC], o[VEC];
    if (VEC == 4) {
      float4 t = reinterpret_cast<const float4*>(y)[i], u = reinterpret_cast<const float4*>(da + r * da_pitch)[g];
      yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3 % VEC] = t.w;
      dv[0] = u.x; dv[1] = u.y; dv[2] = u.z; dv[3 % VEC] = u.w;
    } else {
      yv[0] = y[i]; dv[0] = da[r * da_pitch + g];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int c = g * VEC + v;
      float sc = scale[c];
      float dz = (yv[v] * sc + shift[c] > 0.f) ? dv[v] : 0.f;
      if (sums) {
        float xh = (yv[v] - mean[c]) * rstd[c];
        float gm = gamma ? gamma[c] : 1.f;
        o[v] = gm * rstd[c] * (dz - (float)sums[c] * inv - xh * (float)sums[C + c] * inv);
      } else {
        o[v] = dz * sc;
      }
    }
    if (VEC == 4) reinterpret_cast<float4*>(dy)[i] = make_float4(o[0], o[1], o[2], o[3 % VEC]);
    else dy[i] = o[0];
  }
}

// MaxPool3d(2,2,0) backward: first maximum in (d,h,w) scan order receives the gradient (ATen semantics: val > max).
template <int VEC>
__global__ void k_maxpool2_bwd(const float* __restrict__ a, const float* __restrict__ dpooled, float* __restrict__ da,
                               int N, int D, int H, int W, int C) {
  const int groups = C / VEC;
  const int pd = D / 2, ph = H / 2, pw = W / 2;
  const long long total = (long long)N * pd * ph * pw * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    long long r = i / groups;
    int x = (int)(r % pw); r /= pw;
    int yy = (int)(r % ph); r /= ph;
    int z = (int)(r % pd);
    int n = (int)(r / pd);
    float mx[VEC];
    int arg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { mx[v] = -INFINITY; arg[v] = 0; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      long long off = ((((long long)n * D + 2 * z + (k >> 2)) * H + 2 * yy + ((k >> 1) & 1)) * W + 2 * x + (k & 1)) * C + (long long)g * VEC;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float t = a[off + v];
        if (t > mx[v] || t != t) { mx[v] = t; arg[v] = k; }
      }
    }
    long long po = ((((long long)n * pd + z) * ph + yy) * pw + x) * C + (long long)g * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int k = arg[v];
      long long off = ((((long long)n * D + 2 * z + (k >> 2)) * H + 2 * yy + ((k >> 1) & 1)) * W + 2 * x + (k & 1)) * C + (long long)g * VEC + v;
      da[off] += dpooled[po + v];
    }
  }
}

// ------------------------------------------------------------------------------------------------ resize / upsample+concat
// dst[n][Z][Y][X][co + c] = trilinear(src)[c] for c < Csrc.  Generic in the destination channel stride so that the
// same kernel writes the "up" half of the concat buffer.
template <int VEC>
__global__ void k_trilinear_fwd(const float* __restrict__ src, float* __restrict__ dst, int N, int d, int h, int w,
                                int D, int H, int W, int C, int dstC, int dstOff, float sz, float sy, float sx) {
  const int groups = C / VEC;
  const long long total = (long long)N * D * H * W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    long long r = i / groups;
    int X = (int)(r % W); r /= W;
    int Y = (int)(r % H); r /= H;
    int Z = (int)(r % D);
    int n = (int)(r / D);
    Lerp lz = lerp_setup(Z, sz, d), ly = lerp_setup(Y, sy, h), lx = lerp_setup(X, sx, w);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int zi = (k & 4) ? lz.i1 : lz.i0, yi = (k & 2) ? ly.i1 : ly.i0, xi = (k & 1) ? lx.i1 : lx.i0;
      float wt = ((k & 4) ? lz.w1 : lz.w0) * ((k & 2) ? ly.w1 : ly.w0) * ((k & 1) ? lx.w1 : lx.w0);
      const float* p = src + ((((long long)n * d + zi) * h + yi) * w + xi) * C + (long long)g * VEC;
      if (VEC == 4) {
        float4 t = *reinterpret_cast<const float4*>(p);
        acc[0] += wt * t.x; acc[1] += wt * t.y; acc[2] += wt * t.z; acc[3 % VEC] += wt * t.w;
      } else {
        acc[0] += wt * p[0];
      }
    }
    float* q = dst + ((((long long)n * D + Z) * H + Y) * W + X) * dstC + dstOff + (long long)g * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(acc[0], acc[1], acc[2], acc[3 % VEC]);
    else q[0] = acc[0];
  }
}

// candidate destination range along one axis whose taps may touch source index i: o*scale in (i-1, i+1), one index of
// slack on both sides for the float rounding of o*scale
__device__ __forceinline__ void adj_range(int i, float scale, int out_size, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }
  float inv = 1.0f / scale;
  lo = (int)ceilf((float)(i - 1) * inv) - 1;
  hi = (int)floorf((float)(i + 1) * inv) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}
// weight with which destination index O contributes to source index i (0 if it does not)
__device__ __forceinline__ float adj_weight(int O, int i, float scale, int in_size) {
  Lerp l = lerp_setup(O, scale, in_size);
  return (l.i0 == i ? l.w0 : 0.f) + (l.i1 == i ? l.w1 : 0.f);
}

// exact adjoint of k_trilinear_fwd in gather form: dsrc[n][z][y][x][c] = sum over dst voxels of weight * ddst.
// The per-axis candidate weights are evaluated once (<= kAdj candidates per axis when the resize factor is <= ~3,
// which covers every resize on the training path); wider ranges take the generic loop.
constexpr int kAdj = 8;
template <int VEC>
__global__ void k_trilinear_bwd(const float* __restrict__ ddst, float* __restrict__ dsrc, int N, int d, int h, int w,
                                int D, int H, int W, int C, int dstC, int dstOff, float sz, float sy, float sx) {
  const int groups = C / VEC;
  const long long total = (long long)N * d * h * w * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % groups);
    long long r = i / groups;
    int x = (int)(r % w); r /= w;
    int y = (int)(r % h); r /= h;
    int z = (int)(r % d);
    int n = (int)(r / d);
    int zl, zh, yl, yh, xl, xh;
    adj_range(z, sz, D, zl, zh); adj_range(y, sy, H, yl, yh); adj_range(x, sx, W, xl, xh);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    if (zh - zl < kAdj && yh - yl < kAdj && xh - xl < kAdj) {
      float wy[kAdj], wx[kAdj];
#pragma unroll
      for (int t = 0; t < kAdj; ++t) {
        wy[t] = (yl + t <= yh) ? adj_weight(yl + t, y, sy, h) : 0.f;
        wx[t] = (xl + t <= xh) ? adj_weight(xl + t, x, sx, w) : 0.f;
      }
      for (int Z = zl; Z <= zh; ++Z) {
        const float wz = adj_weight(Z, z, sz, d);
        if (wz == 0.f) continue;
#pragma unroll
        for (int ty = 0; ty < kAdj; ++ty) {
          if (wy[ty] == 0.f) continue;
          const float wzy = wz * wy[ty];
          const float* row = ddst + (((long long)n * D + Z) * H + (yl + ty)) * (long long)W * dstC + dstOff + (long long)g * VEC;
#pragma unroll
          for (int tx = 0; tx < kAdj; ++tx) {
            if (wx[tx] == 0.f) continue;
            const float wt = wzy * wx[tx];
            const float* p = row + (long long)(xl + tx) * dstC;
            if (VEC == 4) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p));
              acc[0] += wt * t.x; acc[1] += wt * t.y; acc[2] += wt * t.z; acc[3 % VEC] += wt * t.w;
            } else {
              acc[0] += wt * p[0];
            }
          }
        }
      }
    } else {
      for (int Z = zl; Z <= zh; ++Z) {
        const float wz = adj_weight(Z, z, sz, d);
        if (wz == 0.f) continue;
        for (int Y = yl; Y <= yh; ++Y) {
          const float wy = adj_weight(Y, y, sy, h);
          if (wy == 0.f) continue;
          for (int X = xl; X <= xh; ++X) {
            const float wx = adj_weight(X, x, sx, w);
            if (wx == 0.f) continue;
            float wt = wz * wy * wx;
            const float* p = ddst + ((((long long)n * D + Z) * H + Y) * W + X) * dstC + dstOff + (long long)g * VEC;
            if (VEC == 4) {
              float4 t = *reinterpret_cast<const float4*>(p);
              acc[0] += wt * t.x; acc[1] += wt * t.y; acc[2] += wt * t.z; acc[3 % VEC] += wt * t.w;
            } else {
              acc[0] += wt * p[0];
            }
          }
        }
      }
    }
    float* q = dsrc + i * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(acc[0], acc[1], acc[2], acc[3 % VEC]);
    else q[0] = acc[0];
  }
}

// Adjoint of the exact x2 upsample (D = 2d ...), gather form, register-blocked: one thread produces a 2x2x2 block of
// source voxels for 4 channels.  With align_corners=True and an exact factor 2, source index i receives from destination
// indices 2i-1 .. 2i+2 only, so the block needs the 6x6x6 destination neighbourhood (27 float4 loads per output instead of
// 64) and the three axes are reduced one after the other in registers.  Weights come from the forward's own lerp_setup,
// so this is the exact transpose of k_trilinear_fwd (a destination index that touches a source index only through float
// rounding of o*scale, weight ~1e-7, is ignored).
__global__ void __launch_bounds__(128)
k_up2x_adjoint(const float* __restrict__ ddst, float* __restrict__ dsrc, int N, int d, int h, int w, int C, int dstC,
               int dstOff, float sz, float sy, float sx) {
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const int groups = C / 4;
  const int bd = (d + 1) / 2, bh = (h + 1) / 2, bw = (w + 1) / 2;
  const long long blocks = (long long)N * bd * bh * bw;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long db = ((long long)gridDim.x * blockDim.x) / groups;     // launch: total threads % groups == 0
  const int g = (int)(tid % groups);
  for (long long b = tid / groups; b < blocks; b += db) {
    unsigned r = (unsigned)b;
    const int x0 = 2 * (int)(r % (unsigned)bw); r /= (unsigned)bw;
    const int y0 = 2 * (int)(r % (unsigned)bh); r /= (unsigned)bh;
    const int z0 = 2 * (int)(r % (unsigned)bd);
    const int n = (int)(r / (unsigned)bd);
    float wz[6][2], wy[6][2], wx[6][2];
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int Z = 2 * z0 - 1 + t, Y = 2 * y0 - 1 + t, X = 2 * x0 - 1 + t;
        wz[t][j] = (Z >= 0 && Z < D && z0 + j < d) ? adj_weight(Z, z0 + j, sz, d) : 0.f;
        wy[t][j] = (Y >= 0 && Y < H && y0 + j < h) ? adj_weight(Y, y0 + j, sy, h) : 0.f;
        wx[t][j] = (X >= 0 && X < W && x0 + j < w) ? adj_weight(X, x0 + j, sx, w) : 0.f;
      }
    // Destination indices are clamped into the volume and every load is unconditional (an out-of-range candidate has
    // weight 0): no data-dependent branch stands between the loads, so the 6 loads of a row and the rows of a plane
    // overlap instead of each paying a full memory latency.
    int Xc[6], Yc[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      Xc[t] = min(max(2 * x0 - 1 + t, 0), W - 1);
      Yc[t] = min(max(2 * y0 - 1 + t, 0), H - 1);
    }
    float4 out[2][2][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i >> 2][(i >> 1) & 1][i & 1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int tz = 0; tz < 6; ++tz) {
      const int Z = min(max(2 * z0 - 1 + tz, 0), D - 1);
      const long long zbase = ((long long)n * D + Z) * H;
      float4 u[2][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i >> 1][i & 1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ty = 0; ty < 6; ++ty) {
        const float* row = ddst + (zbase + Yc[ty]) * (long long)W * dstC + dstOff + g * 4;
        float4 v[6];
#pragma unroll
        for (int tx = 0; tx < 6; ++tx) v[tx] = __ldg(reinterpret_cast<const float4*>(row + (long long)Xc[tx] * dstC));
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
#pragma unroll
        for (int tx = 0; tx < 6; ++tx) {
          t0.x = fmaf(wx[tx][0], v[tx].x, t0.x); t0.y = fmaf(wx[tx][0], v[tx].y, t0.y); t0.z = fmaf(wx[tx][0], v[tx].z, t0.z); t0.w = fmaf(wx[tx][0], v[tx].w, t0.w);
          t1.x = fmaf(wx[tx][1], v[tx].x, t1.x); t1.y = fmaf(wx[tx][1], v[tx].y, t1.y); t1.z = fmaf(wx[tx][1], v[tx].z, t1.z); t1.w = fmaf(wx[tx][1], v[tx].w, t1.w);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float a = wy[ty][j];
          u[j][0].x = fmaf(a, t0.x, u[j][0].x); u[j][0].y = fmaf(a, t0.y, u[j][0].y); u[j][0].z = fmaf(a, t0.z, u[j][0].z); u[j][0].w = fmaf(a, t0.w, u[j][0].w);
          u[j][1].x = fmaf(a, t1.x, u[j][1].x); u[j][1].y = fmaf(a, t1.y, u[j][1].y); u[j][1].z = fmaf(a, t1.z, u[j][1].z); u[j][1].w = fmaf(a, t1.w, u[j][1].w);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float a = wz[tz][j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4& o = out[j][i >> 1][i & 1];
          const float4 uu = u[i >> 1][i & 1];
          o.x = fmaf(a, uu.x, o.x); o.y = fmaf(a, uu.y, o.y); o.z = fmaf(a, uu.z, o.z); o.w = fmaf(a, uu.w, o.w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int z = z0 + (i >> 2), y = y0 + ((i >> 1) & 1), x = x0 + (i & 1);
      if (z < d && y < h && x < w)
        *reinterpret_cast<float4*>(dsrc + ((((long long)n * d + z) * h + y) * w + x) * C + g * 4) = out[i >> 2][(i >> 1) & 1][i & 1];
    }
  }
}

// Adjoint of the exact x2 upsample, shared-memory tiled, marching along z (round 2).  The register-blocked kernel above
// pulls every destination value through L1/L2 27/8 times and ran at 0.37 of the HBM roofline.  Here a block owns an 8 x 8
// (y, x) tile of SOURCE voxels x 32 channels (128 contiguous bytes per voxel) x a z segment and walks the destination planes
// Z of that segment: the 18 x 18 destination patch of a plane is brought in ONCE with cp.async (double buffered), reduced
// along x into [18][8] and along y into [8][8] through shared memory (4 taps each: source index i receives from destination
// indices 2i-1 .. 2i+2 only), and the plane's result is scattered along z exactly as the forward gathered it
// (lerp_setup(Z): out[Z] = w0 in[i0] + w1 in[i1]  =>  din[i0] += w0 u, din[i1] += w1 u) into two rolling accumulators per
// thread; a source plane is written once its last contributing destination plane has passed.  Same weights as
// k_trilinear_fwd, so this is its exact transpose; only the order of the additions differs from k_up2x_adjoint.
constexpr int kAdjT = 8;                       // source tile edge
constexpr int kAdjP = 2 * kAdjT + 2;           // destination patch edge (18)
constexpr int kAdjC = 32;                      // channels per block
constexpr int kAdjTileFloats = kAdjP * kAdjP * kAdjC;
__global__ void __launch_bounds__(256, 2)
k_up2x_adjoint_tiled(const float* __restrict__ ddst, float* __restrict__ dsrc, int N, int d, int h, int w, int C, int dstC,
                     int dstOff, float sz, float sy, float sx, int zsegs) {
  extern __shared__ __align__(16) float sm_adj[];
  float* tile = sm_adj;                                        // [2][18][18][32]
  float* tbuf = sm_adj + 2 * kAdjTileFloats;                   // [18][8][32]  (x-reduced)
  float* wxs = tbuf + kAdjP * kAdjT * kAdjC;                   // [8][4]
  float* wys = wxs + kAdjT * 4;                                // [8][4]
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const int tx_n = (w + kAdjT - 1) / kAdjT, ty_n = (h + kAdjT - 1) / kAdjT, cg_n = C / kAdjC;
  unsigned r = blockIdx.x;
  const int x0 = (int)(r % (unsigned)tx_n) * kAdjT; r /= (unsigned)tx_n;
  const int y0 = (int)(r % (unsigned)ty_n) * kAdjT; r /= (unsigned)ty_n;
  const int c0 = (int)(r % (unsigned)cg_n) * kAdjC; r /= (unsigned)cg_n;
  const int seg = (int)(r % (unsigned)zsegs);
  const int n = (int)(r / (unsigned)zsegs);
  const int cz = (d + zsegs - 1) / zsegs, zc0 = seg * cz, zc1 = min(d, zc0 + cz);
  if (zc0 >= zc1) return;
  const int tid = threadIdx.x;
  if (tid < 2 * kAdjT * 4) {                                   // per-axis tap weights of the tile (0 outside the volume)
    const int axis = tid >> 5, i = (tid >> 2) & 7, t = tid & 3;
    const int src = (axis ? y0 : x0) + i, O = 2 * src - 1 + t, in_size = axis ? h : w;
    const float wgt = (src < in_size && O >= 0 && O < 2 * in_size) ? adj_weight(O, src, axis ? sy : sx, in_size) : 0.f;
    (axis ? wys : wxs)[i * 4 + t] = wgt;
  }
  const int Zbeg = max(0, 2 * zc0 - 2), Zend = min(D - 1, 2 * zc1 + 1);
  auto issue = [&](int Z, int buf) {                           // destination plane Z -> tile[buf], 16 bytes per cp.async
    float* dstp = tile + buf * kAdjTileFloats;
    const long long zbase = ((long long)n * D + Z) * H;
    for (int it = tid; it < kAdjP * kAdjP * (kAdjC / 4); it += 256) {
      const int g = it & 7, pos = it >> 3, lx = pos % kAdjP, ly = pos / kAdjP;
      const int Y = min(max(2 * y0 - 1 + ly, 0), H - 1), X = min(max(2 * x0 - 1 + lx, 0), W - 1);
      const float* src = ddst + ((zbase + Y) * W + X) * (long long)dstC + dstOff + c0 + g * 4;
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dstp + (pos * (kAdjC / 4) + g) * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // stage-2 items of this thread: (yc, xc, g) for item = tid and tid + 256
  float4 accA[2], accB[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) accA[k] = accB[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int curA = lerp_setup(Zbeg, sz, d).i0;
  auto flush_one = [&]() {                                     // source plane curA is complete: store it, roll the accumulators
    if (curA >= zc0 && curA < zc1) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int item = tid + k * 256, g = item & 7, xc = (item >> 3) & 7, yc = item >> 6;
        const int y = y0 + yc, x = x0 + xc;
        if (y < h && x < w)
          *reinterpret_cast<float4*>(dsrc + ((((long long)n * d + curA) * h + y) * w + x) * (long long)C + c0 + g * 4) = accA[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) { accA[k] = accB[k]; accB[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    ++curA;
  };
  issue(Zbeg, 0);
  for (int Z = Zbeg; Z <= Zend; ++Z) {
    const int buf = (Z - Zbeg) & 1;
    if (Z < Zend) issue(Z + 1, buf ^ 1);
    if (Z < Zend) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                           // tile[buf] complete; the previous plane's tbuf reads are done
    const float4* tp = reinterpret_cast<const float4*>(tile + buf * kAdjTileFloats);
    for (int it = tid; it < kAdjP * kAdjT * (kAdjC / 4); it += 256) {          // reduce along x: [18][18] -> [18][8]
      const int g = it & 7, xc = (it >> 3) & 7, ly = it >> 6;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float wv = wxs[xc * 4 + t];
        const float4 v = tp[(ly * kAdjP + 2 * xc + t) * (kAdjC / 4) + g];
        a.x = fmaf(wv, v.x, a.x); a.y = fmaf(wv, v.y, a.y); a.z = fmaf(wv, v.z, a.z); a.w = fmaf(wv, v.w, a.w);
      }
      reinterpret_cast<float4*>(tbuf)[it] = a;
    }
    __syncthreads();
    const Lerp lz = lerp_setup(Z, sz, d);
    while (curA < lz.i0) flush_one();
#pragma unroll
    for (int k = 0; k < 2; ++k) {                                              // reduce along y, scatter along z
      const int item = tid + k * 256, g = item & 7, xc = (item >> 3) & 7, yc = item >> 6;
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float wv = wys[yc * 4 + t];
        const float4 v = reinterpret_cast<const float4*>(tbuf)[((2 * yc + t) * kAdjT + xc) * (kAdjC / 4) + g];
        u.x = fmaf(wv, v.x, u.x); u.y = fmaf(wv, v.y, u.y); u.z = fmaf(wv, v.z, u.z); u.w = fmaf(wv, v.w, u.w);
      }
      accA[k].x = fmaf(lz.w0, u.x, accA[k].x); accA[k].y = fmaf(lz.w0, u.y, accA[k].y);
      accA[k].z = fmaf(lz.w0, u.z, accA[k].z); accA[k].w = fmaf(lz.w0, u.w, accA[k].w);
      float4& b = (lz.i1 == lz.i0) ? accA[k] : accB[k];
      b.x = fmaf(lz.w1, u.x, b.x); b.y = fmaf(lz.w1, u.y, b.y); b.z = fmaf(lz.w1, u.z, b.z); b.w = fmaf(lz.w1, u.w, b.w);
    }
  }
  flush_one();
  flush_one();
}

// copy the (centre-cropped) skip tensor into / out of channels [C1, C1+C2) of the concat buffer
template <int VEC, bool BWD>
__global__ void k_concat_skip(const float* __restrict__ in, float* __restrict__ out, int N, int D, int H, int W,
                              int Ds, int Hs, int Ws, int oz, int oy, int ox, int C1, int C2) {
  // FWD: in = skip [N][Ds][Hs][Ws][C2], out = cat [N][D][H][W][C1+C2]
  // BWD: in = dcat, out = dskip (every voxel of dskip is written; outside the crop -> 0)
  const int groups = C2 / VEC;
  const int Ct = C1 + C2;
  if (!BWD) {
    const long long total = (long long)N * D * H * W * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      int g = (int)(i % groups);
      long long r = i / groups;
      int X = (int)(r % W); r /= W;
      int Y = (int)(r % H); r /= H;
      int Z = (int)(r % D);
      int n = (int)(r / D);
      const float* p = in + ((((long long)n * Ds + Z + oz) * Hs + Y + oy) * Ws + X + ox) * C2 + (long long)g * VEC;
      float* q = out + ((((long long)n * D + Z) * H + Y) * W + X) * Ct + C1 + (long long)g * VEC;
      if (VEC == 4) *reinterpret_cast<float4*>(q) = *reinterpret_cast<const float4*>(p);
      else q[0] = p[0];
    }
  } else {
    const long long total = (long long)N * Ds * Hs * Ws * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      int g = (int)(i % groups);
      long long r = i / groups;
      int X = (int)(r % Ws); r /= Ws;
      int Y = (int)(r % Hs); r /= Hs;
      int Z = (int)(r % Ds);
      int n = (int)(r / Ds);
      int z = Z - oz, y = Y - oy, x = X - ox;
      bool inside = z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W;
      float* q = out + i * VEC;
      if (inside) {
        const float* p = in + ((((long long)n * D + z) * H + y) * W + x) * Ct + C1 + (long long)g * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(q) = *reinterpret_cast<const float4*>(p);
        else q[0] = p[0];
      } else {
        if (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(0.f, 0.f, 0.f, 0.f);
        else q[0] = 0.f;
      }
    }
  }
}

}  // namespace dram

using namespace dram;

#define VEC_DISPATCH(C, CALL4, CALL1) \
  do { if ((C) % 4 == 0) { CALL4; } else { CALL1; } } while (0)

extern "C" {

int dram_version(void) { return 100; }
int dram_sm_arch(void) { return 100; }
const char* dram_last_error(void) { return dram::last_error(); }
int dram_device_check(void) {
  int dev = 0, major = 0;
  DRAM_CUDA(cudaGetDevice(&dev));
  DRAM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("libdram_b200 is built for sm_100a only; current device has compute capability major %d", major);
    return DRAM_E_ARCH;
  }
  return DRAM_OK;
}

int dram_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, long long S, void* stream) {
  DRAM_REQUIRE(src && dst && N > 0 && C > 0 && S > 0, "ncdhw_to_ndhwc: bad arguments");
  long long total = (long long)N * C * S;
  k_nc2cl<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, C, S, total);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}
int dram_ndhwc_to_ncdhw(const float* src, float* dst, int N, int C, long long S, void* stream) {
  DRAM_REQUIRE(src && dst && N > 0 && C > 0 && S > 0, "ndhwc_to_ncdhw: bad arguments");
  long long total = (long long)N * C * S;
  k_cl2nc<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, C, S, total);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static int stat_grid(long long rows) {
  // whole waves of the 148 SMs, >= 256 rows per block, <= ~2048 rows per block (short fp32 partial sums)
  long long want = (rows + 2047) / 2048;
  long long waves = (want + kNumSMs - 1) / kNumSMs;
  if (waves > 8) waves = 8;
  long long g = waves * kNumSMs;
  if (g * 256 > rows) g = (rows + 255) / 256;
  return (int)(g < 1 ? 1 : g);
}

int dram_bn_stats(const float* y, double* sums, long long rows, int C, void* stream) {
  DRAM_REQUIRE(y && sums && rows > 0 && C > 0, "bn_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  if (C % 4 == 0)
    k_bn_reduce<4, 0><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * 4 * sizeof(float), st>>>(
        y, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, sums, rows, C);
  else
    k_bn_reduce<1, 0><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * sizeof(float), st>>>(
        y, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, sums, rows, C);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int n_updates, float* mean, float* rstd,
                     float* scale, float* shift, int C, void* stream) {
  DRAM_REQUIRE(sums && mean && rstd && scale && shift && C > 0 && count > 0, "bn_finalize: bad arguments");
  DRAM_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats must both be set");
  k_bn_finalize<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, beta, running_mean, running_var,
                                                                  momentum, eps, n_updates, mean, rstd, scale, shift, C);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, float* scale, float* shift, int C, void* stream) {
  DRAM_REQUIRE(running_mean && running_var && scale && shift && C > 0, "bn_fold_eval: bad arguments");
  k_bn_fold_eval<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, scale,
                                                                   shift, C);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_relu_apply(const float* y, const float* scale, const float* shift, float* a, float* pooled, int N, int D,
                       int H, int W, int C, void* stream) {
  DRAM_REQUIRE(y && scale && shift && a && N > 0 && D > 0 && H > 0 && W > 0 && C > 0, "bn_relu_apply: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)N * D * H * W;
  if (!pooled) {
    VEC_DISPATCH(C, (k_bn_relu_apply<4><<<grid_for(rows * (C / 4), 256), 256, 0, st>>>(y, scale, shift, a, rows, C)),
                 (k_bn_relu_apply<1><<<grid_for(rows * C, 256), 256, 0, st>>>(y, scale, shift, a, rows, C)));
  } else {
    DRAM_REQUIRE(D >= 2 && H >= 2 && W >= 2, "bn_relu_apply: pooling needs every spatial size >= 2");
    long long cells = (long long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
    VEC_DISPATCH(C, (k_bn_relu_pool<4><<<grid_for(cells * (C / 4), 256), 256, 0, st>>>(y, scale, shift, a, pooled, N, D, H, W, C)),
                 (k_bn_relu_pool<1><<<grid_for(cells * C, 256), 256, 0, st>>>(y, scale, shift, a, pooled, N, D, H, W, C)));
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_relu_bwd_reduce(const float* da, long long da_pitch, const float* wtop, const float* y, const float* scale,
                            const float* shift, const float* mean, const float* rstd, double* sums, long long rows, int C,
                            void* stream) {
  DRAM_REQUIRE(da && y && scale && shift && mean && rstd && sums && rows > 0 && C > 0, "bn_relu_bwd_reduce: bad arguments");
  if (da_pitch == 0) da_pitch = C;
  DRAM_REQUIRE(wtop || da_pitch >= C, "bn_relu_bwd_reduce: da_pitch %lld < C %d", da_pitch, C);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = C % 4 == 0 && da_pitch % 4 == 0;
  if (wtop) {
    DRAM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (3 * C + 1), st));
    if (vec)
      k_bn_reduce<4, 2><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * 4 * sizeof(float), st>>>(
          y, da, 0, wtop, scale, shift, mean, rstd, sums, rows, C);
    else
      k_bn_reduce<1, 2><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * sizeof(float), st>>>(
          y, da, 0, wtop, scale, shift, mean, rstd, sums, rows, C);
  } else {
    DRAM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    if (vec)
      k_bn_reduce<4, 1><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * 4 * sizeof(float), st>>>(
          y, da, da_pitch, nullptr, scale, shift, mean, rstd, sums, rows, C);
    else
      k_bn_reduce<1, 1><<<stat_grid(rows), kStatThreads, 2 * kStatThreads * sizeof(float), st>>>(
          y, da, da_pitch, nullptr, scale, shift, mean, rstd, sums, rows, C);
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_relu_bwd_apply(const float* da, long long da_pitch, const float* y, const float* scale, const float* shift,
                           const float* mean, const float* rstd, const float* gamma, const double* sums, double count,
                           float* dy, long long rows, int C, void* stream) {
  DRAM_REQUIRE(da && y && scale && shift && dy && rows > 0 && C > 0, "bn_relu_bwd_apply: bad arguments");
  DRAM_REQUIRE(!sums || (mean && rstd && count > 0), "bn_relu_bwd_apply: training mode needs mean/rstd/count");
  if (da_pitch == 0) da_pitch = C;
  DRAM_REQUIRE(da_pitch >= C, "bn_relu_bwd_apply: da_pitch %lld < C %d", da_pitch, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 == 0 && da_pitch % 4 == 0)
    k_bn_relu_bwd_apply<4><<<grid_for(rows * (C / 4), 256), 256, 0, st>>>(da, y, scale, shift, mean, rstd, gamma, sums, count, dy, rows, C, da_pitch);
  else
    k_bn_relu_bwd_apply<1><<<grid_for(rows * C, 256), 256, 0, st>>>(da, y, scale, shift, mean, rstd, gamma, sums, count, dy, rows, C, da_pitch);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_maxpool2_bwd(const float* a, const float* dpooled, float* da, int N, int D, int H, int W, int C, void* stream) {
  DRAM_REQUIRE(a && dpooled && da && N > 0 && D >= 2 && H >= 2 && W >= 2 && C > 0, "maxpool2_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  long long cells = (long long)N * (D / 2) * (H / 2) * (W / 2);
  VEC_DISPATCH(C, (k_maxpool2_bwd<4><<<grid_for(cells * (C / 4), 256), 256, 0, st>>>(a, dpooled, da, N, D, H, W, C)),
               (k_maxpool2_bwd<1><<<grid_for(cells * C, 256), 256, 0, st>>>(a, dpooled, da, N, D, H, W, C)));
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static int launch_trilinear_fwd(const float* src, float* dst, int N, int d, int h, int w, int D, int H, int W, int C,
                                int dstC, int dstOff, cudaStream_t st) {
  float sz = ac_scale(d, D), sy = ac_scale(h, H), sx = ac_scale(w, W);
  long long vox = (long long)N * D * H * W;
  if (C % 4 == 0 && dstC % 4 == 0 && dstOff % 4 == 0)
    k_trilinear_fwd<4><<<grid_for(vox * (C / 4), 256), 256, 0, st>>>(src, dst, N, d, h, w, D, H, W, C, dstC, dstOff, sz, sy, sx);
  else
    k_trilinear_fwd<1><<<grid_for(vox * C, 256), 256, 0, st>>>(src, dst, N, d, h, w, D, H, W, C, dstC, dstOff, sz, sy, sx);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}
static int launch_trilinear_bwd(const float* ddst, float* dsrc, int N, int d, int h, int w, int D, int H, int W, int C,
                                int dstC, int dstOff, cudaStream_t st) {
  float sz = ac_scale(d, D), sy = ac_scale(h, H), sx = ac_scale(w, W);
  long long vox = (long long)N * d * h * w;
  if (C % 4 == 0 && dstC % 4 == 0 && dstOff % 4 == 0)
    k_trilinear_bwd<4><<<grid_for(vox * (C / 4), 128), 128, 0, st>>>(ddst, dsrc, N, d, h, w, D, H, W, C, dstC, dstOff, sz, sy, sx);
  else
    k_trilinear_bwd<1><<<grid_for(vox * C, 128), 128, 0, st>>>(ddst, dsrc, N, d, h, w, D, H, W, C, dstC, dstOff, sz, sy, sx);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_trilinear_resize_fwd(const float* src, float* dst, int N, int d, int h, int w, int D, int H, int W, int C,
                              void* stream) {
  DRAM_REQUIRE(src && dst && N > 0 && d > 0 && h > 0 && w > 0 && D > 0 && H > 0 && W > 0 && C > 0, "trilinear_resize_fwd: bad arguments");
  return launch_trilinear_fwd(src, dst, N, d, h, w, D, H, W, C, C, 0, (cudaStream_t)stream);
}
int dram_trilinear_resize_bwd(const float* ddst, float* dsrc, int N, int d, int h, int w, int D, int H, int W, int C,
                              void* stream) {
  DRAM_REQUIRE(ddst && dsrc && N > 0 && d > 0 && h > 0 && w > 0 && D > 0 && H > 0 && W > 0 && C > 0, "trilinear_resize_bwd: bad arguments");
  return launch_trilinear_bwd(ddst, dsrc, N, d, h, w, D, H, W, C, C, 0, (cudaStream_t)stream);
}

static inline int ceil_half(int a) { return (a + 1) / 2; }  // int(np.ceil((b - a) / 2)) for b >= a

int dram_upsample2x_concat_fwd(const float* x, const float* skip, float* cat, int N, int d, int h, int w, int C1,
                               int Ds, int Hs, int Ws, int C2, void* stream) {
  DRAM_REQUIRE(x && skip && cat && N > 0 && d > 0 && h > 0 && w > 0 && C1 > 0 && C2 > 0, "upsample2x_concat_fwd: bad arguments");
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  DRAM_REQUIRE(Ds >= D && Hs >= H && Ws >= W, "upsample2x_concat_fwd: skip (%d,%d,%d) smaller than upsampled (%d,%d,%d)", Ds, Hs, Ws, D, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_trilinear_fwd(x, cat, N, d, h, w, D, H, W, C1, C1 + C2, 0, st);
  if (rc) return rc;
  int oz = ceil_half(Ds - D), oy = ceil_half(Hs - H), ox = ceil_half(Ws - W);
  long long vox = (long long)N * D * H * W;
  if (C1 % 4 == 0 && C2 % 4 == 0)
    k_concat_skip<4, false><<<grid_for(vox * (C2 / 4), 256), 256, 0, st>>>(skip, cat, N, D, H, W, Ds, Hs, Ws, oz, oy, ox, C1, C2);
  else
    k_concat_skip<1, false><<<grid_for(vox * C2, 256), 256, 0, st>>>(skip, cat, N, D, H, W, Ds, Hs, Ws, oz, oy, ox, C1, C2);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_upsample2x_concat_bwd(const float* dcat, float* dx, float* dskip, int N, int d, int h, int w, int C1, int Ds,
                               int Hs, int Ws, int C2, void* stream) {
  DRAM_REQUIRE(dcat && dx && N > 0 && d > 0 && h > 0 && w > 0 && C1 > 0 && C2 > 0, "upsample2x_concat_bwd: bad arguments");
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  DRAM_REQUIRE(Ds >= D && Hs >= H && Ws >= W, "upsample2x_concat_bwd: bad skip size");
  cudaStream_t st = (cudaStream_t)stream;
  if (C1 % 4 == 0 && C2 % 4 == 0 && (long long)N * d * h * w < (1ll << 31)) {
    const bool tiled_ok = getenv("DRAM_UP2X_ADJ_BLOCKED") == nullptr;
    const long long tiles = (long long)N * ((h + kAdjT - 1) / kAdjT) * ((w + kAdjT - 1) / kAdjT) * (C1 / kAdjC);
    if (tiled_ok && C1 % kAdjC == 0 && d >= 2 && tiles > 0) {
      // z segments: a few waves of 2 blocks per SM, at least 4 source planes per block (each segment re-reads ~4 halo planes)
      int zsegs = (int)((4ll * 2 * kNumSMs + tiles - 1) / tiles);
      if (zsegs > d / 4) zsegs = d / 4;
      if (zsegs < 1) zsegs = 1;
      const size_t smem = (2 * kAdjTileFloats + kAdjP * kAdjT * kAdjC + 2 * kAdjT * 4) * sizeof(float);
      static std::once_flag once;
      std::call_once(once, [] { cudaFuncSetAttribute(k_up2x_adjoint_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); });
      k_up2x_adjoint_tiled<<<(unsigned)(tiles * zsegs), 256, smem, st>>>(dcat, dx, N, d, h, w, C1, C1 + C2, 0, ac_scale(d, D),
                                                                        ac_scale(h, H), ac_scale(w, W), zsegs);
    } else {
      const long long blocks = (long long)N * ((d + 1) / 2) * ((h + 1) / 2) * ((w + 1) / 2);
      k_up2x_adjoint<<<grid_fixed_group(blocks, C1 / 4, 128, 64), 128, 0, st>>>(dcat, dx, N, d, h, w, C1, C1 + C2, 0, ac_scale(d, D),
                                                                             ac_scale(h, H), ac_scale(w, W));
    }
    DRAM_LAUNCH_CHECK();
  } else {
    int rc = launch_trilinear_bwd(dcat, dx, N, d, h, w, D, H, W, C1, C1 + C2, 0, st);
    if (rc) return rc;
  }
  if (!dskip) return DRAM_OK;            // the caller reads dcat[..., C1:] in place (no crop: Ds == D etc.)
  int oz = ceil_half(Ds - D), oy = ceil_half(Hs - H), ox = ceil_half(Ws - W);
  long long vox = (long long)N * Ds * Hs * Ws;
  if (C1 % 4 == 0 && C2 % 4 == 0)
    k_concat_skip<4, true><<<grid_for(vox * (C2 / 4), 256), 256, 0, st>>>(dcat, dskip, N, D, H, W, Ds, Hs, Ws, oz, oy, ox, C1, C2);
  else
    k_concat_skip<1, true><<<grid_for(vox * C2, 256), 256, 0, st>>>(dcat, dskip, N, D, H, W, Ds, Hs, Ws, oz, oy, ox, C1, C2);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
