// CUDA-core fp32 implicit-GEMM convolution (k in {1,3}, pad k/2, stride 1) on channels-last volumes.
// Role on the path: the bandwidth-bound first layer (Cin = 1, K = 27), the 1x1x1 heads (models.py:109,490) and any
// channel count the tcgen05 kernel does not cover.  GEMM view: M = N*D*H*W voxels, N = Cout, K = taps*Cin with the K
// index kk = tap*Cin + ci flattened so that tiny Cin does not waste the K tile.
#include "common.cuh"

namespace dram {

constexpr int BM = 64, BN = 64, BK = 16, APAD = 4;

struct VoxelCoord {
  int n, z, y, x;
  bool ok;
};
__device__ __forceinline__ VoxelCoord decode(long long m, long long M, int D, int H, int W) {
  VoxelCoord v;
  v.ok = m < M;
  long long r = v.ok ? m : 0;
  v.x = (int)(r % W); r /= W;
  v.y = (int)(r % H); r /= H;
  v.z = (int)(r % D);
  v.n = (int)(r / D);
  return v;
}

// value of the im2col matrix A[m][kk]
template <int KS>
__device__ __forceinline__ const float* a_ptr(const float* x, const VoxelCoord& v, int tap, int ci, int D, int H, int W,
                                              int Cin) {
  int z = v.z, y = v.y, xx = v.x;
  if (KS == 3) {
    z += tap / 9 - 1; y += (tap / 3) % 3 - 1; xx += tap % 3 - 1;
    if (z < 0 || z >= D || y < 0 || y >= H || xx < 0 || xx >= W) return nullptr;
  }
  return x + ((((long long)v.n * D + z) * H + y) * W + xx) * Cin + ci;
}

template <int KS>
__global__ void __launch_bounds__(256)
k_conv_simt_fwd(const float* __restrict__ x, const float* __restrict__ pack, const float* __restrict__ bias,
                float* __restrict__ y, int N, int D, int H, int W, int Cin, int Cout) {
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const long long M = (long long)N * D * H * W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  const int Ktot = KS * KS * KS * Cin;
  const bool vecA = (Cin % 4 == 0), vecB = (Cout % 4 == 0);

  const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;
  const VoxelCoord av = decode(m0 + a_row, M, D, H, W);
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;
  const int tm = tid >> 4, tn = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    // ---- A tile (im2col gather)
    float va[4] = {0.f, 0.f, 0.f, 0.f};
    int kk = k0 + a_k4;
    if (av.ok && kk < Ktot) {
      if (vecA) {
        const float* p = a_ptr<KS>(x, av, kk / Cin, kk % Cin, D, H, W, Cin);
        if (p) {
          float4 t = *reinterpret_cast<const float4*>(p);
          va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (kk + j < Ktot) {
            const float* p = a_ptr<KS>(x, av, (kk + j) / Cin, (kk + j) % Cin, D, H, W, Cin);
            if (p) va[j] = *p;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) As[a_k4 + j][a_row] = va[j];
    // ---- B tile
    float vb[4] = {0.f, 0.f, 0.f, 0.f};
    int bk = k0 + b_k, bc = co0 + b_n4;
    if (bk < Ktot) {
      const float* p = pack + (long long)bk * Cout + bc;
      if (vecB && bc + 3 < Cout) {
        float4 t = *reinterpret_cast<const float4*>(p);
        vb[0] = t.x; vb[1] = t.y; vb[2] = t.z; vb[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (bc + j < Cout) vb[j] = p[j];
      }
    }
    *reinterpret_cast<float4*>(&Bs[b_k][b_n4]) = make_float4(vb[0], vb[1], vb[2], vb[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int c = co0 + tn * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + tm * 4 + i;
    if (m >= M) continue;
    float* q = y + m * Cout + c;
    if (vecB && c + 3 < Cout) {
      float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (bias) { o.x += bias[c]; o.y += bias[c + 1]; o.z += bias[c + 2]; o.w += bias[c + 3]; }
      *reinterpret_cast<float4*>(q) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < Cout) q[j] = acc[i][j] + (bias ? bias[c + j] : 0.f);
    }
  }
}

// dpack[kk][co] += sum_{m in slab} A[m][kk] * dy[m][co]
constexpr int WK = 64, WN = 64, WM = 16, kWgradSlab = 4096;

template <int KS>
__global__ void __launch_bounds__(256)
k_conv_simt_wgrad(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dpack, int N, int D,
                  int H, int W, int Cin, int Cout) {
  __shared__ __align__(16) float As[WM][WK + APAD];
  __shared__ __align__(16) float Bs[WM][WN];
  const int tid = threadIdx.x;
  const long long M = (long long)N * D * H * W;
  const int Ktot = KS * KS * KS * Cin;
  const int kk0 = blockIdx.x * WK, co0 = blockIdx.y * WN;
  const bool vecA = (Cin % 4 == 0), vecB = (Cout % 4 == 0);
  const int l_m = tid >> 4, l_4 = (tid & 15) * 4;
  const int tk = tid >> 4, tn = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long slab = (long long)blockIdx.z * kWgradSlab; slab < M; slab += (long long)gridDim.z * kWgradSlab) {
    long long end = slab + kWgradSlab < M ? slab + kWgradSlab : M;
    for (long long mb = slab; mb < end; mb += WM) {
      long long m = mb + l_m;
      VoxelCoord v = decode(m, end, D, H, W);
      float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
      int kk = kk0 + l_4;
      if (v.ok && kk < Ktot) {
        if (vecA) {
          const float* p = a_ptr<KS>(x, v, kk / Cin, kk % Cin, D, H, W, Cin);
          if (p) { float4 t = *reinterpret_cast<const float4*>(p); va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w; }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kk + j < Ktot) {
              const float* p = a_ptr<KS>(x, v, (kk + j) / Cin, (kk + j) % Cin, D, H, W, Cin);
              if (p) va[j] = *p;
            }
        }
      }
      int bc = co0 + l_4;
      if (v.ok) {
        const float* p = dy + m * Cout + bc;
        if (vecB && bc + 3 < Cout) { float4 t = *reinterpret_cast<const float4*>(p); vb[0] = t.x; vb[1] = t.y; vb[2] = t.z; vb[3] = t.w; }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (bc + j < Cout) vb[j] = p[j];
        }
      }
      *reinterpret_cast<float4*>(&As[l_m][l_4]) = make_float4(va[0], va[1], va[2], va[3]);
      *reinterpret_cast<float4*>(&Bs[l_m][l_4]) = make_float4(vb[0], vb[1], vb[2], vb[3]);
      __syncthreads();
#pragma unroll
      for (int mi = 0; mi < WM; ++mi) {
        float4 a = *reinterpret_cast<const float4*>(&As[mi][tk * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[mi][tn * 4]);
        float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = kk0 + tk * 4 + i;
    if (kk >= Ktot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = co0 + tn * 4 + j;
      if (c < Cout) atomicAdd(&dpack[(long long)kk * Cout + c], acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ first layer (Cin = 1)
// ds0.c0 (1 -> 32 channels, k = 3): K = 27, 13 FLOP/B — bandwidth-bound on the write of y (forward) / the read of dy
// (wgrad), so no GEMM tiling: every thread owns a few output channels of one voxel and keeps the 27 taps in registers.
// The launch guarantees (gridDim.x * blockDim.x) % groups == 0: a thread keeps its channel group.
constexpr int kC1MaxCout = 64, kC1MaxCoutWgrad = 32;

__device__ __forceinline__ void c1_taps(const float* __restrict__ x, unsigned m, int D, int H, int W, float (&xv)[27]) {
  unsigned r = m;
  const int xx = (int)(r % (unsigned)W); r /= (unsigned)W;
  const int yy = (int)(r % (unsigned)H); r /= (unsigned)H;
  const int zz = (int)(r % (unsigned)D);
  const long long base = (long long)m;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const int dz = t / 9 - 1, dy = (t / 3) % 3 - 1, dx = t % 3 - 1;
    const bool ok = (unsigned)(zz + dz) < (unsigned)D && (unsigned)(yy + dy) < (unsigned)H && (unsigned)(xx + dx) < (unsigned)W;
    xv[t] = ok ? __ldg(x + base + ((long long)dz * H + dy) * W + dx) : 0.f;
  }
}

// thread = (voxel, 8 output channels); y [M][Cout]
__global__ void __launch_bounds__(256)
k_conv_c1_fwd(const float* __restrict__ x, const float* __restrict__ pack, const float* __restrict__ bias,
              float* __restrict__ y, long long M, int D, int H, int W, int Cout) {
  __shared__ __align__(16) float ws[27 * kC1MaxCout];
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) ws[i] = pack[i];
  __syncthreads();
  const int groups = Cout / 8;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dm = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);
  float b8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b8[j] = bias ? bias[g * 8 + j] : 0.f;
  for (long long m = tid / groups; m < M; m += dm) {
    float xv[27], acc[8];
    c1_taps(x, (unsigned)m, D, H, W, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = b8[j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[t * Cout + g * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[t * Cout + g * 8 + 4]);
      acc[0] = fmaf(xv[t], w0.x, acc[0]); acc[1] = fmaf(xv[t], w0.y, acc[1]);
      acc[2] = fmaf(xv[t], w0.z, acc[2]); acc[3] = fmaf(xv[t], w0.w, acc[3]);
      acc[4] = fmaf(xv[t], w1.x, acc[4]); acc[5] = fmaf(xv[t], w1.y, acc[5]);
      acc[6] = fmaf(xv[t], w1.z, acc[6]); acc[7] = fmaf(xv[t], w1.w, acc[7]);
    }
    float4* q = reinterpret_cast<float4*>(y + m * Cout + g * 8);
    q[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    q[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// thread = (voxel lane, 4 output channels) accumulating dpack[27][4] over its voxels; lanes of a warp that share the
// channel group are folded with shuffles, warps through shared memory, blocks with one atomicAdd per output.
// groups = Cout / 4 must be a power of two <= 32.
__global__ void __launch_bounds__(256)
k_conv_c1_wgrad(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dpack, long long M, int D,
                int H, int W, int Cout) {
  __shared__ float red[8][27 * kC1MaxCoutWgrad];
  const int groups = Cout / 4;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dm = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);             // == lane % groups (256 and 32 are multiples of groups)
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  for (long long m = tid / groups; m < M; m += dm) {
    float xv[27];
    c1_taps(x, (unsigned)m, D, H, W, xv);
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(dy + m * Cout) + g);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      acc[t][0] = fmaf(xv[t], d4.x, acc[t][0]); acc[t][1] = fmaf(xv[t], d4.y, acc[t][1]);
      acc[t][2] = fmaf(xv[t], d4.z, acc[t][2]); acc[t][3] = fmaf(xv[t], d4.w, acc[t][3]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[t][j];
      for (int o = 16; o >= groups; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < groups) red[warp][t * Cout + lane * 4 + j] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) v += red[wp][i];
    atomicAdd(&dpack[i], v);
  }
}

// ------------------------------------------------------------------------------------------------ first layer, 4 voxels per thread
// W % 4 == 0: a thread owns FOUR consecutive-w voxels and one channel group.  The 3x3x6 input window of the quad is 27
// loads (one aligned float4 + two edge scalars per (dz,dy)) instead of 4 x 27 gathers, and the weights / accumulators are
// touched once per quad.
__device__ __forceinline__ void c1_window(const float* __restrict__ x, unsigned quad, int D, int H, int W4, float (&xv)[9][6],
                                          long long& m0) {
  unsigned r = quad;
  const int x0 = (int)(r % (unsigned)W4) * 4; r /= (unsigned)W4;
  const int yy = (int)(r % (unsigned)H); r /= (unsigned)H;
  const int zz = (int)(r % (unsigned)D);
  const int W = W4 * 4;
  m0 = (long long)quad * 4;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int dz = t / 3 - 1, dy = t % 3 - 1;
    const bool ok = (unsigned)(zz + dz) < (unsigned)D && (unsigned)(yy + dy) < (unsigned)H;
    const float* row = x + m0 + ((long long)dz * H + dy) * W;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f, rr = 0.f;
    if (ok) {
      c = __ldg(reinterpret_cast<const float4*>(row));
      if (x0 > 0) l = __ldg(row - 1);
      if (x0 + 4 < W) rr = __ldg(row + 4);
    }
    xv[t][0] = l; xv[t][1] = c.x; xv[t][2] = c.y; xv[t][3] = c.z; xv[t][4] = c.w; xv[t][5] = rr;
  }
}

// thread = (quad of voxels, 8 output channels)
__global__ void __launch_bounds__(256)
k_conv_c1_fwd_q(const float* __restrict__ x, const float* __restrict__ pack, const float* __restrict__ bias,
                float* __restrict__ y, long long quads, int D, int H, int W4, int Cout) {
  __shared__ __align__(16) float ws[27 * kC1MaxCout];
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) ws[i] = pack[i];
  __syncthreads();
  const int groups = Cout / 8;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dq = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);
  float b8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b8[j] = bias ? bias[g * 8 + j] : 0.f;
  for (long long qd = tid / groups; qd < quads; qd += dq) {
    float xv[9][6];
    float2 acc[4][4];                                   // packed fp32 pairs (fma.rn.f32x2, sm_100): half the FMA issue slots,
    long long m0;                                       // the same roundings as 864 scalar FMAs per quad
    c1_window(x, (unsigned)qd, D, H, W4, xv, m0);
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(b8[2 * j], b8[2 * j + 1]);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[t * Cout + g * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[t * Cout + g * 8 + 4]);
      const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w), wc = make_float2(w1.x, w1.y), wd = make_float2(w1.z, w1.w);
      const int r9 = t / 3, dx = t % 3;                 // tap t = (dz,dy) row r9, column dx: voxel v reads xv[r9][v + dx]
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float2 xs = make_float2(xv[r9][v + dx], xv[r9][v + dx]);
        acc[v][0] = __ffma2_rn(xs, wa, acc[v][0]); acc[v][1] = __ffma2_rn(xs, wb, acc[v][1]);
        acc[v][2] = __ffma2_rn(xs, wc, acc[v][2]); acc[v][3] = __ffma2_rn(xs, wd, acc[v][3]);
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float4* o = reinterpret_cast<float4*>(y + (m0 + v) * Cout + g * 8);
      o[0] = make_float4(acc[v][0].x, acc[v][0].y, acc[v][1].x, acc[v][1].y);
      o[1] = make_float4(acc[v][2].x, acc[v][2].y, acc[v][3].x, acc[v][3].y);
    }
  }
}

// thread = (quad lane, 4 output channels) accumulating dpack[27][4] over its quads; same reduction tree as k_conv_c1_wgrad
__global__ void __launch_bounds__(256)
k_conv_c1_wgrad_q(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dpack, long long quads, int D,
                  int H, int W4, int Cout) {
  __shared__ float red[8][27 * kC1MaxCoutWgrad];
  const int groups = Cout / 4;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long dq = ((long long)gridDim.x * blockDim.x) / groups;
  const int g = (int)(tid % groups);
  float2 acc2[27][2];                                   // packed fp32 pairs (fma.rn.f32x2): same roundings, half the issue slots
#pragma unroll
  for (int t = 0; t < 27; ++t) acc2[t][0] = acc2[t][1] = make_float2(0.f, 0.f);
  for (long long qd = tid / groups; qd < quads; qd += dq) {
    float xv[9][6];
    long long m0;
    c1_window(x, (unsigned)qd, D, H, W4, xv, m0);
    float4 d4[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) d4[v] = __ldg(reinterpret_cast<const float4*>(dy + (m0 + v) * Cout) + g);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int r9 = t / 3, dx = t % 3;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float2 xs = make_float2(xv[r9][v + dx], xv[r9][v + dx]);
        acc2[t][0] = __ffma2_rn(xs, make_float2(d4[v].x, d4[v].y), acc2[t][0]);
        acc2[t][1] = __ffma2_rn(xs, make_float2(d4[v].z, d4[v].w), acc2[t][1]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = (j & 1) ? acc2[t][j >> 1].y : acc2[t][j >> 1].x;
      for (int o = 16; o >= groups; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < groups) red[warp][t * Cout + lane * 4 + j] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) v += red[wp][i];
    atomicAdd(&dpack[i], v);
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// w [Cout][Cin][T]  ->  mode 0: pack[t][ci][co]   mode 1: pack[t][co][ci] with flipped taps
__global__ void k_pack_weight_f32(const float* __restrict__ w, float* __restrict__ pack, int Cout, int Cin, int T, int mode) {
  long long total = (long long)Cout * Cin * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      int co = (int)(i % Cout); long long r = i / Cout;
      int ci = (int)(r % Cin); int t = (int)(r / Cin);
      pack[i] = w[((long long)co * Cin + ci) * T + t];
    } else {
      int ci = (int)(i % Cin); long long r = i / Cin;
      int co = (int)(r % Cout); int t = (int)(r / Cout);
      pack[i] = w[((long long)co * Cin + ci) * T + (T - 1 - t)];
    }
  }
}
__global__ void k_unpack_wgrad_f32(const float* __restrict__ pack, float* __restrict__ dw, int Cout, int Cin, int T) {
  long long total = (long long)Cout * Cin * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % T); long long r = i / T;
    int ci = (int)(r % Cin); int co = (int)(r / Cin);
    dw[i] = pack[((long long)t * Cin + ci) * Cout + co];
  }
}

}  // namespace dram

using namespace dram;

extern "C" {

int dram_pack_weight_f32(const float* w, float* pack, int Cout, int Cin, int ksize, int mode, void* stream) {
  DRAM_REQUIRE(w && pack && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3) && (mode == 0 || mode == 1), "pack_weight_f32: bad arguments");
  int T = ksize * ksize * ksize;
  k_pack_weight_f32<<<grid_for((long long)Cout * Cin * T, 256), 256, 0, (cudaStream_t)stream>>>(w, pack, Cout, Cin, T, mode);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_unpack_wgrad_f32(const float* pack, float* dw, int Cout, int Cin, int ksize, void* stream) {
  DRAM_REQUIRE(pack && dw && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "unpack_wgrad_f32: bad arguments");
  int T = ksize * ksize * ksize;
  k_unpack_wgrad_f32<<<grid_for((long long)Cout * Cin * T, 256), 256, 0, (cudaStream_t)stream>>>(pack, dw, Cout, Cin, T);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_conv3d_simt_fwd(const float* x, const float* pack, const float* bias, float* y, int N, int D, int H, int W,
                         int Cin, int Cout, int ksize, void* stream) {
  DRAM_REQUIRE(x && pack && y && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_simt_fwd: bad arguments");
  DRAM_REQUIRE(ksize == 1 || ksize == 3, "conv3d_simt_fwd: kernel size %d unsupported (1 or 3)", ksize);
  long long M = (long long)N * D * H * W;
  if (Cin == 1 && ksize == 3 && Cout % 8 == 0 && Cout <= kC1MaxCout && M < (1ll << 31)) {
    if (W % 4 == 0 && ((uintptr_t)x & 15) == 0)
      k_conv_c1_fwd_q<<<grid_fixed_group(M / 4, Cout / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, pack, bias, y, M / 4, D, H, W / 4, Cout);
    else
      k_conv_c1_fwd<<<grid_fixed_group(M, Cout / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, pack, bias, y, M, D, H, W, Cout);
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN));
  if (ksize == 3) k_conv_simt_fwd<3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pack, bias, y, N, D, H, W, Cin, Cout);
  else k_conv_simt_fwd<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pack, bias, y, N, D, H, W, Cin, Cout);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_conv3d_simt_wgrad(const float* x, const float* dy, float* dpack, int N, int D, int H, int W, int Cin, int Cout,
                           int ksize, void* stream) {
  DRAM_REQUIRE(x && dy && dpack && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_simt_wgrad: bad arguments");
  DRAM_REQUIRE(ksize == 1 || ksize == 3, "conv3d_simt_wgrad: kernel size %d unsupported (1 or 3)", ksize);
  long long M = (long long)N * D * H * W;
  const int c1g = Cout / 4;
  if (Cin == 1 && ksize == 3 && Cout % 4 == 0 && Cout <= kC1MaxCoutWgrad && c1g <= 32 && (c1g & (c1g - 1)) == 0 && M < (1ll << 31)) {
    if (W % 4 == 0 && ((uintptr_t)x & 15) == 0)
      k_conv_c1_wgrad_q<<<grid_fixed_group(M / 4, c1g, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, dy, dpack, M / 4, D, H, W / 4, Cout);
    else
      k_conv_c1_wgrad<<<grid_fixed_group(M, c1g, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, dy, dpack, M, D, H, W, Cout);
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  int Ktot = ksize * ksize * ksize * Cin;
  long long slabs = (M + kWgradSlab - 1) / kWgradSlab;
  long long tiles = (long long)((Ktot + WK - 1) / WK) * ((Cout + WN - 1) / WN);
  long long want = ((long long)kNumSMs * 8 + tiles - 1) / tiles;       // >= 8 waves worth of CTAs
  unsigned gz = (unsigned)(slabs < want ? slabs : want);
  if (gz < 1) gz = 1;
  dim3 grid((unsigned)((Ktot + WK - 1) / WK), (unsigned)((Cout + WN - 1) / WN), gz);
  if (ksize == 3) k_conv_simt_wgrad<3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, dpack, N, D, H, W, Cin, Cout);
  else k_conv_simt_wgrad<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, dpack, N, D, H, W, Cin, Cout);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
