// PCM graph attention (models.py:150-411) as a stencil attention on the voxel grid.
//
// The reference builds a DGL graph with one node per voxel of `at_spatial_size` and an edge from every in-grid
// neighbour of the (connectivity, k_size=3) structuring element (models.py:223-259); `update_all` gathers the
// neighbours' features into a mailbox and reduce_func (models.py:355-389) computes, per node x,
//     s_o = act(<theta f_x, phi f_{x+o}>) / T(x),   a = softmax_o(s),   y = sum_o a_o G(cam_{x+o}),   out = r(y).
// G and r are affine in the scalar cam and sum_o a_o = 1, so out = (w_r . w_G) * sum_o a_o cam_{x+o} + w_r.b_G + b_r;
// this file computes sum_o a_o cam_{x+o} (the affine is two scalars, applied by the host wrapper).
// T(x) = sqrt(#in-grid neighbours of x) for the 'scaled_dot_product*' merge types (models.py:274,277: the mailbox
// degree, not the feature dim).  No graph structure is materialised: neighbours are address arithmetic.
//
// Data (R = B*V rows, V = D*H*W): f [R][Cf] channels-last, cam/out [R];
//   qk    [ceil(R/32)][2F][32]  theta|phi projections in blocks of 32 consecutive voxels (array of structures of arrays):
//                  a warp = 32 consecutive x of one grid row, so one feature of one neighbour offset is one coalesced
//                  128-byte request, and the 2F features of a voxel sit at COMPILE-TIME offsets (j*128 bytes) from one
//                  address (the [R][2F] rows of the first version cost 16 L1 wavefronts per request, plain planes a
//                  64-bit address computation per load);
//   stats [R][4]   (m, 1/l, 1/T, s): running max and inverse normaliser of the node's softmax, its inverse temperature
//                  and its output - what the backward keeps INSTEAD of the [R][O] attention weights:
//                  a_{x,o} = exp(s_{x,o}/T_x - m_x) / l_x is recomputed from q_x and k_{x+o} (F FMAs + 1 exp).
#include "common.cuh"

namespace dram {

constexpr int kMaxOff = 27, kMaxF = 16, kMaxCf = 64;

struct PcmGeom {
  int B, D, H, W, Cf, F, O, flags;     // flags: bit0 relu, bits1-2 temperature mode, bit7 self loop, bits8+ connectivity
};

static int count_offsets(int connectivity, int self_loop) {
  int n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        int l1 = abs(dz) + abs(dy) + abs(dx);
        n += (l1 == 0) ? (self_loop != 0) : (l1 <= connectivity);
      }
  return n;
}

// 1/T; mode 1 (scaled_dot_product*): T = sqrt(degree) -> one MUFU.RSQ (the forward and the backward use the same value)
__device__ __forceinline__ float inv_temperature(int flags, int deg) {
  const int mode = (flags >> 1) & 3;
  return mode == 1 ? rsqrtf((float)deg) : (mode == 2 ? 100.f : 1.f);
}

// number of in-grid neighbours in closed form: with a_z = #{dz in {-1,+1} : z+dz in grid} etc. the stencil
// {|dz|+|dy|+|dx| <= connectivity} has a_z+a_y+a_x members at L1 distance 1, a_z a_y + a_z a_x + a_y a_x at 2, a_z a_y a_x at 3
__device__ __forceinline__ int degree_from(int flags, int az, int ay, int ax) {
  const int conn = flags >> 8;
  int deg = ((flags >> 7) & 1) + az + ay + ax;
  if (conn >= 2) deg += az * ay + az * ax + ay * ax;
  if (conn >= 3) deg += az * ay * ax;
  return deg;
}

// The stencil as 27 fully unrolled candidates in (dz, dy, dx) ascending order: dz/dy/dx are compile-time constants inside
// the body, membership is a warp-uniform test (indexing a by-value offset table would put it in local memory).
#define PCM_FOR_EACH_OFFSET(flags_)                                                                                 \
  _Pragma("unroll") for (int t_ = 0; t_ < 27; ++t_)                                                                 \
    if (const int dz = t_ / 9 - 1, dy = (t_ / 3) % 3 - 1, dx = t_ % 3 - 1, l1_ = (dz != 0) + (dy != 0) + (dx != 0); \
        l1_ == 0 ? (((flags_) >> 7) & 1) != 0 : l1_ <= ((flags_) >> 8))

// per-thread position: one z-plane per blockIdx.y (b, z uniform), one 32-bit division for (y, x)
struct PcmPos {
  int i, x, y, z;               // i: global row index
  bool zm, zp, ym, yp, xm, xp;  // neighbour at -1 / +1 along the axis is inside the grid
  bool live;
  __device__ __forceinline__ PcmPos(const PcmGeom& g) {
    const int HW = g.H * g.W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    live = p < HW;
    const int pp = live ? p : HW - 1;
    z = blockIdx.y % g.D;
    y = pp / g.W;
    x = pp - y * g.W;
    i = blockIdx.y * HW + pp;
    zm = z > 0; zp = z < g.D - 1; ym = y > 0; yp = y < g.H - 1; xm = x > 0; xp = x < g.W - 1;
  }
  // neighbour (sz*dz, sz*dy, sz*dx), sz = +1 (x + o) or -1 (x - o), inside the grid?
  template <int S>
  __device__ __forceinline__ bool inside(int dz, int dy, int dx) const {
    const int ez = S * dz, ey = S * dy, ex = S * dx;
    return (ez < 0 ? zm : (ez > 0 ? zp : true)) && (ey < 0 ? ym : (ey > 0 ? yp : true)) && (ex < 0 ? xm : (ex > 0 ? xp : true));
  }
};

template <int F>
__device__ __forceinline__ const float* qk_at(const float* qk, int n) { return qk + (size_t)(n >> 5) * (2 * F * 32) + (n & 31); }

// qk block layout: feature j (theta: j < F, phi: F <= j < 2F) of row r at [(r >> 5)][j][r & 31].  A warp stages its 32
// consecutive rows (32*Cf contiguous floats) through shared memory with coalesced loads; a lane then walks its row at an
// odd pitch (conflict-free); the weights of one input channel are J4 float4 broadcast reads.
__global__ void __launch_bounds__(256)
k_pcm_project(const float* __restrict__ f, const float* __restrict__ tw, const float* __restrict__ tb,
              const float* __restrict__ pw, const float* __restrict__ pb, float* __restrict__ qk, long long rows, int Cf, int F) {
  extern __shared__ float4 psm4[];
  float* psm = reinterpret_cast<float*>(psm4);
  const int Cfp = Cf | 1, J4 = (2 * F + 3) >> 2;
  float* sw = psm;                                  // [Cf][4*J4]  (transposed, zero padded)
  float* sb = sw + Cf * 4 * J4;                     // [4*J4]
  float* st = sb + 4 * J4 + (threadIdx.x >> 5) * 32 * Cfp;   // this warp's [32][Cfp] tile
  for (int i = threadIdx.x; i < Cf * 4 * J4; i += blockDim.x) {
    const int c = i / (4 * J4), j = i - c * 4 * J4;
    sw[i] = j < F ? tw[j * Cf + c] : (j < 2 * F ? pw[(j - F) * Cf + c] : 0.f);
  }
  for (int i = threadIdx.x; i < 4 * J4; i += blockDim.x) sb[i] = i < F ? tb[i] : (i < 2 * F ? pb[i - F] : 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned inv_cf = (1u << 20) / (unsigned)Cf + 1u;          // i / Cf == (i * inv_cf) >> 20 for i < 32 * Cf
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5, tiles = (rows + 31) >> 5;
  for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < tiles; t += warps) {
    const long long r0 = t << 5;
    const int n = (int)(rows - r0 < 32 ? rows - r0 : 32);
    const float* src = f + r0 * Cf;
    __syncwarp();
    for (int i0 = lane; i0 < n * Cf; i0 += 32 * 6) {                 // 6 independent loads in flight, then their stores
      float v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) v[u] = i0 + 32 * u < n * Cf ? __ldg(src + i0 + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + 32 * u;
        if (i < n * Cf) { const int r = (int)(((unsigned)i * inv_cf) >> 20); st[r * Cfp + (i - r * Cf)] = v[u]; }
      }
    }
    __syncwarp();
    if (lane < n) {
      float4 acc[kMaxF / 2];
#pragma unroll
      for (int j = 0; j < kMaxF / 2; ++j) acc[j] = j < J4 ? reinterpret_cast<const float4*>(sb)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float* fp = st + lane * Cfp;
      for (int c = 0; c < Cf; ++c) {
        const float v = fp[c];
        const float4* wc = reinterpret_cast<const float4*>(sw) + c * J4;
#pragma unroll
        for (int j = 0; j < kMaxF / 2; ++j)
          if (j < J4) {
            const float4 w4 = wc[j];
            acc[j].x = fmaf(v, w4.x, acc[j].x); acc[j].y = fmaf(v, w4.y, acc[j].y);
            acc[j].z = fmaf(v, w4.z, acc[j].z); acc[j].w = fmaf(v, w4.w, acc[j].w);
          }
      }
      float* dst = qk + t * (2 * F * 32) + lane;
#pragma unroll
      for (int j = 0; j < kMaxF / 2; ++j) {
        if (4 * j + 0 < 2 * F) dst[(4 * j + 0) * 32] = acc[j].x;
        if (4 * j + 1 < 2 * F) dst[(4 * j + 1) * 32] = acc[j].y;
        if (4 * j + 2 < 2 * F) dst[(4 * j + 2) * 32] = acc[j].z;
        if (4 * j + 3 < 2 * F) dst[(4 * j + 3) * 32] = acc[j].w;
      }
    }
  }
}

// forward: one thread per node; scores of the <= 26(+1) neighbours in registers, two-pass softmax.  SAVE: keep stats.
template <int F, bool SAVE>
__global__ void __launch_bounds__(256, F <= 8 ? 3 : 2)
k_pcm_attend(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, float4* __restrict__ stats,
             float* __restrict__ out) {
  const PcmPos P(g);
  const int W = g.W, HW = g.H * g.W, flags = g.flags;
  const bool relu = flags & 1;
  const float* qp = qk_at<F>(qk, P.i);
  float q[F];
#pragma unroll
  for (int j = 0; j < F; ++j) q[j] = __ldg(qp + j * 32);
  const int deg = degree_from(flags, P.zm + P.zp, P.ym + P.yp, P.xm + P.xp);
  const float invT = inv_temperature(flags, deg);
  float s[27], c[27];
  float m = -INFINITY;
  PCM_FOR_EACH_OFFSET(flags) {
    float d = -INFINITY, cv = 0.f;
    if (P.inside<1>(dz, dy, dx)) {
      const int n = P.i + dz * HW + dy * W + dx;
      const float* kp = qk_at<F>(qk, n) + F * 32;
      d = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) d = fmaf(q[j], __ldg(kp + j * 32), d);
      if (relu) d = fmaxf(d, 0.f);
      d *= invT;
      cv = __ldg(cam + n);
    }
    s[t_] = d; c[t_] = cv;
    m = fmaxf(m, d);
  }
  float l = 0.f, acc = 0.f;
  PCM_FOR_EACH_OFFSET(flags) {
    const float e = __expf(s[t_] - m);                 // absent neighbour: exp(-inf) = 0
    l += e;
    acc = fmaf(e, c[t_], acc);
  }
  if (!P.live) return;
  const bool any = deg > 0;                             // degree-0 nodes (1x1x1 grids) keep 0 like DGL's zero-filled result
  const float il = any ? 1.f / l : 0.f, o = any ? acc * il : 0.f;
  out[P.i] = o;
  if (SAVE) stats[P.i] = make_float4(any ? m : 0.f, il, invT, o);
}

// ------------------------------------------------------------------------------------------------ fused forward (inference)
// Projection + attention in ONE kernel for the no-grad forward (process_pipeline / evaluate_scan): nothing is kept for a
// backward, so the theta|phi projections never go to global memory (k_pcm_project writes 64 B/voxel that k_pcm_attend reads
// back; the two launches together moved 204 B/voxel for 76 algorithmic and were latency-bound on global gathers).
// A block owns a (kPfYT x kPfXT) column of the grid and marches along z:
//   * per new z-plane every thread projects ITS voxel (q for itself in registers, k and cam into shared memory) and the
//     first 140 threads project the halo ring (k only): planes z-1, z, z+1 live in a 3-slot ring,
//     kp[slot][F/4][halo voxel] as float4 (a warp = 32 consecutive x reads 512 contiguous bytes: conflict-free);
//   * the 6/18/26(+1) neighbour dot products then read shared memory; all multiply-adds are packed fp32 pairs
//     (fma.rn.f32x2, sm_100) — half the FMA issue slots of the scalar kernels;
//   * same stencil macro, degree temperature and two-pass softmax as k_pcm_attend.
constexpr int kPfYT = 4, kPfXT = 64, kPfHX = kPfXT + 2, kPfHV = (kPfYT + 2) * kPfHX, kPfRing = kPfHV - kPfYT * kPfXT;
constexpr int kPfBatch = 9;                       // independent feature loads in flight per voxel (Cf = 17: two batches)

template <int F>
__device__ __forceinline__ void pcm_project_voxel(const float* __restrict__ fp, int Cf, const float4* __restrict__ sw,
                                                  const float4* __restrict__ sb, bool want_q, float2 (&q)[F / 2],
                                                  float2 (&k)[F / 2]) {
  constexpr int Q4 = F / 4;                       // float4 per projection
#pragma unroll
  for (int j = 0; j < Q4; ++j) {
    const float4 bq = sb[j], bk = sb[Q4 + j];
    q[2 * j] = make_float2(bq.x, bq.y); q[2 * j + 1] = make_float2(bq.z, bq.w);
    k[2 * j] = make_float2(bk.x, bk.y); k[2 * j + 1] = make_float2(bk.z, bk.w);
  }
  // the feature row in batches of kPfBatch INDEPENDENT loads (a plain `for c` loop issues one load at a time and every FMA
  // waits for its own: ~17 exposed global latencies per voxel)
  for (int c0 = 0; c0 < Cf; c0 += kPfBatch) {
    float v[kPfBatch];
#pragma unroll
    for (int u = 0; u < kPfBatch; ++u) v[u] = c0 + u < Cf ? __ldg(fp + c0 + u) : 0.f;
#pragma unroll
    for (int u = 0; u < kPfBatch; ++u) {
      if (c0 + u < Cf) {
        const float2 vv = make_float2(v[u], v[u]);
        const float4* wc = sw + (c0 + u) * 2 * Q4;
#pragma unroll
        for (int j = 0; j < Q4; ++j) {
          const float4 wk = wc[Q4 + j];
          k[2 * j] = __ffma2_rn(vv, make_float2(wk.x, wk.y), k[2 * j]);
          k[2 * j + 1] = __ffma2_rn(vv, make_float2(wk.z, wk.w), k[2 * j + 1]);
          if (want_q) {
            const float4 wq = wc[j];
            q[2 * j] = __ffma2_rn(vv, make_float2(wq.x, wq.y), q[2 * j]);
            q[2 * j + 1] = __ffma2_rn(vv, make_float2(wq.z, wq.w), q[2 * j + 1]);
          }
        }
      }
    }
  }
}

template <int F>
__global__ void __launch_bounds__(kPfYT * kPfXT)
k_pcm_fused(const PcmGeom g, const float* __restrict__ f, const float* __restrict__ cam, const float* __restrict__ tw,
            const float* __restrict__ tb, const float* __restrict__ pw, const float* __restrict__ pb, float* __restrict__ out,
            int zseg, int tiles_x) {
  constexpr int Q4 = F / 4;
  extern __shared__ float4 pf_sm[];
  float4* sw = pf_sm;                                   // [Cf][2*Q4]: theta | phi weights of one input channel
  float4* sb = sw + g.Cf * 2 * Q4;                      // [2*Q4] biases
  float4* kp = sb + 2 * Q4;                             // [3][Q4][kPfHV]
  float* cp = reinterpret_cast<float*>(kp + 3 * Q4 * kPfHV);   // [3][kPfHV]
  const int tid = threadIdx.x;
  for (int i = tid; i < g.Cf * 2 * F; i += blockDim.x) {
    const int c = i / (2 * F), j = i - c * 2 * F;
    reinterpret_cast<float*>(sw)[i] = j < F ? tw[j * g.Cf + c] : pw[(j - F) * g.Cf + c];
  }
  for (int i = tid; i < 2 * F; i += blockDim.x) reinterpret_cast<float*>(sb)[i] = i < F ? tb[i] : pb[i - F];
  __syncthreads();

  const int W = g.W, H = g.H, D = g.D, flags = g.flags;
  const bool relu = flags & 1;
  const int x0 = (blockIdx.x % tiles_x) * kPfXT, y0 = (blockIdx.x / tiles_x) * kPfYT;
  const int z0 = blockIdx.y * zseg, z1 = min(z0 + zseg, D), b = blockIdx.z;
  const int tx = tid % kPfXT, ty = tid / kPfXT;
  const int x = x0 + tx, y = y0 + ty;
  const bool live = x < W && y < H;
  const int hv = (ty + 1) * kPfHX + tx + 1;
  // halo ring element of this thread (tid < kPfRing): top row, bottom row, left column, right column
  int rhy = 0, rhx = 0;
  if (tid < kPfHX) { rhy = 0; rhx = tid; }
  else if (tid < 2 * kPfHX) { rhy = kPfYT + 1; rhx = tid - kPfHX; }
  else { const int r = tid - 2 * kPfHX; rhy = 1 + r % kPfYT; rhx = (r / kPfYT) ? kPfXT + 1 : 0; }
  const int ry = y0 + rhy - 1, rx = x0 + rhx - 1;
  const bool ring = tid < kPfRing && ry >= 0 && ry < H && rx >= 0 && rx < W;
  const int rhv = rhy * kPfHX + rhx;
  const long long HW = (long long)H * W, vol = HW * D;

  float2 q[F / 2], qn[F / 2], kk[F / 2];
  auto project_plane = [&](int z, bool keep_q) {
    const int slot = z % 3;
    float4* kps = kp + slot * Q4 * kPfHV;
    float* cps = cp + slot * kPfHV;
    const long long base = b * vol + z * HW;
    if (live) {
      const long long i = base + (long long)y * W + x;
      pcm_project_voxel<F>(f + i * g.Cf, g.Cf, sw, sb, keep_q, qn, kk);
#pragma unroll
      for (int j = 0; j < Q4; ++j) kps[j * kPfHV + hv] = make_float4(kk[2 * j].x, kk[2 * j].y, kk[2 * j + 1].x, kk[2 * j + 1].y);
      cps[hv] = __ldg(cam + i);
    }
    if (ring) {
      const long long i = base + (long long)ry * W + rx;
      float2 dummy[F / 2];
      pcm_project_voxel<F>(f + i * g.Cf, g.Cf, sw, sb, false, dummy, kk);
#pragma unroll
      for (int j = 0; j < Q4; ++j) kps[j * kPfHV + rhv] = make_float4(kk[2 * j].x, kk[2 * j].y, kk[2 * j + 1].x, kk[2 * j + 1].y);
      cps[rhv] = __ldg(cam + i);
    }
  };

  if (z0 > 0) project_plane(z0 - 1, false);
  project_plane(z0, true);
#pragma unroll
  for (int j = 0; j < F / 2; ++j) q[j] = qn[j];
  const bool ym = y > 0, yp = y < H - 1, xm = x > 0, xp = x < W - 1;
  for (int z = z0; z < z1; ++z) {
    if (z + 1 < D) project_plane(z + 1, true);
    __syncthreads();
    if (live) {
      const bool zm = z > 0, zp = z < D - 1;
      const int deg = degree_from(flags, zm + zp, ym + yp, xm + xp);
      const float invT = inv_temperature(flags, deg);
      // softmax bound instead of a first pass for the maximum: relu'd logits are >= 0 and <= |q| |k|; with the scores scaled by
      // 1/T the exponent is taken relative to the FIRST in-grid neighbour's score (any reference cancels in acc / l; the
      // first score keeps the exponents near 0 like the true maximum does for these O(1) logits) — no score arrays, half
      // the registers of k_pcm_attend, one exp per neighbour
      float ref = 0.f, l = 0.f, acc = 0.f;
      bool first = true;
      PCM_FOR_EACH_OFFSET(flags) {
        const bool in = (dz < 0 ? zm : (dz > 0 ? zp : true)) && (dy < 0 ? ym : (dy > 0 ? yp : true)) && (dx < 0 ? xm : (dx > 0 ? xp : true));
        if (in) {
          const int slot = (z + dz + 3) % 3, n = hv + dy * kPfHX + dx;
          const float4* kps = kp + slot * Q4 * kPfHV + n;
          float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < Q4; ++j) {
            const float4 k4 = kps[j * kPfHV];
            a2 = __ffma2_rn(q[2 * j], make_float2(k4.x, k4.y), a2);
            a2 = __ffma2_rn(q[2 * j + 1], make_float2(k4.z, k4.w), a2);
          }
          float d = a2.x + a2.y;
          if (relu) d = fmaxf(d, 0.f);
          d *= invT;
          if (first) { ref = d; first = false; }
          // online update keeps the result exact for any logits: rescale when a larger score shows up
          if (d > ref + 60.f) { const float sc = __expf(ref - d); l *= sc; acc *= sc; ref = d; }
          const float e = __expf(d - ref);
          l += e;
          acc = fmaf(e, cp[slot * kPfHV + n], acc);
        }
      }
      out[b * vol + z * HW + (long long)y * W + x] = deg > 0 ? acc / l : 0.f;
    }
    __syncthreads();                                       // plane z-1's slot is overwritten by the next projection
#pragma unroll
    for (int j = 0; j < F / 2; ++j) q[j] = qn[j];
  }
}

// backward, one thread per voxel i in both of its roles, nothing but `stats` kept from the forward:
//  node role   (i = x):  a_o = exp(s_o/T_x - m_x)/l_x,  ds_o = a_o (g_x cam_{x+o} - g_x s_x) / T_x [logit > 0],
//                        dq_x = sum_o ds_o k_{x+o}
//  gather role (i = y):  over the nodes x = y - o that list y as their neighbour o: the same a, ds from q_x, stats_x, g_x and
//                        the own k_y, cam_y:  dcam_y = sum g_x a,  dk_y = sum ds q_x
// (sum_o a_o g cam_{x+o} = g_x s_x is the softmax-Jacobian dot product.)   dqk: blocks of 32 rows like qk.
template <int F>
__global__ void __launch_bounds__(256, F <= 8 ? 3 : 1)
k_pcm_bwd(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, const float4* __restrict__ stats,
          const float* __restrict__ dout, float* __restrict__ dcam, float* __restrict__ dqk) {
  const PcmPos P(g);
  const int W = g.W, HW = g.H * g.W, flags = g.flags;
  const bool relu = flags & 1;
  const float* qp = qk_at<F>(qk, P.i);
  float q[F], k[F], dq[F], dk[F];
#pragma unroll
  for (int j = 0; j < F; ++j) { q[j] = __ldg(qp + j * 32); k[j] = __ldg(qp + (F + j) * 32); dq[j] = 0.f; dk[j] = 0.f; }
  const float4 st = __ldg(stats + P.i);                 // m, 1/l, 1/T, s
  const float go = __ldg(dout + P.i), cam_i = __ldg(cam + P.i);
  const float dot = go * st.w;
  float dc = 0.f;
  PCM_FOR_EACH_OFFSET(flags) {
    const int d_i = dz * HW + dy * W + dx;
    if (P.inside<1>(dz, dy, dx)) {                      // node role: neighbour n = i + o
      const int n = P.i + d_i;
      const float* kp = qk_at<F>(qk, n) + F * 32;
      float kn[F];
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) { kn[j] = __ldg(kp + j * 32); d = fmaf(q[j], kn[j], d); }
      const bool dead = relu && !(d > 0.f);
      if (relu) d = fmaxf(d, 0.f);
      const float a = __expf(fmaf(d, st.z, -st.x)) * st.y;
      const float ds = dead ? 0.f : a * fmaf(go, __ldg(cam + n), -dot) * st.z;
#pragma unroll
      for (int j = 0; j < F; ++j) dq[j] = fmaf(ds, kn[j], dq[j]);
    }
    if (P.inside<-1>(dz, dy, dx)) {                     // gather role: node n = i - o has i as its neighbour o
      const int n = P.i - d_i;
      const float* np = qk_at<F>(qk, n);
      float qn[F];
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) { qn[j] = __ldg(np + j * 32); d = fmaf(qn[j], k[j], d); }
      const bool dead = relu && !(d > 0.f);
      if (relu) d = fmaxf(d, 0.f);
      const float4 sn = __ldg(stats + n);
      const float gn = __ldg(dout + n);
      const float a = __expf(fmaf(d, sn.z, -sn.x)) * sn.y;
      dc = fmaf(gn, a, dc);
      const float ds = dead ? 0.f : a * gn * (cam_i - sn.w) * sn.z;
#pragma unroll
      for (int j = 0; j < F; ++j) dk[j] = fmaf(ds, qn[j], dk[j]);
    }
  }
  if (!P.live) return;
  dcam[P.i] = dc;
  float* dp = dqk + (size_t)(P.i >> 5) * (2 * F * 32) + (P.i & 31);
#pragma unroll
  for (int j = 0; j < F; ++j) { dp[j * 32] = dq[j]; dp[(F + j) * 32] = dk[j]; }
}

// backward of the projections over chunks of 256 rows (dqk = [dq | dk] from k_pcm_bwd, G = 2F columns):
//   df[r][c]  = sum_j W[j][c] dqk[r][j]              one thread per row, dqk row in registers, W as broadcast float4 reads;
//                                                   rows leave through shared memory as one contiguous, coalesced block
//   dW[j][c] += sum_r dqk[r][j] f[r][c], db[j] += sum_r dqk[r][j]
//                                                   register tiles of 4 (j) x 4 (c) per thread over a [256][G] x [256][Cf+1]
//                                                   pair of shared-memory tiles (f gets a column of ones for the bias);
//                                                   the JG*CG tile owners form a row group, 256 / (JG*CG) groups split the rows
__global__ void __launch_bounds__(256)
k_pcm_bwd_params(const float* __restrict__ f, const float* __restrict__ tw, const float* __restrict__ pw,
                 const float* __restrict__ dqk, float* __restrict__ df, float* __restrict__ partial, long long rows,
                 int Cf, int F) {
  extern __shared__ float4 bsm4[];
  float* bsm = reinterpret_cast<float*>(bsm4);
  const int G = 2 * F, JG = (G + 3) >> 2, CG = (Cf + 4) >> 2;      // CG covers Cf + 1 columns
  const int GP = 4 * JG + 4, CP = 4 * CG + 4;                       // row pitches (floats), multiples of 4
  float* sf = bsm;                       // [256][CP]  f rows, column Cf = 1, rest 0
  float* sg = sf + 256 * CP;             // [256][GP]  dqk rows
  float* sw = sg + 256 * GP;             // [Cf][4*JG] W^T (theta | phi)
  float* sd = sf;                        // df rows [256][Cf] reuse sf after the tile phase
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned inv_cf = (1u << 20) / (unsigned)Cf + 1u;          // i / Cf == (i * inv_cf) >> 20 for i < 256 * Cf <= 2^14
  for (int i = tid; i < Cf * 4 * JG; i += blockDim.x) {
    const int c = i / (4 * JG), j = i - c * 4 * JG;
    sw[i] = j < F ? tw[j * Cf + c] : (j < G ? pw[(j - F) * Cf + c] : 0.f);
  }
  const int NT = JG * CG, groups = 256 / NT;
  const int grp = tid / NT, tt = tid - grp * NT, jg = tt / CG, cg = tt - jg * CG;
  const bool owner = grp < groups;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long chunks = (rows + 255) >> 8;
  for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const long long base = ch << 8;
    const int n = (int)(rows - base < 256 ? rows - base : 256);
    __syncthreads();                                                // previous chunk's sd / sg readers are done
    for (int i = tid; i < 256 * CP; i += blockDim.x) sf[i] = 0.f;
    __syncthreads();
    for (int i0 = tid; i0 < n * Cf; i0 += 256 * 6) {                 // contiguous, coalesced, 6 loads in flight
      float v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) v[u] = i0 + 256 * u < n * Cf ? __ldg(f + base * Cf + i0 + 256 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + 256 * u;
        if (i < n * Cf) { const int r = (int)(((unsigned)i * inv_cf) >> 20); sf[r * CP + (i - r * Cf)] = v[u]; }
      }
    }
    float gr[2 * kMaxF];
    if (tid < n) {
      sf[tid * CP + Cf] = 1.f;
      const float* gp = dqk + (size_t)((base + tid) >> 5) * (G * 32) + lane;     // base is a multiple of 32
#pragma unroll
      for (int j = 0; j < 2 * kMaxF; ++j) gr[j] = j < G ? __ldg(gp + j * 32) : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 2 * kMaxF; ++j) gr[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2 * kMaxF; ++j)
      if (j < GP) sg[tid * GP + j] = gr[j];                         // rows >= n and columns >= G are zeros
    __syncthreads();
    if (owner) {
      for (int r = grp; r < n; r += groups) {
        const float4 gv = *reinterpret_cast<const float4*>(sg + r * GP + 4 * jg);
        const float4 fv = *reinterpret_cast<const float4*>(sf + r * CP + 4 * cg);
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, fa[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], fa[b], acc[a][b]);
      }
    }
    __syncthreads();                                                // sf is free: df rows go there
    if (tid < n) {
      for (int c = 0; c < Cf; ++c) {
        const float4* wc = reinterpret_cast<const float4*>(sw) + c * JG;
        float d = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < kMaxF / 2; ++j4)
          if (j4 < JG) {
            const float4 w4 = wc[j4];
            d = fmaf(w4.x, gr[4 * j4], d); d = fmaf(w4.y, gr[4 * j4 + 1], d);
            d = fmaf(w4.z, gr[4 * j4 + 2], d); d = fmaf(w4.w, gr[4 * j4 + 3], d);
          }
        sd[tid * (Cf | 1) + c] = d;
      }
    }
    __syncthreads();
    for (int i = tid; i < n * Cf; i += 256) {
      const int r = (int)(((unsigned)i * inv_cf) >> 20), c = i - r * Cf;
      df[base * Cf + i] = sd[r * (Cf | 1) + c];
    }
  }
  // tile owners -> shared memory -> ONE partial [4*JG][4*CG] per block (row groups summed in order);
  // k_pcm_params_reduce sums the blocks in a fixed order: deterministic
  __syncthreads();
  float* sp = bsm;                                                  // [groups][16*NT] <= 256*16 floats
  if (owner) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) sp[grp * (16 * NT) + (4 * jg + a) * (4 * CG) + 4 * cg + b] = acc[a][b];
  }
  __syncthreads();
  for (int e = tid; e < 16 * NT; e += 256) {
    float t = 0.f;
    for (int gI = 0; gI < groups; ++gI) t += sp[gI * (16 * NT) + e];
    partial[(size_t)blockIdx.x * (16 * NT) + e] = t;
  }
}

// dparams[which][F*Cf weights | F biases] (double) from the per-block tiles: one warp per output element, lanes stride
// over the blocks, fixed-order shuffle tree
__global__ void __launch_bounds__(256)
k_pcm_params_reduce(const float* __restrict__ partial, int nparts, int Cf, int F, double* __restrict__ dparams) {
  const int G = 2 * F, JG = (G + 3) >> 2, CG = (Cf + 4) >> 2;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= G * (Cf + 1)) return;
  const int j = e / (Cf + 1), c = e - j * (Cf + 1);
  double t = 0.0;
  for (int p = lane; p < nparts; p += 32) t += (double)partial[(size_t)p * (16 * JG * CG) + j * (4 * CG) + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  const int which = j / F, jj = j - which * F;
  if (lane == 0) dparams[which * F * (Cf + 1) + (c < Cf ? jj * Cf + c : F * Cf + jj)] = t;
}

}  // namespace dram

using namespace dram;

static int pcm_geom(PcmGeom& g, int B, int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags) {
  DRAM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "pcm: bad grid");
  DRAM_REQUIRE((long long)B * D * H * W < (1ll << 31) - 64 && (long long)B * D <= 65535, "pcm: grid too large (B*V < 2^31, B*D <= 65535)");
  DRAM_REQUIRE(Cf > 0 && Cf <= kMaxCf, "pcm: Cf=%d (<=%d) unsupported", Cf, kMaxCf);
  DRAM_REQUIRE(F == 4 || F == 8 || F == 16, "pcm: F=%d unsupported (4 | 8 | 16)", F);
  DRAM_REQUIRE(connectivity >= 1 && connectivity <= 3, "pcm: connectivity %d unsupported", connectivity);
  g.B = B; g.D = D; g.H = H; g.W = W; g.Cf = Cf; g.F = F;
  g.flags = (flags & 0x7f) | ((self_loop ? 1 : 0) << 7) | (connectivity << 8);
  g.O = count_offsets(connectivity, self_loop);
  return DRAM_OK;
}

extern "C" {

int dram_pcm_num_offsets(int connectivity, int self_loop) { return count_offsets(connectivity, self_loop); }

size_t dram_pcm_qk_floats(long long rows, int F) { return (size_t)((rows + 31) / 32) * 32 * 2 * (size_t)F; }

static const int kPcmParamBlocks = kNumSMs * 3;
/* scratch of dram_pcm_bwd: the [dq | dk] blocks + one [4*JG][4*CG] tile of parameter-gradient partials per block */
size_t dram_pcm_bwd_ws_floats(long long rows, int Cf, int F) {
  const int JG = (2 * F + 3) / 4, CG = (Cf + 4) / 4;
  return dram_pcm_qk_floats(rows, F) + (size_t)kPcmParamBlocks * (size_t)(16 * JG * CG);
}

int dram_pcm_fwd(const float* f, const float* cam, const float* theta_w, const float* theta_b, const float* phi_w,
                 const float* phi_b, float* qk, float* stats, float* out, int B, int D, int H, int W, int Cf, int F,
                 int connectivity, int self_loop, int flags, void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && theta_b && phi_w && phi_b && out, "pcm_fwd: null pointer");
  DRAM_REQUIRE(qk || (!stats && !getenv("DRAM_PCM_TWO_KERNELS")), "pcm_fwd: qk workspace is required whenever statistics are kept for the backward");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  if (!stats && !getenv("DRAM_PCM_TWO_KERNELS")) {
    // inference: nothing kept for a backward -> projection + attention in one kernel, projections stay on chip
    const int tiles_x = (W + kPfXT - 1) / kPfXT, tiles_y = (H + kPfYT - 1) / kPfYT;
    int zsegs = (3 * kNumSMs) / (B * tiles_x * tiles_y);      // at most ONE wave of 3 resident blocks per SM (a second, nearly empty wave doubles the time)
    if (zsegs > D / 4) zsegs = D / 4;
    if (zsegs < 1) zsegs = 1;
    const int zseg = (D + zsegs - 1) / zsegs;
    zsegs = (D + zseg - 1) / zseg;
    const size_t sm = sizeof(float4) * ((size_t)Cf * 2 * (F / 4) + 2 * (F / 4) + 3 * (size_t)(F / 4) * kPfHV) + sizeof(float) * 3 * kPfHV;
    const dim3 gridf((unsigned)(tiles_x * tiles_y), (unsigned)zsegs, (unsigned)B);
    DRAM_REQUIRE(B <= 65535 && zsegs <= 65535, "pcm_fwd: batch too large for the fused kernel's grid");
#define PCM_FUSED(FF)                                                                                              \
  do {                                                                                                             \
    if (sm > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_fused<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_pcm_fused<FF><<<gridf, kPfYT * kPfXT, sm, st>>>(g, f, cam, theta_w, theta_b, phi_w, phi_b, out, zseg, tiles_x);            \
  } while (0)
    if (F == 4) PCM_FUSED(4);
    else if (F == 8) PCM_FUSED(8);
    else PCM_FUSED(16);
#undef PCM_FUSED
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  const int J4 = (2 * F + 3) / 4;
  const size_t smem = sizeof(float) * ((size_t)Cf * 4 * J4 + 4 * J4 + (size_t)8 * 32 * (Cf | 1));
  if (smem > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pcm_project<<<grid_for(rows, 256, 8), 256, smem, st>>>(f, theta_w, theta_b, phi_w, phi_b, qk, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  const dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)(B * D));
  float4* s4 = reinterpret_cast<float4*>(stats);
#define PCM_ATTEND(FF)                                                                  \
  do {                                                                                  \
    if (stats) k_pcm_attend<FF, true><<<grid, 256, 0, st>>>(g, qk, cam, s4, out);       \
    else k_pcm_attend<FF, false><<<grid, 256, 0, st>>>(g, qk, cam, nullptr, out);       \
  } while (0)
  if (F == 4) PCM_ATTEND(4);
  else if (F == 8) PCM_ATTEND(8);
  else PCM_ATTEND(16);
#undef PCM_ATTEND
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_pcm_bwd(const float* f, const float* cam, const float* theta_w, const float* phi_w, const float* qk,
                 const float* stats, const float* dout, float* dqk_ws, float* dcam, float* df, double* dparams, int B,
                 int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags, void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && phi_w && qk && stats && dout && dqk_ws && dcam && df && dparams, "pcm_bwd: null pointer");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  const dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)(B * D));
  const float4* s4 = reinterpret_cast<const float4*>(stats);
  if (F == 4) k_pcm_bwd<4><<<grid, 256, 0, st>>>(g, qk, cam, s4, dout, dcam, dqk_ws);
  else if (F == 8) k_pcm_bwd<8><<<grid, 256, 0, st>>>(g, qk, cam, s4, dout, dcam, dqk_ws);
  else k_pcm_bwd<16><<<grid, 256, 0, st>>>(g, qk, cam, s4, dout, dcam, dqk_ws);
  DRAM_LAUNCH_CHECK();
  const int JG = (2 * F + 3) / 4, CG = (Cf + 4) / 4;
  DRAM_REQUIRE(JG * CG <= 256, "pcm_bwd: 2F=%d x Cf=%d tiles exceed one block", 2 * F, Cf);
  const size_t smem = sizeof(float) * ((size_t)256 * (4 * CG + 4) + (size_t)256 * (4 * JG + 4) + (size_t)Cf * 4 * JG);
  const long long chunks = (rows + 255) / 256;
  const int pgrid = (int)(chunks < (long long)kPcmParamBlocks ? chunks : (long long)kPcmParamBlocks);
  float* partial = dqk_ws + dram_pcm_qk_floats(rows, F);
  if (smem > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_bwd_params, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pcm_bwd_params<<<pgrid, 256, smem, st>>>(f, theta_w, phi_w, dqk_ws, df, partial, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  const int nout = 2 * F * (Cf + 1);
  k_pcm_params_reduce<<<(nout * 32 + 255) / 256, 256, 0, st>>>(partial, pgrid, Cf, F, dparams);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
