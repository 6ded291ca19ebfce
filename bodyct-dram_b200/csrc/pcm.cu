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
//   qk    [2F][R]  theta|phi projections as PLANES (structure of arrays): a warp = 32 consecutive x of one grid row, so every
//                  load of one feature of one neighbour offset is one coalesced 128-byte request (the [R][2F] rows of the
//                  first version cost 16 L1 wavefronts per request - the kernels were LSU-bound at 2-4 % of HBM peak);
//   stats [2][R]   running max m and normaliser l of the node's softmax, kept for the backward INSTEAD of the [R][O]
//                  attention weights: a_{x,o} = exp(s_{x,o} - m_x) / l_x is recomputed from q_x and k_{x+o} (8 FMAs + 1 exp).
#include "common.cuh"

namespace dram {

constexpr int kMaxOff = 27, kMaxF = 16, kMaxCf = 64;

struct PcmGeom {
  int B, D, H, W, Cf, F, O, flags;
  signed char off[kMaxOff][3];
};

static int build_offsets(int connectivity, int self_loop, signed char off[kMaxOff][3]) {
  int n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        int l1 = abs(dz) + abs(dy) + abs(dx);
        bool take = (l1 == 0) ? (self_loop != 0) : (l1 <= connectivity);
        if (take) { off[n][0] = (signed char)dz; off[n][1] = (signed char)dy; off[n][2] = (signed char)dx; ++n; }
      }
  return n;
}

__device__ __forceinline__ bool nb_index(const PcmGeom& g, int z, int y, int x, int o, long long& idx) {
  int zz = z + g.off[o][0], yy = y + g.off[o][1], xx = x + g.off[o][2];
  if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) return false;
  idx = ((long long)zz * g.H + yy) * g.W + xx;
  return true;
}
__device__ __forceinline__ float temperature(const PcmGeom& g, int deg) {
  int mode = (g.flags >> 1) & 3;
  return mode == 1 ? sqrtf((float)deg) : (mode == 2 ? 0.01f : 1.f);
}

// number of in-grid neighbours of (z,y,x) in closed form: with a_z = #{dz in {-1,+1} : z+dz in grid} etc. the stencil
// {|dz|+|dy|+|dx| <= connectivity} has a_z+a_y+a_x members at L1 distance 1, a_z a_y + a_z a_x + a_y a_x at 2, a_z a_y a_x at 3
__device__ __forceinline__ int degree_of(const PcmGeom& g, int z, int y, int x) {
  const int az = (z > 0) + (z < g.D - 1), ay = (y > 0) + (y < g.H - 1), ax = (x > 0) + (x < g.W - 1);
  const int conn = g.flags >> 8;
  int deg = (g.flags >> 7) & 1;                               // self loop
  deg += az + ay + ax;
  if (conn >= 2) deg += az * ay + az * ax + ay * ax;
  if (conn >= 3) deg += az * ay * ax;
  return deg;
}

// The stencil as 27 fully unrolled candidates in build_offsets order (dz, dy, dx ascending): dz/dy/dx are compile-time
// constants inside the body, membership is a warp-uniform test (indexing the by-value offset table would put it in local memory).
#define PCM_FOR_EACH_OFFSET(g)                                                                                      \
  _Pragma("unroll") for (int t_ = 0; t_ < 27; ++t_)                                                                 \
    if (const int dz = t_ / 9 - 1, dy = (t_ / 3) % 3 - 1, dx = t_ % 3 - 1, l1_ = (dz != 0) + (dy != 0) + (dx != 0); \
        l1_ == 0 ? (((g).flags >> 7) & 1) != 0 : l1_ <= ((g).flags >> 8))

// qk[j][r] = theta f_r + b (j < F), phi f_r + b (F <= j < 2F).  A warp stages its 32 consecutive rows (32*Cf contiguous
// floats) through shared memory with coalesced loads; a lane then walks its row at an odd pitch (conflict-free).
__global__ void __launch_bounds__(256)
k_pcm_project(const float* __restrict__ f, const float* __restrict__ tw, const float* __restrict__ tb,
              const float* __restrict__ pw, const float* __restrict__ pb, float* __restrict__ qk, long long rows, int Cf, int F) {
  extern __shared__ float4 psm4[];
  float* psm = reinterpret_cast<float*>(psm4);
  const int Cfp = Cf | 1, J4 = (2 * F + 3) >> 2;    // outputs in groups of 4: the weights of one input channel are J4 float4
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
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5, tiles = (rows + 31) >> 5;
  for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < tiles; t += warps) {
    const long long r0 = t << 5;
    const int n = (int)(rows - r0 < 32 ? rows - r0 : 32);
    const float* src = f + r0 * Cf;
    __syncwarp();
    for (int i = lane; i < n * Cf; i += 32) { const int r = i / Cf; st[r * Cfp + (i - r * Cf)] = __ldg(src + i); }
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
      float* dst = qk + r0 + lane;
#pragma unroll
      for (int j = 0; j < kMaxF / 2; ++j) {
        if (4 * j + 0 < 2 * F) dst[(long long)(4 * j + 0) * rows] = acc[j].x;
        if (4 * j + 1 < 2 * F) dst[(long long)(4 * j + 1) * rows] = acc[j].y;
        if (4 * j + 2 < 2 * F) dst[(long long)(4 * j + 2) * rows] = acc[j].z;
        if (4 * j + 3 < 2 * F) dst[(long long)(4 * j + 3) * rows] = acc[j].w;
      }
    }
  }
}

// forward: one thread per node, online softmax over the in-grid neighbours (no per-offset arrays).  SAVE: keep (m, l).
template <int FT, bool SAVE>
__global__ void __launch_bounds__(256)
k_pcm_attend(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, float* __restrict__ stats,
             float* __restrict__ out) {
  const long long V = (long long)g.D * g.H * g.W, R = (long long)g.B * V;
  const int F = g.F;
  const bool relu = g.flags & 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    float q[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) q[j] = j < F ? __ldg(qk + (long long)j * R + i) : 0.f;
    const int deg = degree_of(g, z, y, x);
    const float invT = 1.f / temperature(g, deg);
    const float* kbase = qk + (long long)F * R + i;
    float m = -INFINITY, l = 0.f, acc = 0.f;
    PCM_FOR_EACH_OFFSET(g) {
      const int zz = z + dz, yy = y + dy, xx = x + dx;
      if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
      const long long d_i = ((long long)dz * g.H + dy) * g.W + dx;
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < FT; ++j)
        if (j < F) d = fmaf(q[j], __ldg(kbase + (long long)j * R + d_i), d);
      if (relu) d = fmaxf(d, 0.f);
      const float sv = d * invT, cv = __ldg(cam + i + d_i);
      const float mn = fmaxf(m, sv);
      const float corr = __expf(m - mn), e = __expf(sv - mn);      // m = -inf on the first neighbour -> corr = 0
      l = l * corr + e;
      acc = acc * corr + e * cv;
      m = mn;
    }
    out[i] = deg > 0 ? acc / l : 0.f;     // degree-0 nodes (1x1x1 grids) keep 0 like DGL's zero-filled result
    if (SAVE) { stats[i] = deg > 0 ? m : 0.f; stats[R + i] = deg > 0 ? l : 1.f; }
  }
}

// backward, one thread per voxel i in both of its roles, nothing but (m, l) and the forward output s kept from the forward:
//  node role   (i = x):  a_o = exp(s_o - m_x)/l_x,  ds_o = a_o (g_x cam_{x+o} - g_x s_x) / T_x [logit > 0],
//                        dq_x = sum_o ds_o k_{x+o}
//  gather role (i = y):  over the nodes x = y - o that list y as their neighbour o: the same a, ds from q_x, (m, l, g, s)_x and
//                        the own k_y, cam_y:  dcam_y = sum g_x a,  dk_y = sum ds q_x
// (sum_o a_o g cam_{x+o} = g_x s_x is the softmax-Jacobian dot product.)   dqk: [2F][R] planes like qk.
template <int FT>
__global__ void __launch_bounds__(256, FT <= 8 ? 2 : 1)
k_pcm_bwd(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, const float* __restrict__ stats,
          const float* __restrict__ s_out, const float* __restrict__ dout, float* __restrict__ dcam, float* __restrict__ dqk) {
  const long long V = (long long)g.D * g.H * g.W, R = (long long)g.B * V;
  const int F = g.F;
  const bool relu = g.flags & 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    float q[FT], k[FT], dq[FT], dk[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) {
      q[j] = j < F ? __ldg(qk + (long long)j * R + i) : 0.f;
      k[j] = j < F ? __ldg(qk + (long long)(F + j) * R + i) : 0.f;
      dq[j] = 0.f; dk[j] = 0.f;
    }
    const int deg = degree_of(g, z, y, x);
    const float invT = 1.f / temperature(g, deg);
    const float go = __ldg(dout + i), m = __ldg(stats + i), il = 1.f / __ldg(stats + R + i), dot = go * __ldg(s_out + i);
    const float cam_i = __ldg(cam + i);
    float dc = 0.f;
    PCM_FOR_EACH_OFFSET(g) {
      const long long d_i = ((long long)dz * g.H + dy) * g.W + dx;
      {   // node role: neighbour n = i + o
        const int zz = z + dz, yy = y + dy, xx = x + dx;
        if (zz >= 0 && zz < g.D && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) {
          float kn[FT];
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < FT; ++j) {
            kn[j] = j < F ? __ldg(qk + (long long)(F + j) * R + i + d_i) : 0.f;
            d = fmaf(q[j], kn[j], d);
          }
          const bool dead = relu && !(d > 0.f);
          if (relu) d = fmaxf(d, 0.f);
          const float a = __expf(d * invT - m) * il;
          const float ds = dead ? 0.f : a * (go * __ldg(cam + i + d_i) - dot) * invT;
#pragma unroll
          for (int j = 0; j < FT; ++j) dq[j] = fmaf(ds, kn[j], dq[j]);
        }
      }
      {   // gather role: node n = i - o has i as its neighbour o
        const int zz = z - dz, yy = y - dy, xx = x - dx;
        if (zz >= 0 && zz < g.D && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) {
          const long long n = i - d_i;
          float qn[FT];
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < FT; ++j) {
            qn[j] = j < F ? __ldg(qk + (long long)j * R + n) : 0.f;
            d = fmaf(qn[j], k[j], d);
          }
          const bool dead = relu && !(d > 0.f);
          if (relu) d = fmaxf(d, 0.f);
          const float invTn = 1.f / temperature(g, degree_of(g, zz, yy, xx));
          const float gn = __ldg(dout + n);
          const float a = __expf(d * invTn - __ldg(stats + n)) / __ldg(stats + R + n);
          dc = fmaf(gn, a, dc);
          const float ds = dead ? 0.f : a * (gn * cam_i - gn * __ldg(s_out + n)) * invTn;
#pragma unroll
          for (int j = 0; j < FT; ++j) dk[j] = fmaf(ds, qn[j], dk[j]);
        }
      }
    }
    dcam[i] = dc;
#pragma unroll
    for (int j = 0; j < FT; ++j)
      if (j < F) { dqk[(long long)j * R + i] = dq[j]; dqk[(long long)(F + j) * R + i] = dk[j]; }
  }
}

// backward pass 3: df = theta^T dq + phi^T dk; dparams += [dq (x) f, dq, dk (x) f, dk] reduced over all voxels
constexpr int kPcmChunk = 128;
__global__ void __launch_bounds__(320)
k_pcm_bwd_params(const float* __restrict__ f, const float* __restrict__ tw, const float* __restrict__ pw,
                 const float* __restrict__ dqk, float* __restrict__ df, double* __restrict__ dparams, long long rows,
                 int Cf, int F) {
  extern __shared__ float sm[];
  float* sf = sm;                                 // [kPcmChunk][Cf]
  const int SG = 2 * F + 1;
  float* sg = sf + kPcmChunk * Cf;                // [kPcmChunk][2F + 1]
  float* sw = sg + kPcmChunk * SG;                // [2F][Cf]
  for (int i = threadIdx.x; i < F * Cf; i += blockDim.x) { sw[i] = tw[i]; sw[F * Cf + i] = pw[i]; }
  const int nout = 2 * F * (Cf + 1);              // outputs: for m in {theta,phi}: F*Cf weights then F biases
  float acc = 0.f;
  int t = threadIdx.x;
  int which = 0, j = 0, c = 0;
  bool owner = t < nout;
  if (owner) {
    int per = F * (Cf + 1);
    which = t / per;
    int r = t % per;
    if (r < F * Cf) { j = r / Cf; c = r % Cf; } else { j = r - F * Cf; c = -1; }
  }
  for (long long base = (long long)blockIdx.x * kPcmChunk; base < rows; base += (long long)gridDim.x * kPcmChunk) {
    int n = (int)(rows - base < kPcmChunk ? rows - base : kPcmChunk);
    __syncthreads();
    for (int i = threadIdx.x; i < n * Cf; i += blockDim.x) sf[i] = f[base * Cf + i];
    for (int i = threadIdx.x; i < n * 2 * F; i += blockDim.x) {           // dqk planes [2F][rows] -> sg[r][SG], SG odd
      const int jj = i / n, r = i - jj * n;
      sg[r * SG + jj] = dqk[(long long)jj * rows + base + r];
    }
    __syncthreads();
    if (owner) {
      for (int r = 0; r < n; ++r) {
        float gq = sg[r * SG + which * F + j];
        acc = c >= 0 ? fmaf(gq, sf[r * Cf + c], acc) : acc + gq;
      }
    }
    for (int i = threadIdx.x; i < n * Cf; i += blockDim.x) {
      int r = i / Cf, cc = i % Cf;
      float d = 0.f;
      for (int jj = 0; jj < 2 * F; ++jj) d = fmaf(sw[jj * Cf + cc], sg[r * SG + jj], d);
      df[base * Cf + i] = d;
    }
  }
  if (owner) atomicAdd(&dparams[t], (double)acc);
}

}  // namespace dram

using namespace dram;

static int pcm_geom(PcmGeom& g, int B, int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags) {
  DRAM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "pcm: bad grid");
  DRAM_REQUIRE(Cf > 0 && Cf <= kMaxCf && F > 0 && F <= kMaxF, "pcm: Cf=%d (<=%d) / F=%d (<=%d) unsupported", Cf, kMaxCf, F, kMaxF);
  DRAM_REQUIRE(connectivity >= 1 && connectivity <= 3, "pcm: connectivity %d unsupported", connectivity);
  g.B = B; g.D = D; g.H = H; g.W = W; g.Cf = Cf; g.F = F;
  g.flags = (flags & 0x7f) | ((self_loop ? 1 : 0) << 7) | (connectivity << 8);   // degree_of reads bits 7 and 8+
  g.O = build_offsets(connectivity, self_loop, g.off);
  return DRAM_OK;
}

extern "C" {

int dram_pcm_num_offsets(int connectivity, int self_loop) {
  signed char off[kMaxOff][3];
  return build_offsets(connectivity, self_loop, off);
}

int dram_pcm_fwd(const float* f, const float* cam, const float* theta_w, const float* theta_b, const float* phi_w,
                 const float* phi_b, float* qk, float* stats, float* out, int B, int D, int H, int W, int Cf, int F,
                 int connectivity, int self_loop, int flags, void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && theta_b && phi_w && phi_b && qk && out, "pcm_fwd: null pointer");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  const int J4 = (2 * F + 3) / 4;
  const size_t smem = sizeof(float) * ((size_t)Cf * 4 * J4 + 4 * J4 + (size_t)8 * 32 * (Cf | 1));
  if (smem > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pcm_project<<<grid_for(rows, 256, 8), 256, smem, st>>>(f, theta_w, theta_b, phi_w, phi_b, qk, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  const int grid = grid_for(rows, 256, 16);
  if (F <= 8) {
    if (stats) k_pcm_attend<8, true><<<grid, 256, 0, st>>>(g, qk, cam, stats, out);
    else k_pcm_attend<8, false><<<grid, 256, 0, st>>>(g, qk, cam, nullptr, out);
  } else {
    if (stats) k_pcm_attend<kMaxF, true><<<grid, 256, 0, st>>>(g, qk, cam, stats, out);
    else k_pcm_attend<kMaxF, false><<<grid, 256, 0, st>>>(g, qk, cam, nullptr, out);
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_pcm_bwd(const float* f, const float* cam, const float* theta_w, const float* phi_w, const float* qk,
                 const float* stats, const float* s_out, const float* dout, float* dqk_ws, float* dcam, float* df,
                 double* dparams, int B, int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags,
                 void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && phi_w && qk && stats && s_out && dout && dqk_ws && dcam && df && dparams, "pcm_bwd: null pointer");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  DRAM_REQUIRE(2 * F * (Cf + 1) <= 320, "pcm_bwd: 2F(Cf+1)=%d > 320 unsupported", 2 * F * (Cf + 1));
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  if (F <= 8) k_pcm_bwd<8><<<grid_for(rows, 256, 16), 256, 0, st>>>(g, qk, cam, stats, s_out, dout, dcam, dqk_ws);
  else k_pcm_bwd<kMaxF><<<grid_for(rows, 256, 16), 256, 0, st>>>(g, qk, cam, stats, s_out, dout, dcam, dqk_ws);
  DRAM_LAUNCH_CHECK();
  DRAM_CUDA(cudaMemsetAsync(dparams, 0, sizeof(double) * 2 * F * (Cf + 1), st));
  size_t smem = sizeof(float) * ((size_t)kPcmChunk * Cf + (size_t)kPcmChunk * (2 * F + 1) + (size_t)2 * F * Cf);
  long long chunks = (rows + kPcmChunk - 1) / kPcmChunk;
  int grid = (int)(chunks < (long long)kNumSMs * 4 ? chunks : (long long)kNumSMs * 4);
  if (smem > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_bwd_params, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pcm_bwd_params<<<grid, 320, smem, st>>>(f, theta_w, phi_w, dqk_ws, df, dparams, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
