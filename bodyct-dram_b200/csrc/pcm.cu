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
// Data: f [B][V][Cf], cam/out [B][V] (V = D*H*W, channels-last), qk [B][V][2F] (theta|phi projections),
// att [B][V][O] softmax weights (0 for absent neighbours) kept for the backward pass.
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

// qk[v][0:F] = theta f + b, qk[v][F:2F] = phi f + b
__global__ void __launch_bounds__(256)
k_pcm_project(const float* __restrict__ f, const float* __restrict__ tw, const float* __restrict__ tb,
              const float* __restrict__ pw, const float* __restrict__ pb, float* __restrict__ qk, long long rows, int Cf, int F) {
  __shared__ float sw[2 * kMaxF * kMaxCf];
  __shared__ float sb[2 * kMaxF];
  for (int i = threadIdx.x; i < F * Cf; i += blockDim.x) { sw[i] = tw[i]; sw[F * Cf + i] = pw[i]; }
  for (int i = threadIdx.x; i < F; i += blockDim.x) { sb[i] = tb[i]; sb[F + i] = pb[i]; }
  __syncthreads();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float acc[2 * kMaxF];
#pragma unroll
    for (int j = 0; j < 2 * kMaxF; ++j) acc[j] = j < 2 * F ? sb[j] : 0.f;
    const float* fp = f + r * Cf;
    for (int c = 0; c < Cf; ++c) {
      float v = __ldg(fp + c);
#pragma unroll
      for (int j = 0; j < 2 * kMaxF; ++j)
        if (j < 2 * F) acc[j] = fmaf(v, sw[j * Cf + c], acc[j]);
    }
    float* q = qk + r * 2 * F;
#pragma unroll
    for (int j = 0; j < 2 * kMaxF; ++j)
      if (j < 2 * F) q[j] = acc[j];
  }
}

__global__ void __launch_bounds__(256)
k_pcm_attend(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, float* __restrict__ att,
             float* __restrict__ out) {
  const long long V = (long long)g.D * g.H * g.W;
  const long long total = (long long)g.B * V;
  const int F = g.F, O = g.O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i % V, b = i / V;
    int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    float q[kMaxF];
#pragma unroll
    for (int j = 0; j < kMaxF; ++j) q[j] = j < F ? qk[i * 2 * F + j] : 0.f;
    float s[kMaxOff], c[kMaxOff];
    int deg = 0;
    for (int o = 0; o < O; ++o) {
      long long nb;
      if (nb_index(g, z, y, x, o, nb)) {
        const float* kp = qk + (b * V + nb) * 2 * F + F;
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxF; ++j)
          if (j < F) d = fmaf(q[j], __ldg(kp + j), d);
        if (g.flags & 1) d = fmaxf(d, 0.f);
        s[o] = d;
        c[o] = __ldg(cam + b * V + nb);
        ++deg;
      } else {
        s[o] = -INFINITY;
        c[o] = 0.f;
      }
    }
    float invT = 1.f / temperature(g, deg);
    float mx = -INFINITY;
    for (int o = 0; o < O; ++o) { s[o] *= invT; mx = fmaxf(mx, s[o]); }
    float den = 0.f;
    for (int o = 0; o < O; ++o) { s[o] = (s[o] == -INFINITY) ? 0.f : __expf(s[o] - mx); den += s[o]; }
    float inv = deg > 0 ? 1.f / den : 0.f, acc = 0.f;
    for (int o = 0; o < O; ++o) {
      float a = s[o] * inv;
      att[i * O + o] = a;
      acc = fmaf(a, c[o], acc);
    }
    out[i] = acc;     // degree-0 nodes (1x1x1 grids) keep 0 like DGL's zero-filled result
  }
}

// Inference variant (no attention weights kept): F = 8, online softmax in registers, neighbour keys as two float4 loads.
// Warp-level sharing: consecutive lanes hold consecutive x of one grid row, so for the three offsets (dz,dy,-1|0|+1) the
// key vector and the cam value of the (dz,dy) row are loaded ONCE per lane (at the lane's own x) and the x-1 / x+1
// neighbours come from the adjacent lanes through __shfl_sync; only the lanes at a warp or row edge load them directly.
__global__ void __launch_bounds__(256)
k_pcm_attend_infer(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, float* __restrict__ out) {
  const long long V = (long long)g.D * g.H * g.W;
  const long long total = (long long)g.B * V;
  const int lane = threadIdx.x & 31;
  const int O = g.O;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long iters = (total + stride - 1) / stride;          // warp-uniform trip count: shuffles need converged warps
  for (long long it = 0; it < iters; ++it) {
    const long long i = first + it * stride;
    const bool live = i < total;
    const long long ii = live ? i : total - 1;
    const long long v = ii % V, b = ii / V;
    const int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    const float4* qp = reinterpret_cast<const float4*>(qk + ii * 16);
    const float4 q0 = __ldg(qp), q1 = __ldg(qp + 1);
    // lanes lane-1 / lane+1 hold x-1 / x+1 of the same row?
    const int xl = __shfl_up_sync(0xffffffffu, x, 1), xr = __shfl_down_sync(0xffffffffu, x, 1);
    const long long il = __shfl_up_sync(0xffffffffu, ii, 1), ir = __shfl_down_sync(0xffffffffu, ii, 1);
    const bool left_ok = lane > 0 && xl == x - 1 && il == ii - 1;
    const bool right_ok = lane < 31 && xr == x + 1 && ir == ii + 1;
    int deg = 0;
    for (int o = 0; o < O; ++o) {
      int zz = z + g.off[o][0], yy = y + g.off[o][1], xx = x + g.off[o][2];
      deg += (zz >= 0 && zz < g.D && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? 1 : 0;
    }
    const float invT = 1.f / temperature(g, deg);
    float m = -INFINITY, l = 0.f, acc = 0.f;
    int cur_dz = 127, cur_dy = 127;
    float4 kc0 = make_float4(0.f, 0.f, 0.f, 0.f), kc1 = kc0;
    float cc = 0.f;
    for (int o = 0; o < O; ++o) {
      const int dz = g.off[o][0], dy = g.off[o][1], dx = g.off[o][2];
      const int zz = z + dz, yy = y + dy, xx = x + dx;
      const bool row_ok = zz >= 0 && zz < g.D && yy >= 0 && yy < g.H;
      if (dz != cur_dz || dy != cur_dy) {                       // new (dz,dy) row: load key / cam at the lane's own x
        cur_dz = dz; cur_dy = dy;
        if (row_ok) {
          const long long nb = b * V + ((long long)zz * g.H + yy) * g.W + x;
          const float4* kp = reinterpret_cast<const float4*>(qk + nb * 16 + 8);
          kc0 = __ldg(kp); kc1 = __ldg(kp + 1);
          cc = __ldg(cam + nb);
        }
      }
      // neighbour at x+dx of that row: own registers, an adjacent lane, or (edges) a direct load
      float4 k0 = kc0, k1 = kc1;
      float cv = cc;
      if (dx != 0) {                                             // uniform across the warp (same offset table)
        float4 s0, s1;
        float sc;
        if (dx < 0) {
          s0.x = __shfl_up_sync(0xffffffffu, kc0.x, 1); s0.y = __shfl_up_sync(0xffffffffu, kc0.y, 1);
          s0.z = __shfl_up_sync(0xffffffffu, kc0.z, 1); s0.w = __shfl_up_sync(0xffffffffu, kc0.w, 1);
          s1.x = __shfl_up_sync(0xffffffffu, kc1.x, 1); s1.y = __shfl_up_sync(0xffffffffu, kc1.y, 1);
          s1.z = __shfl_up_sync(0xffffffffu, kc1.z, 1); s1.w = __shfl_up_sync(0xffffffffu, kc1.w, 1);
          sc = __shfl_up_sync(0xffffffffu, cc, 1);
        } else {
          s0.x = __shfl_down_sync(0xffffffffu, kc0.x, 1); s0.y = __shfl_down_sync(0xffffffffu, kc0.y, 1);
          s0.z = __shfl_down_sync(0xffffffffu, kc0.z, 1); s0.w = __shfl_down_sync(0xffffffffu, kc0.w, 1);
          s1.x = __shfl_down_sync(0xffffffffu, kc1.x, 1); s1.y = __shfl_down_sync(0xffffffffu, kc1.y, 1);
          s1.z = __shfl_down_sync(0xffffffffu, kc1.z, 1); s1.w = __shfl_down_sync(0xffffffffu, kc1.w, 1);
          sc = __shfl_down_sync(0xffffffffu, cc, 1);
        }
        const bool via_lane = dx < 0 ? left_ok : right_ok;
        if (via_lane) { k0 = s0; k1 = s1; cv = sc; }
        else if (row_ok && xx >= 0 && xx < g.W) {
          const long long nb = b * V + ((long long)zz * g.H + yy) * g.W + xx;
          const float4* kp = reinterpret_cast<const float4*>(qk + nb * 16 + 8);
          k0 = __ldg(kp); k1 = __ldg(kp + 1);
          cv = __ldg(cam + nb);
        }
      }
      if (row_ok && xx >= 0 && xx < g.W) {
        float d = q0.x * k0.x + q0.y * k0.y + q0.z * k0.z + q0.w * k0.w + q1.x * k1.x + q1.y * k1.y + q1.z * k1.z + q1.w * k1.w;
        if (g.flags & 1) d = fmaxf(d, 0.f);
        const float sv = d * invT;
        const float mn = fmaxf(m, sv);
        const float corr = __expf(m - mn), e = __expf(sv - mn);     // m = -inf on the first neighbour -> corr = 0
        l = l * corr + e;
        acc = acc * corr + e * cv;
        m = mn;
      }
    }
    if (live) out[i] = deg > 0 ? acc / l : 0.f;
  }
}

// backward pass 1 (per node x): dd[x][o] = d loss / d <q_x,k_{x+o}>,  dq[x] = sum_o dd_o k_{x+o}
__global__ void __launch_bounds__(256)
k_pcm_bwd_node(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ cam, const float* __restrict__ att,
               const float* __restrict__ dout, float* __restrict__ dd, float* __restrict__ dqk) {
  const long long V = (long long)g.D * g.H * g.W;
  const long long total = (long long)g.B * V;
  const int F = g.F, O = g.O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i % V, b = i / V;
    int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    const float go = dout[i];
    float q[kMaxF], dq[kMaxF];
#pragma unroll
    for (int j = 0; j < kMaxF; ++j) { q[j] = j < F ? qk[i * 2 * F + j] : 0.f; dq[j] = 0.f; }
    float a[kMaxOff], da[kMaxOff];
    long long nbs[kMaxOff];
    int deg = 0;
    float dot = 0.f;
    for (int o = 0; o < O; ++o) {
      a[o] = att[i * O + o];
      if (nb_index(g, z, y, x, o, nbs[o])) {
        da[o] = go * __ldg(cam + b * V + nbs[o]);
        dot = fmaf(a[o], da[o], dot);
        ++deg;
      } else {
        nbs[o] = -1;
        da[o] = 0.f;
      }
    }
    float invT = 1.f / temperature(g, deg);
    for (int o = 0; o < O; ++o) {
      float r = 0.f;
      if (nbs[o] >= 0) {
        const float* kp = qk + (b * V + nbs[o]) * 2 * F + F;
        float ds = a[o] * (da[o] - dot) * invT;
        if (g.flags & 1) {
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < kMaxF; ++j)
            if (j < F) d = fmaf(q[j], __ldg(kp + j), d);
          if (!(d > 0.f)) ds = 0.f;
        }
        r = ds;
#pragma unroll
        for (int j = 0; j < kMaxF; ++j)
          if (j < F) dq[j] = fmaf(ds, __ldg(kp + j), dq[j]);
      }
      dd[i * O + o] = r;
    }
#pragma unroll
    for (int j = 0; j < kMaxF; ++j)
      if (j < F) dqk[i * 2 * F + j] = dq[j];
  }
}

// backward pass 2 (per node y, gather over the nodes x = y - o that list y as neighbour o):
//   dcam[y] = sum_o dout[x] att[x][o],   dk[y] = sum_o dd[x][o] q[x]
__global__ void __launch_bounds__(256)
k_pcm_bwd_gather(const PcmGeom g, const float* __restrict__ qk, const float* __restrict__ att, const float* __restrict__ dd,
                 const float* __restrict__ dout, float* __restrict__ dcam, float* __restrict__ dqk) {
  const long long V = (long long)g.D * g.H * g.W;
  const long long total = (long long)g.B * V;
  const int F = g.F, O = g.O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long v = i % V, b = i / V;
    int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / ((long long)g.W * g.H));
    float dk[kMaxF];
#pragma unroll
    for (int j = 0; j < kMaxF; ++j) dk[j] = 0.f;
    float dc = 0.f;
    for (int o = 0; o < O; ++o) {
      int zz = z - g.off[o][0], yy = y - g.off[o][1], xx = x - g.off[o][2];
      if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
      long long src = b * V + ((long long)zz * g.H + yy) * g.W + xx;
      dc = fmaf(__ldg(dout + src), __ldg(att + src * O + o), dc);
      float w = __ldg(dd + src * O + o);
      const float* qp = qk + src * 2 * F;
#pragma unroll
      for (int j = 0; j < kMaxF; ++j)
        if (j < F) dk[j] = fmaf(w, __ldg(qp + j), dk[j]);
    }
    dcam[i] = dc;
#pragma unroll
    for (int j = 0; j < kMaxF; ++j)
      if (j < F) dqk[i * 2 * F + F + j] = dk[j];
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
  float* sg = sf + kPcmChunk * Cf;                // [kPcmChunk][2F]
  float* sw = sg + kPcmChunk * 2 * F;             // [2F][Cf]
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
    for (int i = threadIdx.x; i < n * 2 * F; i += blockDim.x) sg[i] = dqk[base * 2 * F + i];
    __syncthreads();
    if (owner) {
      for (int r = 0; r < n; ++r) {
        float gq = sg[r * 2 * F + which * F + j];
        acc = c >= 0 ? fmaf(gq, sf[r * Cf + c], acc) : acc + gq;
      }
    }
    for (int i = threadIdx.x; i < n * Cf; i += blockDim.x) {
      int r = i / Cf, cc = i % Cf;
      float d = 0.f;
      for (int jj = 0; jj < 2 * F; ++jj) d = fmaf(sw[jj * Cf + cc], sg[r * 2 * F + jj], d);
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
  g.B = B; g.D = D; g.H = H; g.W = W; g.Cf = Cf; g.F = F; g.flags = flags;
  g.O = build_offsets(connectivity, self_loop, g.off);
  return DRAM_OK;
}

extern "C" {

int dram_pcm_num_offsets(int connectivity, int self_loop) {
  signed char off[kMaxOff][3];
  return build_offsets(connectivity, self_loop, off);
}

int dram_pcm_fwd(const float* f, const float* cam, const float* theta_w, const float* theta_b, const float* phi_w,
                 const float* phi_b, float* qk, float* att, float* out, int B, int D, int H, int W, int Cf, int F,
                 int connectivity, int self_loop, int flags, void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && theta_b && phi_w && phi_b && qk && out, "pcm_fwd: null pointer");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  k_pcm_project<<<grid_for(rows, 256, 8), 256, 0, st>>>(f, theta_w, theta_b, phi_w, phi_b, qk, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  if (att == nullptr && F == 8) {       // inference: no attention weights kept, warp-shuffle neighbour sharing
    k_pcm_attend_infer<<<grid_for(rows, 256, 8), 256, 0, st>>>(g, qk, cam, out);
  } else {
    DRAM_REQUIRE(att != nullptr, "pcm_fwd: the attention-weight buffer may only be NULL for F == 8 (inference kernel)");
    k_pcm_attend<<<grid_for(rows, 256, 8), 256, 0, st>>>(g, qk, cam, att, out);
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_pcm_bwd(const float* f, const float* cam, const float* theta_w, const float* phi_w, const float* qk,
                 const float* att, const float* dout, float* dd_ws, float* dqk_ws, float* dcam, float* df,
                 double* dparams, int B, int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags,
                 void* stream) {
  DRAM_REQUIRE(f && cam && theta_w && phi_w && qk && att && dout && dd_ws && dqk_ws && dcam && df && dparams, "pcm_bwd: null pointer");
  PcmGeom g;
  int rc = pcm_geom(g, B, D, H, W, Cf, F, connectivity, self_loop, flags);
  if (rc) return rc;
  DRAM_REQUIRE(2 * F * (Cf + 1) <= 320, "pcm_bwd: 2F(Cf+1)=%d > 320 unsupported", 2 * F * (Cf + 1));
  cudaStream_t st = (cudaStream_t)stream;
  long long rows = (long long)B * D * H * W;
  k_pcm_bwd_node<<<grid_for(rows, 256, 8), 256, 0, st>>>(g, qk, cam, att, dout, dd_ws, dqk_ws);
  DRAM_LAUNCH_CHECK();
  k_pcm_bwd_gather<<<grid_for(rows, 256, 8), 256, 0, st>>>(g, qk, att, dd_ws, dout, dcam, dqk_ws);
  DRAM_LAUNCH_CHECK();
  DRAM_CUDA(cudaMemsetAsync(dparams, 0, sizeof(double) * 2 * F * (Cf + 1), st));
  size_t smem = sizeof(float) * ((size_t)kPcmChunk * Cf + (size_t)kPcmChunk * 2 * F + (size_t)2 * F * Cf);
  long long chunks = (rows + kPcmChunk - 1) / kPcmChunk;
  int grid = (int)(chunks < (long long)kNumSMs * 4 ? chunks : (long long)kNumSMs * 4);
  if (smem > 48 * 1024) DRAM_CUDA(cudaFuncSetAttribute(k_pcm_bwd_params, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pcm_bwd_params<<<grid, 320, smem, st>>>(f, theta_w, phi_w, dqk_ws, df, dparams, rows, Cf, F);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
