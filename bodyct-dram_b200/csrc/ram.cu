// Dense regression-activation-map (RAM) head: channel reduce with the regression weights (models.py:109-110,145),
// per-lobe masked pooling (models.py:37-49, metrics.py:160-165) and the inference epilogue that upsamples a chunk's
// RAM to the lobe crop and pastes it into the scan-sized heat map under the lobe mask (job_runner.py:765-770,993-1004).
// All kernels are HBM-bound: algorithmic bytes = one read of the features (C*4 B/voxel) + one write of the map.
#include "common.cuh"

namespace dram {

// feat [rows][C] (C % 4 == 0, C <= 128*... ), w [O][C].  A group of G = C/4 lanes owns one row (float4 per lane),
// 32/G rows per warp per iteration, shuffle-reduced inside the group.  Fully coalesced 512 B per warp load.
template <int G>
__global__ void __launch_bounds__(256)
k_ram_reduce_fwd(const float* __restrict__ feat, const float* __restrict__ scale, const float* __restrict__ shift,
                 const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ ram, long long rows,
                 int C, int O) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, rsub = lane / G;
  constexpr int RPW = 32 / G;                      // rows per warp-iteration
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool fused = scale != nullptr;
  if (fused) { sc = reinterpret_cast<const float4*>(scale)[sub]; sh = reinterpret_cast<const float4*>(shift)[sub]; }
  constexpr int U = 4;                             // independent row batches in flight per warp (pure HBM stream)
  for (long long r0 = warp * RPW * U; r0 < rows; r0 += nwarps * RPW * U) {
    float4 t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * RPW + rsub;
      t[u] = r < rows ? __ldg(reinterpret_cast<const float4*>(feat + r * C) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * RPW + rsub;
      if (fused) {
        t[u].x = fmaxf(t[u].x * sc.x + sh.x, 0.f); t[u].y = fmaxf(t[u].y * sc.y + sh.y, 0.f);
        t[u].z = fmaxf(t[u].z * sc.z + sh.z, 0.f); t[u].w = fmaxf(t[u].w * sc.w + sh.w, 0.f);
      }
      for (int o = 0; o < O; ++o) {
        float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)o * C) + sub);
        float p = t[u].x * wv.x + t[u].y * wv.y + t[u].z * wv.z + t[u].w * wv.w;
#pragma unroll
        for (int sft = G / 2; sft > 0; sft >>= 1) p += __shfl_xor_sync(0xffffffffu, p, sft);
        if (sub == 0 && r < rows) ram[r * O + o] = p + b[o];
      }
    }
  }
}

// generic (any C): one thread per (row, o)
__global__ void k_ram_reduce_fwd_generic(const float* __restrict__ feat, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ w,
                                         const float* __restrict__ b, float* __restrict__ ram, long long rows, int C, int O) {
  long long total = rows * O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / O;
    int o = (int)(i % O);
    float acc = b[o];
    for (int c = 0; c < C; ++c) {
      float t = feat[r * C + c];
      if (scale) t = fmaxf(t * scale[c] + shift[c], 0.f);
      acc = fmaf(t, w[(long long)o * C + c], acc);
    }
    ram[i] = acc;
  }
}

// dfeat[m][c] = sum_o dram[m][o] w[o][c];  dwb[o*C+c] += sum_m dram[m][o] feat[m][c];  dwb[O*C+o] += sum_m dram[m][o]
// One thread per channel c (blockDim = C rounded up), blocks stride over row slabs.
constexpr int kRamSlab = 1024;
__global__ void k_ram_reduce_bwd(const float* __restrict__ dram, const float* __restrict__ feat,
                                 const float* __restrict__ w, float* __restrict__ dfeat, double* __restrict__ dwb,
                                 long long rows, int C, int O) {
  const int c = threadIdx.x % C;                    // blockDim.x is a multiple of C
  const int rl = threadIdx.x / C, rlanes = blockDim.x / C;
  extern __shared__ float red[];                    // [blockDim.x]
  for (int o = 0; o < O; ++o) {
    const float wv = w[(long long)o * C + c];
    float accw = 0.f, accb = 0.f;
    for (long long slab = (long long)blockIdx.x * kRamSlab; slab < rows; slab += (long long)gridDim.x * kRamSlab) {
      long long end = slab + kRamSlab < rows ? slab + kRamSlab : rows;
      for (long long r = slab + rl; r < end; r += rlanes) {
        float g = dram[r * O + o];
        float f = feat[r * C + c];
        accw = fmaf(g, f, accw);
        accb += g;
        if (O == 1) dfeat[r * C + c] = g * wv;
      }
    }
    red[threadIdx.x] = accw;
    __syncthreads();
    if (rl == 0) {
      double t = 0.0;
      for (int l = 0; l < rlanes; ++l) t += (double)red[l * C + c];
      atomicAdd(&dwb[(long long)o * C + c], t);
    }
    __syncthreads();
    if (c == 0) red[rl] = accb;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int l = 0; l < rlanes; ++l) t += (double)red[l];
      atomicAdd(&dwb[(long long)O * C + o], t);
    }
    __syncthreads();
  }
}
__global__ void k_ram_dfeat_multi(const float* __restrict__ dram, const float* __restrict__ w, float* __restrict__ dfeat,
                                  long long rows, int C, int O) {
  long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / C;
    int c = (int)(i % C);
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc = fmaf(dram[r * O + o], w[(long long)o * C + c], acc);
    dfeat[i] = acc;
  }
}

// masked pooling: per sample b: sum f(x)*m and sum m (float4 loads when the volume size allows)
template <int VEC>
__global__ void __launch_bounds__(256)
k_masked_pool_fwd(const float* __restrict__ x, const float* __restrict__ mask, double* __restrict__ out, long long V,
                  int use_sigmoid, int mode_gt0) {
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * V;
  const float* mb = mask + (long long)b * V;
  float s = 0.f, cnt = 0.f;
  const long long n = V / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float xv[VEC], mv[VEC];
    if (VEC == 4) {
      float4 a = __ldg(reinterpret_cast<const float4*>(xb) + i), m4 = __ldg(reinterpret_cast<const float4*>(mb) + i);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3 % VEC] = a.w;
      mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3 % VEC] = m4.w;
    } else {
      xv[0] = xb[i]; mv[0] = mb[i];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float m = mv[v];
      if (mode_gt0) m = m > 0.f ? 1.f : 0.f;
      float val = use_sigmoid ? sigmoidf_(xv[v]) : xv[v];
      s = fmaf(val, m, s);
      cnt += m;
    }
  }
  s = warp_sum(s);
  cnt = warp_sum(cnt);
  __shared__ float sh[2][8];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][wid] = s; sh[1][wid] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += (double)sh[0][i]; c += (double)sh[1][i]; }
    atomicAdd(&out[2 * b], a);
    atomicAdd(&out[2 * b + 1], c);
  }
}

__global__ void k_masked_pool_bwd(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ g,
                                  float* __restrict__ dx, long long V, int use_sigmoid, int mode_gt0) {
  const int b = blockIdx.y;
  const float gb = g[b];
  const float* xb = x + (long long)b * V;
  const float* mb = mask + (long long)b * V;
  float* db = dx + (long long)b * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    float m = mb[i];
    if (mode_gt0) m = m > 0.f ? 1.f : 0.f;
    float d = gb * m;
    if (use_sigmoid) { float s = sigmoidf_(xb[i]); d *= s * (1.f - s); }
    db[i] = d;
  }
}

// interval-regression term from the pooled sums (metrics.py:121-137, 158-177): one thread per sample, thread 0 adds the B
// terms in index order (deterministic); arithmetic follows the reference's dtypes step by step (fp32 tensors, the interval in
// Python floats = double, rounded back to fp32)
__global__ void k_int_reg_loss(const double* __restrict__ pp, const double* __restrict__ pr, const double* __restrict__ band,
                               const float* __restrict__ w, double bw, float* __restrict__ loss, float* __restrict__ g, int B) {
  extern __shared__ float terms[];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float pred = (float)(pp[2 * b] / pp[2 * b + 1]);
    const double p = (double)(float)(pr[2 * b] / pr[2 * b + 1]);
    const double clb = band[2 * b], cub = band[2 * b + 1];
    const double lb = fmax(p - bw, 0.0), ub = fmin(p + bw, 1.0);
    double lo = fmax(clb, lb), hi = fmin(cub, ub);
    const bool empty = hi < lo, below = empty && (ub <= clb);
    lo = below ? lb : (empty ? clb : lo);
    hi = below ? ub : (empty ? cub : hi);
    const float t0 = (float)lo, t1 = (float)hi;
    const float half = 0.5f * (t1 - t0), K = half * half, mid = (t1 + t0) / 2.0f;
    const float dlt = pred - mid, unh = dlt * dlt - K;
    terms[b] = fmaxf(unh, 0.f) / w[b];
    const float dpred = unh > 0.f ? 2.f * dlt / w[b] : 0.f;
    g[b] = (float)((double)dpred / pp[2 * b + 1]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += terms[b];
    loss[0] = acc;
  }
}

__global__ void k_ram_activation(const float* __restrict__ in, float* __restrict__ out, long long n, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = in[i];
    out[i] = act == 1 ? sigmoidf_(v) : (act == 2 ? fmaxf(v, 0.f) : v);
  }
}

// inference epilogue: one thread per crop voxel, x fastest (coalesced mask read / heat write)
__device__ __forceinline__ float atomic_max_pos(float* addr, float v) {  // v >= 0
  return __int_as_float(atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)));
}
__global__ void __launch_bounds__(256)
k_ram_upsample_mask_scatter(const float* __restrict__ ram, const uint8_t* __restrict__ crop_mask, float* __restrict__ heat,
                            float* maxval, int d, int h, int w, int cd, int ch, int cw, int SD, int SH, int SW, int oz,
                            int oy, int ox, int act, float gain, float sz, float sy, float sx) {
  const long long total = (long long)cd * ch * cw;
  float local_max = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int X = (int)(i % cw);
    long long r = i / cw;
    int Y = (int)(r % ch);
    int Z = (int)(r / ch);
    const bool inside = crop_mask[i] != 0;
    if (!inside && maxval == nullptr) continue;
    Lerp lz = lerp_setup(Z, sz, d), ly = lerp_setup(Y, sy, h), lx = lerp_setup(X, sx, w);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int zi = (k & 4) ? lz.i1 : lz.i0, yi = (k & 2) ? ly.i1 : ly.i0, xi = (k & 1) ? lx.i1 : lx.i0;
      float wt = ((k & 4) ? lz.w1 : lz.w0) * ((k & 2) ? ly.w1 : ly.w0) * ((k & 1) ? lx.w1 : lx.w0);
      if (act == 1) acc += wt * sigmoidf_(__ldg(ram + ((long long)zi * h + yi) * w + xi));   // sigmoid BEFORE interpolation
      else acc += wt * __ldg(ram + ((long long)zi * h + yi) * w + xi);
    }
    if (act == 2) acc = fmaxf(acc, 0.f);                                                         // relu AFTER interpolation
    if (maxval) local_max = fmaxf(local_max, acc);
    if (inside && heat) heat[((long long)(Z + oz) * SH + (Y + oy)) * SW + (X + ox)] = acc * gain;
  }
  if (maxval) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0 && local_max > 0.f) atomic_max_pos(maxval, local_max);
  }
}

// ------------------------------------------------------------------------------------------------ bootstrapped BCE (seg loss)
// metrics.py:10-51 (BootBinCrossEntropy) + :325-354 (thresholded pseudo labels) as ONE reduction pass and ONE gradient
// pass over [B][V]: pseudo label t = (sigmoid(dense) > 0.5) & (lobe != 0) & (lesion > 0), zeroed where keep[b] == 0;
// p = sigmoid(refined); inside = lobe > 0.  The seven sums below are everything the loss needs:
//   s[0] = #outside, s[1] = #inside, s[2] = sum t*inside,
//   s[3] = sum nll*outside, s[4] = sum nll*t*inside, s[5] = sum nll*(1-t)*inside, s[6] = sum -log(pt_hat)*inside
// with nll = -log(clamp(p*t + (1-p)*(1-t), eps, 1-eps)) and pt_hat the same with t_hat = (p > 0.5).
struct BceVoxel {
  float inside, t, p, pt, pt_hat, that;
};
__device__ __forceinline__ BceVoxel bce_voxel(float dense, float refined, float lobe, float lesion, float keep) {
  BceVoxel o;
  o.inside = lobe > 0.f ? 1.f : 0.f;
  const float pd = 1.f / (1.f + expf(-dense));
  o.t = (pd > 0.5f && lobe != 0.f && lesion > 0.f) ? keep : 0.f;
  o.p = 1.f / (1.f + expf(-refined));
  o.pt = o.p * o.t + (1.f - o.p) * (1.f - o.t);
  o.that = o.p > 0.5f ? 1.f : 0.f;
  o.pt_hat = o.p * o.that + (1.f - o.p) * (1.f - o.that);
  return o;
}

__global__ void __launch_bounds__(256)
k_boot_bce_fwd(const float* __restrict__ dense, const float* __restrict__ refined, const float* __restrict__ lobes,
               const float* __restrict__ lesions, const float* __restrict__ keep, double* __restrict__ sums, long long V,
               float eps) {
  const int b = blockIdx.y;
  const float kp = keep[b];
  const long long base = (long long)b * V;
  float a[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const BceVoxel x = bce_voxel(__ldg(dense + base + v), __ldg(refined + base + v), __ldg(lobes + base + v),
                                 __ldg(lesions + base + v), kp);
    const float nll = -logf(fminf(fmaxf(x.pt, eps), 1.f - eps));
    const float nllh = -logf(fminf(fmaxf(x.pt_hat, eps), 1.f - eps));
    a[0] += 1.f - x.inside; a[1] += x.inside; a[2] += x.t * x.inside;
    a[3] += nll * (1.f - x.inside); a[4] += nll * x.t * x.inside; a[5] += nll * (1.f - x.t) * x.inside;
    a[6] += nllh * x.inside;
  }
  __shared__ float red[7][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float w = warp_sum(a[j]);
    if (lane == 0) red[j][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[threadIdx.x][w];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

// d loss / d refined: coef = (c_out, c_t, c_nt, c_boot), already multiplied by the upstream gradient (device floats):
//   loss = c_out * s[3] + c_t * s[4] + c_nt * s[5] + c_boot * s[6];  clamp passes gradient on [eps, 1-eps] (ATen).
__global__ void __launch_bounds__(256)
k_boot_bce_bwd(const float* __restrict__ dense, const float* __restrict__ refined, const float* __restrict__ lobes,
               const float* __restrict__ lesions, const float* __restrict__ keep, const float* __restrict__ coef,
               float* __restrict__ dref, long long V, float eps) {
  const int b = blockIdx.y;
  const float kp = keep[b];
  const float c_out = coef[0], c_t = coef[1], c_nt = coef[2], c_b = coef[3];
  const long long base = (long long)b * V;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
    const BceVoxel x = bce_voxel(__ldg(dense + base + v), __ldg(refined + base + v), __ldg(lobes + base + v),
                                 __ldg(lesions + base + v), kp);
    const float wv = c_out * (1.f - x.inside) + (c_t * x.t + c_nt * (1.f - x.t)) * x.inside;
    float g = 0.f;
    if (x.pt >= eps && x.pt <= 1.f - eps) g += wv * (-(2.f * x.t - 1.f) / x.pt);
    if (x.pt_hat >= eps && x.pt_hat <= 1.f - eps) g += c_b * x.inside * (-(2.f * x.that - 1.f) / x.pt_hat);
    dref[base + v] = g * x.p * (1.f - x.p);
  }
}

}  // namespace dram

using namespace dram;

extern "C" {

int dram_ram_reduce_fwd(const float* feat, const float* scale, const float* shift, const float* w, const float* b,
                        float* ram, long long rows, int C, int O, void* stream) {
  DRAM_REQUIRE(feat && w && b && ram && rows > 0 && C > 0 && O > 0, "ram_reduce_fwd: bad arguments");
  DRAM_REQUIRE((scale == nullptr) == (shift == nullptr), "ram_reduce_fwd: scale and shift must come together");
  cudaStream_t st = (cudaStream_t)stream;
  const int G = C / 4;
  if (C % 4 == 0 && (G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32)) {
    long long warps = (rows + (32 / G) - 1) / (32 / G);
    int grid = grid_for(warps * 32, 256, 16);
    switch (G) {
      case 1: k_ram_reduce_fwd<1><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
      case 2: k_ram_reduce_fwd<2><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
      case 4: k_ram_reduce_fwd<4><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
      case 8: k_ram_reduce_fwd<8><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
      case 16: k_ram_reduce_fwd<16><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
      default: k_ram_reduce_fwd<32><<<grid, 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O); break;
    }
  } else {
    k_ram_reduce_fwd_generic<<<grid_for(rows * O, 256), 256, 0, st>>>(feat, scale, shift, w, b, ram, rows, C, O);
  }
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_ram_reduce_bwd(const float* dram_, const float* feat, const float* w, float* dfeat, double* dwb, long long rows,
                        int C, int O, void* stream) {
  DRAM_REQUIRE(dram_ && feat && w && dfeat && dwb && rows > 0 && C > 0 && O > 0, "ram_reduce_bwd: bad arguments");
  DRAM_REQUIRE(C <= 1024, "ram_reduce_bwd: C=%d > 1024 unsupported", C);
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(dwb, 0, sizeof(double) * ((size_t)O * C + O), st));
  int rl = 256 / C; if (rl < 1) rl = 1;
  int block = rl * C;
  long long slabs = (rows + kRamSlab - 1) / kRamSlab;
  int grid = (int)(slabs < (long long)kNumSMs * 8 ? slabs : (long long)kNumSMs * 8);
  k_ram_reduce_bwd<<<grid, block, block * sizeof(float), st>>>(dram_, feat, w, dfeat, dwb, rows, C, O);
  DRAM_LAUNCH_CHECK();
  if (O > 1) {
    k_ram_dfeat_multi<<<grid_for(rows * C, 256), 256, 0, st>>>(dram_, w, dfeat, rows, C, O);
    DRAM_LAUNCH_CHECK();
  }
  return DRAM_OK;
}

int dram_masked_pool_fwd(const float* x, const float* mask, double* out, int B, long long V, int use_sigmoid,
                         int mode_gt0, void* stream) {
  DRAM_REQUIRE(x && mask && out && B > 0 && V > 0, "masked_pool_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * B, st));
  const bool vec = (V % 4 == 0) && (((uintptr_t)x | (uintptr_t)mask) % 16 == 0);
  long long per_sample = (long long)kNumSMs * 4 / B;
  if (per_sample < 1) per_sample = 1;
  long long need = (V / (vec ? 4 : 1) + 255) / 256;
  int gx = (int)(need < per_sample ? (need < 1 ? 1 : need) : per_sample);
  if (vec) k_masked_pool_fwd<4><<<dim3(gx, B), 256, 0, st>>>(x, mask, out, V, use_sigmoid, mode_gt0);
  else k_masked_pool_fwd<1><<<dim3(gx, B), 256, 0, st>>>(x, mask, out, V, use_sigmoid, mode_gt0);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_masked_pool_bwd(const float* x, const float* mask, const float* g, float* dx, int B, long long V,
                         int use_sigmoid, int mode_gt0, void* stream) {
  DRAM_REQUIRE(x && mask && g && dx && B > 0 && V > 0, "masked_pool_bwd: bad arguments");
  int gx = grid_for(V, 256 * 4, 4);
  k_masked_pool_bwd<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(x, mask, g, dx, V, use_sigmoid, mode_gt0);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_int_reg_loss(const double* pool_pred, const double* pool_rub, const double* band, const float* w, double band_width,
                      float* loss, float* g, int B, void* stream) {
  DRAM_REQUIRE(pool_pred && pool_rub && band && w && loss && g && B > 0 && B <= 8192, "int_reg_loss: bad arguments (B=%d)", B);
  k_int_reg_loss<<<1, 128, sizeof(float) * B, (cudaStream_t)stream>>>(pool_pred, pool_rub, band, w, band_width, loss, g, B);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_ram_activation(const float* in, float* out, long long n, int act, void* stream) {
  DRAM_REQUIRE(in && out && n > 0 && act >= 0 && act <= 2, "ram_activation: bad arguments");
  k_ram_activation<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n, act);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_ram_upsample_mask_scatter(const float* ram, const uint8_t* crop_mask, float* heat, float* maxval, int d, int h,
                                   int w, int cd, int ch, int cw, int SD, int SH, int SW, int oz, int oy, int ox, int act,
                                   float gain, void* stream) {
  DRAM_REQUIRE(ram && crop_mask && (heat || maxval), "ram_upsample_mask_scatter: bad pointers");
  DRAM_REQUIRE(d > 0 && h > 0 && w > 0 && cd > 0 && ch > 0 && cw > 0, "ram_upsample_mask_scatter: bad sizes");
  DRAM_REQUIRE(act >= 0 && act <= 2, "ram_upsample_mask_scatter: act %d unknown", act);
  DRAM_REQUIRE(oz >= 0 && oy >= 0 && ox >= 0 && oz + cd <= SD && oy + ch <= SH && ox + cw <= SW,
               "ram_upsample_mask_scatter: crop [%d+%d,%d+%d,%d+%d] outside scan [%d,%d,%d]", oz, cd, oy, ch, ox, cw, SD, SH, SW);
  long long total = (long long)cd * ch * cw;
  k_ram_upsample_mask_scatter<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      ram, crop_mask, heat, maxval, d, h, w, cd, ch, cw, SD, SH, SW, oz, oy, ox, act, gain, ac_scale(d, cd),
      ac_scale(h, ch), ac_scale(w, cw));
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_boot_bce_fwd(const float* dense, const float* refined, const float* lobes, const float* lesions, const float* keep,
                      double* sums, int B, long long V, float eps, void* stream) {
  DRAM_REQUIRE(dense && refined && lobes && lesions && keep && sums && B > 0 && V > 0, "boot_bce_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 7, st));
  long long per_sample = (long long)kNumSMs * 4 / B;
  if (per_sample < 1) per_sample = 1;
  const long long need = (V + 255) / 256;
  const int gx = (int)(need < per_sample ? need : per_sample);
  k_boot_bce_fwd<<<dim3(gx, B), 256, 0, st>>>(dense, refined, lobes, lesions, keep, sums, V, eps);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_boot_bce_bwd(const float* dense, const float* refined, const float* lobes, const float* lesions, const float* keep,
                      const float* coef, float* drefined, int B, long long V, float eps, void* stream) {
  DRAM_REQUIRE(dense && refined && lobes && lesions && keep && coef && drefined && B > 0 && V > 0, "boot_bce_bwd: bad arguments");
  const int gx = grid_for(V, 256, 4);
  k_boot_bce_bwd<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(dense, refined, lobes, lesions, keep, coef, drefined, V, eps);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
