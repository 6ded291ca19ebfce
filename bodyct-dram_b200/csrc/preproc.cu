// Scan-level pre/post-processing around the chunk model: everything LesionSegTest.run does on the
// host with numpy / SimpleITK / skimage between reading a scan and writing a lesion mask (job_runner.py:951-1015),
// as HBM-bound kernels so that a full CT scan never leaves the GPU between upload and mask download.
#include "common.cuh"

namespace dram {

// ------------------------------------------------------------------------------------------------ lobe bounding boxes
// utils.find_crops (utils.py:244-254) for all labels at once: out[l][0..2] = min z,y,x; out[l][3..5] = max z,y,x (inclusive)
// A thread reads 4 labels as one aligned word of the flat volume and skips all-background words (most of a chest CT);
// only words that hold a label pay for coordinates (one 32-bit division pair per word) and update the thread's own
// register boxes (the label loop is unrolled: no dynamic register indexing); shared / global atomics once per thread / block.
template <int NL>
__global__ void __launch_bounds__(256)
k_label_bboxes(const uint8_t* __restrict__ labels, int D, int H, int W, int nlabels, int* __restrict__ out) {
  __shared__ int smin[8][3], smax[8][3];                       // labels 1..7 supported per launch
  for (int i = threadIdx.x; i < 8 * 3; i += blockDim.x) { (&smin[0][0])[i] = 0x7fffffff; (&smax[0][0])[i] = -1; }
  __syncthreads();
  int mn[NL][3], mx[NL][3];
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[l][a] = 0x7fffffff; mx[l][a] = -1; }
  const unsigned total = (unsigned)D * H * W, HW = (unsigned)H * W, words = (total + 3) >> 2;
  const unsigned* lw = reinterpret_cast<const unsigned*>(labels);
  for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < words; g += gridDim.x * blockDim.x) {
    const unsigned i0 = g << 2;
    unsigned word;
    if (i0 + 3 < total) word = __ldg(lw + g);
    else {
      word = 0;
      for (int j = 0; j < 4 && i0 + j < total; ++j) word |= (unsigned)labels[i0 + j] << (8 * j);
    }
    if (word == 0) continue;
    int z = (int)(i0 / HW);
    const unsigned rem = i0 - (unsigned)z * HW;
    int y = (int)(rem / (unsigned)W), x = (int)(rem - (unsigned)y * W);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lab = (int)((word >> (8 * j)) & 0xffu);
#pragma unroll
      for (int l = 0; l < NL; ++l)
        if (lab == l + 1 && l < nlabels) {
          mn[l][0] = min(mn[l][0], z); mn[l][1] = min(mn[l][1], y); mn[l][2] = min(mn[l][2], x);
          mx[l][0] = max(mx[l][0], z); mx[l][1] = max(mx[l][1], y); mx[l][2] = max(mx[l][2], x);
        }
      if (++x == W) { x = 0; if (++y == H) { y = 0; ++z; } }
    }
  }
#pragma unroll
  for (int l = 0; l < NL; ++l)
    if (mx[l][0] >= 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { atomicMin(&smin[l + 1][a], mn[l][a]); atomicMax(&smax[l + 1][a], mx[l][a]); }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 3; i += blockDim.x) {
    int l = i / 3, a = i % 3;
    if (l >= 1 && l <= nlabels) {
      if (smax[l][a] >= 0) { atomicMin(&out[l * 6 + a], smin[l][a]); atomicMax(&out[l * 6 + 3 + a], smax[l][a]); }
    }
  }
}
__global__ void k_bbox_init(int* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (i % 6) < 3 ? 0x7fffffff : -1;
}

// ------------------------------------------------------------------------------------------------ ITK-style resampling
// SimpleITK ResampleImageFilter with identity transform / identity direction / shared origin (utils.py:414-434,
// data_transforms.py:170-175), computed the way ITK 4.13 computes it — in DOUBLE, operation by operation, so that integer
// and float32 outputs are bit-identical to the float64 oracle (oracle/dram_oracle.py::itk_resample, which restates the ITK
// classes, not this kernel):
//   coordinates   itk::ImageBase::TransformIndexToPhysicalPoint / TransformPhysicalPointToContinuousIndex:
//                   cidx = (1 / spacing_in) * (spacing_out * index)            (y, z)
//                 itk::ResampleImageFilter::LinearThreadedGenerateData, along a scanline of `size` pixels (x):
//                   alpha = i / double(size); cidx = start + alpha * (end - start), start/end = mapped indices 0 and `size`
//   inside test   itk::InterpolateImageFunction::IsInsideBuffer: -0.5 <= cidx < n - 0.5, else the default value 0
//   linear        itk::LinearInterpolateImageFunction::EvaluateOptimized: base = floor(cidx) clamped to 0, distance = cidx - base,
//                 neighbours beyond the last index fall back to the base sample, nested lerps a + (b - a) * d over x, y, z
//   nearest       itk::NearestNeighborInterpolateImageFunction: floor(cidx + 0.5)
//   output cast   ResampleImageFilter::CastPixelWithBoundsChecking: clamp + static_cast (truncation for integers)
// Every double operation is an explicit round-to-nearest intrinsic: the compiler must not contract a*b+c into an FMA.
struct AxisD { int i0, i1; double dist; bool inside; };
__device__ __forceinline__ AxisD itk_axis(double c, int n) {
  AxisD a;
  a.inside = (c >= -0.5) && (c < (double)n - 0.5);
  int b = (int)floor(c);
  if (b < 0) b = 0;
  double dist = __dsub_rn(c, (double)b);
  a.dist = dist > 0.0 ? dist : 0.0;
  a.i0 = min(b, n - 1);
  a.i1 = min(b + 1, n - 1);
  return a;
}
__device__ __forceinline__ int itk_nearest(double c, int n, bool& inside) {
  inside = inside && (c >= -0.5) && (c < (double)n - 0.5);
  const int i = (int)floor(__dadd_rn(c, 0.5));
  return min(max(i, 0), n - 1);
}
// y / z: the mapped index itself; x: ITK's per-scanline start/end interpolation
__device__ __forceinline__ double itk_cidx(int i, double inv_in, double sp_out) { return __dmul_rn(inv_in, __dmul_rn(sp_out, (double)i)); }
__device__ __forceinline__ double itk_cidx_line(int i, int size, double inv_in, double sp_out) {
  const double start = __dmul_rn(inv_in, __dmul_rn(sp_out, 0.0));
  const double end = __dmul_rn(inv_in, __dmul_rn(sp_out, (double)size));
  const double alpha = __ddiv_rn((double)i, (double)size);
  return __dadd_rn(start, __dmul_rn(alpha, __dsub_rn(end, start)));
}
__device__ __forceinline__ double itk_lerp(double a, double b, double d) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), d)); }

struct ItkGeom { double inv_z, inv_y, inv_x, out_z, out_y, out_x; };   // 1 / input spacing, output spacing (z, y, x)

// one lobe chunk: crop [cz..cz+cd) x ... of the scan, blank voxels outside `label` to pad_value, window to [0,1] in float32
// (Windowing, data_transforms.py:37-54: numpy float32 arithmetic), linear-resample to (d,h,w); mask = nearest-resampled
// (labels == label) as float {0,1}   (job_runner.py:961-984)
__global__ void __launch_bounds__(256)
k_lobe_chunk_preprocess(const short* __restrict__ scan, const uint8_t* __restrict__ labels, int SH, int SW, int label,
                        int cz, int cy, int cx, int cd, int ch, int cw, float win_lo, float win_hi, float pad_value,
                        float* __restrict__ img, float* __restrict__ msk, int d, int h, int w, ItkGeom g) {
  const long long total = (long long)d * h * w;
  const float span = win_hi - win_lo;              // numpy divides (utils.py:197), keep the same rounding
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % w), Y = (int)((i / w) % h), Z = (int)(i / ((long long)w * h));
    const double c_z = itk_cidx(Z, g.inv_z, g.out_z), c_y = itk_cidx(Y, g.inv_y, g.out_y), c_x = itk_cidx_line(X, w, g.inv_x, g.out_x);
    const AxisD az = itk_axis(c_z, cd), ay = itk_axis(c_y, ch), ax = itk_axis(c_x, cw);
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zi = ((k & 4) ? az.i1 : az.i0) + cz, yi = ((k & 2) ? ay.i1 : ay.i0) + cy, xi = ((k & 1) ? ax.i1 : ax.i0) + cx;
      const long long off = ((long long)zi * SH + yi) * SW + xi;
      float t = (labels[off] == label) ? (float)scan[off] : pad_value;
      t = fminf(fmaxf(t, win_lo), win_hi);
      v[k] = (double)__fdiv_rn(__fsub_rn(t, win_lo), span);
    }
    const double lo = itk_lerp(itk_lerp(v[0], v[1], ax.dist), itk_lerp(v[2], v[3], ax.dist), ay.dist);
    const double hi = itk_lerp(itk_lerp(v[4], v[5], ax.dist), itk_lerp(v[6], v[7], ax.dist), ay.dist);
    bool in = az.inside && ay.inside && ax.inside;
    img[i] = in ? (float)itk_lerp(lo, hi, az.dist) : 0.f;
    bool inn = true;
    const int zn = itk_nearest(c_z, cd, inn) + cz, yn = itk_nearest(c_y, ch, inn) + cy, xn = itk_nearest(c_x, cw, inn) + cx;
    msk[i] = (inn && labels[((long long)zn * SH + yn) * SW + xn] == label) ? 1.f : 0.f;
  }
}

// generic volume resample (scan <-> working grid): T in {short, uint8, float}; linear (mode 0) or nearest (mode 1)
template <typename T>
__device__ __forceinline__ T cast_out(double v);
template <> __device__ __forceinline__ float cast_out<float>(double v) { return (float)v; }                       // round to nearest
template <> __device__ __forceinline__ short cast_out<short>(double v) { return (short)(int)fmin(fmax(v, -32768.0), 32767.0); }   // trunc
template <> __device__ __forceinline__ uint8_t cast_out<uint8_t>(double v) { return (uint8_t)(int)fmin(fmax(v, 0.0), 255.0); }

template <typename T>
__global__ void __launch_bounds__(256)
k_itk_resample(const T* __restrict__ src, T* __restrict__ dst, int d, int h, int w, int D, int H, int W, ItkGeom g, int mode) {
  const long long total = (long long)D * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W), Y = (int)((i / W) % H), Z = (int)(i / ((long long)W * H));
    const double c_z = itk_cidx(Z, g.inv_z, g.out_z), c_y = itk_cidx(Y, g.inv_y, g.out_y), c_x = itk_cidx_line(X, W, g.inv_x, g.out_x);
    if (mode == 1) {
      bool in = true;
      const int zn = itk_nearest(c_z, d, in), yn = itk_nearest(c_y, h, in), xn = itk_nearest(c_x, w, in);
      dst[i] = in ? src[((long long)zn * h + yn) * w + xn] : (T)0;
    } else {
      const AxisD az = itk_axis(c_z, d), ay = itk_axis(c_y, h), ax = itk_axis(c_x, w);
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = (double)src[((long long)((k & 4) ? az.i1 : az.i0) * h + ((k & 2) ? ay.i1 : ay.i0)) * w + ((k & 1) ? ax.i1 : ax.i0)];
      const double lo = itk_lerp(itk_lerp(v[0], v[1], ax.dist), itk_lerp(v[2], v[3], ax.dist), ay.dist);
      const double hi = itk_lerp(itk_lerp(v[4], v[5], ax.dist), itk_lerp(v[6], v[7], ax.dist), ay.dist);
      dst[i] = (az.inside && ay.inside && ax.inside) ? cast_out<T>(itk_lerp(lo, hi, az.dist)) : (T)0;
    }
  }
}

// Same arithmetic, 4 consecutive output voxels of the FLAT output array per thread (an aligned 4-element store; the
// group may straddle a row end, so the row set-up is redone when x wraps): one 32-bit division pair per 4 outputs, the
// z / y part of the interpolation set-up shared by the group.  Used when the output has < 2^31 voxels and a 16-byte aligned base.
template <typename T> struct alignas(4 * sizeof(T)) Vec4 { T v[4]; };
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_itk_resample_v4(const T* __restrict__ src, T* __restrict__ dst, int d, int h, int w, int D, int H, int W, ItkGeom g) {
  const unsigned total = (unsigned)D * H * W, HW = (unsigned)H * W;
  const unsigned groups = (total + 3) >> 2;
  for (unsigned gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const unsigned i0 = gi << 2;
    int Z = (int)(i0 / HW);
    const unsigned rem = i0 - (unsigned)Z * HW;
    int Y = (int)(rem / (unsigned)W), X = (int)(rem - (unsigned)Y * W);
    Vec4<T> out;
    bool fresh = true;
    // row state: nearest -> (row pointer, inside); linear -> four row pointers + z/y distances
    const T *r00 = src, *r01 = src, *r10 = src, *r11 = src;
    double dz = 0.0, dy = 0.0;
    bool row_in = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j >= total) { out.v[j] = (T)0; continue; }
      if (fresh) {
        fresh = false;
        const double c_z = itk_cidx(Z, g.inv_z, g.out_z), c_y = itk_cidx(Y, g.inv_y, g.out_y);
        if (MODE == 1) {
          row_in = true;
          const int zn = itk_nearest(c_z, d, row_in), yn = itk_nearest(c_y, h, row_in);
          r00 = src + ((long long)zn * h + yn) * w;
        } else {
          const AxisD az = itk_axis(c_z, d), ay = itk_axis(c_y, h);
          row_in = az.inside && ay.inside;
          dz = az.dist; dy = ay.dist;
          r00 = src + ((long long)az.i0 * h + ay.i0) * w; r01 = src + ((long long)az.i0 * h + ay.i1) * w;
          r10 = src + ((long long)az.i1 * h + ay.i0) * w; r11 = src + ((long long)az.i1 * h + ay.i1) * w;
        }
      }
      const double c_x = itk_cidx_line(X, W, g.inv_x, g.out_x);
      if (MODE == 1) {
        bool in = row_in;
        const int xn = itk_nearest(c_x, w, in);
        out.v[j] = in ? r00[xn] : (T)0;
      } else {
        const AxisD ax = itk_axis(c_x, w);
        const double lo = itk_lerp(itk_lerp((double)r00[ax.i0], (double)r00[ax.i1], ax.dist),
                                   itk_lerp((double)r01[ax.i0], (double)r01[ax.i1], ax.dist), dy);
        const double hi = itk_lerp(itk_lerp((double)r10[ax.i0], (double)r10[ax.i1], ax.dist),
                                   itk_lerp((double)r11[ax.i0], (double)r11[ax.i1], ax.dist), dy);
        out.v[j] = (row_in && ax.inside) ? cast_out<T>(itk_lerp(lo, hi, dz)) : (T)0;
      }
      if (++X == W) { X = 0; fresh = true; if (++Y == H) { Y = 0; ++Z; } }
    }
    if (i0 + 3 < total) {
      *reinterpret_cast<Vec4<T>*>(dst + i0) = out;
    } else {
      for (int j = 0; j < 4 && i0 + j < total; ++j) dst[i0 + j] = out.v[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------ RAM -> heat map
// as k_ram_upsample_mask_scatter (ram.cu) but the lobe mask is read from the scan-sized label volume.
// Block = 32 (x) x 8 (y) threads, kScatterZ consecutive z per thread: a warp reads 32 consecutive labels and writes up to
// 128 contiguous bytes of the heat map per z; the x / y interpolation set-up is shared by the thread's z steps; the eight
// source taps come from the 2 MB chunk RAM (L1 / L2 resident).  Algorithmic HBM bytes: 1 (label) + 4 (heat) per crop voxel.
constexpr int kScatterZ = 8;
__global__ void __launch_bounds__(256)
k_ram_upsample_label_scatter(const float* __restrict__ ram, const uint8_t* __restrict__ labels, int label,
                             float* __restrict__ heat, int d, int h, int w, int cd, int ch, int cw, int SH, int SW, int oz,
                             int oy, int ox, int act, float gain, float sz, float sy, float sx) {
  const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y, Z0 = blockIdx.z * kScatterZ;
  if (X >= cw || Y >= ch) return;
  const Lerp lx = lerp_setup(X, sx, w), ly = lerp_setup(Y, sy, h);
  const int o00 = ly.i0 * w + lx.i0, o01 = ly.i0 * w + lx.i1, o10 = ly.i1 * w + lx.i0, o11 = ly.i1 * w + lx.i1;
  const long long col = (long long)(Y + oy) * SW + (X + ox);
  const long long plane = (long long)SH * SW;
  // all label loads of the column first, then the source taps of ALL z steps unconditionally (clamped, L1/L2-resident 2 MB
  // chunk): the loads of the eight steps are independent and in flight together instead of one dependent chain per step
  uint8_t lab[kScatterZ];
#pragma unroll
  for (int k = 0; k < kScatterZ; ++k)
    lab[k] = (Z0 + k < cd) ? __ldg(labels + (long long)(Z0 + k + oz) * plane + col) : (uint8_t)0;
  bool any = false;
#pragma unroll
  for (int k = 0; k < kScatterZ; ++k) any = any || (lab[k] == label);
  if (!any) return;
  float v[kScatterZ];
#pragma unroll
  for (int k = 0; k < kScatterZ; ++k) {
    const Lerp lz = lerp_setup(min(Z0 + k, cd - 1), sz, d);
    const float* p0 = ram + (long long)lz.i0 * h * w;
    const float* p1 = ram + (long long)lz.i1 * h * w;
    float a00 = __ldg(p0 + o00), b00 = __ldg(p0 + o01), a01 = __ldg(p0 + o10), b01 = __ldg(p0 + o11);
    float a10 = __ldg(p1 + o00), b10 = __ldg(p1 + o01), a11 = __ldg(p1 + o10), b11 = __ldg(p1 + o11);
    if (act == 1) {
      a00 = sigmoidf_(a00); b00 = sigmoidf_(b00); a01 = sigmoidf_(a01); b01 = sigmoidf_(b01);
      a10 = sigmoidf_(a10); b10 = sigmoidf_(b10); a11 = sigmoidf_(a11); b11 = sigmoidf_(b11);
    }
    // same nesting as ATen's upsample_trilinear3d: d(h(w))
    float t = lz.w0 * (ly.w0 * (lx.w0 * a00 + lx.w1 * b00) + ly.w1 * (lx.w0 * a01 + lx.w1 * b01)) +
              lz.w1 * (ly.w0 * (lx.w0 * a10 + lx.w1 * b10) + ly.w1 * (lx.w0 * a11 + lx.w1 * b11));
    if (act == 2) t = fmaxf(t, 0.f);
    v[k] = t * gain;
  }
#pragma unroll
  for (int k = 0; k < kScatterZ; ++k)
    if (lab[k] == label) heat[(long long)(Z0 + k + oz) * plane + col] = v[k];
}

// ------------------------------------------------------------------------------------------------ Otsu inputs / masks
// histogram of uint8(window(v, lo, hi) * 255) over voxels with labels > 0  (utils.binary_cam utils.py:226-242);
// integer-exact: the uint8 conversion is the same truncation numpy's astype performs
template <typename T>
__device__ __forceinline__ int hist_bin(T raw, float lo, float hi) {
  // numpy evaluates the windowing in the array's float type: float32 for the heat map, float64 for the int16 scan
  if (sizeof(T) == 4) {
    float x = (float)raw;
    x = fminf(fmaxf(x, lo), hi);
    return (int)(((x - lo) / (hi - lo)) * 255.0f) & 255;
  }
  double x = (double)raw;
  x = x < lo ? (double)lo : (x > hi ? (double)hi : x);
  return (int)(((x - (double)lo) / ((double)hi - (double)lo)) * 255.0) & 255;
}
// a thread reads 4 labels as one word and skips all-background words before touching the values
template <typename T>
__global__ void __launch_bounds__(256)
k_masked_hist_u8(const T* __restrict__ v, const uint8_t* __restrict__ labels, long long n, float lo, float hi,
                 unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const long long words = n >> 2;
  const unsigned* lw = reinterpret_cast<const unsigned*>(labels);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < words; g += (long long)gridDim.x * blockDim.x) {
    const unsigned word = __ldg(lw + g);
    if (word == 0) continue;
    const Vec4<T> raw = *reinterpret_cast<const Vec4<T>*>(v + (g << 2));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((word >> (8 * j)) & 0xffu) atomicAdd(&sh[hist_bin<T>(raw.v[j], lo, hi)], 1u);
  }
  for (long long i = (words << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (labels[i] != 0) atomicAdd(&sh[hist_bin<T>(v[i], lo, hi)], 1u);
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// lesion = heat > th;  post = lesion && (window(scan) > th2) && !vessel      (job_runner.py:1009-1015)
__device__ __forceinline__ void threshold_voxel(float hv, short sv, uint8_t vv, float thf, double th2, double lo, double hi,
                                                double span, bool has_post, uint8_t& les_o, uint8_t& post_o) {
  const bool les = hv > thf;               // numpy compares the float32 heat map against th cast to float32
  les_o = les ? 1 : 0;
  post_o = 0;
  if (has_post && les) {                   // the float64 windowing only where the lesion mask is set
    double x = (double)sv;
    x = x < lo ? lo : (x > hi ? hi : x);
    post_o = (((x - lo) / span > th2) && !(vv > 0)) ? 1 : 0;
  }
}
// 4 voxels per thread: float4 / short4 / uchar4 loads, uchar4 stores
__global__ void __launch_bounds__(256)
k_threshold_masks(const float* __restrict__ heat, const short* __restrict__ scan, const uint8_t* __restrict__ vessel,
                  long long n, double th, double th2, float win_lo, float win_hi, uint8_t* __restrict__ lesion,
                  uint8_t* __restrict__ post) {
  const double lo = (double)win_lo, hi = (double)win_hi, span = hi - lo;
  const float thf = (float)th;
  const bool has_post = post != nullptr;
  const long long words = n >> 2;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < words; g += (long long)gridDim.x * blockDim.x) {
    const float4 h4 = __ldg(reinterpret_cast<const float4*>(heat) + g);
    const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
    Vec4<short> sv = {{0, 0, 0, 0}};
    Vec4<uint8_t> vv = {{0, 0, 0, 0}};
    if (has_post) {
      sv = *(reinterpret_cast<const Vec4<short>*>(scan) + g);
      if (vessel) vv = *(reinterpret_cast<const Vec4<uint8_t>*>(vessel) + g);
    }
    Vec4<uint8_t> lo4, po4;
#pragma unroll
    for (int j = 0; j < 4; ++j) threshold_voxel(hv[j], sv.v[j], vv.v[j], thf, th2, lo, hi, span, has_post, lo4.v[j], po4.v[j]);
    *(reinterpret_cast<Vec4<uint8_t>*>(lesion) + g) = lo4;
    if (has_post) *(reinterpret_cast<Vec4<uint8_t>*>(post) + g) = po4;
  }
  for (long long i = (words << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t l, p;
    threshold_voxel(heat[i], has_post ? scan[i] : (short)0, (has_post && vessel) ? vessel[i] : (uint8_t)0, thf, th2, lo, hi, span,
                    has_post, l, p);
    lesion[i] = l;
    if (has_post) post[i] = p;
  }
}

// lesion ratio of a scan (job_runner.py:772): sum of the heat map over voxels with a lobe label, and their count.
// 4 voxels per thread, all-background words skipped; per-block partials in double, summed in block order (deterministic).
__global__ void __launch_bounds__(256)
k_labelled_sum(const float* __restrict__ v, const uint8_t* __restrict__ labels, long long n, double* __restrict__ partial) {
  float s = 0.f;
  unsigned cnt = 0;
  const long long words = n >> 2;
  const unsigned* lw = reinterpret_cast<const unsigned*>(labels);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < words; g += (long long)gridDim.x * blockDim.x) {
    const unsigned word = __ldg(lw + g);
    if (word == 0) continue;
    const float4 x = __ldg(reinterpret_cast<const float4*>(v) + g);
    if (word & 0x000000ffu) { s += x.x; ++cnt; }
    if (word & 0x0000ff00u) { s += x.y; ++cnt; }
    if (word & 0x00ff0000u) { s += x.z; ++cnt; }
    if (word & 0xff000000u) { s += x.w; ++cnt; }
  }
  for (long long i = (words << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (labels[i]) { s += v[i]; ++cnt; }
  __shared__ double sh[2][8];
  double ds = (double)s, dc = (double)cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dc += __shfl_xor_sync(0xffffffffu, dc, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ds; sh[1][threadIdx.x >> 5] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; c += sh[1][w]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = c;
  }
}
__global__ void k_labelled_sum_finish(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  double a = 0.0, c = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) { a += partial[2 * b]; c += partial[2 * b + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if (threadIdx.x == 0) { out[0] = a; out[1] = c; }
}

// small result read-back without the copy engines: the kernel stores straight into PINNED host memory (unified virtual
// addressing: the host pointer is valid on the device), so a 1 KB histogram does not queue behind a bulk D2H transfer of
// another stream (LesionSegTest.run_scans downloads the previous scan's masks while this scan computes).
__global__ void k_store_to_host(const unsigned* __restrict__ src, unsigned* __restrict__ dst, int nwords) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

}  // namespace dram

using namespace dram;

extern "C" {

int dram_label_bboxes(const uint8_t* labels, int D, int H, int W, int nlabels, int* out, void* stream) {
  DRAM_REQUIRE(labels && out && D > 0 && H > 0 && W > 0 && nlabels >= 1 && nlabels <= 7, "label_bboxes: bad arguments (1..7 labels)");
  cudaStream_t st = (cudaStream_t)stream;
  k_bbox_init<<<1, 64, 0, st>>>(out, (nlabels + 1) * 6);
  DRAM_LAUNCH_CHECK();
  DRAM_REQUIRE((long long)D * H * W < (1ll << 31) - 4 && ((uintptr_t)labels % 4) == 0, "label_bboxes: volume too large or label base not 4-byte aligned");
  const int grid = grid_for(((long long)D * H * W + 3) / 4, 256 * 4, 8);
  if (nlabels <= 5) k_label_bboxes<5><<<grid, 256, 0, st>>>(labels, D, H, W, nlabels, out);
  else k_label_bboxes<7><<<grid, 256, 0, st>>>(labels, D, H, W, nlabels, out);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_lobe_chunk_preprocess(const short* scan, const uint8_t* labels, int SD, int SH, int SW, int label, int cz, int cy,
                               int cx, int cd, int ch, int cw, float win_lo, float win_hi, float pad_value, double sp_z,
                               double sp_y, double sp_x, float* img, float* msk, int d, int h, int w, void* stream) {
  DRAM_REQUIRE(scan && labels && img && msk && d > 0 && h > 0 && w > 0 && cd > 0 && ch > 0 && cw > 0, "lobe_chunk_preprocess: bad arguments");
  DRAM_REQUIRE(cz >= 0 && cy >= 0 && cx >= 0 && cz + cd <= SD && cy + ch <= SH && cx + cw <= SW, "lobe_chunk_preprocess: crop outside the scan");
  DRAM_REQUIRE(win_hi > win_lo, "lobe_chunk_preprocess: empty window");
  DRAM_REQUIRE(sp_z > 0.0 && sp_y > 0.0 && sp_x > 0.0, "lobe_chunk_preprocess: spacing must be positive");
  // Resample('fixed_size'), data_transforms.py:170-175: require_spacing = spacing * (current_size / size), in double
  const ItkGeom g = {1.0 / sp_z, 1.0 / sp_y, 1.0 / sp_x, sp_z * ((double)cd / (double)d), sp_y * ((double)ch / (double)h),
                     sp_x * ((double)cw / (double)w)};
  k_lobe_chunk_preprocess<<<grid_for((long long)d * h * w, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      scan, labels, SH, SW, label, cz, cy, cx, cd, ch, cw, win_lo, win_hi, pad_value, img, msk, d, h, w, g);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_itk_resample(const void* src, void* dst, int dtype, int d, int h, int w, int D, int H, int W, double in_sp_z,
                      double in_sp_y, double in_sp_x, double out_sp_z, double out_sp_y, double out_sp_x, int mode, void* stream) {
  DRAM_REQUIRE(src && dst && d > 0 && h > 0 && w > 0 && D > 0 && H > 0 && W > 0 && (mode == 0 || mode == 1), "itk_resample: bad arguments");
  DRAM_REQUIRE(in_sp_z > 0.0 && in_sp_y > 0.0 && in_sp_x > 0.0 && out_sp_z > 0.0 && out_sp_y > 0.0 && out_sp_x > 0.0,
               "itk_resample: spacings must be positive");
  const ItkGeom g = {1.0 / in_sp_z, 1.0 / in_sp_y, 1.0 / in_sp_x, out_sp_z, out_sp_y, out_sp_x};
  cudaStream_t st = (cudaStream_t)stream;
  if ((long long)D * H * W < (1ll << 31) - 4 && ((uintptr_t)dst % 16) == 0 && !getenv("DRAM_RESAMPLE_SCALAR")) {
    const int g4 = grid_for(((long long)D * H * W + 3) / 4, 256, 16);
#define ITK_V4(T)                                                                                                  \
  do {                                                                                                             \
    if (mode == 0) k_itk_resample_v4<T, 0><<<g4, 256, 0, st>>>((const T*)src, (T*)dst, d, h, w, D, H, W, g);        \
    else k_itk_resample_v4<T, 1><<<g4, 256, 0, st>>>((const T*)src, (T*)dst, d, h, w, D, H, W, g);                  \
  } while (0)
    if (dtype == 0) ITK_V4(float);
    else if (dtype == 1) ITK_V4(short);
    else if (dtype == 2) ITK_V4(uint8_t);
    else DRAM_REQUIRE(false, "itk_resample: dtype %d unknown (0 f32, 1 i16, 2 u8)", dtype);
#undef ITK_V4
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  int grid = grid_for((long long)D * H * W, 256, 16);
  if (dtype == 0) k_itk_resample<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, d, h, w, D, H, W, g, mode);
  else if (dtype == 1) k_itk_resample<short><<<grid, 256, 0, st>>>((const short*)src, (short*)dst, d, h, w, D, H, W, g, mode);
  else if (dtype == 2) k_itk_resample<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, d, h, w, D, H, W, g, mode);
  else DRAM_REQUIRE(false, "itk_resample: dtype %d unknown (0 f32, 1 i16, 2 u8)", dtype);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_ram_upsample_label_scatter(const float* ram, const uint8_t* labels, int label, float* heat, int d, int h, int w,
                                    int cd, int ch, int cw, int SD, int SH, int SW, int oz, int oy, int ox, int act,
                                    float gain, void* stream) {
  DRAM_REQUIRE(ram && labels && heat && d > 0 && h > 0 && w > 0 && cd > 0 && ch > 0 && cw > 0, "ram_upsample_label_scatter: bad arguments");
  DRAM_REQUIRE(act >= 0 && act <= 2, "ram_upsample_label_scatter: act %d unknown", act);
  DRAM_REQUIRE(oz >= 0 && oy >= 0 && ox >= 0 && oz + cd <= SD && oy + ch <= SH && ox + cw <= SW, "ram_upsample_label_scatter: crop outside the scan");
  DRAM_REQUIRE((cd + kScatterZ - 1) / kScatterZ <= 65535 && (ch + 7) / 8 <= 65535, "ram_upsample_label_scatter: crop too large");
  const dim3 grid((unsigned)((cw + 31) / 32), (unsigned)((ch + 7) / 8), (unsigned)((cd + kScatterZ - 1) / kScatterZ));
  k_ram_upsample_label_scatter<<<grid, dim3(32, 8, 1), 0, (cudaStream_t)stream>>>(
      ram, labels, label, heat, d, h, w, cd, ch, cw, SH, SW, oz, oy, ox, act, gain, ac_scale(d, cd), ac_scale(h, ch), ac_scale(w, cw));
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_masked_hist_u8(const void* values, int dtype, const uint8_t* labels, long long n, float lo, float hi,
                        unsigned int* hist, void* stream) {
  DRAM_REQUIRE(values && labels && hist && n > 0 && hi > lo, "masked_hist_u8: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DRAM_CUDA(cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned int), st));
  DRAM_REQUIRE(((uintptr_t)values % 16) == 0 && ((uintptr_t)labels % 4) == 0, "masked_hist_u8: values must be 16-byte, labels 4-byte aligned");
  int grid = grid_for((n + 3) / 4, 256 * 4, 8);
  if (dtype == 0) k_masked_hist_u8<float><<<grid, 256, 0, st>>>((const float*)values, labels, n, lo, hi, hist);
  else if (dtype == 1) k_masked_hist_u8<short><<<grid, 256, 0, st>>>((const short*)values, labels, n, lo, hi, hist);
  else DRAM_REQUIRE(false, "masked_hist_u8: dtype %d unknown (0 f32, 1 i16)", dtype);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_threshold_masks(const float* heat, const short* scan, const uint8_t* vessel, long long n, double th, double th2,
                         float win_lo, float win_hi, uint8_t* lesion, uint8_t* post, void* stream) {
  DRAM_REQUIRE(heat && lesion && n > 0, "threshold_masks: bad arguments");
  DRAM_REQUIRE(!post || (scan && win_hi > win_lo), "threshold_masks: post mask needs the scan and a window");
  DRAM_REQUIRE(((uintptr_t)heat % 16) == 0 && ((uintptr_t)lesion % 4) == 0 && (!post || (((uintptr_t)post % 4) == 0 && ((uintptr_t)scan % 8) == 0)) &&
                   (!vessel || ((uintptr_t)vessel % 4) == 0), "threshold_masks: volumes must be allocation-aligned (16 B)");
  k_threshold_masks<<<grid_for((n + 3) / 4, 256 * 2, 8), 256, 0, (cudaStream_t)stream>>>(heat, scan, vessel, n, th, th2, win_lo, win_hi, lesion, post);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_store_to_host(const void* src, void* dst_pinned_host, int nbytes, void* stream) {
  DRAM_REQUIRE(src && dst_pinned_host && nbytes > 0 && nbytes % 4 == 0 && nbytes <= (1 << 20), "store_to_host: 4..1 MiB, multiple of 4 bytes");
  DRAM_REQUIRE(((uintptr_t)src % 4) == 0 && ((uintptr_t)dst_pinned_host % 4) == 0, "store_to_host: unaligned pointers");
  cudaPointerAttributes attr;
  DRAM_CUDA(cudaPointerGetAttributes(&attr, dst_pinned_host));
  DRAM_REQUIRE(attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr, "store_to_host: destination is not pinned (device-mapped) host memory");
  const int nwords = nbytes / 4;
  k_store_to_host<<<(nwords + 255) / 256 > 64 ? 64 : (nwords + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      (const unsigned*)src, (unsigned*)attr.devicePointer, nwords);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static const int kLabelledSumBlocks = kNumSMs * 4;
size_t dram_labelled_sum_workspace_bytes(void) { return sizeof(double) * 2 * (size_t)kLabelledSumBlocks; }

int dram_labelled_sum(const float* values, const uint8_t* labels, long long n, double* out2, void* workspace, void* stream) {
  DRAM_REQUIRE(values && labels && out2 && workspace && n > 0, "labelled_sum: bad arguments");
  DRAM_REQUIRE(((uintptr_t)values % 16) == 0 && ((uintptr_t)labels % 4) == 0, "labelled_sum: values must be 16-byte, labels 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  k_labelled_sum<<<kLabelledSumBlocks, 256, 0, st>>>(values, labels, n, (double*)workspace);
  DRAM_LAUNCH_CHECK();
  k_labelled_sum_finish<<<1, 32, 0, st>>>((const double*)workspace, kLabelledSumBlocks, out2);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
