// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution for sm_100a (nn.Conv3d k3 p1 / k1, parts.py:105-106,133,185-186).
//
// Forward / dgrad (k_conv_umma_fwd)
//   GEMM view: D[m][co] = sum_{tap,ci} X[m + off(tap)][ci] * Wp[tap][co][ci];  M tile = 128 output voxels arranged as
//   a (TW,TH,TD) box of the volume, N tile = BN output channels, K loop = taps x (Cin_pad/64) blocks.
//   Per K block the producer warp issues ONE 5-D TMA box load per operand plane: the box is the output tile shifted by
//   the tap offset, and TMA's out-of-bounds zero fill implements the pad=1 halo, so no im2col buffer and no bounds
//   code exists anywhere.  The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which is
//   exactly the canonical K-major UMMA operand layout.  A single elected thread issues tcgen05.mma (M=128, N=BN,
//   K=16) into a double-buffered TMEM accumulator; four epilogue warps drain TMEM with tcgen05.ld and write fp32.
//   dgrad is the same kernel run on dy with the flipped/transposed weight pack (dram_pack_weight_bf16 mode 1).
//
// Split-bf16 ("bf16x3"): each fp32 operand is carried as hi = bf16(x), lo = bf16(x - hi); the MMA warp issues
//   hi*hi + hi*lo + lo*hi into the same fp32 accumulator (lo*lo ~ 2^-32 dropped).  This is what meets the 1e-3
//   parity target through 14 train-mode BatchNorm layers; lo == NULL selects 1-pass bf16.
//
// Wgrad (k_conv_umma_wgrad)
//   GEMM view: dW[(tap,ci)][co] = sum_m X[m + off(tap)][ci] * dY[m][co]: the reduction runs over voxels, so both
//   operands are MN-major: a TMA box of 64 voxels x 64 channels IS the canonical MN-major SW128 atom column.  The M
//   dimension of one MMA (128) is two 64-channel blocks that may belong to different taps (they are separate TMA boxes
//   LBO bytes apart), N = Cout tile.  Voxel chunks are split across CTAs (split-K); partial tiles go to a workspace and
//   a second kernel reduces them in a fixed order into the nn.Parameter layout [Cout][Cin][taps] (deterministic).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace dram {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap (context error, GPU released) instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One lane of a fully converged warp.  The loops around the TMA / MMA issue stay warp-uniform and only the issue itself
// is predicated on this: with `if (lane == 0)` around the whole loop nest the compiler keeps descriptors in vector
// registers and wraps every UTCHMMA / UTMALDG in a waterfall loop (R2UR + ELECT + BRA.ANY, ~70 issue cycles per MMA),
// which made the single issuing thread the bottleneck of every N <= 128 MMA sequence.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle, sm_100 version field = 1.
//   K-major : rows of 128 B, 8-row groups SBO = 1024 B apart, LBO unused (1).
//   MN-major: K rows of 128 B (= 64 MN elements), 8-row groups SBO = 1024 B apart, 64-element MN blocks LBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
constexpr int kFwdThreads = 192;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kATileBytes = 128 * 128;    // 128 voxel rows x 64 bf16

// fp32 -> (hi, lo) bf16 split, the same rounding as planes.cu's split2 (round to nearest even, lo = bf16(v - hi))
__device__ __forceinline__ void split1(float v, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
// 16 consecutive channels of one voxel -> 32 bytes per plane
__device__ __forceinline__ void store_planes16(uint16_t* hi, uint16_t* lo, const float (&v)[16]) {
  uint32_t H[8], L[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint16_t h0, l0, h1, l1;
    split1(v[2 * j], h0, l0);
    split1(v[2 * j + 1], h1, l1);
    H[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
    L[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
  }
  reinterpret_cast<uint4*>(hi)[0] = make_uint4(H[0], H[1], H[2], H[3]);
  reinterpret_cast<uint4*>(hi)[1] = make_uint4(H[4], H[5], H[6], H[7]);
  if (lo) {
    reinterpret_cast<uint4*>(lo)[0] = make_uint4(L[0], L[1], L[2], L[3]);
    reinterpret_cast<uint4*>(lo)[1] = make_uint4(L[4], L[5], L[6], L[7]);
  }
}

struct FwdParams {
  float* y;
  uint16_t* o_hi;   // eval mode: folded BatchNorm + ReLU output written as bf16 split planes [rows][Cout] instead of y
  uint16_t* o_lo;
  const float* scale;
  const float* shift;
  int N, D, H, W, Cout, BN, kblocks_c, taps, pad;
  int ksteps;   // K = 16 steps to issue per 64-channel block (< 4 only with ONE block whose upper channels are zero padding)
  int TW, TH, TD, tiles_w, tiles_h, tiles_d, n_mtiles, n_ntiles;
  int passes, stages, stage_bytes, tmem_cols, acc_cols;
  // kw-reuse mode (w3 = 1): one A box of (TW+2) x TH voxels per (kd,kh,channel block) serves the three kw taps.  The box is
  // loaded through a tensor map whose dimension order is (C,H,W,D,N), so shared-memory rows are ordered [w][h]: with
  // TH = 8 every w column is one 8-row / 1024-byte swizzle group and a kw shift is a 1024-byte-aligned descriptor offset.
  int w3, a_plane_bytes, taps_per_stage, b_tap_bytes, a_lo;
  int cb_split;   // virtual concat (parts.py:153 without the copy): channel blocks [0, cb_split) come from the first activation
                  // operand (tmA_*), blocks [cb_split, kblocks_c) from the second (tmA2_*); cb_split = kblocks_c: one operand
};

__global__ void __launch_bounds__(kFwdThreads, 1)
k_conv_umma_fwd(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), tfull0 = smem_u32(bars + 2 * S),
                 tempty0 = smem_u32(bars + 2 * S + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool three = p.passes == 3;            // the weight operand has a lo plane
  const bool a_lo = p.a_lo != 0;               // the activation operand has a lo plane (0: single-plane dy in dgrad)
  const uint32_t b_tile_bytes = (uint32_t)p.BN * 128u;
  const uint32_t offAlo = (uint32_t)p.a_plane_bytes, offBhi = (a_lo ? 2u : 1u) * (uint32_t)p.a_plane_bytes,
                 offBlo = offBhi + b_tile_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = p.n_mtiles * p.n_ntiles;
  const int tps = p.taps_per_stage;
  const int kblocks = (p.taps / tps) * p.kblocks_c;

  if (warp == 0) {
    // whole warp converged, one elected lane issues (see elect_one)
    if (elect_one()) {
      tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmB_hi);
      if (three) tma_prefetch_desc(&tmB_lo);
      if (a_lo) tma_prefetch_desc(&tmA_lo);
    }
    {
      // bytes TMA will actually deliver: the A box has TW*TH*TD rows (<= 128), zero-filled halo included
      const uint32_t a_box_bytes = (uint32_t)((p.w3 ? (p.TW + 2) : p.TW) * p.TH * p.TD) * 128u;
      const uint32_t stage_tx = (a_lo ? 2u : 1u) * a_box_bytes + (three ? 2u : 1u) * (uint32_t)tps * b_tile_bytes;
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_ntiles;
        int mt = tile / p.n_ntiles;
        const int w0 = (mt % p.tiles_w) * p.TW; mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.TH; mt /= p.tiles_h;
        const int d0 = (mt % p.tiles_d) * p.TD;
        const int n = mt / p.tiles_d;
        for (int g = 0; g < p.taps / tps; ++g) {
          // g = tap (generic) or (kd,kh) pair (kw-reuse)
          const int tap0 = g * tps;
          const int kd = p.taps == 1 ? 0 : tap0 / 9, kh = p.taps == 1 ? 0 : (tap0 / 3) % 3, kw = p.taps == 1 ? 0 : tap0 % 3;
          for (int cb = 0; cb < p.kblocks_c; ++cb) {
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (elect_one()) {
              const uint32_t sb = smem_u32(smem) + s * (uint32_t)p.stage_bytes, fb = full0 + 8 * s;
              mbar_expect_tx(fb, stage_tx);
              const bool second = cb >= p.cb_split;       // virtual concat: which activation tensor owns this channel block
              const CUtensorMap* mh = second ? &tmA2_hi : &tmA_hi;
              const CUtensorMap* ml = second ? &tmA2_lo : &tmA_lo;
              const int cc = (second ? cb - p.cb_split : cb) * 64;
              if (p.w3) {                                 // map dims (C,H,W,D,N); kw = 0 here, box starts at w0 - 1
                tma_load_5d(sb, mh, fb, cc, h0 + kh - 1, w0 - 1, d0 + kd - 1, n);
                if (a_lo) tma_load_5d(sb + offAlo, ml, fb, cc, h0 + kh - 1, w0 - 1, d0 + kd - 1, n);
              } else {
                tma_load_5d(sb, mh, fb, cc, w0 + kw - p.pad, h0 + kh - p.pad, d0 + kd - p.pad, n);
                if (a_lo) tma_load_5d(sb + offAlo, ml, fb, cc, w0 + kw - p.pad, h0 + kh - p.pad, d0 + kd - p.pad, n);
              }
              for (int t = 0; t < tps; ++t) {
                const uint32_t bo = (uint32_t)t * (uint32_t)p.b_tap_bytes;
                tma_load_2d(sb + offBhi + bo, &tmB_hi, fb, cb * 64, (tap0 + t) * p.Cout + nt * p.BN);
                if (three) tma_load_2d(sb + offBlo + bo, &tmB_lo, fb, cb * 64, (tap0 + t) * p.Cout + nt * p.BN);
              }
            }
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // split-bf16: the B_lo tile sits directly behind B_hi in shared memory, so A_hi x [B_hi | B_lo] is ONE MMA with
      // N = 2*BN (accumulator columns [0,BN) = hi*hi, [BN,2BN) = hi*lo) followed by A_lo x B_hi into [0,BN).
      // SS-mode operand reads cost 64 + 8192/N bytes per cycle against ~128 B/cycle of shared-memory bandwidth:
      // doubling N is what lifts the Cout = 64 layers off that bound.  The epilogue adds the two halves.
      const uint32_t idesc = umma_idesc(p.BN, 0, 0), idesc2 = umma_idesc(2 * p.BN, 0, 0);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.acc_cols;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)p.stage_bytes;
            for (int t = 0; t < tps; ++t) {            // kw-reuse: tap t reads the halo box shifted by t columns = t*1024 B
              const uint32_t ao = (uint32_t)t * 1024u, bo = (uint32_t)t * (uint32_t)p.b_tap_bytes;
              const uint64_t a_hi = umma_desc(sb + ao, 16, 1024), b_hi = umma_desc(sb + offBhi + bo, 16, 1024);
              const uint64_t a_lod = umma_desc(sb + offAlo + ao, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {           // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle row
                if (k >= p.ksteps) break;             // all-zero padding channels: nothing to accumulate
                const uint64_t adv = (uint64_t)(k * 2);
                const uint32_t accum = (kb | t | k) ? 1u : 0u;
                if (three) {
                  umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc2, accum);
                  if (a_lo) umma_bf16(d_tmem, a_lod + adv, b_hi + adv, idesc, 1u);
                } else {
                  umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
                }
              }
            }
            umma_commit(empty0 + 8 * s);               // frees the smem stage when these MMAs retire
            if (kb == kblocks - 1) umma_commit(tfull0 + 8 * acc);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int tw = p.w3 ? (row >> 3) : row % p.TW, th = p.w3 ? (row & 7) : (row / p.TW) % p.TH,
              td = p.w3 ? 0 : row / (p.TW * p.TH);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int nt = tile % p.n_ntiles;
      int mt = tile / p.n_ntiles;
      const int w = (mt % p.tiles_w) * p.TW + tw; mt /= p.tiles_w;
      const int h = (mt % p.tiles_h) * p.TH + th; mt /= p.tiles_h;
      const int d = (mt % p.tiles_d) * p.TD + td;
      const int n = mt / p.tiles_d;
      const bool valid = td < p.TD && w < p.W && h < p.H && d < p.D;
      const long long oo = ((((long long)n * p.D + d) * p.H + h) * p.W + w) * p.Cout + nt * p.BN;   // element offset in y / in the planes
      float* out = p.y + oo;
      mbar_wait(tfull0 + 8 * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.acc_cols;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16], r2[16];
        tmem_ld16(taddr + c0, r);
        if (three) tmem_ld16(taddr + p.BN + c0, r2);
        tmem_ld_wait();
        if (valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = three ? __uint_as_float(r[j]) + __uint_as_float(r2[j]) : __uint_as_float(r[j]);
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = nt * p.BN + c0 + j;
              v[j] = fmaxf(fmaf(v[j], __ldg(p.scale + c), __ldg(p.shift + c)), 0.f);
            }
          }
          if (p.o_hi) {
            const long long po = oo + c0;
            store_planes16(p.o_hi + po, p.o_lo ? p.o_lo + po : nullptr, v);
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ forward / dgrad, weight-sharing tile pairs
// ncu on k_conv_umma_fwd (profiles/r01_conv_fwd_us2c0_ncu_full.md): tensor pipe 44 % active, HBM 5 %, and
// l1tex__m_xbar2l1tex_read_bytes = 9.2 TB/s = 34.5 B/cycle/SM, i.e. the kernel sits on the chip-wide L2 -> SM limit
// (~6300 B/cycle).  Per 128-voxel tile and (kd,kh,channel block) it pulls 37 KB of activations but 49 KB of weights.
// This kernel halves the weight traffic and keeps the kw reuse of the activations:
//   * a work item is a PAIR of M tiles (two accumulators in TMEM) that share every weight tile;
//   * activations and weights travel through separate rings: an "A stage" holds the halo boxes of both tiles for one
//     (kd,kh,channel block) and serves the three kw taps through descriptor offsets; a "B stage" holds the weights of one tap;
//   * the tile is 8(h) x TDD(d) x TW(w) voxels with TDD*TW = 16, loaded through a tensor map with dimension order
//     (C,H,D,W,N): shared-memory rows are ordered [w][d][h], every (w,d) column is one 8-row / 1024-byte swizzle group
//     and a kw shift is TDD*1024 bytes, so (TW,TDD) = (16,1) covers W % 16 == 0 and (8,2) covers the 40^3 level.
// L2 bytes per tile and 3 taps: 37 + 3*B/2 KB against 3*(32 + B) KB (generic) or 37 + 3*B KB (kw reuse only).
// Accumulators: 2 sets x 2 tiles x <= 128 columns.  BN <= 64 keeps the [B_hi | B_lo] N-concatenation; BN = 96 / 128
// issues hi*hi, hi*lo, lo*hi as three N = BN MMAs into the same columns.
struct Fwd2Params {
  float* y;
  uint16_t* o_hi;   // eval mode: folded BatchNorm + ReLU output written as bf16 split planes [rows][Cout] instead of y
  uint16_t* o_lo;
  const float* scale;
  const float* shift;
  int N, D, H, W, Cout, BN, kblocks_c, ksteps;
  int TW, TDD, tiles_w, tiles_h, tiles_d, n_mtiles, n_ntiles, n_items;
  int SB, a_plane_bytes, a_tile_bytes, a_stage_bytes, b_stage_bytes, acc_cols, tmem_cols;
  long long* prof;   // DRAM_CONV_PROF: per-CTA cycle counters [8] (diagnostics only)
  float* stat;       // training: BatchNorm partial sums of y, [n_mtiles * 4][2][Cout] (row = M tile x epilogue warp), or NULL
  int cb_split;      // virtual concat: channel blocks >= cb_split are read from the second activation operand (tmA2_*)
  int dbg;           // DRAM_CONV_DBG (diagnostics, results are garbage): bit 0 = no activation loads after the first ring fill, bit 1 = no weight loads
};

// Column sums over the 32 lanes of a warp for 32 values per lane: after the five exchange rounds lane l holds the total
// of element l (fixed tree, deterministic).  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum32(float (&x)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? x[i] : x[i + 16], keep = up ? x[i + 16] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? x[i] : x[i + 8], keep = up ? x[i + 8] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? x[i] : x[i + 4], keep = up ? x[i + 4] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? x[i] : x[i + 2], keep = up ? x[i + 2] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? x[0] : x[1], keep = up ? x[1] : x[0];
    x[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return x[0];
}

// MODE 0: one pass (bf16); 1: split-bf16 with the [B_hi | B_lo] N-concatenation (BN <= 64); 2: split-bf16, three N = BN MMAs;
// 3 / 4: as 1 / 2 with a SINGLE-plane A operand (dgrad with dy carried as one bf16 plane): A x [B_hi | B_lo], resp. A x B_hi, A x B_lo
template <int MODE>
__global__ void __launch_bounds__(kFwdThreads, 1)
k_conv_umma_fwd2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                 const Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool three = MODE != 0 /* B has a lo plane */, a2 = (MODE == 1 || MODE == 2) /* A has a lo plane */,
                 concat = (MODE == 1 || MODE == 3);
  const int SB = p.SB;
  uint8_t* smemB = smem + 2 * (size_t)p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + (size_t)SB * p.b_stage_bytes);
  const uint32_t fullA0 = smem_u32(bars), emptyA0 = smem_u32(bars + 2), tfull0 = smem_u32(bars + 4), tempty0 = smem_u32(bars + 6),
                 fullB0 = smem_u32(bars + 8), emptyB0 = smem_u32(bars + 8 + SB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * SB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smemA_u = smem_u32(smem), smemB_u = smem_u32(smemB);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullA0 + 8 * s, 1); mbar_init(emptyA0 + 8 * s, 1);
      mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 4);
    }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp converged, one lane issues)
    if (elect_one()) {
      tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmB_hi);
      if (three) tma_prefetch_desc(&tmB_lo);
      if (a2) tma_prefetch_desc(&tmA_lo);
    }
    constexpr uint32_t planes = a2 ? 2u : 1u;
    const uint32_t b_tx = (three ? 2u : 1u) * (uint32_t)p.BN * 128u;
    uint32_t itA = 0, sB = 0, phB = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int nt = item % p.n_ntiles, pair = item / p.n_ntiles;
      const int ntile = (2 * pair + 1 < p.n_mtiles) ? 2 : 1;
      int mt0 = 2 * pair, mt1 = min(2 * pair + 1, p.n_mtiles - 1);
      const int w00 = (mt0 % p.tiles_w) * p.TW - 1; mt0 /= p.tiles_w;
      const int h00 = (mt0 % p.tiles_h) * 8 - 1; mt0 /= p.tiles_h;
      const int d00 = (mt0 % p.tiles_d) * p.TDD - 1;
      const int n0 = mt0 / p.tiles_d;
      const int w01 = (mt1 % p.tiles_w) * p.TW - 1; mt1 /= p.tiles_w;
      const int h01 = (mt1 % p.tiles_h) * 8 - 1; mt1 /= p.tiles_h;
      const int d01 = (mt1 % p.tiles_d) * p.TDD - 1;
      const int n1 = mt1 / p.tiles_d;
      const uint32_t a_tx = (uint32_t)ntile * planes * (uint32_t)p.a_plane_bytes;
      for (int g = 0; g < 9; ++g) {
        const int kd = g / 3, kh = g - 3 * kd;
        for (int cb = 0; cb < p.kblocks_c; ++cb, ++itA) {
          const uint32_t sA = itA & 1;
          mbar_wait(emptyA0 + 8 * sA, ((itA >> 1) & 1) ^ 1);
          if ((p.dbg & 1) && itA >= 2) {
            if (elect_one()) mbar_arrive(fullA0 + 8 * sA);
          } else if (elect_one()) {                          // map dims (C,H,D,W,N); the box starts at w0 - 1 (kw = 0)
            const uint32_t ab = smemA_u + sA * (uint32_t)p.a_stage_bytes, fa = fullA0 + 8 * sA;
            mbar_expect_tx(fa, a_tx);
            const bool second = cb >= p.cb_split;         // virtual concat: which activation tensor owns this channel block
            const CUtensorMap* mh = second ? &tmA2_hi : &tmA_hi;
            const CUtensorMap* ml = second ? &tmA2_lo : &tmA_lo;
            const int cc = (second ? cb - p.cb_split : cb) * 64;
            tma_load_5d(ab, mh, fa, cc, h00 + kh, d00 + kd, w00, n0);
            if (a2) tma_load_5d(ab + p.a_plane_bytes, ml, fa, cc, h00 + kh, d00 + kd, w00, n0);
            if (ntile == 2) {
              tma_load_5d(ab + p.a_tile_bytes, mh, fa, cc, h01 + kh, d01 + kd, w01, n1);
              if (a2) tma_load_5d(ab + p.a_tile_bytes + p.a_plane_bytes, ml, fa, cc, h01 + kh, d01 + kd, w01, n1);
            }
          }
          __syncwarp();
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait(emptyB0 + 8 * sB, phB ^ 1);
            if ((p.dbg & 2) && (itA > 8 || item != (int)blockIdx.x)) {
              if (elect_one()) mbar_arrive(fullB0 + 8 * sB);
            } else if (elect_one()) {
              const uint32_t bb = smemB_u + sB * (uint32_t)p.b_stage_bytes, fb = fullB0 + 8 * sB;
              mbar_expect_tx(fb, b_tx);
              const int brow = (g * 3 + kw) * p.Cout + nt * p.BN;
              tma_load_2d(bb, &tmB_hi, fb, cb * 64, brow);
              if (three) tma_load_2d(bb + p.BN * 128, &tmB_lo, fb, cb * 64, brow);
            }
            __syncwarp();
            if (++sB == (uint32_t)SB) { sB = 0; phB ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp converged, one lane issues)
    const uint32_t idesc = umma_idesc(p.BN, 0, 0), idesc2 = umma_idesc(2 * p.BN, 0, 0);
    const uint32_t kw_shift = (uint32_t)p.TDD * 1024u;
    const int n_g = 9 * p.kblocks_c;
    uint32_t itA = 0, sB = 0, phB = 0, icount = 0;
    long long t_te = 0, t_fa = 0, t_fb = 0, t0 = 0, t_start = clock64();
    const bool prof = p.prof != nullptr;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++icount) {
      const int pair = item / p.n_ntiles;
      const int ntile = (2 * pair + 1 < p.n_mtiles) ? 2 : 1;
      const uint32_t set = icount & 1;
      if (prof) t0 = clock64();
      mbar_wait(tempty0 + 8 * set, ((icount >> 1) & 1) ^ 1);
      if (prof) t_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_set = tmem_base + set * 2u * (uint32_t)p.acc_cols;
      for (int g = 0; g < n_g; ++g, ++itA) {
        const uint32_t sA = itA & 1;
        if (prof) t0 = clock64();
        mbar_wait(fullA0 + 8 * sA, (itA >> 1) & 1);
        if (prof) t_fa += clock64() - t0;
        tc_fence_after();
        const uint32_t ab = smemA_u + sA * (uint32_t)p.a_stage_bytes;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (prof) t0 = clock64();
          mbar_wait(fullB0 + 8 * sB, phB);
          if (prof) t_fb += clock64() - t0;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bb = smemB_u + sB * (uint32_t)p.b_stage_bytes;
            const uint64_t b_hi = umma_desc(bb, 16, 1024), b_lo = umma_desc(bb + (uint32_t)p.BN * 128u, 16, 1024);
            for (int j = 0; j < ntile; ++j) {
              const uint32_t aj = ab + (uint32_t)j * (uint32_t)p.a_tile_bytes + (uint32_t)kw * kw_shift;
              const uint64_t a_hi = umma_desc(aj, 16, 1024), a_lo = umma_desc(aj + (uint32_t)p.a_plane_bytes, 16, 1024);
              const uint32_t d_tmem = d_set + (uint32_t)j * (uint32_t)p.acc_cols;
#pragma unroll
              for (int k = 0; k < 4; ++k) {          // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle row
                if (k >= p.ksteps) break;            // all-zero padding channels: nothing to accumulate
                const uint64_t adv = (uint64_t)(k * 2);
                const uint32_t accum = (g | kw | k) ? 1u : 0u;
                if (concat) {
                  umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc2, accum);
                  if (a2) umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                } else {
                  umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
                  if (three) umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                  if (a2) umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                }
              }
            }
            umma_commit(emptyB0 + 8 * sB);             // frees the weight stage when these MMAs retire
            if (kw == 2) umma_commit(emptyA0 + 8 * sA);
            if (kw == 2 && g == n_g - 1) umma_commit(tfull0 + 8 * set);   // accumulators complete -> epilogue
          }
          __syncwarp();
          if (++sB == (uint32_t)SB) { sB = 0; phB ^= 1; }
        }
      }
    }
    if (prof && lane == 0) {
      long long* o = p.prof + blockIdx.x * 8;
      o[0] = clock64() - t_start; o[1] = t_te; o[2] = t_fa; o[3] = t_fb;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> fp32 channels-last
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row / (8 * p.TDD), tdd = (row >> 3) % p.TDD, th = row & 7;
    uint32_t icount = 0;
    long long e_wait = 0, e_work = 0, e0 = 0;
    const bool prof = p.prof != nullptr;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++icount) {
      const int nt = item % p.n_ntiles, pair = item / p.n_ntiles;
      const int ntile = (2 * pair + 1 < p.n_mtiles) ? 2 : 1;
      const uint32_t set = icount & 1;
      if (prof) e0 = clock64();
      mbar_wait(tfull0 + 8 * set, (icount >> 1) & 1);
      if (prof) { const long long t = clock64(); e_wait += t - e0; e0 = t; }
      tc_fence_after();
      for (int j = 0; j < ntile; ++j) {
        int mt = 2 * pair + j;
        const int w = (mt % p.tiles_w) * p.TW + tw; mt /= p.tiles_w;
        const int h = (mt % p.tiles_h) * 8 + th; mt /= p.tiles_h;
        const int d = (mt % p.tiles_d) * p.TDD + tdd;
        const int n = mt / p.tiles_d;
        const long long oo = ((((long long)n * p.D + d) * p.H + h) * p.W + w) * p.Cout + nt * p.BN;   // element offset in y / in the planes
      float* out = p.y + oo;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (set * 2u + (uint32_t)j) * (uint32_t)p.acc_cols;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t r[16], r2[16];
          tmem_ld16(taddr + c0, r);
          if (concat) tmem_ld16(taddr + p.BN + c0, r2);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            v[jj] = concat ? __uint_as_float(r[jj]) + __uint_as_float(r2[jj]) : __uint_as_float(r[jj]);
          if (p.stat) {
            // BatchNorm batch statistics of the raw output (parts.py:19, train mode) as a by-product of the epilogue: per
            // channel sum and sum of squares over this warp's 32 voxel rows, one partial row per (M tile, warp); the
            // partial rows are combined in double by k_bn_partials_reduce — no separate pass over y
            float x[32];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) { x[jj] = v[jj]; x[jj + 16] = v[jj] * v[jj]; }
            const float tot = warp_transpose_sum32(x, lane);
            const long long row = (long long)(2 * pair + j) * 4 + q;
            p.stat[(row * 2 + (lane >> 4)) * p.Cout + nt * p.BN + c0 + (lane & 15)] = tot;
          }
          if (p.scale) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int c = nt * p.BN + c0 + jj;
              v[jj] = fmaxf(fmaf(v[jj], __ldg(p.scale + c), __ldg(p.shift + c)), 0.f);
            }
          }
          if (p.o_hi) {
            const long long po = oo + c0;
            store_planes16(p.o_hi + po, p.o_lo ? p.o_lo + po : nullptr, v);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4)
              *reinterpret_cast<float4*>(out + c0 + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * set);
      if (prof) e_work += clock64() - e0;
    }
    if (prof && warp == 2 && lane == 0) { p.prof[blockIdx.x * 8 + 4] = e_wait; p.prof[blockIdx.x * 8 + 5] = e_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ forward / dgrad, channels on M
// ncu on k_conv_umma_fwd2 (profiles/r01b_conv_umma_fwd2_ncu_full.md): the busiest unit is the shared-memory operand path
// of the MMAs.  With voxels on M an MMA of N output channels reads (128 + N) x 32 B for N/2 tensor cycles: 192 B/cycle at
// N = 64, 128 B/cycle at N = 128, against the 128 B/cycle an SM delivers.  This kernel swaps the roles:
//   D^T[co][voxel] = sum_k W[co][k] * X[voxel][k]:  A operand = weights (M = 128 rows), B operand = the activation halo
//   box (N = 256 voxels), so every MMA reads 4 KB + 8 KB for 128 tensor cycles = 96 B/cycle whatever Cout is.
//   MODE 0 (64-channel tiles): A = [W_hi ; W_lo] stacked on M; two MMAs per K step (x X_hi, x X_lo) give all FOUR split
//            products (the lo*lo term comes for free); the epilogue adds accumulator rows co and co + 64.
//   MODE 1 (128-channel tiles): A = W_hi or W_lo; three MMAs per K step (hi*hi, lo*hi, hi*lo) into the same accumulator.
//   MODE 2 / 3: the same with a SINGLE-plane activation operand (dgrad with dy carried as one bf16 plane): MODE 2 is one
//            MMA per K step ([W_hi ; W_lo] x X = exactly the two products), MODE 3 two (W_hi x X, W_lo x X).
// Tile = 8(h) x TDD(d) x TW(w) = 256 voxels ((16,2) at 80^3, (8,4) at 40^3), same [w][d][h] row order and kw-shift trick
// as k_conv_umma_fwd2; accumulator 128 lanes x 256 columns, double buffered (all 512 TMEM columns).
struct Fwd3Params {
  float* y;
  uint16_t* o_hi;   // eval mode: folded BatchNorm + ReLU output written as bf16 split planes [rows][Cout] instead of y
  uint16_t* o_lo;
  const float* scale;
  const float* shift;
  int N, D, H, W, Cout, CT, kblocks_c, ksteps;
  int TW, TDD, tiles_w, tiles_h, tiles_d, n_vtiles, n_ctiles, n_items;
  int SW, x_plane_bytes, x_stage_bytes, w_stage_bytes;
  long long* prof;   // DRAM_CONV_PROF diagnostics
  float* stat;       // training: BatchNorm partial sums of y, [rows][2][Cout]; rows = n_vtiles (x 2 warps in the stacked modes)
  int cb_split;      // virtual concat: channel blocks >= cb_split are read from the second activation operand (tmX2_*)
  int dbg;           // DRAM_CONV_DBG (diagnostics, results are garbage): bit 0 = no activation loads after the first ring fill, bit 1 = no weight loads
};

template <int MODE>
__global__ void __launch_bounds__(kFwdThreads, 1)
k_conv_umma_fwd3(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const __grid_constant__ CUtensorMap tmX2_hi, const __grid_constant__ CUtensorMap tmX2_lo,
                 const Fwd3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool stacked = (MODE == 0 || MODE == 2), x2 = MODE < 2;   // x2: the activation operand has a lo plane
  constexpr int kXchgBytes = stacked ? 8192 : 0;     // 2 pairs x 2 parities x 2 directions x [8 columns][32 lanes] floats
  const int SW = p.SW;
  uint8_t* smemW = smem + 2 * (size_t)p.x_stage_bytes;
  float* xchg = reinterpret_cast<float*>(smemW + (size_t)SW * p.w_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + kXchgBytes);
  const uint32_t fullX0 = smem_u32(bars), emptyX0 = smem_u32(bars + 2), tfull0 = smem_u32(bars + 4), tempty0 = smem_u32(bars + 6),
                 fullW0 = smem_u32(bars + 8), emptyW0 = smem_u32(bars + 8 + SW);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * SW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smemX_u = smem_u32(smem), smemW_u = smem_u32(smemW);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullX0 + 8 * s, 1); mbar_init(emptyX0 + 8 * s, 1);
      mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 4);
    }
    for (int s = 0; s < SW; ++s) { mbar_init(fullW0 + 8 * s, 1); mbar_init(emptyW0 + 8 * s, 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_g = 9 * p.kblocks_c;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) { tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo); if (x2) tma_prefetch_desc(&tmX_lo); }
    const uint32_t x_tx = (x2 ? 2u : 1u) * (uint32_t)p.x_plane_bytes, w_tx = (uint32_t)p.w_stage_bytes;
    uint32_t itX = 0, sW = 0, phW = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int ct = item % p.n_ctiles;
      int vt = item / p.n_ctiles;
      const int w0 = (vt % p.tiles_w) * p.TW - 1; vt /= p.tiles_w;
      const int h0 = (vt % p.tiles_h) * 8 - 1; vt /= p.tiles_h;
      const int d0 = (vt % p.tiles_d) * p.TDD - 1;
      const int n = vt / p.tiles_d;
      for (int g = 0; g < 9; ++g) {
        const int kd = g / 3, kh = g - 3 * kd;
        for (int cb = 0; cb < p.kblocks_c; ++cb, ++itX) {
          const uint32_t sX = itX & 1;
          mbar_wait(emptyX0 + 8 * sX, ((itX >> 1) & 1) ^ 1);
          if ((p.dbg & 1) && itX >= 2) {
            if (elect_one()) mbar_arrive(fullX0 + 8 * sX);
          } else if (elect_one()) {                          // map dims (C,H,D,W,N); the box starts at w0 - 1 (kw = 0)
            const uint32_t xb = smemX_u + sX * (uint32_t)p.x_stage_bytes, fx = fullX0 + 8 * sX;
            mbar_expect_tx(fx, x_tx);
            const bool second = cb >= p.cb_split;         // virtual concat: which activation tensor owns this channel block
            const int cc = (second ? cb - p.cb_split : cb) * 64;
            tma_load_5d(xb, second ? &tmX2_hi : &tmX_hi, fx, cc, h0 + kh, d0 + kd, w0, n);
            if (x2) tma_load_5d(xb + p.x_plane_bytes, second ? &tmX2_lo : &tmX_lo, fx, cc, h0 + kh, d0 + kd, w0, n);
          }
          __syncwarp();
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait(emptyW0 + 8 * sW, phW ^ 1);
            if ((p.dbg & 2) && (itX > 8 || item != (int)blockIdx.x)) {
              if (elect_one()) mbar_arrive(fullW0 + 8 * sW);
            } else if (elect_one()) {
              const uint32_t wb = smemW_u + sW * (uint32_t)p.w_stage_bytes, fw = fullW0 + 8 * sW;
              mbar_expect_tx(fw, w_tx);
              const int wrow = (g * 3 + kw) * p.Cout + ct * p.CT;
              tma_load_2d(wb, &tmW_hi, fw, cb * 64, wrow);
              tma_load_2d(wb + (p.w_stage_bytes >> 1), &tmW_lo, fw, cb * 64, wrow);
            }
            __syncwarp();
            if (++sW == (uint32_t)SW) { sW = 0; phW ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc(256, 0, 0);
    const uint32_t kw_shift = (uint32_t)p.TDD * 1024u;
    uint32_t itX = 0, sW = 0, phW = 0, icount = 0;
    long long t_te = 0, t_fx = 0, t_fw = 0, t0 = 0, t_start = clock64();
    const bool prof = p.prof != nullptr;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++icount) {
      const uint32_t buf = icount & 1;
      if (prof) t0 = clock64();
      mbar_wait(tempty0 + 8 * buf, ((icount >> 1) & 1) ^ 1);
      if (prof) t_te += clock64() - t0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256u;
      for (int g = 0; g < n_g; ++g, ++itX) {
        const uint32_t sX = itX & 1;
        if (prof) t0 = clock64();
        mbar_wait(fullX0 + 8 * sX, (itX >> 1) & 1);
        if (prof) t_fx += clock64() - t0;
        tc_fence_after();
        const uint32_t xb = smemX_u + sX * (uint32_t)p.x_stage_bytes;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (prof) t0 = clock64();
          mbar_wait(fullW0 + 8 * sW, phW);
          if (prof) t_fw += clock64() - t0;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t wb = smemW_u + sW * (uint32_t)p.w_stage_bytes;
            const uint64_t x_hi = umma_desc(xb + (uint32_t)kw * kw_shift, 16, 1024);
            const uint64_t x_lo = umma_desc(xb + (uint32_t)p.x_plane_bytes + (uint32_t)kw * kw_shift, 16, 1024);
            const uint64_t w_a = umma_desc(wb, 16, 1024);                                     // stacked: [W_hi ; W_lo]; plain: W_hi
            const uint64_t w_b = umma_desc(wb + ((uint32_t)p.w_stage_bytes >> 1), 16, 1024);  // plain: W_lo
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k >= p.ksteps) break;              // all-zero padding channels: nothing to accumulate
              const uint64_t adv = (uint64_t)(k * 2);
              const uint32_t accum = (g | kw | k) ? 1u : 0u;
              umma_bf16(d_tmem, w_a + adv, x_hi + adv, idesc, accum);
              if (!stacked) umma_bf16(d_tmem, w_b + adv, x_hi + adv, idesc, 1u);
              if (x2) umma_bf16(d_tmem, w_a + adv, x_lo + adv, idesc, 1u);
            }
            umma_commit(emptyW0 + 8 * sW);
            if (kw == 2) umma_commit(emptyX0 + 8 * sX);
            if (kw == 2 && g == n_g - 1) umma_commit(tfull0 + 8 * buf);
          }
          __syncwarp();
          if (++sW == (uint32_t)SW) { sW = 0; phW ^= 1; }
        }
      }
    }
    if (prof && lane == 0) {
      long long* o = p.prof + blockIdx.x * 8;
      o[0] = clock64() - t_start; o[1] = t_te; o[2] = t_fx; o[3] = t_fw;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: accumulator row = channel, column = voxel
    const int q = warp & 3;
    const int lg = p.TDD == 4 ? 2 : (p.TDD == 2 ? 1 : 0);     // swizzle groups (8 voxels along h) per w column = TDD = 1 << lg
    const long long hstride = (long long)p.W * p.Cout;
    uint32_t icount = 0;
    long long e_wait = 0, e_work = 0, e0 = 0;
    const bool prof = p.prof != nullptr;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++icount) {
      const int ct = item % p.n_ctiles;
      int vt = item / p.n_ctiles;
      const int w0 = (vt % p.tiles_w) * p.TW; vt /= p.tiles_w;
      const int h0 = (vt % p.tiles_h) * 8; vt /= p.tiles_h;
      const int d0 = (vt % p.tiles_d) * p.TDD;
      const int n = vt / p.tiles_d;
      const uint32_t buf = icount & 1;
      if (prof) e0 = clock64();
      mbar_wait(tfull0 + 8 * buf, (icount >> 1) & 1);
      if (prof) { const long long t = clock64(); e_wait += t - e0; e0 = t; }
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u;
      if (!stacked) {
        const int co = ct * p.CT + q * 32 + lane;
        const float sc = p.scale ? __ldg(p.scale + co) : 1.f, sh = p.scale ? __ldg(p.shift + co) : 0.f;
        float s1 = 0.f, s2 = 0.f;            // BatchNorm partial sums: a thread owns a channel, no cross-thread reduction
        for (int c0 = 0; c0 < 256; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            const int grp = (c0 >> 3) + gi;
            const int w = w0 + (grp >> lg), d = d0 + (grp & (p.TDD - 1));
            const long long oo = ((((long long)n * p.D + d) * p.H + h0) * p.W + w) * p.Cout + co;
            float* out = p.y + oo;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = __uint_as_float(r[gi * 8 + j]);
              s1 += v; s2 = fmaf(v, v, s2);
              if (p.scale) v = fmaxf(fmaf(v, sc, sh), 0.f);
              if (p.o_hi) {                            // planes: 32 lanes = 64 contiguous bytes per plane
                const long long po = oo + j * hstride;
                uint16_t hh, ll;
                split1(v, hh, ll);
                p.o_hi[po] = hh;
                if (p.o_lo) p.o_lo[po] = ll;
              } else {
                out[j * hstride] = v;                  // 32 lanes = 32 consecutive channels of one voxel: 128 B
              }
            }
          }
        }
        if (p.stat) {
          float* sp = p.stat + (long long)(item / p.n_ctiles) * 2 * p.Cout + co;
          sp[0] = s1;
          sp[p.Cout] = s2;
        }
      } else {
        // rows 0..63 hold W_hi * X, rows 64..127 hold W_lo * X for the same channels: warps q and q+2 form a pair, each
        // owns one of the two 8-voxel groups of every 16 columns and receives the partner's values through shared memory
        const int pair = q & 1, upper = q >> 1;
        const int co = ct * p.CT + pair * 32 + lane;
        const float sc = p.scale ? __ldg(p.scale + co) : 1.f, sh = p.scale ? __ldg(p.shift + co) : 0.f;
        float s1 = 0.f, s2 = 0.f;            // this warp's half of the tile's voxels (partial row = 2 * voxel tile + upper)
        for (int c0 = 0; c0 < 256; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          const int par = (c0 >> 4) & 1;
          float* snd = xchg + (((pair * 2 + par) * 2 + upper) << 8);        // what this warp sends
          const float* rcv = xchg + (((pair * 2 + par) * 2 + (upper ^ 1)) << 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) snd[j * 32 + lane] = __uint_as_float(upper ? r[j] : r[j + 8]);
          asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
          const int grp = (c0 >> 3) + upper;                                 // the group this warp owns
          const int w = w0 + (grp >> lg), d = d0 + (grp & (p.TDD - 1));
          const long long oo = ((((long long)n * p.D + d) * p.H + h0) * p.W + w) * p.Cout + co;
          float* out = p.y + oo;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v = __uint_as_float(upper ? r[j + 8] : r[j]) + rcv[j * 32 + lane];
            s1 += v; s2 = fmaf(v, v, s2);
            if (p.scale) v = fmaxf(fmaf(v, sc, sh), 0.f);
            if (p.o_hi) {
              const long long po = oo + j * hstride;
              uint16_t hh, ll;
              split1(v, hh, ll);
              p.o_hi[po] = hh;
              if (p.o_lo) p.o_lo[po] = ll;
            } else {
              out[j * hstride] = v;
            }
          }
        }
        if (p.stat) {
          float* sp = p.stat + ((long long)(item / p.n_ctiles) * 2 + upper) * 2 * p.Cout + co;
          sp[0] = s1;
          sp[p.Cout] = s2;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      if (prof) e_work += clock64() - e0;
    }
    if (prof && warp == 2 && lane == 0) { p.prof[blockIdx.x * 8 + 4] = e_wait; p.prof[blockIdx.x * 8 + 5] = e_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------------ forward / dgrad, SM pairs (cta_group::2)
// `DRAM_CONV_DBG=3 python tests/micro_conv.py 8 us2.c0` (profiles/r02b_dbg_skip.txt): with NO TMA traffic at all the tile-pair
// kernel keeps its time (6.47 -> 6.32 ms) at 58 % tensor-pipe activity: k_conv_umma_fwd2 is bound by the shared-memory operand
// reads of its own MMAs (an M=128 x N MMA reads (128 + N) x 32 B for N/2 tensor cycles: 128 B/cycle at N = 128, 192 at N = 64).
// This kernel runs the same voxels-on-M GEMM on a PAIR of SMs (thread-block cluster of 2, tcgen05.mma.cta_group::2, M = 256):
//   * each CTA supplies the activation rows of its own 128-voxel tile(s) and only HALF of the weight columns of every MMA, so
//     the operand bytes per SM and MMA drop to (128 + N/2) x 32 B: 96 B/cycle for the N = 2*BN MMA below, and each weight tile
//     is fetched from L2 once per pair instead of once per CTA;
//   * split-bf16 with exactly three products for every BN: CTA r keeps the weight rows [W_hi[r*h, (r+1)*h) ; W_lo[r*h, (r+1)*h)]
//     (h = BN/2) of the N tile, so   MMA1: X_hi x [W_hi a | W_lo a | W_hi b | W_lo b]  (N = 2*BN; a / b = lower / upper half
//     of the tile's channels) fills accumulator columns [HH a | HL a | HH b | HL b], and   MMA2: X_lo x [W_hi a | W_hi b]
//     (N = BN: the first h rows of each CTA's weight stage, the SAME shared-memory descriptor) accumulates at column offset h,
//     i.e. LH a on top of HL a and LH b on top of HH b.  The epilogue adds column c and c + h of each half.
//   * leader CTA (cluster rank 0): one elected thread issues every MMA of the pair; both CTAs run a TMA producer (2-SM loads
//     that signal the LEADER's full barriers) and four epilogue warps on their own TMEM lanes; tcgen05.commit multicasts the
//     stage-free / accumulator-full arrivals to both CTAs; the epilogues of both CTAs arrive on the leader's tmem-empty barrier.
// Work item = 2*TPC M tiles (TPC per CTA, sharing every weight tile) x one N tile; accumulators 2 sets x TPC x 2*BN columns.
struct Fwd4Params {
  float* y;
  uint16_t* o_hi;   // eval mode: folded BatchNorm + ReLU output written as bf16 split planes [rows][Cout] instead of y
  uint16_t* o_lo;
  const float* scale;
  const float* shift;
  int N, D, H, W, Cout, BN, kblocks_c, ksteps;
  int co_base;       // this launch covers output channels [co_base, co_base + n_ntiles * BN) of the layer's Cout (Cout = 192 runs
                     // as one launch of 128-channel tiles and one of a 64-channel tile)
  int TW, TDD, tiles_w, tiles_h, tiles_d, n_mtiles, n_ntiles, n_items, TPC;
  int SA, SB, a_plane_bytes, a_tile_bytes, a_stage_bytes, b_stage_bytes;
  // generic-tile mode (gen = 1; the 20^3 / 10^3 levels, whose extents do not tile into 8-row boxes): the M tile is a
  // (TW, TH, TD) box of <= 128 voxels (rows ordered [d][h][w], rows beyond TW*TH*TD unused), every tap loads its own
  // shifted box through a (C,W,H,D,N) tensor map (no kw re-use) and the weights follow tap by tap
  int gen, TH, TD, a_box_bytes;
  long long* prof;   // DRAM_CONV_PROF: per-cluster cycle counters [8] (diagnostics only)
  float* stat;       // training: BatchNorm partial sums of y, [n_mtiles * 4][2][Cout] (row = M tile x epilogue warp), or NULL
  int cb_split;      // virtual concat: channel blocks >= cb_split are read from the second activation operand (tmA2_*)
  int dbg;           // DRAM_CONV_DBG (diagnostics, results are garbage): bit 0 = no activation loads after the first ring fill, bit 1 = no weight loads
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA loads: the data lands in the executing CTA's shared memory, the transaction bytes are counted on `bar`, a
// shared::cluster address that may belong to the peer CTA (the pair's leader)
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// collective over the pair: the same warp of both CTAs executes it with the same arguments
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// kind::f16 instruction descriptor of the pair: D=f32, A=B=bf16 K-major, M=256 (128 rows per CTA), N=n (a multiple of 16)
__device__ __forceinline__ uint32_t umma_idesc_2sm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
k_conv_umma_fwd4(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                 const Fwd4Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int SA = p.SA, SB = p.SB, TPC = p.TPC;
  uint8_t* smemB = smem + (size_t)SA * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + (size_t)SB * p.b_stage_bytes);
  // fullA / fullB / tempty are used in the leader only; emptyA / emptyB / tfull exist (and are signalled) in both CTAs
  const uint32_t fullA0 = smem_u32(bars), emptyA0 = fullA0 + 8 * SA, fullB0 = emptyA0 + 8 * SA, emptyB0 = fullB0 + 8 * SB,
                 tfull0 = emptyB0 + 8 * SB, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint32_t smemA_u = smem_u32(smem), smemB_u = smem_u32(smemB);
  const int h = p.BN >> 1, acc_cols = 2 * p.BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(fullA0 + 8 * s, 1); mbar_init(emptyA0 + 8 * s, 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // barriers of both CTAs initialised before any remote arrive / 2-SM load / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_g = (p.gen ? 27 : 9) * p.kblocks_c;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs: own tiles, own weight half)
    if (elect_one()) {
      tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    }
    const uint32_t a_tx = 2u * (uint32_t)TPC * 2u * (uint32_t)(p.gen ? p.a_box_bytes : p.a_plane_bytes);     // both CTAs of the pair
    const uint32_t b_tx = 2u * (uint32_t)p.b_stage_bytes;
    const int n_grp = p.gen ? 27 : 9, kw_n = p.gen ? 1 : 3;
    uint32_t sA = 0, phA = 0, sB = 0, phB = 0;
    for (int item = cid; item < p.n_items; item += n_clusters) {
      const int nt = item % p.n_ntiles, grp = item / p.n_ntiles;
      int w0[2], h0[2], d0[2], n0[2];
      for (int j = 0; j < TPC; ++j) {
        int mt = (grp * 2 + (int)rank) * TPC + j;
        w0[j] = (mt % p.tiles_w) * p.TW - 1; mt /= p.tiles_w;
        h0[j] = (mt % p.tiles_h) * (p.gen ? p.TH : 8) - 1; mt /= p.tiles_h;
        d0[j] = (mt % p.tiles_d) * (p.gen ? p.TD : p.TDD) - 1;
        n0[j] = mt / p.tiles_d;
      }
      for (int g = 0; g < n_grp; ++g) {
        const int kd = p.gen ? g / 9 : g / 3, kh = p.gen ? (g / 3) % 3 : g - 3 * kd, kwg = p.gen ? g % 3 : 0;
        for (int cb = 0; cb < p.kblocks_c; ++cb) {
          mbar_wait(emptyA0 + 8 * sA, phA ^ 1);
          if ((p.dbg & 1) && (phA || item != cid)) {
            if (leader && elect_one()) mbar_arrive(fullA0 + 8 * sA);
          } else if (elect_one()) {                          // map dims (C,H,D,W,N); the box starts at w0 - 1 (kw = 0)
            const uint32_t ab = smemA_u + sA * (uint32_t)p.a_stage_bytes, fa = mapa_u32(fullA0 + 8 * sA, 0);
            if (leader) mbar_expect_tx(fullA0 + 8 * sA, a_tx);
            const bool second = cb >= p.cb_split;         // virtual concat: which activation tensor owns this channel block
            const CUtensorMap* mh = second ? &tmA2_hi : &tmA_hi;
            const CUtensorMap* ml = second ? &tmA2_lo : &tmA_lo;
            const int cc = (second ? cb - p.cb_split : cb) * 64;
            for (int j = 0; j < TPC; ++j) {
              if (p.gen) {                              // map dims (C,W,H,D,N): this tap's own shifted box
                tma_load_5d_2sm(ab + j * p.a_tile_bytes, mh, fa, cc, w0[j] + kwg, h0[j] + kh, d0[j] + kd, n0[j]);
                tma_load_5d_2sm(ab + j * p.a_tile_bytes + p.a_plane_bytes, ml, fa, cc, w0[j] + kwg, h0[j] + kh, d0[j] + kd, n0[j]);
              } else {
                tma_load_5d_2sm(ab + j * p.a_tile_bytes, mh, fa, cc, h0[j] + kh, d0[j] + kd, w0[j], n0[j]);
                tma_load_5d_2sm(ab + j * p.a_tile_bytes + p.a_plane_bytes, ml, fa, cc, h0[j] + kh, d0[j] + kd, w0[j], n0[j]);
              }
            }
          }
          __syncwarp();
          if (++sA == (uint32_t)SA) { sA = 0; phA ^= 1; }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (kw >= kw_n) break;
            mbar_wait(emptyB0 + 8 * sB, phB ^ 1);
            if ((p.dbg & 2) && (phB || item != cid)) {
              if (leader && elect_one()) mbar_arrive(fullB0 + 8 * sB);
            } else if (elect_one()) {
              const uint32_t bb = smemB_u + sB * (uint32_t)p.b_stage_bytes, fb = mapa_u32(fullB0 + 8 * sB, 0);
              if (leader) mbar_expect_tx(fullB0 + 8 * sB, b_tx);
              const int brow = (p.gen ? g : g * 3 + kw) * p.Cout + p.co_base + nt * p.BN + (int)rank * h;
              tma_load_2d_2sm(bb, &tmB_hi, fb, cb * 64, brow);
              tma_load_2d_2sm(bb + h * 128, &tmB_lo, fb, cb * 64, brow);
            }
            __syncwarp();
            if (++sB == (uint32_t)SB) { sB = 0; phB ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------------- MMA issuer of the pair (one elected lane issues)
      // DRAM_CONV_DBG bit 2 (diagnostics, garbage results): issue the N = BN MMA with N = 2*BN at column 0 — does a pair MMA of
      // N = 64 cost less than one of N = 128?
      const uint32_t idesc1 = umma_idesc_2sm(2 * p.BN), idesc2 = umma_idesc_2sm((p.dbg & 4) ? 2 * p.BN : p.BN);
      const uint32_t d2_off = (p.dbg & 4) ? 0u : (uint32_t)h;
      const uint32_t kw_shift = (uint32_t)p.TDD * 1024u;
      uint32_t sA = 0, phA = 0, sB = 0, phB = 0, icount = 0;
      long long t_te = 0, t_fa = 0, t_fb = 0, t0 = 0, t_start = clock64();
      const bool prof = p.prof != nullptr;
      for (int item = cid; item < p.n_items; item += n_clusters, ++icount) {
        const uint32_t set = icount & 1;
        if (prof) t0 = clock64();
        mbar_wait(tempty0 + 8 * set, ((icount >> 1) & 1) ^ 1);
        if (prof) t_te += clock64() - t0;
        tc_fence_after();
        const uint32_t d_set = tmem_base + set * (uint32_t)(TPC * acc_cols);
        for (int g = 0; g < n_g; ++g) {
          if (prof) t0 = clock64();
          mbar_wait(fullA0 + 8 * sA, phA);
          if (prof) t_fa += clock64() - t0;
          tc_fence_after();
          const uint32_t ab = smemA_u + sA * (uint32_t)p.a_stage_bytes;
          const int kw_n = p.gen ? 1 : 3;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            if (kw >= kw_n) break;
            const bool last_kw = kw == kw_n - 1;
            if (prof) t0 = clock64();
            mbar_wait(fullB0 + 8 * sB, phB);
            if (prof) t_fb += clock64() - t0;
            tc_fence_after();
            if (elect_one()) {
              const uint64_t b_d = umma_desc(smemB_u + sB * (uint32_t)p.b_stage_bytes, 16, 1024);
              for (int j = 0; j < TPC; ++j) {
                const uint32_t aj = ab + (uint32_t)j * (uint32_t)p.a_tile_bytes + (uint32_t)kw * kw_shift;
                const uint64_t a_hi = umma_desc(aj, 16, 1024), a_lo = umma_desc(aj + (uint32_t)p.a_plane_bytes, 16, 1024);
                const uint32_t d_tmem = d_set + (uint32_t)(j * acc_cols);
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle row
                  if (k >= p.ksteps) break;            // all-zero padding channels: nothing to accumulate
                  const uint64_t adv = (uint64_t)(k * 2);
                  umma_bf16_2sm(d_tmem, a_hi + adv, b_d + adv, idesc1, (g | kw | k) ? 1u : 0u);
                  umma_bf16_2sm(d_tmem + d2_off, a_lo + adv, b_d + adv, idesc2, 1u);
                }
              }
              umma_commit_2sm(emptyB0 + 8 * sB);             // frees the weight stage of both CTAs when these MMAs retire
              if (last_kw) umma_commit_2sm(emptyA0 + 8 * sA);
              if (last_kw && g == n_g - 1) umma_commit_2sm(tfull0 + 8 * set);   // accumulators complete -> both epilogues
            }
            __syncwarp();
            if (++sB == (uint32_t)SB) { sB = 0; phB ^= 1; }
          }
          if (++sA == (uint32_t)SA) { sA = 0; phA ^= 1; }
        }
      }
      if (prof && lane == 0) {
        long long* o = p.prof + cid * 8;
        o[0] = clock64() - t_start; o[1] = t_te; o[2] = t_fa; o[3] = t_fb;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs): own TMEM lanes -> fp32 channels-last
    const int q = warp & 3;
    const int row = q * 32 + lane;
    // kw-reuse tiles: rows ordered [w][d][h]; generic tiles: [d][h][w] with rows >= TW*TH*TD unused
    const int tw = p.gen ? row % p.TW : row / (8 * p.TDD), tdd = p.gen ? row / (p.TW * p.TH) : (row >> 3) % p.TDD,
              th = p.gen ? (row / p.TW) % p.TH : row & 7;
    const uint32_t tempty_leader = mapa_u32(tempty0, 0);
    uint32_t icount = 0;
    long long e_wait = 0, e_work = 0, e0 = 0;
    const bool prof = p.prof != nullptr;
    for (int item = cid; item < p.n_items; item += n_clusters, ++icount) {
      const int nt = item % p.n_ntiles, grp = item / p.n_ntiles;
      const uint32_t set = icount & 1;
      if (prof) e0 = clock64();
      mbar_wait(tfull0 + 8 * set, (icount >> 1) & 1);
      if (prof) { const long long t = clock64(); e_wait += t - e0; e0 = t; }
      tc_fence_after();
      for (int j = 0; j < TPC; ++j) {
        const int mtile = (grp * 2 + (int)rank) * TPC + j;
        int mt = mtile;
        const int w = (mt % p.tiles_w) * p.TW + tw; mt /= p.tiles_w;
        const int hh = (mt % p.tiles_h) * (p.gen ? p.TH : 8) + th; mt /= p.tiles_h;
        const int d = (mt % p.tiles_d) * (p.gen ? p.TD : p.TDD) + tdd;
        const int n = mt / p.tiles_d;
        const bool valid = !p.gen || (tdd < p.TD && w < p.W && hh < p.H && d < p.D);    // ragged generic tiles: masked rows
        const int cbase = p.co_base + nt * p.BN;
        const long long oo = ((((long long)n * p.D + d) * p.H + hh) * p.W + w) * p.Cout + cbase;   // element offset in y / in the planes
        float* out = p.y + oo;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (set * (uint32_t)TPC + (uint32_t)j) * (uint32_t)acc_cols;
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          // channels [c0, c0 + 16) of the N tile: lower half -> columns [c0 | h + c0], upper half -> [2h + (c0 - h) | 3h + (c0 - h)]
          const uint32_t col = (uint32_t)(c0 < h ? c0 : c0 + h);
          uint32_t r[16], r2[16];
          tmem_ld16(taddr + col, r);
          tmem_ld16(taddr + col + (uint32_t)h, r2);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(r[jj]) + __uint_as_float(r2[jj]);
          if (p.stat) {                              // BatchNorm batch statistics as in k_conv_umma_fwd2
            float x[32];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) { x[jj] = v[jj]; x[jj + 16] = v[jj] * v[jj]; }
            const float tot = warp_transpose_sum32(x, lane);
            const long long srow = (long long)mtile * 4 + q;
            p.stat[(srow * 2 + (lane >> 4)) * p.Cout + cbase + c0 + (lane & 15)] = tot;
          }
          if (p.scale) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int c = cbase + c0 + jj;
              v[jj] = fmaxf(fmaf(v[jj], __ldg(p.scale + c), __ldg(p.shift + c)), 0.f);
            }
          }
          if (valid) {
            if (p.o_hi) {
              const long long po = oo + c0;
              store_planes16(p.o_hi + po, p.o_lo ? p.o_lo + po : nullptr, v);
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4)
                *reinterpret_cast<float4*>(out + c0 + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * set);
      if (prof) e_work += clock64() - e0;
    }
    if (prof && leader && warp == 2 && lane == 0) { p.prof[cid * 8 + 4] = e_wait; p.prof[cid * 8 + 5] = e_work; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int kWgKV = 64;                       // voxels per K chunk (one TMA box)
constexpr int kWgBlkBytes = kWgKV * 128;        // one 64-voxel x 64-channel MN-major block

struct WgParams {
  float* ws;                                    // [slabs][n_mtiles][n_ntiles][128][BN]
  int N, D, H, W, taps, pad, CB /*Cin_pad/64*/, MB /*taps*CB*/, n_mtiles, n_ntiles, BN;
  int TW, TH, TD, TN, tiles_w, tiles_h, tiles_d, tiles_n, n_chunks, n_slabs, spg;
  int passes, stages, stage_bytes, tmem_cols, concat, y_lo;
  int cb_split;   // virtual concat: input channel blocks >= cb_split are read from the second X operand (tmX2_*)
};
// slab -> chunks first, first + step, ... (count of them): the slabs of a group of `spg` consecutive slabs (about one wave of
// CTAs) interleave over the group's contiguous chunk range, so that all CTAs in flight read neighbouring chunks (see Wg3Params)
struct WgRange { int first, step, count; };
__device__ __forceinline__ WgRange wg_range(int n_chunks, int n_slabs, int spg, int slab) {
  const int g = slab / spg, s0 = g * spg, s1 = min(s0 + spg, n_slabs);
  const int lo = (int)((long long)n_chunks * s0 / n_slabs), hi = (int)((long long)n_chunks * s1 / n_slabs);
  WgRange r;
  r.step = s1 - s0;
  r.first = lo + (slab - s0);
  r.count = r.first < hi ? (hi - r.first + r.step - 1) / r.step : 0;
  return r;
}

__global__ void __launch_bounds__(kFwdThreads, 1)
k_conv_umma_wgrad(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                  const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                  const __grid_constant__ CUtensorMap tmX2_hi, const __grid_constant__ CUtensorMap tmX2_lo,
                  const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), tfull0 = smem_u32(bars + 2 * S),
                 tempty0 = smem_u32(bars + 2 * S + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool three = p.passes == 3;             // X (layer input) has a lo plane
  const bool y2 = p.y_lo != 0;                  // dY has a lo plane (0: gradient carried as one bf16 plane)
  const int nb = p.BN / 64;                                            // 64-channel blocks of the N operand
  const uint32_t a_bytes = 2u * kWgBlkBytes, b_bytes = (uint32_t)nb * kWgBlkBytes;
  const uint32_t offAlo = a_bytes, offBhi = three ? 2 * a_bytes : a_bytes, offBlo = offBhi + b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_stride = p.tmem_cols / 2;
  const int n_items = p.n_slabs * p.n_mtiles * p.n_ntiles;

  if (warp == 0) {
    // whole warp converged, one elected lane issues (see elect_one)
    if (elect_one()) {
      tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmY_hi);
      if (three) tma_prefetch_desc(&tmX_lo);
      if (y2) tma_prefetch_desc(&tmY_lo);
    }
    {
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nt = item % p.n_ntiles;
        const int mt = (item / p.n_ntiles) % p.n_mtiles;
        const int slab = item / (p.n_ntiles * p.n_mtiles);
        const int mb0 = 2 * mt, mb1 = 2 * mt + 1;
        const bool has1 = mb1 < p.MB;
        const uint32_t tx = (three ? 2u : 1u) * (has1 ? 2u : 1u) * kWgBlkBytes + (y2 ? 2u : 1u) * b_bytes;
        const int tap0 = mb0 / p.CB, cb0 = mb0 % p.CB, tap1 = has1 ? mb1 / p.CB : 0, cb1 = has1 ? mb1 % p.CB : 0;
        const int kd0 = p.taps == 1 ? 0 : tap0 / 9 - p.pad, kh0 = p.taps == 1 ? 0 : (tap0 / 3) % 3 - p.pad, kw0 = p.taps == 1 ? 0 : tap0 % 3 - p.pad;
        const int kd1 = p.taps == 1 ? 0 : tap1 / 9 - p.pad, kh1 = p.taps == 1 ? 0 : (tap1 / 3) % 3 - p.pad, kw1 = p.taps == 1 ? 0 : tap1 % 3 - p.pad;
        const WgRange rg = wg_range(p.n_chunks, p.n_slabs, p.spg, slab);
        for (int k = 0; k < rg.count; ++k) {
          int t = rg.first + k * rg.step;            // depth fastest (see k_conv_umma_wgrad_w3)
          const int d0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
          const int w0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
          const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
          const int n0 = t * p.TN;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          if (elect_one()) {
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)p.stage_bytes, fb = full0 + 8 * s;
            mbar_expect_tx(fb, tx);
            const bool s0 = cb0 >= p.cb_split, s1 = cb1 >= p.cb_split;       // virtual concat: second X operand
            const int cc0 = (s0 ? cb0 - p.cb_split : cb0) * 64, cc1 = (s1 ? cb1 - p.cb_split : cb1) * 64;
            tma_load_5d(sb, s0 ? &tmX2_hi : &tmX_hi, fb, cc0, w0 + kw0, h0 + kh0, d0 + kd0, n0);
            if (three) tma_load_5d(sb + offAlo, s0 ? &tmX2_lo : &tmX_lo, fb, cc0, w0 + kw0, h0 + kh0, d0 + kd0, n0);
            if (has1) {
              tma_load_5d(sb + kWgBlkBytes, s1 ? &tmX2_hi : &tmX_hi, fb, cc1, w0 + kw1, h0 + kh1, d0 + kd1, n0);
              if (three) tma_load_5d(sb + offAlo + kWgBlkBytes, s1 ? &tmX2_lo : &tmX_lo, fb, cc1, w0 + kw1, h0 + kh1, d0 + kd1, n0);
            }
            for (int j = 0; j < nb; ++j) {
              tma_load_5d(sb + offBhi + j * kWgBlkBytes, &tmY_hi, fb, nt * p.BN + j * 64, w0, h0, d0, n0);
              if (y2) tma_load_5d(sb + offBlo + j * kWgBlkBytes, &tmY_lo, fb, nt * p.BN + j * 64, w0, h0, d0, n0);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = umma_idesc(p.BN, 1, 1), idesc2 = umma_idesc(2 * p.BN, 1, 1);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
        const int slab = item / (p.n_ntiles * p.n_mtiles);
        const int c_begin = 0, c_end = wg_range(p.n_chunks, p.n_slabs, p.spg, slab).count;
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)acc_stride;
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)p.stage_bytes;
            const uint64_t a_hi = umma_desc(sb, kWgBlkBytes, 1024), b_hi = umma_desc(sb + offBhi, kWgBlkBytes, 1024);
            const uint64_t a_lo = umma_desc(sb + offAlo, kWgBlkBytes, 1024), b_lo = umma_desc(sb + offBlo, kWgBlkBytes, 1024);
#pragma unroll
            for (int k = 0; k < kWgKV / 16; ++k) {      // 16 voxel rows (2048 B) per MMA
              const uint64_t adv = (uint64_t)(k * (2048 >> 4));
              const uint32_t first = (ch > c_begin || k) ? 1u : 0u;
              if (p.concat) {                           // X_hi x [dY_hi | dY_lo] (N = 2*BN), then X_lo x dY_hi
                umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc2, first);
                umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              } else {
                umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, first);
                if (y2) umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                if (three) umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            }
            umma_commit(empty0 + 8 * s);
            if (ch == c_end - 1) umma_commit(tfull0 + 8 * acc);
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int mt = (item / p.n_ntiles) % p.n_mtiles;
      const bool valid = (2 * mt + row / 64) < p.MB;
      float* out = p.ws + ((long long)item * 128 + row) * p.BN;
      mbar_wait(tfull0 + 8 * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)acc_stride;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16], r2[16];
        tmem_ld16(taddr + c0, r);
        if (p.concat) tmem_ld16(taddr + p.BN + c0, r2);
        tmem_ld_wait();
        if (valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = p.concat ? __uint_as_float(r[j]) + __uint_as_float(r2[j]) : __uint_as_float(r[j]);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ wgrad, SM pairs (Cout tiles of 128)
// The generic kernel above issues three M = 128, N = 128 MMAs per K step: 8 KB of shared-memory operands per 64 tensor cycles
// (tensor pipe 78 % active, profiles/r01b_conv_umma_fwd2_ncu_full.md section 5) and every CTA streams the whole dY tile.
// Pair version (cta_group::2, see k_conv_umma_fwd4): M = 256 = FOUR (tap, channel-block) blocks of X^T, two per CTA; the dY
// tile of 128 channels is split across the pair — CTA r keeps the 64-channel block r as [dY_hi r | dY_lo r] — so that
//   MMA1: X_hi^T x [dY_hi a | dY_lo a | dY_hi b | dY_lo b] (N = 256)  and  MMA2: X_lo^T x [dY_hi a | dY_hi b] (N = 128, the
//   first block of each CTA's stage, accumulated at column offset 64: LH a on top of HL a, LH b on top of HH b)
// give exactly the three split products with 64 / 96 B/cycle of operand reads per SM, and each CTA fetches half of dY.
// Work item = (slab, group of four M blocks, N tile); the partial tiles land in the generic kernel's workspace layout
// [slab][M tile = pair of blocks][N tile][128][128], so k_wgrad_reduce is shared.
struct Wg2Params {
  float* ws;
  int N, D, H, W, taps, pad, CB /*Cin_pad/64*/, MB /*taps*CB*/, n_mtiles /*pairs of blocks*/, n_mt4 /*groups of four*/, n_ntiles;
  int TW, TH, TD, TN, tiles_w, tiles_h, tiles_d, tiles_n, n_chunks, n_slabs, spg, stages;
  int cb_split;   // virtual concat: input channel blocks >= cb_split are read from the second X operand (tmX2_*)
};
constexpr int kWg2Stage = 6 * kWgBlkBytes;      // X: 2 blocks x (hi, lo); dY: (hi, lo) of this CTA's 64-channel block

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
k_conv_umma_wgrad2(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                   const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                   const __grid_constant__ CUtensorMap tmX2_hi, const __grid_constant__ CUtensorMap tmX2_lo,
                   const Wg2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * kWg2Stage);
  // full / tempty are used in the leader only; empty / tfull exist (and are signalled) in both CTAs
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull0 = empty0 + 8 * S, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  constexpr uint32_t offXlo = 2 * kWgBlkBytes, offYhi = 4 * kWgBlkBytes, offYlo = 5 * kWgBlkBytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_slabs * p.n_mt4 * p.n_ntiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs: own X blocks, own dY block)
    if (elect_one()) { tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo); tma_prefetch_desc(&tmY_hi); tma_prefetch_desc(&tmY_lo); }
    uint32_t s = 0, ph = 0;
    for (int item = cid; item < n_items; item += n_clusters) {
      const int nt = item % p.n_ntiles;
      const int mt4 = (item / p.n_ntiles) % p.n_mt4;
      const int slab = item / (p.n_ntiles * p.n_mt4);
      const int mb0 = 4 * mt4 + 2 * (int)rank, mb1 = mb0 + 1;
      const bool has0 = mb0 < p.MB, has1 = mb1 < p.MB;
      // blocks of the whole pair (the leader announces the bytes of both CTAs): 4 mt4 .. 4 mt4 + 3, cut at MB
      const int nblk = min(4, p.MB - 4 * mt4);
      const uint32_t tx = 2u * (uint32_t)nblk * kWgBlkBytes + 4u * kWgBlkBytes;
      const int tap0 = has0 ? mb0 / p.CB : 0, cb0 = has0 ? mb0 % p.CB : 0, tap1 = has1 ? mb1 / p.CB : 0, cb1 = has1 ? mb1 % p.CB : 0;
      const int kd0 = p.taps == 1 ? 0 : tap0 / 9 - p.pad, kh0 = p.taps == 1 ? 0 : (tap0 / 3) % 3 - p.pad, kw0 = p.taps == 1 ? 0 : tap0 % 3 - p.pad;
      const int kd1 = p.taps == 1 ? 0 : tap1 / 9 - p.pad, kh1 = p.taps == 1 ? 0 : (tap1 / 3) % 3 - p.pad, kw1 = p.taps == 1 ? 0 : tap1 % 3 - p.pad;
      const WgRange rg = wg_range(p.n_chunks, p.n_slabs, p.spg, slab);
      for (int k = 0; k < rg.count; ++k) {
        int t = rg.first + k * rg.step;            // depth fastest (see k_conv_umma_wgrad_w3)
        const int d0 = (t % p.tiles_d) * p.TD; t /= p.tiles_d;
        const int w0 = (t % p.tiles_w) * p.TW; t /= p.tiles_w;
        const int h0 = (t % p.tiles_h) * p.TH; t /= p.tiles_h;
        const int n0 = t * p.TN;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          const uint32_t sb = smem_u32(smem) + s * (uint32_t)kWg2Stage, fb = mapa_u32(full0 + 8 * s, 0);
          if (leader) mbar_expect_tx(full0 + 8 * s, tx);
          const bool s0 = cb0 >= p.cb_split, s1 = cb1 >= p.cb_split;       // virtual concat: second X operand
          const int cc0 = (s0 ? cb0 - p.cb_split : cb0) * 64, cc1 = (s1 ? cb1 - p.cb_split : cb1) * 64;
          if (has0) {
            tma_load_5d_2sm(sb, s0 ? &tmX2_hi : &tmX_hi, fb, cc0, w0 + kw0, h0 + kh0, d0 + kd0, n0);
            tma_load_5d_2sm(sb + offXlo, s0 ? &tmX2_lo : &tmX_lo, fb, cc0, w0 + kw0, h0 + kh0, d0 + kd0, n0);
          }
          if (has1) {
            tma_load_5d_2sm(sb + kWgBlkBytes, s1 ? &tmX2_hi : &tmX_hi, fb, cc1, w0 + kw1, h0 + kh1, d0 + kd1, n0);
            tma_load_5d_2sm(sb + offXlo + kWgBlkBytes, s1 ? &tmX2_lo : &tmX_lo, fb, cc1, w0 + kw1, h0 + kh1, d0 + kd1, n0);
          }
          tma_load_5d_2sm(sb + offYhi, &tmY_hi, fb, nt * 128 + (int)rank * 64, w0, h0, d0, n0);
          tma_load_5d_2sm(sb + offYlo, &tmY_lo, fb, nt * 128 + (int)rank * 64, w0, h0, d0, n0);
        }
        __syncwarp();
        if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------------- MMA issuer of the pair
      const uint32_t idesc1 = umma_idesc_2sm(256) | (1u << 15) | (1u << 16), idesc2 = umma_idesc_2sm(128) | (1u << 15) | (1u << 16);
      uint32_t s = 0, ph = 0, tcount = 0;
      for (int item = cid; item < n_items; item += n_clusters, ++tcount) {
        const int slab = item / (p.n_ntiles * p.n_mt4);
        const int c_end = wg_range(p.n_chunks, p.n_slabs, p.spg, slab).count;
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        for (int ch = 0; ch < c_end; ++ch) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)kWg2Stage;
            const uint64_t a_hi = umma_desc(sb, kWgBlkBytes, 1024), a_lo = umma_desc(sb + offXlo, kWgBlkBytes, 1024);
            const uint64_t b_d = umma_desc(sb + offYhi, kWgBlkBytes, 1024);    // [dY_hi r | dY_lo r]: lo is LBO behind hi
#pragma unroll
            for (int k = 0; k < kWgKV / 16; ++k) {      // 16 voxel rows (2048 B) per MMA
              const uint64_t adv = (uint64_t)(k * (2048 >> 4));
              umma_bf16_2sm(d_tmem, a_hi + adv, b_d + adv, idesc1, (ch | k) ? 1u : 0u);
              umma_bf16_2sm(d_tmem + 64u, a_lo + adv, b_d + adv, idesc2, 1u);
            }
            umma_commit_2sm(empty0 + 8 * s);
            if (ch == c_end - 1) umma_commit_2sm(tfull0 + 8 * acc);
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs): own 128 rows -> workspace
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tempty_leader = mapa_u32(tempty0, 0);
    uint32_t tcount = 0;
    for (int item = cid; item < n_items; item += n_clusters, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int nt = item % p.n_ntiles;
      const int mt4 = (item / p.n_ntiles) % p.n_mt4;
      const int slab = item / (p.n_ntiles * p.n_mt4);
      const int mt = 2 * mt4 + (int)rank;                                  // this CTA's pair of blocks = one M tile of the workspace
      const bool valid = mt < p.n_mtiles && (2 * mt + row / 64) < p.MB;
      float* out = p.ws + ((((long long)slab * p.n_mtiles + mt) * p.n_ntiles + nt) * 128 + row) * 128;
      mbar_wait(tfull0 + 8 * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
      for (int c0 = 0; c0 < 128; c0 += 16) {
        const uint32_t col = (uint32_t)(c0 < 64 ? c0 : c0 + 64);            // channels 0-63: [c | 64 + c]; 64-127: [128 + c' | 192 + c']
        uint32_t r[16], r2[16];
        tmem_ld16(taddr + col, r);
        tmem_ld16(taddr + col + 64u, r2);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + c0 + j) =
                make_float4(__uint_as_float(r[j]) + __uint_as_float(r2[j]), __uint_as_float(r[j + 1]) + __uint_as_float(r2[j + 1]),
                            __uint_as_float(r[j + 2]) + __uint_as_float(r2[j + 2]), __uint_as_float(r[j + 3]) + __uint_as_float(r2[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512u);
}

// ------------------------------------------------------------------------------------------------ wgrad, kw-reuse (Cout <= 64)
// For the 64-channel layers of the 80^3 / 40^3 levels the generic kernel above is bound by the L2 -> shared-memory fill:
// an N = 64 tile needs 48 KB of operands per 384 tensor-pipe cycles (125 B/cycle/SM).  Here one work item owns up to SIX
// 64-row blocks of dW — the three kw taps of two "sources" (kd, kh, 64-channel block) — over a slab of voxel chunks:
//   * a chunk is 8(w) x 8(h) voxels of one (n, d) plane, its rows ordered [w][h] (tensor map with H before W), so that a
//     kw shift is exactly one 8-row / 1024-byte swizzle group: ONE halo box of 10 x 8 voxels per source serves its three
//     taps as three descriptor start addresses 1024 B apart;
//   * the dY tile of the chunk is loaded once and reused by all six blocks;
//   * the six blocks form three M = 128 tiles (blocks b, b+1 are LBO bytes apart: 1024 inside a source, 8192 across the
//     two), each accumulating X_hi^T x [dY_hi | dY_lo] (N = 128) + X_lo^T x dY_hi (N = 64) in its own 128 TMEM columns.
// 56 KB of operands per 1152 tensor-pipe cycles = 50 B/cycle/SM.  Partial tiles go to the workspace
// [item][tile][128][64] and k_wgrad_w3_reduce sums the slabs in a fixed order (deterministic).
constexpr int kW3XBox = 10 * 8 * 128;           // one halo box: 10 w-columns x 8 h-rows x 128 B
constexpr int kW3YBox = 8 * 8 * 128;            // one dY tile
constexpr int kW3Stage = 4 * kW3XBox + 2 * kW3YBox;
constexpr int kW3Window = 8;                    // chunks a single-source item may run ahead of the item it follows

struct Wg3Params {
  float* ws;
  int* prog;      // [n_items] chunks issued so far per work item (pacing of the single-source items, see below)
  int N, D, H, W, CB, n_src, n_pairs, tiles_w, tiles_h, n_chunks, n_slabs, spg, stages, dfast, y_lo;
  int cb_split;   // virtual concat: input channel blocks >= cb_split are read from the second X operand (tmX2_*)
  int dbg;        // DRAM_CONV_DBG (diagnostics, garbage results): bit 0 = no X loads after the first ring fill, bit 1 = no dY loads
};

// Chunk schedule of the kw-reuse wgrad.  The voxel chunks are split over `n_slabs` slabs (split-K); a work item is
// (slab, source pair) and item i runs on CTA i % gridDim.x.  Round 1 gave every slab a CONTIGUOUS range of chunks: ~10-30
// slabs were in flight at once, each streaming its own region, and the neighbouring (w,h) columns of a slab (halo re-use)
// were a whole column of chunks x all live slabs apart — far beyond the L2.  ncu: 2.4-2.6x the operand bytes from DRAM.
// Now the slabs of a GROUP (`spg` consecutive slabs ~ one wave of CTAs) INTERLEAVE over the group's contiguous range:
// slab r of the group takes chunks r, r + spg, r + 2 spg, ...  All CTAs of a wave sweep one common front through the
// volume, so the dY tile shared by the pairs of a slab, the kd-1/kd/kd+1 planes shared by its sources and the w/h halos
// shared with the neighbouring slabs are all a few chunk steps apart.
typedef WgRange Wg3Range;
__device__ __forceinline__ Wg3Range wg3_range(const Wg3Params& p, int slab) { return wg_range(p.n_chunks, p.n_slabs, p.spg, slab); }

__global__ void __launch_bounds__(kFwdThreads, 1)
k_conv_umma_wgrad_w3(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                     const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                     const __grid_constant__ CUtensorMap tmX2_hi, const __grid_constant__ CUtensorMap tmX2_lo,
                     const Wg3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * kW3Stage);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), tfull = smem_u32(bars + 2 * S),
                 tempty = smem_u32(bars + 2 * S + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t offXlo = 2 * kW3XBox, offYhi = 4 * kW3XBox, offYlo = 4 * kW3XBox + kW3YBox;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_slabs * p.n_pairs;

  if (warp == 0) {
    // whole warp converged, one elected lane issues (see elect_one)
    if (elect_one()) { tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo); tma_prefetch_desc(&tmY_hi); if (p.y_lo) tma_prefetch_desc(&tmY_lo); }
    {
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pair = item % p.n_pairs, slab = item / p.n_pairs;
        const int nsrc = (2 * pair + 1 < p.n_src) ? 2 : 1;
        const uint32_t tx = 2u * (uint32_t)nsrc * kW3XBox + (p.y_lo ? 2u : 1u) * kW3YBox;
        const int s0 = 2 * pair, s1 = min(2 * pair + 1, p.n_src - 1);
        const int cb0 = s0 % p.CB, kh0 = (s0 / p.CB) % 3 - 1, kd0 = s0 / (3 * p.CB) - 1;
        const int cb1 = s1 % p.CB, kh1 = (s1 / p.CB) % 3 - 1, kd1 = s1 / (3 * p.CB) - 1;
        const Wg3Range rg = wg3_range(p, slab);
        // Pacing.  With an odd number of sources the last pair of a slab has ONE source: 2 MMA tiles per chunk instead of 3,
        // so its CTA runs 1.5x faster than its slab mates and reads dY and its X planes from DRAM long before they do — the
        // lines are evicted again by the time the others arrive (measured: +1 x dY + 1.25 x X[source] per launch).  It now
        // follows the pair before it: that item publishes its chunk count, this one stays at most kW3Window chunks ahead.
        // The followed item has a lower index (same or earlier wave) and never waits itself; the wait is bounded, so a CTA
        // that is not resident yet can only cost locality, never progress.
        const bool paced = (p.n_src & 1) && pair == p.n_pairs - 1 && p.n_pairs > 1;
        const bool paces = (p.n_src & 1) && pair == p.n_pairs - 2;
        volatile int* prog = p.prog;
        bool pacing = paced;
        for (int k = 0; k < rg.count; ++k) {
          const int ch = rg.first + k * rg.step;
          if (paces && (k & 3) == 0 && lane == 0) prog[item] = k;
          if (pacing && (k & 3) == 0) {
            uint32_t spins = 0;
            while (prog[item - 1] + kW3Window < k) {
              __nanosleep(128);
              if (++spins > (1u << 14)) { pacing = false; break; }
            }
          }
          // chunk order: depth fastest.  The kd = 0,1,2 sources of a chunk read planes d-1, d, d+1; walking d first puts
          // their re-use one or two chunks of the common front apart (L2 hits) instead of a whole plane of chunks apart.
          int t = ch, w0, h0, d0, n0;
          if (p.dfast) {
            d0 = t % p.D; t /= p.D;
            w0 = (t % p.tiles_w) * 8; t /= p.tiles_w;
            h0 = (t % p.tiles_h) * 8;
            n0 = t / p.tiles_h;
          } else {
            w0 = (t % p.tiles_w) * 8; t /= p.tiles_w;
            h0 = (t % p.tiles_h) * 8; t /= p.tiles_h;
            d0 = t % p.D;
            n0 = t / p.D;
          }
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          if (elect_one()) {                         // map dims (C,H,W,D,N): the box starts at w0 - 1 (kw = 0)
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)kW3Stage, fb = full0 + 8 * s;
            const bool warm = ph || item != (int)blockIdx.x;                  // DRAM_CONV_DBG: the ring has been filled once
            const bool skipx = (p.dbg & 1) && warm, skipy = (p.dbg & 2) && warm;
            mbar_expect_tx(fb, (skipx ? 0u : 2u * (uint32_t)nsrc * kW3XBox) + (skipy ? 0u : (p.y_lo ? 2u : 1u) * kW3YBox));
            const bool v0 = cb0 >= p.cb_split, v1 = cb1 >= p.cb_split;       // virtual concat: second X operand
            const int cc0 = (v0 ? cb0 - p.cb_split : cb0) * 64, cc1 = (v1 ? cb1 - p.cb_split : cb1) * 64;
            if (!skipx) {
              tma_load_5d(sb, v0 ? &tmX2_hi : &tmX_hi, fb, cc0, h0 + kh0, w0 - 1, d0 + kd0, n0);
              tma_load_5d(sb + offXlo, v0 ? &tmX2_lo : &tmX_lo, fb, cc0, h0 + kh0, w0 - 1, d0 + kd0, n0);
              if (nsrc == 2) {
                tma_load_5d(sb + kW3XBox, v1 ? &tmX2_hi : &tmX_hi, fb, cc1, h0 + kh1, w0 - 1, d0 + kd1, n0);
                tma_load_5d(sb + offXlo + kW3XBox, v1 ? &tmX2_lo : &tmX_lo, fb, cc1, h0 + kh1, w0 - 1, d0 + kd1, n0);
              }
            }
            if (!skipy) {
              tma_load_5d(sb + offYhi, &tmY_hi, fb, 0, h0, w0, d0, n0);
              if (p.y_lo) tma_load_5d(sb + offYlo, &tmY_lo, fb, 0, h0, w0, d0, n0);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
        if (paces && lane == 0) prog[item] = 0x3fffffff;        // done: the follower runs free
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = umma_idesc(64, 1, 1), idesc2 = umma_idesc(128, 1, 1);
      // block b = src*3 + kw sits at src*kW3XBox + kw*1024; tile t = blocks (2t, 2t+1):
      //   tile 0 at 0 (LBO 1024), tile 1 at 2048 (LBO kW3XBox - 2048), tile 2 at kW3XBox + 1024 (LBO 1024)
      uint32_t s = 0, ph = 0, icount = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++icount) {
        const int pair = item % p.n_pairs, slab = item / p.n_pairs;
        const int ntiles = (2 * pair + 1 < p.n_src) ? 3 : 2;
        const int c_begin = 0, c_end = wg3_range(p, slab).count;
        mbar_wait(tempty, (icount & 1) ^ 1);
        tc_fence_after();
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_u32(smem) + s * (uint32_t)kW3Stage;
            const uint64_t b_hi = umma_desc(sb + offYhi, kW3YBox, 1024);    // [dY_hi | dY_lo]: lo is LBO = 8192 B behind hi
            const uint32_t first = ch > c_begin ? 1u : 0u;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              if (t < ntiles) {
                const uint32_t toff = t == 0 ? 0u : (t == 1 ? 2048u : (uint32_t)kW3XBox + 1024u);
                const uint32_t tlbo = t == 1 ? (uint32_t)kW3XBox - 2048u : 1024u;
                const uint64_t a_hi = umma_desc(sb + toff, tlbo, 1024);
                const uint64_t a_lo = umma_desc(sb + offXlo + toff, tlbo, 1024);
                const uint32_t d_tmem = tmem_base + (uint32_t)t * 128u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {           // 16 voxel rows (two w columns, 2048 B) per MMA
                  const uint64_t adv = (uint64_t)(k * (2048 >> 4));
                  umma_bf16(d_tmem, a_hi + adv, b_hi + adv, p.y_lo ? idesc2 : idesc, k ? 1u : first);   // N = 128: [dY_hi | dY_lo]
                  umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                }
              }
            }
            umma_commit(empty0 + 8 * s);
            if (ch == c_end - 1) umma_commit(tfull);
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t icount = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++icount) {
      const int pair = item % p.n_pairs;
      const int ntiles = (2 * pair + 1 < p.n_src) ? 3 : 2;
      mbar_wait(tfull, icount & 1);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) {
        float* out = p.ws + (((long long)item * 3 + t) * 128 + row) * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 128u;
        const bool valid = (t < 2) || (ntiles == 3);         // the last tile of a single-source item has no second block
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16], r2[16];
          tmem_ld16(taddr + c0, r);
          if (p.y_lo) tmem_ld16(taddr + 64 + c0, r2);
          tmem_ld_wait();
          if (!p.y_lo) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r2[j] = 0u;                       // +0.0f
          }
          if (valid || row < 64) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(out + c0 + j) =
                  make_float4(__uint_as_float(r[j]) + __uint_as_float(r2[j]), __uint_as_float(r[j + 1]) + __uint_as_float(r2[j + 1]),
                              __uint_as_float(r[j + 2]) + __uint_as_float(r2[j + 2]), __uint_as_float(r[j + 3]) + __uint_as_float(r2[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// dw[co][ci][tap] = sum_slab ws[slab*n_pairs + pair][tile][row][co]
__global__ void k_wgrad_w3_reduce(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int CB, int n_pairs,
                                  int n_slabs) {
  const long long total = 27ll * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int ci = (int)(r % Cin), tap = (int)(r / Cin);
    const int kw = tap % 3, g = tap / 3;                 // g = kd*3 + kh
    const int s = g * CB + ci / 64;
    const int pair = s >> 1, b = (s & 1) * 3 + kw;
    const int tile = b >> 1, row = (b & 1) * 64 + (ci & 63);
    const float* src = ws + (((long long)pair * 3 + tile) * 128 + row) * 64 + co;
    const long long slab_stride = (long long)n_pairs * 3 * 128 * 64;
    float acc = 0.f;
    for (int sl = 0; sl < n_slabs; ++sl) acc += src[sl * slab_stride];
    dw[((long long)co * Cin + ci) * 27 + tap] = acc;
  }
}

// dw[co][ci][tap] = sum_slab ws[slab][mt][nt][row][n]   (fixed summation order => deterministic)
__global__ void k_wgrad_reduce(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int taps, int CB,
                               int n_mtiles, int n_ntiles, int BN, int n_slabs) {
  const long long total = (long long)taps * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int ci = (int)(r % Cin), tap = (int)(r / Cin);
    const int mb = tap * CB + ci / 64;
    const int mt = mb >> 1, row = (mb & 1) * 64 + (ci & 63), nt = co / BN, n = co % BN;
    float acc = 0.f;
    for (int s = 0; s < n_slabs; ++s)
      acc += ws[((((long long)s * n_mtiles + mt) * n_ntiles + nt) * 128 + row) * BN + n];
    dw[((long long)co * Cin + ci) * taps + tap] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ operand preparation
// x fp32 [rows][C] -> hi/lo bf16 [rows][Cpad]; 8 channels (two float4 in, one 16-byte store per plane) per thread
__global__ void k_split_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                             long long rows, int C, int Cpad) {
  const int groups = Cpad / 8;
  const long long total = rows * groups;
  const bool vec = (C % 8 == 0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long r = i / groups;
    float v[8];
    const int c0 = g * 8;
    if (vec && c0 + 8 <= C) {
      const float4* p = reinterpret_cast<const float4*>(x + r * C + c0);
      float4 a = __ldg(p), b = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? x[r * C + c0 + j] : 0.f;
    }
    __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __float2bfloat16_rn(v[j]);
      l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint4*>(hi + r * Cpad + c0) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + r * Cpad + c0) = *reinterpret_cast<const uint4*>(l);
  }
}

// w [Cout][Cin][T] -> mode 0: out[t][co][k = ci]  (rows = Cout, K = Cin  -> Kpad)
//                     mode 1: out[t][ci][k = co]  (rows = Cin,  K = Cout -> Kpad), taps flipped
__global__ void k_pack_weight_bf16(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                   int Cout, int Cin, int Kpad, int T, int mode) {
  const int R = mode == 0 ? Cout : Cin, K = mode == 0 ? Cin : Cout;
  const long long total = (long long)T * R * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    long long r = i / Kpad;
    const int row = (int)(r % R), t = (int)(r / R);
    float v = 0.f;
    if (k < K) v = mode == 0 ? w[((long long)row * Cin + k) * T + t] : w[((long long)k * Cin + row) * T + (T - 1 - t)];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 5-D map over a channels-last bf16 volume [N][D][H][W][C]; box = (64 channels, bw, bh, bd, bn), 128-byte swizzle
static int make_volume_map(CUtensorMap* m, const void* base, int N, int D, int H, int W, int C, int bw, int bh, int bd, int bn,
                           bool h_fastest = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DRAM_E_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, (cuuint32_t)bn};
  if (h_fastest) {       // dimension order (C,H,W,D,N): box rows land in shared memory ordered [w][h]
    dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)W;
    strides[0] = (cuuint64_t)W * C * 2; strides[1] = (cuuint64_t)C * 2;
    box[1] = (cuuint32_t)bh; box[2] = (cuuint32_t)bw;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(volume %dx%dx%dx%dx%d box %dx%dx%dx%d) failed: %d", N, D, H, W, C, bn, bd, bh, bw, (int)r); return DRAM_E_CUDA; }
  return DRAM_OK;
}
// 2-D map over packed weights [rows][K] bf16; box = (64, box_rows)
static int make_weight_map(CUtensorMap* m, const void* base, long long rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DRAM_E_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %lldx%d box %d) failed: %d", rows, K, box_rows, (int)r); return DRAM_E_CUDA; }
  return DRAM_OK;
}

// 5-D map with dimension order (C,H,D,W,N) over the same channels-last volume: box = (64 channels, 8 h, bd, bw, 1);
// shared-memory rows are ordered [w][d][h] (k_conv_umma_fwd2)
static int make_volume_map_hdw(CUtensorMap* m, const void* base, int N, int D, int H, int W, int C, int bw, int bd) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DRAM_E_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)W, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {64, 8, (cuuint32_t)bd, (cuuint32_t)bw, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(volume hdw %dx%dx%dx%dx%d box %dx%d) failed: %d", N, D, H, W, C, bd, bw, (int)r); return DRAM_E_CUDA; }
  return DRAM_OK;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// output tile (TW,TH,TD), TW*TH*TD <= 128, maximising the fraction of useful MMA rows
static void pick_fwd_tile(int D, int H, int W, int& TW, int& TH, int& TD) {
  double best = -1.0;
  TW = TH = TD = 1;
  for (int tw = 1; tw <= 128 && tw <= W; ++tw)
    for (int th = 1; tw * th <= 128 && th <= H; ++th) {
      int td = 128 / (tw * th);
      if (td > D) td = D;
      if (td < 1) continue;
      double eff = ((double)W * H * D) / ((double)cdiv(W, tw) * cdiv(H, th) * cdiv(D, td) * 128.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && (tw > TW || (tw == TW && th > TH)))) { best = eff; TW = tw; TH = th; TD = td; }
    }
}
// K chunk (TW,TH,TD,TN) with product exactly 64 (powers of two): least zero-filled rows first, then the most COMPACT box —
// smallest halo volume TN*(TW+2)*(TH+2)*(TD+2), i.e. the 27 shifted boxes of a chunk and the boxes of the neighbouring chunks
// overlap as much as possible in L2.  (Round 1 broke ties by the largest TW and ended up with 8(w) x 1 x 1 voxels x 8 SAMPLES
// at the 40^3 level: every chunk touched eight volumes, the kh / kd neighbours of a row were hundreds of chunk steps away
// and us1.c0 read 4.1x its operands from DRAM; 4 x 4 x 4 voxels of one sample now.)
static void pick_wgrad_chunk(int N, int D, int H, int W, int& TW, int& TH, int& TD, int& TN) {
  double best = 1e300, best_halo = 1e300;
  TW = 64; TH = TD = TN = 1;
  for (int tw = 1; tw <= 64; tw *= 2)
    for (int th = 1; tw * th <= 64; th *= 2)
      for (int td = 1; tw * th * td <= 64; td *= 2) {
        int tn = 64 / (tw * th * td);
        double vol = (double)cdiv(W, tw) * tw * cdiv(H, th) * th * (double)cdiv(D, td) * td * cdiv(N, tn) * tn;
        double halo = (double)tn * (tw + 2) * (th + 2) * (td + 2);
        if (vol < best - 0.5 || (vol < best + 0.5 && halo < best_halo - 0.5)) {
          best = vol; best_halo = halo; TW = tw; TH = th; TD = td; TN = tn;
        }
      }
}
static int pick_bn(int Cout) {
  const int cand[6] = {128, 96, 64, 48, 32, 16};
  for (int i = 0; i < 6; ++i)
    if (Cout % cand[i] == 0) return cand[i];
  return 0;
}
static int pow2_cols(int c) { int v = 32; while (v < c) v *= 2; return v; }

constexpr int kSmemBudget = 200 * 1024;

// which forward / dgrad kernel runs a layer (shared by the launcher and by the query for the BatchNorm partial rows)
enum { kFwdGeneric = 0, kFwdPairs = 2, kFwdChannelsOnM = 3, kFwdSmPairs = 4, kFwdSmPairsGen = 5 };
// SM-pair kernel (k_conv_umma_fwd4): 128-channel N tiles plus a 64-channel tail tile
static int fwd_kernel_kind(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, bool x_lo, bool w_lo) {
  // DRAM_CONV_V4=0 switches it off.  It takes every split-bf16 3x3x3 layer whose volume tiles into 8(h) x TDD(d) x TW(w) boxes
  // and whose Cout is a multiple of 64 (interleaved A/B on one box, profiles/r02c_pairs.txt: 11-18 % faster than the single-SM
  // kernels on all of them), except a single short K block (Cin <= 32: ds0.c1, 2 K steps per tap — too little MMA time per
  // halo box for its two activation stages)
  {
    const char* v4_env = getenv("DRAM_CONV_V4");
    const bool allow_v4 = !(v4_env && atoi(v4_env) == 0);
    if (allow_v4 && Cout % 32 == 0 && x_lo && w_lo && ksize == 3 && H % 8 == 0 && (W % 16 == 0 || (W % 8 == 0 && D % 2 == 0)) &&
        !(Cin_pad == 64 && Cin <= 32)) {
      const int TW = (W % 16 == 0) ? 16 : 8, TDD = (W % 16 == 0) ? 1 : 2;
      const long long mt = (long long)N * (D / TDD) * (H / 8) * (W / TW);
      if (mt % 2 == 0) return kFwdSmPairs;
    }
    // the same kernel on generic (TW,TH,TD) tiles for volumes that do not tile into 8-row boxes (the 20^3 / 10^3 levels):
    // Cout a multiple of 128 only (DRAM_CONV_V4=1 also allows a 64-channel tail)
    if (allow_v4 && Cout % 64 == 0 && (Cout % 128 == 0 || (v4_env && atoi(v4_env) == 1)) && x_lo && w_lo && ksize == 3 &&
        !(H % 8 == 0 && (W % 16 == 0 || (W % 8 == 0 && D % 2 == 0)))) {
      int TW, TH, TD;
      pick_fwd_tile(D, H, W, TW, TH, TD);
      const long long mt = (long long)N * cdiv(D, TD) * cdiv(H, TH) * cdiv(W, TW);
      if (mt % 2 == 0 && mt >= 2) return kFwdSmPairsGen;
    }
  }
  // channels-on-M kernel (k_conv_umma_fwd3) for split-bf16 layers with 64- or 128-channel output tiles
  // DRAM_CONV_V3: 0 = never, 1 = wherever it applies, unset = where it measured faster in an interleaved A/B on one box
  // (profiles/r01b_conv_fwd2_vs_fwd3.txt): 128-channel tiles, or a single 64-channel tile with at most two K blocks per tap
  const char* v3_env = getenv("DRAM_CONV_V3");
  int use_v3 = v3_env ? atoi(v3_env) : ((Cout % 128 == 0 || (Cout == 64 && (Cin_pad <= 128 || !x_lo))) ? 1 : 0);
  if (use_v3 == 2) use_v3 = (Cout % 128 == 0) ? 1 : 0;       // experiment: 128-channel tiles only
  if (use_v3 && w_lo && ksize == 3 && H % 8 == 0 && Cout % 64 == 0 &&
      ((W % 16 == 0 && D % 2 == 0) || (W % 8 == 0 && D % 4 == 0)))
    return kFwdChannelsOnM;
  // weight-sharing tile pairs (k_conv_umma_fwd2) wherever the volume tiles into 8(h) x TDD(d) x TW(w) boxes
  static const bool allow_v2 = getenv("DRAM_CONV_NO_V2") == nullptr;
  if (allow_v2 && ksize == 3 && H % 8 == 0 && (W % 16 == 0 || (W % 8 == 0 && D % 2 == 0)) && Cout % 32 == 0) return kFwdPairs;
  return kFwdGeneric;
}
// rows of the BatchNorm partial-sum buffer [rows][2][Cout] the kernel's epilogue fills (0: this kernel has no such epilogue)
static long long fwd_stat_rows(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, bool x_lo, bool w_lo) {
  const int kind = fwd_kernel_kind(N, D, H, W, Cin, Cin_pad, Cout, ksize, x_lo, w_lo);
  if (kind == kFwdChannelsOnM) {
    const int TW = (W % 16 == 0 && D % 2 == 0) ? 16 : 8, TDD = (W % 16 == 0 && D % 2 == 0) ? 2 : 4;
    const long long vt = (long long)N * (D / TDD) * (H / 8) * (W / TW);
    return (Cout % 128 == 0) ? vt : 2 * vt;
  }
  if (kind == kFwdPairs || kind == kFwdSmPairs) {
    const int TW = (W % 16 == 0) ? 16 : 8, TDD = (W % 16 == 0) ? 1 : 2;
    return 4ll * N * (D / TDD) * (H / 8) * (W / TW);
  }
  return 0;
}

// sums[c] = sum over rows of partial[row][0][c], sums[C + c] = ... [1][c], in double, in a fixed order (deterministic):
// level 1, grid (column blocks of 32, kStatChunks row chunks): a warp owns a column block, its lanes the columns, the 8 warps
// of a block stride the chunk's rows -> ws[chunk][2C]; level 2 adds the chunks per column.
constexpr int kStatChunks = 192;
__global__ void __launch_bounds__(256)
k_bn_partials_reduce(const float* __restrict__ partial, long long rows, int C, double* __restrict__ ws) {
  __shared__ double sh[8][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;                   // column of the [rows][2C] matrix
  const long long rpc = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rpc, r1 = r0 + rpc < rows ? r0 + rpc : rows;
  double a = 0.0;
  if (col < 2 * C) {
    const float* p = partial + col;
    long long r = r0 + wp;
    for (; r + 24 < r1; r += 32) {                          // four independent loads in flight
      const float v0 = __ldg(p + r * 2 * C), v1 = __ldg(p + (r + 8) * 2 * C), v2 = __ldg(p + (r + 16) * 2 * C),
                  v3 = __ldg(p + (r + 24) * 2 * C);
      a += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; r < r1; r += 8) a += (double)__ldg(p + r * 2 * C);
  }
  sh[wp][lane] = a;
  __syncthreads();
  if (wp == 0 && col < 2 * C) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][lane];
    ws[(long long)blockIdx.y * 2 * C + col] = t;
  }
}
// level 2: a warp per column, lanes stride the chunks, fixed shuffle tree (deterministic); the one-thread-per-column version
// walked the 192 chunks serially: 21-24 us per layer for 6 KB of data
__global__ void __launch_bounds__(256)
k_bn_partials_finish(const double* __restrict__ ws, int chunks, int C, double* __restrict__ sums) {
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (col >= 2 * C) return;
  double t = 0.0;
  for (int k = lane; k < chunks; k += 32) t += ws[(long long)k * 2 * C + col];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) sums[col] = t;
}

}  // namespace dram

using namespace dram;

extern "C" {

int dram_split_bf16(const float* x, void* hi, void* lo, long long rows, int C, int Cpad, void* stream) {
  DRAM_REQUIRE(x && hi && rows > 0 && C > 0 && Cpad >= C && Cpad % 8 == 0, "split_bf16: bad arguments (C=%d Cpad=%d)", C, Cpad);
  k_split_bf16<<<grid_for(rows * (Cpad / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, C, Cpad);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_pack_weight_bf16(const float* w, void* w_hi, void* w_lo, int Cout, int Cin, int Kpad, int ksize, int mode,
                          void* stream) {
  DRAM_REQUIRE(w && w_hi && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3) && (mode == 0 || mode == 1), "pack_weight_bf16: bad arguments");
  DRAM_REQUIRE(Kpad >= (mode == 0 ? Cin : Cout) && Kpad % 64 == 0, "pack_weight_bf16: Kpad=%d must be a multiple of 64 covering K", Kpad);
  const int T = ksize * ksize * ksize;
  long long total = (long long)T * (mode == 0 ? Cout : Cin) * Kpad;
  k_pack_weight_bf16<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo,
                                                                            Cout, Cin, Kpad, T, mode);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

long long dram_conv3d_umma_fwd_stat_rows(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, int has_x_lo, int has_w_lo) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin_pad < Cin || Cin_pad % 64 || Cout <= 0) return 0;
  return fwd_stat_rows(N, D, H, W, Cin, Cin_pad, Cout, ksize, has_x_lo != 0, has_w_lo != 0);
}

int dram_conv3d_umma_fwd_kernel(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, int has_x_lo, int has_w_lo) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin_pad < Cin || Cin_pad % 64 || Cout <= 0) return -1;
  return fwd_kernel_kind(N, D, H, W, Cin, Cin_pad, Cout, ksize, has_x_lo != 0, has_w_lo != 0);
}

size_t dram_bn_stats_from_partials_workspace_bytes(int C) { return C > 0 ? sizeof(double) * 2 * (size_t)C * kStatChunks : 0; }

int dram_bn_stats_from_partials(const float* partials, long long rows, int C, double* sums, void* workspace, void* stream) {
  DRAM_REQUIRE(partials && sums && workspace && rows > 0 && C > 0, "bn_stats_from_partials: bad arguments");
  const int chunks = rows < kStatChunks ? (int)rows : kStatChunks;
  cudaStream_t st = (cudaStream_t)stream;
  k_bn_partials_reduce<<<dim3((2 * C + 31) / 32, chunks), 256, 0, st>>>(partials, rows, C, (double*)workspace);
  DRAM_LAUNCH_CHECK();
  k_bn_partials_finish<<<(2 * C + 7) / 8, 256, 0, st>>>((const double*)workspace, chunks, C, sums);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_conv3d_umma_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* scale,
                         const float* shift, float* y, void* out_hi, void* out_lo, float* bn_partials, const void* x2_hi,
                         const void* x2_lo, int Cin1_pad, int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize,
                         void* stream) {
  // virtual concat (UpsampleConvBlock5d, parts.py:151-155, without materialising cat([up, skip])): x holds the first Cin1_pad
  // channels (its own row pitch), x2 the remaining Cin_pad - Cin1_pad; the K loop switches tensor maps at the block boundary
  const int C1p = x2_hi ? Cin1_pad : Cin_pad, C2p = Cin_pad - C1p;
  DRAM_REQUIRE(!x2_hi || (Cin1_pad > 0 && Cin1_pad % 64 == 0 && C2p > 0 && (x2_lo != nullptr) == (x_lo != nullptr)),
               "conv3d_umma_fwd: second operand needs 0 < Cin1_pad (%d, multiple of 64) < Cin_pad (%d) and the same planes as x", Cin1_pad, Cin_pad);
  DRAM_REQUIRE(x_hi && w_hi && (y || out_hi) && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_umma_fwd: bad arguments");
  DRAM_REQUIRE(!out_hi || (scale && Cout % 64 == 0), "conv3d_umma_fwd: plane output needs scale/shift (eval mode) and Cout %% 64 == 0 (no channel padding)");
  DRAM_REQUIRE(out_hi || !out_lo, "conv3d_umma_fwd: out_lo without out_hi");
  DRAM_REQUIRE(!(x_lo != nullptr && w_lo == nullptr), "conv3d_umma_fwd: x_lo needs w_lo (modes: both = bf16x3, w_lo only = single-plane activations x split weights, neither = bf16)");
  DRAM_REQUIRE((scale == nullptr) == (shift == nullptr), "conv3d_umma_fwd: scale and shift must come together");
  DRAM_REQUIRE(ksize == 1 || ksize == 3, "conv3d_umma_fwd: kernel size %d unsupported", ksize);
  DRAM_REQUIRE(Cin_pad > 0 && Cin_pad % 64 == 0, "conv3d_umma_fwd: Cin_pad=%d must be a multiple of 64", Cin_pad);
  DRAM_REQUIRE(Cin > 0 && Cin <= Cin_pad, "conv3d_umma_fwd: Cin=%d must be in (0, Cin_pad=%d]", Cin, Cin_pad);
  // channels [Cin, Cin_pad) are zeros in x and w: with a single 64-channel block the K = 16 steps that would only multiply
  // padding are not issued (ds0.c1, Cin = 32: 2 of 4 steps)
  const int ksteps = (Cin_pad == 64 && !getenv("DRAM_CONV_FULL_K")) ? (Cin + 15) / 16 : 4;
  const int kind = fwd_kernel_kind(N, D, H, W, Cin, Cin_pad, Cout, ksize, x_lo != nullptr, w_lo != nullptr);
  DRAM_REQUIRE(!bn_partials || (!scale && fwd_stat_rows(N, D, H, W, Cin, Cin_pad, Cout, ksize, x_lo != nullptr, w_lo != nullptr) > 0),
               "conv3d_umma_fwd: bn_partials needs a raw (no scale/shift) output and a kernel with the statistics epilogue "
               "(dram_conv3d_umma_fwd_stat_rows > 0)");
  if (kind == kFwdSmPairs || kind == kFwdSmPairsGen) {
    Fwd4Params q;
    q.y = y; q.scale = scale; q.shift = shift; q.o_hi = (uint16_t*)out_hi; q.o_lo = (uint16_t*)out_lo; q.stat = bn_partials;
    q.N = N; q.D = D; q.H = H; q.W = W; q.Cout = Cout; q.kblocks_c = Cin_pad / 64; q.ksteps = ksteps;
    q.gen = kind == kFwdSmPairsGen ? 1 : 0;
    q.cb_split = C1p / 64;
    CUtensorMap mA_hi, mA_lo, mA2_hi, mA2_lo;
    int rc4;
    if (q.gen) {
      pick_fwd_tile(D, H, W, q.TW, q.TH, q.TD);
      q.TDD = 1;
      q.tiles_w = cdiv(W, q.TW); q.tiles_h = cdiv(H, q.TH); q.tiles_d = cdiv(D, q.TD);
      q.a_plane_bytes = kATileBytes;                      // 128 rows reserved, TW*TH*TD of them written by TMA
      q.a_box_bytes = q.TW * q.TH * q.TD * 128;
      if ((rc4 = make_volume_map(&mA_hi, x_hi, N, D, H, W, C1p, q.TW, q.TH, q.TD, 1))) return rc4;
      if ((rc4 = make_volume_map(&mA_lo, x_lo, N, D, H, W, C1p, q.TW, q.TH, q.TD, 1))) return rc4;
      mA2_hi = mA_hi; mA2_lo = mA_lo;
      if (x2_hi && (rc4 = make_volume_map(&mA2_hi, x2_hi, N, D, H, W, C2p, q.TW, q.TH, q.TD, 1))) return rc4;
      if (x2_lo && (rc4 = make_volume_map(&mA2_lo, x2_lo, N, D, H, W, C2p, q.TW, q.TH, q.TD, 1))) return rc4;
    } else {
      if (W % 16 == 0) { q.TW = 16; q.TDD = 1; } else { q.TW = 8; q.TDD = 2; }
      q.TH = 8; q.TD = q.TDD;
      q.tiles_w = W / q.TW; q.tiles_h = H / 8; q.tiles_d = D / q.TDD;
      q.a_plane_bytes = (q.TW + 2) * q.TDD * 1024;
      q.a_box_bytes = q.a_plane_bytes;
      if ((rc4 = make_volume_map_hdw(&mA_hi, x_hi, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc4;
      if ((rc4 = make_volume_map_hdw(&mA_lo, x_lo, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc4;
      mA2_hi = mA_hi; mA2_lo = mA_lo;
      if (x2_hi && (rc4 = make_volume_map_hdw(&mA2_hi, x2_hi, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc4;
      if (x2_lo && (rc4 = make_volume_map_hdw(&mA2_lo, x2_lo, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc4;
    }
    q.n_mtiles = N * q.tiles_d * q.tiles_h * q.tiles_w;
    q.a_tile_bytes = 2 * q.a_plane_bytes;
    static int max_clusters = 0;
    static std::once_flag once4;
    std::call_once(once4, [&] {
      cudaFuncSetAttribute(k_conv_umma_fwd4, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kNumSMs); cfg.blockDim = dim3(kFwdThreads); cfg.dynamicSmemBytes = 227 * 1024 - 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, k_conv_umma_fwd4, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = kNumSMs / 2; }
      max_clusters = nc < kNumSMs / 2 ? nc : kNumSMs / 2;
    });
    static long long* prof_buf4 = nullptr;
    static const bool want_prof4 = getenv("DRAM_CONV_PROF") != nullptr;
    if (want_prof4 && !prof_buf4) cudaMalloc(&prof_buf4, kNumSMs * 8 * sizeof(long long));
    q.prof = want_prof4 ? prof_buf4 : nullptr;
    { const char* e = getenv("DRAM_CONV_DBG"); q.dbg = e ? atoi(e) : 0; }
    // Cout = 128 * n128 (+ 64): one launch of 128-channel N tiles (both MMAs at N >= 128: the MMA pipe of the pair stays ~98 %
    // busy), one launch for a 64-channel tail tile (its N = 64 MMA costs almost as much as an N = 128 one: ~77 %)
    // a 32-channel tail (ds0.c1's dgrad, 64 -> 32) runs the same scheme at BN = 32 (N = 64 / 32 MMAs: ~48 % MMA efficiency, against
    // 33 % tensor-pipe activity of the single-SM tile-pair kernel on that layer, profiles/r02u_all_conv_launches.md #39)
    const int n128 = Cout / 128, tail = Cout % 128;
    for (int part = 0; part < 3; ++part) {
      const int bn = part == 0 ? 128 : (part == 1 ? 64 : 32);
      const bool run = part == 0 ? n128 > 0 : (part == 1 ? (tail & 64) != 0 : (tail & 32) != 0);
      if (!run) continue;
      q.BN = bn;
      q.co_base = part == 0 ? 0 : (part == 1 ? 128 * n128 : 128 * n128 + (tail & 64));
      q.n_ntiles = part == 0 ? n128 : 1;
      q.TPC = (q.BN <= 64 && q.n_mtiles % 4 == 0) ? 2 : 1;      // two M tiles per CTA share every weight tile (2 x 2 x 128 TMEM columns)
      q.n_items = q.n_mtiles / (2 * q.TPC) * q.n_ntiles;
      q.a_stage_bytes = q.TPC * q.a_tile_bytes;
      q.b_stage_bytes = q.BN * 128;                        // this CTA's half of the N = 2*BN weight columns: BN/2 hi + BN/2 lo rows
      q.SA = q.TPC == 2 ? 2 : 3;
      q.SB = (227 * 1024 - 1024 - 512 - q.SA * q.a_stage_bytes) / q.b_stage_bytes;
      if (q.SB > 8) q.SB = 8;
      DRAM_REQUIRE(q.SB >= 3, "conv3d_umma_fwd: SM-pair pipeline does not fit in shared memory (SB=%d)", q.SB);
      CUtensorMap mB_hi, mB_lo;
      if ((rc4 = make_weight_map(&mB_hi, w_hi, 27ll * Cout, Cin_pad, q.BN / 2))) return rc4;
      if ((rc4 = make_weight_map(&mB_lo, w_lo, 27ll * Cout, Cin_pad, q.BN / 2))) return rc4;
      const size_t smem4 = (size_t)q.SA * q.a_stage_bytes + (size_t)q.SB * q.b_stage_bytes + 1024 + 512;
      const int n_cl = q.n_items < max_clusters ? q.n_items : max_clusters;
      k_conv_umma_fwd4<<<2 * n_cl, kFwdThreads, smem4, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
      DRAM_LAUNCH_CHECK();
      if (want_prof4) {
        long long hbuf[kNumSMs * 8];
        cudaMemcpy(hbuf, prof_buf4, sizeof(hbuf), cudaMemcpyDeviceToHost);
        double a[6] = {0, 0, 0, 0, 0, 0};
        for (int bi = 0; bi < n_cl; ++bi) for (int j = 0; j < 6; ++j) a[j] += (double)hbuf[bi * 8 + j] / n_cl;
        fprintf(stderr, "[fwd4 prof] clusters %d items/cluster %.1f BN %d TPC %d cb %d SA %d SB %d | mma thread: total %.0f, wait tmem-empty %.0f, wait A %.0f, wait B %.0f | epilogue: wait %.0f, work %.0f (cycles)\n",
                n_cl, (double)q.n_items / n_cl, q.BN, q.TPC, q.kblocks_c, q.SA, q.SB, a[0], a[1], a[2], a[3], a[4], a[5]);
      }
    }
    return DRAM_OK;
  }
  if (kind == kFwdChannelsOnM) {
    Fwd3Params q;
    const int plain = (Cout % 128 == 0) ? 1 : 0;
    const int mode = plain + (x_lo ? 0 : 2);
    q.y = y; q.scale = scale; q.shift = shift; q.o_hi = (uint16_t*)out_hi; q.o_lo = (uint16_t*)out_lo; q.stat = bn_partials;
    q.N = N; q.D = D; q.H = H; q.W = W; q.Cout = Cout; q.CT = plain ? 128 : 64; q.kblocks_c = Cin_pad / 64; q.ksteps = ksteps;
    if (W % 16 == 0 && D % 2 == 0) { q.TW = 16; q.TDD = 2; } else { q.TW = 8; q.TDD = 4; }
    q.tiles_w = W / q.TW; q.tiles_h = H / 8; q.tiles_d = D / q.TDD;
    q.n_vtiles = N * q.tiles_d * q.tiles_h * q.tiles_w;
    q.n_ctiles = Cout / q.CT;
    q.n_items = q.n_vtiles * q.n_ctiles;
    q.x_plane_bytes = (q.TW + 2) * q.TDD * 1024;
    q.x_stage_bytes = (x_lo ? 2 : 1) * q.x_plane_bytes;
    q.w_stage_bytes = 2 * 128 * 128 / (plain ? 1 : 2);         // stacked: 64 hi + 64 lo rows; plain: 128 + 128 rows
    const int xchg = plain ? 0 : 8192;
    q.SW = (227 * 1024 - 1024 - 512 - xchg - 2 * q.x_stage_bytes) / q.w_stage_bytes;
    if (q.SW > 6) q.SW = 6;
    DRAM_REQUIRE(q.SW >= 2, "conv3d_umma_fwd: channels-on-M pipeline does not fit in shared memory");
    CUtensorMap mX_hi, mX_lo, mW_hi, mW_lo;
    int rc3;
    CUtensorMap mX2_hi, mX2_lo;
    q.cb_split = C1p / 64;
    if ((rc3 = make_volume_map_hdw(&mX_hi, x_hi, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc3;
    if (x_lo) { if ((rc3 = make_volume_map_hdw(&mX_lo, x_lo, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc3; }
    else mX_lo = mX_hi;
    mX2_hi = mX_hi; mX2_lo = mX_lo;
    if (x2_hi && (rc3 = make_volume_map_hdw(&mX2_hi, x2_hi, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc3;
    if (x2_lo && (rc3 = make_volume_map_hdw(&mX2_lo, x2_lo, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc3;
    if ((rc3 = make_weight_map(&mW_hi, w_hi, 27ll * Cout, Cin_pad, q.CT))) return rc3;
    if ((rc3 = make_weight_map(&mW_lo, w_lo, 27ll * Cout, Cin_pad, q.CT))) return rc3;
    const size_t smem3 = 2 * (size_t)q.x_stage_bytes + (size_t)q.SW * q.w_stage_bytes + xchg + 1024 + 512;
    static std::once_flag once3;
    std::call_once(once3, [] {
      cudaFuncSetAttribute(k_conv_umma_fwd3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    const int grid3 = q.n_items < kNumSMs ? q.n_items : kNumSMs;
    static long long* prof_buf3 = nullptr;
    static const bool want_prof3 = getenv("DRAM_CONV_PROF") != nullptr;
    if (want_prof3 && !prof_buf3) cudaMalloc(&prof_buf3, kNumSMs * 8 * sizeof(long long));
    q.prof = want_prof3 ? prof_buf3 : nullptr;
    { const char* e = getenv("DRAM_CONV_DBG"); q.dbg = e ? atoi(e) : 0; }
    if (mode == 0) k_conv_umma_fwd3<0><<<grid3, kFwdThreads, smem3, (cudaStream_t)stream>>>(mX_hi, mX_lo, mW_hi, mW_lo, mX2_hi, mX2_lo, q);
    else if (mode == 1) k_conv_umma_fwd3<1><<<grid3, kFwdThreads, smem3, (cudaStream_t)stream>>>(mX_hi, mX_lo, mW_hi, mW_lo, mX2_hi, mX2_lo, q);
    else if (mode == 2) k_conv_umma_fwd3<2><<<grid3, kFwdThreads, smem3, (cudaStream_t)stream>>>(mX_hi, mX_lo, mW_hi, mW_lo, mX2_hi, mX2_lo, q);
    else k_conv_umma_fwd3<3><<<grid3, kFwdThreads, smem3, (cudaStream_t)stream>>>(mX_hi, mX_lo, mW_hi, mW_lo, mX2_hi, mX2_lo, q);
    DRAM_LAUNCH_CHECK();
    if (want_prof3) {
      long long h[kNumSMs * 8];
      cudaMemcpy(h, prof_buf3, sizeof(h), cudaMemcpyDeviceToHost);
      double a[6] = {0, 0, 0, 0, 0, 0};
      for (int b = 0; b < grid3; ++b) for (int j = 0; j < 6; ++j) a[j] += (double)h[b * 8 + j] / grid3;
      fprintf(stderr, "[fwd3 prof] mode %d items/CTA %.1f Cout %d cb %d SW %d | mma warp: total %.0f, wait tmem-empty %.0f, wait X %.0f, wait W %.0f | epilogue: wait %.0f, work %.0f (cycles)\n",
              mode, (double)q.n_items / grid3, Cout, q.kblocks_c, q.SW, a[0], a[1], a[2], a[3], a[4], a[5]);
    }
    return DRAM_OK;
  }
  if (kind == kFwdPairs) {
    Fwd2Params q;
    q.BN = Cout <= 64 ? Cout : (Cout % 128 == 0 ? 128 : (Cout % 96 == 0 ? 96 : (Cout % 64 == 0 ? 64 : 32)));
    const int mode = !w_lo ? 0 : ((q.BN <= 64 ? 1 : 2) + (x_lo ? 0 : 2));
    q.acc_cols = (mode == 1 || mode == 3) ? 2 * q.BN : q.BN;
    q.tmem_cols = pow2_cols(4 * q.acc_cols);
    q.y = y; q.scale = scale; q.shift = shift; q.o_hi = (uint16_t*)out_hi; q.o_lo = (uint16_t*)out_lo; q.stat = bn_partials;
    q.N = N; q.D = D; q.H = H; q.W = W; q.Cout = Cout; q.kblocks_c = Cin_pad / 64; q.ksteps = ksteps;
    if (W % 16 == 0) { q.TW = 16; q.TDD = 1; } else { q.TW = 8; q.TDD = 2; }
    q.tiles_w = W / q.TW; q.tiles_h = H / 8; q.tiles_d = D / q.TDD;
    q.n_mtiles = N * q.tiles_d * q.tiles_h * q.tiles_w;
    q.n_ntiles = Cout / q.BN;
    q.n_items = cdiv(q.n_mtiles, 2) * q.n_ntiles;
    q.a_plane_bytes = (q.TW + 2) * q.TDD * 1024;
    q.a_tile_bytes = (x_lo ? 2 : 1) * q.a_plane_bytes;
    q.a_stage_bytes = 2 * q.a_tile_bytes;
    q.b_stage_bytes = (w_lo ? 2 : 1) * q.BN * 128;
    const int smem_max = 227 * 1024;
    q.SB = (smem_max - 1024 - 512 - 2 * q.a_stage_bytes) / q.b_stage_bytes;
    if (q.SB > 8) q.SB = 8;
    DRAM_REQUIRE(q.SB >= 2 && q.tmem_cols <= 512, "conv3d_umma_fwd: tile-pair pipeline does not fit (SB=%d, tmem=%d)", q.SB, q.tmem_cols);
    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
    int rc2;
    CUtensorMap mA2_hi, mA2_lo;
    q.cb_split = C1p / 64;
    if ((rc2 = make_volume_map_hdw(&mA_hi, x_hi, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc2;
    if ((rc2 = make_weight_map(&mB_hi, w_hi, 27ll * Cout, Cin_pad, q.BN))) return rc2;
    mA_lo = mA_hi; mB_lo = mB_hi;
    if (x_lo && (rc2 = make_volume_map_hdw(&mA_lo, x_lo, N, D, H, W, C1p, q.TW + 2, q.TDD))) return rc2;
    mA2_hi = mA_hi; mA2_lo = mA_lo;
    if (x2_hi && (rc2 = make_volume_map_hdw(&mA2_hi, x2_hi, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc2;
    if (x2_lo && (rc2 = make_volume_map_hdw(&mA2_lo, x2_lo, N, D, H, W, C2p, q.TW + 2, q.TDD))) return rc2;
    if (w_lo && (rc2 = make_weight_map(&mB_lo, w_lo, 27ll * Cout, Cin_pad, q.BN))) return rc2;
    const size_t smem2 = 2 * (size_t)q.a_stage_bytes + (size_t)q.SB * q.b_stage_bytes + 1024 + 512;
    static std::once_flag once2;
    std::call_once(once2, [] {
      cudaFuncSetAttribute(k_conv_umma_fwd2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(k_conv_umma_fwd2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    static long long* prof_buf = nullptr;
    static const bool want_prof = getenv("DRAM_CONV_PROF") != nullptr;
    if (want_prof && !prof_buf) cudaMalloc(&prof_buf, kNumSMs * 8 * sizeof(long long));
    q.prof = want_prof ? prof_buf : nullptr;
    { const char* e = getenv("DRAM_CONV_DBG"); q.dbg = e ? atoi(e) : 0; }
    const int grid2 = q.n_items < kNumSMs ? q.n_items : kNumSMs;
    if (mode == 0) k_conv_umma_fwd2<0><<<grid2, kFwdThreads, smem2, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
    else if (mode == 1) k_conv_umma_fwd2<1><<<grid2, kFwdThreads, smem2, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
    else if (mode == 2) k_conv_umma_fwd2<2><<<grid2, kFwdThreads, smem2, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
    else if (mode == 3) k_conv_umma_fwd2<3><<<grid2, kFwdThreads, smem2, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
    else k_conv_umma_fwd2<4><<<grid2, kFwdThreads, smem2, (cudaStream_t)stream>>>(mA_hi, mA_lo, mB_hi, mB_lo, mA2_hi, mA2_lo, q);
    DRAM_LAUNCH_CHECK();
    if (want_prof) {                         // diagnostics: per-CTA averages of the MMA thread's and one epilogue warp's cycle split
      long long h[kNumSMs * 8];
      cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
      double a[6] = {0, 0, 0, 0, 0, 0};
      const int nb = q.n_items < kNumSMs ? q.n_items : kNumSMs;
      for (int b = 0; b < nb; ++b) for (int j = 0; j < 6; ++j) a[j] += (double)h[b * 8 + j] / nb;
      fprintf(stderr, "[fwd2 prof] items/CTA %.1f BN %d cb %d | mma thread: total %.0f, wait tmem-empty %.0f, wait A %.0f, wait B %.0f | epilogue: wait %.0f, work %.0f (cycles)\n",
              (double)q.n_items / nb, q.BN, q.kblocks_c, a[0], a[1], a[2], a[3], a[4], a[5]);
    }
    return DRAM_OK;
  }
  FwdParams p;
  p.BN = pick_bn(Cout);
  DRAM_REQUIRE(p.BN > 0, "conv3d_umma_fwd: Cout=%d must be a multiple of 16", Cout);
  p.y = y; p.scale = scale; p.shift = shift; p.o_hi = (uint16_t*)out_hi; p.o_lo = (uint16_t*)out_lo;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cout = Cout;
  p.kblocks_c = Cin_pad / 64; p.taps = ksize * ksize * ksize; p.pad = ksize / 2; p.ksteps = ksteps;
  // kw-reuse for the layers whose L2->smem fill rate is the bound (N tile <= 64: 125 B/cycle/SM needed, ~75-80 sustained)
  static const bool allow_w3 = getenv("DRAM_CONV_NO_W3") == nullptr;
  p.w3 = (allow_w3 && ksize == 3 && p.BN <= 64 && W % 16 == 0 && H % 8 == 0) ? 1 : 0;
  if (p.w3) { p.TW = 16; p.TH = 8; p.TD = 1; }
  else pick_fwd_tile(D, H, W, p.TW, p.TH, p.TD);
  p.tiles_w = cdiv(W, p.TW); p.tiles_h = cdiv(H, p.TH); p.tiles_d = cdiv(D, p.TD);
  p.n_mtiles = N * p.tiles_d * p.tiles_h * p.tiles_w;
  p.n_ntiles = Cout / p.BN;
  p.passes = w_lo ? 3 : 1;
  p.a_lo = x_lo ? 1 : 0;
  p.taps_per_stage = p.w3 ? 3 : 1;
  p.a_plane_bytes = p.w3 ? (p.TW + 2) * p.TH * 128 : kATileBytes;
  p.b_tap_bytes = (w_lo ? 2 : 1) * p.BN * 128;
  p.stage_bytes = (x_lo ? 2 : 1) * p.a_plane_bytes + p.taps_per_stage * p.b_tap_bytes;
  p.stages = (kSmemBudget - 1024) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  DRAM_REQUIRE(p.stages >= 2, "conv3d_umma_fwd: pipeline does not fit in shared memory");
  p.acc_cols = (w_lo ? 2 : 1) * p.BN;
  p.tmem_cols = pow2_cols(2 * p.acc_cols);
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  int rc;
  const int abw = p.w3 ? p.TW + 2 : p.TW;
  CUtensorMap tmA2_hi, tmA2_lo;
  p.cb_split = C1p / 64;
  if ((rc = make_volume_map(&tmA_hi, x_hi, N, D, H, W, C1p, abw, p.TH, p.TD, 1, p.w3))) return rc;
  if ((rc = make_weight_map(&tmB_hi, w_hi, (long long)p.taps * Cout, Cin_pad, p.BN))) return rc;
  tmA_lo = tmA_hi; tmB_lo = tmB_hi;
  if (x_lo && (rc = make_volume_map(&tmA_lo, x_lo, N, D, H, W, C1p, abw, p.TH, p.TD, 1, p.w3))) return rc;
  if (w_lo && (rc = make_weight_map(&tmB_lo, w_lo, (long long)p.taps * Cout, Cin_pad, p.BN))) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(k_conv_umma_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  const long long tiles = (long long)p.n_mtiles * p.n_ntiles;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  tmA2_hi = tmA_hi; tmA2_lo = tmA_lo;
  if (x2_hi && (rc = make_volume_map(&tmA2_hi, x2_hi, N, D, H, W, C2p, abw, p.TH, p.TD, 1, p.w3))) return rc;
  if (x2_lo && (rc = make_volume_map(&tmA2_lo, x2_lo, N, D, H, W, C2p, abw, p.TH, p.TD, 1, p.w3))) return rc;
  k_conv_umma_fwd<<<grid, kFwdThreads, smem, (cudaStream_t)stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmA2_hi, tmA2_lo, p);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

static int wgrad_plan(WgParams& p, int N, int D, int H, int W, int Cin_pad, int Cout_pad, int ksize, int passes, int y_lo = 1) {
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.taps = ksize * ksize * ksize; p.pad = ksize / 2;
  p.CB = Cin_pad / 64; p.MB = p.taps * p.CB; p.n_mtiles = cdiv(p.MB, 2);
  p.BN = Cout_pad <= 256 ? Cout_pad : (Cout_pad % 256 == 0 ? 256 : (Cout_pad % 128 == 0 ? 128 : 64));
  p.n_ntiles = Cout_pad / p.BN;
  pick_wgrad_chunk(N, D, H, W, p.TW, p.TH, p.TD, p.TN);
  p.tiles_w = cdiv(W, p.TW); p.tiles_h = cdiv(H, p.TH); p.tiles_d = cdiv(D, p.TD); p.tiles_n = cdiv(N, p.TN);
  p.n_chunks = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  // split-K: pick the number of voxel slabs so that the work items fill whole waves of the 148 persistent CTAs
  // (308 items = 2.08 waves costs a third wave); prefer fewer slabs on ties (less workspace, shorter reduction).
  const int tiles = p.n_mtiles * p.n_ntiles;
  const int max_slabs = cdiv(p.n_chunks, 8) < 48 ? cdiv(p.n_chunks, 8) : 48;    // >= 8 chunks per item
  int best = 1;
  double best_score = -1.0;
  for (int sl = 1; sl <= (max_slabs > 1 ? max_slabs : 1); ++sl) {
    const int cps = cdiv(p.n_chunks, sl), ns = cdiv(p.n_chunks, cps);
    const long long items = (long long)ns * tiles;
    const long long waves = (items + kNumSMs - 1) / kNumSMs;
    const double lastslab = (double)(p.n_chunks - (ns - 1) * cps) / cps;          // a short last slab idles its CTAs
    const double eff = ((double)items - tiles * (1.0 - lastslab)) / (double)(waves * kNumSMs);
    const double score = eff - 0.003 * sl;
    if (score > best_score) { best_score = score; best = sl; }
  }
  p.n_slabs = cdiv(p.n_chunks, cdiv(p.n_chunks, best));
  p.spg = cdiv(kNumSMs, tiles);
  { const char* e = getenv("DRAM_WGRAD_SPG"); if (e && atoi(e) > 0) p.spg = atoi(e); }
  p.passes = passes;
  p.y_lo = (passes == 3 && y_lo) ? 1 : 0;
  p.stage_bytes = (passes == 3 ? 2 : 1) * 2 * kWgBlkBytes + (p.y_lo ? 2 : 1) * (p.BN / 64) * kWgBlkBytes;
  p.stages = (kSmemBudget - 1024) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.concat = (passes == 3 && p.y_lo && p.BN <= 128) ? 1 : 0;
  p.tmem_cols = pow2_cols(2 * (p.concat ? 2 * p.BN : p.BN));
  return DRAM_OK;
}

// SM-pair wgrad (k_conv_umma_wgrad2): full split-bf16 layers whose Cout is a multiple of 128; DRAM_WGRAD_V2=0 switches it off
static bool wgrad2_ok(int Cout_pad, int passes, int y_lo) {
  const char* e = getenv("DRAM_WGRAD_V2");
  return !(e && atoi(e) == 0) && passes == 3 && y_lo && Cout_pad % 128 == 0;
}
static void wgrad2_plan(Wg2Params& p, int N, int D, int H, int W, int Cin_pad, int Cout_pad, int ksize) {
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.taps = ksize * ksize * ksize; p.pad = ksize / 2;
  p.CB = Cin_pad / 64; p.MB = p.taps * p.CB; p.n_mtiles = cdiv(p.MB, 2); p.n_mt4 = cdiv(p.MB, 4);
  p.n_ntiles = Cout_pad / 128;
  pick_wgrad_chunk(N, D, H, W, p.TW, p.TH, p.TD, p.TN);
  p.tiles_w = cdiv(W, p.TW); p.tiles_h = cdiv(H, p.TH); p.tiles_d = cdiv(D, p.TD); p.tiles_n = cdiv(N, p.TN);
  p.n_chunks = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  // split-K over voxel slabs: whole waves of the 74 persistent clusters (as wgrad_plan does for 148 CTAs)
  const int units = kNumSMs / 2, tiles = p.n_mt4 * p.n_ntiles;
  const int max_slabs = cdiv(p.n_chunks, 8) < 48 ? cdiv(p.n_chunks, 8) : 48;    // >= 8 chunks per item
  int best = 1;
  double best_score = -1.0;
  for (int sl = 1; sl <= (max_slabs > 1 ? max_slabs : 1); ++sl) {
    const int cps = cdiv(p.n_chunks, sl), ns = cdiv(p.n_chunks, cps);
    const long long items = (long long)ns * tiles;
    const long long waves = (items + units - 1) / units;
    const double lastslab = (double)(p.n_chunks - (ns - 1) * cps) / cps;
    const double eff = ((double)items - tiles * (1.0 - lastslab)) / (double)(waves * units);
    const double score = eff - 0.003 * sl;
    if (score > best_score) { best_score = score; best = sl; }
  }
  p.n_slabs = cdiv(p.n_chunks, cdiv(p.n_chunks, best));
  p.spg = cdiv(units, tiles);
  { const char* e = getenv("DRAM_WGRAD_SPG"); if (e && atoi(e) > 0) p.spg = atoi(e); }
  p.stages = (kSmemBudget - 1024) / kWg2Stage;
  if (p.stages > 8) p.stages = 8;
}

// kw-reuse plan: slabs chosen so that the static round-robin schedule (item i -> CTA i % 148; a single-source item costs
// 2 tiles, a full one 3) has the shortest makespan; more slabs cost workspace and reduction time.
static bool wgrad_w3_ok(int H, int W, int Cout_pad, int ksize, int passes) {
  static const bool allow = getenv("DRAM_WGRAD_NO_W3") == nullptr;
  return allow && ksize == 3 && Cout_pad == 64 && passes == 3 && H % 8 == 0 && W % 8 == 0;
}
static void wgrad_w3_plan(Wg3Params& p, int N, int D, int H, int W, int Cin_pad) {
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.CB = Cin_pad / 64; p.n_src = 9 * p.CB; p.n_pairs = (p.n_src + 1) / 2;
  p.tiles_w = W / 8; p.tiles_h = H / 8;
  p.n_chunks = p.tiles_w * p.tiles_h * D * N;
  const bool odd = p.n_src & 1;
  int best = 1;
  double best_cost = 1e300;
  const int max_slabs = p.n_chunks / 16 < 128 ? (p.n_chunks / 16 > 1 ? p.n_chunks / 16 : 1) : 128;
  for (int sl = 1; sl <= max_slabs; ++sl) {
    const int cps = cdiv(p.n_chunks, sl), ns = cdiv(p.n_chunks, cps);
    const long long items = (long long)ns * p.n_pairs;
    double worst = 0.0;
    const int ctas = items < kNumSMs ? (int)items : kNumSMs;
    for (int c = 0; c < ctas; ++c) {
      double load = 0.0;
      for (long long it = c; it < items; it += ctas) {
        const int pair = (int)(it % p.n_pairs), slab = (int)(it / p.n_pairs);
        const int chunks = slab == ns - 1 ? p.n_chunks - (ns - 1) * cps : cps;
        load += (double)chunks * ((odd && pair == p.n_pairs - 1) ? 2.0 : 3.0) + 40.0;   // + epilogue / pipeline fill
      }
      if (load > worst) worst = load;
    }
    const double cost = worst * (1.0 + 0.002 * sl);
    if (cost < best_cost) { best_cost = cost; best = sl; }
  }
  p.n_slabs = cdiv(p.n_chunks, cdiv(p.n_chunks, best));
  // slabs per interleave group = the slabs one wave of CTAs holds (wg3_range); DRAM_WGRAD_SPG=1 restores contiguous slabs
  p.spg = cdiv(kNumSMs, p.n_pairs);
  { const char* e = getenv("DRAM_WGRAD_SPG"); if (e && atoi(e) > 0) p.spg = atoi(e); }
  p.stages = (kSmemBudget - 1024) / kW3Stage;
  { const char* e = getenv("DRAM_WGRAD_ORDER"); p.dfast = e ? atoi(e) : 1; }
}

int dram_conv3d_umma_wgrad_kernel(int H, int W, int Cout_pad, int ksize, int has_x_lo, int has_dy_lo) {
  if (H <= 0 || W <= 0 || Cout_pad <= 0 || Cout_pad % 64) return -1;
  const int passes = has_x_lo ? 3 : 1;
  if (wgrad_w3_ok(H, W, Cout_pad, ksize, passes)) return 1;
  if (wgrad2_ok(Cout_pad, passes, has_dy_lo ? 1 : 0)) return 2;
  return 0;
}

size_t dram_conv3d_umma_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin_pad, int Cout_pad, int ksize) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin_pad <= 0 || Cin_pad % 64 || Cout_pad <= 0 || Cout_pad % 64 ||
      (ksize != 1 && ksize != 3))
    return 0;
  size_t w3 = 0;
  if (wgrad_w3_ok(H, W, Cout_pad, ksize, 3)) {
    Wg3Params q;
    wgrad_w3_plan(q, N, D, H, W, Cin_pad);
    w3 = (size_t)q.n_slabs * q.n_pairs * (3 * 128 * 64 * sizeof(float) + sizeof(int)) + 16;   // partial tiles + pacing counters
  }
  WgParams p;
  wgrad_plan(p, N, D, H, W, Cin_pad, Cout_pad, ksize, 3);
  size_t generic = (size_t)p.n_slabs * p.n_mtiles * p.n_ntiles * 128 * p.BN * sizeof(float);
  if (Cout_pad % 128 == 0) {                      // the SM-pair kernel plans its own slab count
    Wg2Params q2;
    wgrad2_plan(q2, N, D, H, W, Cin_pad, Cout_pad, ksize);
    const size_t pairs = (size_t)q2.n_slabs * q2.n_mtiles * q2.n_ntiles * 128 * 128 * sizeof(float);
    if (pairs > generic) generic = pairs;
  }
  return generic > w3 ? generic : w3;
}

int dram_conv3d_umma_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* dw,
                           void* workspace, const void* x2_hi, const void* x2_lo, int Cin1_pad, int N, int D, int H, int W,
                           int Cin, int Cin_pad, int Cout, int Cout_pad, int ksize, void* stream) {
  // virtual concat of the layer input, see dram_conv3d_umma_fwd
  const int C1p = x2_hi ? Cin1_pad : Cin_pad, C2p = Cin_pad - C1p;
  DRAM_REQUIRE(!x2_hi || (Cin1_pad > 0 && Cin1_pad % 64 == 0 && C2p > 0 && (x2_lo != nullptr) == (x_lo != nullptr)),
               "conv3d_umma_wgrad: second operand needs 0 < Cin1_pad (%d, multiple of 64) < Cin_pad (%d) and the same planes as x", Cin1_pad, Cin_pad);
  DRAM_REQUIRE(dy_hi && x_hi && dw && workspace && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_umma_wgrad: bad arguments");
  DRAM_REQUIRE(!(dy_lo != nullptr && x_lo == nullptr), "conv3d_umma_wgrad: dy_lo needs x_lo (modes: both = bf16x3, x_lo only = single-plane gradient x split input, neither = bf16)");
  DRAM_REQUIRE(ksize == 1 || ksize == 3, "conv3d_umma_wgrad: kernel size %d unsupported", ksize);
  DRAM_REQUIRE(Cin > 0 && Cin_pad >= Cin && Cin_pad % 64 == 0 && Cout > 0 && Cout_pad >= Cout && Cout_pad % 64 == 0,
               "conv3d_umma_wgrad: channel pads must be multiples of 64 (Cin %d/%d, Cout %d/%d)", Cin, Cin_pad, Cout, Cout_pad);
  if (wgrad_w3_ok(H, W, Cout_pad, ksize, x_lo ? 3 : 1)) {
    Wg3Params q;
    wgrad_w3_plan(q, N, D, H, W, Cin_pad);
    q.ws = (float*)workspace;
    q.prog = (int*)(q.ws + (size_t)q.n_slabs * q.n_pairs * 3 * 128 * 64);        // uninitialised on purpose (see the kernel)
    { const char* e = getenv("DRAM_CONV_DBG"); q.dbg = e ? atoi(e) : 0; }
    q.y_lo = dy_lo ? 1 : 0;
    CUtensorMap mX_hi, mX_lo, mY_hi, mY_lo;
    int rc3;
    CUtensorMap mX2_hi, mX2_lo;
    q.cb_split = C1p / 64;
    if ((rc3 = make_volume_map(&mX_hi, x_hi, N, D, H, W, C1p, 10, 8, 1, 1, true))) return rc3;
    if ((rc3 = make_volume_map(&mX_lo, x_lo, N, D, H, W, C1p, 10, 8, 1, 1, true))) return rc3;
    mX2_hi = mX_hi; mX2_lo = mX_lo;
    if (x2_hi && (rc3 = make_volume_map(&mX2_hi, x2_hi, N, D, H, W, C2p, 10, 8, 1, 1, true))) return rc3;
    if (x2_lo && (rc3 = make_volume_map(&mX2_lo, x2_lo, N, D, H, W, C2p, 10, 8, 1, 1, true))) return rc3;
    if ((rc3 = make_volume_map(&mY_hi, dy_hi, N, D, H, W, Cout_pad, 8, 8, 1, 1, true))) return rc3;
    if (dy_lo) { if ((rc3 = make_volume_map(&mY_lo, dy_lo, N, D, H, W, Cout_pad, 8, 8, 1, 1, true))) return rc3; }
    else mY_lo = mY_hi;
    const size_t smem3 = (size_t)q.stages * kW3Stage + 1024 + 256;
    static std::once_flag once3;
    std::call_once(once3, [] { cudaFuncSetAttribute(k_conv_umma_wgrad_w3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int items3 = q.n_slabs * q.n_pairs;
    cudaStream_t st3 = (cudaStream_t)stream;
    k_conv_umma_wgrad_w3<<<items3 < kNumSMs ? items3 : kNumSMs, kFwdThreads, smem3, st3>>>(mX_hi, mX_lo, mY_hi, mY_lo, mX2_hi, mX2_lo, q);
    DRAM_LAUNCH_CHECK();
    k_wgrad_w3_reduce<<<grid_for(27ll * Cin * Cout, 256), 256, 0, st3>>>(q.ws, dw, Cout, Cin, q.CB, q.n_pairs, q.n_slabs);
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  if (wgrad2_ok(Cout_pad, x_lo ? 3 : 1, dy_lo ? 1 : 0)) {
    Wg2Params q;
    wgrad2_plan(q, N, D, H, W, Cin_pad, Cout_pad, ksize);
    q.ws = (float*)workspace;
    q.cb_split = C1p / 64;
    CUtensorMap mX_hi, mX_lo, mY_hi, mY_lo, mX2_hi, mX2_lo;
    int rc2;
    if ((rc2 = make_volume_map(&mX_hi, x_hi, N, D, H, W, C1p, q.TW, q.TH, q.TD, q.TN))) return rc2;
    if ((rc2 = make_volume_map(&mX_lo, x_lo, N, D, H, W, C1p, q.TW, q.TH, q.TD, q.TN))) return rc2;
    if ((rc2 = make_volume_map(&mY_hi, dy_hi, N, D, H, W, Cout_pad, q.TW, q.TH, q.TD, q.TN))) return rc2;
    if ((rc2 = make_volume_map(&mY_lo, dy_lo, N, D, H, W, Cout_pad, q.TW, q.TH, q.TD, q.TN))) return rc2;
    mX2_hi = mX_hi; mX2_lo = mX_lo;
    if (x2_hi && (rc2 = make_volume_map(&mX2_hi, x2_hi, N, D, H, W, C2p, q.TW, q.TH, q.TD, q.TN))) return rc2;
    if (x2_lo && (rc2 = make_volume_map(&mX2_lo, x2_lo, N, D, H, W, C2p, q.TW, q.TH, q.TD, q.TN))) return rc2;
    const size_t smem2 = (size_t)q.stages * kWg2Stage + 1024 + 256;
    static std::once_flag once2;
    std::call_once(once2, [] { cudaFuncSetAttribute(k_conv_umma_wgrad2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int items2 = q.n_slabs * q.n_mt4 * q.n_ntiles;
    const int n_cl = items2 < kNumSMs / 2 ? items2 : kNumSMs / 2;
    cudaStream_t st2 = (cudaStream_t)stream;
    k_conv_umma_wgrad2<<<2 * n_cl, kFwdThreads, smem2, st2>>>(mX_hi, mX_lo, mY_hi, mY_lo, mX2_hi, mX2_lo, q);
    DRAM_LAUNCH_CHECK();
    k_wgrad_reduce<<<grid_for((long long)q.taps * Cin * Cout, 256), 256, 0, st2>>>(q.ws, dw, Cout, Cin, q.taps, q.CB, q.n_mtiles,
                                                                                  q.n_ntiles, 128, q.n_slabs);
    DRAM_LAUNCH_CHECK();
    return DRAM_OK;
  }
  WgParams p;
  wgrad_plan(p, N, D, H, W, Cin_pad, Cout_pad, ksize, x_lo ? 3 : 1, dy_lo ? 1 : 0);
  DRAM_REQUIRE(p.stages >= 2, "conv3d_umma_wgrad: pipeline does not fit in shared memory");
  DRAM_REQUIRE(p.tmem_cols <= 512, "conv3d_umma_wgrad: accumulator does not fit in TMEM");
  p.ws = (float*)workspace;
  CUtensorMap tmX_hi, tmX_lo, tmY_hi, tmY_lo;
  int rc;
  CUtensorMap tmX2_hi, tmX2_lo;
  p.cb_split = C1p / 64;
  if ((rc = make_volume_map(&tmX_hi, x_hi, N, D, H, W, C1p, p.TW, p.TH, p.TD, p.TN))) return rc;
  if ((rc = make_volume_map(&tmY_hi, dy_hi, N, D, H, W, Cout_pad, p.TW, p.TH, p.TD, p.TN))) return rc;
  tmX_lo = tmX_hi; tmY_lo = tmY_hi;
  if (x_lo && (rc = make_volume_map(&tmX_lo, x_lo, N, D, H, W, C1p, p.TW, p.TH, p.TD, p.TN))) return rc;
  if (dy_lo && (rc = make_volume_map(&tmY_lo, dy_lo, N, D, H, W, Cout_pad, p.TW, p.TH, p.TD, p.TN))) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(k_conv_umma_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  const int items = p.n_slabs * p.n_mtiles * p.n_ntiles;
  const int grid = items < kNumSMs ? items : kNumSMs;
  cudaStream_t st = (cudaStream_t)stream;
  tmX2_hi = tmX_hi; tmX2_lo = tmX_lo;
  if (x2_hi && (rc = make_volume_map(&tmX2_hi, x2_hi, N, D, H, W, C2p, p.TW, p.TH, p.TD, p.TN))) return rc;
  if (x2_lo && (rc = make_volume_map(&tmX2_lo, x2_lo, N, D, H, W, C2p, p.TW, p.TH, p.TD, p.TN))) return rc;
  k_conv_umma_wgrad<<<grid, kFwdThreads, smem, st>>>(tmX_hi, tmX_lo, tmY_hi, tmY_lo, tmX2_hi, tmX2_lo, p);
  DRAM_LAUNCH_CHECK();
  k_wgrad_reduce<<<grid_for((long long)p.taps * Cin * Cout, 256), 256, 0, st>>>(p.ws, dw, Cout, Cin, p.taps, p.CB, p.n_mtiles,
                                                                                p.n_ntiles, p.BN, p.n_slabs);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
