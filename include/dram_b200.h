/* dram_b200.h — C ABI of the B200-native DRAM hot path (libdram_b200.so).
 *
 * The reference (DIAGNijmegen/bodyct-dram) is pure Python and has NO FFI of its own: every device op on this path is a
 * PyTorch/cuDNN/DGL library call made from dram/parts.py and dram/models.py.  Each entry point below therefore cites the
 * reference call site (file:line under /root/reference/dram/) whose library call it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (the PyTorch caching allocator in practice);
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it and returns;
 *   - no allocation happens inside; scratch space is caller-provided (see the *_workspace_bytes queries);
 *   - return value: 0 = ok, negative = error (DRAM_E_*); dram_last_error() gives a thread-local message;
 *   - volumes are channels-last:  [N][D][H][W][C]  ("NDHWC"), fp32 unless a name says bf16;
 *   - "bf16 split planes": a tensor x is carried as two bf16 tensors (hi, lo) with hi = bf16(x), lo = bf16(x - hi);
 *     tensor-core convolutions accumulate hi*hi + hi*lo + lo*hi in fp32 (3 passes, ~2^-16 relative operand error).
 */
#ifndef DRAM_B200_H
#define DRAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRAM_OK 0
#define DRAM_E_INVALID (-1) /* bad argument / unsupported shape (no silent fallback) */
#define DRAM_E_CUDA (-2)    /* CUDA runtime / driver error */
#define DRAM_E_ARCH (-3)    /* device is not sm_100 */

/* ------------------------------------------------------------------------------------------------ library */
int dram_version(void);              /* 100*major + minor */
int dram_sm_arch(void);              /* 100 — the only architecture this library is built for */
const char* dram_last_error(void);   /* thread-local, never NULL */
int dram_device_check(void);         /* DRAM_OK iff the current device is compute capability 10.x */

/* ------------------------------------------------------------------------------------------------ layout */
/* [N][C][S] <-> [N][S][C] (S = D*H*W).  Boundary converters for nn.Module inputs with C > 1 (models.py:120). */
int dram_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, long long S, void* stream);
int dram_ndhwc_to_ncdhw(const float* src, float* dst, int N, int C, long long S, void* stream);

/* ------------------------------------------------------------------------------------------------ convolution
 * nn.Conv3d(k in {1,3}, padding=k/2, stride 1, bias=False|True): parts.py:105-106,133,185-186; models.py:109-110,490.
 *
 * Weight packs (produced on device from the nn.Parameter [Cout][Cin][k][k][k], every step):
 *   mode 0 (forward): pack[tap][ci][co] = w[co][ci][tap]
 *   mode 1 (dgrad)  : pack[tap][co][ci] = w[co][ci][ntaps-1-tap]   (flipped taps, channels swapped)
 * so that dgrad is the SAME kernel run on dy with the mode-1 pack. */
int dram_pack_weight_f32(const float* w, float* pack, int Cout, int Cin, int ksize, int mode, void* stream);
/* dw[co][ci][tap] = pack[tap][ci][co] (wgrad output -> nn.Parameter.grad layout) */
int dram_unpack_wgrad_f32(const float* pack, float* dw, int Cout, int Cin, int ksize, void* stream);

/* CUDA-core fp32 implicit GEMM: any Cin/Cout.  Used for Cin=1 (K=27, bandwidth-bound), the 1x1x1 heads and shapes
 * the tcgen05 kernel does not cover.  y[n,d,h,w,co] = bias[co] + sum_{tap,ci} x[n,d+kd-p,..,ci] * pack[tap][ci][co] */
int dram_conv3d_simt_fwd(const float* x, const float* pack, const float* bias /*nullable*/, float* y,
                         int N, int D, int H, int W, int Cin, int Cout, int ksize, void* stream);
/* dpack[tap][ci][co] = sum_m x[m+tap][ci] * dy[m][co]; dpack must be zeroed by the caller (atomic split-K). */
int dram_conv3d_simt_wgrad(const float* x, const float* dy, float* dpack,
                           int N, int D, int H, int W, int Cin, int Cout, int ksize, void* stream);

/* tcgen05 / TMEM / TMA implicit GEMM (sm_100a).  Operands are bf16 split planes, channels padded to a multiple of 64.
 *   x_hi/x_lo : [N][D][H][W][Cin_pad] bf16
 *   w_hi/w_lo : [taps][Cout][Cin_pad] bf16 (K-major, see dram_pack_weight_bf16)
 *   precision : x_lo && w_lo  -> split-bf16, x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (forward; parity mode)
 *               !x_lo && w_lo -> single-plane activations x split weights, x*w_hi + x*w_lo (dgrad with the gradient
 *                                carried as one bf16 plane: DRAM_BWD_PRECISION=bf16x2)
 *               !x_lo && !w_lo -> single-pass bf16 ("fast" mode, reported only);  x_lo without w_lo is an error
 *   y         : [N][D][H][W][Cout] fp32 raw accumulators; if scale/shift != NULL the epilogue applies
 *               y = max(0, acc*scale[co] + shift[co])  (eval-mode folded BatchNorm + ReLU, parts.py:107-108)
 *   out_hi/lo : optional (needs scale/shift, Cout % 64 == 0): that activation is written as bf16 split planes
 *               [N][D][H][W][Cout] - the next convolution's operand - instead of y (inference: no fp32 round trip)
 * Cin = real input channels; channels [Cin, Cin_pad) of x and w are zeros (with one 64-channel block the K = 16 steps that
 * would only multiply that padding are not issued).
 * Cout must be a multiple of 16.  K loop = taps x Cin_pad/64 stages of {A 128x64, B BNx64} fed by TMA (5-D tensor map
 * with zero fill for the padding halo), accumulators double-buffered in TMEM, persistent over output tiles.
 * Kernel choice is internal and shape-driven (no caller knob; environment overrides in INTEGRATION.md are diagnostics):
 * full split-bf16 3x3x3 layers with Cout % 64 == 0 run on PAIRS of SMs (thread-block clusters of 2, tcgen05.mma
 * cta_group::2, M = 256, weight columns split across the pair: k_conv_umma_fwd4; likewise k_conv_umma_wgrad2 for
 * dram_conv3d_umma_wgrad with Cout_pad % 128 == 0); everything else runs on the single-SM kernels.  All variants compute
 * the same products; results differ only in the order of the fp32 additions (2e-5 of max |y| in the tests). */
int dram_split_bf16(const float* x, void* hi, void* lo /*nullable*/, long long rows, int C, int Cpad, void* stream);
int dram_pack_weight_bf16(const float* w, void* w_hi, void* w_lo /*nullable*/, int Cout, int Cin, int Cin_pad,
                          int ksize, int mode, void* stream);
int dram_conv3d_umma_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                         const float* scale /*nullable*/, const float* shift /*nullable*/, float* y /*nullable with out_hi*/,
                         void* out_hi /*nullable*/, void* out_lo /*nullable*/, float* bn_partials /*nullable*/,
                         const void* x2_hi /*nullable*/, const void* x2_lo /*nullable*/, int Cin1_pad,
                         int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, void* stream);
/* x2_hi != NULL: VIRTUAL channel concat of the input (UpsampleConvBlock5d: cat([upsampled, skip]), parts.py:151-155, without
 * materialising the 768/384/192-channel tensor): x holds channels [0, Cin1_pad) with row pitch Cin1_pad (multiple of 64),
 * x2 holds channels [Cin1_pad, Cin_pad) with row pitch Cin_pad - Cin1_pad; the weights are packed for the concatenated
 * input as usual.  The K loop of every kernel switches tensor maps at the 64-channel block boundary. */
/* Train-mode BatchNorm statistics (parts.py:19) as a by-product of the convolution's epilogue: with
 * bn_partials != NULL (raw output only) the epilogue also writes per-channel partial sums of y and y*y,
 * [rows][2][Cout] floats, rows = dram_conv3d_umma_fwd_stat_rows(...) (one row per output tile and epilogue warp; 0 = the
 * kernel that runs this shape has no such epilogue: use dram_bn_stats).  dram_bn_stats_from_partials adds the rows up in
 * double in a fixed order (two levels; workspace of dram_bn_stats_from_partials_workspace_bytes(C) bytes) ->
 * sums[0..C) = sum y, sums[C..2C) = sum y*y, what dram_bn_finalize takes. */
long long dram_conv3d_umma_fwd_stat_rows(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, int has_x_lo, int has_w_lo);
/* Which kernel dram_conv3d_umma_fwd / _wgrad run for a shape (host logic only, no device work; for tests and diagnostics):
 * fwd: 0 = generic single-SM tiles (k_conv_umma_fwd), 2 = weight-sharing tile pairs (k_conv_umma_fwd2), 3 = channels on M
 * (k_conv_umma_fwd3), 4 = SM pairs with kw re-use (k_conv_umma_fwd4), 5 = SM pairs on generic tiles (k_conv_umma_fwd4);
 * wgrad: 0 = generic (k_conv_umma_wgrad), 1 = kw re-use for Cout <= 64 (k_conv_umma_wgrad_w3), 2 = SM pairs
 * (k_conv_umma_wgrad2).  -1 = invalid arguments. */
int dram_conv3d_umma_fwd_kernel(int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int ksize, int has_x_lo, int has_w_lo);
int dram_conv3d_umma_wgrad_kernel(int H, int W, int Cout_pad, int ksize, int has_x_lo, int has_dy_lo);
size_t dram_bn_stats_from_partials_workspace_bytes(int C);
int dram_bn_stats_from_partials(const float* partials, long long rows, int C, double* sums, void* workspace, void* stream);
/* wgrad on tensor cores: dw[co][ci][tap] (+)= sum_m dy[m][co] * x[m+tap][ci]; operands as split planes
 *   dy_hi/dy_lo : [N][D][H][W][Cout_pad]   x_hi/x_lo : [N][D][H][W][Cin_pad]   (pads are multiples of 64)
 *   precision: dy_lo && x_lo -> three products; !dy_lo && x_lo -> x_hi*dy + x_lo*dy (single-plane gradient,
 *   DRAM_BWD_PRECISION=bf16x2); neither -> single pass;  dy_lo without x_lo is an error
 * partial sums over voxel ranges are written to `workspace` (dram_conv3d_umma_wgrad_workspace_bytes) and reduced
 * deterministically into dw in nn.Parameter layout [Cout][Cin][taps]. */
size_t dram_conv3d_umma_wgrad_workspace_bytes(int N, int D, int H, int W, int Cin_pad, int Cout_pad, int ksize);
int dram_conv3d_umma_wgrad(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, float* dw,
                           void* workspace, const void* x2_hi /*nullable*/, const void* x2_lo /*nullable*/, int Cin1_pad,
                           int N, int D, int H, int W, int Cin, int Cin_pad, int Cout, int Cout_pad, int ksize, void* stream);
/* (x2_*, Cin1_pad: virtual concat of the layer input, as in dram_conv3d_umma_fwd) */

/* ------------------------------------------------------------------------------------------------ BatchNorm + ReLU (+ pool)
 * nn.BatchNorm3d (eps 1e-5, momentum 0.1) + nn.ReLU: parts.py:19,50,107-108.  rows = N*D*H*W, y is [rows][C]. */
int dram_bn_stats(const float* y, double* sums /*[2*C]: sum, sumsq; overwritten*/, long long rows, int C, void* stream);
/* sums -> mean/rstd/scale/shift; running stats updated `n_updates` times (checkpointed blocks update twice per step,
 * models.py:123-143).  `count` = rows (global rows under data parallelism, after the caller all-reduced `sums`). */
int dram_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int n_updates, float* mean, float* rstd,
                     float* scale, float* shift, int C, void* stream);
/* eval mode: scale/shift from running statistics */
int dram_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, float* scale, float* shift, int C, void* stream);
/* a = max(0, y*scale + shift); if pooled != NULL also MaxPool3d(2,2,0) of a (parts.py:191,195), floor semantics */
int dram_bn_relu_apply(const float* y, const float* scale, const float* shift, float* a, float* pooled /*nullable*/,
                       int N, int D, int H, int W, int C, void* stream);
/* backward, pass 1: dz = da * (y*scale+shift > 0); sums[0:C] = sum dz, sums[C:2C] = sum dz * xhat.
 * `da` rows are `da_pitch` floats apart (0 = C): the skip half of a decoder gradient is read in place.
 * wtop != NULL: fused RAM head (models.py:145) — `da` is g [rows] and da[r][c] = g[r]*wtop[c]; sums is then
 * double[3*C+1] with sums[2C+c] = sum g*relu(y*scale+shift)[c] (= d top_layer.weight), sums[3C] = sum g (= d bias). */
int dram_bn_relu_bwd_reduce(const float* da, long long da_pitch, const float* wtop /*nullable*/, const float* y,
                            const float* scale, const float* shift, const float* mean, const float* rstd, double* sums,
                            long long rows, int C, void* stream);
/* backward, pass 2 (training): dy = gamma*rstd*(dz - sum_dz/count - xhat*sum_dz_xhat/count);
 * eval (sums == NULL): dy = dz*scale */
int dram_bn_relu_bwd_apply(const float* da, long long da_pitch, const float* y, const float* scale, const float* shift,
                           const float* mean, const float* rstd, const float* gamma, const double* sums /*nullable*/,
                           double count, float* dy, long long rows, int C, void* stream);
/* MaxPool3d(2,2,0) backward: da[first argmax of each window] += dpooled (da is read-modify-write; windows are disjoint) */
int dram_maxpool2_bwd(const float* a, const float* dpooled, float* da, int N, int D, int H, int W, int C, void* stream);

/* ------------------------------------------------------------------------------------------------ the same units on bf16 split planes
 * Between two tensor-core convolutions an activation exists ONLY as split planes [rows][Cpad] (hi, lo; lo == NULL in the
 * single-pass bf16 mode; channels [C,Cpad) zero): the kernels below write the next convolution's TMA operand directly, so
 * no fp32 activation and no separate split pass exist.  C must be a multiple of 8. */
/* parts.py:107-108 (+191,195): a = max(0, y*scale+shift) -> planes; p_hi != NULL: also MaxPool3d(2,2,0)(a) -> planes */
int dram_bn_relu_apply_planes(const float* y, const float* scale, const float* shift, void* a_hi, void* a_lo /*nullable*/,
                              void* p_hi /*nullable*/, void* p_lo /*nullable*/, int N, int D, int H, int W, int C, int Cpad,
                              void* stream);
/* dram_bn_relu_bwd_apply writing dy as planes (the operand of dgrad and wgrad); wtop != NULL: da[r][c] = g[r]*wtop[c] */
int dram_bn_relu_bwd_apply_planes(const float* da, long long da_pitch, const float* wtop /*nullable*/, const float* y,
                                  const float* scale, const float* shift, const float* mean, const float* rstd,
                                  const float* gamma, const double* sums /*nullable*/, double count, void* dy_hi,
                                  void* dy_lo /*nullable*/, long long rows, int C, int Cpad, void* stream);
/* Backward of a pooled unit (parts.py:193-196: returns (a, maxpool(a))): da = ga (nullable, rows ga_pitch apart) + gp
 * routed to the first maximum of each 2x2x2 window, recomputed from y (ATen max_pool3d backward semantics); fused with
 * the BatchNorm+ReLU backward so that neither `a` nor a dense `da` is ever stored.  C must be a multiple of 4. */
int dram_bn_pool_bwd_reduce(const float* ga /*nullable*/, long long ga_pitch, const float* gp /*nullable*/, const float* y,
                            const float* scale, const float* shift, const float* mean, const float* rstd, double* sums,
                            int N, int D, int H, int W, int C, void* stream);
int dram_bn_pool_bwd_apply_planes(const float* ga /*nullable*/, long long ga_pitch, const float* gp /*nullable*/,
                                  const float* y, const float* scale, const float* shift, const float* mean,
                                  const float* rstd, const float* gamma, const double* sums /*nullable*/, double count,
                                  void* dy_hi, void* dy_lo /*nullable*/, int N, int D, int H, int W, int C, int Cpad,
                                  void* stream);
/* parts.py:149-153 on planes: cat = [trilinear x2 (align_corners=True) of x | centre-cropped skip]; P1/P2/Pc = channel
 * pitches of the x / skip / cat planes */
int dram_upsample2x_concat_planes(const void* x_hi, const void* x_lo, const void* skip_hi, const void* skip_lo, void* cat_hi,
                                  void* cat_lo, int N, int d, int h, int w, int C1, int P1, int Ds, int Hs, int Ws, int C2,
                                  int P2, int Pc, void* stream);
/* the upsampled half alone, planes -> planes [N][2d][2h][2w][Pout] (channels [C1, Pout) zeroed): the first operand of the
 * virtual concat; the skip planes are the second operand as they are (no crop: skip size == 2x the input size) */
int dram_upsample2x_planes(const void* x_hi, const void* x_lo, void* out_hi, void* out_lo, int N, int d, int h, int w, int C1,
                           int P1, int Pout, void* stream);
/* planes -> fp32 [rows][C]: materialises an activation for a consumer outside the tensor-core path */
int dram_merge_planes(const void* hi, const void* lo /*nullable*/, float* out, long long rows, int C, int Cpad, void* stream);

/* Conv3d(C, 8, kernel_size=1) on split planes - the attention reshape heads (models.py:488-494, applied at :564 to a
 * detached decoder feature map).  x planes [rows][Cin_pad] bf16 (x_lo nullable), w [8][Cin] fp32 (nn.Conv3d layout),
 * bias [8] (nullable), y [rows][8] fp32.  Supported: Cin_pad 64 | 128 (query dram_pointwise8_planes_supported). */
int dram_pointwise8_planes_supported(int Cin_pad, int Cout);
int dram_pointwise8_planes_fwd(const void* x_hi, const void* x_lo, const float* w, const float* bias, float* y, long long rows,
                               int Cin, int Cin_pad, void* stream);
/* dw [8][Cin], dbias [8] (nullable): overwritten; deterministic (per-block partials in `workspace`, fixed-order sum) */
size_t dram_pointwise8_planes_wgrad_workspace_bytes(int Cin_pad);
int dram_pointwise8_planes_wgrad(const void* x_hi, const void* x_lo, const float* dy, float* dw, float* dbias, void* workspace,
                                 long long rows, int Cin, int Cin_pad, void* stream);

/* ------------------------------------------------------------------------------------------------ decoder glue
 * nn.Upsample(scale_factor=2, trilinear, align_corners=True) + crop_concat_5d: parts.py:149-153,37-46.
 * cat[..., 0:C1] = up(x) ; cat[..., C1:C1+C2] = skip centre-cropped with ceil offsets.  x: [N][d][h][w][C1],
 * skip: [N][Ds][Hs][Ws][C2], cat: [N][2d][2h][2w][C1+C2]. */
int dram_upsample2x_concat_fwd(const float* x, const float* skip, float* cat, int N, int d, int h, int w, int C1,
                               int Ds, int Hs, int Ws, int C2, void* stream);
/* dx (gather form of the transposed interpolation, no atomics) and dskip (zero outside the crop; NULL = not wanted,
 * the caller reads dcat[..., C1:] in place) */
int dram_upsample2x_concat_bwd(const float* dcat, float* dx, float* dskip /*nullable*/, int N, int d, int h, int w, int C1,
                               int Ds, int Hs, int Ws, int C2, void* stream);
/* F.interpolate(size=..., trilinear, align_corners=True): models.py:146,514-518,588,591-592; job_runner.py:766,993.
 * src [N][d][h][w][C] -> dst [N][D][H][W][C]; the backward is the exact adjoint, gather form. */
int dram_trilinear_resize_fwd(const float* src, float* dst, int N, int d, int h, int w, int D, int H, int W, int C,
                              void* stream);
int dram_trilinear_resize_bwd(const float* ddst, float* dsrc, int N, int d, int h, int w, int D, int H, int W, int C,
                              void* stream);

/* ------------------------------------------------------------------------------------------------ RAM head
 * top_layer = nn.Conv3d(64 -> out_ch, k=1) (models.py:109-110,145): the regression-weight channel reduce.
 * Fused with the last BatchNorm+ReLU when scale/shift != NULL:  ram[m][o] = b[o] + sum_c relu(y*scale+shift)[c]*w[o][c] */
int dram_ram_reduce_fwd(const float* feat, const float* scale /*nullable*/, const float* shift /*nullable*/,
                        const float* w /*[O][C]*/, const float* b /*[O]*/, float* ram /*[rows][O]*/, long long rows,
                        int C, int O, void* stream);
/* dfeat[m][c] = sum_o dram[m][o]*w[o][c]; dwb[o*C+c] = sum_m dram[m][o]*feat[m][c]; dwb[O*C+o] = sum_m dram[m][o]
 * (dwb: double[O*C+O], overwritten) */
int dram_ram_reduce_bwd(const float* dram, const float* feat, const float* w, float* dfeat, double* dwb,
                        long long rows, int C, int O, void* stream);
/* pooling_dense_features / masked mean (models.py:37-49; metrics.py:160-165):
 * out[b*2+0] = sum_v f(x[b][v]) * mask[b][v], out[b*2+1] = sum_v mask[b][v]; f = sigmoid if use_sigmoid.
 * mode_gt0: use (mask > 0) instead of the mask value (metrics.py:162).  out: double[2*B], overwritten. */
int dram_masked_pool_fwd(const float* x, const float* mask, double* out, int B, long long V, int use_sigmoid,
                         int mode_gt0, void* stream);
/* dx[b][v] = g[b] * f'(x) * m[b][v]   (g already divided by the mask count by the caller) */
int dram_masked_pool_bwd(const float* x, const float* mask, const float* g, float* dx, int B, long long V,
                         int use_sigmoid, int mode_gt0, void* stream);
/* Interval-regression term of IntRegLoss from the two masked-pool results of a batch (metrics.py:121-137 get_labels +
 * :158-177 compute_reg_loss_with_probs), one tiny kernel instead of ~40 [B]-sized tensor ops:
 *   pred[b] = pool_pred[b][0] / pool_pred[b][1] (mean probability over the lobe), rub[b] = pool_rub[b][0] / pool_rub[b][1]
 *   (lesion-candidate ratio), both rounded to fp32 as the reference's tensors are; the target interval [lo, hi] =
 *   band[b] = ctss_ratio_map[ctss] intersected with [rub - band_width, rub + band_width] clamped to [0, 1] (with the
 *   reference's two fall-backs for an empty intersection), evaluated in double and rounded to fp32;
 *   loss = sum_b max((pred - (hi + lo) / 2)^2 - (0.5 (hi - lo))^2, 0) / w[b]                 -> loss[0] (fp32)
 *   g[b] = d loss / d pool_pred[b][0] = [hinge active] 2 (pred - mid) / w[b] / pool_pred[b][1]   -> g (fp32), which is what
 *   dram_masked_pool_bwd takes.  band: double [B][2], w: fp32 [B] (already clamped), pools: double [B][2]. */
int dram_int_reg_loss(const double* pool_pred, const double* pool_rub, const double* band, const float* w, double band_width,
                      float* loss, float* g, int B, void* stream);
/* Segmentation term of IntRegRefineLoss in two passes (metrics.py:10-51 BootBinCrossEntropy over the thresholded pseudo
 * labels of metrics.py:325-354): dense / refined are the RAM and refined-RAM logits [B][V], lobes / lesions the float
 * masks, keep[b] = 0 zeroes the pseudo labels of sample b (ctss == 0, metrics.py:327-328).
 * fwd: sums[0..6] (double, overwritten) = #outside, #inside, sum t*inside, sum nll*outside, sum nll*t*inside,
 *      sum nll*(1-t)*inside, sum -log(pt_hat)*inside;  the caller forms the loss from these scalars (and all-reduces
 *      sums[0..2] under data parallelism).
 * bwd: drefined[b][v] = d loss / d refined given coef = (c_out, c_t, c_nt, c_boot) on the device with
 *      loss = c_out*sums[3] + c_t*sums[4] + c_nt*sums[5] + c_boot*sums[6]. */
int dram_boot_bce_fwd(const float* dense, const float* refined, const float* lobes, const float* lesions, const float* keep,
                      double* sums, int B, long long V, float eps, void* stream);
int dram_boot_bce_bwd(const float* dense, const float* refined, const float* lobes, const float* lesions, const float* keep,
                      const float* coef, float* drefined, int B, long long V, float eps, void* stream);
/* out = act(in) elementwise (0 identity, 1 sigmoid, 2 relu): job_runner.py:765 `probs = F.sigmoid(dense_outs)` */
int dram_ram_activation(const float* in, float* out, long long n, int act, void* stream);
/* Inference epilogue (job_runner.py:765-770 / 993-1004): trilinear-upsample one chunk's RAM [d][h][w] to the lobe crop
 * [cd][ch][cw] (align_corners=True), apply `act` (0 = identity, 1 = sigmoid, 2 = relu), multiply by `gain`
 * (1/max for the max-normalised head) and write it into the scan-sized heat map at offset (oz,oy,ox) ONLY where
 * crop_mask != 0 (bit-exact lobe masking).  If maxval != NULL, also atomically tracks max(act(value)) over the crop. */
int dram_ram_upsample_mask_scatter(const float* ram, const uint8_t* crop_mask, float* heat, float* maxval /*nullable*/,
                                   int d, int h, int w, int cd, int ch, int cw, int SD, int SH, int SW, int oz, int oy,
                                   int ox, int act, float gain, void* stream);

/* ------------------------------------------------------------------------------------------------ PCM stencil attention
 * PCM.forward via DGL update_all (models.py:322-411) on the voxel-grid graph of models.py:223-259, restated as an
 * 18/6/26(+1)-point stencil attention on a [D][H][W] grid:
 *   q = theta(f_x), k_o = phi(f_{x+o}); s_o = act(<q,k_o>) / T(x); a = softmax_o over in-grid neighbours;
 *   out_x = sum_o a_o * cam_{x+o}                      (the affine G/r maps collapse to a scalar affine, applied by the host)
 * f: [B][D][H][W][Cf] fp32, cam/out: [B][D][H][W].  theta/phi: [F][Cf] weights + [F] bias.
 * flags bit0: relu on logits; bits1-2: temperature (0 none, 1 sqrt(degree) — models.py:274-277, 2 = 0.01);
 * connectivity 1|2|3, self_loop 0|1. */
int dram_pcm_num_offsets(int connectivity, int self_loop); /* O: 18 for (2, no self loop) */
/* R = B*V rows.  qk (out, dram_pcm_qk_floats(R, F) floats): theta|phi projections in blocks of 32 consecutive voxels,
 * [ceil(R/32)][2F][32] (coalesced neighbour loads, compile-time feature offsets).  F: 4 | 8 | 16.
 * stats [R][4] (out, may be NULL at inference): (max, 1/normaliser, 1/T, output) of each node's softmax - what the backward
 * keeps instead of the [R][O] attention weights (they are recomputed from qk).
 * stats == NULL (no-grad forward): projection and attention run as ONE kernel with the projections in shared memory (a block
 * marches a 4 x 64 column of the grid along z, three z-planes of phi(f) and cam in a ring); qk may then be NULL. */
size_t dram_pcm_qk_floats(long long rows, int F);
int dram_pcm_fwd(const float* f, const float* cam, const float* theta_w, const float* theta_b, const float* phi_w,
                 const float* phi_b, float* qk, float* stats, float* out, int B, int D, int H, int W, int Cf, int F,
                 int connectivity, int self_loop, int flags, void* stream);
/* dqk_ws (dram_pcm_bwd_ws_floats floats): scratch.  dcam [R], df [R][Cf]: overwritten.
 * dparams double[2*F*(Cf+1)] = dtheta_w, dtheta_b, dphi_w, dphi_b: overwritten (per-block partials summed in a fixed
 * order: deterministic). */
size_t dram_pcm_bwd_ws_floats(long long rows, int Cf, int F);
int dram_pcm_bwd(const float* f, const float* cam, const float* theta_w, const float* phi_w, const float* qk,
                 const float* stats, const float* dout, float* dqk_ws, float* dcam, float* df, double* dparams, int B,
                 int D, int H, int W, int Cf, int F, int connectivity, int self_loop, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------ scan pre/post-processing
 * What LesionSegTest.run / evaluate_scan do on the host around the model (job_runner.py:730-772, 951-1015), on the GPU.
 * scan: int16 HU [SD][SH][SW]; labels: uint8 lobe labels (0 = background, 1..5 = lobes); heat: fp32 [SD][SH][SW]. */
/* utils.find_crops (utils.py:244-254) for all labels at once: out[l*6 + {0,1,2}] = min z,y,x, out[l*6 + {3,4,5}] = max
 * z,y,x (inclusive) for l in 1..nlabels (<= 7); empty labels keep min = INT_MAX, max = -1.  out: int[(nlabels+1)*6]. */
int dram_label_bboxes(const uint8_t* labels, int D, int H, int W, int nlabels, int* out, void* stream);
/* job_runner.py:772 lesion ratio: out2[0] = sum of values over voxels with labels > 0, out2[1] = their count (doubles);
 * deterministic (per-block partials in `workspace`, dram_labelled_sum_workspace_bytes(), summed in block order) */
size_t dram_labelled_sum_workspace_bytes(void);
int dram_labelled_sum(const float* values, const uint8_t* labels, long long n, double* out2, void* workspace, void* stream);
/* Read back a SMALL device result (bounding boxes, a histogram) with a kernel that stores into pinned, device-mapped host
 * memory instead of a cudaMemcpy: the copy engines stay free for the bulk transfers of neighbouring scans (run_scans).
 * The caller synchronises the stream (or an event) before reading dst.  nbytes: multiple of 4, <= 1 MiB. */
int dram_store_to_host(const void* src, void* dst_pinned_host, int nbytes, void* stream);
/* job_runner.py:961-984: crop [cz,cz+cd)x[cy,cy+ch)x[cx,cx+cw), voxels outside `label` -> pad_value (-2048), Windowing
 * (data_transforms.py:37-54, float32) to [0,1], Resample('fixed_size') (data_transforms.py:170-175: SimpleITK resample with
 * require_spacing = spacing * crop_size / chunk_size, linear for the image, nearest for the mask) to the chunk grid (d,h,w).
 * sp_* = spacing of the scan grid (z, y, x).  Coordinates and interpolation in double exactly as ITK 4.13 computes them
 * (see dram_itk_resample): bit-identical to the float64 oracle. */
int dram_lobe_chunk_preprocess(const short* scan, const uint8_t* labels, int SD, int SH, int SW, int label, int cz, int cy,
                               int cx, int cd, int ch, int cw, float win_lo, float win_hi, float pad_value, double sp_z,
                               double sp_y, double sp_x, float* img, float* msk, int d, int h, int w, void* stream);
/* utils.resample (utils.py:414-434) -> SimpleITK ResampleImageFilter (identity transform, identity direction, shared
 * origin, default value 0, output pixel type = input pixel type) of a whole volume [d][h][w] with spacing in_sp_* onto the
 * grid [D][H][W] with spacing out_sp_* (z, y, x).  Restates ITK 4.13 in double, operation by operation: continuous index =
 * (1/in_sp) * (out_sp * index) (ImageBase::TransformIndexToPhysicalPoint / TransformPhysicalPointToContinuousIndex), along x
 * through ResampleImageFilter::LinearThreadedGenerateData's start + (i/size) * (end - start); outside [-0.5, n-0.5) -> 0;
 * LinearInterpolateImageFunction::EvaluateOptimized nested lerps in double; NearestNeighbor = floor(c + 0.5);
 * CastPixelWithBoundsChecking (clamp + truncating cast for integers, round-to-nearest for f32).
 * dtype 0 = f32, 1 = i16, 2 = u8; mode 0 = linear, 1 = nearest. */
int dram_itk_resample(const void* src, void* dst, int dtype, int d, int h, int w, int D, int H, int W, double in_sp_z,
                      double in_sp_y, double in_sp_x, double out_sp_z, double out_sp_y, double out_sp_x, int mode, void* stream);
/* dram_ram_upsample_mask_scatter with the mask read from the scan-sized label volume (labels == label) */
int dram_ram_upsample_label_scatter(const float* ram, const uint8_t* labels, int label, float* heat, int d, int h, int w,
                                    int cd, int ch, int cw, int SD, int SH, int SW, int oz, int oy, int ox, int act,
                                    float gain, void* stream);
/* utils.binary_cam (utils.py:226-242): 256-bin histogram of uint8(windowing(v, (lo,hi)) -> [0,255]) over labels > 0,
 * integer-exact (dtype 0 = f32 values, 1 = i16 values); hist: unsigned[256], overwritten.  Otsu runs on the host. */
int dram_masked_hist_u8(const void* values, int dtype, const uint8_t* labels, long long n, float lo, float hi,
                        unsigned int* hist, void* stream);
/* job_runner.py:1009-1015: lesion = heat > th; post = lesion && windowing(scan,(win_lo,win_hi)->(0,1)) > th2 && !vessel
 * (post / scan / vessel may be NULL) */
int dram_threshold_masks(const float* heat, const short* scan, const uint8_t* vessel, long long n, double th, double th2,
                         float win_lo, float win_hi, uint8_t* lesion, uint8_t* post, void* stream);

/* ---- cross-GPU exchanges of the data-parallel step over NVLink peer memory (new functionality: the reference has no
 * distributed code).  Parity target = the single-process reference on the GLOBAL batch, so every train-mode
 * nn.BatchNorm3d (parts.py:19) needs all ranks' batch statistics, forward and backward, and IntRegRefineLoss needs the
 * batch-global normalisers of metrics.py:30,37,42,48.  Each rank owns a mailbox (dram_peer_alloc) that every other rank maps
 * through CUDA IPC (dram_peer_export -> 64-byte handle -> dram_peer_open); one call = ONE kernel: push the payload into every
 * mailbox, release a flag, wait for every peer's flag, sum in rank order (bit-identical on all ranks).  `mailboxes` is a HOST
 * array of `world` device pointers (entry `rank` = the own mailbox); all ranks must issue the same sequence of calls.
 * nvirt = 1 in production.  nvirt = world plays ALL ranks as the blocks of one launch on one GPU (in/out/... are host arrays
 * of nvirt per-rank pointers) - the single-GPU test of the protocol. */
size_t dram_peer_mailbox_bytes(void);
int dram_peer_max_doubles(void);
int dram_peer_max_ranks(void);
int dram_peer_alloc(void** mailbox);
int dram_peer_free(void* mailbox);
int dram_peer_export(const void* mailbox, void* handle64);
int dram_peer_open(const void* handle64, void** mailbox);
int dram_peer_close(void* mailbox);
/* out[r][i] = sum over ranks q (in order) of in_q[i], n <= dram_peer_max_doubles(); out may alias in */
int dram_peer_allreduce_f64(void* const* mailboxes, const double* const* in, double* const* out, int n, int rank, int world,
                            int nvirt, void* stream);
/* dram_bn_finalize on the GLOBAL statistics, fused with their exchange: payload = local (sum y [C], sum y^2 [C]) + the local
 * element count counts[r]; sums_out[r] receives the global [2C+1]; then mean / rstd / scale / shift and the running-statistic
 * update exactly as dram_bn_finalize (unbiased variance with the GLOBAL count). */
int dram_bn_finalize_peer(void* const* mailboxes, const double* const* sums_in, const double* counts, double* const* sums_out,
                          int rank, int world, int nvirt, const float* const* gamma, const float* const* beta,
                          float* const* running_mean, float* const* running_var, float momentum, float eps, int n_updates,
                          float* const* mean, float* const* rstd, float* const* scale, float* const* shift, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRAM_B200_H */
