"""Tensor-level wrappers over the C ABI (include/dram_b200.h).

Volumes are torch tensors of LOGICAL shape [N, C, D, H, W] (the reference's layout at every nn.Module boundary) whose
MEMORY is channels-last ([N][D][H][W][C], `torch.channels_last_3d`), the layout every kernel works in.  PyTorch is used
for device memory and streams only; every op below is a call into libdram_b200.so and raises if that is impossible.
"""
import os

import torch

from . import lib as _lib

CL = torch.channels_last_3d


def _L():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DramLibraryError(f"{name}: the DRAM B200 path needs CUDA tensors (got {type(t).__name__} on "
                                    f"{getattr(t, 'device', '?')}); there is no CPU fallback")
    if t.dtype != dtype:
        raise _lib.DramLibraryError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def new_volume(N, C, D, H, W, device, dtype=torch.float32, zero=False):
    buf = (torch.zeros if zero else torch.empty)((N, D, H, W, C), device=device, dtype=dtype)
    return buf.permute(0, 4, 1, 2, 3)


def is_cl(x):
    return x.dim() == 5 and x.permute(0, 2, 3, 4, 1).is_contiguous()


def to_cl(x, name="input"):
    """Return x (logical NCDHW) with channels-last memory; converts with our own kernel when needed."""
    _req(x, name)
    if x.dim() != 5:
        raise _lib.DramLibraryError(f"{name}: expected a 5-d NCDHW tensor, got {tuple(x.shape)}")
    if is_cl(x):
        return x
    x = x.contiguous()
    N, C, D, H, W = x.shape
    out = new_volume(N, C, D, H, W, x.device)
    _lib.check(_L().dram_ncdhw_to_ndhwc(x.data_ptr(), out.data_ptr(), N, C, D * H * W, _stream()), "ncdhw_to_ndhwc")
    return out


def to_ncdhw(x):
    """Channels-last volume -> standard contiguous NCDHW (only needed when a caller insists on that memory layout)."""
    _req(x, "to_ncdhw")
    if x.is_contiguous():
        return x
    N, C, D, H, W = x.shape
    src = to_cl(x)
    out = torch.empty((N, C, D, H, W), device=x.device, dtype=x.dtype)
    _lib.check(_L().dram_ndhwc_to_ncdhw(src.data_ptr(), out.data_ptr(), N, C, D * H * W, _stream()), "ndhwc_to_ncdhw")
    return out


# ------------------------------------------------------------------------------------------------ configuration
def conv_path():
    """'umma' (tcgen05 tensor cores where the shape allows, default) or 'simt' (CUDA-core fp32 everywhere)."""
    return os.environ.get("DRAM_CONV_PATH", "umma")


def precision():
    """'bf16x3' (split-bf16, parity mode, default) or 'bf16' (single pass, fast mode)."""
    return os.environ.get("DRAM_PRECISION", "bf16x3")


def grad_planes_three():
    """Does the GRADIENT operand of dgrad / wgrad (dy) carry a lo plane?  Default yes (three products, the parity mode every
    gradient tolerance in tests/ is stated for).  DRAM_BWD_PRECISION=bf16x2 carries dy as ONE bf16 plane, multiplied with the
    split weights (dgrad: dy*w_hi + dy*w_lo) and the split layer input (wgrad: x_hi*dy + x_lo*dy): exact products with the
    bf16-rounded gradient (tests/test_kernels_gpu.py), ~6 ms of the 70 ms training step, per-parameter gradient error against
    the fp32 CPU oracle 1e-3 -> 4e-3 at 32^3 and 5.0e-3 -> 5.6e-3 at 80^3 (tests/diag_bwd_x2.py).
    Forward quantities are untouched either way."""
    return precision() == "bf16x3" and os.environ.get("DRAM_BWD_PRECISION", "bf16x3") != "bf16x2"


def _pad64(c):
    return (c + 63) // 64 * 64


def umma_ok_fwd(Cin, Cout, ksize):
    return conv_path() == "umma" and ksize in (1, 3) and Cin >= 16 and Cout % 16 == 0


# ------------------------------------------------------------------------------------------------ convolution
def pack_weight_f32(w, mode):
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    pack = torch.empty(k ** 3 * Cin * Cout, device=w.device, dtype=torch.float32)
    _lib.check(_L().dram_pack_weight_f32(w.data_ptr(), pack.data_ptr(), Cout, Cin, k, mode, _stream()), "pack_weight_f32")
    return pack


def conv_simt(x, pack, bias, Cout, ksize):
    """x CL volume [N,Cin,D,H,W]; pack [taps][Cin][Cout] -> CL volume [N,Cout,D,H,W]"""
    N, Cin, D, H, W = x.shape
    y = new_volume(N, Cout, D, H, W, x.device)
    _lib.PROFILE.note(flops=2.0 * N * D * H * W * Cin * Cout * ksize ** 3, bytes=4.0 * N * D * H * W * (Cin + Cout))
    _lib.check(_L().dram_conv3d_simt_fwd(x.data_ptr(), pack.data_ptr(), _p(bias), y.data_ptr(), N, D, H, W, Cin, Cout,
                                         ksize, _stream()), "conv3d_simt_fwd")
    return y


def conv_simt_wgrad(x, dy, ksize):
    """-> dw in nn.Parameter layout [Cout, Cin, k, k, k]"""
    N, Cin, D, H, W = x.shape
    Cout = dy.shape[1]
    taps = ksize ** 3
    dpack = torch.zeros(taps * Cin * Cout, device=x.device, dtype=torch.float32)
    _lib.check(_L().dram_conv3d_simt_wgrad(x.data_ptr(), dy.data_ptr(), dpack.data_ptr(), N, D, H, W, Cin, Cout, ksize,
                                           _stream()), "conv3d_simt_wgrad")
    dw = torch.empty((Cout, Cin, ksize, ksize, ksize), device=x.device, dtype=torch.float32)
    _lib.check(_L().dram_unpack_wgrad_f32(dpack.data_ptr(), dw.data_ptr(), Cout, Cin, ksize, _stream()), "unpack_wgrad_f32")
    return dw


class SplitPlanes:
    """bf16 split planes of a CL volume: hi (+ lo in bf16x3 mode), channels zero-padded to a multiple of 64.
    `second`: a VIRTUAL channel concat — this object holds the first `C1` channels (pitch Cpad1 = hi.shape[-1]) and
    `second` (another SplitPlanes of the same spatial size) the rest; `shape` / `Cpad` describe the concatenated volume.
    The tensor-core convolutions read both through two tensor maps (parts.py:153 without the copy)."""
    __slots__ = ("hi", "lo", "shape", "Cpad", "second")

    def __init__(self, hi, lo, shape, Cpad, second=None):
        self.hi, self.lo, self.shape, self.Cpad, self.second = hi, lo, shape, Cpad, second

    @property
    def Cpad1(self):
        return self.hi.shape[-1]

    def x2(self):
        """(x2_hi ptr, x2_lo ptr, Cin1_pad) for the C ABI"""
        if self.second is None:
            return 0, 0, 0
        return self.second.hi.data_ptr(), _p(self.second.lo), self.Cpad1


def split_bf16(x, three=None):
    N, C, D, H, W = x.shape
    three = (precision() == "bf16x3") if three is None else three
    Cpad = _pad64(C)
    rows = N * D * H * W
    hi = torch.empty((rows, Cpad), device=x.device, dtype=torch.bfloat16)
    lo = torch.empty((rows, Cpad), device=x.device, dtype=torch.bfloat16) if three else None
    _lib.check(_L().dram_split_bf16(x.data_ptr(), hi.data_ptr(), _p(lo), rows, C, Cpad, _stream()), "split_bf16")
    return SplitPlanes(hi, lo, (N, C, D, H, W), Cpad)


def pack_weight_bf16(w, mode, three=None):
    """mode 0: [taps][Cout][pad64(Cin)]; mode 1 (dgrad): [taps][Cin][pad64(Cout)] -> (hi, lo|None, Kpad)"""
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    three = (precision() == "bf16x3") if three is None else three
    rows, K = (Cout, Cin) if mode == 0 else (Cin, Cout)
    Kpad = _pad64(K)
    hi = torch.empty((k ** 3 * rows, Kpad), device=w.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if three else None
    _lib.check(_L().dram_pack_weight_bf16(w.data_ptr(), hi.data_ptr(), _p(lo), Cout, Cin, Kpad, k, mode, _stream()),
               "pack_weight_bf16")
    return hi, lo, Kpad


def conv_umma(xs, w_hi, w_lo, Cout, ksize, scale=None, shift=None, out_planes=False, want_stats=False):
    """xs: SplitPlanes of the input volume; weights packed by pack_weight_bf16 -> CL volume [N,Cout,D,H,W] fp32, or (with
    out_planes, eval mode: scale / shift = folded BatchNorm) the ReLU'd activation as SplitPlanes, no fp32 tensor written.
    want_stats (train mode): -> (y, sums | None); sums = double [2*Cout] (sum y, sum y^2 per channel) accumulated in the
    convolution's epilogue when the kernel that runs this shape supports it (None otherwise: use bn_stats(y))."""
    N, Cin, D, H, W = xs.shape
    _lib.PROFILE.note(flops=2.0 * N * D * H * W * Cin * Cout * ksize ** 3, tag=f"{Cin}->{Cout}@{D}")   # algorithmic (unpadded, 1 pass)
    if out_planes:
        if scale is None or Cout % 64:
            raise _lib.DramLibraryError("conv_umma: plane output needs folded BatchNorm scale/shift and Cout % 64 == 0")
        out = alloc_planes((N, Cout, D, H, W), three=xs.lo is not None)
        _lib.check(_L().dram_conv3d_umma_fwd(xs.hi.data_ptr(), _p(xs.lo), w_hi.data_ptr(), _p(w_lo), scale.data_ptr(),
                                             shift.data_ptr(), None, out.hi.data_ptr(), _p(out.lo), None, *xs.x2(), N, D, H, W,
                                             Cin, xs.Cpad, Cout, ksize, _stream()), "conv3d_umma_fwd")
        return out
    y = new_volume(N, Cout, D, H, W, xs.hi.device)
    partials, rows = None, 0
    if want_stats and scale is None and os.environ.get("DRAM_BN_EPILOGUE", "1") == "1":
        rows = _L().dram_conv3d_umma_fwd_stat_rows(N, D, H, W, Cin, xs.Cpad, Cout, ksize, int(xs.lo is not None), int(w_lo is not None))
        if rows > 0:
            partials = torch.empty((rows, 2, Cout), device=xs.hi.device, dtype=torch.float32)
    _lib.check(_L().dram_conv3d_umma_fwd(xs.hi.data_ptr(), _p(xs.lo), w_hi.data_ptr(), _p(w_lo), _p(scale), _p(shift),
                                         y.data_ptr(), None, None, _p(partials), *xs.x2(), N, D, H, W, Cin, xs.Cpad, Cout,
                                         ksize, _stream()), "conv3d_umma_fwd")
    if not want_stats:
        return y
    if partials is None:
        return y, None
    sums = torch.empty(2 * Cout, device=y.device, dtype=torch.float64)
    ws = torch.empty(_L().dram_bn_stats_from_partials_workspace_bytes(Cout), device=y.device, dtype=torch.uint8)
    _lib.check(_L().dram_bn_stats_from_partials(partials.data_ptr(), rows, Cout, sums.data_ptr(), ws.data_ptr(), _stream()),
               "bn_stats_from_partials")
    return y, sums


def conv_umma_wgrad(dys, xs, Cin, Cout, ksize, stream=None, keep=None, out=None):
    """dys / xs: SplitPlanes of dy and of the layer input -> dw [Cout, Cin, k, k, k] fp32.
    `stream`: raw cudaStream_t to enqueue on (default: the current stream); memory is always allocated on the current
    stream.  `keep`: list that receives the workspace so the caller can keep it alive until that stream is joined."""
    N, _, D, H, W = xs.shape
    nbytes = _L().dram_conv3d_umma_wgrad_workspace_bytes(N, D, H, W, xs.Cpad, dys.Cpad, ksize)
    if nbytes == 0:
        raise _lib.DramLibraryError("conv3d_umma_wgrad: unsupported shape")
    ws = torch.empty(nbytes // 4, device=xs.hi.device, dtype=torch.float32)
    dw = out if out is not None else torch.empty((Cout, Cin, ksize, ksize, ksize), device=xs.hi.device, dtype=torch.float32)
    _lib.PROFILE.note(flops=2.0 * N * D * H * W * Cin * Cout * ksize ** 3, tag=f"{Cin}->{Cout}@{D}")
    _lib.check(_L().dram_conv3d_umma_wgrad(dys.hi.data_ptr(), _p(dys.lo), xs.hi.data_ptr(), _p(xs.lo), dw.data_ptr(),
                                           ws.data_ptr(), *xs.x2(), N, D, H, W, Cin, xs.Cpad, Cout, dys.Cpad, ksize,
                                           _stream() if stream is None else stream), "conv3d_umma_wgrad")
    if keep is not None:
        keep.append(ws)
    return dw


# ------------------------------------------------------------------------------------------------ BN / ReLU / pool
def bn_stats(y):
    N, C, D, H, W = y.shape
    sums = torch.empty(2 * C, device=y.device, dtype=torch.float64)
    _lib.check(_L().dram_bn_stats(y.data_ptr(), sums.data_ptr(), N * D * H * W, C, _stream()), "bn_stats")
    return sums


def bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps, n_updates):
    C = sums.numel() // 2
    out = torch.empty((4, C), device=sums.device, dtype=torch.float32)     # mean, rstd, scale, shift
    _lib.check(_L().dram_bn_finalize(sums.data_ptr(), float(count), _p(gamma), _p(beta), _p(running_mean),
                                     _p(running_var), float(momentum), float(eps), int(n_updates), out[0].data_ptr(),
                                     out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), C, _stream()), "bn_finalize")
    return out[0], out[1], out[2], out[3]


def bn_fold_eval(gamma, beta, running_mean, running_var, eps):
    C = running_mean.numel()
    out = torch.empty((2, C), device=running_mean.device, dtype=torch.float32)
    _lib.check(_L().dram_bn_fold_eval(_p(gamma), _p(beta), running_mean.data_ptr(), running_var.data_ptr(), float(eps),
                                      out[0].data_ptr(), out[1].data_ptr(), C, _stream()), "bn_fold_eval")
    return out[0], out[1]


def bn_relu_apply(y, scale, shift, pool=False):
    N, C, D, H, W = y.shape
    a = new_volume(N, C, D, H, W, y.device)
    pooled = new_volume(N, C, D // 2, H // 2, W // 2, y.device) if pool else None
    _lib.check(_L().dram_bn_relu_apply(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), a.data_ptr(), _p(pooled), N, D,
                                       H, W, C, _stream()), "bn_relu_apply")
    return a, pooled


def alloc_planes(shape, three=None):
    """Uninitialised split planes for a volume of logical shape [N,C,D,H,W] (every kernel that fills them also zeroes
    the pad channels)."""
    N, C, D, H, W = shape
    three = (precision() == "bf16x3") if three is None else three
    Cpad = _pad64(C)
    rows = N * D * H * W
    dev = torch.device("cuda", torch.cuda.current_device())
    hi = torch.empty((rows, Cpad), device=dev, dtype=torch.bfloat16)
    lo = torch.empty((rows, Cpad), device=dev, dtype=torch.bfloat16) if three else None
    return SplitPlanes(hi, lo, (N, C, D, H, W), Cpad)


def bn_relu_apply_planes(y, scale, shift, pool=False):
    """a = relu(y*scale+shift) (and its 2x2x2 max-pool) written directly as split planes -> (a, pooled | None)"""
    N, C, D, H, W = y.shape
    a = alloc_planes((N, C, D, H, W))
    p = alloc_planes((N, C, D // 2, H // 2, W // 2)) if pool else None
    _lib.check(_L().dram_bn_relu_apply_planes(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), a.hi.data_ptr(), _p(a.lo),
                                              _p(p.hi) if pool else 0, _p(p.lo) if pool else 0, N, D, H, W, C, a.Cpad,
                                              _stream()), "bn_relu_apply_planes")
    return a, p


def merge_planes(xs):
    """split planes -> fp32 channels-last volume"""
    if xs.second is not None:
        N, C, D, H, W = xs.shape
        C1 = xs.Cpad1
        a = merge_planes(SplitPlanes(xs.hi, xs.lo, (N, C1, D, H, W), C1))
        return torch.cat([a, merge_planes(xs.second)], dim=1).contiguous(memory_format=CL)
    N, C, D, H, W = xs.shape
    out = new_volume(N, C, D, H, W, xs.hi.device)
    _lib.check(_L().dram_merge_planes(xs.hi.data_ptr(), _p(xs.lo), out.data_ptr(), N * D * H * W, C, xs.Cpad, _stream()),
               "merge_planes")
    return out


def pointwise8_ok(xs, Cout, k):
    """1x1x1 conv to 8 channels straight from planes (the attention reshape heads)"""
    return xs.second is None and k == 1 and isinstance(xs, SplitPlanes) and bool(_L().dram_pointwise8_planes_supported(int(xs.Cpad), int(Cout)))


def pointwise8_planes(xs, w, bias):
    """y [N,8,D,H,W] (channels-last) = Conv3d(C, 8, 1)(planes)"""
    N, C, D, H, W = xs.shape
    y = new_volume(N, 8, D, H, W, xs.hi.device)
    wf = w.detach().reshape(8, C).contiguous()
    _lib.PROFILE.note(bytes=float(N * D * H * W) * (4.0 * C + 32.0))
    _lib.check(_L().dram_pointwise8_planes_fwd(xs.hi.data_ptr(), _p(xs.lo), wf.data_ptr(), _p(bias), y.data_ptr(),
                                               N * D * H * W, C, xs.Cpad, _stream()), "pointwise8_planes_fwd")
    return y


def pointwise8_planes_wgrad(xs, dy, want_bias=True):
    """-> (dw [8,C,1,1,1], dbias [8] | None) for the layer above; dy [N,8,D,H,W] channels-last fp32"""
    N, C, D, H, W = xs.shape
    dw = torch.empty((8, C, 1, 1, 1), device=dy.device, dtype=torch.float32)
    db = torch.empty(8, device=dy.device, dtype=torch.float32) if want_bias else None
    ws = torch.empty(_L().dram_pointwise8_planes_wgrad_workspace_bytes(int(xs.Cpad)), device=dy.device, dtype=torch.uint8)
    _lib.PROFILE.note(bytes=float(N * D * H * W) * (4.0 * C + 32.0))
    _lib.check(_L().dram_pointwise8_planes_wgrad(xs.hi.data_ptr(), _p(xs.lo), dy.data_ptr(), dw.data_ptr(), _p(db),
                                                 ws.data_ptr(), N * D * H * W, C, xs.Cpad, _stream()), "pointwise8_planes_wgrad")
    return dw, db


def bn_relu_bwd_reduce(da, y, scale, shift, mean, rstd, pitch=0, wtop=None):
    """-> double sums [2C] (sum dz, sum dz*xhat); with wtop (fused RAM head, da = g [rows]): [3C+1], see the header"""
    N, C, D, H, W = y.shape
    sums = torch.empty(2 * C if wtop is None else 3 * C + 1, device=y.device, dtype=torch.float64)
    _lib.check(_L().dram_bn_relu_bwd_reduce(da.data_ptr(), int(pitch), _p(wtop), y.data_ptr(), scale.data_ptr(),
                                            shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(),
                                            N * D * H * W, C, _stream()), "bn_relu_bwd_reduce")
    return sums


def bn_relu_bwd_apply(da, y, scale, shift, mean, rstd, gamma, sums, count, pitch=0):
    N, C, D, H, W = y.shape
    dy = new_volume(N, C, D, H, W, y.device)
    _lib.check(_L().dram_bn_relu_bwd_apply(da.data_ptr(), int(pitch), y.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                           _p(mean), _p(rstd), _p(gamma), _p(sums), float(count), dy.data_ptr(),
                                           N * D * H * W, C, _stream()), "bn_relu_bwd_apply")
    return dy


def bn_relu_bwd_apply_planes(da, pitch, wtop, y, scale, shift, mean, rstd, gamma, sums, count):
    """dy of the BatchNorm+ReLU backward as split planes (operand of dgrad / wgrad)"""
    N, C, D, H, W = y.shape
    dys = alloc_planes((N, C, D, H, W), three=grad_planes_three())
    _lib.check(_L().dram_bn_relu_bwd_apply_planes(da.data_ptr(), int(pitch), _p(wtop), y.data_ptr(), scale.data_ptr(),
                                                  shift.data_ptr(), _p(mean), _p(rstd), _p(gamma), _p(sums), float(count),
                                                  dys.hi.data_ptr(), _p(dys.lo), N * D * H * W, C, dys.Cpad, _stream()),
               "bn_relu_bwd_apply_planes")
    return dys


def bn_pool_bwd_reduce(ga, pitch, gp, y, scale, shift, mean, rstd):
    N, C, D, H, W = y.shape
    sums = torch.empty(2 * C, device=y.device, dtype=torch.float64)
    _lib.check(_L().dram_bn_pool_bwd_reduce(_p(ga), int(pitch), _p(gp), y.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                            mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(), N, D, H, W, C, _stream()),
               "bn_pool_bwd_reduce")
    return sums


def bn_pool_bwd_apply_planes(ga, pitch, gp, y, scale, shift, mean, rstd, gamma, sums, count):
    N, C, D, H, W = y.shape
    dys = alloc_planes((N, C, D, H, W), three=grad_planes_three())
    _lib.check(_L().dram_bn_pool_bwd_apply_planes(_p(ga), int(pitch), _p(gp), y.data_ptr(), scale.data_ptr(),
                                                  shift.data_ptr(), _p(mean), _p(rstd), _p(gamma), _p(sums), float(count),
                                                  dys.hi.data_ptr(), _p(dys.lo), N, D, H, W, C, dys.Cpad, _stream()),
               "bn_pool_bwd_apply_planes")
    return dys


def maxpool2_bwd(a, dpooled, da):
    N, C, D, H, W = a.shape
    _lib.check(_L().dram_maxpool2_bwd(a.data_ptr(), dpooled.data_ptr(), da.data_ptr(), N, D, H, W, C, _stream()),
               "maxpool2_bwd")
    return da


# ------------------------------------------------------------------------------------------------ decoder glue
def upsample2x_concat(x, skip):
    N, C1, d, h, w = x.shape
    _, C2, Ds, Hs, Ws = skip.shape
    cat = new_volume(N, C1 + C2, 2 * d, 2 * h, 2 * w, x.device)
    _lib.check(_L().dram_upsample2x_concat_fwd(x.data_ptr(), skip.data_ptr(), cat.data_ptr(), N, d, h, w, C1, Ds, Hs, Ws,
                                               C2, _stream()), "upsample2x_concat_fwd")
    return cat


def upsample2x_concat_planes(xs, skips):
    """planes of x [N,C1,d,h,w] and skip [N,C2,Ds,Hs,Ws] -> planes of cat [N,C1+C2,2d,2h,2w]"""
    N, C1, d, h, w = xs.shape
    _, C2, Ds, Hs, Ws = skips.shape
    cat = alloc_planes((N, C1 + C2, 2 * d, 2 * h, 2 * w), three=xs.lo is not None)
    _lib.check(_L().dram_upsample2x_concat_planes(xs.hi.data_ptr(), _p(xs.lo), skips.hi.data_ptr(), _p(skips.lo),
                                                  cat.hi.data_ptr(), _p(cat.lo), N, d, h, w, C1, xs.Cpad, Ds, Hs, Ws, C2,
                                                  skips.Cpad, cat.Cpad, _stream()), "upsample2x_concat_planes")
    return cat


def virtual_concat_ok(xs, skips):
    """can cat([up2x(x), skip]) stay virtual?  exact x2 sizes (no crop), the upsampled half fills whole 64-channel blocks"""
    N, C1, d, h, w = xs.shape
    _, C2, Ds, Hs, Ws = skips.shape
    return (os.environ.get("DRAM_VIRTUAL_CONCAT", "1") == "1" and (Ds, Hs, Ws) == (2 * d, 2 * h, 2 * w) and C1 % 64 == 0
            and skips.second is None and xs.second is None and (xs.lo is None) == (skips.lo is None))


def upsample2x_virtual_concat(xs, skips):
    """planes of x [N,C1,d,h,w] and skip [N,C2,2d,2h,2w] -> SplitPlanes of cat [N,C1+C2,2d,2h,2w] whose first half is the freshly
    upsampled tensor and whose `second` half IS the skip planes (no copy)"""
    N, C1, d, h, w = xs.shape
    C2 = skips.shape[1]
    up = alloc_planes((N, C1, 2 * d, 2 * h, 2 * w), three=xs.lo is not None)
    _lib.check(_L().dram_upsample2x_planes(xs.hi.data_ptr(), _p(xs.lo), up.hi.data_ptr(), _p(up.lo), N, d, h, w, C1, xs.Cpad,
                                           up.Cpad, _stream()), "upsample2x_planes")
    return SplitPlanes(up.hi, up.lo, (N, C1 + C2, 2 * d, 2 * h, 2 * w), up.Cpad + skips.Cpad, second=skips)


def upsample2x_concat_bwd(dcat, x_shape, skip_shape, want_dskip=True):
    N, C1, d, h, w = x_shape
    _, C2, Ds, Hs, Ws = skip_shape
    dx = new_volume(N, C1, d, h, w, dcat.device)
    dskip = new_volume(N, C2, Ds, Hs, Ws, dcat.device) if want_dskip else None
    _lib.check(_L().dram_upsample2x_concat_bwd(dcat.data_ptr(), dx.data_ptr(), _p(dskip), N, d, h, w, C1, Ds, Hs,
                                               Ws, C2, _stream()), "upsample2x_concat_bwd")
    return dx, dskip


def trilinear_resize(x, size):
    N, C, d, h, w = x.shape
    D, H, W = size
    out = new_volume(N, C, D, H, W, x.device)
    _lib.check(_L().dram_trilinear_resize_fwd(x.data_ptr(), out.data_ptr(), N, d, h, w, D, H, W, C, _stream()),
               "trilinear_resize_fwd")
    return out


def trilinear_resize_bwd(dout, src_size):
    N, C, D, H, W = dout.shape
    d, h, w = src_size
    dx = new_volume(N, C, d, h, w, dout.device)
    _lib.check(_L().dram_trilinear_resize_bwd(dout.data_ptr(), dx.data_ptr(), N, d, h, w, D, H, W, C, _stream()),
               "trilinear_resize_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ RAM head
def ram_reduce(feat, w, b, scale=None, shift=None):
    """feat CL volume [N,C,D,H,W]; w [O,C]; b [O] -> CL volume [N,O,D,H,W]"""
    N, C, D, H, W = feat.shape
    O = w.shape[0]
    ram = new_volume(N, O, D, H, W, feat.device)
    _lib.PROFILE.note(bytes=4.0 * N * D * H * W * (C + O))                           # read features once, write the map
    _lib.check(_L().dram_ram_reduce_fwd(feat.data_ptr(), _p(scale), _p(shift), w.data_ptr(), b.data_ptr(), ram.data_ptr(),
                                        N * D * H * W, C, O, _stream()), "ram_reduce_fwd")
    return ram


def ram_reduce_bwd(dram, feat, w):
    N, C, D, H, W = feat.shape
    O = w.shape[0]
    dfeat = new_volume(N, C, D, H, W, feat.device)
    dwb = torch.empty(O * C + O, device=feat.device, dtype=torch.float64)
    _lib.check(_L().dram_ram_reduce_bwd(dram.data_ptr(), feat.data_ptr(), w.data_ptr(), dfeat.data_ptr(), dwb.data_ptr(),
                                        N * D * H * W, C, O, _stream()), "ram_reduce_bwd")
    return dfeat, dwb[:O * C].view(O, C).float(), dwb[O * C:].float()


def masked_pool(x, mask, use_sigmoid=False, mode_gt0=False):
    """x, mask: [B, V] contiguous fp32 -> double [B, 2] (sum f(x)*m, sum m)"""
    B, V = x.shape
    out = torch.empty((B, 2), device=x.device, dtype=torch.float64)
    _lib.check(_L().dram_masked_pool_fwd(x.data_ptr(), mask.data_ptr(), out.data_ptr(), B, V, int(use_sigmoid),
                                         int(mode_gt0), _stream()), "masked_pool_fwd")
    return out


def masked_pool_bwd(x, mask, g, use_sigmoid=False, mode_gt0=False):
    B, V = x.shape
    dx = torch.empty_like(x)
    _lib.check(_L().dram_masked_pool_bwd(x.data_ptr(), mask.data_ptr(), g.data_ptr(), dx.data_ptr(), B, V,
                                         int(use_sigmoid), int(mode_gt0), _stream()), "masked_pool_bwd")
    return dx


def int_reg_loss(pool_pred, pool_rub, band, w, band_width):
    """pooled sums [B,2] double (x2), band [B,2] double, w [B] fp32 -> (loss [1] fp32, g [B] fp32 = d loss / d pool_pred[:,0])"""
    B = pool_pred.shape[0]
    loss = torch.empty(1, device=pool_pred.device, dtype=torch.float32)
    g = torch.empty(B, device=pool_pred.device, dtype=torch.float32)
    _lib.check(_L().dram_int_reg_loss(pool_pred.data_ptr(), pool_rub.data_ptr(), band.data_ptr(), w.data_ptr(),
                                      float(band_width), loss.data_ptr(), g.data_ptr(), B, _stream()), "int_reg_loss")
    return loss, g


def boot_bce_sums(dense, refined, lobes, lesions, keep, eps=1e-7):
    """[B,V] fp32 logits / masks, keep [B] -> double [7] (see dram_boot_bce_fwd)"""
    B, V = refined.shape
    sums = torch.empty(7, device=refined.device, dtype=torch.float64)
    _lib.check(_L().dram_boot_bce_fwd(dense.data_ptr(), refined.data_ptr(), lobes.data_ptr(), lesions.data_ptr(),
                                      keep.data_ptr(), sums.data_ptr(), B, V, float(eps), _stream()), "boot_bce_fwd")
    return sums


def boot_bce_grad(dense, refined, lobes, lesions, keep, coef, eps=1e-7):
    """-> d loss / d refined [B,V] for loss = coef . sums[3:7]; coef: float32 [4] on the device"""
    B, V = refined.shape
    out = torch.empty_like(refined)
    _lib.check(_L().dram_boot_bce_bwd(dense.data_ptr(), refined.data_ptr(), lobes.data_ptr(), lesions.data_ptr(),
                                      keep.data_ptr(), coef.data_ptr(), out.data_ptr(), B, V, float(eps), _stream()),
               "boot_bce_bwd")
    return out


def ram_activation(x, act):
    """act(x) elementwise: 1 = sigmoid, 2 = relu"""
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(_L().dram_ram_activation(x.data_ptr(), out.data_ptr(), x.numel(), int(act), _stream()), "ram_activation")
    return out


def ram_upsample_mask_scatter(ram, crop_mask, heat, offset, act, gain=1.0, maxval=None):
    """ram [d,h,w] fp32; crop_mask [cd,ch,cw] uint8; heat [SD,SH,SW] fp32 (in place) or None."""
    d, h, w = ram.shape
    cd, ch, cw = crop_mask.shape
    if crop_mask.dtype != torch.uint8 or not crop_mask.is_contiguous():
        raise _lib.DramLibraryError("ram_upsample_mask_scatter: crop_mask must be a contiguous uint8 tensor")
    SD, SH, SW = heat.shape if heat is not None else (cd, ch, cw)
    oz, oy, ox = offset
    _lib.check(_L().dram_ram_upsample_mask_scatter(ram.data_ptr(), crop_mask.data_ptr(), _p(heat), _p(maxval), d, h, w,
                                                   cd, ch, cw, SD, SH, SW, oz, oy, ox, int(act), float(gain), _stream()),
               "ram_upsample_mask_scatter")


# ------------------------------------------------------------------------------------------------ PCM
MERGE_FLAGS = {"sm": 0, "smrelu": 1, "scaled_dot_product": 2, "scaled_dot_product_relu": 3, "smscaled": 4}


def pcm_num_offsets(connectivity, self_loop):
    return _L().dram_pcm_num_offsets(int(connectivity), int(bool(self_loop)))


def pcm_fwd(f, cam, tw, tb, pw, pb, connectivity, self_loop, flags, keep_stats=True):
    """f CL volume [B,Cf,D,H,W]; cam [B,1,D,H,W] -> (out [B,1,D,H,W], qk (blocks of 32 voxels x 2F features),
    softmax stats [B*V, 4] | None)"""
    B, Cf, D, H, W = f.shape
    F = tw.shape[0]
    V = D * H * W
    two_kernels = keep_stats or bool(os.environ.get("DRAM_PCM_TWO_KERNELS"))
    qk = torch.empty(_L().dram_pcm_qk_floats(B * V, F), device=f.device, dtype=torch.float32) if two_kernels else None
    stats = torch.empty((B * V, 4), device=f.device, dtype=torch.float32) if keep_stats else None
    out = torch.empty((B, 1, D, H, W), device=f.device, dtype=torch.float32)
    _lib.PROFILE.note(bytes=4.0 * B * V * (Cf + 2))
    _lib.check(_L().dram_pcm_fwd(f.data_ptr(), cam.data_ptr(), tw.data_ptr(), tb.data_ptr(), pw.data_ptr(), pb.data_ptr(),
                                 _p(qk), _p(stats), out.data_ptr(), B, D, H, W, Cf, F, int(connectivity),
                                 int(bool(self_loop)), int(flags), _stream()), "pcm_fwd")
    return out, qk, stats


def pcm_bwd(f, cam, tw, pw, qk, stats, dout, connectivity, self_loop, flags):
    B, Cf, D, H, W = f.shape
    F = tw.shape[0]
    dqk = torch.empty(_L().dram_pcm_bwd_ws_floats(B * D * H * W, Cf, F), device=f.device, dtype=torch.float32)
    dcam = torch.empty((B, 1, D, H, W), device=f.device, dtype=torch.float32)
    df = new_volume(B, Cf, D, H, W, f.device)
    dparams = torch.empty(2 * F * (Cf + 1), device=f.device, dtype=torch.float64)
    _lib.PROFILE.note(bytes=4.0 * B * D * H * W * (2 * Cf + 3))
    _lib.check(_L().dram_pcm_bwd(f.data_ptr(), cam.data_ptr(), tw.data_ptr(), pw.data_ptr(), qk.data_ptr(), stats.data_ptr(),
                                 dout.data_ptr(), dqk.data_ptr(), dcam.data_ptr(), df.data_ptr(),
                                 dparams.data_ptr(), B, D, H, W, Cf, F, int(connectivity), int(bool(self_loop)),
                                 int(flags), _stream()), "pcm_bwd")
    dp = dparams.float()
    n = F * Cf
    return dcam, df, dp[:n].view(F, Cf), dp[n:n + F], dp[n + F:2 * n + F].view(F, Cf), dp[2 * n + F:]


# ------------------------------------------------------------------------------------------------ scan pre/post-processing
_DTYPE_CODE = {torch.float32: 0, torch.int16: 1, torch.uint8: 2}


_HOST_BUFS = {}


def read_small(t):
    """device tensor (<= 1 MiB, 4-byte multiple) -> CPU tensor through pinned memory written by a kernel (no copy engine,
    see dram_store_to_host); blocks the host until the values are there."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    key = (nbytes, t.dtype)
    buf = _HOST_BUFS.get(key)
    if buf is None:
        buf = _HOST_BUFS[key] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
    _lib.check(_L().dram_store_to_host(t.data_ptr(), buf.data_ptr(), nbytes, _stream()), "store_to_host")
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()
    return buf.clone().view(t.shape)


def label_bboxes(labels, nlabels=5):
    """labels: uint8 [D,H,W] CUDA -> int32 [(nlabels+1), 6] (min z,y,x, max z,y,x inclusive); row 0 unused."""
    _req(labels, "labels", torch.uint8)
    D, H, W = labels.shape
    out = torch.empty((nlabels + 1, 6), device=labels.device, dtype=torch.int32)
    _lib.check(_L().dram_label_bboxes(labels.data_ptr(), D, H, W, nlabels, out.data_ptr(), _stream()), "label_bboxes")
    return out


def lobe_chunk_preprocess(scan, labels, label, crop, window, pad_value, img_out, msk_out, spacing=(1.0, 1.0, 1.0)):
    """scan int16 / labels uint8 [SD,SH,SW] with voxel `spacing` (z,y,x); crop = ((z0,z1),(y0,y1),(x0,x1)); writes
    img_out/msk_out [d,h,w] fp32."""
    _req(scan, "scan", torch.int16)
    _req(labels, "labels", torch.uint8)
    SD, SH, SW = scan.shape
    (z0, z1), (y0, y1), (x0, x1) = crop
    d, h, w = img_out.shape
    _lib.check(_L().dram_lobe_chunk_preprocess(scan.data_ptr(), labels.data_ptr(), SD, SH, SW, int(label), z0, y0, x0,
                                               z1 - z0, y1 - y0, x1 - x0, float(window[0]), float(window[1]),
                                               float(pad_value), float(spacing[0]), float(spacing[1]), float(spacing[2]),
                                               img_out.data_ptr(), msk_out.data_ptr(), d, h, w, _stream()),
               "lobe_chunk_preprocess")


def itk_resample(src, new_size, mode="linear", in_spacing=None, out_spacing=None):
    """Volume resample with SimpleITK ResampleImageFilter semantics (identity transform, shared origin): `src` has voxel
    spacing in_spacing (z,y,x; default 1), the result has `new_size` voxels of out_spacing (default: Resample('fixed_size'),
    in_spacing * in_size / out_size, data_transforms.py:170-175).  Bit-identical to the float64 oracle."""
    if src.dtype not in _DTYPE_CODE or not src.is_cuda or not src.is_contiguous():
        raise _lib.DramLibraryError("itk_resample: need a contiguous CUDA tensor of dtype float32 / int16 / uint8")
    d, h, w = src.shape
    D, H, W = (int(s) for s in new_size)
    isp = [1.0, 1.0, 1.0] if in_spacing is None else [float(v) for v in in_spacing]
    osp = [s * (float(n_in) / float(n_out)) for s, n_in, n_out in zip(isp, (d, h, w), (D, H, W))] if out_spacing is None \
        else [float(v) for v in out_spacing]
    dst = torch.empty((D, H, W), device=src.device, dtype=src.dtype)
    _lib.check(_L().dram_itk_resample(src.data_ptr(), dst.data_ptr(), _DTYPE_CODE[src.dtype], d, h, w, D, H, W,
                                      isp[0], isp[1], isp[2], osp[0], osp[1], osp[2], 0 if mode == "linear" else 1,
                                      _stream()), "itk_resample")
    return dst


def ram_upsample_label_scatter(ram, labels, label, heat, crop, act, gain=1.0):
    """ram [d,h,w] fp32 -> heat[crop] (scan-sized fp32, in place) where labels == label."""
    d, h, w = ram.shape
    SD, SH, SW = heat.shape
    (z0, z1), (y0, y1), (x0, x1) = crop
    _lib.PROFILE.note(bytes=4.0 * d * h * w + 5.0 * (z1 - z0) * (y1 - y0) * (x1 - x0))
    _lib.check(_L().dram_ram_upsample_label_scatter(ram.data_ptr(), labels.data_ptr(), int(label), heat.data_ptr(), d, h, w,
                                                    z1 - z0, y1 - y0, x1 - x0, SD, SH, SW, z0, y0, x0, int(act), float(gain),
                                                    _stream()), "ram_upsample_label_scatter")


def masked_hist_u8(values, labels, lo, hi):
    """-> int64 [256] histogram of uint8(window(values,(lo,hi))*255) over labels > 0 (device tensor)."""
    if values.dtype not in (torch.float32, torch.int16):
        raise _lib.DramLibraryError("masked_hist_u8: values must be float32 or int16")
    hist = torch.empty(256, device=values.device, dtype=torch.int32)
    _lib.check(_L().dram_masked_hist_u8(values.data_ptr(), _DTYPE_CODE[values.dtype], labels.data_ptr(), values.numel(),
                                        float(lo), float(hi), hist.data_ptr(), _stream()), "masked_hist_u8")
    return hist


def labelled_mean(values, labels):
    """mean of `values` (fp32) over voxels with labels > 0 (0 when there are none) -> 0-dim fp32 device tensor"""
    _req(labels, "labels", torch.uint8)
    if values.dtype != torch.float32 or not values.is_contiguous() or values.numel() != labels.numel():
        raise _lib.DramLibraryError("labelled_mean: values must be a contiguous float32 tensor of the labels' size")
    out = torch.empty(2, device=values.device, dtype=torch.float64)
    ws = torch.empty(_L().dram_labelled_sum_workspace_bytes(), device=values.device, dtype=torch.uint8)
    _lib.check(_L().dram_labelled_sum(values.data_ptr(), labels.data_ptr(), values.numel(), out.data_ptr(), ws.data_ptr(),
                                      _stream()), "labelled_sum")
    return (out[0] / out[1].clamp_min(1.0)).float()


def threshold_masks(heat, th, scan=None, vessel=None, th2=0.0, window=(-1150.0, 350.0), want_post=True):
    lesion = torch.empty(heat.shape, device=heat.device, dtype=torch.uint8)
    post = torch.empty_like(lesion) if (want_post and scan is not None) else None
    _lib.check(_L().dram_threshold_masks(heat.data_ptr(), _p(scan), _p(vessel), heat.numel(), float(th), float(th2),
                                         float(window[0]), float(window[1]), lesion.data_ptr(), _p(post), _stream()),
               "threshold_masks")
    return lesion, post
