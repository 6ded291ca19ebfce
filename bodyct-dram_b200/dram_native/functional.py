"""torch.autograd.Function wrappers: each forward/backward is a short sequence of libdram_b200 kernel launches.

Granularity follows the reference's building blocks (parts.py): one Function per [Conv3d -> BatchNorm3d -> ReLU
(-> MaxPool3d)] unit, one for upsample+concat, one for the RAM reduce, etc.  Tensors crossing Function boundaries
are logical NCDHW / channels-last memory fp32.
"""
import os

import torch
import torch.optim.optimizer as _torch_optimizer

from . import dist as ddist
from . import ops


def _cl_grad(g, like):
    if g is None:
        return None
    return ops.to_cl(g, "grad")


class _WeightCache:
    """Packed weights live ON the parameter object (attribute `_dram_packs`) and are rebuilt whenever the parameter
    may have changed: `_version` bumps on load_state_dict / copy_ / foreach optimizers, and `generation` bumps after
    EVERY optimizer step (global post-step hook below: fused Adam updates parameters without touching `_version`),
    after CUDA-graph replays and before a capture (so the pack kernels are part of the captured step).  Nothing is
    keyed on addresses alone: the caching allocator reuses them across models."""

    def __init__(self):
        self.misses = 0
        self.generation = 0

    def invalidate(self):
        self.generation += 1

    def get(self, w, kind, builder):
        packs = getattr(w, "_dram_packs", None)
        if packs is None:
            packs = {}
            try:
                w._dram_packs = packs
            except AttributeError:
                pass
        key = (kind, ops.precision())
        hit = packs.get(key)
        if hit is not None and hit[0] == (w._version, self.generation) and hit[1] == w.data_ptr():
            return hit[2]
        self.misses += 1
        val = builder()
        packs[key] = ((w._version, self.generation), w.data_ptr(), val)
        return val


WEIGHTS = _WeightCache()
_torch_optimizer.register_optimizer_step_post_hook(lambda *_a, **_k: WEIGHTS.invalidate())


def conv_forward(x, w, bias=None):
    """Raw convolution (k in {1,3}, pad k//2).  Returns (y, saved_input) where saved_input is what wgrad needs."""
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    if ops.umma_ok_fwd(Cin, Cout, k) and bias is None:
        xs = ops.split_bf16(x)
        w_hi, w_lo, _ = WEIGHTS.get(w, "bf16_fwd", lambda: ops.pack_weight_bf16(w.detach(), 0))
        return ops.conv_umma(xs, w_hi, w_lo, Cout, k), xs
    pack = WEIGHTS.get(w, "f32_fwd", lambda: ops.pack_weight_f32(w.detach(), 0))
    return ops.conv_simt(x, pack, bias, Cout, k), x


def conv_dgrad(dy, w, dys=None):
    """dx = conv(dy, flipped/transposed w).  `dys` = already split planes of dy (shared with wgrad)."""
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    if ops.umma_ok_fwd(Cout, Cin, k):
        dys = dys if dys is not None else ops.split_bf16(dy, three=ops.grad_planes_three())
        w_hi, w_lo, _ = WEIGHTS.get(w, "bf16_dgrad", lambda: ops.pack_weight_bf16(w.detach(), 1))
        return ops.conv_umma(dys, w_hi, w_lo, Cin, k)
    pack = WEIGHTS.get(w, "f32_dgrad", lambda: ops.pack_weight_f32(w.detach(), 1))
    return ops.conv_simt(dy, pack, None, Cin, k)


class _SideStream:
    """Weight gradients are off the critical path of the backward pass (nothing but the optimizer reads them), so the
    tensor-core wgrad kernels are enqueued on a second stream: the HBM-bound BatchNorm-backward / upsample-adjoint kernels
    of the NEXT layer then run concurrently with them (they fit next to a persistent 1-CTA-per-SM conv kernel: few
    registers, no dynamic shared memory).  The main stream joins at the end of every backward pass (autograd engine
    callback) and before a data-parallel gradient bucket is reduced; operands are kept alive until then.  Works inside
    CUDA-graph capture (the fork and the join are captured as graph edges).
    OFF by default (DRAM_WGRAD_STREAM=1 enables it): measured on B200 at the bench configuration the step is limited by the
    board power cap (sw_power_cap, ~1 kW) — with the overlap the SM clock drops from 1740 to 1635 MHz and the step time is
    unchanged (70.55 vs 70.34 ms), see DESIGN.md section 5."""

    def __init__(self):
        self.streams = {}
        self.keep = []
        self.join_queued = False
        self.used = False
        self.main = None

    def enabled(self):
        return os.environ.get("DRAM_WGRAD_STREAM", "0") == "1"

    def stream(self):
        dev = torch.cuda.current_device()
        if dev not in self.streams:
            self.streams[dev] = torch.cuda.Stream(device=dev)
        return self.streams[dev]

    def fork(self):
        """-> side stream ordered after everything enqueued on the current stream so far, or None outside a backward
        pass (no engine callback to join it: the caller stays on one stream)"""
        if not self.join_queued:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self.join)
                self.join_queued = True
            except RuntimeError:
                return None
        side = self.stream()
        self.main = torch.cuda.current_stream()      # the stream the autograd engine runs this node on
        side.wait_stream(self.main)
        self.used = True
        return side

    def join(self):
        self.join_queued = False
        if self.used:
            # not current_stream(): final callbacks may run on an engine thread whose current stream is the default one
            self.main.wait_stream(self.streams[self.main.device.index])
            self.used = False
        self.keep.clear()


SIDE = _SideStream()


def conv_wgrad(saved_in, dy, w, dys=None):
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    if isinstance(saved_in, ops.SplitPlanes):
        dys = dys if dys is not None else ops.split_bf16(dy, three=ops.grad_planes_three())
        # an existing .grad means AccumulateGrad will `+=` this result on the main stream before the end-of-backward join:
        # stay on one stream then (gradient accumulation); likewise while per-kernel events are being recorded
        if SIDE.enabled() and not ops._lib.PROFILE.enabled and w.grad is None:
            side = SIDE.fork()
            if side is not None:
                SIDE.keep.append((dys, saved_in))
                return ops.conv_umma_wgrad(dys, saved_in, Cin, Cout, k, stream=side.cuda_stream, keep=SIDE.keep,
                                           out=ddist.grad_out(w))
        # data parallel: dw is written straight into the flat all-reduce buffer (dist.GradReducer.grad_view)
        return ops.conv_umma_wgrad(dys, saved_in, Cin, Cout, k, out=ddist.grad_out(w) if ddist.active() else None)
    return ops.conv_simt_wgrad(saved_in, dy, k)


class Act:
    """An activation volume that exists ONLY as bf16 split planes (the tensor-core convolutions' operand format).

    `planes` holds the data; `t` is the autograd handle: an fp32 tensor of the logical shape [N, C, D, H, W] with no
    storage behind it (one element, expanded), so the graph, shapes and gradient flow are ordinary PyTorch while the
    fp32 activation is never written.  Gradients w.r.t. `t` are real dense fp32 tensors."""
    __slots__ = ("t", "planes")

    def __init__(self, t, planes):
        self.t, self.planes = t, planes

    @property
    def shape(self):
        return self.planes.shape

    def detach(self):
        return Act(self.t.detach(), self.planes)


def _handle(shape, device):
    return torch.empty(1, device=device, dtype=torch.float32).expand(tuple(shape))


def planes_enabled():
    return ops.conv_path() == "umma"


def _grad_rows(g, name="grad"):
    """-> (tensor to take the pointer from, row pitch in floats) for a gradient that is channels-last dense OR a channel
    slice of a wider channels-last tensor (read in place); anything else is converted with our layout kernel."""
    if g.dim() == 5 and g.stride(1) == 1 and g.is_cuda and g.dtype == torch.float32:
        N, C, D, H, W = g.shape
        P = g.stride(4)
        if P >= C and P % 4 == 0 and g.stride(3) == W * P and g.stride(2) == H * W * P and g.stride(0) == D * H * W * P \
                and g.data_ptr() % 16 == 0:
            return g, P
    return ops.to_cl(g, name), g.shape[1]


def _bn_fold_cached(gamma, beta, running_mean, running_var, eps):
    """eval-mode scale / shift of a BatchNorm.  Deliberately NOT cached: the running statistics are updated in place by
    kernels (also from a replayed CUDA graph, where no Python runs), so no host-side key can tell a stale fold from a fresh
    one; the fold is one tiny launch."""
    return ops.bn_fold_eval(gamma, beta, running_mean, running_var, eps)


# Grad mode is always off INSIDE autograd.Function.forward and ctx.needs_input_grad reflects requires_grad of the parameters
# whatever the mode: the wrappers record the caller's grad mode here right before .apply (single-threaded).
_GRAD_OFF = [False]


class ConvBnRelu(torch.autograd.Function):
    """[Conv3d(bias optional) -> BatchNorm3d -> ReLU (-> MaxPool3d(2,2,0))]  — parts.py:103-110,184-196.

    Input: fp32 volume `x`, or split planes (x_hi, x_lo) with `x` their storage-less handle.  Output: fp32 volumes, or
    (out_planes) planes + handles.  Returns (a, a_hi, a_lo, pooled, p_hi, p_lo); unused slots are None."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, w, bias, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates,
                pool, out_planes):
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        ctx.set_materialize_grads(False)       # the plane outputs carry no gradient: do not let autograd zero-fill GBs for them
        umma = ops.umma_ok_fwd(Cin, Cout, k) and bias is None
        # the backward writes dy straight to planes when both consumers (dgrad, wgrad) are tensor-core kernels
        planes_dy = umma and Cout % 8 == 0 and (ops.umma_ok_fwd(Cout, Cin, k) or not ctx.needs_input_grad[0])
        xs = x_real = None
        pointwise = False
        if x_hi is not None:
            xs = _planes_of(x, x_hi, x_lo)                     # incl. the second operand of a virtual concat
            pointwise = not umma and not ctx.needs_input_grad[0] and ops.pointwise8_ok(xs, Cout, k)
            if not umma and not pointwise:
                x_real = ops.merge_planes(xs)
        else:
            x_real = ops.to_cl(x, "conv input")
        sums = None
        if umma:
            if xs is None:
                xs = ops.split_bf16(x_real)
            w_hi, w_lo, _ = WEIGHTS.get(w, "bf16_fwd", lambda: ops.pack_weight_bf16(w.detach(), 0))
            if (not training and not pool and bool(out_planes) and planes_enabled() and Cout % 64 == 0
                    and _GRAD_OFF[0] and os.environ.get("DRAM_EVAL_EPILOGUE", "1") == "1"):
                # inference: folded BatchNorm + ReLU in the convolution's epilogue, written as the next layer's planes
                scale, shift = _bn_fold_cached(gamma, beta, running_mean, running_var, eps)
                ap = ops.conv_umma(xs, w_hi, w_lo, Cout, k, scale, shift, out_planes=True)
                ctx.mark_non_differentiable(ap.hi, *([ap.lo] if ap.lo is not None else []))
                return _handle(ap.shape, ap.hi.device), ap.hi, ap.lo, None, None, None
            y, sums = ops.conv_umma(xs, w_hi, w_lo, Cout, k, want_stats=True) if training else (ops.conv_umma(xs, w_hi, w_lo, Cout, k), None)
        elif pointwise:
            y = ops.pointwise8_planes(xs, w, bias)       # reshape heads: straight from the planes, no fp32 copy of the input
        else:
            pack = WEIGHTS.get(w, "f32_fwd", lambda: ops.pack_weight_f32(w.detach(), 0))
            y = ops.conv_simt(x_real, pack, bias, Cout, k)
        N, C, D, H, W = y.shape
        count = N * D * H * W
        if training:
            if sums is None:                       # kernels without the statistics epilogue (CUDA-core layers, ragged tiles)
                sums = ops.bn_stats(y)
            pg = ddist.peer()
            if pg is not None and sums.numel() < pg.max_doubles:
                # data parallel: the statistics exchange over NVLink peer memory and the finalize are ONE kernel
                mean, rstd, scale, shift = pg.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps,
                                                          n_updates)
                count = float(count) * ddist.world_size()
            else:
                count = ddist.allreduce_stats(sums, count)
                mean, rstd, scale, shift = ops.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum,
                                                           eps, n_updates)
        else:
            scale, shift = _bn_fold_cached(gamma, beta, running_mean, running_var, eps)
            if not _GRAD_OFF[0]:
                mean, rstd = running_mean.clone(), torch.rsqrt(running_var + eps)  # only for dgamma/dbeta in eval mode
            else:
                mean = rstd = None                                                  # inference: nothing is kept
        planes_out = bool(out_planes) and planes_enabled() and C % 8 == 0 and (planes_dy or not pool)
        a_hi = a_lo = p_hi = p_lo = pooled = None
        if planes_out:
            ap, pp = ops.bn_relu_apply_planes(y, scale, shift, pool)
            a, a_hi, a_lo = _handle(y.shape, y.device), ap.hi, ap.lo
            if pool:
                pooled, p_hi, p_lo = _handle(pp.shape, y.device), pp.hi, pp.lo
        else:
            a, pooled = ops.bn_relu_apply(y, scale, shift, pool)
        ctx.training, ctx.pool, ctx.count, ctx.has_bias, ctx.umma = training, pool, count, bias is not None, umma
        ctx.planes_dy = planes_dy
        ctx.pointwise = pointwise
        ctx.saved_planes = xs if (umma or pointwise) else None
        ctx.save_for_backward(w, gamma, y, scale, shift, a if (pool and not planes_dy) else None, mean, rstd,
                              None if (umma or pointwise) else x_real)
        ctx.mark_non_differentiable(*[t for t in (a_hi, a_lo, p_hi, p_lo) if t is not None])
        return a, a_hi, a_lo, pooled, p_hi, p_lo

    @staticmethod
    def backward(ctx, ga, _g1, _g2, gp, _g3, _g4):
        w, gamma, y, scale, shift, a, mean, rstd, x_plain = ctx.saved_tensors
        if ga is None and gp is None:
            return (None,) * 15
        N, C, D, H, W = y.shape
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        want_gb = ctx.needs_input_grad[5] or ctx.needs_input_grad[6]
        dgamma = dbeta = dys = dy = None
        if ctx.planes_dy:
            # dy goes straight to split planes, the operand of dgrad and wgrad
            if ctx.pool:
                ga_t, pitch = _grad_rows(ga) if ga is not None else (None, 0)
                gp_t = ops.to_cl(gp, "grad") if gp is not None else None
                if ga_t is None and gp_t is None:
                    ga_t, pitch = torch.zeros_like(y), C
                red = lambda: ops.bn_pool_bwd_reduce(ga_t, pitch, gp_t, y, scale, shift, mean, rstd)
                app = lambda gs, cnt: ops.bn_pool_bwd_apply_planes(ga_t, pitch, gp_t, y, scale, shift, mean if gs is not None else None,
                                                                   rstd if gs is not None else None, gamma, gs, cnt)
            else:
                ga_t, pitch = _grad_rows(ga)
                red = lambda: ops.bn_relu_bwd_reduce(ga_t, y, scale, shift, mean, rstd, pitch=pitch)
                app = lambda gs, cnt: ops.bn_relu_bwd_apply_planes(ga_t, pitch, None, y, scale, shift, mean if gs is not None else None,
                                                                   rstd if gs is not None else None, gamma, gs, cnt)
            if ctx.training:
                sums = red()
                dbeta, dgamma = sums[:C].float(), sums[C:].float()
                dys = app(ddist.allreduce_sums(sums), ctx.count)
            else:
                dys = app(None, 1.0)
                if want_gb:
                    sums = red()
                    dbeta, dgamma = sums[:C].float(), sums[C:].float()
        else:
            if ctx.pool:
                if ga is None:
                    da = torch.zeros_like(y)
                else:
                    da = ops.to_cl(ga, "grad").clone() if gp is not None else ops.to_cl(ga, "grad")
                if gp is not None:
                    ops.maxpool2_bwd(a, ops.to_cl(gp, "grad"), da)
                pitch = C
            else:
                da, pitch = _grad_rows(ga)
            if ctx.training:
                sums = ops.bn_relu_bwd_reduce(da, y, scale, shift, mean, rstd, pitch=pitch)
                dbeta, dgamma = sums[:C].float(), sums[C:].float()
                gsums = ddist.allreduce_sums(sums)
                dy = ops.bn_relu_bwd_apply(da, y, scale, shift, mean, rstd, gamma, gsums, ctx.count, pitch=pitch)
            else:
                dy = ops.bn_relu_bwd_apply(da, y, scale, shift, None, None, gamma, None, 1.0, pitch=pitch)
                if want_gb:
                    sums = ops.bn_relu_bwd_reduce(da, y, scale, shift, mean, rstd, pitch=pitch)
                    dbeta, dgamma = sums[:C].float(), sums[C:].float()
        saved_in = ctx.saved_planes if ctx.umma else x_plain
        if dys is None and (ctx.umma or (ctx.needs_input_grad[0] and ops.umma_ok_fwd(Cout, Cin, k))):
            dys = ops.split_bf16(dy, three=ops.grad_planes_three())
        if ctx.pointwise:
            dw, dbias = ops.pointwise8_planes_wgrad(ctx.saved_planes, dy, ctx.has_bias)
            return None, None, None, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None
        dx = conv_dgrad(dy, w, dys) if ctx.needs_input_grad[0] else None
        dw = conv_wgrad(saved_in, dy, w, dys) if ctx.needs_input_grad[3] else None
        dbias = dy.sum(dim=(0, 2, 3, 4)) if ctx.has_bias else None
        return dx, None, None, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None


def _wrap(t, hi, lo):
    if hi is None:
        return t
    return Act(t, ops.SplitPlanes(hi, lo, tuple(t.shape), hi.shape[-1]))


def _unwrap(x):
    return (x.t, x.planes.hi, x.planes.lo) if isinstance(x, Act) else (x, None, None)


def conv_bn_relu(x, w, bias, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates=1, pool=False,
                 out_planes=False):
    """One [conv -> BN -> ReLU (-> pool)] unit on a tensor or an `Act`; returns tensors, or `Act`s when out_planes is set
    and the shape allows it.  With pool: (a, pooled)."""
    _GRAD_OFF[0] = not torch.is_grad_enabled()
    a, a_hi, a_lo, p, p_hi, p_lo = ConvBnRelu.apply(*_unwrap(x), w, bias, gamma, beta, running_mean, running_var, training,
                                                    momentum, eps, n_updates, pool, out_planes)
    if pool:
        return _wrap(a, a_hi, a_lo), _wrap(p, p_hi, p_lo)
    return _wrap(a, a_hi, a_lo)


def _planes_of(t, hi, lo):
    """(handle, hi, lo) as they travel through the autograd Functions -> SplitPlanes; a virtual concat's second operand rides
    on the hi tensor (upsample_concat)"""
    second = getattr(hi, "_dram_second", None)
    return ops.SplitPlanes(hi, lo, tuple(t.shape), hi.shape[-1] + (second.Cpad if second is not None else 0), second)


class Materialize(torch.autograd.Function):
    """planes -> fp32 channels-last volume (for consumers outside the tensor-core path); gradient passes through."""

    @staticmethod
    def forward(ctx, t, hi, lo):
        return ops.merge_planes(_planes_of(t, hi, lo))

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def as_tensor(x):
    return Materialize.apply(x.t, x.planes.hi, x.planes.lo) if isinstance(x, Act) else x


class ConvBnReluRam(torch.autograd.Function):
    """Last decoder unit fused with the RAM head: [Conv3d -> BatchNorm3d -> ReLU] -> top_layer (1x1x1, C -> 1)
    (parts.py:103-110 + models.py:109-110,145).  The unit's activation is never stored: the RAM reduce applies BN+ReLU
    to the raw conv output on the fly, and the backward forms da = g (x) w_top inside the BatchNorm backward kernels."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, w, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates, w_top,
                b_top):
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        xs = ops.SplitPlanes(x_hi, x_lo, tuple(x.shape), x_hi.shape[-1]) if x_hi is not None else \
            ops.split_bf16(ops.to_cl(x, "conv input"))
        w_hi, w_lo, _ = WEIGHTS.get(w, "bf16_fwd", lambda: ops.pack_weight_bf16(w.detach(), 0))
        y, sums = ops.conv_umma(xs, w_hi, w_lo, Cout, k, want_stats=True) if training else (ops.conv_umma(xs, w_hi, w_lo, Cout, k), None)
        N, C, D, H, W = y.shape
        count = N * D * H * W
        if training:
            if sums is None:
                sums = ops.bn_stats(y)
            pg = ddist.peer()
            if pg is not None and sums.numel() < pg.max_doubles:
                # data parallel: the statistics exchange over NVLink peer memory and the finalize are ONE kernel
                mean, rstd, scale, shift = pg.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps,
                                                          n_updates)
                count = float(count) * ddist.world_size()
            else:
                count = ddist.allreduce_stats(sums, count)
                mean, rstd, scale, shift = ops.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum,
                                                           eps, n_updates)
        else:
            scale, shift = _bn_fold_cached(gamma, beta, running_mean, running_var, eps)
            if not _GRAD_OFF[0]:
                mean, rstd = running_mean.clone(), torch.rsqrt(running_var + eps)
            else:
                mean = rstd = None
        w2 = w_top.reshape(1, C).contiguous()
        ram = ops.ram_reduce(y, w2, b_top.contiguous(), scale, shift)
        ctx.training, ctx.count, ctx.saved_planes, ctx.wshape = training, count, xs, tuple(w_top.shape)
        ctx.save_for_backward(w, gamma, y, scale, shift, mean, rstd, w2)
        return ram

    @staticmethod
    def backward(ctx, g):
        w, gamma, y, scale, shift, mean, rstd, w2 = ctx.saved_tensors
        N, C, D, H, W = y.shape
        g = g.contiguous()                                   # [N,1,D,H,W]: one float per voxel row
        sums = ops.bn_relu_bwd_reduce(g, y, scale, shift, mean, rstd, wtop=w2)
        dbeta, dgamma = sums[:C].float(), sums[C:2 * C].float()
        dw_top, db_top = sums[2 * C:3 * C].float().view(ctx.wshape), sums[3 * C:3 * C + 1].float()
        if ctx.training:
            dys = ops.bn_relu_bwd_apply_planes(g, 0, w2, y, scale, shift, mean, rstd, gamma,
                                               ddist.allreduce_sums(sums[:2 * C].contiguous()), ctx.count)
        else:
            dys = ops.bn_relu_bwd_apply_planes(g, 0, w2, y, scale, shift, None, None, gamma, None, 1.0)
        dx = conv_dgrad(None, w, dys) if ctx.needs_input_grad[0] else None
        dw = conv_wgrad(ctx.saved_planes, None, w, dys) if ctx.needs_input_grad[3] else None
        return dx, None, None, dw, dgamma, dbeta, None, None, None, None, None, None, dw_top, db_top


def ram_fusable(w, w_top, bias):
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    return (planes_enabled() and bias is None and ops.umma_ok_fwd(Cin, Cout, k) and ops.umma_ok_fwd(Cout, Cin, k)
            and Cout % 8 == 0 and w_top.shape[0] == 1 and tuple(w_top.shape[2:]) == (1, 1, 1))


def conv_bn_relu_ram(x, w, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates, w_top, b_top):
    _GRAD_OFF[0] = not torch.is_grad_enabled()
    return ConvBnReluRam.apply(*_unwrap(x), w, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates,
                               w_top, b_top)


class UpsampleConcat(torch.autograd.Function):
    """nn.Upsample(x2, trilinear, align_corners=True) + crop_concat_5d([up, skip]) — parts.py:149-153.
    Planes in -> planes out when both inputs are planes; fp32 otherwise."""

    @staticmethod
    def forward(ctx, x, x_hi, x_lo, skip, s_hi, s_lo):
        ctx.shapes = (tuple(x.shape), tuple(skip.shape))
        ctx.set_materialize_grads(False)
        ctx.planes = x_hi is not None and s_hi is not None
        if ctx.planes:
            xs = ops.SplitPlanes(x_hi, x_lo, tuple(x.shape), x_hi.shape[-1])
            ss = ops.SplitPlanes(s_hi, s_lo, tuple(skip.shape), s_hi.shape[-1])
            if ops.virtual_concat_ok(xs, ss):
                # VIRTUAL concat: only the upsampled half is written; the skip planes are handed on as they are and the
                # convolution reads both through two tensor maps (the caller re-attaches them, see upsample_concat)
                cat = ops.upsample2x_virtual_concat(xs, ss)
            else:
                cat = ops.upsample2x_concat_planes(xs, ss)
            ctx.mark_non_differentiable(*[t for t in (cat.hi, cat.lo) if t is not None])
            return _handle(cat.shape, x_hi.device), cat.hi, cat.lo
        if x_hi is not None:
            x = ops.merge_planes(ops.SplitPlanes(x_hi, x_lo, tuple(x.shape), x_hi.shape[-1]))
        if s_hi is not None:
            skip = ops.merge_planes(ops.SplitPlanes(s_hi, s_lo, tuple(skip.shape), s_hi.shape[-1]))
        return ops.upsample2x_concat(ops.to_cl(x, "upsample input"), ops.to_cl(skip, "skip")), None, None

    @staticmethod
    def backward(ctx, g, _g1, _g2):
        (N, C1, d, h, w), (_, C2, Ds, Hs, Ws) = ctx.shapes
        if g is None:
            return (None,) * 6
        g = ops.to_cl(g, "grad")
        if (Ds, Hs, Ws) == (2 * d, 2 * h, 2 * w) and C1 % 4 == 0 and C2 % 4 == 0:
            # no crop: the skip gradient is the channel slice [C1:] of g, handed on as a view and read in place by the
            # BatchNorm backward kernels of the encoder unit (row pitch C1 + C2)
            dx, _ = ops.upsample2x_concat_bwd(g, *ctx.shapes, want_dskip=False)
            dskip = g[:, C1:]
        else:
            dx, dskip = ops.upsample2x_concat_bwd(g, *ctx.shapes)
        return dx, None, None, dskip, None, None


def upsample_concat(x, skip):
    t, hi, lo = UpsampleConcat.apply(*_unwrap(x), *_unwrap(skip))
    out = _wrap(t, hi, lo)
    if isinstance(out, Act) and isinstance(skip, Act) and hi.shape[-1] < ops._pad64(t.shape[1]):
        # the planes returned hold only the upsampled channels: virtual concat with the skip planes as second operand.  The
        # autograd Functions take (handle, hi, lo) triples, so the second operand travels as an attribute of the hi tensor.
        out.planes = ops.SplitPlanes(hi, lo, tuple(t.shape), hi.shape[-1] + skip.planes.Cpad, second=skip.planes)
        hi._dram_second = skip.planes
    return out


class TrilinearResize(torch.autograd.Function):
    """F.interpolate(size=..., mode='trilinear', align_corners=True) — models.py:146,514-518,588,591."""

    @staticmethod
    def forward(ctx, x, size):
        x = ops.to_cl(x, "resize input")
        ctx.src = tuple(x.shape[2:])
        size = tuple(int(s) for s in size)
        if size == ctx.src:
            ctx.identity = True
            return x.view_as(x)
        ctx.identity = False
        return ops.trilinear_resize(x, size)

    @staticmethod
    def backward(ctx, g):
        if ctx.identity:
            return g, None
        return ops.trilinear_resize_bwd(ops.to_cl(g, "grad"), ctx.src), None


class RamReduce(torch.autograd.Function):
    """top_layer = nn.Conv3d(C -> out_ch, k=1) — the regression-weight channel reduce, models.py:109-110,145."""

    @staticmethod
    def forward(ctx, feat, w, b):
        feat = ops.to_cl(feat, "RAM features")
        w2 = w.reshape(w.shape[0], w.shape[1]).contiguous()
        ctx.save_for_backward(feat, w2)
        ctx.wshape = tuple(w.shape)
        return ops.ram_reduce(feat, w2, b.contiguous())

    @staticmethod
    def backward(ctx, g):
        feat, w2 = ctx.saved_tensors
        dfeat, dw, db = ops.ram_reduce_bwd(ops.to_cl(g, "grad"), feat, w2)
        return dfeat, dw.view(ctx.wshape), db


class MaskedMean(torch.autograd.Function):
    """sum_v f(x)*m / sum_v m per (sample, channel); f = sigmoid optionally — models.py:45-47, metrics.py:160-165.
    x [B,C,D,H,W] with C == 1 or mask broadcast over C handled by the caller; returns (mean [B], count [B])."""

    @staticmethod
    def forward(ctx, x, mask, use_sigmoid, mode_gt0):
        B = x.shape[0]
        x2 = x.reshape(B, -1).contiguous()
        m2 = mask.reshape(B, -1).contiguous()
        out = ops.masked_pool(x2, m2, use_sigmoid, mode_gt0)
        ctx.save_for_backward(x2, m2, out)
        ctx.flags = (use_sigmoid, mode_gt0)
        ctx.xshape = tuple(x.shape)
        mean = (out[:, 0] / out[:, 1]).float()
        count = out[:, 1].float()
        ctx.mark_non_differentiable(count)
        return mean, count

    @staticmethod
    def backward(ctx, gmean, _gcount):
        x2, m2, out = ctx.saved_tensors
        g = (gmean.double() / out[:, 1]).float().contiguous()
        dx = ops.masked_pool_bwd(x2, m2, g, *ctx.flags)
        return dx.view(ctx.xshape), None, None, None


class IntRegHinge(torch.autograd.Function):
    """Interval-regression term of IntRegLoss (metrics.py:121-137, 158-177) on a batch: two masked-pool reductions (the
    prediction with the sigmoid fused, the lesion-candidate ratio), ONE tiny kernel for the interval / hinge / weighting of
    the B samples, and in the backward one small multiply + the masked-pool gradient kernel — instead of ~40 [B]-sized ATen
    launches around the same two reductions."""

    @staticmethod
    def forward(ctx, values, lobes, lesion_candidates, band, w, band_width, use_sigmoid):
        B = values.shape[0]
        x2 = values.reshape(B, -1).contiguous()
        m2 = lobes.reshape(B, -1).contiguous()
        l2 = lesion_candidates.reshape(B, -1).contiguous()
        pool_p = ops.masked_pool(x2, m2, use_sigmoid, True)         # mean of the probabilities over lobe > 0
        pool_r = ops.masked_pool(l2, m2, False, False)              # sum(lesion * lobe) / sum(lobe)
        loss, g = ops.int_reg_loss(pool_p, pool_r, band.contiguous(), w.contiguous(), band_width)
        ctx.save_for_backward(x2, m2, g)
        ctx.use_sigmoid = use_sigmoid
        ctx.xshape = tuple(values.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        x2, m2, g = ctx.saved_tensors
        dx = ops.masked_pool_bwd(x2, m2, (g * gout).contiguous(), ctx.use_sigmoid, True)
        return dx.view(ctx.xshape), None, None, None, None, None, None


class BootBce(torch.autograd.Function):
    """Segmentation term of IntRegRefineLoss — pseudo labels from the RAM (metrics.py:325-354) + BootBinCrossEntropy on
    the refined RAM (metrics.py:10-51) — as one reduction kernel and one gradient kernel.  Gradient flows to the refined
    logits only (the pseudo labels are built under no_grad in the reference)."""

    @staticmethod
    def forward(ctx, dense, refined, lobes, lesions, keep, smoothing, eps):
        B = refined.shape[0]
        flat = lambda t: t.reshape(B, -1).contiguous().float()
        d2, r2, lo2, le2 = flat(dense), flat(refined), flat(lobes), flat(lesions)
        keep = keep.contiguous().float()
        s = ops.boot_bce_sums(d2, r2, lo2, le2, keep, eps)
        counts = s[:3].clone()
        if ddist.active():                       # batch-global normalisers (metrics.py:30,37,42,48)
            ddist.all_reduce_(counts)
        n_out, n_in, t_in = counts[0], counts[1], counts[2]
        alpha = (1.0 - t_in / n_in.clamp_min(1.0)).clamp(0.25, 0.75).float().double()    # fp32 like the reference's alpha
        w_sum = (alpha * t_in + (1.0 - alpha) * (n_in - t_in)).clamp_min(1e-30)
        has = (n_in > 0).double()
        coef = torch.stack([1.0 / n_out, has * (1.0 - smoothing) * alpha / w_sum,
                            has * (1.0 - smoothing) * (1.0 - alpha) / w_sum, has * smoothing / n_in.clamp_min(1.0)])
        loss = (coef * s[3:7]).sum().float()
        ctx.save_for_backward(d2, r2, lo2, le2, keep, coef.float())
        ctx.eps, ctx.shape = eps, tuple(refined.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        d2, r2, lo2, le2, keep, coef = ctx.saved_tensors
        dref = ops.boot_bce_grad(d2, r2, lo2, le2, keep, (coef * g).contiguous(), ctx.eps)
        return None, dref.view(ctx.shape), None, None, None, None, None


class PcmAttend(torch.autograd.Function):
    """sum_o softmax_o(act(<theta f_x, phi f_{x+o}>)/T) * cam_{x+o} — the DGL update_all of models.py:322-411."""

    @staticmethod
    def forward(ctx, cam, f, tw, tb, pw, pb, connectivity, self_loop, flags, keep_for_backward=True):
        f = ops.to_cl(f, "attention features")
        cam = cam.contiguous()
        tw, tb, pw, pb = tw.contiguous(), tb.contiguous(), pw.contiguous(), pb.contiguous()
        need_grad = keep_for_backward and any(ctx.needs_input_grad[:6])   # grad mode is off inside forward: the caller tells us
        out, qk, stats = ops.pcm_fwd(f, cam, tw, tb, pw, pb, connectivity, self_loop, flags, keep_stats=need_grad)
        if need_grad:
            ctx.save_for_backward(f, cam, tw, pw, qk, stats)
        ctx.cfg = (connectivity, self_loop, flags)
        return out

    @staticmethod
    def backward(ctx, g):
        f, cam, tw, pw, qk, stats = ctx.saved_tensors
        dcam, df, dtw, dtb, dpw, dpb = ops.pcm_bwd(f, cam, tw, pw, qk, stats, g.contiguous(), *ctx.cfg)
        return dcam, df, dtw, dtb, dpw, dpb, None, None, None, None
