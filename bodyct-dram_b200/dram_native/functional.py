"""torch.autograd.Function wrappers: each forward/backward is a short sequence of libdram_b200 kernel launches.

Granularity follows the reference's building blocks (parts.py): one Function per [Conv3d -> BatchNorm3d -> ReLU
(-> MaxPool3d)] unit, one for upsample+concat, one for the RAM reduce, etc.  Tensors crossing Function boundaries
are logical NCDHW / channels-last memory fp32.
"""
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
        dys = dys if dys is not None else ops.split_bf16(dy)
        w_hi, w_lo, _ = WEIGHTS.get(w, "bf16_dgrad", lambda: ops.pack_weight_bf16(w.detach(), 1))
        return ops.conv_umma(dys, w_hi, w_lo, Cin, k)
    pack = WEIGHTS.get(w, "f32_dgrad", lambda: ops.pack_weight_f32(w.detach(), 1))
    return ops.conv_simt(dy, pack, None, Cin, k)


def conv_wgrad(saved_in, dy, w, dys=None):
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
    if isinstance(saved_in, ops.SplitPlanes):
        dys = dys if dys is not None else ops.split_bf16(dy)
        return ops.conv_umma_wgrad(dys, saved_in, Cin, Cout, k)
    return ops.conv_simt_wgrad(saved_in, dy, k)


class ConvBnRelu(torch.autograd.Function):
    """[Conv3d(bias optional) -> BatchNorm3d -> ReLU (-> MaxPool3d(2,2,0))]  — parts.py:103-110,184-196."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, running_mean, running_var, training, momentum, eps, n_updates, pool):
        x = ops.to_cl(x, "conv input")
        y, saved_in = conv_forward(x, w, bias)
        N, C, D, H, W = y.shape
        count = N * D * H * W
        if training:
            sums = ops.bn_stats(y)
            count = ddist.allreduce_stats(sums, count)
            mean, rstd, scale, shift = ops.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps,
                                                       n_updates)
        else:
            scale, shift = ops.bn_fold_eval(gamma, beta, running_mean, running_var, eps)
            mean, rstd = running_mean.clone(), torch.rsqrt(running_var + eps)      # only for dgamma/dbeta in eval mode
        a, pooled = ops.bn_relu_apply(y, scale, shift, pool)
        ctx.training, ctx.pool, ctx.count, ctx.has_bias = training, pool, count, bias is not None
        ctx.saved_in = saved_in if isinstance(saved_in, ops.SplitPlanes) else None
        tensors = [w, gamma, y, scale, shift, a if pool else None, mean, rstd,
                   None if isinstance(saved_in, ops.SplitPlanes) else saved_in]
        ctx.save_for_backward(*tensors)
        if pool:
            return a, pooled
        return a

    @staticmethod
    def backward(ctx, *grads):
        w, gamma, y, scale, shift, a, mean, rstd, x_plain = ctx.saved_tensors
        saved_in = ctx.saved_in if ctx.saved_in is not None else x_plain
        ga = grads[0]
        if ctx.pool:
            gp = grads[1]
            if ga is None:
                da = torch.zeros_like(y)
            else:
                da = ops.to_cl(ga, "grad").clone() if gp is not None else ops.to_cl(ga, "grad")
            if gp is not None:
                ops.maxpool2_bwd(a, ops.to_cl(gp, "grad"), da)
        else:
            da = ops.to_cl(ga, "grad")
        N, C, D, H, W = y.shape
        if ctx.training:
            sums = ops.bn_relu_bwd_reduce(da, y, scale, shift, mean, rstd)
            dbeta = sums[:C].float()
            dgamma = sums[C:].float()
            gsums = ddist.allreduce_sums(sums)
            dy = ops.bn_relu_bwd_apply(da, y, scale, shift, mean, rstd, gamma, gsums, ctx.count)
        else:
            dy = ops.bn_relu_bwd_apply(da, y, scale, shift, None, None, gamma, None, 1.0)
            dgamma = dbeta = None
            if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
                sums = ops.bn_relu_bwd_reduce(da, y, scale, shift, mean, rstd)
                dbeta, dgamma = sums[:C].float(), sums[C:].float()
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        dys = None
        if isinstance(saved_in, ops.SplitPlanes) or (ctx.needs_input_grad[0] and ops.umma_ok_fwd(Cout, Cin, k)):
            dys = ops.split_bf16(dy)
        dx = conv_dgrad(dy, w, dys) if ctx.needs_input_grad[0] else None
        dw = conv_wgrad(saved_in, dy, w, dys) if ctx.needs_input_grad[1] else None
        dbias = dy.sum(dim=(0, 2, 3, 4)) if ctx.has_bias else None
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None


class UpsampleConcat(torch.autograd.Function):
    """nn.Upsample(x2, trilinear, align_corners=True) + crop_concat_5d([up, skip]) — parts.py:149-153."""

    @staticmethod
    def forward(ctx, x, skip):
        x, skip = ops.to_cl(x, "upsample input"), ops.to_cl(skip, "skip")
        ctx.shapes = (tuple(x.shape), tuple(skip.shape))
        return ops.upsample2x_concat(x, skip)

    @staticmethod
    def backward(ctx, g):
        dx, dskip = ops.upsample2x_concat_bwd(ops.to_cl(g, "grad"), *ctx.shapes)
        return dx, dskip


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


class PcmAttend(torch.autograd.Function):
    """sum_o softmax_o(act(<theta f_x, phi f_{x+o}>)/T) * cam_{x+o} — the DGL update_all of models.py:322-411."""

    @staticmethod
    def forward(ctx, cam, f, tw, tb, pw, pb, connectivity, self_loop, flags, keep_for_backward=True):
        f = ops.to_cl(f, "attention features")
        cam = cam.contiguous()
        tw, tb, pw, pb = tw.contiguous(), tb.contiguous(), pw.contiguous(), pb.contiguous()
        need_grad = keep_for_backward and any(ctx.needs_input_grad[:6])   # grad mode is off inside forward: the caller tells us
        out, qk, att = ops.pcm_fwd(f, cam, tw, tb, pw, pb, connectivity, self_loop, flags, keep_att=need_grad)
        if need_grad:
            ctx.save_for_backward(f, cam, tw, pw, qk, att)
        ctx.cfg = (connectivity, self_loop, flags)
        return out

    @staticmethod
    def backward(ctx, g):
        f, cam, tw, pw, qk, att = ctx.saved_tensors
        dcam, df, dtw, dtb, dpw, dpb = ops.pcm_bwd(f, cam, tw, pw, qk, att, g.contiguous(), *ctx.cfg)
        return dcam, df, dtw, dtb, dpw, dpb, None, None, None, None
