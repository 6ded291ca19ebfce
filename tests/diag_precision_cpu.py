"""Diagnostic (CPU only, not a pytest file): which operand precision does each phase of the path need?

The tensor-core convolutions multiply rounded operands and accumulate in fp32.  This script replays that on the CPU oracle:
every 3x3x3 convolution with more than one input channel gets its operands rounded the way a candidate kernel mode would
(the first layer and the 1x1x1 heads stay fp32, as in the library) and the outputs are compared with the fp32 oracle.
   python tests/diag_precision_cpu.py [size] [batch]
Modes:  bf16    one product, 8-bit mantissas            (DRAM_PRECISION=bf16, "fast mode")
        fp16    one product, 11-bit mantissas           (tcgen05 kind::f16 at the bf16 rate; = TF32 operand precision)
        fp16x2  x as two fp16 planes x w as one         (two products)
        bf16x3  hi*hi + hi*lo + lo*hi                   (three products: the library's parity mode)
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bodyct-dram_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import dram_oracle as O  # noqa: E402
import models  # noqa: E402

_conv3d = F.conv3d
MODE = ["fp32"]


def _split(t, dt):
    hi = t.to(dt).float()
    return hi, (t - hi).to(dt).float()


def conv3d_rounded(x, w, b=None, *a, **k):
    if MODE[0] == "fp32" or w.shape[1] == 1 or w.shape[2] == 1:
        return _conv3d(x, w, b, *a, **k)
    if MODE[0] == "bf16":
        return _conv3d(x.bfloat16().float(), w.bfloat16().float(), b, *a, **k)
    if MODE[0] == "fp16":
        return _conv3d(x.half().float(), w.half().float(), b, *a, **k)
    if MODE[0] == "fp16x2":
        xh, xl = _split(x, torch.float16)
        wh = w.half().float()
        return _conv3d(xh, wh, b, *a, **k) + _conv3d(xl, wh, None, *a, **k)
    if MODE[0] == "bf16x3":
        xh, xl = _split(x, torch.bfloat16)
        wh, wl = _split(w, torch.bfloat16)
        return _conv3d(xh, wh, b, *a, **k) + _conv3d(xh, wl, None, *a, **k) + _conv3d(xl, wh, None, *a, **k)
    raise ValueError(MODE[0])


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = dict(n_layers=3, in_ch_list=[1, 64, 128, 256, 768, 384, 192], base_ch_list=[32, 64, 128, 256, 256, 128, 64],
               end_ch_list=[64, 128, 256, 512, 256, 128, 64], kernel_sizes=[(3, 3)] * 7, stacking=3,
               padding_list=[(1, 1)] * 7, checkpoint_layers=[0, 1, 0, 1, 0, 1, 0], dropout=0.0, upsample_ksize=(3, 3, 3),
               upsample_sf=(2, 2, 2), out_ch=1)
    torch.set_num_threads(os.cpu_count())
    F.conv3d = conv3d_rounded
    O.F.conv3d = conv3d_rounded
    print(f"DC3D, reference widths, HeNorm init, {B} x {size}^3 synthetic lobe chunks; errors against the fp32 oracle")
    print(f"{'BN mode':28s} {'operands':8s} {'RAM normwise':>13s} {'pooled score':>13s} {'Dice(sigmoid(RAM)>0.5)':>24s}")
    for bn_state in ("eval, running stats at init", "eval, trained-like stats", "train (batch stats)"):
        torch.manual_seed(0)
        m = models.DC3D(**cfg)
        m.init(models.HeNorm(mode="fan_in"))
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        images, lobes, _, _ = O.synthetic_batch(B, (size,) * 3, seed=1)
        training = bn_state.startswith("train")
        if bn_state == "eval, trained-like stats":           # a few training-mode passes move the running statistics
            MODE[0] = "fp32"
            with torch.no_grad():
                for s in range(30):
                    O.dc3d_forward(sd, O.synthetic_batch(B, (size,) * 3, seed=10 + s)[0], cfg, True)
        ref = None
        for mode in ("fp32", "bf16", "fp16", "fp16x2", "bf16x3"):
            MODE[0] = mode
            sdc = {k: v.clone() for k, v in sd.items()}
            with torch.no_grad():
                d, _ = O.dc3d_forward(sdc, images, cfg, training)
                pooled = O.masked_pool(d, lobes)
            if ref is None:
                ref = (d, pooled)
                continue
            e_ram = ((d - ref[0]).norm() / ref[0].norm()).item()
            e_pool = ((pooled - ref[1]).abs().max() / ref[1].abs().max()).item()
            a, b = torch.sigmoid(d) > 0.5, torch.sigmoid(ref[0]) > 0.5
            dice = ((2.0 * (a & b).sum() + 1e-5) / (a.sum() + b.sum() + 1e-5)).item()
            print(f"{bn_state:28s} {mode:8s} {e_ram:13.2e} {e_pool:13.2e} {dice:24.5f}", flush=True)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "grad"):
    main()


# ------------------------------------------------------------------------------------------------ gradients
# Second part (python tests/diag_precision_cpu.py grad [size] [batch]): per-parameter gradient error of the training step
# for the backward modes of the library and for TF32 operands everywhere - what the reference itself computes on a GPU
# with PyTorch's default torch.backends.cudnn.allow_tf32 = True (convolutions in TF32, 10-bit mantissas).
def _tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class RoundedConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, fwd, bwd):
        ctx.save_for_backward(x, w)
        ctx.cfg = (padding, bwd, b is not None)
        if fwd == "fp32":
            return _conv3d(x, w, b, padding=padding)
        if fwd == "tf32":
            return _conv3d(_tf32(x), _tf32(w), b, padding=padding)
        xh, xl = _split(x, torch.bfloat16)
        wh, wl = _split(w, torch.bfloat16)
        return _conv3d(xh, wh, b, padding=padding) + _conv3d(xh, wl, None, padding=padding) + _conv3d(xl, wh, None, padding=padding)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        padding, bwd, has_b = ctx.cfg
        gi = lambda g, ww: torch.nn.grad.conv3d_input(x.shape, ww, g, padding=padding)
        gw = lambda xx, g: torch.nn.grad.conv3d_weight(xx, w.shape, g, padding=padding)
        if bwd == "fp32":
            dx, dw = gi(dy, w), gw(x, dy)
        elif bwd == "tf32":
            dx, dw = gi(_tf32(dy), _tf32(w)), gw(_tf32(x), _tf32(dy))
        else:
            xh, xl = _split(x, torch.bfloat16)
            wh, wl = _split(w, torch.bfloat16)
            gh, gl = _split(dy, torch.bfloat16)
            if bwd == "bf16x3":
                dx = gi(gh, wh) + gi(gh, wl) + gi(gl, wh)
                dw = gw(xh, gh) + gw(xh, gl) + gw(xl, gh)
            else:                                            # bf16x2: the gradient operand as one bf16 plane
                dx = gi(gh, wh) + gi(gh, wl)
                dw = gw(xh, gh) + gw(xl, gh)
        db = dy.sum(dim=(0, 2, 3, 4)) if has_b else None
        return (dx if ctx.needs_input_grad[0] else None), dw, db, None, None, None


def grad_main():
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    cfg = dict(n_layers=3, in_ch_list=[1, 64, 128, 256, 768, 384, 192], base_ch_list=[32, 64, 128, 256, 256, 128, 64],
               end_ch_list=[64, 128, 256, 512, 256, 128, 64], kernel_sizes=[(3, 3)] * 7, stacking=3,
               padding_list=[(1, 1)] * 7, checkpoint_layers=[0, 1, 0, 1, 0, 1, 0], dropout=0.0, upsample_ksize=(3, 3, 3),
               upsample_sf=(2, 2, 2), out_ch=1)
    torch.set_num_threads(os.cpu_count())
    modes = [("fp32", "fp32")]

    def conv(x, w, b=None, *a, padding=0, **k):
        if w.shape[1] == 1 or w.shape[2] == 1:
            return _conv3d(x, w, b, *a, padding=padding, **k)
        return RoundedConv.apply(x, w, b, padding, *modes[0])

    O.F.conv3d = conv
    torch.manual_seed(0)
    m = models.DC3D(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, lobes, lesions, ctsses = O.synthetic_batch(B, (size,) * 3, seed=1)
    freq = {k: 1.0 / 6 for k in range(6)}
    print(f"DC3D training step (IntRegRefineLoss), {B} x {size}^3; per-parameter gradient error ||g - g_fp32|| / ||g_fp32||")
    ref = None
    for fwd, bwd, label in (("fp32", "fp32", "fp32"), ("bf16x3", "bf16x3", "library default: bf16x3 forward + bf16x3 backward"),
                            ("bf16x3", "bf16x2", "DRAM_BWD_PRECISION=bf16x2: bf16x3 forward + one-plane gradient"),
                            ("tf32", "tf32", "TF32 operands everywhere (the reference on a GPU, cudnn.allow_tf32)")):
        modes[0] = (fwd, bwd)
        sd = {k: v.clone() for k, v in sd0.items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        d, r = O.dc3d_forward(sd, images, cfg, True)
        rl, sl = O.int_reg_refine_loss(d, r, lobes, lesions, ctsses, freq)
        (2.0 * rl + sl).backward()
        g = {k: v.grad.detach().clone() for k, v in sd.items() if v.requires_grad and v.grad is not None}
        if ref is None:
            ref = g
            continue
        errs = sorted(((g[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)).item() for k in ref)
        print(f"{label:72s} median {errs[len(errs) // 2]:.2e}  90% {errs[int(len(errs) * 0.9)]:.2e}  worst {errs[-1]:.2e}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "grad":
    grad_main()
