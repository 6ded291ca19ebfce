"""Stand-alone timing of the HBM-bound kernels of the training step at the bench configuration (B=8, 80^3 chunks):
algorithmic bytes (each tensor read / written once) / CUDA-event time against the measured copy bandwidth.
Not a test: `python tests/micro_elementwise.py [filter]` on the GPU box (also the command profiled with ncu)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
from dram_native import ops  # noqa: E402

PEAK = 6554.9
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps=int(os.environ.get('MICRO_REPS', '5'))):
    for _ in range(int(os.environ.get('MICRO_WARM', '2'))):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def report(name, ms, gbytes):
    print(f"{name:44s} {ms * 1000:9.1f} us  {gbytes:7.3f} GB  {gbytes / ms * 1000:7.0f} GB/s  {gbytes / ms * 1000 / PEAK:5.2f} of measured copy peak")


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    B = 8
    dev = "cuda"
    vol = lambda C, D: ops.new_volume(B, C, D, D, D, dev).normal_()
    vec = lambda C: (torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1)
    GB = lambda *els: sum(els) / 1e9
    cases = []

    def case(name, fn, gbytes):
        if flt in name:
            cases.append((name, fn, gbytes))

    for C, D in [(64, 80), (128, 40)]:
        y, da = vol(C, D), vol(C, D)
        V = B * D ** 3
        sc, sh = vec(C)
        mean, rstd = vec(C)[1], vec(C)[0]
        gamma = vec(C)[0]
        sums = torch.randn(2 * C, device=dev, dtype=torch.float64)
        gp = vol(C, D // 2)
        g1 = torch.randn(V, device=dev)
        wtop = torch.randn(C, device=dev)
        wide = vol(3 * C, D)
        t = f"{C}@{D}^3"
        case(f"bn_stats {t}", lambda y=y: ops.bn_stats(y), GB(4 * V * C))
        case(f"bn_relu_apply_planes {t}", lambda y=y, sc=sc, sh=sh: ops.bn_relu_apply_planes(y, sc, sh, False), GB(8 * V * C))
        case(f"bn_relu_apply_planes+pool {t}", lambda y=y, sc=sc, sh=sh: ops.bn_relu_apply_planes(y, sc, sh, True), GB(8.5 * V * C))
        case(f"bn_relu_bwd_reduce {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_relu_bwd_reduce(da, y, sc, sh, mean, rstd), GB(8 * V * C))
        case(f"bn_relu_bwd_reduce pitch {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_relu_bwd_reduce(wide[:, 2 * C:], y, sc, sh, mean, rstd, pitch=3 * C), GB(8 * V * C))
        case(f"bn_relu_bwd_reduce rank1 {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_relu_bwd_reduce(g1, y, sc, sh, mean, rstd, wtop=wtop), GB(4 * V * C))
        case(f"bn_relu_bwd_apply_planes {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_relu_bwd_apply_planes(da, 0, None, y, sc, sh, mean, rstd, gamma, sums, V), GB(12 * V * C))
        case(f"bn_relu_bwd_apply_planes rank1 {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_relu_bwd_apply_planes(g1, 0, wtop, y, sc, sh, mean, rstd, gamma, sums, V), GB(8 * V * C))
        case(f"bn_pool_bwd_reduce {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_pool_bwd_reduce(da, C, gp, y, sc, sh, mean, rstd), GB(8.5 * V * C))
        case(f"bn_pool_bwd_apply_planes {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.bn_pool_bwd_apply_planes(da, C, gp, y, sc, sh, mean, rstd, gamma, sums, V), GB(12.5 * V * C))
        case(f"ram_reduce_fused {t}", lambda da=da, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, gamma=gamma, sums=sums, gp=gp, g1=g1, wtop=wtop, wide=wide, C=C, V=V: ops.ram_reduce(y, wtop.view(1, C), torch.zeros(1, device=dev), sc, sh), GB(4 * V * C + 4 * V))
    for C1, C2, d in [(128, 64, 40), (256, 128, 20)]:
        xs = ops.split_bf16(ops.new_volume(B, C1, d, d, d, dev).normal_())
        ss = ops.split_bf16(ops.new_volume(B, C2, 2 * d, 2 * d, 2 * d, dev).normal_())
        g = ops.new_volume(B, C1 + C2, 2 * d, 2 * d, 2 * d, dev).normal_()
        v, V = B * d ** 3, B * (2 * d) ** 3
        t = f"{C1}@{d}^3+{C2}@{2 * d}^3"
        case(f"upsample2x_concat_planes {t}", lambda xs=xs, ss=ss: ops.upsample2x_concat_planes(xs, ss), GB(4 * v * C1 + 4 * V * C2 + 4 * V * (C1 + C2)))
        case(f"upsample2x_concat_bwd(dx) {t}", lambda g=g, d=d, C1=C1, C2=C2: ops.upsample2x_concat_bwd(g, (B, C1, d, d, d), (B, C2, 2 * d, 2 * d, 2 * d), want_dskip=False),
             GB(4 * V * C1 + 4 * v * C1))
    x1 = ops.new_volume(B, 1, 80, 80, 80, dev).normal_()
    w1 = torch.randn(32, 1, 3, 3, 3, device=dev)
    pack = ops.pack_weight_f32(w1, 0)
    dy1 = vol(32, 80)
    V = B * 80 ** 3
    case("conv_first_fwd 1->32@80^3", lambda: ops.conv_simt(x1, pack, None, 32, 3), GB(4 * V * 33))
    case("conv_first_wgrad 1->32@80^3", lambda: ops.conv_simt_wgrad(x1, dy1, 3), GB(4 * V * 33))
    xs = vol(64, 80)
    case("split_bf16 64@80^3", lambda: ops.split_bf16(xs), GB(8 * V * 64))
    for name, fn, gb in cases:
        report(name, timed(fn), gb)


if __name__ == "__main__":
    main()
