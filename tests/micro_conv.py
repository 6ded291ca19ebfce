"""Per-layer timing of the tcgen05 conv kernels at the training configuration (B=8, 80^3 chunks): forward, dgrad, wgrad.
Not a test: run on the GPU box, `python tests/micro_conv.py [B]`.  TF/s columns are ALGORITHMIC (1 pass); the split-bf16
kernels issue 3x that in tensor work."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
from dram_native import ops  # noqa: E402

LAYERS = [("ds0.c1", 32, 64, 80), ("ds1.c0", 64, 64, 40), ("ds1.c1", 64, 128, 40), ("ds2.c0", 128, 128, 20),
          ("ds2.c1", 128, 256, 20), ("bg.c0", 256, 256, 10), ("bg.c1", 256, 512, 10), ("us0.c0", 768, 256, 20),
          ("us0.c1", 256, 256, 20), ("us1.c0", 384, 128, 40), ("us1.c1", 128, 128, 40), ("us2.c0", 192, 64, 80),
          ("us2.c1", 64, 64, 80)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    only = sys.argv[2] if len(sys.argv) > 2 else None
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = 0.0
    print(f"{'layer':8s} {'Cin->Cout@D':>14s} {'GF':>8s} | {'fwd ms':>8s} {'TF/s':>6s} | {'dgrad ms':>8s} {'TF/s':>6s} | {'wgrad ms':>8s} {'TF/s':>6s}")
    for name, ci, co, d in LAYERS:
        if only and only != name:
            continue
        x = ops.new_volume(B, ci, d, d, d, "cuda").normal_()
        dy = ops.new_volume(B, co, d, d, d, "cuda").normal_()
        w = torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.05
        xs, dys = ops.split_bf16(x), ops.split_bf16(dy)
        wf = ops.pack_weight_bf16(w, 0)
        wd = ops.pack_weight_bf16(w, 1)
        gf = 2.0 * B * d ** 3 * ci * co * 27 / 1e9
        if os.environ.get("MICRO_AB"):            # interleaved A/B of the two forward kernels on the same box state
            res = {}
            for rep in range(3):
                for v in ("0", "1"):
                    os.environ["DRAM_CONV_V3"] = v
                    res.setdefault(v, []).append((timed(lambda: ops.conv_umma(xs, wf[0], wf[1], co, 3)),
                                                  timed(lambda: ops.conv_umma(dys, wd[0], wd[1], ci, 3))))
            os.environ.pop("DRAM_CONV_V3")
            best = {v: (min(r[0] for r in res[v]), min(r[1] for r in res[v])) for v in res}
            print(f"{name:8s} fwd  tile-pair {best['0'][0]:7.3f} ms ({gf / best['0'][0]:4.0f})  channels-on-M {best['1'][0]:7.3f} ms ({gf / best['1'][0]:4.0f}) | "
                  f"dgrad tile-pair {best['0'][1]:7.3f} ({gf / best['0'][1]:4.0f})  channels-on-M {best['1'][1]:7.3f} ({gf / best['1'][1]:4.0f})")
            continue
        t_f = timed(lambda: ops.conv_umma(xs, wf[0], wf[1], co, 3))
        t_d = timed(lambda: ops.conv_umma(dys, wd[0], wd[1], ci, 3))
        t_w = timed(lambda: ops.conv_umma_wgrad(dys, xs, ci, co, 3))
        tot["fwd"] += t_f; tot["dgrad"] += t_d; tot["wgrad"] += t_w; totf += gf
        print(f"{name:8s} {f'{ci}->{co}@{d}':>14s} {gf:8.1f} | {t_f:8.3f} {gf / t_f:6.0f} | {t_d:8.3f} {gf / t_d:6.0f} | {t_w:8.3f} {gf / t_w:6.0f}")
        del x, dy, xs, dys
    if os.environ.get("MICRO_AB"):
        return
    print(f"{'total':8s} {'':>14s} {totf:8.1f} | {tot['fwd']:8.3f} {totf / tot['fwd']:6.0f} | {tot['dgrad']:8.3f} {totf / tot['dgrad']:6.0f} | {tot['wgrad']:8.3f} {totf / tot['wgrad']:6.0f}")


if __name__ == "__main__":
    main()
