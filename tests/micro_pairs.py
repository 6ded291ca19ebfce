"""SM-pair kernel (k_conv_umma_fwd4) against the single-SM kernels on the same inputs, then per-layer timings of both.
Not a test: `python tests/micro_pairs.py` on the GPU box."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
from dram_native import ops  # noqa: E402
from micro_conv import timed  # noqa: E402


def run(xs, w_hi, w_lo, Cout, v4, **kw):
    os.environ["DRAM_CONV_V4"] = v4
    try:
        return ops.conv_umma(xs, w_hi, w_lo, Cout, 3, **kw)
    finally:
        os.environ.pop("DRAM_CONV_V4")


def main():
    torch.manual_seed(0)
    for N, Cin, Cout, S in [(1, 64, 64, (16, 16, 16)), (2, 32, 64, (8, 8, 16)), (1, 192, 64, (8, 8, 8)), (1, 64, 192, (8, 8, 8)),
                            (2, 64, 128, (8, 8, 16)), (2, 192, 64, (3, 16, 32)), (2, 64, 64, (16, 40, 40)), (1, 128, 384, (2, 8, 16)),
                            (1, 64, 64, (1, 8, 32)), (2, 64, 64, (80, 80, 80))]:
        x = torch.randn(N, Cin, *S)
        w = torch.randn(Cout, Cin, 3, 3, 3) * (2.0 / (Cin * 27)) ** 0.5
        xs = ops.split_bf16(x.cuda().contiguous(memory_format=torch.channels_last_3d), True)
        w_hi, w_lo, _ = ops.pack_weight_bf16(w.cuda(), 0, True)
        y1, s1 = run(xs, w_hi, w_lo, Cout, "1", want_stats=True)
        y0, s0 = run(xs, w_hi, w_lo, Cout, "0", want_stats=True)
        torch.cuda.synchronize()
        err = (y1 - y0).abs().max().item() / y0.abs().max().item()
        serr = ((s1 - s0).abs() / s0.abs().clamp_min(1e-3)).max().item() if s0 is not None and s1 is not None else float("nan")
        line = f"N{N} {Cin}->{Cout} {S}: pair vs single-SM max rel diff {err:.2e}, stats {serr:.2e}"
        if S[0] * S[1] * S[2] <= 16 ** 3 * 2:
            ref = F.conv3d(x, w, None, padding=1)
            line += f", vs fp32 torch {(y1.cpu() - ref).abs().max().item() / ref.abs().max().item():.2e}"
        print(line, flush=True)
        assert err < 1e-5
    B = 8
    for name, ci, co, d in [("us2.c0", 192, 64, 80), ("us2.c0 dgrad", 64, 192, 80), ("us2.c1", 64, 64, 80),
                            ("ds1.c0", 64, 64, 40), ("ds1.c1", 64, 128, 40), ("ds1.c1 dgrad", 128, 64, 40), ("us1.c1", 128, 128, 40),
                            ("us1.c0", 384, 128, 40), ("us1.c0 dgrad", 128, 384, 40)]:
        x = ops.new_volume(B, ci, d, d, d, "cuda").normal_()
        w = torch.randn(co, ci, 3, 3, 3, device="cuda") * 0.05
        xs = ops.split_bf16(x)
        w_hi, w_lo, _ = ops.pack_weight_bf16(w, 0)
        gf = 2.0 * B * d ** 3 * ci * co * 27 / 1e9
        res = {}
        for rep in range(2):
            for v in ("0", "1"):
                res.setdefault(v, []).append(timed(lambda: run(xs, w_hi, w_lo, co, v)))
        t0, t1 = min(res["0"]), min(res["1"])
        print(f"{name:14s} {ci}->{co}@{d}: single-SM {t0:7.3f} ms ({gf / t0:4.0f} TF/s)   SM pair {t1:7.3f} ms ({gf / t1:4.0f} TF/s)", flush=True)
        del x, xs


if __name__ == "__main__":
    main()
