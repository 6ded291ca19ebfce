"""SM-pair wgrad (k_conv_umma_wgrad2) against the single-SM generic wgrad on the same inputs, then per-layer timings of both.
Not a test: `python tests/micro_wgrad_pairs.py` on the GPU box."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
from dram_native import ops  # noqa: E402
from micro_conv import timed  # noqa: E402


def run(dys, xs, ci, co, v2):
    os.environ["DRAM_WGRAD_V2"] = v2
    try:
        return ops.conv_umma_wgrad(dys, xs, ci, co, 3)
    finally:
        os.environ.pop("DRAM_WGRAD_V2")


def main():
    torch.manual_seed(0)
    for N, Cin, Cout, S in [(1, 64, 128, (6, 8, 24)), (2, 128, 256, (10, 10, 10)), (1, 128, 384, (2, 8, 16)), (1, 384, 128, (4, 16, 8)),
                            (2, 96, 128, (5, 6, 7)), (1, 64, 128, (40, 40, 40))]:
        x, dy = torch.randn(N, Cin, *S, requires_grad=True), torch.randn(N, Cout, *S)
        xs = ops.split_bf16(x.detach().cuda().contiguous(memory_format=torch.channels_last_3d), True)
        dys = ops.split_bf16(dy.cuda().contiguous(memory_format=torch.channels_last_3d), True)
        d1 = run(dys, xs, Cin, Cout, "1")
        d0 = run(dys, xs, Cin, Cout, "0")
        torch.cuda.synchronize()
        line = f"N{N} {Cin}->{Cout} {S}: pair vs single-SM max rel diff {(d1 - d0).abs().max().item() / d0.abs().max().item():.2e}"
        if S[0] * S[1] * S[2] <= 4096:
            w = torch.zeros(Cout, Cin, 3, 3, 3, requires_grad=True)
            F.conv3d(x, w, None, padding=1).backward(dy)
            line += f", vs fp32 torch {(d1.cpu() - w.grad).abs().max().item() / w.grad.abs().max().item():.2e}"
        print(line, flush=True)
    B = 8
    for name, ci, co, d in [("ds1.c1", 64, 128, 40), ("us1.c1", 128, 128, 40), ("us1.c0", 384, 128, 40), ("ds2.c0", 128, 128, 20),
                            ("ds2.c1", 128, 256, 20), ("us0.c0", 768, 256, 20), ("us0.c1", 256, 256, 20), ("bg.c0", 256, 256, 10),
                            ("bg.c1", 256, 512, 10)]:
        x = ops.new_volume(B, ci, d, d, d, "cuda").normal_()
        dy = ops.new_volume(B, co, d, d, d, "cuda").normal_()
        xs, dys = ops.split_bf16(x), ops.split_bf16(dy)
        gf = 2.0 * B * d ** 3 * ci * co * 27 / 1e9
        res = {}
        for rep in range(2):
            for v in ("0", "1"):
                res.setdefault(v, []).append(timed(lambda: run(dys, xs, ci, co, v)))
        t0, t1 = min(res["0"]), min(res["1"])
        print(f"{name:8s} {ci}->{co}@{d}: single-SM {t0:7.3f} ms ({gf / t0:4.0f} TF/s)   SM pair {t1:7.3f} ms ({gf / t1:4.0f} TF/s)", flush=True)
        del x, dy, xs, dys


if __name__ == "__main__":
    main()
