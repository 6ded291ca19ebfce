"""A/B of the wgrad chunk schedules at the training configuration (B=8): contiguous slabs (DRAM_WGRAD_SPG=1, round 1) vs
slabs interleaved within a wave (default).  Not a test: `python tests/micro_wgrad.py [reps]` on the GPU box; under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:wgrad` with reps=1 it yields the DRAM bytes per launch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dram_native import ops  # noqa: E402
from micro_conv import LAYERS, timed  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    B = 8
    print(f"{'layer':8s} {'Cin->Cout@D':>14s} {'operand MB':>10s} | contiguous ms  TF/s | interleaved ms  TF/s | max |dw diff|")
    for name, ci, co, d in LAYERS:
        if d < 40:
            continue
        x = ops.new_volume(B, ci, d, d, d, "cuda").normal_()
        dy = ops.new_volume(B, co, d, d, d, "cuda").normal_()
        xs, dys = ops.split_bf16(x), ops.split_bf16(dy)
        gf = 2.0 * B * d ** 3 * ci * co * 27 / 1e9
        mb = 4.0 * B * d ** 3 * (xs.Cpad + dys.Cpad) / 1e6
        res, outs = {}, {}
        for rep in range(2 if reps > 1 else 1):
            for spg in ("1", ""):
                if spg:
                    os.environ["DRAM_WGRAD_SPG"] = spg
                else:
                    os.environ.pop("DRAM_WGRAD_SPG", None)
                fn = lambda: ops.conv_umma_wgrad(dys, xs, ci, co, 3)
                if reps > 1:
                    res.setdefault(spg, []).append(timed(fn, reps))
                outs[spg] = fn()
        torch.cuda.synchronize()
        diff = (outs["1"] - outs[""]).abs().max().item()
        if reps > 1:
            a, b = min(res["1"]), min(res[""])
            print(f"{name:8s} {f'{ci}->{co}@{d}':>14s} {mb:10.0f} | {a:10.3f} {gf / a:6.0f} | {b:11.3f} {gf / b:6.0f} | {diff:.2e}")
        else:
            print(f"{name:8s} {f'{ci}->{co}@{d}':>14s} {mb:10.0f} | one launch each (contiguous, interleaved) | {diff:.2e}")
        del x, dy, xs, dys


if __name__ == "__main__":
    main()
