"""Microbenchmark of the PCM stencil-attention kernels and the 1x1x1 reshape-head kernels (not a pytest file).
   python tests/micro_pcm.py [B]      -> per-call CUDA-event times with an L2 flush between repetitions, against the
   measured HBM copy bandwidth (MEASURED_PEAKS.json) and the algorithmic bytes of SURVEY section 8d."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
from dram_native import ops  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    ts = []
    for _ in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts[3:])


def line(name, ms, nbytes):
    print(f"{name:44s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB algorithmic  {nbytes / ms / 1e6:8.0f} GB/s  {nbytes / ms / 1e6 / PEAK:6.3f} of HBM peak")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    G, Cf, F = 64, 17, 8
    V = G ** 3
    f = ops.new_volume(B, Cf, G, G, G, "cuda"); f.normal_()
    cam = torch.randn(B, 1, G, G, G, device="cuda")
    tw, tb = torch.randn(F, Cf, device="cuda") * 0.3, torch.randn(F, device="cuda") * 0.1
    pw, pb = torch.randn(F, Cf, device="cuda") * 0.3, torch.randn(F, device="cuda") * 0.1
    flags = ops.MERGE_FLAGS["scaled_dot_product_relu"]
    print(f"PCM, B = {B}, grid {G}^3, Cf = {Cf}, F = {F}, 18 neighbours")
    line("pcm_fwd (inference: project + attend)", timed(lambda: ops.pcm_fwd(f, cam, tw, tb, pw, pb, 2, False, flags, False)), 76.0 * B * V)
    line("pcm_fwd (training: + softmax stats)", timed(lambda: ops.pcm_fwd(f, cam, tw, tb, pw, pb, 2, False, flags, True)), 76.0 * B * V)
    out, qk, stats = ops.pcm_fwd(f, cam, tw, tb, pw, pb, 2, False, flags, True)
    g = torch.randn_like(out)
    # backward, algorithmic: read f, cam, dout (4*(Cf+2)), write df, dcam (4*(Cf+1))
    line("pcm_bwd (attention + params/df)", timed(lambda: ops.pcm_bwd(f, cam, tw, pw, qk, stats, g, 2, False, flags)), 4.0 * (2 * Cf + 3) * B * V)
    for C, S in ((64, 80), (128, 40)):
        x = ops.new_volume(B, C, S, S, S, "cuda"); x.normal_()
        xs = ops.split_bf16(x)
        del x
        w, b = torch.randn(8, C, 1, 1, 1, device="cuda"), torch.randn(8, device="cuda")
        rows = B * S ** 3
        line(f"pointwise8_planes_fwd {C}->8 @ {S}^3", timed(lambda: ops.pointwise8_planes(xs, w, b)), rows * (4.0 * C + 32))
        dy = ops.new_volume(B, 8, S, S, S, "cuda"); dy.normal_()
        line(f"pointwise8_planes_wgrad {C}->8 @ {S}^3", timed(lambda: ops.pointwise8_planes_wgrad(xs, dy)), rows * (4.0 * C + 32))
        del xs, dy


if __name__ == "__main__":
    main()
