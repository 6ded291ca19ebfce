"""Diagnostic (not a test): per-parameter gradient error of the CUDA paths against the CPU oracle at full channel width."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bodyct-dram_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import dram_oracle as O  # noqa: E402
import metrics  # noqa: E402
import models  # noqa: E402
from util import rel_err  # noqa: E402


class Host:
    ctss_frequency_map, debug_path, epoch_n = {k: 1.0 / 6 for k in range(6)}, "/tmp/x", 0


def main(size, B):
    cfg = dict(n_layers=3, in_ch_list=[1, 64, 128, 256, 768, 384, 192], base_ch_list=[32, 64, 128, 256, 256, 128, 64],
               end_ch_list=[64, 128, 256, 512, 256, 128, 64], kernel_sizes=[(3, 3)] * 7, stacking=3,
               padding_list=[(1, 1)] * 7, checkpoint_layers=[0, 1, 0, 1, 0, 1, 0], dropout=0.0, upsample_ksize=(3, 3, 3),
               upsample_sf=(2, 2, 2), out_ch=1)
    torch.manual_seed(21)
    m = models.DC3D(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, lobes, lesions, ctsses = O.synthetic_batch(B, size, seed=22)
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    d_ref, r_ref = O.dc3d_forward(sd, images, cfg, True)
    rl_ref, sl_ref = O.int_reg_refine_loss(d_ref, r_ref, lobes, lesions, ctsses, Host.ctss_frequency_map)
    (2 * rl_ref + sl_ref).backward()
    for path, prec in (("simt", "bf16x3"), ("umma", "bf16x3"), ("umma", "bf16")):
        os.environ["DRAM_CONV_PATH"], os.environ["DRAM_PRECISION"] = path, prec
        mm = models.DC3D(**cfg)
        mm.load_state_dict(sd0)
        mm = mm.cuda().train()
        rl, sl = metrics.IntRegRefineLoss()(mm, images.cuda(), lobes.cuda(), lesions.cuda(), ctsses, obj=Host(), metas={})
        (2 * rl + sl).backward()
        errs = sorted(((rel_err(p.grad, sd[k].grad), k) for k, p in mm.named_parameters()), reverse=True)
        with torch.no_grad():
            dd, _ = mm(images.cuda(), lobes.cuda())
        print(f"size {size} B {B} path {path}/{prec}: reg {rl.item():.6f}/{rl_ref.item():.6f} seg {sl.item():.6f}/{sl_ref.item():.6f}"
              f" dense(train-mode 2nd fwd) vs oracle {rel_err(dd, d_ref):.2e}")
        print("   worst grads:", ", ".join(f"{k}:{e:.1e}" for e, k in errs[:6]))
        print("   median grad err: %.2e" % errs[len(errs) // 2][0])


if __name__ == "__main__":
    main((16, 16, 16), 2)
    main((32, 32, 32), 2)
