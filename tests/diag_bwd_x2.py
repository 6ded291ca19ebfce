"""Diagnostic (not a pytest file): per-parameter gradient error of the training step against the fp32 CPU oracle with the
three-product backward (default) and with the single-plane gradient operand (DRAM_BWD_PRECISION=bf16x2).
   python tests/diag_bwd_x2.py"""
import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bodyct-dram_b200","oracle","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import dram_oracle as O, metrics, models
from util import rel_err
class Host: ctss_frequency_map, debug_path, epoch_n = {k:1.0/6 for k in range(6)}, "/tmp/x", 0
def main(size,B,seed):
    cfg = dict(n_layers=3, in_ch_list=[1, 64, 128, 256, 768, 384, 192], base_ch_list=[32, 64, 128, 256, 256, 128, 64],
               end_ch_list=[64, 128, 256, 512, 256, 128, 64], kernel_sizes=[(3, 3)] * 7, stacking=3,
               padding_list=[(1, 1)] * 7, checkpoint_layers=[0, 1, 0, 1, 0, 1, 0], dropout=0.0, upsample_ksize=(3, 3, 3),
               upsample_sf=(2, 2, 2), out_ch=1)
    torch.manual_seed(seed)
    m = models.DC3D(**cfg); m.init(models.HeNorm(mode="fan_in"))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, lobes, lesions, ctsses = O.synthetic_batch(B, size, seed=seed+1)
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
    d_ref, r_ref = O.dc3d_forward(sd, images, cfg, True)
    rl_ref, sl_ref = O.int_reg_refine_loss(d_ref, r_ref, lobes, lesions, ctsses, Host.ctss_frequency_map)
    (2 * rl_ref + sl_ref).backward()
    for mode in ("bf16x3", "bf16x2"):
        os.environ["DRAM_BWD_PRECISION"] = mode
        mm = models.DC3D(**cfg); mm.load_state_dict(sd0); mm = mm.cuda().train()
        rl, sl = metrics.IntRegRefineLoss()(mm, images.cuda(), lobes.cuda(), lesions.cuda(), ctsses, obj=Host(), metas={})
        (2 * rl + sl).backward()
        errs = sorted(((rel_err(p.grad, sd[k].grad), k) for k, p in mm.named_parameters()), reverse=True)
        print(f"size {size} B {B} DRAM_BWD_PRECISION={mode}: worst", ", ".join(f"{k}:{e:.1e}" for e, k in errs[:4]), "| median %.2e" % errs[len(errs)//2][0], flush=True)
main((32,32,32),2,21)
main((48,48,48),2,5)
main((80,80,80),1,33)
