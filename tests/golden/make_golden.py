"""Generate the golden vectors under tests/golden/ by executing the UNMODIFIED reference (via oracle/ref_shims.py).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference ships no tests or fixtures, so these files are the pinned ground truth that travels to
the GPU box: seeded inputs, the reference-initialised state_dict, reference outputs, losses, gradients and updated
BatchNorm running statistics for one training step and one eval forward of
  * DC3D            (reference settings st_dram_ref.py with channels / 8,  chunk 16^3, batch 2)
  * DC3DATGeneric   (reference settings st_dram_ref_att.py with channels / 16, chunk 16^3, attention grid 12^3, batch 2)
and a stand-alone PCM case executed through the fake-DGL degree-bucketed update_all.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")

import dram_oracle as O  # noqa: E402
import ref_shims  # noqa: E402


def tiny_cfg(settings_name, div, at_size=None):
    s = ref_shims.load_settings(settings_name)
    cfg = dict(s.MODEL)
    for k in ("in_ch_list", "base_ch_list", "end_ch_list"):
        cfg[k] = [max(1, c // div) for c in cfg[k]]
    cfg["in_ch_list"][0] = 1
    if at_size is not None:
        cfg["at_spatial_size"] = at_size
    return cfg, s


def run_case(settings_name, div, at_size, size, B, seed):
    ref = ref_shims.load_reference()
    cfg, s = tiny_cfg(settings_name, div, at_size)
    model = ref_shims.build_reference_model(cfg, seed=seed)
    images, lobes, lesions, ctsses = O.synthetic_batch(B, size, seed=seed + 1)
    out = {"cfg": cfg, "loss_cfg": {k: v for k, v in s.LOSS_FUNC.items() if k != "method"},
           "loss_factors": list(s.LOSS_FACTORS), "images": images, "lobes": lobes, "lesions": lesions, "ctsses": ctsses,
           "state_dict": {k: v.clone() for k, v in model.state_dict().items()}}
    host = ref_shims.LossHost()
    out["freq_map"] = dict(host.ctss_frequency_map)
    model.eval()
    with torch.no_grad():
        d, r = model(images, lobes)
        out["eval_dense"], out["eval_refined"] = d.clone(), r.clone()
        out["eval_pool"] = model.pooling_dense_features(d, lobes).clone()
    model.train()
    loss = ref.metrics.IntRegRefineLoss(**out["loss_cfg"])
    with ref_shims.cpu_cuda_identity():
        rl, sl = loss(model, images, lobes, lesions, ctsses, obj=host, metas=ref_shims.trace_metas(B, size))
    total = rl * s.LOSS_FACTORS[0] + sl * s.LOSS_FACTORS[1]
    total.backward()
    out["train_reg_loss"], out["train_seg_loss"] = rl.detach().clone(), sl.detach().clone()
    out["grads"] = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    out["state_dict_after"] = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
    return out


def run_pcm(seed=3):
    ref = ref_shims.load_reference()
    torch.manual_seed(seed)
    grid, B, Cf = (6, 5, 7), 2, 17
    pcm = ref.models.PCM(grid, Cf, 1, 8, 0, 8, 1, 3, "scaled_dot_product_relu", False, p_enc_dim=0)
    cam = torch.randn(B, 1, *grid, requires_grad=True)
    f = torch.randn(B, Cf, *grid, requires_grad=True)
    out = pcm(cam, f)
    g = torch.randn_like(out)
    out.backward(g)
    return {"grid": grid, "state_dict": {k: v.clone() for k, v in pcm.state_dict().items()}, "cam": cam.detach().clone(),
            "f": f.detach().clone(), "out": out.detach().clone(), "gout": g, "dcam": cam.grad.clone(), "df": f.grad.clone(),
            "grads": {k: p.grad.clone() for k, p in pcm.named_parameters()},
            "n_edges": pcm.graph.number_of_edges(), "n_nodes": pcm.graph.number_of_nodes()}


if __name__ == "__main__":
    torch.save(run_case("st_dram_ref.py", 8, None, (16, 16, 16), 2, seed=0), os.path.join(HERE, "dc3d_div8_16.pt"))
    torch.save(run_case("st_dram_ref_att.py", 16, (12, 12, 12), (16, 16, 16), 2, seed=10), os.path.join(HERE, "dc3dat_div16_16.pt"))
    torch.save(run_pcm(), os.path.join(HERE, "pcm_6x5x7.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
