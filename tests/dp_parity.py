"""Data-parallel parity on real GPUs (run under torchrun, not collected by pytest):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_parity.py
N ranks each take a shard of one global batch; losses (summed over ranks), parameter gradients (after the bucketed
all-reduce) and BatchNorm running statistics must equal a single-process run on the GLOBAL batch."""
import os
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bodyct-dram_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))


class Host:
    ctss_frequency_map, debug_path, epoch_n = {k: 1.0 / 6 for k in range(6)}, "/tmp/x", 0


def step(model, batch, reducer=None):
    import metrics
    images, lobes, lesions, ctsses = batch
    rl, sl = metrics.IntRegRefineLoss()(model, images.cuda(), lobes.cuda(), lesions.cuda(), ctsses, obj=Host(), metas={})
    (2.0 * rl + sl).backward()
    if reducer is not None:
        reducer.finish()
    return rl.detach(), sl.detach()


def main():
    import dram_oracle as O
    import models
    from dram_native import dist as ddist
    from util import rel_err
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    td.init_process_group("nccl")
    if os.environ.get("DRAM_PEER", "1") == "1":
        assert ddist.init_peer() is not None, "peer mailboxes (CUDA IPC over NVLink) could not be set up"
    att = len(sys.argv) > 1 and sys.argv[1] == "att"
    cfg = dict(n_layers=3, in_ch_list=[1, 64, 128, 256, 768, 384, 192], base_ch_list=[32, 64, 128, 256, 256, 128, 64],
               end_ch_list=[64, 128, 256, 512, 256, 128, 64], kernel_sizes=[(3, 3)] * 7, stacking=3,
               padding_list=[(1, 1)] * 7, checkpoint_layers=[0, 1, 0, 1, 0, 1, 0], dropout=0.0, upsample_ksize=(3, 3, 3),
               upsample_sf=(2, 2, 2), out_ch=1)
    if att:
        cfg.update(at_spatial_size=(24, 24, 24), at_f_dim=8, at_g_dim=8, at_g_iter=1, at_k_size=3,
                   at_merge_type="scaled_dot_product_relu", at_self_loop=False, at_layers=[-1, 0, 1], at_p_enc_dim=0,
                   at_geo_f_dim=0)
    cls = models.DC3DATGeneric if att else models.DC3D
    per = 2
    torch.manual_seed(3)
    base = cls(**cfg)
    base.init(models.HeNorm(mode="fan_in"))
    sd0 = {k: v.clone() for k, v in base.state_dict().items()}
    images, lobes, lesions, ctsses = O.synthetic_batch(per * world, (32, 32, 32), seed=4)
    sl_ = slice(rank * per, (rank + 1) * per)
    shard = (images[sl_], lobes[sl_], lesions[sl_], ctsses[sl_])

    m = cls(**cfg)
    m.load_state_dict(sd0)
    m = m.cuda().train()
    reducer = ddist.GradReducer(m.parameters(), bucket_mb=8.0)
    rl, sl = step(m, shard, reducer)
    tot = torch.stack([rl, sl])
    td.all_reduce(tot)
    ok = True
    if rank == 0:
        ddist.configure(sync_bn=False)                       # single-process reference run: no exchanges
        os.environ["DRAM_SINGLE"] = "1"
    td.barrier()
    if rank == 0:
        # emulate world_size 1 by making the collectives no-ops
        ddist.active = lambda: False
        ref = cls(**cfg)
        ref.load_state_dict(sd0)
        ref = ref.cuda().train()
        rl1, sl1 = step(ref, (images, lobes, lesions, ctsses))
        e_loss = max(abs(tot[0].item() - rl1.item()) / abs(rl1.item()), abs(tot[1].item() - sl1.item()) / abs(sl1.item()))
        errs = sorted((rel_err(p.grad, dict(ref.named_parameters())[k].grad), k) for k, p in m.named_parameters()
                      if dict(ref.named_parameters())[k].grad is not None and not (k.startswith("reshape.") and k.endswith(".0.bias")))
        e_stat = max(rel_err(v, ref.state_dict()[k]) for k, v in m.state_dict().items() if "running" in k)
        print(f"DP parity world={world} att={att}: loss rel err {e_loss:.2e}; grad err median {errs[len(errs)//2][0]:.2e} "
              f"worst {errs[-1][0]:.2e} ({errs[-1][1]}); running-stat err {e_stat:.2e}")
        ok = e_loss < 1e-4 and errs[-1][0] < 2e-2 and errs[len(errs) // 2][0] < 5e-3 and e_stat < 1e-4
        print("DP PARITY", "OK" if ok else "FAILED")
    td.barrier()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
