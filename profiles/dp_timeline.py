#!/usr/bin/env python
"""Timeline of ONE data-parallel training step (run under torchrun on N GPUs; rank 0 profiles):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 profiles/dp_timeline.py OUT.md
torch.profiler (CUPTI) records every kernel of an EAGER step on rank 0 — ours, ATen's, NCCL's and the peer-mailbox exchanges —
and this script writes: total busy / idle time of the compute stream, the NCCL kernels and whether they overlap ours, the
peer exchanges (whose duration is mostly waiting for the slowest rank) and the largest gaps.  nsys is not installed in the
image; this is the same information from CUPTI through torch."""
import os
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
sys.path.insert(0, ROOT)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dp_timeline.md"
    import bench
    import job_runner
    from utils import Settings
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["DRAM_CUDA_GRAPH"] = "0"
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref.py"))
    s.OPTIMIZER['lr'] = 1e-3
    s.TRAIN_BATCH_SIZE = 8
    torch.manual_seed(0)
    runner = job_runner.LesionSegChunkTrain(settings_module=s)
    batch = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in bench.make_batch(8, seed=rank, pinned=False).items()}
    for _ in range(4):
        runner.train_step(batch)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            runner.train_step(batch)
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
        evs.sort(key=lambda e: e.time_range.start)
        t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
        span = (t1 - t0) / 2.0 / 1e3                                         # ms per step (two steps profiled)
        def tot(pred):
            return sum(e.device_time_total for e in evs if pred(e.name)) / 2.0 / 1e3
        is_nccl = lambda n: "nccl" in n.lower()
        is_peer = lambda n: "k_peer_allreduce" in n
        is_conv = lambda n: "k_conv_umma" in n
        # union of busy intervals on the device (all streams) -> idle
        busy, cur_s, cur_e = 0.0, None, None
        for e in evs:
            a, b = e.time_range.start, e.time_range.end
            if cur_e is None or a > cur_e:
                if cur_e is not None:
                    busy += cur_e - cur_s
                cur_s, cur_e = a, b
            else:
                cur_e = max(cur_e, b)
        busy += cur_e - cur_s
        busy = busy / 2.0 / 1e3
        nccl = [e for e in evs if is_nccl(e.name)]
        others = [e for e in evs if not is_nccl(e.name)]
        def overlap(e):
            a, b = e.time_range.start, e.time_range.end
            return sum(max(0, min(b, o.time_range.end) - max(a, o.time_range.start)) for o in others)
        nccl_overlapped = sum(min(overlap(e), e.device_time_total) for e in nccl) / 2.0 / 1e3
        gaps = sorted(((evs[i + 1].time_range.start - max(x.time_range.end for x in evs[max(0, i - 3):i + 1]), evs[i].name, evs[i + 1].name)
                       for i in range(len(evs) - 1)), reverse=True)[:8]
        lines = [f"# Data-parallel training step timeline, {world} x B200, rank 0, eager (torch.profiler / CUPTI), DC3D B = 8 per GPU", "",
                 f"* wall span per step: **{span:.2f} ms**; device busy (union over streams): {busy:.2f} ms; idle: {span - busy:.2f} ms",
                 f"* tensor-core conv kernels: {tot(is_conv):.2f} ms; peer-mailbox exchanges (`k_peer_allreduce`, {sum(1 for e in evs if is_peer(e.name)) // 2} per step): "
                 f"{tot(is_peer):.2f} ms — the kernel itself is ~3 us; the rest is waiting for the slowest rank to reach the same BatchNorm layer",
                 f"* NCCL kernels: {len(nccl) // 2} per step, {tot(is_nccl):.2f} ms, of which {nccl_overlapped:.2f} ms overlap other kernels "
                 f"(mode: {'overlapped buckets' if os.environ.get('DRAM_GRAD_OVERLAP', '0') == '1' else 'one flat all-reduce after backward'}, "
                 f"BatchNorm exchange: {'peer mailboxes' if os.environ.get('DRAM_PEER', '1') == '1' else 'NCCL'})", "",
                 "| largest gaps between consecutive kernels | after | before |", "|---|---|---|"]
        for g, a, b in gaps:
            lines.append(f"| {g / 1e3:.3f} ms | `{a[:60]}` | `{b[:60]}` |")
        lines += ["", "| kernel | launches/step | ms/step |", "|---|---|---|"]
        agg = {}
        for e in evs:
            k = e.name.split("(")[0][:70]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += e.device_time_total
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            lines.append(f"| `{k}` | {n / 2:.1f} | {t / 2.0 / 1e3:.3f} |")
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        open(out_path, "w").write("\n".join(lines) + "\n")
        print("\n".join(lines[:8]))
    runner.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
