#!/usr/bin/env python
"""Benchmark of the DRAM hot path on B200 — one JSON line per run (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1]): DRAM training step — DC3D forward + IntRegRefineLoss (interval regression + seg
term) + backward + Adam step — on synthetic lobe chunks, 80^3, per-GPU batch 8, random-init (HeNorm) weights.
  value : lobe-chunks/s, whole job, inputs already resident in HBM, timed on the device (CUDA events, max over ranks)
  e2e   : the same metric through the public runner API (job_runner.LesionSegChunkTrain.train_step) with PINNED HOST
          batches: the host->device copy of the step's inputs and the device->host read of the loss are inside the
          timed region
  roofline     : dominant kernel, algorithmic FLOPs / CUDA-event time of its launches inside the timed region
  cpu_baseline : the CPU oracle (a port of the reference's arithmetic; oracle/dram_oracle.py) on the box's host cores,
                 bounded sample, rank 0 at N=1 only
`--impl reference` times that CPU arm alone (the reference has no other implementation of this path: it is PyTorch
library calls; /root/reference does not exist on the GPU box, so the oracle port stands in — kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))

CHUNK = (80, 80, 80)
FWD_GFLOP_PER_CHUNK = 924.61           # SURVEY §8d, 2*MAC over the 14 convs + 1x1x1 head
TRAIN_GFLOP_PER_CHUNK = 2772.96        # fwd + dgrad + wgrad, no dgrad for layer 0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if mx and c > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batch(B, seed, pinned):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))
    from train import synthetic_loader
    batch = next(iter(synthetic_loader(1, B, CHUNK, seed=seed)))
    if pinned:
        for k in ("#image", "#lobe_reference", "#pseudo_lesion_reference"):
            batch[k] = batch[k].pin_memory()
    return batch


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_chunks_per_s(steps, warmup, batch=1, threads=None):
    """The CPU oracle (port of the reference arithmetic) running the same training step: forward, IntRegRefineLoss,
    backward, Adam — torch CPU fp32 with all host threads."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dram_oracle as O
    import models
    from utils import Settings
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref.py"))
    cfg = dict(s.MODEL)
    cfg.pop("method")
    torch.manual_seed(0)
    m = models.DC3D(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    opt = torch.optim.Adam(params, lr=1e-3)
    images, lobes, lesions, ctsses = O.synthetic_batch(batch, CHUNK, seed=0)
    freq = {k: 1.0 / 6 for k in range(6)}

    def step():
        opt.zero_grad()
        d, r = O.dc3d_forward(sd, images, cfg, True)
        rl, sl = O.int_reg_refine_loss(d, r, lobes, lesions, ctsses, freq)
        (2.0 * rl + sl).backward()
        opt.step()
        return float(rl) + float(sl)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, s_per_step, threads = cpu_train_chunks_per_s(args.steps, args.warmup, batch=1)
    sample = f"{args.steps} timed training steps of batch 1 (80^3 chunk) after {args.warmup} warm-up, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": "train_lobe_chunks_per_s", "value": v, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DRAM training step (DC3D fwd + IntRegRefineLoss + bwd + Adam), 80^3 lobe chunks, "
                               "CPU arm runs batch 1 per step", "chunk": list(CHUNK)},
        "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dram_native import lib
    import job_runner
    from utils import Settings

    B = args.batch
    settings = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref.py"))
    settings.OPTIMIZER['lr'] = 1e-3                 # train.py default (train.py:33-34)
    settings.TRAIN_BATCH_SIZE = B
    torch.manual_seed(0)
    runner = job_runner.LesionSegChunkTrain(settings_module=settings)
    dev = torch.device("cuda", local)

    host_batch = make_batch(B, seed=rank, pinned=True)
    dev_batch = {k: (v.to(dev) if hasattr(v, "to") else v) for k, v in host_batch.items()}

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (value)
    for _ in range(args.warmup):
        runner.train_step(dev_batch)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.PROFILE.reset()
    lib.PROFILE.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = runner.train_step(dev_batch)
    e1.record()
    barrier()
    lib.PROFILE.enabled = False
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    ms_total = ms.item()
    launches = lib.PROFILE.launches
    prof = lib.PROFILE.summary()
    final_loss = loss.item()

    # ---- end-to-end leg: pinned host batch -> runner.train_step -> loss.item()
    nbytes_in = sum(host_batch[k].numel() * host_batch[k].element_size()
                    for k in ("#image", "#lobe_reference", "#pseudo_lesion_reference"))
    runner.train_step(host_batch)[0].item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        l, _ = runner.train_step(host_batch)
        l.item()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_value = world * B * args.steps / e2e_s.item()

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peaks = load_peaks()
    value = world * B * args.steps / (ms_total / 1e3)
    kern = {k: v for k, v in prof.items()}
    conv_names = ("dram_conv3d_umma_fwd", "dram_conv3d_umma_wgrad")
    dom = max(kern, key=lambda k: kern[k]["ms"])
    roof_k = dom if dom in conv_names else max(conv_names, key=lambda k: kern.get(k, {"ms": 0})["ms"])
    rk = kern.get(roof_k, {"ms": 0.0, "flops": 0.0, "calls": 0})
    achieved = rk["flops"] / (rk["ms"] / 1e3) / 1e12 if rk["ms"] > 0 else 0.0
    total_kernel_ms = sum(v["ms"] for v in kern.values())
    breakdown = {k.replace("dram_", ""): {"calls_per_step": v["calls"] / args.steps, "ms_per_step": v["ms"] / args.steps,
                                            "share": v["ms"] / total_kernel_ms,
                                            **({"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12} if v["flops"] else {})}
                 for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])[:12]}

    cpu = None
    if world == 1 and not args.no_cpu:
        v, s_per, threads = cpu_train_chunks_per_s(1, 0, batch=1)
        cpu = {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
               "sample": f"1 training step, batch 1, 80^3 chunk ({s_per:.1f} s), torch CPU fp32 oracle"}

    out = {
        "metric": "train_lobe_chunks_per_s", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3 (split-bf16 tensor-core operands, fp32 accumulate; fp32 elsewhere)",
        "data": "synthetic",
        "config": {"workload": "DRAM training step: DC3D fwd + IntRegRefineLoss + bwd + Adam, synthetic 80^3 lobe chunks, "
                               f"per-GPU batch {B} (BASELINE configs[1])",
                   "per_gpu_batch": B, "global_batch": B * world, "chunk": list(CHUNK), "parallelism": f"dp{world}",
                   "precision_mode": os.environ.get("DRAM_PRECISION", "bf16x3"),
                   "l2": "no flush needed: ~13 GB of activations stream through the 126 MB L2 every step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": {"dram_conv3d_umma_fwd": "k_conv_umma_fwd (forward + dgrad launches)",
                                                    "dram_conv3d_umma_wgrad": "k_conv_umma_wgrad"}.get(roof_k, roof_k),
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     "note": "achieved = algorithmic FLOPs (unpadded channels, one pass; the split-bf16 kernel issues 3 MMAs "
                             "per algorithmic MAC) / CUDA-event time of this kernel's launches in the timed region",
                     "whole_step_tflops_per_gpu": B * TRAIN_GFLOP_PER_CHUNK / 1e3 / (ms_total / args.steps / 1e3)},
        "kernels": breakdown,
        "cpu_baseline": cpu,
        "loss": final_loss,
    }
    print(json.dumps(out))
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch of lobe chunks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
