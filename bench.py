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


_REAL_STDOUT = None


def emit(line):
    """the run's ONE stdout line"""
    out = _REAL_STDOUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line), sampled in-process through
    NVML every 25 ms (an external `nvidia-smi -lms` poller measurably stalls kernel launches on these hosts)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.mode = os.environ.get("DRAM_BENCH_SAMPLER", "nvml")

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                self._stop.wait(0.025)
        except Exception as e:                                  # noqa: BLE001
            self.reasons.add(f"nvml unavailable: {type(e).__name__}")

    def start(self):
        if self.mode == "none":
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler disabled"]}
        self._stop.set()
        self.thread.join(timeout=2)
        busy = [c for c in self.samples if self.max_mhz and c > 0.3 * self.max_mhz] or self.samples
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic_per_launch(kernel_prefixes):
    """Mean DRAM bytes (read + write) per launch of the given kernels, from the committed ncu launch list of this same
    command (profiles/*_launches.json, written by profiles/summarize_launches.py); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_launches.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    n = b = 0
    for k, v in d["kernels"].items():
        if any(k.startswith(pfx) for pfx in kernel_prefixes):
            n += v["launches"]
            b += v["dram_bytes_per_launch"] * v["launches"]
    return (b / n if n else None), os.path.basename(files[-1])


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


TRAIN_WORKLOAD = ("DRAM training step: DC3D fwd + IntRegRefineLoss + bwd + Adam, synthetic 80^3 lobe chunks, "
                  "per-GPU batch {B} (BASELINE configs[1])")
INFER_WORKLOAD = ("DC3DATGeneric eval forward (U-Net + RAM head + PCM refinement) + per-lobe pooling on the 5 lobe chunks of "
                  "one scan, 80^3, batch 5 per GPU")


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: /root/reference does not exist on the GPU box), all
    host threads, on this arm's metric / unit / config; every step is a bounded sample of the workload (one chunk)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        v, s_per_step, threads = cpu_train_chunks_per_s(args.steps, args.warmup, batch=1)
        metric, unit, hib, ms = "train_lobe_chunks_per_s", "chunks/s", True, s_per_step * 1e3
        workload = TRAIN_WORKLOAD.format(B=args.batch)
        sample = f"{args.steps} timed training steps of batch 1 (one 80^3 chunk per step) after {args.warmup} warm-up, torch CPU fp32"
    elif args.workload == "infer":
        v, dt, threads = cpu_infer_chunks_per_s(True, 1)
        metric, unit, hib, ms, workload = "infer_lobe_chunks_per_s", "chunks/s", True, dt * 1e3, INFER_WORKLOAD
        sample = "1 eval forward of DC3DATGeneric + pooling, batch 1 (one 80^3 chunk), torch CPU fp32"
    else:
        _, dt, threads = cpu_infer_chunks_per_s(True, 1)
        v, metric, unit, hib, ms = 5.0 * dt, "seconds_per_ct_scan", "s/scan (wall time per scan of the whole job)", False, 5.0 * dt * 1e3
        workload = "process_pipeline full-CT inference, synthetic scan 400x512x512 @ (1.0,0.7,0.7) mm + 5-lobe mask, one scan per step per GPU"
        sample = ("model part only: one of the scan's 5 lobe chunks through the CPU oracle x 5; the reference additionally "
                  "spends CPU time in SimpleITK resampling and numpy masking")
    emit(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": hib,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "chunk": list(CHUNK), "sample": sample},
        "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    att = args.model == "att"                       # BASELINE configs[2]: DC3DATGeneric (PCM lobe-graph attention head)
    settings = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py" if att else "st_dram_ref.py"))
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

    # ---- device-resident leg (value): the runner captures the step in a CUDA graph after its eager warm-up steps
    for _ in range(max(args.warmup, 3)):
        runner.train_step(dev_batch)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.PROFILE.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = runner.train_step(dev_batch)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    ms_total = ms.item()
    graphed = runner._graph is not None
    launches = (runner.kernels_per_step * args.steps) if graphed else lib.PROFILE.launches
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

    # ---- per-kernel CUDA events: the same step run eagerly (a captured graph cannot carry event records per launch)
    os.environ["DRAM_CUDA_GRAPH"] = "0"
    runner.train_step(dev_batch)
    barrier()
    lib.PROFILE.reset()
    lib.PROFILE.enabled = True
    ev_steps = min(args.steps, 3)
    for _ in range(ev_steps):
        runner.train_step(dev_batch)
    barrier()
    lib.PROFILE.enabled = False
    prof = lib.PROFILE.summary()
    if os.environ.get("DRAM_BENCH_LAYERS") and rank == 0:
        for k, v in sorted(lib.PROFILE.summary(by_tag=True).items(), key=lambda kv: -kv[1]["ms"])[:45]:
            tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0.0
            print(f"# {k:46s} calls/step {v['calls'] / ev_steps:4.1f} ms/step {v['ms'] / ev_steps:8.3f} TFLOP/s {tf:7.1f}", file=sys.stderr)

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peaks = load_peaks()
    value = world * B * args.steps / (ms_total / 1e3)
    kern = {k: v for k, v in prof.items()} or {"none": {"ms": 0.0, "flops": 0.0, "calls": 0, "bytes": 0.0}}
    conv_names = ("dram_conv3d_umma_fwd", "dram_conv3d_umma_wgrad")
    dom = max(kern, key=lambda k: kern[k]["ms"])
    roof_k = dom if dom in conv_names else max(conv_names, key=lambda k: kern.get(k, {"ms": 0})["ms"])
    rk = kern.get(roof_k, {"ms": 0.0, "flops": 0.0, "calls": 0})
    achieved = rk["flops"] / (rk["ms"] / 1e3) / 1e12 if rk["ms"] > 0 else 0.0
    total_kernel_ms = sum(v["ms"] for v in kern.values()) or 1.0
    breakdown = {k.replace("dram_", ""): {"calls_per_step": v["calls"] / ev_steps, "ms_per_step": v["ms"] / ev_steps,
                                            "share": v["ms"] / total_kernel_ms,
                                            **({"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12} if v["flops"] else {})}
                 for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])[:12]}

    traffic, traffic_src = ncu_traffic_per_launch(
        ("dram::k_conv_umma_fwd",) if roof_k == "dram_conv3d_umma_fwd" else ("dram::k_conv_umma_wgrad",))
    extra = None
    if world == 1 and not args.no_extra:
        del runner
        torch.cuda.empty_cache()
        extra = {"hbm_kernels": hbm_kernel_rooflines(peaks)}
        if os.environ.get("DRAM_BWD_PRECISION", "bf16x3") != "bf16x2":
            # the same step with the gradient operand of dgrad / wgrad carried as ONE bf16 plane (ops.grad_planes_three):
            # reported next to the headline, never as it - the parity tolerances on gradients are stated for bf16x3
            os.environ["DRAM_BWD_PRECISION"], os.environ["DRAM_CUDA_GRAPH"] = "bf16x2", "1"
            torch.manual_seed(0)
            r2 = job_runner.LesionSegChunkTrain(settings_module=settings)
            ms2 = time_cuda(lambda: r2.train_step(dev_batch), args.steps, max(args.warmup, 3))
            extra["train_bwd_bf16x2"] = {"metric": "train_lobe_chunks_per_s", "value": B / (ms2 / 1e3), "ms_per_step": ms2,
                                         "note": "DRAM_BWD_PRECISION=bf16x2: dy as one bf16 plane x split weights / split "
                                                 "layer input (2 MMAs per MAC in dgrad and wgrad); forward unchanged"}
            os.environ["DRAM_BWD_PRECISION"] = "bf16x3"
            del r2
            torch.cuda.empty_cache()
        att = build_att_runner()
        m = att.model.eval()
        b5 = make_batch(5, seed=1, pinned=False)
        i5, l5 = b5["#image"].unsqueeze(1).cuda(), b5["#lobe_reference"].unsqueeze(1).cuda()

        def infer5():
            with torch.no_grad():
                _, r = m(i5, l5)
                return m.pooling_dense_features(r, l5)
        ms5 = time_cuda(infer5, 5, 3)
        extra["infer"] = {"metric": "infer_lobe_chunks_per_s", "value": 5 / (ms5 / 1e3), "ms_per_batch_of_5": ms5,
                          "workload": "DC3DATGeneric eval forward + pooling, 5 lobe chunks of one scan (see --workload infer)"}
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
        "config": {"workload": TRAIN_WORKLOAD.format(B=B) if not att else
                               "DRAM training step: DC3DATGeneric (PCM attention head) fwd + IntRegRefineLoss + bwd + Adam, "
                               f"synthetic 80^3 lobe chunks, per-GPU batch {B} (BASELINE configs[2])",
                   "per_gpu_batch": B, "global_batch": B * world, "chunk": list(CHUNK), "parallelism": f"dp{world}",
                   "precision_mode": os.environ.get("DRAM_PRECISION", "bf16x3"),
                   "backward_precision": os.environ.get("DRAM_BWD_PRECISION", "bf16x3"),
                   "cuda_graph": graphed,
                   "l2": "no flush needed: ~13 GB of activations stream through the 126 MB L2 every step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": {"dram_conv3d_umma_fwd": "k_conv_umma_fwd3 / k_conv_umma_fwd2 / k_conv_umma_fwd (forward + dgrad launches)",
                                                    "dram_conv3d_umma_wgrad": "k_conv_umma_wgrad(_w3)"}.get(roof_k, roof_k),
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                     "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over this kernel's "
                                     f"launches in profiles/{traffic_src})" if traffic else None,
                     "algorithmic_flops_per_launch": rk["flops"] / max(rk["calls"], 1),
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     "note": "achieved = algorithmic FLOPs (unpadded channels, one pass; the split-bf16 kernel issues 3 MMAs "
                             "per algorithmic MAC) / CUDA-event time of this kernel's launches, events taken on the launching "
                             "stream over %d eager repetitions of the timed step (the timed step itself is one CUDA-graph launch)" % ev_steps,
                     "whole_step_tflops_per_gpu": B * TRAIN_GFLOP_PER_CHUNK / 1e3 / (ms_total / args.steps / 1e3)},
        "kernels": breakdown,
        "cpu_baseline": cpu,
        "extra": extra,
        "loss": final_loss,
    }
    emit(json.dumps(out))
    if world > 1:
        td.destroy_process_group()



# ------------------------------------------------------------------------------------------------ inference workloads
def build_att_runner(head="sigmoid"):
    import torch
    import job_runner
    from utils import Settings
    settings = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    torch.manual_seed(0)
    return job_runner.LesionSegTest(None, None, None, settings, None, head=head)


def time_cuda(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def hbm_kernel_rooflines(peaks):
    """Stand-alone HBM rooflines of the RAM-head and PCM kernels: algorithmic bytes / CUDA-event time."""
    import torch
    from dram_native import ops
    import models
    out = {}
    B, C, V = 8, 64, 80 * 80 * 80
    feat = ops.new_volume(B, C, 80, 80, 80, "cuda")
    feat.normal_()
    w, b = torch.randn(1, C, device="cuda"), torch.randn(1, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=10):
        ts = []
        for _ in range(reps + 3):
            flush.zero_()                                   # evict L2 (126 MB) between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts[3:])

    ms = timed(lambda: ops.ram_reduce(feat, w, b))
    nbytes = 4.0 * B * V * (C + 1)
    out["ram_reduce_fwd"] = {"bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6, "unit": "GB/s",
                             "peak": peaks["hbm_gbs"], "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"]}
    ram = torch.randn(B, V, device="cuda")
    mask = (torch.rand(B, V, device="cuda") > 0.4).float()
    ms = timed(lambda: ops.masked_pool(ram, mask, True, True))
    nbytes = 8.0 * B * V
    out["masked_pool_fwd"] = {"bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6, "unit": "GB/s",
                              "peak": peaks["hbm_gbs"], "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"]}
    SD, SH, SW = 400, 358, 358
    labels = torch.ones((SD, SH, SW), dtype=torch.uint8, device="cuda")
    heat = torch.zeros((SD, SH, SW), device="cuda")
    crop = ((20, 260), (30, 230), (40, 220))
    vc = 240 * 200 * 180
    r1 = ops.ram_activation(torch.randn(1, 1, 80, 80, 80, device="cuda"), 1)[0, 0]   # sigmoid first, as LesionSegTest.run does
    ms = timed(lambda: ops.ram_upsample_label_scatter(r1, labels, 1, heat, crop, 0, 1.0))
    nbytes = 4.0 * V + 5.0 * vc
    out["ram_upsample_label_scatter"] = {"bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6,
                                         "unit": "GB/s", "peak": peaks["hbm_gbs"], "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"],
                                         "crop_voxels": vc}
    Bp, G = 5, 64
    f = ops.new_volume(Bp, 17, G, G, G, "cuda")
    f.normal_()
    cam = torch.randn(Bp, 1, G, G, G, device="cuda")
    pcm = models.PCM((G, G, G), 17, 1, 8, 0, 8, 1, 3, "scaled_dot_product_relu", False, p_enc_dim=0).cuda()
    with torch.no_grad():
        ms = timed(lambda: pcm(cam, f))
    nbytes = 76.0 * Bp * G ** 3
    out["pcm_fwd"] = {"bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6, "unit": "GB/s",
                      "peak": peaks["hbm_gbs"], "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"],
                      "note": "algorithmic 76 B/voxel (f 17x4 + cam 4 in, 4 out); this round's kernels also write/read the "
                              "theta|phi projections (64 B) and the softmax weights (72 B) kept for the backward"}
    return out


def cpu_infer_chunks_per_s(att=True, batch=1, threads=None):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dram_oracle as O
    import models
    from utils import Settings
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py" if att else "st_dram_ref.py"))
    cfg = dict(s.MODEL)
    cls = getattr(models, cfg.pop("method").split(".")[-1])
    torch.manual_seed(0)
    m = cls(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, lobes, _, _ = O.synthetic_batch(batch, CHUNK, seed=0)
    fwd = O.dc3dat_forward if att else O.dc3d_forward
    t0 = time.perf_counter()
    with torch.no_grad():
        d, r = fwd(sd, images, dict(s.MODEL), False)
        O.masked_pool(r, lobes)
    dt = time.perf_counter() - t0
    return batch / dt, dt, threads


def run_infer(args):
    """BASELINE configs[0]/[3]-[4] building block: DC3DATGeneric eval forward + per-lobe pooling on the 5 lobe chunks of
    a scan (batch 5), chunks/s."""
    import torch
    import torch.distributed as td
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dram_native import lib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    runner = build_att_runner()
    model = runner.model.eval()
    B = 5
    batch = make_batch(B, seed=rank, pinned=True)
    img_h, lobe_h = batch["#image"].unsqueeze(1), batch["#lobe_reference"].unsqueeze(1)
    img_h, lobe_h = img_h.contiguous().pin_memory(), lobe_h.contiguous().pin_memory()
    img_d, lobe_d = img_h.cuda(), lobe_h.cuda()

    def fwd_dev():
        with torch.no_grad():
            _, r = model(img_d, lobe_d)
            return model.pooling_dense_features(r, lobe_d)

    def fwd_host():
        with torch.no_grad():
            i, l = img_h.cuda(non_blocking=True), lobe_h.cuda(non_blocking=True)
            _, r = model(i, l)
            return model.pooling_dense_features(r, l).cpu()

    for _ in range(args.warmup):
        fwd_dev()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.PROFILE.reset()
    lib.PROFILE.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fwd_dev()
    e1.record()
    torch.cuda.synchronize()
    lib.PROFILE.enabled = False
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    launches, prof = lib.PROFILE.launches, lib.PROFILE.summary()
    fwd_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd_host()
    torch.cuda.synchronize()
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        td.all_reduce(e2e, op=td.ReduceOp.MAX)
    if rank == 0:
        peaks = load_peaks()
        k = prof.get("dram_conv3d_umma_fwd", {"ms": 0.0, "flops": 0.0})
        achieved = k["flops"] / (k["ms"] / 1e3) / 1e12 if k["ms"] else 0.0
        cpu = None
        if world == 1 and not args.no_cpu:
            v, dt, threads = cpu_infer_chunks_per_s(True, 1)
            cpu = {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                   "sample": f"1 eval forward of DC3DATGeneric + pooling, batch 1, 80^3 chunk ({dt:.1f} s), torch CPU fp32 oracle"}
        emit(json.dumps({
            "metric": "infer_lobe_chunks_per_s", "value": world * B * args.steps / (ms.item() / 1e3), "unit": "chunks/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item() / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 tensor-core operands, fp32 accumulate; fp32 elsewhere)", "data": "synthetic",
            "config": {"workload": INFER_WORKLOAD, "per_gpu_batch": B, "chunk": list(CHUNK),
                       "l2": "activations (~5 GB per batch) stream through L2"},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / e2e.item(), "unit": "chunks/s",
                    "h2d_bytes_per_step": 2 * img_h.numel() * 4, "d2h_bytes_per_step": B * 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "k_conv_umma_fwd", "achieved": achieved, "peak": peaks["tflops_sustained"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"], "traffic": None,
                         "peak_source": peaks["source"]},
            "kernels": {k.replace("dram_", ""): {"calls_per_step": v["calls"] / args.steps, "ms_per_step": v["ms"] / args.steps,
                                                 **({"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12} if v["flops"] else {})}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]},
            "cpu_baseline": cpu}))
    if world > 1:
        td.destroy_process_group()


def run_scan(args):
    """BASELINE configs[3]/[4]: process_pipeline full-CT inference on a synthetic 512x512x400 scan (spacing 0.7 mm in-plane,
    1.0 mm slices): resample to 1 mm, 5 lobe chunks, DC3DATGeneric, RAM -> heat map -> Otsu -> lesion masks, back-resample.
    N > 1: scans are sharded across GPUs, no collective."""
    import numpy as np
    import torch
    import torch.distributed as td
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dram_oracle as O
    import job_runner
    from dram_native import lib
    runner = build_att_runner()
    shape = (400, 512, 512) if not args.small else (100, 128, 128)
    spacing = [1.0, 0.7, 0.7]
    scan, lobe, _, _ = O.synthetic_scan(shape, spacing, seed=rank)
    scan_h, lobe_h = torch.from_numpy(scan).pin_memory(), torch.from_numpy(lobe).pin_memory()
    new_sp = [1.0, 1.0, 1.0]

    pinned_out = {}

    def one_scan(scan_in, lobe_in, download):
        s_t = runner.resample_to_working_grid(scan_in.cuda(non_blocking=True), spacing, "linear")
        l_t = runner.resample_to_working_grid(lobe_in.cuda(non_blocking=True), spacing, "nearest")
        out = runner.run_scan(s_t, l_t, new_sp, return_device=True)
        les = job_runner.ops_itk_back(out["lesion"], shape, new_sp, spacing, "nearest")
        post = job_runner.ops_itk_back(out["lesion_post"], shape, new_sp, spacing, "nearest")
        if download:                                    # masks land in pinned host buffers (allocated once)
            if "les" not in pinned_out:
                pinned_out["les"] = torch.empty(les.shape, dtype=les.dtype).pin_memory()
                pinned_out["post"] = torch.empty(post.shape, dtype=post.dtype).pin_memory()
            pinned_out["les"].copy_(les, non_blocking=True)
            pinned_out["post"].copy_(post, non_blocking=True)
            return pinned_out["les"], pinned_out["post"], float(out["ratio"].item())     # .item() syncs the stream
        return les, post, out["ratio"]

    scan_d, lobe_d = scan_h.cuda(), lobe_h.cuda()
    for _ in range(args.warmup):
        one_scan(scan_d, lobe_d, False)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.PROFILE.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_scan(scan_d, lobe_d, False)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.PROFILE.launches
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    lib.PROFILE.reset()                                  # per-kernel CUDA events of one more scan (outside the timed region)
    lib.PROFILE.enabled = True
    one_scan(scan_d, lobe_d, False)
    torch.cuda.synchronize()
    lib.PROFILE.enabled = False
    prof = lib.PROFILE.summary()
    # end to end through the public API: LesionSegTest.run_scans over PINNED HOST scans (upload of scan i+1 and download
    # of the masks of scan i-1 overlap the kernels of scan i); every scan's H2D and D2H is inside the timed region
    for _ in runner.run_scans([(scan_h, lobe_h, spacing)] * max(args.steps, 3)):     # warm-up: slots, pinned buffers, allocator pools
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_done = sum(1 for _ in runner.run_scans([(scan_h, lobe_h, spacing)] * args.steps))
    torch.cuda.synchronize()
    assert n_done == args.steps
    e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    one_scan(scan_h, lobe_h, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one_scan(scan_h, lobe_h, True)
    torch.cuda.synchronize()
    e2e_single = time.perf_counter() - t0                                # one scan alone: upload + kernels + download in series
    if world > 1:
        td.all_reduce(e2e, op=td.ReduceOp.MAX)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, dt, threads = cpu_infer_chunks_per_s(True, 1)
            cpu = {"value": 5.0 * dt, "unit": "s/scan", "cores": threads, "kind": "port",
                   "sample": f"model part only: one of the scan's 5 lobe chunks through the CPU oracle ({dt:.1f} s) x 5; the "
                             "reference additionally spends CPU time in SimpleITK resampling and numpy masking"}
        s_per_scan = ms.item() / 1e3 / args.steps
        emit(json.dumps({
            "metric": "seconds_per_ct_scan", "value": s_per_scan / world, "unit": "s/scan (wall time per scan of the whole job)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item() / args.steps,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 tensor-core operands, fp32 accumulate; fp32 elsewhere)", "data": "synthetic",
            "config": {"workload": "process_pipeline full-CT inference, synthetic scan %dx%dx%d @ (1.0,0.7,0.7) mm + 5-lobe mask, "
                                   "one scan per step per GPU" % shape, "scans_per_s": world / s_per_scan},
            "clocks": clocks,
            "e2e": {"value": e2e.item() / args.steps / world, "unit": "s/scan",
                    "h2d_bytes_per_step": int(scan_h.numel() * 2 + lobe_h.numel()), "d2h_bytes_per_step": int(2 * lobe_h.numel() + 4),
                    "api": "LesionSegTest.run_scans (pipelined copies)", "single_scan_latency_s": e2e_single},
            "gpu_launches": launches, "cpu_baseline": cpu,
            "kernels": {k.replace("dram_", ""): {"calls_per_scan": v["calls"], "ms_per_scan": v["ms"]}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:16]},
            "roofline": None}))
    if world > 1:
        td.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for the duration of the run (NCCL prints
    # its version banner and debug lines to fd 1 from native code) and the line is written to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch of lobe chunks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "scan"],
                    help="train = BASELINE configs[1] (default, the driver's line); infer = chunk inference; scan = full-CT pipeline")
    ap.add_argument("--model", default="dc3d", choices=["dc3d", "att"],
                    help="train workload: dc3d = BASELINE configs[1]; att = DC3DATGeneric with the PCM head (configs[2])")
    ap.add_argument("--small", action="store_true", help="scan workload: 128x128x100 scan (quick check)")
    ap.add_argument("--no-extra", action="store_true", help="train workload: skip the extra inference / HBM-roofline figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer":
        run_infer(args)
    elif args.workload == "scan":
        run_scan(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
