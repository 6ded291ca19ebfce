#!/usr/bin/env python
"""Benchmark of the DRAM hot path on B200 — ONE JSON line per run (contract: task statement / DESIGN.md section 5).

The line's headline is BASELINE.json configs[1]: the DRAM training step — DC3D forward + IntRegRefineLoss (interval
regression + seg term) + backward + Adam — on synthetic lobe chunks, 80^3, per-GPU batch 8, HeNorm weights:
  value : lobe-chunks/s, whole job, inputs already resident in HBM, timed on the device (CUDA events, max over ranks)
  e2e   : the same metric through the public runner API (job_runner.LesionSegChunkTrain.train_step) with PINNED HOST
          batches: the host->device copy of the step's inputs and the device->host read of the loss are timed
  roofline     : dominant kernel family, algorithmic FLOPs / CUDA-event time of its launches
  cpu_baseline : the CPU oracle (a port of the reference's arithmetic; oracle/dram_oracle.py) on the box's host cores,
                 bounded sample, rank 0 at N=1 only
BASELINE.json's metric has three parts (train chunks/s, inference chunks/s, s per full CT scan) and five configs; the
driver runs only `bench.py --gpus N`, so the SAME line carries the other parts under `extra`, measured at every N:
  extra.train_att : configs[2], DC3DATGeneric (PCM attention head) training step, data parallel       (value + e2e)
  extra.infer     : DC3DATGeneric eval forward + per-lobe pooling on the 5 lobe chunks of a scan      (value + e2e)
  extra.scan      : configs[3]/[4], process_pipeline full-CT inference, scans sharded across the GPUs (value + e2e)
each with its own roofline and cpu_baseline at N=1.  At N=1 `extra` also holds the stand-alone HBM rooflines of the RAM
head / PCM kernels, the two-product backward variant and `cudnn_baseline`: stock PyTorch + cuDNN running the same step
on the same GPU (what the reference executes on a GPU, train.py:44-45) — informational, never a ratio's denominator.
`--workload infer|scan` / `--model att` print the corresponding part as a line of its own.
`--impl reference` times the CPU arm alone (the reference has no other implementation of this path: it is PyTorch
library calls; /root/reference does not exist on the GPU box, so the oracle port stands in — kind "port").
"""
import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "bodyct-dram_b200"))

CHUNK = (80, 80, 80)
FWD_GFLOP_PER_CHUNK = 924.61           # SURVEY §8d, 2*MAC over the 14 convs + 1x1x1 head
TRAIN_GFLOP_PER_CHUNK = 2772.96        # fwd + dgrad + wgrad, no dgrad for layer 0
DTYPE = "bf16x3 (split-bf16 tensor-core operands, fp32 accumulate; fp32 elsewhere)"
SCAN_SHAPE, SCAN_SPACING = (400, 512, 512), [1.0, 0.7, 0.7]

TRAIN_WORKLOAD = ("DRAM training step: DC3D fwd + IntRegRefineLoss + bwd + Adam, synthetic 80^3 lobe chunks, "
                  "per-GPU batch {B} (BASELINE configs[1])")
TRAIN_ATT_WORKLOAD = ("DRAM training step: DC3DATGeneric (PCM attention head) fwd + IntRegRefineLoss + bwd + Adam, "
                      "synthetic 80^3 lobe chunks, per-GPU batch {B} (BASELINE configs[2])")
INFER_WORKLOAD = ("DC3DATGeneric eval forward (U-Net + RAM head + PCM refinement) + per-lobe pooling on the 5 lobe chunks of "
                  "one scan, 80^3, batch 5 per GPU (BASELINE configs[0] on the GPU)")
SCAN_WORKLOAD = ("process_pipeline full-CT inference, synthetic scan %dx%dx%d @ (1.0,0.7,0.7) mm + 5-lobe mask -> lesion masks, "
                 "one scan per step per GPU, scans sharded across GPUs (BASELINE configs[3]/[4])")

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
            return self
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler disabled"]}
        self._stop.set()
        self.thread.join(timeout=2)
        busy = [c for c in self.samples if self.max_mhz and c > 0.3 * self.max_mhz] or self.samples
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def source_sha():
    """fingerprint of the kernel sources an ncu capture belongs to (profiles/summarize_launches.py stamps the same value)"""
    h = hashlib.sha256()
    for f in ("conv_umma.cu", "planes.cu", "elementwise.cu"):
        h.update(open(os.path.join(ROOT, "bodyct-dram_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic_per_launch(kernel_prefixes):
    """Mean DRAM bytes (read + write) per launch of the given kernels, from the newest committed ncu launch list of this same
    command (profiles/*_launches.json, written by profiles/summarize_launches.py) -> (bytes, file, stale).  ncu cannot run
    inside the timed run, so the figure is a capture's; `stale` says whether the kernel sources changed since it was taken."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_train_b8_launches.json")))
    if not files:
        return None, None, None
    d = json.load(open(files[-1]))
    n = b = 0
    for k, v in d["kernels"].items():
        if any(k.startswith(pfx) for pfx in kernel_prefixes):
            n += v["launches"]
            b += v["dram_bytes_per_launch"] * v["launches"]
    stale = d.get("source_sha") != source_sha()
    return (b / n if n else None), os.path.basename(files[-1]), stale


def make_batch(B, seed, pinned):
    from train import synthetic_loader
    batch = next(iter(synthetic_loader(1, B, CHUNK, seed=seed)))
    if pinned:
        for k in ("#image", "#lobe_reference", "#pseudo_lesion_reference"):
            batch[k] = batch[k].pin_memory()
    return batch


class Dist:
    """rank / world / barrier / max-over-ranks of a host or device time"""

    def __init__(self):
        import torch
        import torch.distributed as td
        self.td = td
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not td.is_initialized():
            td.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        if self.world > 1:
            self.td.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        import torch
        t = torch.tensor([float(x)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return t.item()

    def close(self):
        if self.world > 1 and self.td.is_initialized():
            self.td.destroy_process_group()


def timed_device(dist, fn, steps, clocks=False):
    """barrier + sync, CUDA events around exactly `steps` calls, barrier + sync, max over ranks -> (ms_total, clocks)"""
    import torch
    dist.barrier()
    sampler = ClockSampler(dist.local).start() if (clocks and dist.rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier()
    ck = sampler.stop() if sampler else None
    return dist.max(e0.elapsed_time(e1)), ck


def timed_host(dist, fn, steps):
    """wall clock around `steps` calls of an API that ends in a device->host read, barriers on both sides, max over ranks"""
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dist.barrier()
    return dist.max(time.perf_counter() - t0)


def kernel_breakdown(prof, per, top=12):
    total = sum(v["ms"] for v in prof.values()) or 1.0
    return {k.replace("dram_", ""): {"calls_per_step": v["calls"] / per, "ms_per_step": v["ms"] / per, "share": v["ms"] / total,
                                     **({"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12} if v["flops"] else {}),
                                     **({"gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9} if v["bytes"] and not v["flops"] else {})}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:top]}


def conv_roofline(prof, peaks, per_steps, note):
    """tensor-pipe roofline of the dominant convolution family from per-kernel CUDA events"""
    conv_names = ("dram_conv3d_umma_fwd", "dram_conv3d_umma_wgrad")
    roof_k = max(conv_names, key=lambda k: prof.get(k, {"ms": 0.0})["ms"])
    rk = prof.get(roof_k, {"ms": 0.0, "flops": 0.0, "calls": 0})
    achieved = rk["flops"] / (rk["ms"] / 1e3) / 1e12 if rk["ms"] > 0 else 0.0
    traffic, src, stale = ncu_traffic_per_launch(("dram::k_conv_umma_fwd",) if roof_k == "dram_conv3d_umma_fwd" else ("dram::k_conv_umma_wgrad",))
    return {"bound": "tensor",
            "kernel": {"dram_conv3d_umma_fwd": "k_conv_umma_fwd4 (SM pairs, cta_group::2) / k_conv_umma_fwd / _fwd2 / _fwd3 (forward + dgrad launches)",
                       "dram_conv3d_umma_wgrad": "k_conv_umma_wgrad_w3 / k_conv_umma_wgrad2 (SM pairs)"}[roof_k],
            "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
            "traffic": traffic,
            "traffic_source": (f"bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), mean over this kernel family's launches "
                               f"in profiles/{src}; capture {'OLDER than' if stale else 'of'} the current kernel sources "
                               f"(source_sha {source_sha()})") if traffic else None,
            "traffic_stale": stale,
            "algorithmic_flops_per_launch": rk["flops"] / max(rk["calls"], 1),
            "ms_per_step": rk["ms"] / per_steps, "launches_per_step": rk["calls"] / per_steps,
            "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
            "note": "achieved = algorithmic FLOPs (unpadded channels, one pass; the split-bf16 kernels issue 3 MMAs per algorithmic "
                    "MAC, so 1/3 of the tensor peak AT THE CLOCK THE KERNELS RUN AT is the ceiling; `peak` is the sustained cuBLAS figure measured at ~1.3 GHz, the pair kernels hold 1.5-1.6 GHz) / CUDA-event time of this family's launches on the launching stream; " + note}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dram_oracle as O
    return O


def _oracle_model(att, device="cpu"):
    import torch
    import models
    from utils import Settings
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py" if att else "st_dram_ref.py"))
    cfg = dict(s.MODEL)
    cls = getattr(models, cfg.pop("method").split(".")[-1])
    torch.manual_seed(0)
    m = cls(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    sd = {k: v.detach().clone().to(device) for k, v in m.state_dict().items()}
    return sd, cfg


def oracle_train_chunks_per_s(steps, warmup, batch=1, att=False, threads=None, device="cpu"):
    """The oracle (port of the reference arithmetic) running the training step — forward, IntRegRefineLoss, backward, Adam
    — as plain torch fp32 ops: on the CPU with all host threads (the CPU arm), or on `device="cuda"` = stock PyTorch +
    cuDNN on this GPU (extra.cudnn_baseline)."""
    import torch
    O = _oracle()
    if device == "cpu":
        threads = threads or os.cpu_count()
        torch.set_num_threads(threads)
    sd, cfg = _oracle_model(att, device)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    opt = torch.optim.Adam(params, lr=1e-3)
    images, lobes, lesions, ctsses = O.synthetic_batch(batch, CHUNK, seed=0)
    images, lobes, lesions = images.to(device), lobes.to(device), lesions.to(device)
    freq = {k: 1.0 / 6 for k in range(6)}
    fwd = O.dc3dat_forward if att else O.dc3d_forward

    def step():
        opt.zero_grad()
        d, r = fwd(sd, images, cfg, True)
        rl, sl = O.int_reg_refine_loss(d, r, lobes, lesions, ctsses, freq)
        (2.0 * rl + sl).backward()
        opt.step()
        return float(rl.detach()) + float(sl.detach())

    for _ in range(warmup):
        step()
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def cpu_infer_chunks_per_s(att=True, batch=1, threads=None):
    import torch
    O = _oracle()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd, cfg = _oracle_model(att)
    images, lobes, _, _ = O.synthetic_batch(batch, CHUNK, seed=0)
    fwd = O.dc3dat_forward if att else O.dc3d_forward
    t0 = time.perf_counter()
    with torch.no_grad():
        d, r = fwd(sd, images, cfg, False)
        O.masked_pool(r, lobes)
    dt = time.perf_counter() - t0
    return batch / dt, dt, threads


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: /root/reference does not exist on the GPU box), all
    host threads, on this arm's metric / unit / config; every step is a bounded SAMPLE of the workload: ONE 80^3 chunk."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        att = args.model == "att"
        v, s_per_step, threads = oracle_train_chunks_per_s(args.steps, args.warmup, batch=1, att=att)
        metric, unit, hib, ms = "train_lobe_chunks_per_s", "chunks/s", True, s_per_step * 1e3
        workload = (TRAIN_ATT_WORKLOAD if att else TRAIN_WORKLOAD).format(B=args.batch)
        sample = (f"batch-1 sample of the workload: {args.steps} timed training steps of ONE 80^3 chunk each after {args.warmup} "
                  "warm-up, torch CPU fp32 (the B200 arm steps batch %d per GPU)" % args.batch)
    elif args.workload == "infer":
        v, dt, threads = cpu_infer_chunks_per_s(True, 1)
        metric, unit, hib, ms, workload = "infer_lobe_chunks_per_s", "chunks/s", True, dt * 1e3, INFER_WORKLOAD
        sample = "batch-1 sample: 1 eval forward of DC3DATGeneric + pooling on ONE 80^3 chunk, torch CPU fp32"
    else:
        _, dt, threads = cpu_infer_chunks_per_s(True, 1)
        v, metric, unit, hib, ms = 5.0 * dt, "seconds_per_ct_scan", "s/scan (wall time per scan of the whole job)", False, 5.0 * dt * 1e3
        workload = SCAN_WORKLOAD % SCAN_SHAPE
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


# ------------------------------------------------------------------------------------------------ B200 arm: training
def measure_train(dist, args, att, peaks, clocks=True):
    """One data-parallel training workload (DC3D = configs[1], DC3DATGeneric = configs[2]) -> dict of measurements
    (complete on rank 0).  The runner is built, timed device-resident, timed end to end, profiled per kernel and dropped."""
    import torch
    import job_runner
    from dram_native import lib
    from utils import Settings
    B, steps = args.batch, args.steps
    settings = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py" if att else "st_dram_ref.py"))
    settings.OPTIMIZER['lr'] = 1e-3                 # train.py default (train.py:33-34)
    settings.TRAIN_BATCH_SIZE = B
    torch.manual_seed(0)
    os.environ["DRAM_CUDA_GRAPH"] = os.environ.get("DRAM_BENCH_GRAPH", "1")
    runner = job_runner.LesionSegChunkTrain(settings_module=settings)
    host_batch = make_batch(B, seed=dist.rank, pinned=True)
    dev_batch = {k: (v.to(dist.dev) if hasattr(v, "to") else v) for k, v in host_batch.items()}

    # ---- device-resident leg (value): the runner captures the step in a CUDA graph after its eager warm-up steps
    for _ in range(max(args.warmup, 3)):
        runner.train_step(dev_batch)
    lib.PROFILE.reset()
    last = {}

    def step_dev():
        last["loss"] = runner.train_step(dev_batch)[0]
    ms_total, ck = timed_device(dist, step_dev, steps, clocks)
    graphed = runner._graph is not None
    launches = (runner.kernels_per_step * steps) if graphed else lib.PROFILE.launches
    final_loss = last["loss"].item()

    # ---- end-to-end leg: pinned host batch -> runner.train_step -> loss.item()
    nbytes_in = sum(host_batch[k].numel() * host_batch[k].element_size() for k in ("#image", "#lobe_reference", "#pseudo_lesion_reference"))
    runner.train_step(host_batch)[0].item()
    e2e_s = timed_host(dist, lambda: runner.train_step(host_batch)[0].item(), steps)

    # ---- per-kernel CUDA events: the same step run eagerly (a captured graph cannot carry event records per launch)
    os.environ["DRAM_CUDA_GRAPH"] = "0"
    runner.train_step(dev_batch)
    dist.barrier()
    lib.PROFILE.reset()
    lib.PROFILE.enabled = True
    ev_steps = min(steps, 3)
    for _ in range(ev_steps):
        runner.train_step(dev_batch)
    dist.barrier()
    lib.PROFILE.enabled = False
    prof = lib.PROFILE.summary()
    if os.environ.get("DRAM_BENCH_LAYERS") and dist.rank == 0:
        for k, v in sorted(lib.PROFILE.summary(by_tag=True).items(), key=lambda kv: -kv[1]["ms"])[:45]:
            tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0.0
            print(f"# {k:46s} calls/step {v['calls'] / ev_steps:4.1f} ms/step {v['ms'] / ev_steps:8.3f} TFLOP/s {tf:7.1f}", file=sys.stderr)
    os.environ["DRAM_CUDA_GRAPH"] = "1"
    if runner.reducer is not None:
        runner.reducer.remove()
    del runner
    torch.cuda.empty_cache()

    ms_step = ms_total / steps
    gflop = TRAIN_GFLOP_PER_CHUNK + (3 * 0.655 if att else 0.0)
    return {
        "metric": "train_lobe_chunks_per_s", "value": dist.world * B * steps / (ms_total / 1e3), "unit": "chunks/s",
        "ms_per_step": ms_step, "workload": (TRAIN_ATT_WORKLOAD if att else TRAIN_WORKLOAD).format(B=B),
        "e2e": {"value": dist.world * B * steps / e2e_s, "unit": "chunks/s", "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4,
                "api": "job_runner.LesionSegChunkTrain.train_step(pinned host batch) + loss.item()"},
        "clocks": ck, "gpu_launches": launches, "cuda_graph": graphed, "loss": final_loss,
        "roofline": dict(conv_roofline(prof, peaks, ev_steps, "events over %d eager repetitions of the timed step (the timed step "
                                       "itself is one CUDA-graph launch)" % ev_steps),
                         whole_step_tflops_per_gpu=B * gflop / 1e3 / (ms_step / 1e3)),
        "kernels": kernel_breakdown(prof, ev_steps),
    }


def measure_rank_spread(dist, args, sync_ms):
    """N > 1: the same training step with the data-parallel exchanges SUSPENDED — every GPU free-running on its own replica,
    all GPUs of the box loaded at the same time — gathered over the ranks.  The synchronised step cannot be faster than
    the slowest GPU's free-running step (28 BatchNorm exchanges per step make all ranks move in lockstep), so
    max(free-running) / synchronised is what the exchanges themselves cost; the spread between the GPUs (silicon and
    cooling differ under the same 1 kW cap) is what weak scaling loses before any communication."""
    import torch
    import job_runner
    from dram_native import dist as ddist
    from utils import Settings
    ddist.suspend(True)
    try:
        settings = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref.py"))
        settings.OPTIMIZER['lr'] = 1e-3
        settings.TRAIN_BATCH_SIZE = args.batch
        torch.manual_seed(0)
        runner = job_runner.LesionSegChunkTrain(settings_module=settings)
        batch = {k: (v.to(dist.dev) if hasattr(v, "to") else v) for k, v in make_batch(args.batch, seed=dist.rank, pinned=False).items()}
        for _ in range(max(args.warmup, 3)):
            runner.train_step(batch)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            runner.train_step(batch)
        e1.record()
        torch.cuda.synchronize()
        mine = e0.elapsed_time(e1) / args.steps
        del runner
        torch.cuda.empty_cache()
    finally:
        ddist.suspend(False)
    per_rank = [None] * dist.world
    dist.td.all_gather_object(per_rank, mine)
    return {"free_running_ms_per_step_by_rank": [round(v, 3) for v in per_rank], "slowest_gpu_ms": max(per_rank),
            "fastest_gpu_ms": min(per_rank), "synchronised_ms_per_step": sync_ms,
            "exchange_cost_ms": sync_ms - max(per_rank),
            "note": "free-running = same step, data-parallel exchanges suspended, all GPUs loaded simultaneously; the synchronised step "
                    "is bounded below by the slowest GPU"}


# ------------------------------------------------------------------------------------------------ B200 arm: inference
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


def measure_infer(dist, runner, steps, warmup, peaks, clocks=True):
    """DC3DATGeneric eval forward + per-lobe pooling on the 5 lobe chunks of a scan (batch 5 per GPU), chunks/s."""
    import torch
    from dram_native import lib
    model = runner.model.eval()
    B = 5
    batch = make_batch(B, seed=dist.rank, pinned=True)
    img_h = batch["#image"].unsqueeze(1).contiguous().pin_memory()
    lobe_h = batch["#lobe_reference"].unsqueeze(1).contiguous().pin_memory()
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

    for _ in range(max(warmup, 3)):
        fwd_dev()
    lib.PROFILE.reset()
    ms_total, ck = timed_device(dist, fwd_dev, steps, clocks)
    launches = lib.PROFILE.launches
    fwd_host()
    e2e_s = timed_host(dist, fwd_host, steps)
    lib.PROFILE.reset()
    lib.PROFILE.enabled = True
    ev = min(steps, 3)
    for _ in range(ev):
        fwd_dev()
    torch.cuda.synchronize()
    lib.PROFILE.enabled = False
    prof = lib.PROFILE.summary()
    k = prof.get("dram_conv3d_umma_fwd", {"ms": 0.0, "flops": 0.0, "calls": 0})
    achieved = k["flops"] / (k["ms"] / 1e3) / 1e12 if k["ms"] else 0.0
    return {
        "metric": "infer_lobe_chunks_per_s", "value": dist.world * B * steps / (ms_total / 1e3), "unit": "chunks/s",
        "ms_per_step": ms_total / steps, "workload": INFER_WORKLOAD,
        "e2e": {"value": dist.world * B * steps / e2e_s, "unit": "chunks/s", "h2d_bytes_per_step": 2 * img_h.numel() * 4,
                "d2h_bytes_per_step": B * 4, "api": "DC3DATGeneric.forward + pooling_dense_features on pinned host chunks, scores read back"},
        "clocks": ck, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "k_conv_umma_fwd4 (SM pairs) / _fwd / _fwd3 (13 forward launches, folded BN + ReLU epilogues)",
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                     "traffic": None, "ms_per_step": k["ms"] / ev, "peak_source": peaks["source"] + ", sustained figure",
                     "note": "algorithmic FLOPs / CUDA-event time; 1/3 of the peak is the split-bf16 ceiling"},
        "kernels": kernel_breakdown(prof, ev, 14),
    }


def measure_scan(dist, runner, steps, warmup, small=False, clocks=True):
    """BASELINE configs[3]/[4]: process_pipeline full-CT inference on a synthetic 512x512x400 scan (spacing 0.7 mm in-plane,
    1.0 mm slices): resample to 1 mm, 5 lobe chunks, DC3DATGeneric, RAM -> heat map -> Otsu -> lesion masks, back-resample.
    Every rank processes its own stream of scans (scans sharded across GPUs, no collective)."""
    import torch
    import job_runner
    from dram_native import lib
    from synthetic import synthetic_scan
    shape = SCAN_SHAPE if not small else (100, 128, 128)
    spacing = list(SCAN_SPACING)
    scan, lobe, _, _ = synthetic_scan(shape, spacing, seed=dist.rank)
    scan_h, lobe_h = torch.from_numpy(scan).pin_memory(), torch.from_numpy(lobe).pin_memory()
    new_sp = [1.0, 1.0, 1.0]

    def one_scan(scan_in, lobe_in):
        s_t = runner.resample_to_working_grid(scan_in, spacing, "linear")
        l_t = runner.resample_to_working_grid(lobe_in, spacing, "nearest")
        out = runner.run_scan(s_t, l_t, new_sp, return_device=True)
        les = job_runner.ops_itk_back(out["lesion"], shape, new_sp, spacing, "nearest")
        post = job_runner.ops_itk_back(out["lesion_post"], shape, new_sp, spacing, "nearest")
        return les, post, out["ratio"]

    scan_d, lobe_d = scan_h.cuda(), lobe_h.cuda()
    for _ in range(max(warmup, 3)):
        one_scan(scan_d, lobe_d)
    lib.PROFILE.reset()
    ms_total, ck = timed_device(dist, lambda: one_scan(scan_d, lobe_d), steps, clocks)
    launches = lib.PROFILE.launches
    lib.PROFILE.reset()                                  # per-kernel CUDA events of one more scan (outside the timed region)
    lib.PROFILE.enabled = True
    one_scan(scan_d, lobe_d)
    torch.cuda.synchronize()
    lib.PROFILE.enabled = False
    prof = lib.PROFILE.summary()
    # end to end through the public API: LesionSegTest.run_scans over PINNED HOST scans (upload of scan i+1 and download
    # of the masks of scan i-1 overlap the kernels of scan i); every scan's H2D and D2H is inside the timed region
    for _ in runner.run_scans([(scan_h, lobe_h, spacing)] * 3):          # warm-up: slots, pinned buffers, allocator pools
        pass
    done = []
    e2e_s = timed_host(dist, lambda: done.append(sum(1 for _ in runner.run_scans([(scan_h, lobe_h, spacing)] * steps))), 1)
    assert done == [steps]
    t0 = time.perf_counter()                             # one scan alone: upload + kernels + download in series
    for _ in runner.run_scans([(scan_h, lobe_h, spacing)]):
        pass
    torch.cuda.synchronize()
    single = time.perf_counter() - t0
    s_per_scan = ms_total / 1e3 / steps
    k = prof.get("dram_conv3d_umma_fwd", {"ms": 0.0, "flops": 0.0})
    return {
        "metric": "seconds_per_ct_scan", "value": s_per_scan / dist.world, "unit": "s/scan (wall time per scan of the whole job)",
        "higher_is_better": False, "ms_per_step": ms_total / steps, "scans_per_s": dist.world / s_per_scan,
        "workload": SCAN_WORKLOAD % shape,
        "e2e": {"value": e2e_s / steps / dist.world, "unit": "s/scan", "h2d_bytes_per_step": int(scan_h.numel() * 2 + lobe_h.numel()),
                "d2h_bytes_per_step": int(2 * lobe_h.numel() + 4), "api": "LesionSegTest.run_scans (pinned host scans in, host masks out; "
                "pipelined copies)", "single_scan_latency_s": single},
        "clocks": ck, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "k_conv_umma_fwd4 (SM pairs) / _fwd / _fwd3 (13 launches per scan, the dominant kernels of a scan)",
                     "achieved": k["flops"] / (k["ms"] / 1e3) / 1e12 if k["ms"] else 0.0, "unit": "TFLOP/s",
                     "ms_per_scan": k["ms"], "share_of_kernel_time": k["ms"] / (sum(v["ms"] for v in prof.values()) or 1.0),
                     "traffic": None},
        "kernels": {kk.replace("dram_", ""): {"calls_per_scan": v["calls"], "ms_per_scan": v["ms"]}
                    for kk, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:16]},
    }


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

    def row(nbytes, ms, **kw):
        return {"bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"], **kw}

    out["ram_reduce_fwd"] = row(4.0 * B * V * (C + 1), timed(lambda: ops.ram_reduce(feat, w, b)))
    ram = torch.randn(B, V, device="cuda")
    mask = (torch.rand(B, V, device="cuda") > 0.4).float()
    out["masked_pool_fwd"] = row(8.0 * B * V, timed(lambda: ops.masked_pool(ram, mask, True, True)))
    SD, SH, SW = 400, 358, 358
    labels = torch.ones((SD, SH, SW), dtype=torch.uint8, device="cuda")
    heat = torch.zeros((SD, SH, SW), device="cuda")
    crop = ((20, 260), (30, 230), (40, 220))
    vc = 240 * 200 * 180
    r1 = ops.ram_activation(torch.randn(1, 1, 80, 80, 80, device="cuda"), 1)[0, 0]   # sigmoid first, as LesionSegTest.run does
    out["ram_upsample_label_scatter"] = row(4.0 * V + 5.0 * vc, timed(lambda: ops.ram_upsample_label_scatter(r1, labels, 1, heat, crop, 0, 1.0)),
                                            crop_voxels=vc)
    Bp, G = 5, 64
    f = ops.new_volume(Bp, 17, G, G, G, "cuda")
    f.normal_()
    cam = torch.randn(Bp, 1, G, G, G, device="cuda")
    pcm = models.PCM((G, G, G), 17, 1, 8, 0, 8, 1, 3, "scaled_dot_product_relu", False, p_enc_dim=0).cuda()
    with torch.no_grad():
        ms = timed(lambda: pcm(cam, f))
    out["pcm_fwd"] = row(76.0 * Bp * G ** 3, ms, note="algorithmic 76 B/voxel (f 17x4 + cam 4 in, 4 out), inference form (nothing kept for a backward)")
    # x2 trilinear upsample of the decoder (parts.py:149-153) at the us2 level: 128 channels, 40^3 -> 80^3, batch 8
    del feat, ram, mask, labels, heat, f, cam
    C1, C2, d = 128, 64, 40
    xs = ops.split_bf16(ops.new_volume(B, C1, d, d, d, "cuda").normal_())
    sk = ops.split_bf16(ops.new_volume(B, C2, 2 * d, 2 * d, 2 * d, "cuda").normal_())
    vin, vout = B * d ** 3, B * (2 * d) ** 3
    out["upsample2x_planes_fwd"] = row(4.0 * C1 * (vin + vout), timed(lambda: ops.upsample2x_virtual_concat(xs, sk)),
                                       note="bf16 split planes in (4 B/element) and out; the skip half of the concat is never copied")
    g = ops.new_volume(B, C1 + C2, 2 * d, 2 * d, 2 * d, "cuda").normal_()
    out["upsample2x_adjoint"] = row(4.0 * C1 * (vin + vout),
                                    timed(lambda: ops.upsample2x_concat_bwd(g, (B, C1, d, d, d), (B, C2, 2 * d, 2 * d, 2 * d), want_dskip=False)),
                                    note="reads the C1 channels of the fp32 concat gradient once, writes the source gradient")
    return out


def cudnn_baseline(B, steps=3):
    """The 'kernel to beat' of SURVEY §8d / VERDICT r1 #9: the SAME training step as stock PyTorch modules' arithmetic
    (F.conv3d / F.batch_norm / ... = cuDNN + ATen on this B200, via the oracle's functional restatement), with TF32
    allowed (the reference's effective GPU default: cudnn.allow_tf32 = True, train.py:44-45) and disallowed (fp32)."""
    import torch
    out = {}
    torch.backends.cudnn.benchmark = True
    for name, tf32 in (("tf32", True), ("fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            v, s_per, _ = oracle_train_chunks_per_s(steps, 2, batch=B, att=False, device="cuda")
            out[name] = {"value": v, "unit": "chunks/s", "ms_per_step": s_per * 1e3}
        except Exception as e:                                  # noqa: BLE001  (informational leg: never fail the bench)
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    out["what"] = (f"stock PyTorch {torch.__version__} + cuDNN {torch.backends.cudnn.version()} on this GPU: DC3D fwd + IntRegRefineLoss + bwd + Adam, "
                   f"batch {B}, 80^3, fp32 tensors, cudnn.benchmark=True; tf32 = cudnn/matmul allow_tf32 (reference default), fp32 = disallowed; "
                   "eager, no CUDA graph, wall clock over %d steps after 2 warm-up; informational only" % steps)
    return out


def run_b200(args):
    import torch
    dist = Dist()
    peaks = load_peaks()
    out = None
    if args.workload == "train":
        att = args.model == "att"
        main = measure_train(dist, args, att, peaks)
    elif args.workload == "infer":
        main = measure_infer(dist, build_att_runner(), args.steps, args.warmup, peaks)
    else:
        main = measure_scan(dist, build_att_runner(), args.steps, args.warmup, args.small)

    extra = None
    if args.workload == "train" and args.model == "dc3d" and not args.no_extra:
        # the other parts of BASELINE.json's metric, on this same line, at every N (all ranks take part)
        extra = {"train_att": measure_train(dist, args, True, peaks, clocks=False)}
        runner = build_att_runner()
        extra["infer"] = measure_infer(dist, runner, max(args.steps, 10), args.warmup, peaks, clocks=False)
        extra["scan"] = measure_scan(dist, runner, min(max(args.steps, 5), 10), min(args.warmup, 3), clocks=False)
        del runner
        torch.cuda.empty_cache()
        if dist.world > 1:
            extra["rank_spread"] = measure_rank_spread(dist, args, main["ms_per_step"])
        if dist.world == 1:
            extra["hbm_kernels"] = hbm_kernel_rooflines(peaks)
            if os.environ.get("DRAM_BWD_PRECISION", "bf16x3") != "bf16x2":
                # the same step with the gradient operand of dgrad / wgrad carried as ONE bf16 plane (ops.grad_planes_three):
                # reported next to the headline, never as it - the parity tolerances on gradients are stated for bf16x3
                os.environ["DRAM_BWD_PRECISION"] = "bf16x2"
                x2 = measure_train(dist, args, False, peaks, clocks=False)
                os.environ["DRAM_BWD_PRECISION"] = "bf16x3"
                extra["train_bwd_bf16x2"] = {"metric": x2["metric"], "value": x2["value"], "ms_per_step": x2["ms_per_step"],
                                             "note": "DRAM_BWD_PRECISION=bf16x2: dy as one bf16 plane x split weights / split layer "
                                                     "input (2 MMAs per MAC in dgrad and wgrad); forward unchanged"}
            if not args.no_cudnn:
                extra["cudnn_baseline"] = cudnn_baseline(args.batch)

    if dist.rank == 0:
        cpu = None
        if dist.world == 1 and not args.no_cpu:
            if args.workload == "train":
                v, s_per, threads = oracle_train_chunks_per_s(1, 0, batch=1, att=args.model == "att")
                cpu = {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                       "sample": f"batch-1 sample: 1 training step on ONE 80^3 chunk ({s_per:.1f} s), torch CPU fp32 oracle"}
                if extra is not None:
                    v, s_per, threads = oracle_train_chunks_per_s(1, 0, batch=1, att=True)
                    extra["train_att"]["cpu_baseline"] = {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                                                          "sample": f"batch-1 sample: 1 DC3DATGeneric training step on ONE 80^3 chunk ({s_per:.1f} s), torch CPU fp32 oracle"}
            if args.workload != "train" or extra is not None:
                v, dt, threads = cpu_infer_chunks_per_s(True, 1)
                ci = {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                      "sample": f"batch-1 sample: 1 eval forward of DC3DATGeneric + pooling on ONE 80^3 chunk ({dt:.1f} s), torch CPU fp32 oracle"}
                cs = {"value": 5.0 * dt, "unit": "s/scan", "cores": threads, "kind": "port",
                      "sample": f"model part only: one of the scan's 5 lobe chunks through the CPU oracle ({dt:.1f} s) x 5; the reference "
                                "additionally spends CPU time in SimpleITK resampling and numpy masking"}
                if extra is not None:
                    extra["infer"]["cpu_baseline"], extra["scan"]["cpu_baseline"] = ci, cs
                else:
                    cpu = ci if args.workload == "infer" else cs
        B = args.batch
        config = {"workload": main["workload"], "chunk": list(CHUNK), "parallelism": f"dp{dist.world}",
                  "precision_mode": os.environ.get("DRAM_PRECISION", "bf16x3")}
        if args.workload == "train":
            config.update(per_gpu_batch=B, global_batch=B * dist.world, cuda_graph=main["cuda_graph"],
                          backward_precision=os.environ.get("DRAM_BWD_PRECISION", "bf16x3"),
                          grad_allreduce="overlapped buckets" if os.environ.get("DRAM_GRAD_OVERLAP", "0") == "1" else "one flat all-reduce after backward",
                          l2="no flush needed: ~13 GB of activations stream through the 126 MB L2 every step")
        elif args.workload == "infer":
            config.update(per_gpu_batch=5, l2="activations (~5 GB per batch) stream through L2")
        else:
            config.update(scans_per_s=main["scans_per_s"], l2="a scan's volumes (~1 GB) stream through L2")
        out = {
            "metric": main["metric"], "value": main["value"], "unit": main["unit"], "n_gpus": dist.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": main.get("higher_is_better", True),
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config,
            "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
            "kernels": main["kernels"], "cpu_baseline": cpu, "extra": extra,
        }
        if "loss" in main:
            out["loss"] = main["loss"]
        emit(json.dumps(out))
    dist.close()


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-cudnn", action="store_true", help="skip extra.cudnn_baseline")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "scan"],
                    help="train = BASELINE configs[1] (default, the driver's line, carries the others under `extra`); "
                         "infer = chunk inference; scan = full-CT pipeline")
    ap.add_argument("--model", default="dc3d", choices=["dc3d", "att"],
                    help="train workload: dc3d = BASELINE configs[1]; att = DC3DATGeneric with the PCM head (configs[2])")
    ap.add_argument("--small", action="store_true", help="scan workload: 128x128x100 scan (quick check)")
    ap.add_argument("--no-extra", action="store_true", help="train workload: only the headline (no extra.* parts)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
