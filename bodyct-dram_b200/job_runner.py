"""Job runners for the B200 hot path, mirroring the caller contract of the reference's `dram/job_runner.py`
(JobRunner.init :356-413, LesionSegChunkTrain.train :649-681 / evaluate_scan :720-779, LesionSegTest.run :906-1067).

Only what surrounds the hot path is kept: model/optimizer/loss/scheduler construction from a settings module, the
training step, checkpoint save/reload with the reference's dictionary layout, and the per-scan inference loop.
Dataset classes, CSV bookkeeping, tensorboard, confusion-matrix plots and .mha writers are out of scope;
batches / scans are handed in by the caller as arrays or tensors.
"""
import copy
import glob
import logging
import os
import time

import numpy as np
import torch

from dram_native import dist as ddist
from utils import AverageMeter, Settings, get_callable_by_name


class JobRunner:
    """Builds model / initializer / optimizer / loss / scheduler from `settings` exactly as job_runner.py:356-413."""

    def __init__(self, setting_module_file_path=None, settings_module=None, **kwargs):
        if setting_module_file_path is None:
            self.settings = copy.deepcopy(settings_module)
        else:
            self.settings = Settings(setting_module_file_path)
        self.logger = logging.getLogger(type(self).__name__)
        self.exp_path = os.path.join(self.settings.MODEL_ROOT_PATH, self.settings.EXP_NAME)
        self.debug_path = os.path.join(self.settings.DEBUG_PATH, self.settings.EXP_NAME)
        self.current_iteration = 0
        self.epoch_n = 0
        self.saved_model_states = {}
        self.ctss_frequency_map = {k: 1.0 / 6 for k in range(6)}     # set by the sampler in the reference

    def init(self):
        s = self.settings
        init_cfg, model_cfg, opt_cfg = dict(s.INITIALIZER), dict(s.MODEL), dict(s.OPTIMIZER)
        loss_cfg, sched_cfg = dict(s.LOSS_FUNC), dict(s.SCHEDULER)
        self.parameter_initializer = get_callable_by_name(init_cfg.pop('method'))(**init_cfg)
        self.model = get_callable_by_name(model_cfg.pop('method'))(**model_cfg)
        self.is_cuda = bool(s.IS_CUDA) and torch.cuda.is_available()
        if not self.is_cuda:
            raise RuntimeError("the DRAM B200 path needs a CUDA device (IS_CUDA and torch.cuda.is_available())")
        self.model = self.model.cuda()
        self.model.init(self.parameter_initializer)
        opt_cls = get_callable_by_name(opt_cfg.pop('method'))
        if opt_cls is torch.optim.Adam:
            opt_cfg.setdefault("fused", True)
            opt_cfg.setdefault("capturable", True)          # step counter on the device: the step can live in a CUDA graph
            if opt_cfg["capturable"] and not isinstance(opt_cfg.get("lr"), torch.Tensor):
                # the learning rate lives on the device too: a Python float would be baked into the captured kernel
                # arguments and scheduler.step() / param_group['lr'] edits would be ignored by every graph replay
                # (torch's schedulers fill_() a tensor lr in place)
                opt_cfg["lr"] = torch.tensor(float(opt_cfg.get("lr", 1e-3)), dtype=torch.float32, device="cuda")
        self.optimizer = opt_cls(self.model.parameters(), **opt_cfg)
        self.loss_func = get_callable_by_name(loss_cfg.pop('method'))(**loss_cfg)
        self.scheduler = get_callable_by_name(sched_cfg.pop('method'))(self.optimizer, **sched_cfg)
        self.reducer = ddist.GradReducer(self.model.parameters()) if ddist.active() else None
        if ddist.active():                                   # replicas start from rank 0's weights
            ddist.init_peer()                                # NVLink mailboxes for the BatchNorm / loss-normaliser exchanges
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                torch.distributed.broadcast(t.data, src=0)
            self._weights_changed()

    @staticmethod
    def _weights_changed():
        """Writes through `.data` (broadcast, load_state_dict) bump neither `_version` nor the optimizer-step generation:
        drop the packed bf16 copies of the weights explicitly."""
        from dram_native import functional as _DF
        _DF.WEIGHTS.invalidate()

    # -------------------------------------------------------------------------------------------- checkpoints
    def reload_state(self, model_state):
        """Tolerant reload (job_runner.py:477-497): keep only keys that exist with the same shape."""
        own = self.model.state_dict()
        kept = {k: v for k, v in model_state.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        for k, v in model_state.items():
            if k not in own:
                self.logger.warning("checkpoint key %s does not exist in the model: ignored", k)
            elif k not in kept:
                self.logger.warning("in %s, saved tensor size %s does not match current tensor size %s: ignored",
                                    k, tuple(v.shape), tuple(own[k].shape))
        missing = sorted(set(own) - set(kept))
        if missing:
            self.logger.warning("%d model tensors keep their initial values (not in the checkpoint): %s%s", len(missing),
                                ", ".join(missing[:8]), " ..." if len(missing) > 8 else "")
        own.update(kept)
        self.model.load_state_dict(own)
        self._weights_changed()
        return sorted(set(model_state) - set(kept))

    def reload_model_from_cache(self, path=None):
        path = path or getattr(self.settings, "RELOAD_CHECKPOINT_PATH", None)
        if path is None:                 # newest *.pth of the experiment folder (job_runner.py:524-530)
            cands = sorted(glob.glob(os.path.join(self.exp_path, "*.pth")), key=os.path.getmtime)
            if not cands:
                raise RuntimeError("{} has no checkpoint files with pth extensions.".format(self.exp_path))
            path = cands[-1]
        if not os.path.exists(path):
            raise RuntimeError("checkpoint {} does not exist.".format(path))
        self.logger.info("reloading model from %s.", path)
        state = torch.load(path, map_location="cpu")
        self.saved_model_states = state
        if "model" in self.settings.RELOAD_DICT_LIST:
            self.reload_state(state.get("model", state.get("model_dict", {})))
        if "optimizer" in self.settings.RELOAD_DICT_LIST and "optimizer" in state:
            self.optimizer.load_state_dict(state["optimizer"])
            for g in self.optimizer.param_groups:        # keep the learning rate a device tensor (see init)
                if g.get("capturable") and not (isinstance(g["lr"], torch.Tensor) and g["lr"].is_cuda):
                    g["lr"] = torch.tensor(float(g["lr"]), dtype=torch.float32, device="cuda")
        self.epoch_n = state.get("epoch", 0)
        self.current_iteration = state.get("iteration", 0)
        return True

    def save_checkpoint(self, path=None, metrics=None):
        """Same dictionary layout as job_runner.py:707-716."""
        os.makedirs(self.exp_path, exist_ok=True)
        path = path or os.path.join(self.exp_path, f"{self.epoch_n}.pth")
        torch.save({"epoch": self.epoch_n, "iteration": self.current_iteration, "model": self.model.state_dict(),
                    "optimizer": self.optimizer.state_dict(), "metrics": metrics or {}}, path)
        return path


class _ScanPipeline:
    """The per-scan loop shared by validation (LesionSegChunkTrain.evaluate_scan, job_runner.py:720-779) and full-CT
    inference (LesionSegTest.run, job_runner.py:942-1030): one scan stays on the GPU from upload to mask download.
    Needs `self.model`, `self.settings`, `self.head` ('sigmoid' | 'literal') and `self.crop_border` (mm)."""

    # ---- stages -------------------------------------------------------------------------------------------------
    def lobe_crops(self, lobe_t, spacing):
        """utils.find_crops for the 5 lobes; one 120-byte device->host read per scan."""
        from dram_native import ops
        import math
        boxes = ops.read_small(ops.label_bboxes(lobe_t, 5)).numpy()
        crops = {}
        for label in range(1, 6):
            mn, mx = boxes[label, :3], boxes[label, 3:]
            if mx[0] < 0:
                continue
            crops[label] = tuple((max(0, int(mn[a]) - int(math.ceil(self.crop_border / spacing[a]))),
                                  min(int(lobe_t.shape[a]), int(mx[a]) + 1 + int(math.ceil(self.crop_border / spacing[a]))))
                                 for a in range(3))
        return crops

    def preprocess(self, scan_t, lobe_t, crops, spacing=(1.0, 1.0, 1.0)):
        from dram_native import ops
        size = tuple(self.settings.RESAMPLE_SIZE)
        n = len(crops)
        imgs = torch.empty((n, 1) + size, device=scan_t.device, dtype=torch.float32)
        msks = torch.empty((n, 1) + size, device=scan_t.device, dtype=torch.float32)
        window = (self.settings.WINDOWING_MIN, self.settings.WINDOWING_MAX)
        for i, (label, crop) in enumerate(crops.items()):
            ops.lobe_chunk_preprocess(scan_t, lobe_t, label, crop, window, self.settings.PAD_VALUE, imgs[i, 0], msks[i, 0],
                                      spacing)
        return imgs, msks

    def paste(self, dense, msks, lobe_t, crops, heat):
        from dram_native import ops
        import models
        if self.head == 'literal':
            pool = models.pooling_dense_features(dense, msks)
            cls_pred = torch.max(pool, dim=-1)[-1].tolist()                # always 0 for out_ch == 1
        probs = ops.ram_activation(dense, 1) if self.head != 'literal' else None      # F.sigmoid(dense_outs), job_runner.py:765
        for i, (label, crop) in enumerate(crops.items()):
            if self.head == 'literal':
                if cls_pred[i] < 1e-7:
                    continue                                                 # dense_out.zero_() -> heat map stays 0
                mx = torch.zeros(1, device=dense.device)
                (z0, z1), (y0, y1), (x0, x1) = crop
                ops.ram_upsample_mask_scatter(dense[i, cls_pred[i]], (lobe_t[z0:z1, y0:y1, x0:x1] == label).to(torch.uint8).contiguous(),
                                              None, (0, 0, 0), 2, 1.0, mx)
                ops.ram_upsample_label_scatter(dense[i, cls_pred[i]], lobe_t, label, heat, crop, 2, 1.0 / mx.item())
            else:
                ops.ram_upsample_label_scatter(probs[i, 0], lobe_t, label, heat, crop, 0, 1.0)

    def postprocess(self, heat, scan_t, lobe_t, vessel_t=None):
        """binary_cam / Otsu thresholds and lesion masks (job_runner.py:1006-1015)."""
        from dram_native import ops
        from utils import otsu_threshold_from_histogram

        def otsu(values, lo, hi, scaler):
            hist = ops.read_small(ops.masked_hist_u8(values, lobe_t, lo, hi)).numpy().astype(np.int64)   # 1 KB, no copy engine
            if hist.sum() == 0:
                raise ValueError("empty array encountered! cam_probs.size == 0.")
            if np.count_nonzero(hist) < 2:
                return float(np.nonzero(hist)[0][0]) / 255.0
            return min(otsu_threshold_from_histogram(hist) * scaler, 255.0) / 255.0

        th = otsu(heat, 0.0, 1.0, 1.0)
        th2 = otsu(scan_t, -1150.0, 350.0, 0.75)
        lesion, post = ops.threshold_masks(heat, th, scan_t, vessel_t, th2, (-1150.0, 350.0))
        return lesion, post, th, th2

    # ---- one scan -----------------------------------------------------------------------------------------------
    def run_scan(self, scan, lobe, spacing, vessel=None, return_device=False, postprocess=True):
        """scan: int16 [D,H,W] (numpy or tensor, already at TEST_RESAMPLE_SPACING), lobe: uint8 labels 0..5.
        postprocess=False stops after the heat map and the lesion ratio (all evaluate_scan needs, job_runner.py:772)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev, dt, non_blocking=True)
        scan_t, lobe_t = as_t(scan, torch.int16), as_t(lobe, torch.uint8)
        vessel_t = as_t(vessel, torch.uint8) if vessel is not None else None
        self.model.eval()
        with torch.no_grad():
            crops = self.lobe_crops(lobe_t, spacing)
            heat = torch.zeros(scan_t.shape, device=dev, dtype=torch.float32)
            if crops:
                imgs, msks = self.preprocess(scan_t, lobe_t, crops, spacing)
                _, dense = self.model(imgs, msks)
                self.paste(dense, msks, lobe_t, crops, heat)
            from dram_native import ops
            ratio = ops.labelled_mean(heat, lobe_t)                         # job_runner.py:772, one pass over heat + labels
            if not postprocess:
                return {"heatmap": heat if return_device else heat.cpu().numpy(), "crops": crops,
                        "ratio": ratio if return_device else float(ratio.item())}
            lesion, post, th, th2 = self.postprocess(heat, scan_t, lobe_t, vessel_t)
        out = {"heatmap": heat, "lesion": lesion, "lesion_post": post, "threshold": th, "threshold_post": th2,
               "ratio": ratio, "crops": crops}
        if not return_device:
            out.update(heatmap=heat.cpu().numpy(), lesion=lesion.cpu().numpy(), lesion_post=post.cpu().numpy(),
                       ratio=float(ratio.item()))
        return out

    def resample_to_working_grid(self, arr_t, spacing, mode):
        """Resample('fixed_spacing', TEST_RESAMPLE_SPACING) of a whole scan (job_runner.py:827-835)."""
        from dram_native import ops
        new_sp = float(self.settings.TEST_RESAMPLE_SPACING)
        new_size = [int(np.ceil(s * (sp / new_sp))) for s, sp in zip(arr_t.shape, spacing)]     # utils.py:369-371
        return ops.itk_resample(arr_t, new_size, mode, in_spacing=spacing, out_spacing=[new_sp] * 3)

    def scan_to_masks(self, scan_d, lobe_d, spacing):
        """One scan at its ORIGINAL grid, already on the device -> (lesion, lesion_post) uint8 masks at the original grid and
        the lesion ratio (device tensors): resample to the working grid, run_scan, resample back (job_runner.py:827-835,
        942-1030)."""
        new_sp = [float(self.settings.TEST_RESAMPLE_SPACING)] * 3
        s_t = self.resample_to_working_grid(scan_d, spacing, "linear")
        l_t = self.resample_to_working_grid(lobe_d, spacing, "nearest")
        out = self.run_scan(s_t, l_t, new_sp, return_device=True)
        shape = tuple(scan_d.shape)
        return (ops_itk_back(out["lesion"], shape, new_sp, spacing, "nearest"),
                ops_itk_back(out["lesion_post"], shape, new_sp, spacing, "nearest"), out["ratio"])


class LesionSegChunkTrain(JobRunner, _ScanPipeline):
    """Training runner: `train_step(batch)` is the body of the reference's hot loop (job_runner.py:657-674)."""

    GRAPH_WARMUP_STEPS = 2        # eager steps before capture (lazy one-time setup, allocator warm-up, Adam state)

    def __init__(self, settings_module=None, setting_module_file_path=None, loader_factory=None, val_dataset=None):
        """`loader_factory(runner, epoch) -> iterable of batch dicts` and `val_dataset` (iterable of scan dicts, the
        schema of dataset.py:401-407) stand in for the reference's institute datasets (reset_data, job_runner.py:621-647);
        without a factory, `settings.SYNTHETIC_STEPS` synthetic batches per epoch are used (0: nothing to train on)."""
        super().__init__(setting_module_file_path, settings_module)
        self.init()
        self.metrics = {}
        self.head, self.crop_border, self.trace = 'sigmoid', 5, False
        self.loader_factory, self.val_dataset, self.tr_loader = loader_factory, val_dataset, None
        if getattr(self.settings, "RELOAD_CHECKPOINT", False):
            self.reload_model_from_cache()
        self._graph = self._graph_key = self._static_in = self._static_labels = self._static_out = None
        self._eager_steps = 0
        self.kernels_per_step = None
        self._warm_stream = None

    # -------------------------------------------------------------------------------------------- data
    def reset_data(self):
        """job_runner.py:621-647 without the institute archive: (re)build this epoch's loader.  Under data parallelism
        every rank draws its own shard (different seed / `rank::world` of the caller's loader) with the SAME number of
        equally shaped batches — SyncBN statistics and the captured graph rely on it (dist.check_uniform_batch)."""
        if self.loader_factory is not None:
            self.tr_loader = self.loader_factory(self, self.epoch_n)
        else:
            from train import synthetic_loader
            steps = int(getattr(self.settings, "SYNTHETIC_STEPS", 0))
            seed = self.epoch_n * ddist.world_size() + ddist.rank()
            self.tr_loader = synthetic_loader(steps, self.settings.TRAIN_BATCH_SIZE, tuple(self.settings.RESAMPLE_SIZE), seed=seed)
        return self.tr_loader

    def _step_body(self, images, lobes, lesions, ctsses, metas, labels):
        """forward + loss + backward (+ gradient all-reduce) + optimizer step; enqueues only, never syncs."""
        self.optimizer.zero_grad(set_to_none=True)
        loss_tuple = self.loss_func(self.model, images, lobes, lesions, ctsses, obj=self, metas=metas, label_tensors=labels)
        factors = self.settings.LOSS_FACTORS[:len(loss_tuple)]
        loss = torch.stack([l * w for l, w in zip(loss_tuple, factors)]).sum()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return loss, loss_tuple

    def train_step(self, batch_data):
        """batch_data: {"#image", "#lobe_reference", "#pseudo_lesion_reference": [B,D,H,W] host or device tensors,
        "meta": {"cle": [...]}} -> (loss tensor, loss tuple).  Includes the H2D copy, backward, gradient all-reduce
        (data parallel) and the optimizer step; does NOT sync with the host.

        Steps with a stable batch shape are captured once into a CUDA graph (after `GRAPH_WARMUP_STEPS` eager steps) —
        under data parallelism together with their NCCL all-reduces — and replayed, so the ~800 kernel launches of a
        step cost one launch and host jitter cannot starve the GPU; set DRAM_CUDA_GRAPH=0 to stay eager.  The returned
        tensors are then the graph's static outputs."""
        self.model.train()
        dev = torch.device("cuda", torch.cuda.current_device())
        metas = batch_data["meta"]
        ctsses = metas["cle"]
        labels = self.loss_func.label_tensors(ctsses, self.ctss_frequency_map, dev) \
            if hasattr(self.loss_func, "label_tensors") else None
        srcs = [batch_data[k] for k in ("#image", "#lobe_reference", "#pseudo_lesion_reference")]
        key = tuple(tuple(t.shape) for t in srcs)
        if self.reducer is not None and key != getattr(self, "_checked_key", None):
            ddist.check_uniform_batch(key)            # SyncBN counts and graph captures assume equal shapes on every rank
            self._checked_key = key
        use_graph = os.environ.get("DRAM_CUDA_GRAPH", "1") == "1" and labels is not None
        self.current_iteration += 1
        if not use_graph or self._graph_key not in (None, key):
            self._graph = None
            images, lobes, lesions = (t.to(dev, torch.float32, non_blocking=True).unsqueeze(1) for t in srcs)
            return self._step_body(images, lobes, lesions, ctsses, metas, labels)
        if self._graph is None:
            self._graph_key = key
            self._eager_steps += 1
            images, lobes, lesions = (t.to(dev, torch.float32, non_blocking=True).unsqueeze(1) for t in srcs)
            if self._eager_steps <= self.GRAPH_WARMUP_STEPS:
                # warm up on a side stream (torch CUDA-graph recipe): autograd's AccumulateGrad nodes must not be bound to
                # the legacy default stream, or the later capture would have to synchronise with it
                if self._warm_stream is None:
                    self._warm_stream = torch.cuda.Stream()
                self._warm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._warm_stream):
                    out = self._step_body(images, lobes, lesions, ctsses, metas, labels)
                torch.cuda.current_stream().wait_stream(self._warm_stream)
                return out
            self._static_in = [torch.empty_like(t) for t in (images, lobes, lesions)]
            self._static_labels = {k: torch.empty_like(v) for k, v in labels.items()}
            from dram_native import lib as _dlib
            torch.cuda.synchronize()
            was_profiling, _dlib.PROFILE.enabled = _dlib.PROFILE.enabled, False      # no event records inside a capture
            launches0 = _dlib.PROFILE.launches
            from dram_native import functional as _DF
            _DF.WEIGHTS.invalidate()              # the weight re-pack kernels must be part of the captured step
            graph = torch.cuda.CUDAGraph()
            # data parallel: the NCCL all-reduces (BatchNorm sums, gradient buckets, loss normalisers) are captured with
            # the step; NCCL's watchdog thread polls events meanwhile, hence the thread-local capture mode
            mode = "thread_local" if self.reducer is not None else "global"
            with torch.cuda.graph(graph, capture_error_mode=mode):
                self._static_out = self._step_body(*self._static_in, ctsses, metas, self._static_labels)
            self.kernels_per_step = _dlib.PROFILE.launches - launches0             # libdram_b200 kernels in one step
            _dlib.PROFILE.enabled = was_profiling
            self._graph = graph
        for dst, src in zip(self._static_in, srcs):
            dst.copy_(src.unsqueeze(1), non_blocking=True)             # H2D (pinned) or D2D into the graph's inputs
        for k, v in labels.items():
            self._static_labels[k].copy_(v, non_blocking=True)
        self._graph.replay()
        from dram_native import functional as _DF
        _DF.WEIGHTS.invalidate()           # the replay changed the weights behind PyTorch's version counters
        return self._static_out

    def close(self):
        """Drop the captured step (a CUDA graph that holds NCCL work must be gone before the process group is destroyed, or
        the teardown waits on it forever) and the gradient hooks.  The runner can keep training afterwards (it re-captures)."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self._graph = self._graph_key = self._static_in = self._static_labels = self._static_out = None
        self._eager_steps = 0
        if self.reducer is not None:
            self.reducer.remove()
            self.reducer = None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def train(self, loader=None):
        """One epoch over `loader` (default: self.tr_loader) — job_runner.py:649-681."""
        loader = self.tr_loader if loader is None else loader
        batch_time, loss_record = AverageMeter(), AverageMeter()
        end = time.time()
        for step_idx, batch in enumerate(loader):
            loss, loss_tuple = self.train_step(batch)
            if self.current_iteration % self.settings.LOG_STEPS == 0:
                loss_record.update(loss.item(), batch["#image"].shape[0])          # the only host sync of a step
                batch_time.update(time.time() - end)
                end = time.time()
                self.logger.info(f"Epoch: [{self.epoch_n}][{step_idx}], Time {batch_time.val:.3f} "
                                 f"Loss {loss_record.val:.6f} ({loss_record.avg:.6f}), "
                                 f"losses: {[f'{l.item():.5f}' for l in loss_tuple]}")
        return {'tr_loss': loss_record.avg, 'tr_batch_time': batch_time.avg}

    def run(self):
        """The reference's epoch loop (job_runner.py:683-718): reset_data / train every epoch, validate on the reference's
        schedule, step the LR scheduler ONLY on validation epochs (job_runner.py:706), checkpoint every STATE_EPOCHS with
        the reference's dictionary layout.  Rank 0 writes the checkpoints under data parallelism."""
        s = self.settings
        self.logger.info("start running epochs from {} to {}. ".format(self.epoch_n, s.NUM_EPOCHS))
        for epoch_n in range(self.epoch_n, s.NUM_EPOCHS):
            self.epoch_n = epoch_n
            self.reset_data()
            tr_metric_dict = self.train()
            last = self.epoch_n == s.NUM_EPOCHS - 1
            if self.epoch_n % s.VAL_EPOCHS == 0 or last or self.epoch_n < 15:
                val_metric_dict = self.validate()
                if len(val_metric_dict) > 0:
                    self.metrics.update(val_metric_dict)
                    self.metrics.update(tr_metric_dict)
                self.scheduler.step()
            if (self.epoch_n % s.STATE_EPOCHS == 0 or last) and ddist.rank() == 0:
                self.save_checkpoint(metrics=dict(self.metrics))
                self.logger.info(f"Saved epoch {self.epoch_n}/ {s.NUM_EPOCHS}.")
        self.logger.info(f"Training stops at epoch {self.epoch_n}/ {s.NUM_EPOCHS}.")
        return dict(self.metrics)

    def evaluate_scan(self, scan_data):
        """job_runner.py:720-779: scan dict {'#image' int16 [D,H,W], '#lobe_reference' uint8 labels, 'meta': {uid, spacing,
        cle, ...}} -> (reg_cls_pred, scan_cls_target, seconds).  Every lobe is cropped (+5 mm), blanked outside the lobe,
        windowed, resampled to the chunk grid, run through the model; sigmoid(RAM) is trilinearly resized to the crop and
        pasted under the lobe mask; the scan's lesion ratio is mapped to a CT severity class (ratio_to_label)."""
        meta = scan_data['meta']
        now = time.time()
        out = self.run_scan(scan_data['#image'], scan_data['#lobe_reference'], list(meta['spacing']), postprocess=False)
        reg_cls_pred = self.loss_func.ratio_to_label([out["ratio"]])[0]
        scan_cls_target = int(float(meta["cle"]))
        self.logger.info(f"val scan {meta.get('uid')}, , reg_cls_pred: {reg_cls_pred}, scan_cls_target: {scan_cls_target}.")
        return reg_cls_pred, scan_cls_target, time.time() - now

    def validate(self):
        """job_runner.py:781-811: evaluate_scan over the validation scans -> {'val_time', 'val_acc_reg_cls'}; the confusion
        matrix plot is reporting and out of scope.  Data parallel: scans are sharded `rank::world`, the per-scan results
        are gathered, so every rank returns the same metrics."""
        scans = list(self.val_dataset) if self.val_dataset is not None else []
        self.logger.info("\r\n************At {}, we validate {} scans.**************\r\n".format(self.epoch_n, len(scans)))
        if not scans:
            return {}
        was_training = self.model.training
        self.model.eval()
        val_time = AverageMeter()
        mine = []
        for scan_idx in range(ddist.rank(), len(scans), ddist.world_size()):
            pred, target, elapse = self.evaluate_scan(scans[scan_idx])
            mine.append((scan_idx, pred, target, elapse))
            self.logger.info(f"Validation step {scan_idx + 1}/{len(scans)}.")
        rows = ddist.all_gather_object(mine)
        rows = sorted(r for part in rows for r in part)
        for _, _, _, elapse in rows:
            val_time.update(elapse, 1)
        preds, targets = np.asarray([r[1] for r in rows]), np.asarray([r[2] for r in rows])
        v_metrics = {'val_time': val_time.avg, 'val_acc_reg_cls': float((preds == targets).mean())}
        self.logger.info(f"val_metrics: {v_metrics}")
        self.model.train(was_training)
        return v_metrics


class LesionSegTest(JobRunner, _ScanPipeline):
    """Full-CT inference runner (job_runner.py:814-1067) with the working constructor process_pipeline.py assumes
    (SURVEY D3): `LesionSegTest(input_image_path, input_lobe_path, output_path, settings, checkpoint)`.

    `run_scan` keeps one scan on the GPU from upload to mask download: lobe bounding boxes -> per-lobe crop / blank /
    window / ITK-style resample to the chunk grid (one kernel per lobe) -> the five chunks as ONE batch through the model
    (eval-mode BatchNorm is batch independent, so batching does not change results) -> RAM upsample + activation + paste
    under the lobe mask -> 8-bit histograms -> Otsu (256 bins, host) -> lesion masks.
    head='sigmoid' is LesionSegChunkTrain.evaluate_scan's path (job_runner.py:764-770); head='literal' reproduces
    LesionSegTest.run verbatim, which zeroes every heat map when out_ch == 1 (job_runner.py:988-1000)."""

    def __init__(self, input_image_path=None, input_lobe_path=None, output_path=None, settings_module=None,
                 checkpoint=None, task_name='test', head='sigmoid', build_model=True):
        super().__init__(None, settings_module)
        self.scan_path, self.lobe_path, self.output_path, self.task_name, self.head = \
            input_image_path, input_lobe_path, output_path, task_name, head
        self.crop_border = 5                                                 # mm, dataset.py crop_border
        if build_model:
            self.init()
            if checkpoint is not None:
                path = checkpoint if os.path.isabs(checkpoint) else os.path.join(self.exp_path, checkpoint)
                self.reload_model_from_cache(path)            # raises when the file is missing
            else:
                self.logger.warning("LesionSegTest built WITHOUT a checkpoint: the model keeps its random (HeNorm) "
                                    "initialisation — benchmark / test use only")
        self.model_eval = None

    def run_scans(self, items):
        """Pipelined full-CT inference over an iterable of (scan int16 [D,H,W], lobe uint8 [D,H,W], spacing) HOST tensors
        (pinned memory makes the copies asynchronous): the upload of scan i+1 (copy stream) and the download of the masks
        of scan i-1 (second copy stream, into two alternating pinned buffers) overlap the kernels of scan i, so a scan
        costs max(compute, H2D, D2H) instead of their sum.  Yields (lesion, lesion_post, ratio) per scan; the yielded host
        tensors are valid until the next result is requested."""
        cur = torch.cuda.current_stream()
        dev = torch.device("cuda", torch.cuda.current_device())
        if getattr(self, "_pipe", None) is None:     # streams, two device input slots, two pinned result slots: live with the runner
            self._pipe = {"up": torch.cuda.Stream(), "down": torch.cuda.Stream(), "out": [None, None],
                          "in": [None, None], "in_free": [None, None]}
        P = self._pipe
        up_s, down_s, bufs = P["up"], P["down"], P["out"]

        def upload(item, slot):
            """host scan -> persistent device slot (no per-scan allocation); the slot is reused only after the scan that
            read it last has finished"""
            scan_h, lobe_h = item[0], item[1]
            n = scan_h.numel()
            if P["in"][slot] is None or P["in"][slot][0].numel() < n:
                P["in"][slot] = (torch.empty(n, dtype=torch.int16, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
                P["in_free"][slot] = None
                cur.synchronize()                                      # fresh storage: visible to the copy stream
            s_d = P["in"][slot][0][:n].view(scan_h.shape)
            l_d = P["in"][slot][1][:n].view(lobe_h.shape)
            with torch.cuda.stream(up_s):
                if P["in_free"][slot] is not None:
                    up_s.wait_event(P["in_free"][slot])
                s_d.copy_(scan_h, non_blocking=True)
                l_d.copy_(lobe_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up_s)
            return s_d, l_d, ev

        def finish(p):
            p[3].synchronize()
            return p[0], p[1], float(p[2].item())

        it = iter(items)
        item = next(it, None)
        slot = 0
        up = upload(item, slot) if item is not None else None
        pending, k = None, 0
        while item is not None:
            (s_d, l_d, ev), spacing, this_slot = up, item[2], slot
            item = next(it, None)
            slot ^= 1
            up = upload(item, slot) if item is not None else None      # next scan's H2D runs under this scan's kernels
            cur.wait_event(ev)
            les, post, ratio = self.scan_to_masks(s_d, l_d, spacing)
            done = torch.cuda.Event()
            done.record(cur)
            P["in_free"][this_slot] = done
            if bufs[k] is None or bufs[k][0].shape != les.shape:
                bufs[k] = (torch.empty(les.shape, dtype=les.dtype).pin_memory(), torch.empty(post.shape, dtype=post.dtype).pin_memory(),
                           torch.empty((), dtype=torch.float32).pin_memory())
            with torch.cuda.stream(down_s):
                down_s.wait_event(done)
                for dst, src in zip(bufs[k], (les, post, ratio)):
                    dst.copy_(src, non_blocking=True)
                    src.record_stream(down_s)
                dl = torch.cuda.Event()
                dl.record(down_s)
            if pending is not None:
                yield finish(pending)
            pending, k = (bufs[k][0], bufs[k][1], bufs[k][2], dl), k ^ 1
        if pending is not None:
            yield finish(pending)

    def archive_results(self, heatmap, pred, post_pred, meta):
        """job_runner.py:857-890: `<out>/<task>/<uid>.mha` (lesion mask), `heatmap/<uid>.mha` (heat map windowed to uint8),
        `post/<uid>.mha` (post-processed mask), all uint8 with the scan's origin / direction / original spacing."""
        from utils import windowing, write_array_to_mha_itk
        output_path = os.path.join(self.output_path, self.task_name)
        post_path, heatmap_path = os.path.join(output_path, "post"), os.path.join(output_path, "heatmap")
        for d in (output_path, post_path, heatmap_path):
            os.makedirs(d, exist_ok=True)
        geo = dict(type=np.uint8, origin=list(meta["origin"])[::-1], spacing=list(meta["original_spacing"])[::-1],
                   direction=np.asarray(meta["direction"]).reshape(3, 3)[::-1].flatten().tolist())
        write_array_to_mha_itk(output_path, [pred.astype(np.uint8)], [meta["uid"]], **geo)
        write_array_to_mha_itk(heatmap_path, [windowing(heatmap, from_span=(0, 1)).astype(np.uint8)], [meta["uid"]], **geo)
        write_array_to_mha_itk(post_path, [post_pred.astype(np.uint8)], [meta["uid"]], **geo)

    def _process_mha(self, path):
        """One `<uid>.mha` scan + lobe mask -> the three archived .mha files; returns the record row."""
        from utils import read_mha
        uid = os.path.splitext(os.path.basename(path))[0]
        image, geo = read_mha(path)
        lobe, _ = read_mha(os.path.join(self.lobe_path, uid + ".mha"))
        start = time.time()
        dev = torch.device("cuda", torch.cuda.current_device())
        spacing = [float(v) for v in geo["spacing"][::-1]]                      # ITK x-y-z -> z-y-x
        scan_t = self.resample_to_working_grid(torch.from_numpy(image.astype(np.int16)).to(dev), spacing, "linear")
        lobe_t = self.resample_to_working_grid(torch.from_numpy(lobe.astype(np.uint8)).to(dev), spacing, "nearest")
        new_sp = [float(self.settings.TEST_RESAMPLE_SPACING)] * 3
        out = self.run_scan(scan_t, lobe_t, new_sp, return_device=True)
        back = lambda t, mode: ops_itk_back(t, image.shape, new_sp, spacing, mode).cpu().numpy()
        meta = {"uid": uid, "origin": geo["origin"][::-1], "original_spacing": spacing,
                "direction": np.asarray(geo["direction"]).reshape(3, 3)[::-1].flatten().tolist()}
        self.archive_results(back(out["heatmap"], "linear"), back(out["lesion"], "nearest"),
                             back(out["lesion_post"], "nearest"), meta)
        return {"uid": uid, "seconds": time.time() - start, "ratio": float(out["ratio"].item())}

    def run_mha(self):
        """The reference's file contract (job_runner.py:906-1067): every `<uid>.mha` scan under input_image_path with its
        lobe mask `<uid>.mha` under input_lobe_path -> archive_results().  MetaImage I/O is utils.read_mha / write_mha.
        This rank's shard of the sorted list only (dist.shard); a scan that fails is logged and skipped like the
        reference's per-scan try/except (job_runner.py:1056-1059)."""
        records = []
        for path in ddist.shard(sorted(glob.glob(os.path.join(self.scan_path, "*.mha")))):
            uid = os.path.splitext(os.path.basename(path))[0]
            if os.path.exists(os.path.join(self.output_path, self.task_name, uid + ".mha")):
                self.logger.warning("We have already archived results for scan %s", uid)
                continue
            try:
                records.append(self._process_mha(path))
            except Exception:                                                    # noqa: BLE001
                self.logger.exception("scan %s failed", uid)
                continue
            self.logger.info("Finished %s, in %.3f seconds.", uid, records[-1]["seconds"])
        return records

    def run(self):
        """Process every `<uid>.mha` (reference contract, see run_mha) or `<uid>.npz` (keys: image int16, lobe uint8,
        spacing[, vessel]) under input_image_path.  With torch.distributed initialised (torchrun) the sorted scan list is
        sharded `rank::world` — every scan is processed and written exactly once, no collective on the data path."""
        os.makedirs(os.path.join(self.output_path, self.task_name), exist_ok=True)
        records = self.run_mha() if self.lobe_path else []
        for path in ddist.shard(sorted(glob.glob(os.path.join(self.scan_path, "*.npz")))):
            uid = os.path.splitext(os.path.basename(path))[0]
            target = os.path.join(self.output_path, self.task_name, uid + ".npz")
            if os.path.exists(target):
                self.logger.warning("We have already archived results for scan %s", uid)
                continue
            data = np.load(path)
            lobe = np.load(os.path.join(self.lobe_path, uid + ".npz"))["lobe"] if self.lobe_path and "lobe" not in data else data["lobe"]
            start = time.time()
            dev = torch.device("cuda", torch.cuda.current_device())
            spacing = [float(s) for s in data["spacing"]]
            scan_t = self.resample_to_working_grid(torch.from_numpy(data["image"].astype(np.int16)).to(dev), spacing, "linear")
            lobe_t = self.resample_to_working_grid(torch.from_numpy(lobe.astype(np.uint8)).to(dev), spacing, "nearest")
            new_sp = [float(self.settings.TEST_RESAMPLE_SPACING)] * 3
            out = self.run_scan(scan_t, lobe_t, new_sp, return_device=True)
            back = lambda t, mode: ops_itk_back(t, data["image"].shape, new_sp, spacing, mode)
            np.savez_compressed(target, lesion=back(out["lesion"], "nearest").cpu().numpy(),
                                lesion_post=back(out["lesion_post"], "nearest").cpu().numpy(),
                                heatmap=back(out["heatmap"], "linear").cpu().numpy(), threshold=out["threshold"])
            records.append({"uid": uid, "seconds": time.time() - start, "ratio": float(out["ratio"].item())})
            self.logger.info("Finished %s, in %.3f seconds.", uid, records[-1]["seconds"])
        return records


def ops_itk_back(t, original_size, spacing, original_spacing, mode):
    """resample a working-grid volume back to the scan's original grid (job_runner.py:1017-1030)."""
    from dram_native import ops
    return ops.itk_resample(t.contiguous(), tuple(int(s) for s in original_size), mode, in_spacing=spacing,
                            out_spacing=original_spacing)
