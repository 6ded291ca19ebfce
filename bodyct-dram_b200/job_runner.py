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
        self.optimizer = opt_cls(self.model.parameters(), **opt_cfg)
        self.loss_func = get_callable_by_name(loss_cfg.pop('method'))(**loss_cfg)
        self.scheduler = get_callable_by_name(sched_cfg.pop('method'))(self.optimizer, **sched_cfg)
        self.reducer = ddist.GradReducer(self.model.parameters()) if ddist.active() else None
        if ddist.active():                                   # replicas start from rank 0's weights
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                torch.distributed.broadcast(t.data, src=0)

    # -------------------------------------------------------------------------------------------- checkpoints
    def reload_state(self, model_state):
        """Tolerant reload (job_runner.py:477-497): keep only keys that exist with the same shape."""
        own = self.model.state_dict()
        kept = {k: v for k, v in model_state.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        own.update(kept)
        self.model.load_state_dict(own)
        return sorted(set(model_state) - set(kept))

    def reload_model_from_cache(self, path=None):
        path = path or getattr(self.settings, "RELOAD_CHECKPOINT_PATH", None)
        if path is None:
            cands = sorted(glob.glob(os.path.join(self.exp_path, "*.pth")), key=os.path.getmtime)
            if not cands:
                return False
            path = cands[-1]
        state = torch.load(path, map_location="cpu")
        self.saved_model_states = state
        if "model" in self.settings.RELOAD_DICT_LIST:
            self.reload_state(state.get("model", state.get("model_dict", {})))
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


class LesionSegChunkTrain(JobRunner):
    """Training runner: `train_step(batch)` is the body of the reference's hot loop (job_runner.py:657-674)."""

    def __init__(self, settings_module=None, setting_module_file_path=None):
        super().__init__(setting_module_file_path, settings_module)
        self.init()

    def train_step(self, batch_data):
        """batch_data: {"#image", "#lobe_reference", "#pseudo_lesion_reference": [B,D,H,W] host or device tensors,
        "meta": {"cle": [...]}} -> (loss tensor, loss tuple).  Includes the H2D copy, backward, gradient all-reduce
        (data parallel) and the optimizer step; does NOT sync with the host."""
        self.model.train()
        dev = torch.device("cuda", torch.cuda.current_device())
        images = batch_data["#image"].to(dev, torch.float32, non_blocking=True).unsqueeze(1)
        lobes = batch_data["#lobe_reference"].to(dev, torch.float32, non_blocking=True).unsqueeze(1)
        lesions = batch_data["#pseudo_lesion_reference"].to(dev, torch.float32, non_blocking=True).unsqueeze(1)
        metas = batch_data["meta"]
        self.optimizer.zero_grad(set_to_none=True)
        loss_tuple = self.loss_func(self.model, images, lobes, lesions, metas["cle"], obj=self, metas=metas)
        factors = self.settings.LOSS_FACTORS[:len(loss_tuple)]
        loss = torch.stack([l * w for l, w in zip(loss_tuple, factors)]).sum()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        self.current_iteration += 1
        return loss, loss_tuple

    def train(self, loader):
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
