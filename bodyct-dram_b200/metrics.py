"""Drop-in for the loss objects of the reference's `dram/metrics.py` that the two shipped settings files select
(`IntRegRefineLoss`, with `IntRegLoss` and `BootBinCrossEntropy` underneath; metrics.py:10-51,75-210,311-373).

Same call signature — the loss object calls `model(images, lobes)` itself (metrics.py:362) — and same arithmetic, but
the whole objective stays on the device: the reference's per-sample `.item()` syncs (metrics.py:124,163), the boolean
gather `probs[lobes > 0]` and the per-sample `.cpu().numpy()` pseudo-label loop (metrics.py:336-352) are replaced by
masked reductions, so a training step has no host round trip before `loss.item()` logging.
Under data parallelism the batch-global normalisers (metrics.py:30,37,42,48) are all-reduced so that the sum of the
ranks' losses equals the single-process loss on the global batch.
"""
import os
from itertools import permutations

import numpy as np
import torch

from dram_native import dist as ddist
from dram_native import functional as DF


class BootBinCrossEntropy:
    """Bootstrapped binary cross entropy (metrics.py:10-51) written with sums and counts instead of gathers."""

    def __init__(self, smoothing=0.1):
        self.smoothing = smoothing
        self.eps = 1e-7

    def __call__(self, p, t, voi, class_weights=None):
        assert t.size() == p.size()
        eps = self.eps
        inside = (voi > 0).to(p.dtype)
        outside = 1.0 - inside
        t = t.to(p.dtype)
        # counts that are batch-global in the reference -> all-reduced under data parallelism
        counts = torch.stack([outside.sum(), inside.sum(), (t * inside).sum()]).double()
        counts = ddist.all_reduce_(counts.clone()) if ddist.active() else counts
        n_out, n_in, t_in = counts[0], counts[1], counts[2]
        # part I, outside the volume of interest (metrics.py:24-30)
        pt = (p * t + (1.0 - p) * (1.0 - t)).clamp(eps, 1.0 - eps)
        nll = -torch.log(pt)
        bceo = (nll * outside).sum() / n_out.to(p.dtype)
        # part II, inside (metrics.py:33-49)
        alpha = (1.0 - t_in / n_in.clamp_min(1.0)).clamp(0.25, 0.75).to(p.dtype)
        w = (alpha * t + (1.0 - alpha) * (1.0 - t)) * inside
        w_sum = (alpha.double() * t_in + (1.0 - alpha.double()) * (n_in - t_in)).to(p.dtype)
        bce = (nll * w).sum() / w_sum.clamp_min(1e-30)
        t_hat = (p > 0.5).to(p.dtype)
        pt_hat = (p * t_hat + (1.0 - p) * (1.0 - t_hat)).clamp(eps, 1.0 - eps)
        boot = (-torch.log(pt_hat) * inside).sum() / n_in.clamp_min(1.0).to(p.dtype)
        has_inside = (n_in > 0).to(p.dtype)
        return bceo + has_inside * ((1.0 - self.smoothing) * bce + self.smoothing * boot)


class IntRegLoss:
    """Interval regression on the predicted lesion ratio per lobe chunk (metrics.py:75-210)."""
    ctss_ratio_map = {0: (0.0, 0.001), 1: (0.001, 0.01), 2: (0.01, 0.05), 3: (0.05, 0.35), 4: (0.35, 0.5), 5: (0.5, 1.00001)}
    ctss_frequency_map = {0: 0.3, 1: 0.25, 2: 0.23, 3: 0.2, 4: 0.18, 5: 0.15}

    def __init__(self, band_width=5e-2):
        self.band_width = band_width
        self.trace = False
        self.qwk = self.gen_qwk(6)

    def gen_qwk(self, N):
        qwk = np.zeros((N, N), dtype=np.float32)
        for i, j in permutations(range(N), 2):
            qwk[i, j] = (abs(i - j) + 1) ** 2 / (N ** 2)
        return qwk + 1.0

    def ratio_to_label(self, ratios, ratio_map=None):
        ratio_map = self.ctss_ratio_map if ratio_map is None else ratio_map
        labels = []
        for ratio in ratios:
            r = float(ratio)
            labels.append([k for k, (lo, hi) in ratio_map.items() if lo <= r < hi][0])
        return labels

    def get_one_label(self, ctss):
        return self.ctss_ratio_map[int(float(ctss))]

    def label_tensors(self, ctsses, freq_map, device):
        """Everything the loss derives from the per-chunk CT severity scores, as small device tensors.  Built OUTSIDE a
        captured CUDA graph (they come from Python lists) and handed to __call__ through `label_tensors=`."""
        band = torch.tensor([self.ctss_ratio_map[int(float(c))] for c in ctsses], dtype=torch.float64)
        w = torch.tensor([freq_map[int(float(c))] for c in ctsses], dtype=torch.float32).clamp(0.2, 0.8)
        keep = torch.tensor([0.0 if float(c) < 1e-7 else 1.0 for c in ctsses], dtype=torch.float32)
        return {"band": band.to(device), "w": w.to(device), "keep": keep.to(device)}

    def get_labels(self, ctsses, lesion_ps, band=None):
        """metrics.py:121-137 on the device, in float64 like the Python-float arithmetic it replaces, no `.item()`."""
        p = lesion_ps.reshape(-1).double()
        if band is None:
            band = torch.tensor([self.ctss_ratio_map[int(float(c))] for c in ctsses], dtype=torch.float64, device=p.device)
        clb, cub = band[:, 0], band[:, 1]
        lb, ub = (p - self.band_width).clamp_min(0.0), (p + self.band_width).clamp_max(1.0)
        lo, hi = torch.maximum(clb, lb), torch.minimum(cub, ub)
        empty = hi < lo
        below = empty & (ub <= clb)
        lo = torch.where(below, lb, torch.where(empty, clb, lo))
        hi = torch.where(below, ub, torch.where(empty, cub, hi))
        return torch.stack([lo, hi], dim=1).float()

    def _reg_loss(self, values, lobes, lesion_candidates, ctsses, use_sigmoid, **kwargs):
        labels = kwargs.get("label_tensors") or self.label_tensors(ctsses, kwargs.get('obj').ctss_frequency_map, values.device)
        if os.environ.get("DRAM_FUSED_LOSS", "1") == "1":
            return DF.IntRegHinge.apply(values, lobes, lesion_candidates, labels["band"], labels["w"], float(self.band_width),
                                        bool(use_sigmoid))
        with torch.no_grad():
            rub, _ = DF.MaskedMean.apply(lesion_candidates, lobes, False, False)     # sum(lesion*lobe)/sum(lobe)
        pred_ratio, _ = DF.MaskedMean.apply(values, lobes, use_sigmoid, True)       # mean of probs over lobe > 0
        tg = self.get_labels(ctsses, rub, labels["band"])
        K = (0.5 * (tg[:, 1] - tg[:, 0])) ** 2
        loss_unhinge = (pred_ratio - (tg[:, 1] + tg[:, 0]) / 2.0) ** 2 - K
        loss_unweight = torch.clamp_min(loss_unhinge, 0.0)
        return (loss_unweight / labels["w"]).sum()

    def compute_reg_loss_with_probs(self, probs, lobes, lesion_candidates, ctsses, **kwargs):
        """metrics.py:158-177 (same signature: takes sigmoid(RAM))."""
        return self._reg_loss(probs, lobes, lesion_candidates, ctsses, False, **kwargs)

    def compute_reg_loss_with_logits(self, dense_outs, lobes, lesion_candidates, ctsses, **kwargs):
        """Same loss from the RAM logits: sigmoid + masked mean run as ONE fused reduction kernel (fwd and bwd)."""
        return self._reg_loss(dense_outs, lobes, lesion_candidates, ctsses, True, **kwargs)

    def before_call(self, model, **kwargs):
        obj = kwargs.get('obj')
        base = os.path.join(getattr(obj, 'debug_path', '.'), "model_detailed_trace")
        model.trace_path = (os.path.join(base, f"epoch_{getattr(obj, 'epoch_n', 0)}"), kwargs.get('metas'))


class IntRegRefineLoss(IntRegLoss):
    """reg loss on the RAM + bootstrapped BCE of the refined RAM against thresholded pseudo labels (metrics.py:311-373)."""

    def __init__(self, band_width=1e-2, smoothing=0.1, refine_method='th', config_param={}):
        super().__init__(band_width)
        self.smoothing = smoothing
        self.refine_method = refine_method
        self.config_param = config_param
        self.bootstrap_loss = BootBinCrossEntropy(smoothing)

    def pseudo_labels(self, dense_outs, lobes, lesions, ctsses, keep=None):
        """metrics.py:333-354 + threshold_postprocessing :325-329, without leaving the device."""
        if self.refine_method != 'th':
            raise NotImplementedError(f"Do not support refine method :{self.refine_method}!")
        with torch.no_grad():
            pred = (torch.sigmoid(dense_outs) > 0.5) & (lobes != 0)
            if keep is None:
                keep = torch.tensor([0.0 if float(c) < 1e-7 else 1.0 for c in ctsses], dtype=dense_outs.dtype,
                                    device=dense_outs.device)
            return (pred & (lesions > 0)).to(dense_outs.dtype) * keep.view(-1, 1, 1, 1, 1)

    def compute_seg_loss(self, dense_outs, refined_dense_outs, images, lobes, lesions, scores, metas, obj, tag='fixed',
                         keep=None):
        if self.refine_method != 'th':
            raise NotImplementedError(f"Do not support refine method :{self.refine_method}!")
        if keep is None:
            keep = torch.tensor([0.0 if float(c) < 1e-7 else 1.0 for c in scores], dtype=torch.float32,
                                device=dense_outs.device)
        if os.environ.get("DRAM_FUSED_LOSS", "1") == "1":
            # pseudo labels + bootstrapped BCE as one reduction kernel and one gradient kernel
            return DF.BootBce.apply(dense_outs.detach(), refined_dense_outs, lobes, lesions, keep,
                                    float(self.bootstrap_loss.smoothing), float(self.bootstrap_loss.eps))
        t = self.pseudo_labels(dense_outs, lobes, lesions, scores, keep)
        return self.bootstrap_loss(torch.sigmoid(refined_dense_outs), t, lobes > 0)

    def __call__(self, model, images, lobes, lesions, ctsses, **kwargs):
        self.before_call(model, **kwargs)
        dense_outs, refined_dense_outs = model(images, lobes)
        reg_loss = self.compute_reg_loss_with_logits(dense_outs, lobes, lesions, ctsses, **kwargs)
        labels = kwargs.get("label_tensors")
        seg_loss = self.compute_seg_loss(dense_outs, refined_dense_outs, images, lobes, lesions, ctsses,
                                         kwargs.get("metas"), kwargs.get("obj"), keep=labels["keep"] if labels else None)
        return reg_loss, seg_loss
