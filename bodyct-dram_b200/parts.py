"""Drop-in for the reference's `dram/parts.py` building blocks, executing on hand-written sm_100a kernels.

Same class names, constructor arguments, sub-module names (`conv_blocks.<i>.{0: Conv3d, 1: BatchNorm3d, 2: ReLU}`,
`maxpool`, `upsample`) and therefore the same `state_dict` keys and tensor shapes as the reference
(/root/reference/dram/parts.py:66-196), so checkpoints load either way.  The nn.Conv3d / nn.BatchNorm3d children
only HOLD parameters and buffers: `forward` never calls them, it hands their tensors to the fused
[conv -> batch-norm -> ReLU (-> max-pool)] kernels in `dram_native`.  Inputs/outputs are [N, C, D, H, W] fp32 CUDA
tensors (any memory layout in, channels-last memory out).  Unsupported options raise; nothing falls back to PyTorch.
"""
import numpy as np
import torch
import torch.nn as nn

from dram_native import functional as DF
from dram_native import ops


class Identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x, args=None):
        return x


def normal_wrapper(normal_method, in_ch, in_ch_div=2):
    """parts.py:17-35.  Only batch statistics normalisation ('bn', and 'sbn' = the same kernels with the statistics
    all-reduced, see dram_native.dist) is on the B200 path."""
    if normal_method in ("bn", "sbn"):
        return nn.BatchNorm3d(in_ch)
    if normal_method is None:
        return Identity()
    raise NotImplementedError(f"norm method {normal_method!r} is not implemented on the B200 path (bn | sbn)")


def act_wrapper(act_method, num_parameters=1, init=0.25):
    """parts.py:48-54 (the reference's blocks always end up with ReLU: `act_methpd` typo)."""
    if act_method == "relu":
        return nn.ReLU(inplace=True)
    raise NotImplementedError(f"activation {act_method!r} is not implemented on the B200 path (relu)")


def crop_concat_5d(t1, t2):
    """parts.py:37-46: centre-crop t2 to t1's spatial size (ceil offsets) and concatenate on channels.
    Stand-alone helper kept for API parity; the decoder block fuses this into its upsample kernel."""
    assert t1.dim() == t2.dim() == 5
    assert t1.shape[-1] <= t2.shape[-1]
    sl = (slice(None), slice(None)) + tuple(
        slice(int(np.ceil((b - a) / 2)), a + int(np.ceil((b - a) / 2))) for a, b in zip(t1.shape[2:], t2.shape[2:]))
    return torch.cat([t1, t2[sl]], dim=1)


def _as_list(v, n):
    return list(v) if isinstance(v, (tuple, list)) else [v] * n


def _make_units(in_chs, out_chs, ksizes, pads, strides, conv_bias, norm_method, dropout):
    if dropout and dropout > 0:
        raise NotImplementedError("dropout > 0 is not implemented on the B200 path (both reference settings use 0.0)")
    units = []
    for i, (ci, co) in enumerate(zip(in_chs, out_chs)):
        k, p, s = ksizes[i], pads[i], strides[i]
        if s != 1 or k not in (1, 3) or p != k // 2:
            raise NotImplementedError(f"conv k={k} pad={p} stride={s} is not on the B200 path (k in {{1,3}}, pad=k//2, stride 1)")
        units.append(nn.Sequential(nn.Conv3d(ci, co, kernel_size=k, padding=p, bias=conv_bias, stride=s),
                                   normal_wrapper(norm_method, co), act_wrapper("relu")))
    return nn.Sequential(*units)


class _FusedUnits(nn.Module):
    """Shared execution of a stack of [conv, norm, relu] units through the fused kernels.

    Between units (and, with `planes=True`, towards the caller) activations travel as `DF.Act`: bf16 split planes, the
    tensor-core convolutions' operand format, with a storage-less autograd handle.  The public `forward`s return plain
    fp32 tensors; the models in this package pass `planes=True` and keep everything in planes."""

    def _unit_args(self, i):
        conv, bn = self.conv_blocks[i][0], self.conv_blocks[i][1]
        if not isinstance(bn, nn.BatchNorm3d):
            raise NotImplementedError("conv units without batch-norm are not on the B200 path")
        return conv, bn, (bn.training or bn.running_mean is None)

    def _count(self, bn, stat_updates):
        if bn.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += stat_updates

    def _run(self, x, stat_updates=1, pool_last=False, planes=False, skip_last=False, only_last=False):
        """skip_last: stop before the last unit (the caller fuses it with the RAM head); only_last: run just that unit."""
        n = len(self.conv_blocks)
        pooled = None
        for i in range(n - 1 if only_last else 0, n - 1 if skip_last else n):
            conv, bn, use_batch_stats = self._unit_args(i)
            pool = pool_last and i == n - 1
            out = DF.conv_bn_relu(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                  use_batch_stats, bn.momentum, bn.eps, stat_updates, pool,
                                  out_planes=planes or i < n - 1)
            self._count(bn, stat_updates)
            if pool:
                x, pooled = out
            else:
                x = out
        if not planes:
            x, pooled = DF.as_tensor(x), (DF.as_tensor(pooled) if pooled is not None else None)
        return (x, pooled) if pool_last else x

    def run_last_with_ram(self, x, top_layer, stat_updates=1):
        """Last unit + `top_layer` (the RAM reduce, models.py:145) as one fused op, or None if the shapes do not allow it."""
        conv, bn, use_batch_stats = self._unit_args(len(self.conv_blocks) - 1)
        if not DF.ram_fusable(conv.weight, top_layer.weight, conv.bias):
            return None
        ram = DF.conv_bn_relu_ram(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch_stats,
                                  bn.momentum, bn.eps, stat_updates, top_layer.weight, top_layer.bias)
        self._count(bn, stat_updates)
        return ram


class ConvBlock5d(_FusedUnits):
    """2 x [Conv3d -> BatchNorm3d -> ReLU]  (parts.py:66-113)."""

    def __init__(self, in_chs, base_chs, checkpoint_segments, conv_ksize, conv_bias, conv_pad, dropout=0.1,
                 conv_strides=1, norm_method='bn', act_methpd='relu', lite=False, **kwargs):
        super().__init__()
        if lite:
            raise NotImplementedError("lite blocks (no normalisation) are not on the B200 path")
        n = len(in_chs)
        self.checkpoint_segments = checkpoint_segments
        self.conv_blocks = _make_units(in_chs, base_chs, _as_list(conv_ksize, n), _as_list(conv_pad, n),
                                       _as_list(conv_strides, n), conv_bias, norm_method, dropout)

    def forward(self, x, args=None, stat_updates=1, planes=False):
        return self._run(x, stat_updates, planes=planes)


class UpsampleConvBlock5d(_FusedUnits):
    """trilinear x2 (align_corners=True) -> crop_concat([up, skip]) -> 2 x [Conv3d -> BN -> ReLU]  (parts.py:116-155)."""

    def __init__(self, in_chs, base_chs, checkpoint_segments, scale_factor, conv_ksize, conv_bias, conv_pad, dropout=0.1,
                 norm_method='bn', act_methpd='relu', **kwargs):
        super().__init__()
        n = len(in_chs)
        self.checkpoint_segments = checkpoint_segments
        self.scale_factor = scale_factor
        sf = tuple(scale_factor) if isinstance(scale_factor, (tuple, list)) else (scale_factor,) * 3
        if any(int(s) != 2 for s in sf):
            raise NotImplementedError(f"upsample scale factor {scale_factor} is not on the B200 path (2)")
        self.conv_blocks = _make_units(in_chs, base_chs, _as_list(conv_ksize, n), _as_list(conv_pad, n), [1] * n,
                                       conv_bias, norm_method, dropout)
        self.merge_func = kwargs.get('merge_func', crop_concat_5d)
        self.upsample = nn.Upsample(size=None, scale_factor=self.scale_factor, mode='trilinear', align_corners=True)

    def forward(self, inputs, cats, args=None, stat_updates=1, planes=False, skip_last=False):
        x = DF.upsample_concat(inputs, cats)
        return self._run(x, stat_updates, planes=planes, skip_last=skip_last)


class ConvPoolBlock5d(_FusedUnits):
    """2 x [Conv3d -> BN -> ReLU] then MaxPool3d(2,2,0); returns (pre_pool, pooled)  (parts.py:157-196)."""

    def __init__(self, in_ch_list, base_ch_list, checkpoint_segments, conv_ksize, conv_bias, conv_pad, pool_ksize,
                 pool_strides, pool_pad, dropout=0.1, conv_strdes=1, norm_method='bn', act_method="relu", **kwargs):
        super().__init__()
        n = len(in_ch_list)
        self.checkpoint_segments = checkpoint_segments
        if (pool_ksize, pool_strides, pool_pad) != (2, 2, 0):
            raise NotImplementedError("only MaxPool3d(2, 2, 0) is on the B200 path")
        self.conv_blocks = _make_units(in_ch_list, base_ch_list, _as_list(conv_ksize, n), _as_list(conv_pad, n),
                                       _as_list(conv_strdes, n), conv_bias, norm_method, dropout)
        self.maxpool = nn.MaxPool3d(kernel_size=pool_ksize, stride=pool_strides, padding=pool_pad)

    def forward(self, x, args=None, stat_updates=1, planes=False):
        y, pooled = self._run(x, stat_updates, pool_last=True, planes=planes)
        return y, pooled
