"""Drop-in for the reference's `dram/models.py` (DC3D, DC3DATGeneric, PCM, HeNorm, pooling_dense_features).

Same constructor kwargs (the `MODEL` dict of exp_settings/*.py), same sub-module and `state_dict` names, same
`forward(x, lungs=None) -> (dense_outs, refined_dense_outs)` contract as /root/reference/dram/models.py, but every
device op runs in libdram_b200.so (tcgen05 implicit-GEMM convolutions, fused BN/ReLU/pool, RAM reduce, stencil
attention).  No DGL, no graph object: `PCM` computes the same message passing as address arithmetic on the voxel grid.
Gradient checkpointing (`checkpoint_layers`) is not needed on a 180 GB part and is not performed; its one observable
side effect — BatchNorm running statistics updated twice per step in checkpointed blocks — is reproduced.
"""
import torch
import torch.nn as nn

from parts import ConvBlock5d, ConvPoolBlock5d, Identity, UpsampleConvBlock5d
from dram_native import functional as DF
from dram_native import ops


class Initializer:
    def initialize(self, module):
        raise NotImplementedError("need subclassing to implement.")


class HeNorm(Initializer):
    """models.py:17-35: He-normal conv weights (fan_in by default), conv bias 0.01, norm weight 1 / bias 0,
    nn.Linear default reset."""

    def __init__(self, **kwargs):
        self.mode = kwargs.get('mode', 'fan_in')

    def initialize(self, module):
        def init_weights(m):
            if type(m) in (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d):
                nn.init.kaiming_normal_(m.weight, mode=self.mode)
                if m.bias is not None:
                    m.bias.data.fill_(0.01)
            elif isinstance(m, (nn.BatchNorm3d, nn.GroupNorm)):
                m.weight.data.fill_(1)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.reset_parameters()

        module.apply(init_weights)


def pooling_dense_features(dense_outs, lungs, pooling_method='avg'):
    """models.py:37-49 — per-lobe masked average of the RAM: sum(dense * lobe) / sum(lobe) per (sample, channel)."""
    B, C = dense_outs.shape[0], dense_outs.shape[1]
    if pooling_method == 'global_max':
        # models.py:42-43 (F.adaptive_max_pool3d(dense_outs, 1)); no shipped settings file selects it: a plain reduction
        # over the channels-last volume, gradient to the arg-max voxel like the reference's
        return torch.amax(dense_outs.reshape(B, C, -1), dim=2)
    if pooling_method == 'global_avg':
        lungs = torch.ones_like(dense_outs[:, :1])
    x = ops.to_ncdhw(dense_outs).reshape(B * C, -1)
    m = lungs.expand(B, C, *lungs.shape[2:]).reshape(B * C, -1) if C > 1 else lungs.reshape(B, -1)
    mean, _ = DF.MaskedMean.apply(x, m.contiguous().float(), False, False)
    return mean.view(B, C)


def _stat_updates(module, flag):
    """A torch.utils.checkpoint-wrapped block runs its forward twice per training step (models.py:123-143)."""
    return 2 if (flag > 0 and module.training and torch.is_grad_enabled()) else 1


class _UNet(nn.Module):
    """Shared encoder/decoder of DC3D and DC3DATGeneric (models.py:79-110 / 453-481)."""

    def _build_unet(self, n_layers, in_ch_list, base_ch_list, end_ch_list, out_ch, padding_list, checkpoint_layers,
                    dropout, upsample_sf, kernel_sizes, stacking, norm_method, act_method, always_decoder):
        assert len(end_ch_list) == len(base_ch_list) == len(in_ch_list) == len(padding_list)
        conv_bias = norm_method is None
        n = n_layers
        self.ds_modules = nn.ModuleList([
            ConvPoolBlock5d([in_ch_list[i], base_ch_list[i]], [base_ch_list[i], end_ch_list[i]], checkpoint_layers[i],
                            kernel_sizes[i], conv_bias, padding_list[i], 2, 2, 0, norm_method=norm_method,
                            act_method=act_method, dropout=dropout) for i in range(n)])
        self.bg = ConvBlock5d([in_ch_list[n], base_ch_list[n]], [base_ch_list[n], end_ch_list[n]], checkpoint_layers[n],
                              kernel_sizes[n], conv_bias, padding_list[n], dropout, norm_method=norm_method,
                              act_method=act_method)
        if always_decoder or (n + 1) < len(in_ch_list):
            self.us_modules = nn.ModuleList([
                UpsampleConvBlock5d([in_ch_list[n + 1 + i], base_ch_list[n + 1 + i]],
                                    [base_ch_list[n + 1 + i], end_ch_list[n + 1 + i]], checkpoint_layers[n + 1 + i],
                                    upsample_sf, kernel_sizes[n + 1 + i], conv_bias, padding_list[n + 1 + i],
                                    norm_method=norm_method, act_method=act_method, dropout=dropout) for i in range(n)])
        else:
            self.us_modules = None
        self.top_layer = nn.Conv3d(end_ch_list[n + stacking], out_ch, kernel_size=1, padding=0)

    def _encode_decode(self, x, decoder_ckpt_offset, tap=None, size=None):
        """Encoder / bottleneck / decoder + RAM head -> dense_outs (models.py:123-146 / 563-588).  `tap(idx, features)` is
        called with pre-pool encoder / bottleneck / decoder features under the reference's layer numbering
        (models.py:563-585).  Activations stay in split planes (`DF.Act`) from the first unit to the RAM head; the last
        decoder unit is fused with `top_layer` unless its features are tapped."""
        n = self.n_layers
        planes = DF.planes_enabled()
        skips = []
        h = x
        for idx, ds in enumerate(self.ds_modules):
            y, h = ds(h, stat_updates=_stat_updates(self, self.checkpoint_layers[idx]), planes=planes)
            skips.append(y)
            if tap is not None:
                tap(idx, y)
        h = self.bg(h, stat_updates=_stat_updates(self, self.checkpoint_layers[n]), planes=planes)
        if tap is not None:
            tap(n, h)
        dense = None
        if self.us_modules is not None:
            n_dec = min(len(self.us_modules), self.stacking)
            for idx, (us, skip) in enumerate(zip(self.us_modules, reversed(skips))):
                if self.stacking == idx:
                    break
                su = _stat_updates(self, self.checkpoint_layers[decoder_ckpt_offset + idx])
                tapped = tap is not None and (n + idx + 1) in getattr(self, "at_layers", ())
                if planes and idx == n_dec - 1 and not tapped:
                    h = us(h, skip, stat_updates=su, planes=True, skip_last=True)
                    dense = us.run_last_with_ram(h, self.top_layer, su)
                    if dense is None:
                        h = us._run(h, su, planes=True, only_last=True)
                else:
                    h = us(h, skip, stat_updates=su, planes=planes)
                if tap is not None and dense is None:
                    tap(n + idx + 1, h)
        if dense is None:
            dense = DF.RamReduce.apply(DF.as_tensor(h), self.top_layer.weight, self.top_layer.bias)        # models.py:145
        return DF.TrilinearResize.apply(dense, tuple(size))                                                   # models.py:146 / 588

    def init(self, initializer):
        initializer.initialize(self)

    def pooling_dense_features(self, dense_outs, lungs, pooling_method='avg'):
        return pooling_dense_features(dense_outs, lungs, pooling_method)


class DC3D(_UNet):
    """3-level 3-D U-Net + dense RAM head (models.py:52-147)."""

    def __init__(self, n_layers, in_ch_list, base_ch_list, end_ch_list, out_ch, padding_list, checkpoint_layers, dropout,
                 upsample_ksize=3, upsample_sf=2, kernel_sizes=None, stacking=0, norm_method="bn", act_method='relu',
                 pooling_method='avg', out_cls_ch=6):
        super().__init__()
        self.dropout, self.n_layers, self.padding_list = dropout, n_layers, padding_list
        self.in_ch_list, self.base_ch_list, self.end_ch_list = in_ch_list, base_ch_list, end_ch_list
        self.kernel_sizes = [3] * (n_layers * 2 + 1) if kernel_sizes is None else kernel_sizes
        self.upsample_ksize, self.upsample_sf = upsample_ksize, upsample_sf
        self.checkpoint_layers, self.norm_method = checkpoint_layers, norm_method
        self.out_ch, self.stacking, self.out_cls_ch, self.pooling_method = out_ch, stacking, out_cls_ch, pooling_method
        self._build_unet(n_layers, in_ch_list, base_ch_list, end_ch_list, out_ch, padding_list, checkpoint_layers, dropout,
                         upsample_sf, self.kernel_sizes, stacking, norm_method, act_method, always_decoder=False)
        self.dummy = torch.ones(1, requires_grad=True)
        self.trace_path = None

    def forward(self, x, lungs=None):
        x = ops.to_cl(x, "DC3D input")
        dense_outs = self._encode_decode(x, decoder_ckpt_offset=self.n_layers, size=x.shape[-3:])   # models.py:140 indexing
        return dense_outs, dense_outs


class PCM(nn.Module):
    """Graph attention over the voxel grid (models.py:150-411) without a graph: an in-grid-neighbour stencil attention.

    Parameters keep the reference names/shapes: theta, phi: Linear(in_ch, f_dim); G: Linear(g_ch, g_dim);
    r: Linear(g_dim, g_ch).  Supported on the B200 path: k_size 3, connectivity 1..3, self_loop on/off, merge types
    sm | smrelu | scaled_dot_product | scaled_dot_product_relu | smscaled, p_enc_dim 0, g_ch 1."""

    def __init__(self, pool_size, in_ch, g_ch, f_dim, geo_f_dim, g_dim, non_local_iter, k_size, merge_type='l2',
                 self_loop=True, connectivity=2, residual=False, p_enc_dim=32):
        super().__init__()
        self.in_ch, self.g_ch, self.f_dim, self.g_dim = in_ch, g_ch, f_dim, g_dim
        self.pool_size, self.merge_type, self.self_loop = pool_size, merge_type, self_loop
        self.non_local_iter, self.k_size, self.connectivity = non_local_iter, k_size, connectivity
        self.residual, self.p_enc_dim, self.geo_f_dim = residual, p_enc_dim, geo_f_dim
        if self.g_dim > 0:
            self.G = nn.Linear(g_ch, g_dim)
            self.r = nn.Linear(g_dim, g_ch)
        else:
            self.G, self.r, self.g_dim = Identity(), Identity(), g_ch
        if f_dim > 0:
            self.theta = nn.Linear(in_ch, f_dim)
            self.phi = nn.Linear(in_ch, f_dim)
        else:
            self.theta, self.phi, self.f_dim = Identity(), Identity(), in_ch
        if self.p_enc_dim > 0:
            if geo_f_dim > 0:
                self.geo_theta = nn.Linear(p_enc_dim, geo_f_dim)
                self.geo_phi = nn.Linear(p_enc_dim, geo_f_dim)
            else:
                self.geo_theta, self.geo_phi, self.geo_f_dim = Identity(), Identity(), p_enc_dim
        self.graph = None          # kept for attribute parity; never built

    def _check(self):
        if self.merge_type not in ops.MERGE_FLAGS:
            raise NotImplementedError(f"merge type {self.merge_type!r} is not on the B200 path {sorted(ops.MERGE_FLAGS)}")
        if self.k_size != 3 or self.p_enc_dim > 0 or self.g_ch != 1 or not isinstance(self.theta, nn.Linear):
            raise NotImplementedError("PCM on the B200 path needs k_size=3, p_enc_dim=0, g_ch=1 and f_dim>0")
        if self.f_dim not in (4, 8, 16):
            raise NotImplementedError(f"PCM on the B200 path is built for f_dim 4 | 8 | 16 (got {self.f_dim})")

    def forward(self, cam, f, args=None):
        self._check()
        flags = ops.MERGE_FLAGS[self.merge_type]
        if isinstance(self.G, nn.Linear):
            # r(sum_o a_o G(cam_o)) with sum_o a_o = 1  ==  coef * sum_o a_o cam_o + const   (scalars, autograd-tracked)
            coef = (self.r.weight @ self.G.weight).reshape(())
            const = (self.r.weight @ self.G.bias + self.r.bias).reshape(())
        else:
            coef = const = None
        for _ in range(self.non_local_iter):
            s = DF.PcmAttend.apply(cam, f, self.theta.weight, self.theta.bias, self.phi.weight, self.phi.bias,
                                   self.connectivity, bool(self.self_loop), flags, torch.is_grad_enabled())
            refined = s * coef + const if coef is not None else s
            cam = refined + cam if self.residual else refined
        return cam


class DC3DATGeneric(_UNet):
    """U-Net + RAM head + attention refinement of the RAM (models.py:415-597)."""

    def __init__(self, n_layers, in_ch_list, base_ch_list, end_ch_list, out_ch, padding_list, checkpoint_layers, dropout,
                 at_spatial_size, at_f_dim, at_g_dim, at_p_enc_dim, at_geo_f_dim, at_g_iter, at_k_size, at_merge_type,
                 at_self_loop, at_layers, upsample_ksize=3, upsample_sf=2, kernel_sizes=None, stacking=3,
                 norm_method="bn", act_method='relu', pooling_method='avg', out_cls_ch=6):
        super().__init__()
        self.dropout, self.n_layers, self.padding_list = dropout, n_layers, padding_list
        self.in_ch_list, self.base_ch_list, self.end_ch_list = in_ch_list, base_ch_list, end_ch_list
        self.at_spatial_size, self.out_cls_ch = at_spatial_size, out_cls_ch
        self.kernel_sizes = [3] * (n_layers * 2 + 1) if kernel_sizes is None else kernel_sizes
        self.upsample_ksize, self.upsample_sf = upsample_ksize, upsample_sf
        self.checkpoint_layers, self.norm_method = checkpoint_layers, norm_method
        self.out_ch, self.stacking, self.pooling_method = out_ch, stacking, pooling_method
        self.at_f_dim, self.at_g_dim, self.at_g_iter, self.at_k_size = at_f_dim, at_g_dim, at_g_iter, at_k_size
        self.at_p_enc_dim, self.at_geo_f_dim = at_p_enc_dim, at_geo_f_dim
        self.at_merge_type, self.at_self_loop, self.at_layers = at_merge_type, at_self_loop, at_layers
        self._build_unet(n_layers, in_ch_list, base_ch_list, end_ch_list, out_ch, padding_list, checkpoint_layers, dropout,
                         upsample_sf, self.kernel_sizes, stacking, norm_method, act_method, always_decoder=True)
        taps = [s for s in at_layers if s != -1]
        n_at_in_ch = at_f_dim * len(taps) + (1 if -1 in at_layers else 0)
        self.reshape = nn.ModuleList([
            nn.Sequential(nn.Conv3d(end_ch_list[l], at_f_dim, kernel_size=1, padding=0, stride=1),
                          nn.BatchNorm3d(at_f_dim), nn.ReLU(inplace=True)) for l in taps])
        self.attention_module = PCM(at_spatial_size, n_at_in_ch, out_ch, at_f_dim, at_geo_f_dim, at_g_dim, at_g_iter,
                                    at_k_size, at_merge_type, at_self_loop, p_enc_dim=at_p_enc_dim)
        self.dummy = torch.ones(1, requires_grad=True)
        self.trace_path = None       # assigned by the loss every call (metrics.py:202); debug dumps are not produced
        self.n_pcm_layer = 0

    def _reshape_head(self, nc, feats):
        conv, bn = self.reshape[nc][0], self.reshape[nc][1]
        out = DF.conv_bn_relu(feats.detach(), conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                              bn.running_var, bn.training, bn.momentum, bn.eps, 1, False)           # models.py:564
        if bn.training:
            bn.num_batches_tracked += 1
        return out

    def apply_attention(self, x, lungs, dense_out, attention_features):
        """models.py:510-518: RAM -> attention grid -> PCM -> back to the RAM grid (trilinear, align_corners=True)."""
        raw = tuple(dense_out.shape[2:])
        cam = DF.TrilinearResize.apply(dense_out, tuple(self.at_spatial_size))
        refined = self.attention_module(cam, attention_features)
        return DF.TrilinearResize.apply(refined, raw)

    def forward(self, x, lungs=None):
        x = ops.to_cl(x, "DC3DATGeneric input")
        att = [x] if -1 in self.at_layers else []

        def tap(idx, feats):
            if idx in self.at_layers:
                att.append(self._reshape_head(len(att) - (1 if -1 in self.at_layers else 0), feats))

        dense_outs = self._encode_decode(x, decoder_ckpt_offset=self.n_layers + 1, tap=tap,     # models.py:578 indexing
                                         size=x.shape[-3:])
        size = tuple(self.at_spatial_size)
        att = torch.cat([DF.TrilinearResize.apply(f, size) for f in att], dim=1)               # models.py:591-594
        refined = self.apply_attention(x, lungs, dense_outs, att)
        return dense_outs, refined
