"""CPU oracle for the DRAM hot path (TEST INFRASTRUCTURE — never imported by the product).

A from-scratch, *functional* restatement (pure functions over a `state_dict`, torch CPU fp32 + numpy) of what the
reference computes on the path `BASELINE.json::north_star` names.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it, and only as the checker / the timed CPU arm.

Pinning: the reference ships no tests/golden vectors, so this oracle is pinned against the
reference ITSELF, executed in the build container through `oracle/ref_shims.py`
(tests/test_oracle_vs_reference.py, skipped when /root/reference is absent) and against the fixtures under
`tests/golden/` that `tests/golden/make_golden.py` generated from the reference.  Third-party pieces that the
reference calls but does not vendor are restated from their published semantics and are "parity unpinned":
DGL degree-bucketed UDF reduce (we pin against the fake-DGL shim, see ref_shims.py), SimpleITK 1.1.0
ResampleImageFilter, scikit-image threshold_otsu.

Every function cites the reference file:line (relative to /root/reference/dram/) it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm3d default, parts.py:19
BN_MOMENTUM = 0.1


# ============================================================================================ building blocks
def conv_bn_relu(sd, prefix, x, training, stat_updates=1, padding=1):
    """One [Conv3d(bias=False) -> BatchNorm3d -> ReLU] unit.  parts.py:103-110 / 140-146 / 184-190.

    `prefix` names the nn.Sequential holding (0)=conv, (1)=bn.  In training mode the running statistics in `sd`
    are updated `stat_updates` times (a block wrapped in torch.utils.checkpoint re-runs its forward during
    backward, so its BatchNorm momentum update is applied twice per step: models.py:123-143)."""
    w = sd[prefix + ".0.weight"]
    b = sd.get(prefix + ".0.bias")
    y = F.conv3d(x, w, b, padding=padding)
    rm, rv = sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"]
    g, be = sd[prefix + ".1.weight"], sd[prefix + ".1.bias"]
    if training:
        out = F.batch_norm(y, None, None, g, be, True, BN_MOMENTUM, BN_EPS)
        with torch.no_grad():
            n = y.numel() / y.shape[1]
            mean = y.mean(dim=(0, 2, 3, 4))
            var_unbiased = y.var(dim=(0, 2, 3, 4), unbiased=True) if n > 1 else torch.zeros_like(mean)
            for _ in range(stat_updates):
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var_unbiased)
                sd[prefix + ".1.num_batches_tracked"] += 1
    else:
        out = F.batch_norm(y, rm, rv, g, be, False, BN_MOMENTUM, BN_EPS)
    return F.relu(out)


def conv_block(sd, prefix, x, training, stat_updates, paddings=(1, 1)):
    """Two conv_bn_relu units: ConvBlock5d.forward parts.py:112-113."""
    x = conv_bn_relu(sd, prefix + ".conv_blocks.0", x, training, stat_updates, paddings[0])
    return conv_bn_relu(sd, prefix + ".conv_blocks.1", x, training, stat_updates, paddings[1])


def crop_concat(t1, t2):
    """crop_concat_5d parts.py:37-46: centre-crop t2 to t1's DHW with ceil offsets, cat [t1, t2] on C."""
    sl = [slice(None), slice(None)]
    for a, b in zip(t1.shape[2:], t2.shape[2:]):
        off = int(math.ceil((b - a) / 2))
        sl.append(slice(off, off + a))
    return torch.cat([t1, t2[tuple(sl)]], dim=1)


def trilinear(x, size):
    """F.interpolate(..., mode='trilinear', align_corners=True): src = dst*(in-1)/(out-1).  models.py:146,514-518."""
    return F.interpolate(x, size=tuple(size), mode="trilinear", align_corners=True)


def masked_pool(dense, lungs, method="avg"):
    """pooling_dense_features models.py:37-49."""
    B, C = dense.shape[:2]
    if method == "global_avg":
        return dense.reshape(B, C, -1).mean(-1)
    if method == "global_max":
        return dense.reshape(B, C, -1).amax(-1)
    le = lungs.expand_as(dense)
    return (dense * le).reshape(B, C, -1).sum(-1) / le.reshape(B, C, -1).sum(-1)


# ============================================================================================ U-Net (DC3D)
def _ckpt_updates(cfg, idx, training, counting_grad):
    return 2 if (training and counting_grad and cfg["checkpoint_layers"][idx] > 0) else 1


def unet_forward(sd, x, cfg, training=False, attention=False, emulate_checkpoint_stats=True):
    """DC3D.forward models.py:120-147 / the U-Net part of DC3DATGeneric.forward models.py:550-588.

    Returns (dense_outs, taps) where taps = {layer index: pre-pool features} (ds blocks) for the attention heads.
    Note the two classes index `checkpoint_layers` differently for the decoder (models.py:140 vs :578)."""
    n = cfg["n_layers"]
    pads = cfg["padding_list"]
    stacking = cfg.get("stacking", 3 if attention else 0)
    cg = emulate_checkpoint_stats and torch.is_grad_enabled()
    taps = {}
    skips = []
    h = x
    for i in range(n):
        y = conv_block(sd, f"ds_modules.{i}", h, training, _ckpt_updates(cfg, i, training, cg), pads[i])
        taps[i] = y
        skips.append(y)
        h = F.max_pool3d(y, 2, 2, 0)                                   # parts.py:191,195
    h = conv_block(sd, "bg", h, training, _ckpt_updates(cfg, n, training, cg), pads[n])
    taps[n] = h
    for i, skip in enumerate(reversed(skips)):
        if stacking == i:
            break
        ck = n + 1 + i if attention else n + i                          # models.py:578 vs models.py:140
        up = F.interpolate(h, scale_factor=tuple(cfg.get("upsample_sf", (2, 2, 2))), mode="trilinear",
                           align_corners=True)                          # parts.py:149,152
        h = conv_block(sd, f"us_modules.{i}", crop_concat(up, skip), training,
                       _ckpt_updates(cfg, ck, training, cg), pads[n + 1 + i])
        taps[n + 1 + i] = h
    dense = F.conv3d(h, sd["top_layer.weight"], sd["top_layer.bias"])    # models.py:145
    dense = trilinear(dense, x.shape[-3:])                               # models.py:146 (identity at config)
    return dense, taps


def dc3d_forward(sd, x, cfg, training=False):
    dense, _ = unet_forward(sd, x, cfg, training, attention=False)
    return dense, dense                                                  # models.py:147


# ============================================================================================ PCM stencil attention
def stencil_offsets(connectivity=2, self_loop=False):
    """Neighbour offsets of PCM.init_graph models.py:223-259 for k_size=3:
    scipy generate_binary_structure(3, connectivity) minus the centre unless self_loop."""
    offs = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                l1 = abs(dz) + abs(dy) + abs(dx)
                if l1 == 0:
                    if self_loop:
                        offs.append((dz, dy, dx))
                elif l1 <= connectivity:
                    offs.append((dz, dy, dx))
    return offs


def _shift(t, off, fill):
    """t[..., z+dz, y+dy, x+dx] with out-of-grid positions = fill."""
    dz, dy, dx = off
    p = F.pad(t, (1, 1, 1, 1, 1, 1), value=fill)
    D, H, W = t.shape[-3:]
    return p[..., 1 + dz:1 + dz + D, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]


def pcm_forward(sd, cam, f, prefix="attention_module", merge_type="scaled_dot_product_relu", self_loop=False,
                connectivity=2, non_local_iter=1, residual=False):
    """PCM.forward + compute_cross_x + merge_func, dense form.  models.py:322-353,355-389,261-320.

    cam [B,1,D,H,W], f [B,Cf,D,H,W].  Logits s_o = <theta(f_x), phi(f_{x+o})>; the softmax temperature of the
    'scaled_dot_product*' variants is sqrt(number of in-edges of x) (models.py:274,277: f.shape[-1] is the
    mailbox degree); absent neighbours do not exist in the graph -> masked out."""
    B = f.shape[0]
    q = torch.einsum("bcdhw,oc->bodhw", f, sd[prefix + ".theta.weight"]) + sd[prefix + ".theta.bias"].view(1, -1, 1, 1, 1)
    k = torch.einsum("bcdhw,oc->bodhw", f, sd[prefix + ".phi.weight"]) + sd[prefix + ".phi.bias"].view(1, -1, 1, 1, 1)
    offs = stencil_offsets(connectivity, self_loop)
    ones = torch.ones(1, 1, *f.shape[-3:], dtype=f.dtype)
    valid = torch.stack([_shift(ones, o, 0.0) for o in offs], dim=0)            # [O,1,1,D,H,W]
    deg = valid.sum(0)                                                           # [1,1,D,H,W]
    logits = torch.stack([(q * _shift(k, o, 0.0)).sum(1, keepdim=True) for o in offs], dim=0)   # [O,B,1,D,H,W]
    if merge_type in ("scaled_dot_product_relu", "smrelu"):
        logits = F.relu(logits)
    elif merge_type not in ("scaled_dot_product", "sm", "smscaled"):
        raise NotImplementedError(merge_type)
    if merge_type.startswith("scaled_dot_product"):
        logits = logits / torch.sqrt(deg)
    elif merge_type == "smscaled":
        logits = logits / 0.01
    logits = logits.masked_fill(valid.expand_as(logits) == 0, float("-inf"))
    att = torch.softmax(logits, dim=0)
    wG, bG = sd[prefix + ".G.weight"], sd[prefix + ".G.bias"]                    # [g,1], [g]
    wr, br = sd[prefix + ".r.weight"], sd[prefix + ".r.bias"]                    # [1,g], [1]
    for _ in range(non_local_iter):
        nb = torch.stack([_shift(cam, o, 0.0) for o in offs], dim=0)            # [O,B,1,D,H,W]
        xg = nb.unsqueeze(-1) * wG.view(-1) + bG                                 # G(cam_j)  [O,B,1,D,H,W,g]
        y = (att.unsqueeze(-1) * xg).sum(0)                                      # [B,1,D,H,W,g]
        new = (y * wr.view(-1)).sum(-1) + br                                     # r(y)
        cam = new + cam if residual else new
    return cam


def dc3dat_forward(sd, x, cfg, training=False):
    """DC3DATGeneric.forward models.py:550-597 (+ apply_attention 510-518)."""
    dense, taps = unet_forward(sd, x, cfg, training, attention=True)
    at_size = tuple(cfg["at_spatial_size"])
    feats = []
    nc = 0
    if -1 in cfg["at_layers"]:
        feats.append(x)
    for l in sorted(l for l in cfg["at_layers"] if l != -1):
        t = taps[l].detach()                                                     # models.py:564
        p = f"reshape.{nc}"
        y = F.conv3d(t, sd[p + ".0.weight"], sd[p + ".0.bias"])
        if training:
            out = F.batch_norm(y, None, None, sd[p + ".1.weight"], sd[p + ".1.bias"], True, BN_MOMENTUM, BN_EPS)
            with torch.no_grad():
                nel = y.numel() / y.shape[1]
                sd[p + ".1.running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * y.mean(dim=(0, 2, 3, 4)))
                sd[p + ".1.running_var"].mul_(1 - BN_MOMENTUM).add_(
                    BN_MOMENTUM * (y.var(dim=(0, 2, 3, 4), unbiased=True) if nel > 1 else 0.0))
                sd[p + ".1.num_batches_tracked"] += 1
        else:
            out = F.batch_norm(y, sd[p + ".1.running_mean"], sd[p + ".1.running_var"], sd[p + ".1.weight"],
                               sd[p + ".1.bias"], False, BN_MOMENTUM, BN_EPS)
        feats.append(F.relu(out))
        nc += 1
    feats = torch.cat([trilinear(t, at_size) for t in feats], dim=1)             # models.py:591-594
    cam = trilinear(dense, at_size)                                              # models.py:514
    refined = pcm_forward(sd, cam, feats, merge_type=cfg["at_merge_type"], self_loop=cfg["at_self_loop"],
                          non_local_iter=cfg["at_g_iter"])
    refined = trilinear(refined, dense.shape[2:])                                # models.py:517
    return dense, refined


# ============================================================================================ loss (caller)
CTSS_RATIO_MAP = {0: (0.0, 0.001), 1: (0.001, 0.01), 2: (0.01, 0.05), 3: (0.05, 0.35), 4: (0.35, 0.5),
                  5: (0.5, 1.00001)}                                             # metrics.py:76-83


def interval_targets(ctsses, ratio_upper_bound, band_width):
    """IntRegLoss.get_labels metrics.py:121-137, in float64 like the Python-float arithmetic it replaces."""
    out = []
    for c, p in zip(ctsses, ratio_upper_bound.reshape(-1).tolist()):
        lb, ub = max(0.0, p - band_width), min(1.0, p + band_width)
        clb, cub = CTSS_RATIO_MAP[int(float(c))]
        band = (max(clb, lb), min(cub, ub))
        if band[1] < band[0]:
            if ub <= clb:
                band = (lb, ub)
            elif lb >= cub:
                band = (clb, cub)
            else:
                raise RuntimeError("cannot reach here!")
        out.append(band)
    return torch.tensor(out, dtype=torch.float32)


def reg_loss(dense, lobes, lesions, ctsses, freq_map, band_width=1e-2):
    """IntRegLoss.compute_reg_loss_with_probs metrics.py:158-177 (probs = sigmoid(dense), metrics.py:363)."""
    B = dense.shape[0]
    probs = torch.sigmoid(dense)
    rub = (lesions * lobes).reshape(B, -1).sum(-1) / lobes.reshape(B, -1).sum(-1)
    inside = (lobes > 0).to(dense.dtype)
    pred = (probs * inside).reshape(B, -1).sum(-1) / inside.reshape(B, -1).sum(-1)
    tg = interval_targets(ctsses, rub, band_width).to(dense.device)
    K = (0.5 * (tg[:, 1] - tg[:, 0])) ** 2
    unhinge = (pred - (tg[:, 1] + tg[:, 0]) / 2.0) ** 2 - K
    w = torch.tensor([freq_map[int(float(c))] for c in ctsses], dtype=torch.float32).clamp(0.2, 0.8).to(dense.device)
    return (torch.clamp_min(unhinge, 0.0) / w).sum()


def pseudo_labels(dense, lobes, lesions, ctsses):
    """IntRegRefineLoss.compute_seg_loss metrics.py:333-354 + threshold_postprocessing :325-329."""
    pred = (torch.sigmoid(dense.detach()) > 0.5) & (lobes != 0)
    t = (pred & (lesions > 0)).to(dense.dtype)
    keep = torch.tensor([0.0 if float(c) < 1e-7 else 1.0 for c in ctsses], dtype=dense.dtype, device=dense.device)
    return t * keep.view(-1, 1, 1, 1, 1)


def boot_bce(p, t, voi, smoothing=0.1, eps=1e-7):
    """BootBinCrossEntropy.__call__ metrics.py:17-51."""
    tb = ~voi
    po, to = p[tb], t[tb]
    pto = (po * to + (1.0 - po) * (1.0 - to)).clamp(eps, 1.0 - eps)
    bceo = (-torch.log(pto)).mean()
    if voi.sum() > 0:
        pi, ti = p[voi], t[voi]
        alpha = (1.0 - ti.sum() / voi.sum()).clamp(0.25, 0.75)
        pti = (pi * ti + (1.0 - pi) * (1.0 - ti)).clamp(eps, 1.0 - eps)
        w = alpha * ti + (1.0 - alpha) * (1.0 - ti)
        bce = (-torch.log(pti) * w).sum() / w.sum()
        th = (pi > 0.5).to(p.dtype)
        pth = (pi * th + (1.0 - pi) * (1.0 - th)).clamp(eps, 1.0 - eps)
        boot = (-torch.log(pth)).mean()
        return bceo + (1.0 - smoothing) * bce + smoothing * boot
    return bceo


def int_reg_refine_loss(dense, refined, lobes, lesions, ctsses, freq_map, band_width=1e-2, smoothing=0.1):
    """IntRegRefineLoss.__call__ metrics.py:360-373 -> (reg_loss, seg_loss)."""
    r = reg_loss(dense, lobes, lesions, ctsses, freq_map, band_width)
    t = pseudo_labels(dense, lobes, lesions, ctsses)
    s = boot_bce(torch.sigmoid(refined), t, lobes > 0, smoothing)
    return r, s


# ============================================================================================ inference plumbing
def windowing(image, from_span=(-1150, 350), to_span=(0, 255)):
    """utils.windowing utils.py:189-198."""
    lo, hi = (np.min(image), np.max(image)) if from_span is None else from_span
    image = np.clip(image, a_min=lo, a_max=hi)
    return ((image - lo) / float(hi - lo)) * (to_span[1] - to_span[0]) + to_span[0]


def find_crops(mask, spacing, border):
    """utils.find_crops utils.py:244-254: bounding box of mask>0 padded by ceil(border/spacing) voxels, clipped."""
    nz = np.nonzero(mask > 0)
    sl = []
    for ax, (ss, sp) in enumerate(zip(mask.shape, spacing)):
        lo, hi = int(nz[ax].min()), int(nz[ax].max()) + 1
        pad = int(math.ceil(border / sp)) if border > 0 else 0
        sl.append(slice(max(0, lo - pad), min(ss, hi + pad)))
    return tuple(sl)


def itk_continuous_indices(in_size, out_size, in_spacing, out_spacing):
    """Continuous INPUT index of every output index along each axis, as ITK's ResampleImageFilter computes it for an
    identity transform, identity direction and a shared origin — in float64, following (not fitted to any kernel):

      * itk::ImageBase::TransformIndexToPhysicalPoint:            point = spacing_out * index  (+ origin)
      * itk::ImageBase::TransformPhysicalPointToContinuousIndex:  cidx  = (1 / spacing_in) * (point - origin)
      * itk::ResampleImageFilter::LinearThreadedGenerateData (the path taken for a linear — here identity — transform):
        per output scanline (ITK x = our LAST axis) the continuous index of the line's first pixel (`startIndex`) and of
        index `start + size` (`endIndex`) are mapped as above, and pixel i of the line gets
            alpha = i / double(size);  inputIndex = startIndex + alpha * (endIndex - startIndex)
        so the y / z components (equal at both ends of a line) are the mapped values themselves.

    ITK version: SimpleITK 1.1.0 (requirements.in:33) bundles ITK 4.13; `sitk.GetImageFromArray` gives origin 0 and identity
    direction (utils.py:418-419), the matrix inverse of diag(spacing) is taken as 1/spacing.  Parity unpinned (no SimpleITK
    in this image).  Axes are (z, y, x); returns three float64 arrays."""
    in_spacing = [float(v) for v in in_spacing]
    out_spacing = [float(v) for v in out_spacing]
    inv = [1.0 / v for v in in_spacing]
    cz = inv[0] * (out_spacing[0] * np.arange(out_size[0], dtype=np.float64))
    cy = inv[1] * (out_spacing[1] * np.arange(out_size[1], dtype=np.float64))
    start = inv[2] * (out_spacing[2] * 0.0)
    end = inv[2] * (out_spacing[2] * float(out_size[2]))
    alpha = np.arange(out_size[2], dtype=np.float64) / float(out_size[2])
    cx = start + alpha * (end - start)
    return [cz, cy, cx]


def fixed_size_spacing(in_size, out_size, spacing=(1.0, 1.0, 1.0)):
    """Resample('fixed_size') data_transforms.py:170-175: require_spacing = spacing * (current_size / size), float64."""
    ratios = np.asarray(in_size, dtype=np.float64) / np.asarray(out_size, dtype=np.float64)
    return (np.asarray(spacing, dtype=np.float64) * ratios).tolist()


def itk_resample(arr, new_size, interpolator="linear", in_spacing=None, out_spacing=None):
    """SimpleITK ResampleImageFilter as used by utils.resample utils.py:414-434 (+ resample_sitk_image :299-384) and
    Resample('fixed_size') data_transforms.py:170-175: identity transform, same origin, default pixel value 0, output pixel
    type = input pixel type.  An INDEPENDENT float64 restatement of the ITK 4.13 classes (see itk_continuous_indices for the
    coordinate part):

      * itk::InterpolateImageFunction::IsInsideBuffer: a continuous index outside [-0.5, n - 0.5) on any axis -> default 0;
      * itk::LinearInterpolateImageFunction::EvaluateOptimized (3-D): base = floor(cidx) clamped to the first index,
        distance = cidx - base (a non-positive distance means "no interpolation on this axis"), neighbours beyond the last
        index fall back to the base sample; values are combined in RealType = double as nested lerps x, then y, then z:
            vx00 = v000 + (v100 - v000) * dx;  vxx0 = vx00 + (vx10 - vx00) * dy;  out = vxx0 + (vxx1 - vxx0) * dz
        (every special-case branch of the ITK function equals this expression with dx/dy/dz = 0 or equal neighbours);
      * itk::NearestNeighborInterpolateImageFunction: index = Math::RoundHalfIntegerUp(cidx) = floor(cidx + 0.5);
      * itk::ResampleImageFilter::CastPixelWithBoundsChecking: clamp to the output type's range, then static_cast —
        truncation toward zero for integer pixels, round-to-nearest for float32.

    in_spacing / out_spacing (z, y, x) default to Resample('fixed_size'): in = 1, out = in * in_size / out_size."""
    arr = np.asarray(arr)
    if in_spacing is None and out_spacing is None and tuple(arr.shape) == tuple(new_size):
        return arr                                                              # utils.py:415-417
    in_spacing = [1.0, 1.0, 1.0] if in_spacing is None else in_spacing
    out_spacing = fixed_size_spacing(arr.shape, new_size, in_spacing) if out_spacing is None else out_spacing
    coords = itk_continuous_indices(arr.shape, new_size, in_spacing, out_spacing)
    inside = [(c >= -0.5) & (c < n - 0.5) for c, n in zip(coords, arr.shape)]
    inside3 = inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :]
    if interpolator == "nearest":
        idx = [np.clip(np.floor(c + 0.5).astype(np.int64), 0, n - 1) for c, n in zip(coords, arr.shape)]
        out = arr[np.ix_(*idx)].copy()
        out[~inside3] = 0
        return out
    base = [np.maximum(np.floor(c).astype(np.int64), 0) for c in coords]
    dist = [np.maximum(c - b.astype(np.float64), 0.0) for c, b in zip(coords, base)]
    i0 = [np.minimum(b, n - 1) for b, n in zip(base, arr.shape)]                  # only reached when outside the buffer
    i1 = [np.minimum(b + 1, n - 1) for b, n in zip(base, arr.shape)]              # beyond the last index -> the base sample
    src = arr.astype(np.float64)
    dz, dy, dx = dist[0][:, None, None], dist[1][None, :, None], dist[2][None, None, :]

    def g(zi, yi, xi):
        return src[np.ix_(zi, yi, xi)]

    def lerp_x(zi, yi):
        a = g(zi, yi, i0[2])
        return a + (g(zi, yi, i1[2]) - a) * dx

    def lerp_xy(zi):
        a = lerp_x(zi, i0[1])
        return a + (lerp_x(zi, i1[1]) - a) * dy

    lo = lerp_xy(i0[0])
    out = lo + (lerp_xy(i1[0]) - lo) * dz
    out[~inside3] = 0.0
    if np.issubdtype(arr.dtype, np.integer):
        info = np.iinfo(arr.dtype)
        return np.trunc(np.clip(out, info.min, info.max)).astype(arr.dtype)     # static_cast: truncation toward zero
    return out.astype(arr.dtype)


def resample_to_spacing(arr, spacing, new_spacing, interpolator="linear"):
    """Resample('fixed_spacing') data_transforms.py:76-83 + resample_sitk_image utils.py:369-371:
    new_size = ceil(size * spacing / new_spacing), output spacing = new_spacing."""
    spacing, new_spacing = np.asarray(spacing, np.float64), np.asarray(new_spacing, np.float64)
    new_size = np.ceil(np.asarray(arr.shape) * (spacing / new_spacing)).astype(int)
    return itk_resample(arr, tuple(new_size), interpolator, in_spacing=spacing.tolist(), out_spacing=new_spacing.tolist())


def threshold_otsu_u8(values_u8):
    """skimage.filters.threshold_otsu on uint8 data (utils.py:239), restated: for integer images skimage histograms with
    bincount (one bin per integer value between min and max, bin centres = the integer values), maximises
    w1*w2*(mu1-mu2)^2 between neighbouring bins and returns the bin centre (scikit-image <= 0.17, unpinned)."""
    v = np.asarray(values_u8).ravel().astype(np.int64)
    lo, hi = int(v.min()), int(v.max())
    hist = np.bincount(v - lo, minlength=hi - lo + 1).astype(np.float64)
    centers = np.arange(lo, hi + 1, dtype=np.float64)
    w1 = np.cumsum(hist)
    w2 = np.cumsum(hist[::-1])[::-1]
    m1 = np.cumsum(hist * centers) / np.maximum(w1, 1e-300)
    m2 = (np.cumsum((hist * centers)[::-1]) / np.maximum(w2[::-1], 1e-300))[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return centers[:-1][int(np.argmax(var12))]


def binary_cam(values, scaler=1.0, from_span=(0, 1)):
    """utils.binary_cam utils.py:226-242 -> threshold in [0,1] units."""
    w = windowing(np.asarray(values), from_span=from_span).astype(np.uint8)
    u = np.unique(w)
    if len(u) < 2:
        return u[0] / 255.0
    return min(threshold_otsu_u8(w) * scaler, 255.0) / 255.0


def preprocess_lobe_chunk(scan, lobe, label, spacing, window, chunk_size=(80, 80, 80), border=5, pad_value=-2048):
    """job_runner.py:955-984: crop one lobe, blank outside-lobe voxels, window, resample to the chunk grid."""
    lobe_binary = lobe == label
    sl = find_crops(lobe_binary, spacing, border)
    lobe_chunk = lobe_binary[sl]
    scan_chunk = scan[sl].copy()
    scan_chunk[lobe_chunk == 0] = pad_value
    img = windowing(scan_chunk.astype(np.int16).astype(np.float32), from_span=window, to_span=(0, 1))
    img = itk_resample(img.astype(np.float32), chunk_size, "linear", in_spacing=spacing)
    msk = itk_resample(lobe_chunk.astype(np.uint8), chunk_size, "nearest", in_spacing=spacing)
    return sl, lobe_chunk, img.astype(np.float32), msk.astype(np.float32)


def infer_scan(model_fn, scan, lobe, spacing, window=(-1000, -700), chunk_size=(80, 80, 80), head="sigmoid",
               vessel=None):
    """Per-scan inference loop.  head='literal' follows LesionSegTest.run job_runner.py:951-1015 (argmax over the
    out_ch=1 pooled score is always 0, so every heat-map is zeroed: SURVEY D4); head='sigmoid' follows
    LesionSegChunkTrain.evaluate_scan job_runner.py:730-772 (sigmoid -> trilinear to crop -> paste inside lobe).
    `model_fn(image[1,1,D,H,W], lobe[1,1,D,H,W]) -> (dense, refined)`."""
    htp = np.zeros(scan.shape, dtype=np.float32)
    for label in range(1, 6):
        if (lobe == label).sum() < 1e-7:
            continue
        sl, lobe_chunk, img, msk = preprocess_lobe_chunk(scan, lobe, label, spacing, window, chunk_size)
        ti = torch.from_numpy(img)[None, None]
        tl = torch.from_numpy(msk)[None, None]
        with torch.no_grad():
            _, dense = model_fn(ti, tl)
            if head == "literal":
                pool = masked_pool(dense, tl)
                cls_pred = int(torch.max(pool, dim=-1)[-1].item())
                d = F.relu(trilinear(dense, lobe_chunk.shape).squeeze(0))[cls_pred]
                d = d / d.max()
                if cls_pred < 1e-7:
                    d = torch.zeros_like(d)
            else:
                d = trilinear(torch.sigmoid(dense), lobe_chunk.shape)[0, 0]
        view = htp[sl]
        view[lobe_chunk > 0] = d.numpy()[lobe_chunk > 0]
    th = binary_cam(htp[lobe > 0])                                               # job_runner.py:1008
    lesion_pred = htp > th
    w_scan = windowing(scan, to_span=(0, 1))                                     # job_runner.py:1011
    th2 = binary_cam(w_scan[lobe > 0], 0.75)
    post = lesion_pred & (w_scan > th2)
    if vessel is not None:
        post &= ~(vessel > 0)
    ratio = float((htp * (lobe > 0)).sum() / max((lobe > 0).sum(), 1))           # job_runner.py:772
    return {"heatmap": htp, "lesion": lesion_pred.astype(np.uint8), "lesion_post": post.astype(np.uint8),
            "threshold": th, "ratio": ratio}


# ============================================================================================ synthetic inputs
def ellipsoid_lobe(B, size, seed=0):
    """Deterministic non-empty lobe masks: ellipsoid radii 0.45*size, centre jittered by seed."""
    g = np.random.RandomState(seed)
    D, H, W = size
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    out = np.zeros((B, 1, D, H, W), dtype=np.float32)
    for b in range(B):
        c = [(s - 1) / 2.0 + g.uniform(-0.05, 0.05) * s for s in size]
        r = [0.45 * s for s in size]
        out[b, 0] = (((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2) <= 1.0
    return torch.from_numpy(out)


def synthetic_batch(B, size=(80, 80, 80), seed=0):
    """Synthetic training batch of SURVEY §8d: images in [0,1) zeroed outside the lobe, ellipsoid lobes,
    random lesion candidates inside the lobe, cycling CT severity scores."""
    g = torch.Generator().manual_seed(seed)
    lobes = ellipsoid_lobe(B, size, seed)
    images = torch.rand(B, 1, *size, generator=g)
    images[lobes == 0] = 0
    lesions = (torch.rand(B, 1, *size, generator=g) > 0.7).float() * lobes
    ctsses = [str(i % 6) for i in range(B)]
    return images, lobes, lesions, ctsses


def synthetic_scan(shape=(400, 512, 512), spacing=(1.0, 0.7, 0.7), seed=0):
    """Synthetic CT: air -1000, body ellipse +40, two lungs -850+N(0,50) with -300 blobs,
    5 lobe labels (3 right, 2 left).  Returned on the scan's native grid."""
    g = np.random.RandomState(seed)
    D, H, W = shape
    zz = np.arange(D, dtype=np.float32)[:, None, None]
    yy = np.arange(H, dtype=np.float32)[None, :, None]
    xx = np.arange(W, dtype=np.float32)[None, None, :]
    scan = np.full(shape, -1000, dtype=np.int16)
    body = ((yy - H / 2) / (0.42 * H)) ** 2 + ((xx - W / 2) / (0.46 * W)) ** 2 <= 1.0
    scan[np.broadcast_to(body, shape)] = 40
    lobe = np.zeros(shape, dtype=np.uint8)
    for side, cx in ((0, 0.30 * W), (1, 0.70 * W)):
        lung = (((zz - D / 2) / (0.40 * D)) ** 2 + ((yy - H / 2) / (0.28 * H)) ** 2
                + ((xx - cx) / (0.15 * W)) ** 2) <= 1.0
        if side == 0:
            lab = np.where(zz < 0.38 * D, 1, np.where(zz < 0.6 * D, 2, 3)).astype(np.uint8)
        else:
            lab = np.where(zz < 0.5 * D, 4, 5).astype(np.uint8)
        lobe = np.where(lung, np.broadcast_to(lab, shape), lobe)
    lungs = lobe > 0
    noise = (g.randn(*shape).astype(np.float32) * 50.0 - 850.0)
    scan = np.where(lungs, noise, scan).astype(np.int16)
    lesion = np.zeros(shape, dtype=bool)
    nz = np.argwhere(lungs)
    for _ in range(12):
        c = nz[g.randint(len(nz))]
        r = g.uniform(6, 18)
        blob = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= r * r
        lesion |= blob & lungs
    scan = np.where(lesion, np.int16(-300), scan).astype(np.int16)
    return scan, lobe, lesion.astype(np.uint8), np.asarray(spacing, dtype=np.float64)
