"""Pin the CPU oracle: against the committed golden vectors (generated from the unmodified reference) and, when
/root/reference is present (build container only), against the reference executed live."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle_import import O
from util import assert_close, rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sd(g, grad=False):
    sd = {k: v.clone() for k, v in g["state_dict"].items()}
    if grad:
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize("fixture,att", [("dc3d_div8_16.pt", False), ("dc3dat_div16_16.pt", True)])
def test_oracle_matches_golden(fixture, att):
    g = torch.load(os.path.join(GOLDEN, fixture))
    fwd = O.dc3dat_forward if att else O.dc3d_forward
    with torch.no_grad():
        d, r = fwd(_sd(g), g["images"], g["cfg"], False)
    assert_close(d, g["eval_dense"], 1e-6, "dense")
    assert_close(r, g["eval_refined"], 1e-5, "refined")
    assert_close(O.masked_pool(d, g["lobes"]), g["eval_pool"], 1e-6, "pool")
    sd = _sd(g, grad=True)
    d, r = fwd(sd, g["images"], g["cfg"], True)
    rl, sl = O.int_reg_refine_loss(d, r, g["lobes"], g["lesions"], g["ctsses"], g["freq_map"], **{
        k: v for k, v in g["loss_cfg"].items() if k in ("band_width", "smoothing")})
    (rl * g["loss_factors"][0] + sl * g["loss_factors"][1]).backward()
    assert_close(rl, g["train_reg_loss"], 1e-5, "reg loss")
    assert_close(sl, g["train_seg_loss"], 1e-5, "seg loss")
    scale = max(v.abs().max().item() for v in g["grads"].values())
    for k, ref in g["grads"].items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(ref)
        assert (got - ref).abs().max().item() <= 2e-5 * scale, k
    for k, v in g["state_dict_after"].items():
        assert_close(sd[k].float(), v.float(), 1e-5, k)


def test_oracle_pcm_matches_reference_dgl_semantics():
    g = torch.load(os.path.join(GOLDEN, "pcm_6x5x7.pt"))
    sd = {"attention_module." + k: v for k, v in g["state_dict"].items()}
    out = O.pcm_forward(sd, g["cam"], g["f"])
    assert_close(out, g["out"], 1e-5, "pcm")
    D, H, W = g["grid"]
    ones = torch.ones(1, 1, D, H, W)
    deg = sum(O._shift(ones, o, 0.0) for o in O.stencil_offsets(2, False))
    assert int(deg.sum()) == g["n_edges"] and D * H * W == g["n_nodes"]


def test_stencil_matches_survey_graph_statistics():
    """64^3 grid, connectivity 2, no self loop: 4 596 480 edges, in-degrees {18,13,9,6}"""
    ones = torch.ones(1, 1, 64, 64, 64)
    deg = sum(O._shift(ones, o, 0.0) for o in O.stencil_offsets(2, False))
    assert int(deg.sum()) == 4596480
    vals, counts = torch.unique(deg, return_counts=True)
    assert dict(zip(vals.int().tolist(), counts.tolist())) == {18: 238328, 13: 23064, 9: 744, 6: 8}


@pytest.mark.skipif(not os.path.isdir("/root/reference/dram"), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    import ref_shims
    ref = ref_shims.load_reference()
    s = ref_shims.load_settings("st_dram_ref_att.py")
    cfg = dict(s.MODEL)
    cfg["in_ch_list"], cfg["base_ch_list"], cfg["end_ch_list"] = [1, 8, 16, 32, 96, 48, 24], [4, 8, 16, 32, 32, 16, 8], [8, 16, 32, 64, 32, 16, 8]
    cfg["at_spatial_size"] = (10, 12, 9)
    m = ref_shims.build_reference_model(cfg, seed=4)
    x, lobes, lesions, ctss = O.synthetic_batch(3, (16, 16, 16), seed=9)
    m.eval()
    with torch.no_grad():
        rd, rr = m(x, lobes)
        d, r = O.dc3dat_forward({k: v.clone() for k, v in m.state_dict().items()}, x, cfg, False)
    assert_close(d, rd, 1e-6, "dense")
    assert_close(r, rr, 1e-5, "refined")
    # literal inference head of LesionSegTest.run vs the sane sigmoid head
    for k in (ref.utils.windowing, ):
        a = np.random.RandomState(0).randint(-2048, 1000, size=(5, 6, 7)).astype(np.int16)
        assert np.array_equal(k(a), O.windowing(a))
        assert np.array_equal(k(a.astype(np.float32), from_span=(-1000, -700), to_span=(0, 1)),
                              O.windowing(a.astype(np.float32), from_span=(-1000, -700), to_span=(0, 1)))
    # utils.find_crops (utils.py:244-254) cannot be executed here: scipy >= 1.10 rejects the bool array it passes to
    # ndimage.find_objects; the restatement is checked against the documented semantics instead.
    mask = np.zeros((20, 30, 25), bool)
    mask[5:9, 10:22, 3:20] = True
    assert O.find_crops(mask, (1.0, 0.7, 0.7), 5) == (slice(0, 14), slice(2, 30), slice(0, 25))


def test_inference_plumbing_restatement():
    rng = np.random.RandomState(1)
    a = rng.rand(10, 12, 9).astype(np.float32)
    assert np.array_equal(O.itk_resample(a, (10, 12, 9)), a)
    up = O.itk_resample(a, (20, 24, 18), "linear")
    assert up.shape == (20, 24, 18) and np.allclose(up[::2, ::2, ::2], a, atol=1e-6)      # even samples hit source voxels
    m = (a > 0.5).astype(np.uint8)
    nn = O.itk_resample(m, (20, 24, 18), "nearest")
    assert set(np.unique(nn)) <= {0, 1} and np.array_equal(nn[::2, ::2, ::2], m)
    v = np.concatenate([rng.randint(0, 60, 5000), rng.randint(180, 255, 3000)]).astype(np.uint8)
    th = O.threshold_otsu_u8(v)
    assert 58 <= th <= 180      # flat maximum over the empty gap: first bin wins, like skimage
    assert th == int(th)        # integer images: bincount histogram, integer bin centres
    assert 0.2 < O.binary_cam(v / 255.0) < 0.75


def test_infer_scan_heads():
    scan, lobe, lesion, spacing = O.synthetic_scan((24, 32, 32), (1.0, 1.0, 1.0), seed=0)
    assert set(np.unique(lobe)) == {0, 1, 2, 3, 4, 5}
    fn = lambda img, lb: (img * 4 - 2, img * 4 - 2)
    out = O.infer_scan(fn, scan, lobe, spacing, window=(-1000, -300), chunk_size=(8, 8, 8), head="sigmoid")
    assert out["heatmap"].shape == scan.shape and (out["heatmap"][lobe == 0] == 0).all()
    assert out["heatmap"][lobe > 0].max() > 0
    lit = O.infer_scan(fn, scan, lobe, spacing, window=(-1000, -300), chunk_size=(8, 8, 8), head="literal")
    assert (lit["heatmap"] == 0).all()        # the shipped out_ch=1 head zeroes every heat map


def test_itk_resample_known_answers_from_the_itk_definitions():
    """Hand-derived vectors for the float64 ITK restatement: index i of an N -> 2N 'fixed_size' resample reads the continuous
    index i/2; ITK's buffer ends at n - 0.5 (IsInsideBuffer is half-open), so the last output sample takes the default 0;
    integer outputs are static_cast (truncation toward zero), nearest is round-half-up."""
    a = np.array([0, 10, 20, 30], dtype=np.float32).reshape(1, 1, 4)
    assert O.itk_resample(a, (1, 1, 8), "linear").ravel().tolist() == [0, 5, 10, 15, 20, 25, 30, 0]
    b = np.array([-5, 0, 5, -10], dtype=np.int16).reshape(1, 1, 4)
    assert O.itk_resample(b, (1, 1, 8), "linear").ravel().tolist() == [-5, -2, 0, 2, 5, -2, -10, 0]      # -2.5 -> -2, 2.5 -> 2
    assert O.itk_resample(b, (1, 1, 8), "nearest").ravel().tolist() == [-5, 0, 0, 5, 5, -10, -10, 0]       # 0.5 -> 1, 1.5 -> 2, 2.5 -> 3
    c = np.arange(6, dtype=np.float32).reshape(1, 1, 6)                                                    # 6 -> 4: cidx = 1.5 i
    assert O.itk_resample(c, (1, 1, 4), "linear").ravel().tolist() == [0.0, 1.5, 3.0, 4.5]
    # anisotropic spacing, fixed_spacing mode: new_size = ceil(size * spacing / new_spacing)
    d = np.arange(5, dtype=np.float32).reshape(5, 1, 1)
    r = O.resample_to_spacing(d, (1.5, 1.0, 1.0), (1.0, 1.0, 1.0), "linear")
    assert r.shape == (8, 1, 1) and np.allclose(r.ravel()[:7], np.arange(7) / 1.5, atol=1e-12) and r.ravel()[7] == 0.0   # 7/1.5 = 4.67 >= 4.5


def test_otsu_restatements_match_opencv():
    """scikit-image is not in the image, so `threshold_otsu` (utils.py:239) cannot be called; OpenCV's THRESH_OTSU is an
    independent third-party implementation of the same published algorithm on 8-bit data (first maximum of the between-class
    variance over the 256-bin histogram; threshold t: values > t are foreground = skimage's bin centre for integer images).
    The oracle's restatement and the product's histogram form must both return exactly OpenCV's threshold."""
    cv2 = pytest.importorskip("cv2")
    import utils
    rs = np.random.RandomState(0)
    checked = 0
    for trial in range(200):
        n = int(rs.randint(50, 20000))
        kind = trial % 5
        if kind == 0:
            v = rs.randint(0, 256, n)
        elif kind == 1:
            v = np.concatenate([rs.normal(60, 15, n // 2), rs.normal(180, 25, n - n // 2)])
        elif kind == 2:
            v = rs.exponential(20, n)
        elif kind == 3:
            v = np.concatenate([rs.normal(10, 3, n * 9 // 10), rs.normal(240, 5, n - n * 9 // 10)])
        else:
            v = rs.randint(rs.randint(0, 100), rs.randint(101, 256), n)
        v = np.clip(v, 0, 255).astype(np.uint8)
        if len(np.unique(v)) < 2:
            continue
        t_cv, _ = cv2.threshold(v.reshape(-1, 1), 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert O.threshold_otsu_u8(v) == t_cv, (trial, O.threshold_otsu_u8(v), t_cv)
        assert utils.otsu_threshold_from_histogram(np.bincount(v, minlength=256)) == t_cv, trial
        checked += 1
    assert checked > 150


def test_itk_resample_interpolation_against_scipy():
    """No SimpleITK in the image.  The INTERPOLATION part of the float64 restatement (trilinear weights, edge clamping inside the
    half-voxel border, default value 0 outside the buffer) is checked against an independent implementation of linear / nearest
    interpolation, scipy.ndimage.map_coordinates, evaluated at the same ITK continuous indices; the coordinate convention itself
    is covered by the known-answer test above."""
    ndi = pytest.importorskip("scipy.ndimage")
    rs = np.random.RandomState(2)
    for shape, new_size, in_sp in [((10, 12, 9), (20, 24, 18), None), ((23, 17, 31), (16, 16, 16), (1.0, 0.7, 0.7)),
                                   ((8, 8, 8), (11, 5, 13), (2.5, 1.0, 0.5))]:
        a = rs.randn(*shape).astype(np.float32)
        in_sp = [1.0, 1.0, 1.0] if in_sp is None else list(in_sp)
        out_sp = O.fixed_size_spacing(shape, new_size, in_sp)
        cz, cy, cx = O.itk_continuous_indices(shape, new_size, in_sp, out_sp)
        zz, yy, xx = np.meshgrid(cz, cy, cx, indexing="ij")
        inside = ((zz >= -0.5) & (zz < shape[0] - 0.5) & (yy >= -0.5) & (yy < shape[1] - 0.5) & (xx >= -0.5) & (xx < shape[2] - 0.5))
        clip = lambda c, n: np.clip(c, 0.0, n - 1.0)              # ITK: no interpolation beyond the first / last sample
        ref = ndi.map_coordinates(a.astype(np.float64), [clip(zz, shape[0]), clip(yy, shape[1]), clip(xx, shape[2])], order=1,
                                  mode="nearest")
        ref[~inside] = 0.0
        got = O.itk_resample(a, new_size, "linear", in_spacing=in_sp)
        assert got.shape == tuple(new_size) and got.dtype == np.float32
        assert np.abs(got.astype(np.float64) - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
        m = (rs.rand(*shape) > 0.5).astype(np.uint8)
        nn = ndi.map_coordinates(m, [np.floor(clip(zz, shape[0]) + 0.5), np.floor(clip(yy, shape[1]) + 0.5),
                                     np.floor(clip(xx, shape[2]) + 0.5)], order=0, mode="nearest")
        nn[~inside] = 0
        assert np.array_equal(O.itk_resample(m, new_size, "nearest", in_spacing=in_sp), nn)

