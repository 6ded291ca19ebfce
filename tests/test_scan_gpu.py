"""GPU parity tests of the scan-level pre/post-processing kernels and of the whole-scan inference runner
(job_runner.LesionSegTest.run_scan) against the CPU oracle's restatement of job_runner.py:730-772 / 951-1015.
Index / mask / histogram work must be bit-exact; interpolated values within fp32 tolerance."""
import os

import numpy as np
import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ops():
    from dram_native import ops as o
    return o


def test_label_bboxes_match_find_crops():
    from oracle_import import O
    scan, lobe, lesion, spacing = O.synthetic_scan((40, 56, 48), (1.0, 0.7, 0.7), seed=1)
    boxes = ops().label_bboxes(torch.from_numpy(lobe).cuda(), 5).cpu().numpy()
    for label in range(1, 6):
        sl = O.find_crops(lobe == label, spacing, 0)
        assert [int(boxes[label, a]) for a in range(3)] == [s.start for s in sl]
        assert [int(boxes[label, 3 + a]) + 1 for a in range(3)] == [s.stop for s in sl]
    lobe[lobe == 3] = 0                                                      # empty label
    boxes = ops().label_bboxes(torch.from_numpy(lobe).cuda(), 5).cpu().numpy()
    assert boxes[3, 3] < 0


@pytest.mark.parametrize("chunk", [(16, 16, 16), (24, 20, 28)])
def test_lobe_chunk_preprocess(chunk):
    from oracle_import import O
    scan, lobe, lesion, spacing = O.synthetic_scan((40, 56, 48), (1.0, 0.7, 0.7), seed=2)
    scan_t, lobe_t = torch.from_numpy(scan).cuda(), torch.from_numpy(lobe).cuda()
    for label in (1, 4):
        sl, lobe_chunk, img_ref, msk_ref = O.preprocess_lobe_chunk(scan, lobe, label, spacing, (-1000, -700), chunk)
        img = torch.empty(chunk, device="cuda")
        msk = torch.empty(chunk, device="cuda")
        crop = tuple((s.start, s.stop) for s in sl)
        ops().lobe_chunk_preprocess(scan_t, lobe_t, label, crop, (-1000, -700), -2048, img, msk, spacing)
        assert torch.equal(msk.cpu(), torch.from_numpy(msk_ref)), "nearest-resampled lobe mask must be bit-exact"
        assert torch.equal(img.cpu(), torch.from_numpy(img_ref)), "the kernel follows ITK's double arithmetic operation by operation: bit-exact vs the float64 oracle"


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8])
@pytest.mark.parametrize("mode", ["linear", "nearest"])
@pytest.mark.parametrize("src,dst", [((20, 24, 18), (10, 16, 18)), ((9, 10, 11), (20, 31, 15)), ((12, 14, 10), (7, 9, 11))])
def test_itk_resample(dtype, mode, src, dst):
    """vs the float64 restatement of ITK's ResampleImageFilter / LinearInterpolate / NearestNeighbor (the
    oracle is written from the ITK definitions in double, the kernel must match it BIT-EXACTLY for every dtype)"""
    from oracle_import import O
    rng = np.random.RandomState(3)
    a = (rng.rand(*src) * 200 - 50).astype(dtype) if dtype != np.uint8 else rng.randint(0, 256, size=src).astype(np.uint8)
    for in_sp in (None, (1.0, 0.7, 0.7), (2.5, 0.68359375, 0.68359375)):
        ref = O.itk_resample(a, dst, mode, in_spacing=in_sp)
        got = ops().itk_resample(torch.from_numpy(a).cuda(), dst, mode, in_spacing=in_sp).cpu().numpy()
        assert got.shape == ref.shape and got.dtype == ref.dtype
        assert np.array_equal(got, ref), (dtype, mode, in_sp, np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.int16, torch.uint8])
@pytest.mark.parametrize("mode", ["linear", "nearest"])
def test_itk_resample_vector_kernel_equals_scalar_kernel(monkeypatch, dtype, mode):
    """4 outputs per thread with aligned vector stores (groups straddle row ends; ragged tail) == one output per thread"""
    torch.manual_seed(5)
    for src, dst, osp in (((23, 30, 26), (17, 21, 23), None), ((16, 20, 22), (16, 29, 31), (1.0, 0.7, 0.7)), ((5, 6, 7), (3, 3, 3), None)):
        a = (torch.rand(src) * 300 - 100).to(dtype).cuda() if dtype != torch.uint8 else torch.randint(0, 6, src, dtype=torch.uint8).cuda()
        monkeypatch.delenv("DRAM_RESAMPLE_SCALAR", raising=False)
        fast = ops().itk_resample(a, dst, mode, out_spacing=osp)
        monkeypatch.setenv("DRAM_RESAMPLE_SCALAR", "1")
        slow = ops().itk_resample(a, dst, mode, out_spacing=osp)
        assert torch.equal(fast, slow), (src, dst)


@pytest.mark.parametrize("mode,dtype", [("linear", np.int16), ("nearest", np.uint8), ("linear", np.float32)])
def test_itk_resample_fixed_spacing_and_back(mode, dtype):
    """Resample('fixed_spacing') of a scan to the 1 mm working grid (job_runner.py:827-835) and the resample back to the
    original grid (job_runner.py:1017-1030): sizes and every voxel equal to the float64 ITK restatement"""
    from oracle_import import O
    rng = np.random.RandomState(4)
    a = rng.randint(-1000, 400, size=(20, 30, 30)).astype(dtype) if dtype != np.uint8 else rng.randint(0, 6, size=(20, 30, 30)).astype(np.uint8)
    sp = (1.5, 0.7, 0.7)
    ref = O.resample_to_spacing(a, sp, (1.0, 1.0, 1.0), mode)
    assert ref.shape == (30, 21, 21)
    got = ops().itk_resample(torch.from_numpy(a).cuda(), ref.shape, mode, in_spacing=sp, out_spacing=(1.0, 1.0, 1.0))
    assert np.array_equal(got.cpu().numpy(), ref)
    back_ref = O.itk_resample(ref, a.shape, mode, in_spacing=(1.0, 1.0, 1.0), out_spacing=sp)
    back = ops().itk_resample(got, a.shape, mode, in_spacing=(1.0, 1.0, 1.0), out_spacing=sp)
    assert np.array_equal(back.cpu().numpy(), back_ref)


def test_masked_histogram_and_thresholds_are_integer_exact():
    rng = np.random.RandomState(5)
    heat = rng.rand(24, 20, 28).astype(np.float32)
    scan = rng.randint(-1500, 600, size=heat.shape).astype(np.int16)
    lobe = (rng.rand(*heat.shape) > 0.5).astype(np.uint8) * rng.randint(1, 6, size=heat.shape).astype(np.uint8)
    vessel = (rng.rand(*heat.shape) > 0.9).astype(np.uint8)
    o = ops()
    from oracle_import import O
    h = o.masked_hist_u8(torch.from_numpy(heat).cuda(), torch.from_numpy(lobe).cuda(), 0.0, 1.0).cpu().numpy()
    ref = np.bincount(O.windowing(heat[lobe > 0], from_span=(0, 1)).astype(np.uint8), minlength=256)
    assert np.array_equal(h, ref)
    w_scan = O.windowing(scan, to_span=(0, 1))
    h2 = o.masked_hist_u8(torch.from_numpy(scan).cuda(), torch.from_numpy(lobe).cuda(), -1150.0, 350.0).cpu().numpy()
    ref2 = np.bincount(O.windowing(w_scan[lobe > 0], from_span=(0, 1)).astype(np.uint8), minlength=256)
    assert np.array_equal(h2, ref2)
    th, th2 = 0.4392156862745098, 0.3
    lesion, post = o.threshold_masks(torch.from_numpy(heat).cuda(), th, torch.from_numpy(scan).cuda(),
                                     torch.from_numpy(vessel).cuda(), th2)
    ref_les = heat > np.float32(th)
    ref_post = ref_les & (w_scan > th2) & ~(vessel > 0)
    assert np.array_equal(lesion.cpu().numpy().astype(bool), ref_les)
    assert np.array_equal(post.cpu().numpy().astype(bool), ref_post)


@pytest.mark.parametrize("head", ["sigmoid", "literal"])
def test_run_scan_matches_oracle(head):
    """whole-scan inference on a small synthetic CT with the reference-initialised tiny attention model"""
    from oracle_import import O
    import job_runner
    import models
    from utils import Settings
    g = torch.load(os.path.join(GOLDEN, "dc3dat_div16_16.pt"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = Settings(os.path.join(root, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    s.MODEL = dict(g["cfg"])
    s.RESAMPLE_SIZE = (16, 16, 16)
    runner = job_runner.LesionSegTest(None, None, None, s, None, head=head)
    runner.model.load_state_dict(g["state_dict"])
    scan, lobe, lesion, spacing = O.synthetic_scan((40, 56, 48), (1.0, 1.0, 1.0), seed=6)
    out = runner.run_scan(scan, lobe, [1.0, 1.0, 1.0])

    sd = {k: v.clone() for k, v in g["state_dict"].items()}
    model_fn = lambda img, lb: O.dc3dat_forward(sd, img, g["cfg"], False)
    ref = O.infer_scan(model_fn, scan, lobe, np.asarray([1.0, 1.0, 1.0]), window=(s.WINDOWING_MIN, s.WINDOWING_MAX),
                       chunk_size=(16, 16, 16), head=head)
    assert np.array_equal(out["heatmap"] != 0, ref["heatmap"] != 0), "lobe masking / paste indices must be bit-exact"
    assert (out["heatmap"][lobe == 0] == 0).all()
    if head == "literal":
        assert (out["heatmap"] == 0).all() and (ref["heatmap"] == 0).all()
        return
    assert_close(torch.from_numpy(out["heatmap"]), torch.from_numpy(ref["heatmap"]), 2e-4, "heat map")
    assert abs(out["threshold"] - ref["threshold"]) < 1e-12, (out["threshold"], ref["threshold"])
    a, b = out["lesion"].astype(bool), ref["lesion"].astype(bool)
    dice = (2.0 * (a & b).sum() + 1e-5) / (a.sum() + b.sum() + 1e-5)
    assert dice >= 0.999, dice
    a, b = out["lesion_post"].astype(bool), ref["lesion_post"].astype(bool)
    assert (2.0 * (a & b).sum() + 1e-5) / (a.sum() + b.sum() + 1e-5) >= 0.999
    assert abs(out["ratio"] - ref["ratio"]) <= 1e-4 * abs(ref["ratio"])


@pytest.mark.parametrize("shape", [(40, 56, 48), (7, 9, 11), (3, 5, 2)])
def test_labelled_mean(shape):
    """lesion ratio (job_runner.py:772): mean of the heat map over the lobe voxels, deterministic"""
    torch.manual_seed(11)
    heat = torch.rand(shape, device="cuda")
    labels = (torch.rand(shape, device="cuda") > 0.6).to(torch.uint8) * torch.randint(1, 6, shape, device="cuda", dtype=torch.uint8)
    inside = labels > 0
    ref = (heat.double() * inside).sum() / inside.sum().clamp_min(1)
    got = ops().labelled_mean(heat, labels)
    assert got.dtype == torch.float32 and abs(got.item() - ref.item()) <= 1e-6 * max(ref.item(), 1e-6)
    assert ops().labelled_mean(heat, labels).item() == got.item()
    assert ops().labelled_mean(heat, torch.zeros_like(labels)).item() == 0.0


def test_read_small_goes_through_pinned_memory():
    """small results are stored into pinned host memory by a kernel (no cudaMemcpy): same values as .cpu(); misuse is loud"""
    from dram_native import lib
    o = ops()
    for t in (torch.arange(36, dtype=torch.int32, device="cuda").view(6, 6), torch.randint(0, 1 << 20, (256,), dtype=torch.int32, device="cuda"),
              torch.randn(1000, device="cuda")):
        assert torch.equal(o.read_small(t), t.cpu())
    with pytest.raises(lib.DramLibraryError):                       # pageable destination
        lib.check(lib.load().dram_store_to_host(torch.zeros(4, device="cuda").data_ptr(), torch.zeros(4).data_ptr(), 16, 0), "store_to_host")


def test_run_scans_pipeline_equals_one_scan_at_a_time():
    """LesionSegTest.run_scans (uploads / downloads on side streams, overlapped with the kernels) returns, scan by scan,
    exactly what the serial path (scan_to_masks) returns; scans of different sizes and spacings in one stream of work"""
    from oracle_import import O
    import job_runner
    from utils import Settings
    g = torch.load(os.path.join(GOLDEN, "dc3dat_div16_16.pt"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = Settings(os.path.join(root, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    s.MODEL = dict(g["cfg"])
    s.RESAMPLE_SIZE = (16, 16, 16)
    runner = job_runner.LesionSegTest(None, None, None, s, None)
    runner.model.load_state_dict(g["state_dict"])
    items = []
    for seed, shape, sp in ((1, (40, 56, 48), (1.0, 1.0, 1.0)), (2, (36, 64, 64), (1.0, 0.8, 0.8)), (3, (40, 56, 48), (1.0, 1.0, 1.0)),
                            (4, (30, 48, 40), (1.5, 1.0, 1.0))):
        scan, lobe, _, _ = O.synthetic_scan(shape, sp, seed=seed)
        items.append((torch.from_numpy(scan).pin_memory(), torch.from_numpy(lobe).pin_memory(), [float(v) for v in sp]))
    serial = []
    for scan_h, lobe_h, sp in items:
        les, post, ratio = runner.scan_to_masks(scan_h.cuda(), lobe_h.cuda(), sp)
        serial.append((les.cpu().clone(), post.cpu().clone(), float(ratio.item())))
    n = 0
    for (les, post, ratio), (rl, rp, rr) in zip(runner.run_scans(items), serial):
        assert les.shape == rl.shape and torch.equal(les, rl) and torch.equal(post, rp), f"scan {n}: masks differ"
        assert ratio == rr
        n += 1
    assert n == len(items)
    assert list(runner.run_scans([])) == []


def test_run_reads_and_writes_metaimage(tmp_path):
    """LesionSegTest.run() on the reference's file contract: <uid>.mha scan + lobe mask in, lesion / heat map / post .mha out
    (job_runner.py:857-890), equal to run_scan on the same arrays resampled the same way."""
    import job_runner
    import utils
    from oracle_import import O
    from utils import Settings
    g = torch.load(os.path.join(GOLDEN, "dc3dat_div16_16.pt"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = Settings(os.path.join(root, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    s.MODEL = dict(g["cfg"])
    s.RESAMPLE_SIZE = (16, 16, 16)
    scan, lobe, _, _ = O.synthetic_scan((40, 56, 48), (1.0, 0.8, 0.8), seed=7)
    (tmp_path / "scans").mkdir(); (tmp_path / "lobes").mkdir()
    geo = dict(origin=[-100.0, 20.5, 7.0], spacing=[0.8, 0.8, 1.0])             # ITK x-y-z
    utils.write_array_to_mha_itk(str(tmp_path / "scans"), [scan], ["case1"], type=np.int16, **geo)
    utils.write_array_to_mha_itk(str(tmp_path / "lobes"), [lobe], ["case1"], type=np.uint8, **geo)
    runner = job_runner.LesionSegTest(str(tmp_path / "scans"), str(tmp_path / "lobes"), str(tmp_path / "out"), s, None)
    runner.model.load_state_dict(g["state_dict"])
    records = runner.run()
    assert [r["uid"] for r in records] == ["case1"]
    les, meta = utils.read_mha(str(tmp_path / "out" / "test" / "case1.mha"))
    heat, _ = utils.read_mha(str(tmp_path / "out" / "test" / "heatmap" / "case1.mha"))
    post, _ = utils.read_mha(str(tmp_path / "out" / "test" / "post" / "case1.mha"))
    assert les.shape == scan.shape and les.dtype == np.uint8 and heat.shape == scan.shape and post.shape == scan.shape
    assert meta["spacing"] == geo["spacing"] and meta["origin"] == geo["origin"]
    dev = torch.device("cuda")
    sp = [1.0, 0.8, 0.8]
    scan_t = runner.resample_to_working_grid(torch.from_numpy(scan).to(dev), sp, "linear")
    lobe_t = runner.resample_to_working_grid(torch.from_numpy(lobe).to(dev), sp, "nearest")
    out = runner.run_scan(scan_t, lobe_t, [1.0, 1.0, 1.0], return_device=True)
    ref_les = job_runner.ops_itk_back(out["lesion"], scan.shape, [1.0, 1.0, 1.0], sp, "nearest").cpu().numpy()
    assert np.array_equal(les, ref_les)
    ref_heat = job_runner.ops_itk_back(out["heatmap"], scan.shape, [1.0, 1.0, 1.0], sp, "linear").cpu().numpy()
    assert np.array_equal(heat, utils.windowing(ref_heat, from_span=(0, 1)).astype(np.uint8))
    assert runner.run() == []                                                    # already archived -> skipped
