"""CPU tests of the host-side logic: module surface, settings, loss arithmetic (pure torch parts), error behaviour."""
import os

import numpy as np
import pytest
import torch

from oracle_import import O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_state_dict_surface_matches_reference_fixtures():
    import models
    for fixture in ("dc3d_div8_16.pt", "dc3dat_div16_16.pt"):
        g = torch.load(os.path.join(GOLDEN, fixture))
        cfg = dict(g["cfg"])
        cls = getattr(models, cfg.pop("method").split(".")[-1])
        m = cls(**cfg)
        own = m.state_dict()
        assert list(own.keys()) == list(g["state_dict"].keys())
        for k, v in g["state_dict"].items():
            assert tuple(own[k].shape) == tuple(v.shape), k
        m.load_state_dict(g["state_dict"])
        assert m.trace_path is None and hasattr(m, "dummy") and hasattr(m, "top_layer")


def test_full_config_parameter_count_and_init():
    import models
    from utils import Settings
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref.py"))
    cfg = dict(s.MODEL)
    assert cfg.pop("method") == "models.DC3D"
    torch.manual_seed(0)
    m = models.DC3D(**cfg)
    m.init(models.HeNorm(mode="fan_in"))
    assert sum(p.numel() for p in m.parameters()) == 16317921
    w = m.ds_modules[1].conv_blocks[0][0].weight
    assert abs(w.std().item() - (2.0 / (64 * 27)) ** 0.5) < 2e-3          # He normal, fan_in
    assert torch.all(m.top_layer.bias == 0.01)
    s2 = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    cfg2 = dict(s2.MODEL)
    cfg2.pop("method")
    m2 = models.DC3DATGeneric(**cfg2)
    assert sum(p.numel() for p in m2.parameters()) == 16317921 + 1897
    assert m2.attention_module.theta.weight.shape == (8, 17)


def test_unsupported_options_raise_instead_of_falling_back():
    import models
    import parts
    with pytest.raises(NotImplementedError):
        parts.ConvBlock5d([4, 4], [4, 4], 0, 5, False, 2)                 # kernel size 5
    with pytest.raises(NotImplementedError):
        parts.ConvBlock5d([4, 4], [4, 4], 0, 3, False, 1, dropout=0.5)
    with pytest.raises(NotImplementedError):
        parts.normal_wrapper("ln", 8)
    pcm = models.PCM((4, 4, 4), 17, 1, 8, 0, 8, 1, 3, "l2", False, p_enc_dim=0)
    with pytest.raises(NotImplementedError):
        pcm(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 17, 4, 4, 4))


def test_cpu_tensors_are_rejected_loudly():
    from dram_native import lib
    import models
    g = torch.load(os.path.join(GOLDEN, "dc3d_div8_16.pt"))
    cfg = dict(g["cfg"])
    cfg.pop("method")
    m = models.DC3D(**cfg)
    with pytest.raises(lib.DramLibraryError, match="no CPU fallback"):
        m(g["images"], g["lobes"])


def test_interval_targets_match_oracle():
    import metrics
    loss = metrics.IntRegRefineLoss(band_width=1e-2)
    rub = torch.tensor([0.0, 0.0005, 0.02, 0.2, 0.45, 0.9, 0.3, 0.004], dtype=torch.float32)
    ctss = ["0", "1", "2", "3", "4", "5", "0", "5"]
    got = loss.get_labels(ctss, rub)
    ref = O.interval_targets(ctss, rub, 1e-2)
    assert torch.equal(got, ref)


def test_boot_bce_matches_oracle_and_reference_structure():
    import metrics
    torch.manual_seed(0)
    p = torch.rand(2, 1, 6, 6, 6).clamp(0.01, 0.99).requires_grad_(True)
    voi = O.ellipsoid_lobe(2, (6, 6, 6), seed=1) > 0
    t = ((torch.rand(2, 1, 6, 6, 6) > 0.6) & voi).float()
    ref = O.boot_bce(p, t, voi, 0.1)
    gref, = torch.autograd.grad(ref, p)
    got = metrics.BootBinCrossEntropy(0.1)(p, t, voi)
    ggot, = torch.autograd.grad(got, p)
    assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert (ggot - gref).abs().max() <= 1e-6 * gref.abs().max()
    only_out = metrics.BootBinCrossEntropy(0.1)(p, torch.zeros_like(t), torch.zeros_like(voi))
    assert abs(only_out.item() - O.boot_bce(p, torch.zeros_like(t), torch.zeros_like(voi), 0.1).item()) < 1e-6


def test_settings_files_expose_the_reference_names():
    from utils import Settings
    for name, method in (("st_dram_ref.py", "models.DC3D"), ("st_dram_ref_att.py", "models.DC3DATGeneric")):
        s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", name))
        for key in ("MODEL", "INITIALIZER", "OPTIMIZER", "SCHEDULER", "LOSS_FUNC", "LOSS_FACTORS", "RESAMPLE_SIZE",
                    "WINDOWING_MAX", "WINDOWING_MIN", "PAD_VALUE", "TRAIN_BATCH_SIZE", "JOB_RUNNER_CLS", "TEST_JOB_RUNNER_CLS"):
            assert hasattr(s, key), key
        assert s.MODEL["method"] == method and s.RESAMPLE_SIZE == (80, 80, 80) and s.LOSS_FACTORS[:2] == [2.0, 1.0]


def test_utils_helpers():
    import utils
    a = np.random.RandomState(0).randint(-2048, 1000, size=(5, 6, 7)).astype(np.int16)
    assert np.array_equal(utils.windowing(a), O.windowing(a))
    mask = np.zeros((20, 30, 25), bool)
    mask[5:9, 10:22, 3:20] = True
    assert utils.find_crops(mask, (1.0, 0.7, 0.7), 5) == O.find_crops(mask, (1.0, 0.7, 0.7), 5)
    v = np.random.RandomState(1).rand(4000)
    assert abs(utils.binary_cam(v)[1] - O.binary_cam(v)) < 1e-12


def test_kernel_schedule_of_the_benchmark_layers(monkeypatch):
    """Which tensor-core kernel runs each convolution of the DC3D training step at the benchmark size (batch 8, 80^3 chunks):
    host logic of the library only (no device work), asserted here so that a change of the selection rules is a visible
    change.  forward / dgrad: 4 = SM pairs (cta_group::2) with kw re-use, 5 = SM pairs on generic (5,5,5) tiles, 3 = channels
    on M (ds0.c1's forward: Cin = 32, two K steps per tap);  wgrad: 1 = kw re-use (Cout <= 64), 2 = SM pairs."""
    from dram_native import lib
    monkeypatch.delenv("DRAM_CONV_V4", raising=False)
    monkeypatch.delenv("DRAM_WGRAD_V2", raising=False)
    L = lib.load()
    pad = lambda c: (c + 63) // 64 * 64
    layers = [("ds0.c1", 32, 64, 80), ("ds1.c0", 64, 64, 40), ("ds1.c1", 64, 128, 40), ("ds2.c0", 128, 128, 20),
              ("ds2.c1", 128, 256, 20), ("bg.c0", 256, 256, 10), ("bg.c1", 256, 512, 10), ("us0.c0", 768, 256, 20),
              ("us0.c1", 256, 256, 20), ("us1.c0", 384, 128, 40), ("us1.c1", 128, 128, 40), ("us2.c0", 192, 64, 80),
              ("us2.c1", 64, 64, 80)]
    got = {}
    for name, ci, co, s in layers:
        fwd = L.dram_conv3d_umma_fwd_kernel(8, s, s, s, ci, pad(ci), co, 3, 1, 1)
        dgrad = L.dram_conv3d_umma_fwd_kernel(8, s, s, s, co, pad(co), ci, 3, 1, 1)       # the same kernel on dy, channels swapped
        wgrad = L.dram_conv3d_umma_wgrad_kernel(s, s, pad(co), 3, 1, 1)
        got[name] = (fwd, dgrad, wgrad)
    assert got == {"ds0.c1": (3, 4, 1), "ds1.c0": (4, 4, 1), "ds1.c1": (4, 4, 2), "ds2.c0": (5, 5, 2), "ds2.c1": (5, 5, 2),
                   "bg.c0": (5, 5, 2), "bg.c1": (5, 5, 2), "us0.c0": (5, 5, 2), "us0.c1": (5, 5, 2), "us1.c0": (4, 4, 2),
                   "us1.c1": (4, 4, 2), "us2.c0": (4, 4, 1), "us2.c1": (4, 4, 1)}, got
    monkeypatch.setenv("DRAM_CONV_V4", "0")                    # the single-SM kernels the tests compare against
    assert L.dram_conv3d_umma_fwd_kernel(8, 80, 80, 80, 192, 192, 64, 3, 1, 1) == 2
    assert L.dram_conv3d_umma_fwd_kernel(8, 40, 40, 40, 384, 384, 128, 3, 1, 1) == 3
    assert L.dram_conv3d_umma_fwd_kernel(8, 20, 20, 20, 256, 256, 256, 3, 1, 1) == 0
    monkeypatch.delenv("DRAM_CONV_V4")
    # the one-plane gradient backward (DRAM_BWD_PRECISION=bf16x2) stays on the single-SM kernels
    assert L.dram_conv3d_umma_fwd_kernel(8, 80, 80, 80, 64, 64, 64, 3, 0, 1) == 3
    assert L.dram_conv3d_umma_wgrad_kernel(40, 40, 128, 3, 1, 0) == 0
    assert L.dram_conv3d_umma_fwd_kernel(0, 1, 1, 1, 1, 64, 64, 3, 1, 1) == -1


def test_metaimage_round_trip(tmp_path):
    """utils.write_array_to_mha_itk (utils.py:142-159) / read_mha: voxels, element type and geometry survive, compressed
    and uncompressed; the header is what ITK's MetaImageIO writes for a 3-d image."""
    import utils
    rs = np.random.RandomState(3)
    for dtype in (np.uint8, np.int16, np.float32):
        a = (rs.rand(5, 6, 7) * 200 - 50).astype(dtype)
        utils.write_array_to_mha_itk(str(tmp_path), [a], ["vol"], type=dtype, origin=[1.5, -2.0, 3.25],
                                     spacing=[0.7, 0.7, 1.25], direction=[0, 1, 0, 1, 0, 0, 0, 0, 1])
        b, meta = utils.read_mha(str(tmp_path / "vol.mha"))
        assert b.dtype == dtype and np.array_equal(a, b)
        assert meta["spacing"] == [0.7, 0.7, 1.25] and meta["origin"] == [1.5, -2.0, 3.25]
        assert meta["direction"] == [0, 1, 0, 1, 0, 0, 0, 0, 1]
        head = open(tmp_path / "vol.mha", "rb").read(400).decode("latin1")
        assert "DimSize = 7 6 5" in head and "CompressedData = True" in head and "ElementDataFile = LOCAL" in head
        utils.write_mha(str(tmp_path / "raw.mha"), a, compress=False)
        c, _ = utils.read_mha(str(tmp_path / "raw.mha"))
        assert np.array_equal(a, c)


def test_metaimage_reader_on_a_file_assembled_from_the_format_description(tmp_path):
    """A MetaImage put together byte by byte in this test from the published MetaIO format — the header ITK 4.13's MetaImageIO
    writes for a 3-d image (key order, `True`/`False`, `CompressedDataSize`, x-fastest little-endian voxels, zlib stream) —
    not by utils.write_mha: read_mha must return the voxels at [z, y, x], the element type and the geometry, and write_mha must
    emit the same keys in the same order.  (No ITK in the image: the fixture restates the format, it is not an ITK-written file.)"""
    import struct
    import zlib
    import utils
    nx, ny, nz = 5, 4, 3
    vox = [(x - 2) * 100 + y * 10 + z - 1000 for z in range(nz) for y in range(ny) for x in range(nx)]      # x runs fastest
    cases = [("MET_SHORT", "<h", np.int16, True), ("MET_UCHAR", "<B", np.uint8, False), ("MET_FLOAT", "<f", np.float32, True)]
    for met, fmt, dtype, compressed in cases:
        vals = [v % 251 for v in vox] if dtype == np.uint8 else vox
        payload = b"".join(struct.pack(fmt, v) for v in vals)
        data = zlib.compress(payload) if compressed else payload
        lines = ["ObjectType = Image", "NDims = 3", "BinaryData = True", "BinaryDataByteOrderMSB = False",
                 f"CompressedData = {compressed}"] + ([f"CompressedDataSize = {len(data)}"] if compressed else []) + [
                 "TransformMatrix = 1 0 0 0 1 0 0 0 1", "Offset = -170.5 -180.25 -300", "CenterOfRotation = 0 0 0",
                 "AnatomicalOrientation = RAI", "ElementSpacing = 0.7 0.7 1.25", f"DimSize = {nx} {ny} {nz}",
                 f"ElementType = {met}", "ElementDataFile = LOCAL"]
        path = tmp_path / f"{met}.mha"
        path.write_bytes(("\n".join(lines) + "\n").encode("ascii") + data)
        arr, meta = utils.read_mha(str(path))
        assert arr.dtype == dtype and arr.shape == (nz, ny, nx)
        for z in range(nz):
            for y in range(ny):
                for x in range(nx):
                    assert arr[z, y, x] == dtype(vals[(z * ny + y) * nx + x])
        assert meta["spacing"] == [0.7, 0.7, 1.25] and meta["origin"] == [-170.5, -180.25, -300.0]
        assert meta["direction"] == [1, 0, 0, 0, 1, 0, 0, 0, 1]
        # the writer's header: the same keys in the same order, and the same bytes after it
        utils.write_mha(str(tmp_path / "w.mha"), arr, spacing=meta["spacing"], origin=meta["origin"], direction=meta["direction"],
                        compress=compressed)
        raw = (tmp_path / "w.mha").read_bytes()
        head = raw[:raw.index(b"ElementDataFile = LOCAL\n") + len(b"ElementDataFile = LOCAL\n")].decode("ascii")
        assert [ln.split(" = ")[0] for ln in head.strip().split("\n")] == [ln.split(" = ")[0] for ln in lines]
        body = raw[len(head):]
        assert (zlib.decompress(body) if compressed else body) == payload


@pytest.mark.parametrize("workload", ["train", "infer"])
def test_bench_reference_arm_prints_one_json_line(workload):
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): exactly ONE stdout line, the contract's
    keys, the metric / unit / workload string of the B200 arm, no GPU needed"""
    import json
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    if workload == "train":
        assert d["metric"] == "train_lobe_chunks_per_s" and d["config"]["workload"] == bench.TRAIN_WORKLOAD.format(B=8)
    else:
        assert d["metric"] == "infer_lobe_chunks_per_s" and d["config"]["workload"] == bench.INFER_WORKLOAD


def test_process_pipeline_refuses_to_run_without_a_checkpoint(tmp_path):
    """a missing best.pth must not silently run the pipeline on random weights; default folders are the reference's"""
    import inspect
    import process_pipeline
    sig = inspect.signature(process_pipeline.main)
    assert sig.parameters["input_lobe_path"].default == "/input/images/pulmonary-lobes/"      # process_pipeline.py:14-16
    assert sig.parameters["output_path"].default == "/output/images/"
    with pytest.raises(RuntimeError, match="does not exist"):
        process_pipeline.main(str(tmp_path), str(tmp_path), str(tmp_path / "out"), algorithm_path=str(tmp_path))
