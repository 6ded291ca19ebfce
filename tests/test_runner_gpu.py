"""GPU tests of the training runner's caller contract (job_runner.py:683-811): run() / validate() / evaluate_scan(), the
learning-rate schedule reaching a captured CUDA graph, and the data-parallel entry points under torchrun (2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def small_settings(name="st_dram_ref.py", fixture="dc3d_div8_16.pt"):
    from utils import Settings
    g = torch.load(os.path.join(GOLDEN, fixture))
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", name))
    s.MODEL = dict(g["cfg"])
    s.RESAMPLE_SIZE = (16, 16, 16)
    s.OPTIMIZER = dict(s.OPTIMIZER, lr=1e-3)
    return s, g


def val_scans(n, shape=(40, 56, 48)):
    from oracle_import import O
    out = []
    for i in range(n):
        scan, lobe, _, _ = O.synthetic_scan(shape, (1.0, 1.0, 1.0), seed=20 + i)
        out.append({"#image": scan, "#lobe_reference": lobe,
                    "meta": {"uid": f"val{i}", "spacing": (1.0, 1.0, 1.0), "cle": str(i % 6)}})
    return out


def test_evaluate_scan_matches_oracle_restatement():
    """LesionSegChunkTrain.evaluate_scan (job_runner.py:720-779): sigmoid(RAM) -> crop -> paste under the lobe -> lesion
    ratio -> CT severity class, against the CPU oracle's restatement of the same loop"""
    import job_runner
    from oracle_import import O
    s, g = small_settings()
    runner = job_runner.LesionSegChunkTrain(settings_module=s)
    runner.model.load_state_dict(g["state_dict"])
    runner.model.eval()
    sd = {k: v.clone() for k, v in g["state_dict"].items()}
    for data in val_scans(2):
        pred, target, seconds = runner.evaluate_scan(data)
        ref = O.infer_scan(lambda img, lb: O.dc3d_forward(sd, img, g["cfg"], False), data["#image"], data["#lobe_reference"],
                           np.asarray([1.0, 1.0, 1.0]), window=(s.WINDOWING_MIN, s.WINDOWING_MAX), chunk_size=(16, 16, 16))
        out = runner.run_scan(data["#image"], data["#lobe_reference"], [1.0, 1.0, 1.0], postprocess=False)
        assert np.array_equal(out["heatmap"] != 0, ref["heatmap"] != 0)
        assert abs(out["ratio"] - ref["ratio"]) <= 1e-4 * abs(ref["ratio"])
        assert pred == runner.loss_func.ratio_to_label([ref["ratio"]])[0]
        assert target == int(data["meta"]["cle"]) and seconds > 0


def test_run_epoch_loop_validates_schedules_and_checkpoints(tmp_path):
    """run(): epochs, validation on the reference's schedule, scheduler stepped once per validation epoch, checkpoints
    with the reference's dictionary keys, and a reload that restores epoch / iteration / weights"""
    import job_runner
    s, g = small_settings()
    s.MODEL_ROOT_PATH, s.DEBUG_PATH = str(tmp_path / "models"), str(tmp_path / "debug")
    s.NUM_EPOCHS, s.SYNTHETIC_STEPS, s.TRAIN_BATCH_SIZE, s.STATE_EPOCHS = 3, 2, 2, 1
    runner = job_runner.LesionSegChunkTrain(settings_module=s, val_dataset=val_scans(2))
    metrics = runner.run()
    assert runner.epoch_n == 2 and runner.current_iteration == 6
    assert set(metrics) >= {"val_time", "val_acc_reg_cls", "tr_loss"}
    lr = runner.optimizer.param_groups[0]["lr"]
    assert isinstance(lr, torch.Tensor) and lr.is_cuda
    assert abs(float(lr) - 1e-3 * 0.9 ** 3) < 1e-9                          # three validation epochs (epoch < 15)
    ck = torch.load(os.path.join(runner.exp_path, "2.pth"), map_location="cpu")
    assert set(ck) == {"epoch", "iteration", "model", "optimizer", "metrics"}
    s2, _ = small_settings()
    s2.MODEL_ROOT_PATH, s2.DEBUG_PATH, s2.RELOAD_CHECKPOINT = s.MODEL_ROOT_PATH, s.DEBUG_PATH, True
    again = job_runner.LesionSegChunkTrain(settings_module=s2)
    assert again.epoch_n == 2 and again.current_iteration == 6
    for k, v in runner.model.state_dict().items():
        assert torch.equal(v, again.model.state_dict()[k]), k
    s3, _ = small_settings()
    s3.MODEL_ROOT_PATH, s3.RELOAD_CHECKPOINT = str(tmp_path / "nothing_here"), True
    with pytest.raises(RuntimeError, match="no checkpoint files"):
        job_runner.LesionSegChunkTrain(settings_module=s3)


def test_scheduler_step_reaches_the_captured_graph(monkeypatch):
    """the learning rate is a device tensor, so ExponentialLR.step() after the step has been captured into a
    CUDA graph changes what the replays apply — graph and eager runs stay together across a decay"""
    import job_runner
    from oracle_import import O
    finals = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DRAM_CUDA_GRAPH", mode)
        s, g = small_settings()
        s.SCHEDULER = dict(s.SCHEDULER, gamma=0.1)
        runner = job_runner.LesionSegChunkTrain(settings_module=s)
        runner.model.load_state_dict(g["state_dict"])
        for step in range(6):
            images, lobes, lesions, ctsses = O.synthetic_batch(2, (16, 16, 16), seed=60 + step)
            runner.train_step({"#image": images[:, 0], "#lobe_reference": lobes[:, 0],
                               "#pseudo_lesion_reference": lesions[:, 0], "meta": {"cle": ctsses}})
            if step == 3:
                assert (runner._graph is not None) == (mode == "1")
                before = {k: v.detach().clone() for k, v in runner.model.named_parameters()}
                runner.scheduler.step()                                       # lr 1e-3 -> 1e-4, after the capture
        moved = max((v - before[k]).abs().max().item() for k, v in runner.model.named_parameters())
        finals[mode] = (moved, {k: v.detach().clone() for k, v in runner.model.named_parameters()})
    # two Adam steps at lr 1e-4 move a weight by at most ~2e-4 (+ slack); at the stale 1e-3 they would move it ~2e-3
    assert finals["1"][0] < 1e-3 and finals["0"][0] < 1e-3, (finals["0"][0], finals["1"][0])
    # graph and eager runs differ only by summation order; Adam turns a ~0 gradient's last-bit noise into a step of up to lr,
    # so the bound is the step budget of the six updates (4 x 1e-3 + 2 x 1e-4), not rounding error
    for k, v in finals["0"][1].items():
        assert (finals["1"][1][k] - v).abs().max().item() <= 5e-3, k


def test_pooling_global_max_and_avg():
    import models
    x = torch.randn(2, 3, 6, 5, 4, device="cuda")
    lungs = (torch.rand(2, 1, 6, 5, 4, device="cuda") > 0.3).float()
    assert torch.equal(models.pooling_dense_features(x, lungs, "global_max"), torch.amax(x, dim=(2, 3, 4)))
    ref = (x * lungs).sum(dim=(2, 3, 4)) / lungs.sum(dim=(2, 3, 4))
    assert torch.allclose(models.pooling_dense_features(x, lungs, "avg"), ref, rtol=1e-5, atol=1e-6)


def _torchrun(nproc, script, *args, env=None, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), script, *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env={**os.environ, **(env or {})})


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")


@needs2
@pytest.mark.parametrize("model", ["dc3d", "att"])
def test_data_parallel_parity_two_gpus(model):
    """tests/dp_parity.py: 2 ranks on shards of one global batch == one process on the global batch (loss, gradients after
    the all-reduce, BatchNorm running statistics)"""
    r = _torchrun(2, os.path.join(ROOT, "tests", "dp_parity.py"), *(["att"] if model == "att" else []))
    assert r.returncode == 0 and "DP PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@needs2
def test_torchrun_train_entry_point_keeps_replicas_identical(tmp_path):
    """`torchrun --nproc-per-node 2 train.py --synthetic 4`: the entry point initialises NCCL itself, every rank trains on its
    own shard and all replicas end with identical weights"""
    r = _torchrun(2, os.path.join(ROOT, "tests", "dp_entry.py"), "train", str(tmp_path))
    assert r.returncode == 0 and "REPLICAS IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@needs2
def test_torchrun_process_pipeline_writes_each_scan_once(tmp_path):
    r = _torchrun(2, os.path.join(ROOT, "tests", "dp_entry.py"), "pipeline", str(tmp_path))
    assert r.returncode == 0 and "EACH SCAN ONCE" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
