"""world_size-2 gloo tests (CPU) of the data-parallel host logic: gradient bucketing/all-reduce, SyncBN statistic
exchange, and the loss's batch-global normalisers (sum of the ranks' losses == single-process loss on the global batch)."""
import os
import sys

import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in ("bodyct-dram_b200", "oracle"):
        sys.path.insert(0, os.path.join(ROOT, p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from dram_native import dist as ddist
    import dram_oracle as O
    import metrics
    torch.manual_seed(0)
    out = {}
    # --- gradient reducer: SUM over ranks, buckets in reverse order, parameters without grad handled
    params = [torch.nn.Parameter(torch.randn(n)) for n in (5, 70000, 3, 9)]
    red = ddist.GradReducer(params, bucket_mb=0.1)
    assert len(red.buckets) >= 2
    loss = sum((p * (rank + 1)).sum() for p in params[:3])            # params[3] gets no gradient
    loss.backward()
    red.finish()
    out["grads_ok"] = all(torch.allclose(p.grad, torch.full_like(p, 3.0)) for p in params[:3]) and \
        torch.equal(params[3].grad, torch.zeros(9))
    # --- SyncBN statistic exchange
    sums = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
    cnt = ddist.allreduce_stats(sums, 10)
    out["stats_ok"] = bool(torch.equal(sums, torch.tensor([3.0, 4.0], dtype=torch.float64)) and cnt == 20.0)
    local = torch.tensor([1.0, 1.0], dtype=torch.float64)
    g = ddist.allreduce_sums(local)
    out["sums_ok"] = bool(torch.equal(g, torch.tensor([2.0, 2.0], dtype=torch.float64)) and torch.equal(local, torch.ones(2, dtype=torch.float64)))
    # --- loss normalisers: global batch of 4, two per rank
    gen = torch.Generator().manual_seed(5)
    p_all = torch.rand(4, 1, 6, 6, 6, generator=gen).clamp(0.02, 0.98)
    voi_all = O.ellipsoid_lobe(4, (6, 6, 6), seed=2) > 0
    t_all = ((torch.rand(4, 1, 6, 6, 6, generator=gen) > 0.5) & voi_all).float()
    sl = slice(2 * rank, 2 * rank + 2)
    p = p_all[sl].clone().requires_grad_(True)
    mine = metrics.BootBinCrossEntropy(0.1)(p, t_all[sl], voi_all[sl])
    gmine, = torch.autograd.grad(mine, p)
    tot = mine.detach().clone()
    td.all_reduce(tot)
    p_ref = p_all.clone().requires_grad_(True)
    ref = O.boot_bce(p_ref, t_all, voi_all, 0.1)
    gref, = torch.autograd.grad(ref, p_ref)
    out["loss_ok"] = bool(abs(tot.item() - ref.item()) < 1e-5 * abs(ref.item()))
    out["loss_grad_ok"] = bool((gmine - gref[sl]).abs().max() < 1e-6 * gref.abs().max())
    # --- work sharding of the inference entry points: every item exactly once, no collective on the data path
    items = [f"case{i}" for i in range(7)]
    mine = ddist.shard(items)
    every = ddist.all_gather_object(mine)
    out["shard_ok"] = sorted(u for part in every for u in part) == items and mine == items[rank::world]
    try:
        ddist.check_uniform_batch(((2, 16, 16, 16),) * 3)
        out["uniform_ok"] = True
    except RuntimeError:
        out["uniform_ok"] = False
    try:
        ddist.check_uniform_batch(((2 + rank, 16, 16, 16),) * 3)
        out["ragged_raises"] = False
    except RuntimeError:
        out["ragged_raises"] = True
    # --- LesionSegTest.run_mha: the sorted scan list is sharded rank::world, already archived scans are skipped
    import job_runner
    from utils import Settings
    tmp = os.environ["DRAM_TEST_TMP"]
    s = Settings(os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", "st_dram_ref_att.py"))
    runner = job_runner.LesionSegTest(os.path.join(tmp, "scans"), os.path.join(tmp, "lobes"), os.path.join(tmp, "out"), s, None,
                                      build_model=False)

    def fake_process(path):                      # stands in for the GPU work: archive a marker under the scan's uid
        uid = os.path.splitext(os.path.basename(path))[0]
        os.makedirs(os.path.join(tmp, "out", "test"), exist_ok=True)
        with open(os.path.join(tmp, "out", "test", uid + ".mha"), "x") as f:          # "x": fails if written twice
            f.write(str(rank))
        return {"uid": uid, "seconds": 0.0, "ratio": 0.0}

    runner._process_mha = fake_process
    recs = runner.run_mha()
    td.barrier()
    done = sorted(os.listdir(os.path.join(tmp, "out", "test")))
    out["run_mha_ok"] = done == [f"s{i}.mha" for i in range(5)] and [r["uid"] for r in recs] == [f"s{i}" for i in range(5)][rank::world]
    out["rerun_skips"] = runner.run_mha() == []
    q.put((rank, out))
    td.destroy_process_group()


def test_data_parallel_host_logic_world_size_2(tmp_path):
    os.makedirs(tmp_path / "scans")
    for i in range(5):
        (tmp_path / "scans" / f"s{i}.mha").write_text("")
    os.environ["DRAM_TEST_TMP"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results:
        assert all(out.values()), (rank, out)
