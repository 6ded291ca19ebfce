"""Data-parallel entry points on real GPUs (run under torchrun by tests/test_runner_gpu.py; not collected by pytest):
    torchrun --nproc-per-node 2 tests/dp_entry.py train    <tmp>   train.run_training_job: replicas identical after training
    torchrun --nproc-per-node 2 tests/dp_entry.py pipeline <tmp>   process_pipeline.main: every scan archived exactly once"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bodyct-dram_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))


def small_settings_file(tmp, name, fixture):
    """a settings file next to the shipped ones, reduced to the golden fixture's widths / 16^3 chunks"""
    src = os.path.join(ROOT, "bodyct-dram_b200", "exp_settings", name)
    g = torch.load(os.path.join(ROOT, "tests", "golden", fixture))
    path = os.path.join(tmp, "small_" + name)
    with open(path, "w") as f:
        f.write(f"import sys\nsys.path.insert(0, {os.path.dirname(src)!r})\nexec(open({src!r}).read())\n"
                f"MODEL = {dict(g['cfg'])!r}\nRESAMPLE_SIZE = (16, 16, 16)\nNUM_EPOCHS = 2\nSTATE_EPOCHS = 1\n"
                f"MODEL_ROOT_PATH = {os.path.join(tmp, 'models') + '/'!r}\nDEBUG_PATH = {os.path.join(tmp, 'debug') + '/'!r}\n")
    return path, g


def main():
    what, tmp = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ok = True
    if what == "train":
        import train
        smp, _ = small_settings_file(tmp, "st_dram_ref.py", "dc3d_div8_16.pt")
        runner = train.run_training_job(0, 1e-3, 2, smp, None, synthetic_steps=4)
        assert td.is_initialized() and td.get_world_size() == world
        h = hashlib.sha256()
        for k, v in sorted(runner.model.state_dict().items()):
            h.update(v.detach().cpu().numpy().tobytes())
        digests = [None] * world
        td.all_gather_object(digests, h.hexdigest())
        ok = len(set(digests)) == 1 and runner.current_iteration == 8
        if rank == 0:
            ok = ok and os.path.exists(os.path.join(runner.exp_path, "1.pth"))
            print("REPLICAS IDENTICAL" if ok else f"REPLICAS DIFFER {digests}", flush=True)
        runner.close()
    else:
        import process_pipeline
        import utils
        import dram_oracle as O
        scans, lobes, out = (os.path.join(tmp, d) for d in ("scans", "lobes", "out"))
        if rank == 0:
            for d in (scans, lobes):
                os.makedirs(d, exist_ok=True)
            for i in range(5):
                scan, lobe, _, _ = O.synthetic_scan((40, 56, 48), (1.0, 0.8, 0.8), seed=30 + i)
                geo = dict(origin=[0.0, 0.0, 0.0], spacing=[0.8, 0.8, 1.0])
                utils.write_array_to_mha_itk(scans, [scan], [f"case{i}"], type=np.int16, **geo)
                utils.write_array_to_mha_itk(lobes, [lobe], [f"case{i}"], type=np.uint8, **geo)
            open(os.path.join(tmp, "inputs_ready"), "w").close()
        # the pipeline initialises the process group itself: until then the ranks synchronise through the filesystem
        import time
        while not os.path.exists(os.path.join(tmp, "inputs_ready")):
            time.sleep(0.2)
        records = process_pipeline.main(scans, lobes, out, allow_random_weights=True)
        mine = sorted(r["uid"] for r in records)
        allr = [None] * world
        td.all_gather_object(allr, mine)
        flat = sorted(u for part in allr for u in part)
        ok = flat == [f"case{i}" for i in range(5)] and all(len(part) >= 2 for part in allr)
        ok = ok and all(os.path.exists(os.path.join(out, "test", f"case{i}.mha")) for i in range(5))
        if rank == 0:
            print("EACH SCAN ONCE" if ok else f"SHARDING WRONG {allr}", flush=True)
    td.barrier()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
