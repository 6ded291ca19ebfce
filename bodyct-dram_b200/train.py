"""Training entry point — same CLI as the reference's `dram/train.py` (2 positionals + 3 options, train.py:27-43) and the
same flow: Settings(--smp) -> overrides -> `get_callable_by_name(JOB_RUNNER_CLS)(settings_module=settings).run()`.

Data loading from the institute's .mha archive is out of scope: batches come from `--synthetic N` synthetic
steps per epoch, or from a caller-supplied `loader_factory` via `run_training_job(..., loader_factory=...)`.

Data parallel (new functionality): launched under `torchrun --nproc-per-node N train.py ...` the process
group is initialised from RANK / LOCAL_RANK / WORLD_SIZE (NCCL), every rank owns one GPU and draws its own shard of the
data; gradients, BatchNorm statistics and the loss normalisers are all-reduced inside the step (dram_native/dist.py), so
the replicas hold identical weights after every step."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import torch  # noqa: E402

from utils import Settings, get_callable_by_name  # noqa: E402


def synthetic_loader(steps, batch_size, size=(80, 80, 80), seed=0):
    g = torch.Generator().manual_seed(seed)
    zz, yy, xx = torch.meshgrid(*[torch.arange(s, dtype=torch.float32) for s in size], indexing="ij")
    for _ in range(steps):
        c = [(s - 1) / 2.0 for s in size]
        lobe = ((((zz - c[0]) / (0.45 * size[0])) ** 2 + ((yy - c[1]) / (0.45 * size[1])) ** 2
                 + ((xx - c[2]) / (0.45 * size[2])) ** 2) <= 1.0).float()
        lobes = lobe.expand(batch_size, *size).contiguous()
        images = torch.rand(batch_size, *size, generator=g) * lobes
        lesions = (torch.rand(batch_size, *size, generator=g) > 0.7).float() * lobes
        yield {"#image": images, "#lobe_reference": lobes, "#pseudo_lesion_reference": lesions,
               "meta": {"cle": [str(i % 6) for i in range(batch_size)]}}


def init_distributed():
    """One process per GPU under torchrun: NCCL process group from the environment; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or not torch.distributed.is_available() or torch.distributed.is_initialized():
        return world
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world


def run_training_job(pretrain, lr, batch_size, smp, ckp_path, loader=None, synthetic_steps=0, loader_factory=None,
                     val_dataset=None, epochs=None):
    """train.py:11-24.  Returns the runner after `run()` (or after one pass over an explicit `loader`)."""
    init_distributed()
    settings = Settings(smp)
    settings.OPTIMIZER['lr'] = lr
    settings.TRAIN_BATCH_SIZE = batch_size
    settings.RELOAD_CHECKPOINT = bool(pretrain)
    settings.RELOAD_CHECKPOINT_PATH = ckp_path
    settings.SYNTHETIC_STEPS = synthetic_steps
    if epochs is not None:
        settings.NUM_EPOCHS = epochs
    runner_cls = get_callable_by_name(settings.JOB_RUNNER_CLS)
    runner = runner_cls(settings_module=settings, loader_factory=loader_factory, val_dataset=val_dataset)
    if loader is not None:
        runner.train(loader)
    else:
        runner.run()
    return runner


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('pretrain', type=int, nargs='?', default=0, help="if use pretrained model.")
    parser.add_argument('lr', type=float, nargs='?', default=1e-3, help="set up learning rate.")
    parser.add_argument('--batch_size', type=int, nargs='?', default=1)
    parser.add_argument('--smp', type=str, nargs='?', default=os.path.join(HERE, "exp_settings", "st_dram_ref.py"))
    parser.add_argument('--ckp_path', type=str, default=None, help='set checkpoint path.')
    parser.add_argument('--synthetic', type=int, default=0, help="synthetic lobe-chunk batches per epoch (no dataset on this path)")
    parser.add_argument('--epochs', type=int, default=None, help="override settings.NUM_EPOCHS")
    args = parser.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.deterministic = True
    r = run_training_job(args.pretrain, args.lr, args.batch_size, args.smp, args.ckp_path, synthetic_steps=args.synthetic,
                         epochs=args.epochs)
    print({"epoch": r.epoch_n, "iteration": r.current_iteration, **r.metrics})
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        r.close()                                   # the captured step holds NCCL work: drop it before the group goes away
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
