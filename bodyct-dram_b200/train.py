"""Training entry point — same CLI as the reference's `dram/train.py` (2 positionals + 3 options, train.py:27-43).
Data loading from the institute's .mha archive is out of scope: batches come from `--synthetic N` steps of the
synthetic generator, or from a caller-supplied loader via `run_training_job(..., loader=...)`."""
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


def run_training_job(pretrain, lr, batch_size, smp, ckp_path, loader=None, synthetic_steps=0):
    settings = Settings(smp)
    settings.OPTIMIZER['lr'] = lr
    settings.TRAIN_BATCH_SIZE = batch_size
    settings.RELOAD_CHECKPOINT = bool(pretrain)
    settings.RELOAD_CHECKPOINT_PATH = ckp_path
    runner = get_callable_by_name(settings.JOB_RUNNER_CLS)(settings_module=settings)
    if settings.RELOAD_CHECKPOINT:
        runner.reload_model_from_cache(ckp_path)
    if loader is None and synthetic_steps > 0:
        loader = synthetic_loader(synthetic_steps, batch_size, tuple(settings.RESAMPLE_SIZE))
    if loader is not None:
        return runner.train(loader)
    return runner


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('pretrain', type=int, nargs='?', default=0, help="reload a checkpoint before training")
    parser.add_argument('lr', type=float, nargs='?', default=1e-3, help="learning rate")
    parser.add_argument('--batch_size', type=int, default=1)
    parser.add_argument('--smp', type=str, default=os.path.join(HERE, "exp_settings", "st_dram_ref.py"))
    parser.add_argument('--ckp_path', type=str, default=None)
    parser.add_argument('--synthetic', type=int, default=0, help="run N steps on synthetic lobe chunks")
    args = parser.parse_args()
    torch.backends.cudnn.benchmark = True
    print(run_training_job(args.pretrain, args.lr, args.batch_size, args.smp, args.ckp_path, synthetic_steps=args.synthetic))
