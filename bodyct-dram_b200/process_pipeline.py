"""Full-CT inference entry point — same role, default folders and settings keys as the reference's
`dram/process_pipeline.py` (process_pipeline.py:10-26): load `exp_settings/st_dram_ref_att.py`, point MODEL_ROOT_PATH /
DEBUG_PATH / RELOAD_CHECKPOINT_PATH at the run and call `LesionSegTest(image, lobe, out, settings, 'best.pth').run()`.
(The reference's own call raises TypeError as shipped — SURVEY D3; this runner has the 5-argument constructor.)

The checkpoint is looked up as `<algorithm_path>/<EXP_NAME>/best.pth` (algorithm_path: argument, $DRAM_ALGORITHM_PATH or
the output folder like the reference's MODEL_ROOT_PATH = output_path); a missing checkpoint RAISES — pass
`allow_random_weights=True` (tests / benchmarks only) to run the pipeline on the HeNorm initialisation.
Under `torchrun --nproc-per-node N process_pipeline.py` the scans are sharded across the GPUs (no collective)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from utils import Settings  # noqa: E402


def main(input_image_path="/input/images/ct/", input_lobe_path="/input/images/pulmonary-lobes/",
         output_path="/output/images/", algorithm_path=None, checkpoint="best.pth", allow_random_weights=False):
    from job_runner import LesionSegTest
    from train import init_distributed
    init_distributed()
    settings = Settings(os.path.join(HERE, "exp_settings", "st_dram_ref_att.py"))
    algorithm_path = algorithm_path or os.environ.get("DRAM_ALGORITHM_PATH", output_path)
    settings.MODEL_ROOT_PATH = algorithm_path
    settings.DEBUG_PATH = output_path
    settings.RELOAD_CHECKPOINT_PATH = os.path.join(algorithm_path, settings.EXP_NAME, checkpoint)
    ckpt = settings.RELOAD_CHECKPOINT_PATH
    if not os.path.exists(ckpt):
        if not allow_random_weights:
            raise RuntimeError(f"checkpoint {ckpt} does not exist (set DRAM_ALGORITHM_PATH or pass algorithm_path; "
                               "allow_random_weights=True runs on the random initialisation)")
        ckpt = None
    return LesionSegTest(input_image_path, input_lobe_path, output_path, settings, ckpt).run()


if __name__ == "__main__":
    print(main(*sys.argv[1:4], allow_random_weights=os.environ.get("DRAM_ALLOW_RANDOM_WEIGHTS") == "1"))
