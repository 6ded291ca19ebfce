"""Full-CT inference entry point — same role and settings keys as the reference's `dram/process_pipeline.py`
(process_pipeline.py:10-26): load `exp_settings/st_dram_ref_att.py`, point MODEL_ROOT_PATH / DEBUG_PATH /
RELOAD_CHECKPOINT_PATH at the algorithm folder and run `LesionSegTest(image, lobe, out, settings, 'best.pth').run()`.
(The reference's own call raises TypeError as shipped — SURVEY D3; this runner has the 5-argument constructor.)"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from utils import Settings  # noqa: E402


def main(input_image_path="/input/images/ct/", input_lobe_path="/input/images/lobes/", output_path="/output/",
         algorithm_path=None, checkpoint="best.pth"):
    from job_runner import LesionSegTest
    settings = Settings(os.path.join(HERE, "exp_settings", "st_dram_ref_att.py"))
    algorithm_path = algorithm_path or os.environ.get("DRAM_ALGORITHM_PATH", os.path.join(HERE, "algorithm"))
    settings.MODEL_ROOT_PATH = algorithm_path
    settings.DEBUG_PATH = os.path.join(output_path, "debug")
    settings.RELOAD_CHECKPOINT_PATH = os.path.join(algorithm_path, settings.EXP_NAME, checkpoint)
    ckpt = settings.RELOAD_CHECKPOINT_PATH if os.path.exists(settings.RELOAD_CHECKPOINT_PATH) else None
    return LesionSegTest(input_image_path, input_lobe_path, output_path, settings, ckpt).run()


if __name__ == "__main__":
    print(main(*sys.argv[1:4]))
