# DC3D (U-Net + dense RAM head), the default of train.py — same names/values as the reference's st_dram_ref.py
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _common import *  # noqa: F401,F403
from _common import unet_model, logging_config

EXP_NAME = "st_dram_ref"
NUM_EPOCHS = 200
WINDOWING_MAX = -300
MODEL = unet_model("models.DC3D")
LOGGING = logging_config(EXP_NAME)
PROCESSOR_LOGGING = logging_config(EXP_NAME, "processor_info.log")
