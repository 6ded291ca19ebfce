# DC3DATGeneric (U-Net + RAM head + PCM attention refinement), used by process_pipeline.py — same names/values as the
# reference's st_dram_ref_att.py
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _common import *  # noqa: F401,F403
from _common import unet_model, logging_config

EXP_NAME = "st_dram_ref_rw"
NUM_EPOCHS = 300
WINDOWING_MAX = -700
MODEL = unet_model("models.DC3DATGeneric", at_spatial_size=(64, 64, 64), at_f_dim=8, at_g_dim=8, at_g_iter=1, at_k_size=3,
                   at_merge_type="scaled_dot_product_relu", at_self_loop=False, at_layers=[-1, 0, 1], at_p_enc_dim=0,
                   at_geo_f_dim=0)
LOGGING = logging_config(EXP_NAME)
PROCESSOR_LOGGING = logging_config(EXP_NAME, "processor_info.log")
