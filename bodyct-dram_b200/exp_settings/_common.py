"""Values shared by the two shipped experiment settings.  The UPPERCASE names and their values are the reference's
configuration API (dram/exp_settings/st_dram_ref.py, st_dram_ref_att.py); institute-specific paths default to a local
working directory and can be overridden through the DRAM_* environment variables."""
import os

_root = os.environ.get("DRAM_EXP_ROOT", os.path.join(os.getcwd(), "dram_experiments"))

COPY_DATA, ON_PREMISE_LOCATION = False, None
RELOAD_CHECKPOINT, IS_CUDA, RELOAD_CHECKPOINT_PATH, RELOAD_DICT_LIST = False, True, None, ["model"]
DB_PATH = os.environ.get("DRAM_DB_PATH", _root + "/")
TEST_CSV, TRAIN_CSV, VALID_CSV = DB_PATH + "test.csv", DB_PATH + "wss_train.csv", DB_PATH + "val.csv"
DEBUG_PATH, MODEL_ROOT_PATH = DB_PATH + "test_cases/", DB_PATH + "models/"
JOB_RUNNER_CLS, TEST_JOB_RUNNER_CLS = "job_runner.LesionSegChunkTrain", "job_runner.LesionSegTest"

RESAMPLE_MODE = "fixed_size"
VAL_EPOCHS = STATE_EPOCHS = 10
NUM_WORKERS, LOG_STEPS, AUG_RATIO, BALANCED_LABEL_COUNT, TRAIN_BATCH_SIZE = 0, 1, 0.0, 200, 10
RESAMPLE_SPACING = TEST_RESAMPLE_SPACING = 1.0
RESAMPLE_SIZE = (80, 80, 80)
LOSS_FACTORS = [2.0, 1.0, 0.5, 0.5]
RELABEL_MAPPING = {}
LABEL_NAME_MAPPING = {0: 'background', 1: 'emphysema'}
CLASS_WEIGHTS = [0.65, 0.7, 0.7, 0.75, 0.75, 0.8]
PAD_VALUE, WINDOWING_MIN, NR_CLASS = -2048, -1000, 1


def unet_model(method, **extra):
    pairs = 7
    cfg = {"method": method, "n_layers": 3,
           "in_ch_list": [1, 64, 128, 256, 768, 384, 192],
           "base_ch_list": [32, 64, 128, 256, 256, 128, 64],
           "end_ch_list": [64, 128, 256, 512, 256, 128, 64],
           "kernel_sizes": [(3, 3)] * pairs, "stacking": 3, "padding_list": [(1, 1)] * pairs,
           "checkpoint_layers": [0, 1, 0, 1, 0, 1, 0], "dropout": 0.0,
           "upsample_ksize": (3, 3, 3), "upsample_sf": (2, 2, 2), "out_ch": NR_CLASS}
    cfg.update(extra)
    return cfg


TEST_MERGE_PROTOCOLS = [(None, None, None, None)]
INITIALIZER = {"method": "models.HeNorm", "mode": "fan_in"}
OPTIMIZER = {"method": "torch.optim.Adam", "lr": 0.0001}
SCHEDULER = {"method": "torch.optim.lr_scheduler.ExponentialLR", "gamma": 0.9}
LOSS_FUNC = {"method": "metrics.IntRegRefineLoss", "band_width": 1e-2, "smoothing": 0.1}


def logging_config(exp_name, file_name="info.log"):
    fmt = {'standard': {'format': '%(asctime)s [%(levelname)s] %(name)s: %(message)s'}}
    handlers = {'console': {'level': 'INFO', 'formatter': 'standard', 'class': 'logging.StreamHandler',
                            'stream': 'ext://sys.stdout'},
                'file_handler': {'class': 'logging.handlers.RotatingFileHandler', 'level': 'INFO', 'formatter': 'standard',
                                 'filename': "{}/{}/{}".format(MODEL_ROOT_PATH, exp_name, file_name),
                                 'maxBytes': 10485760, 'backupCount': 20, 'encoding': 'utf8'}}
    return {'version': 1, 'disable_existing_loggers': False, 'formatters': fmt, 'handlers': handlers,
            'loggers': {'': {'handlers': ['console', 'file_handler'], 'level': 'INFO', 'propagate': True}}}
