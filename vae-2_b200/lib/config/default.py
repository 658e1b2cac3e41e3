"""Config surface of the reference (lib/config/default.py:17-127), yacs-free.

Key names and default values are the reference's public configuration
surface; ``MODEL.EXTRA`` accepts new keys (STAGE1..4, FINAL_CONV_KERNEL, HD_Z,
Z_DIM come from yaml only; reference default.py:38).
"""
from .node import CfgNode as CN

_DEFAULTS = {
    "OUTPUT_DIR": "", "LOG_DIR": "", "GPUS": (0,), "WORKERS": 4, "PRINT_FREQ": 20,
    "AUTO_RESUME": False, "PIN_MEMORY": True, "RANK": 0,
    "CUDNN": {"BENCHMARK": True, "DETERMINISTIC": False, "ENABLED": True},
    "MODEL": {"NAME": "enc_hrnet", "PRETRAINED": "",
              "EXTRA": {"IS_BASELINE": False, "BASELINE_MODE": "VAE_NATIVE"}},
    "LOSS": {"USE_OHEM": False, "OHEMTHRES": 0.9, "OHEMKEEP": 100000, "CLASS_BALANCE": True},
    "DATASET": {"ROOT": "", "DATASET": "cityscapes", "NUM_CLASSES": 19, "TRAIN_SET": "",
                "EXTRA_TRAIN_SET": "", "TEST_SET": "", "FIXED_LENGTH": False},
    "TRAIN": {"IMAGE_SIZE": [512, 256], "BASE_SIZE": 512, "DOWNSAMPLERATE": 1, "FLIP": False,
              "MULTI_SCALE": False, "SCALE_FACTOR": 16, "CLIP_LENGTH": 3,
              "X1RECON_LAMBDA": 1.0, "X2RECON_LAMBDA": 0.1, "X3RECON_LAMBDA": 1.0,
              "GAN_LAMBDA": 1.0, "USE_X2RECON_MULTIPLIER": False,
              "LR_FACTOR": 0.1, "LR_STEP": [90, 110], "LR": 0.01, "EXTRA_LR": 0.001,
              "OPTIMIZER": "sgd", "MOMENTUM": 0.9, "WD": 0.0001, "NESTEROV": False,
              "IGNORE_LABEL": -1, "BEGIN_EPOCH": 0, "END_EPOCH": 484, "EXTRA_EPOCH": 0,
              "RESUME": False, "BATCH_SIZE_PER_GPU": 32, "SHUFFLE": True, "NUM_SAMPLES": 0},
    "TEST": {"IMAGE_SIZE": [512, 256], "BASE_SIZE": 512, "BATCH_SIZE_PER_GPU": 32,
             "NUM_SAMPLES": 0, "MODEL_FILE": "", "FLIP_TEST": False, "MULTI_SCALE": False,
             "CENTER_CROP_TEST": False, "SCALE_LIST": [1]},
    "DEBUG": {"DEBUG": False, "SAVE_BATCH_IMAGES_GT": False, "SAVE_BATCH_IMAGES_PRED": False,
              "SAVE_HEATMAPS_GT": False, "SAVE_HEATMAPS_PRED": False},
}


def _build():
    root = CN(_DEFAULTS)
    extra = CN(_DEFAULTS["MODEL"]["EXTRA"], new_allowed=True)
    root["MODEL"]["EXTRA"] = extra
    return root


_C = _build()


def get_cfg_defaults():
    return _build()


def update_config(cfg, args):
    """Same contract as reference lib/config/default.py:121-127."""
    cfg.defrost()
    cfg.merge_from_file(args.cfg)
    cfg.merge_from_list(getattr(args, "opts", None))
    cfg.freeze()


def load_config(yaml_path, opts=None):
    """Convenience: defaults <- yaml <- KEY VAL overrides."""
    cfg = get_cfg_defaults()
    cfg.merge_from_file(yaml_path)
    cfg.merge_from_list(opts)
    cfg.freeze()
    return cfg
