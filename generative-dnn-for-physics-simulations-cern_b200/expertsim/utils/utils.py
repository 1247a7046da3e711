"""Experiment-directory naming (reference expertsim/utils/utils.py:48-56)."""
import os
import time


def append_experiment_dir_to_cfg(cfg):
    """Adds cfg.train.dir_experiment / dir_models / dir_info under train.save_experiments_dir."""
    if cfg.train.get("checkpoint_experiment_dir"):
        base = cfg.train.checkpoint_experiment_dir
    else:
        stamp = time.strftime("%d_%m_%Y_%H_%M_%S")
        base = os.path.join(cfg.train.save_experiments_dir, f"{cfg.config.run_name}_{cfg.model.architecture}_{stamp}")
    cfg.train.dir_experiment = base
    cfg.train.dir_models = os.path.join(base, "models")
    cfg.train.dir_info = os.path.join(base, "info")
    return cfg
