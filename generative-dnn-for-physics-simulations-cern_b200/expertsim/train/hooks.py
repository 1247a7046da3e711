"""Training callbacks.  Drop-in for expertsim/train/hooks.py of the reference: same ``Callback.on_epoch_end`` signature
(hooks.py:21-26), ``CheckpointSaver`` (saves when ``metrics[monitor] < ws_threshold``, hooks.py:102-136) and a
``WandBLogger`` that is active only when the wandb package is importable."""
import logging
import os

from .training_utils import save_models_and_architectures

logger = logging.getLogger(__name__)


class Callback:
    def on_epoch_end(self, epoch, metrics, moe, gen_optims, disc_optims, aux_reg_optim, router_optim):
        pass


class WandBLogger(Callback):
    def __init__(self, cfg):
        try:
            import wandb
        except ImportError as e:   # no silent no-op: the user asked for logging
            raise RuntimeError("wandb.log_experiments is set but the wandb package is not installed") from e
        self.wandb = wandb
        if getattr(cfg.wandb, "api_key", ""):
            wandb.login(key=cfg.wandb.api_key)
        self.run = wandb.init(name=cfg.config.run_name, config=cfg.to_dict() if hasattr(cfg, "to_dict") else dict(cfg))

    def on_epoch_end(self, epoch, metrics, moe, gen_optims, disc_optims, aux_reg_optim, router_optim):
        self.wandb.log({k: v for k, v in metrics.items() if isinstance(v, (int, float))}, step=epoch)


class CheckpointSaver(Callback):
    def __init__(self, dir_path, monitor="ws_mean", ws_threshold=3):
        self.dir_path, self.monitor, self.ws_threshold = dir_path, monitor, ws_threshold
        os.makedirs(dir_path, exist_ok=True)

    def on_epoch_end(self, epoch, metrics, moe, gen_optims, disc_optims, aux_reg_optim, router_optim):
        value = metrics.get(self.monitor)
        if value is None or not value < self.ws_threshold:
            return
        save_models_and_architectures(self.dir_path, moe.n_experts, moe.aux_regs, aux_reg_optim, moe.generators, gen_optims,
                                      moe.discriminators, disc_optims, moe.router, router_optim, epoch)
        logger.info("checkpoint saved at epoch %d (%s=%.4f)", epoch, self.monitor, value)


class MetricsTracker(Callback):
    def __init__(self):
        self.history = []

    def on_epoch_end(self, epoch, metrics, *args):
        self.history.append(dict(metrics))
