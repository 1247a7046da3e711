"""Command line entry point.  Drop-in for cli.py of the reference (cli.py:37-118): ``--config <yaml>`` and
``--override key=value ...``; wires config -> loaders -> train.  Unlike the reference it does not force
``CUDA_LAUNCH_BLOCKING=1`` / anomaly detection (cli.py:27-34): the step is asynchronous by design."""
import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from expertsim.config import load_config  # noqa: E402

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
logger = logging.getLogger(__name__)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Train the ExpertSim MoE GAN (B200-native hot path)")
    p.add_argument("--config", type=str, default=None, help="path to a YAML config (default: expertsim/config/default.yaml)")
    p.add_argument("--override", type=str, nargs="*", default=[], help="config overrides, e.g. model.n_experts=8 train.epochs=2")
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    cfg = load_config(args.config, args.override)
    from expertsim.utils.utils import append_experiment_dir_to_cfg
    cfg = append_experiment_dir_to_cfg(cfg)
    from expertsim.train.loop import train
    from expertsim.utils.data import get_train_test_data_loaders
    train_loader, test_loader = get_train_test_data_loaders(cfg)
    history = train(cfg, train_loader, test_loader)
    logger.info("done: %d epochs", len(history))
    return history


if __name__ == "__main__":
    main()
