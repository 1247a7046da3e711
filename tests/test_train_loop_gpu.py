"""The drop-in training surface end to end on the GPU: cli.main -> load_config -> loaders -> train() -> train_epoch /
evaluate_epoch (Wasserstein metric) -> callbacks (checkpoint files with the reference's names)."""
import importlib.util
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli():
    spec = importlib.util.spec_from_file_location("es_cli", os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200", "cli.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("arch,shape", [("proton", "[56,30]"), ("neutron", "[44,44]")])
def test_cli_trains_two_epochs_on_synthetic_showers(tmp_path, arch, shape):
    hist = _cli().main(["--override", f"model.architecture={arch}", f"dataset.zdc_type={arch}", f"dataset.input_image_shape={shape}",
                        "dataset.synthetic_samples=640", "train.epochs=2", "train.batch_size=128", "model.n_experts=3",
                        "train.save_experiment_data=True", f"train.save_experiments_dir={tmp_path}", "train.ws_threshold_model_save=1e9"])
    assert len(hist) == 2
    for h in hist:
        for k in ("gen_loss", "disc_loss", "router_loss", "div_loss", "intensity_loss", "aux_reg_loss", "ws_mean", "ws_std",
                  "gen_loss_0", "n_choosen_experts_mean_epoch_2", "ws_mean_1", "epoch_time"):
            assert k in h and math.isfinite(float(h[k])), (k, h.get(k))
    files = [f for _, _, fs in os.walk(tmp_path) for f in fs]
    for name in ("gen_0_epoch_1.pth", "disc_2_epoch_1.pth", "aux_reg_1_epoch_0.pth", "router_network_epoch_1.pth", "gen_optim_0_epoch_1.pth"):
        assert name in files, files
    # a saved generator loads back into a fresh module (state_dict with the reference's key names)
    from expertsim.models import build_model
    path = next(os.path.join(d, "gen_0_epoch_1.pth") for d, _, fs in os.walk(tmp_path) if "gen_0_epoch_1.pth" in fs)
    g = build_model(f"{arch}.generator", dict(noise_dim=10, cond_dim=9, di_strength=0.1, in_strength=1e-3), "cuda")
    g.load_state_dict(torch.load(path))
    out = g(torch.randn(4, 10, device="cuda"), torch.randn(4, 9, device="cuda"))
    assert torch.isfinite(out).all() and out.shape[0] == 4
