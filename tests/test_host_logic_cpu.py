"""Host-side logic of the drop-in surface that needs no GPU: config loading, registry, state_dict names against the
oracle's (reference-measured) layouts, arena aliasing, deepcopy semantics, loaders, checkpoint round trip, metric helper."""
import copy
import os

import numpy as np
import pytest
import torch

import oracle.expertsim_oracle as orc
from expertsim.config import Config, load_config
from expertsim.models import MODEL_REGISTRY, build_model
from expertsim.models.moe import MoEWrapper


def small_moe(E=2, arch="proton"):
    cfg = load_config(None, [f"model.n_experts={E}", f"model.architecture={arch}"])
    from expertsim.train.loop import setup_moe_system
    return setup_moe_system(cfg, torch.device("cpu")), cfg


def test_config_schema_and_overrides():
    cfg = load_config(None, ["model.n_experts=8", "train.batch_size=64", "model.router.alb_strength=1e-4"])
    assert cfg.model.n_experts == 8 and cfg.train.batch_size == 64 and cfg.model.router.alb_strength == 1e-4
    assert isinstance(cfg.model.generator.lr_g, float) and cfg.model.generator.lr_g == 1e-4    # OmegaConf-style typing
    assert cfg.train.epoch_to_load is None and cfg.dataset.input_image_shape == [56, 30]
    assert set(dict(**cfg.model.generator)) == {"lr_g", "di_strength", "in_strength"}          # ** unpacking works
    cfg.generator_name = "x"                                                                    # struct is off
    assert cfg.generator_name == "x"
    ref_keys = {"version", "lr_r", "ed_strength", "gan_strength", "diff_strength", "util_strength", "alb_strength",
                "stop_router_training_epoch", "alpha", "min_weight", "tau_start", "tau_min", "tau_decay"}
    assert set(cfg.model.router) == ref_keys


def test_registry_keys():
    assert {"proton.generator", "proton.discriminator", "proton.aux_reg", "neutron.generator", "neutron.discriminator",
            "neutron.aux_reg", "router_v1"} <= set(MODEL_REGISTRY)
    with pytest.raises(KeyError):
        build_model("nope", {}, "cpu")


@pytest.mark.parametrize("arch", ["proton", "neutron"])
def test_state_dict_names_and_shapes_match_the_reference(arch):
    kw = {"generator": dict(noise_dim=10, cond_dim=9, di_strength=0.1, in_strength=1e-3), "discriminator": dict(cond_dim=9),
          "aux_reg": dict(strength=1e-3)}
    for kind in ("generator", "discriminator", "aux_reg"):
        m = build_model(f"{arch}.{kind}", kw[kind], "cpu")
        want = orc.make_weights(arch, kind, 0)
        got = m.state_dict()
        assert list(got) == list(want), kind
        for k in want:
            assert tuple(got[k].shape) == tuple(want[k].shape), (kind, k)
        m.load_state_dict(want)                      # reference-layout checkpoints load
        k0 = next(k for k in want if want[k].dtype.is_floating_point)
        assert torch.equal(m.state_dict()[k0], want[k0])
    r = build_model("router_v1", dict(cond_dim=9, n_experts=5), "cpu")
    assert list(r.state_dict()) == list(orc.make_weights(arch, "router", 0, n_experts=5))


def test_wrapper_arenas_alias_the_module_parameters():
    moe, cfg = small_moe(E=3)
    sd = moe.state_dict()
    assert "generators.2.fc2.0.weight" in sd and "discriminators.0.fc3.weight_u" in sd and "router.fc_layers.6.bias" in sd
    a = moe.arena("g")
    assert a.P.shape[0] == 3
    w = moe.generators[1].fc2[0].weight
    assert w.data_ptr() == a.view(a.P, "fc2.0.weight", 1).data_ptr()
    a.P[1].zero_()
    assert float(w.abs().max()) == 0.0 and float(moe.generators[0].fc2[0].weight.abs().max()) > 0
    # identical initial experts (reference deepcopy) but no shared storage
    assert torch.equal(moe.generators[0].fc1[0].weight, moe.generators[2].fc1[0].weight)
    # spectral-norm vectors are buffers, not parameters
    names = [n for n, _ in moe.discriminators[0].named_parameters()]
    assert "fc1.0.weight_orig" in names and "fc1.0.weight_u" not in names
    n_g = sum(p.numel() for p in moe.generators[0].parameters())
    assert n_g == 26_571_841


def test_deepcopy_does_not_alias_and_to_rebinds():
    moe, _ = small_moe(E=2)
    g2 = copy.deepcopy(moe.generators[0])
    g2.fc1[0].weight.data.zero_()
    assert float(moe.generators[0].fc1[0].weight.abs().max()) > 0
    moe.double()                       # any .to()/.float()/.cuda() re-creates storages: arenas must be re-adopted
    assert all(a.owns(a.modules[0]) for a in moe._arenas.values())


def test_product_path_refuses_to_compute_without_cuda():
    moe, _ = small_moe(E=2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        moe.train_step(0, torch.randn(4, 9), torch.zeros(4, 1, 56, 30), torch.zeros(4, 2), torch.rand(4, 1), torch.rand(4, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        moe.generators[0](torch.randn(2, 10), torch.randn(2, 9))


def test_optimizer_handles_and_checkpoint_roundtrip(tmp_path):
    from expertsim.train.training_setup import load_checkpoint_weights, setup_optimizers
    from expertsim.train.training_utils import save_models_and_architectures
    moe, cfg = small_moe(E=2)
    g, d, a, r = setup_optimizers(moe, cfg)
    assert len(g) == len(d) == len(a) == 2 and g[0].param_groups[0]["lr"] == cfg.model.generator.lr_g
    assert d[0].param_groups[0]["lr"] == cfg.model.discriminator.lr_d and r.param_groups[0]["lr"] == cfg.model.router.lr_r
    moe.arena("g").M[1].fill_(0.5)
    save_models_and_architectures(str(tmp_path), 2, moe.aux_regs, a, moe.generators, g, moe.discriminators, d, moe.router, r, 7)
    assert os.path.exists(tmp_path / "gen_1_epoch_7.pth") and os.path.exists(tmp_path / "router_network_epoch_7.pth")
    ref = {k: v.clone() for k, v in moe.state_dict().items()}
    with torch.no_grad():
        for p in moe.parameters():
            p.add_(1.0)
    moe.arena("g").M.zero_()
    load_checkpoint_weights(str(tmp_path), 7, moe, g, d, a, r)
    for k, v in moe.state_dict().items():
        assert torch.equal(v, ref[k]), k
    assert float(moe.arena("g").M[1].min()) == 0.5


def test_device_loader_shards_are_disjoint_and_cover_the_global_batch():
    from expertsim.utils.data import DeviceLoader, synthetic_showers
    data = synthetic_showers("proton", 64, seed=1)
    full = [b for b in DeviceLoader(data, 16, shuffle=True, rank=0, world=1, seed=3)]
    l0 = [b for b in DeviceLoader(data, 8, shuffle=True, rank=0, world=2, seed=3)]
    l1 = [b for b in DeviceLoader(data, 8, shuffle=True, rank=1, world=2, seed=3)]
    assert len(full) == len(l0) == len(l1) == 4
    for f, a, b in zip(full, l0, l1):
        merged = torch.empty_like(f[2])
        merged[0::2], merged[1::2] = a[2], b[2]
        assert torch.equal(merged, f[2])
    x, x2, cond, std, inten, pos = full[0]
    assert x.shape == (16, 56, 30) and cond.shape == (16, 9) and std.shape == (16, 1) and inten.shape == (16, 1) and pos.shape == (16, 2)
    assert torch.allclose(torch.expm1(x).sum((1, 2)), inten[:, 0], rtol=1e-4)


def test_channel_sums_match_the_oracle():
    from expertsim.train.utils import sum_channels_parallel
    x = np.random.RandomState(0).rand(5, 44, 44)
    got = np.array(list(sum_channels_parallel(x)))
    want = orc.sum_channels(torch.from_numpy(x)).numpy()
    assert np.abs(got - want).max() < 1e-10


def test_router_loss_helpers_match_the_oracle():
    from expertsim.train import utils as U
    g = torch.Generator().manual_seed(0)
    gates = torch.rand(12, 4, generator=g).softmax(1)
    m = torch.rand(12, 1, generator=g) * 10
    assert torch.allclose(U.calculate_adaptive_load_balancing_loss(gates.sum(0), 1e-2), orc.adaptive_load_balancing_loss(gates.sum(0), 1e-2))
    assert torch.allclose(U.calculate_expert_utilization_entropy(gates, 0.3), orc.utilization_entropy(gates, 0.3))
    assert torch.allclose(U.calculate_expert_distribution_loss(gates, m), orc.expert_distribution_loss(gates, m))
