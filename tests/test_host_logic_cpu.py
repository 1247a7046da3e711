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
    shape = "[56,30]" if arch == "proton" else "[44,44]"
    cfg = load_config(None, [f"model.n_experts={E}", f"model.architecture={arch}", f"dataset.input_image_shape={shape}"])
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
    moe.arena("g").steps[1] = 4         # torch.optim.Adam keeps no state for parameters that never stepped
    save_models_and_architectures(str(tmp_path), 2, moe.aux_regs, a, moe.generators, g, moe.discriminators, d, moe.router, r, 7)
    assert os.path.exists(tmp_path / "gen_1_epoch_7.pth") and os.path.exists(tmp_path / "router_network_epoch_7.pth")
    ref = {k: v.clone() for k, v in moe.state_dict().items()}
    with torch.no_grad():
        for p in moe.parameters():
            p.add_(1.0)
    moe.arena("g").M.zero_()
    moe.arena("g").steps.zero_()
    load_checkpoint_weights(str(tmp_path), 7, moe, g, d, a, r)
    for k, v in moe.state_dict().items():
        assert torch.equal(v, ref[k]), k
    ag = moe.arena("g")
    assert all(float(ag.view(ag.M, n, 1).min()) == 0.5 for n in ag.off) and ag.steps.tolist() == [0, 4]


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


# ---------------------------------------------------------------------------------------------------- round 2 host rows
def test_arena_adam_state_dict_interchanges_with_torch_adam():
    """optimizer checkpoints in torch.optim.Adam's format, both directions (the reference saves
    ``optimizer.state_dict()`` / pickled optimizers: train/training_utils.py:300-380, training_setup.py:46-49)."""
    from expertsim.train.training_setup import setup_optimizers
    moe, cfg = small_moe(E=2, arch="neutron")
    g_o, d_o, a_o, r_o = setup_optimizers(moe, cfg)
    assert g_o[0].state_dict()["state"] == {}                                   # torch: no state before the first step
    a = moe.arena("d")
    g = torch.Generator().manual_seed(0)
    a.M.copy_(torch.randn(a.M.shape, generator=g))
    a.V.copy_(torch.rand(a.V.shape, generator=g))
    a.steps.copy_(torch.tensor([3, 5], dtype=torch.int32))
    sd = d_o[1].state_dict()
    # -> a stock torch optimizer over a module with the same parameter order
    ref_opt = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in moe.discriminators[1].parameters()], lr=1e-5)
    ref_opt.load_state_dict(copy.deepcopy(sd))
    names = [n for n, _ in moe.discriminators[1].named_parameters()]
    for i, (n, p) in enumerate(zip(names, ref_opt.param_groups[0]["params"])):
        st = ref_opt.state[p]
        assert float(st["step"]) == 5.0
        assert torch.equal(st["exp_avg"], a.view(a.M, n, 1)) and torch.equal(st["exp_avg_sq"], a.view(a.V, n, 1))
    for p in ref_opt.param_groups[0]["params"]:                                 # the loaded state is usable: one torch step
        p.grad = torch.ones_like(p)
    ref_opt.step()
    # <- back into slot 0 of the arena (a reference checkpoint loaded into this build)
    d_o[0].load_state_dict(ref_opt.state_dict())
    assert int(a.steps[0]) == 6
    for n, p in zip(names, ref_opt.param_groups[0]["params"]):
        assert torch.equal(a.view(a.M, n, 0), ref_opt.state[p]["exp_avg"])
    assert "params" in d_o[0].param_groups[0] and isinstance(d_o[0].param_groups[0]["params"][0], torch.nn.Parameter)
    # the round-1 flat layout still loads
    d_o[0].load_state_dict({"step": 2, "exp_avg": torch.zeros(a.n), "exp_avg_sq": torch.ones(a.n), "param_groups": [{"lr": 1e-5}]})
    assert int(a.steps[0]) == 2 and float(a.M[0].abs().max()) == 0.0
    with pytest.raises(ValueError):
        bad = ref_opt.state_dict()
        bad["param_groups"][0]["params"] = bad["param_groups"][0]["params"][:-1]
        d_o[0].load_state_dict(bad)


def test_optimizer_handles_must_agree():
    from expertsim.train.training_setup import setup_optimizers
    moe, cfg = small_moe(E=2)
    g_o, _, _, _ = setup_optimizers(moe, cfg)
    assert MoEWrapper._lr(g_o, 1.0) == cfg.model.generator.lr_g
    g_o[1].param_groups[0]["lr"] = 5e-4
    with pytest.raises(NotImplementedError):
        MoEWrapper._lr(g_o, 1.0)
    g_o[1].param_groups[0]["lr"] = cfg.model.generator.lr_g
    g_o[0].param_groups[0]["betas"] = (0.5, 0.999)
    with pytest.raises(NotImplementedError):
        MoEWrapper._lr(g_o, 1.0)


def test_optimizer_state_moves_with_the_parameters():
    """re-binding the arenas (what .to()/.cuda() triggers) keeps Adam's moments and step counters"""
    moe, _ = small_moe(E=2)
    a = moe.arena("a")
    a.M.fill_(0.25)
    a.V.fill_(0.5)
    a.steps.fill_(7)
    moe._bind()
    b = moe.arena("a")
    assert b is not a and float(b.M.min()) == 0.25 and float(b.V.max()) == 0.5 and b.steps.tolist() == [7, 7]


def test_ema_helper_advances_only_live_experts():
    from expertsim.train.loop import EMAHelper
    moe, _ = small_moe(E=2)
    ema = EMAHelper(moe, decay=0.9, enabled=True)
    P = moe.arena("g").P
    before = P.clone()
    P.add_(1.0)
    ema.update(moe, live=torch.tensor([1, 0], dtype=torch.int32))
    assert torch.allclose(ema.shadow[0], 0.9 * before[0] + 0.1 * P[0], atol=1e-6)
    assert torch.equal(ema.shadow[1], before[1])
    ema.update(moe, updated_indices=[1])
    assert torch.allclose(ema.shadow[1], 0.9 * before[1] + 0.1 * P[1], atol=1e-6)
    v0 = moe.arena("g").version
    ema.apply_shadow(moe)
    assert torch.equal(moe.generators[0].fc1[0].weight, ema.shadow[0][:moe.generators[0].fc1[0].weight.numel()].view(256, 19))
    ema.restore(moe)
    assert torch.equal(moe.arena("g").P, P) and moe.arena("g").version == v0 + 2


def test_loaders_from_the_reference_arrays():
    """the arrays ``transform_data_for_training`` returns -> 6-tuple batches
    (data_transformations.py:269-271), resident or streamed through the pinned double buffer."""
    from expertsim.utils.data import DeviceLoader, PinnedHostLoader, loaders_from_arrays
    rng = np.random.default_rng(0)
    n = 50
    x = rng.random((n, 56, 30)).astype(np.float64)          # the reference's arrays are float64 / float32 mixed
    cond = rng.standard_normal((n, 9)).astype(np.float32)
    std, inten = rng.random(n), rng.random((n, 1)) * 100
    pos = rng.integers(0, 30, (n, 2)).astype(np.float64)
    dl = DeviceLoader.from_arrays(x, cond, std, inten, pos, batch_size=8, device="cpu", shuffle=False, drop_last=False)
    batches = list(dl)
    assert len(batches) == 7 and len(batches[0]) == 6
    xb, x2, c, s, it, p = batches[0]
    assert xb.dtype == torch.float32 and tuple(xb.shape) == (8, 56, 30) and tuple(s.shape) == (8, 1) and tuple(it.shape) == (8, 1)
    assert torch.equal(xb, x2) and np.allclose(c.numpy(), cond[:8]) and np.allclose(p.numpy(), pos[:8])
    assert tuple(batches[-1][0].shape) == (2, 56, 30)
    # rank sharding: ranks 0/1 of 2 see alternating rows of every global batch
    r0 = next(iter(DeviceLoader.from_arrays(x, cond, std, inten, pos, 4, "cpu", shuffle=False, rank=0, world=2)))
    r1 = next(iter(DeviceLoader.from_arrays(x, cond, std, inten, pos, 4, "cpu", shuffle=False, rank=1, world=2)))
    assert np.allclose(r0[2].numpy(), cond[0:8:2]) and np.allclose(r1[2].numpy(), cond[1:8:2])
    with pytest.raises(ValueError):
        DeviceLoader.from_arrays(x, cond[:-1], std, inten, pos, 8, "cpu")
    # host-streamed variant: same batches
    hl = PinnedHostLoader(x, cond, std, inten, pos, 8, device="cpu")
    hb = list(hl)
    assert len(hb) == 6 and all(torch.equal(a[2], b[2]) and torch.equal(a[0], b[0]) for a, b in zip(hb, batches))
    cfg = load_config(None, ["train.batch_size=8"])
    tr, te = loaders_from_arrays(cfg, x[10:], x[:10], cond[10:], cond[:10], std[10:], std[:10], inten[10:], inten[:10],
                                 pos[10:], pos[:10], device="cpu")
    assert len(tr) == 5 and len(list(te)) == 2


def test_zero_pool_alternates_and_clears(monkeypatch):
    """the per-step scratch pool (one memset instead of ~90 fills): 256-byte aligned views, two alternating buffers,
    cleared at begin(), escaping tensors never pooled"""
    from expertsim import _nets as N
    monkeypatch.setattr(N, "_dev", lambda: torch.device("cpu"))
    zp = N.ZeroPool()
    monkeypatch.setattr(N, "ZP", zp)
    zp.begin(torch.device("cpu"))
    a, b = N.zeros(3, 5), N.zeros(7, dtype=torch.float64)          # first step: nothing pooled yet, plain zeros
    zp.end()
    assert zp.need == 512 and zp.bufs == [None, None]
    zp.begin(torch.device("cpu"))
    a, b, c = N.zeros(3, 5), N.zeros(7, dtype=torch.float64), N.zeros(2, 2, escape=True)
    a += 1.0
    assert zp.used == 512 and float(b.sum()) == 0.0 and b.dtype == torch.float64 and tuple(a.shape) == (3, 5)
    assert c.untyped_storage().data_ptr() != a.untyped_storage().data_ptr()
    zp.end()
    zp.begin(torch.device("cpu"))
    a2 = N.zeros(3, 5)
    assert a2.data_ptr() != a.data_ptr() and float(a.sum()) == 15.0          # the previous step's views stay untouched
    zp.end()
    zp.begin(torch.device("cpu"))
    a3 = N.zeros(3, 5)
    assert a3.data_ptr() == a.data_ptr() and float(a3.sum()) == 0.0          # two steps later: same memory, cleared
    big = N.zeros(1000)                                                       # does not fit: falls back, pool grows next step
    zp.end()
    assert zp.need >= 256 + 4096 and float(big.sum()) == 0.0
    assert not zp.active and N.zeros(2).sum() == 0


def test_pipelined_adam_decomposition_equals_one_launch(monkeypatch):
    """MoEWrapper._adam_pipelined (data-parallel steps) runs the fused Adam rectangle by rectangle behind the chunked
    all-reduce.  Host logic pinned here with es_adam_step emulated over HOST pointers (same argument meaning as the C-ABI
    entry: p/g/m/v + n + slot_stride + slots, per-slot counters advanced for live groups only): every parameter is updated
    exactly once with the bias correction of step t + 1, the counters advance once, skipped experts stay untouched —
    bit-identical to one launch over the arena."""
    import ctypes
    from types import SimpleNamespace
    import numpy as np
    from expertsim.models import moe as moe_mod

    def view(addr, count, dt):
        return np.ctypeslib.as_array((dt * count).from_address(addr))

    calls = []

    def fake_call(name, p, g, m, v, n, stride, slots, lr, b1, b2, eps, steps, grp):
        assert name == "es_adam_step"
        calls.append((n, slots))
        st = view(steps if isinstance(steps, int) else steps.data_ptr(), slots, ctypes.c_int32)
        gr = None if grp is None else view(grp if isinstance(grp, int) else grp.data_ptr(), slots * 4, ctypes.c_int32).reshape(slots, 4)
        ptr = lambda a: a if isinstance(a, int) else a.data_ptr()
        for s_ in range(slots):
            if gr is not None and gr[s_, 1] == 0:
                continue
            st[s_] += 1
            t = int(st[s_])
            P, G_, M, V = (view(ptr(a) + 4 * s_ * stride, n, ctypes.c_float) for a in (p, g, m, v))
            M[:] = M + (G_ - M) * np.float32(1 - b1)
            V[:] = V * np.float32(b2) + G_ * G_ * np.float32(1 - b2)
            step = np.float32(lr / (1 - b1 ** t))
            P[:] = P - step * M / (np.sqrt(V) / np.float32(np.sqrt(1 - b2 ** t)) + np.float32(eps))

    monkeypatch.setattr(moe_mod.L, "call", fake_call)
    E, n = 5, 4096
    g = torch.Generator().manual_seed(1)

    def arena():
        g.manual_seed(1)
        return SimpleNamespace(P=torch.randn(E, n, generator=g), G=torch.randn(E, n, generator=g), M=torch.randn(E, n, generator=g) * .1,
                               V=torch.rand(E, n, generator=g) * .1, steps=torch.tensor([3, 0, 7, 7, 1], dtype=torch.int32), n=n, E=E,
                               version=0)

    grp = torch.tensor([[0, 4, 0, 4], [4, 0, 1, 0], [4, 2, 2, 2], [6, 0, 3, 0], [6, 9, 4, 9]], dtype=torch.int32)
    one = arena()
    moe_mod.MoEWrapper._adam(one, 1e-3, grp)
    joins = []
    red = SimpleNamespace(join=lambda: joins.append(len(calls)))
    for rects in ([(0, 2, 1024, 3072, None), (2, 4, 1024, 3072, None), (4, 5, 1024, 3072, None)],        # row blocks, both remainders
                  [(e, e + 1, c, c + 2048, None) for e in range(E) for c in (0, 2048)],                  # column blocks, whole arena
                  [(0, 5, 0, 3072, None)]):                                                              # no left remainder
        calls.clear(), joins.clear()
        pip = arena()
        moe_mod.MoEWrapper._adam_pipelined(SimpleNamespace(), pip, 1e-3, grp, rects, red)
        for k in ("P", "M", "V", "steps"):
            assert torch.equal(getattr(pip, k), getattr(one, k)), k
        assert pip.version == 1 and len(joins) == 1
        hi = max(r[3] for r in rects)
        lo = min(r[2] for r in rects)
        assert len(calls) == len(rects) + (hi < n) + (lo > 0)
        assert joins[0] == len(calls) - (lo > 0)          # only the columns left of the bucket wait for join()
    assert one.steps.tolist() == [4, 0, 8, 7, 2]


def test_committed_bench_lines_carry_the_contract_keys():
    """The measured lines under profiles/ (written by bench.py on the B200) carry every key the bench contract names —
    both arms — so a change of bench.py that drops one shows up here before it reaches the GPU box."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in line, k
    assert line["config"]["workload"] and "model" not in line["config"]
    assert line["gpu_launches"] > 0 and line["warmup"] >= 3 and line["higher_is_better"] is True
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"]) and line["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "executed_frac", "tensor_pipe_active_pct"} <= set(line["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"]) and line["cpu_baseline"]["kind"] == "reference"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    ref = json.loads(open(os.path.join(root, "profiles", "r02_bench_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["config"] == line["config"] and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "reference"
    two = json.loads(open(os.path.join(root, "profiles", "r02_bench_n2_adam_pipelined.json")).read().strip().splitlines()[-1])
    assert two["n_gpus"] == 2 and two["dp_parity"]["replicas_identical"] is True and two["comm"]["adam_pipelined_rects"] == 4
