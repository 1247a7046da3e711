"""Step-level parity: MoEWrapper.train_step / .generate on the B200 (grouped sm_100a kernels through the C-ABI) against
the CPU oracle — and through it the golden vectors recorded from the unmodified reference — on identical weights,
inputs and injected noise.

Bars (BASELINE.json north_star): expert assignment, counts and the token->expert permutation BIT-EXACT; generated
images, losses and gradients within the bf16 tolerance written next to each check (the generator computes in bf16 with
fp32 accumulation, everything else in fp32)."""
import copy
import json
import os

import pytest
import torch

import oracle.expertsim_oracle as orc
from gpu_util import DEV, check, log

pytestmark = pytest.mark.gpu


def make_cfg(arch, E, router_over=None):
    from expertsim.config import Config
    c = copy.deepcopy(orc.DEFAULT_CFG)
    c["model"]["architecture"] = arch
    c["model"]["n_experts"] = E
    c["model"]["router"].update(router_over or {})
    c["dataset"] = {"input_image_shape": list(orc.IMAGE_SHAPE[arch])}
    return c, Config(c)


def build_moe(arch, E, cfg, st=None):
    from expertsim.train.loop import setup_moe_system
    moe = setup_moe_system(cfg, torch.device(DEV))
    if st is not None:
        for e in range(E):
            moe.generators[e].load_state_dict(st.gens[e])
            moe.discriminators[e].load_state_dict(st.discs[e])
            moe.aux_regs[e].load_state_dict(st.auxs[e])
        moe.router.load_state_dict(st.router)
    moe.train()
    return moe


def to_dev(d):
    return {k: v.to(DEV) for k, v in d.items()}


def run_case(arch, E, B, seed, router_over=None, steps=1, tol_img=3e-2, tol_loss=3e-2, tol_grad=6e-2, gold=None):
    ocfg, cfg = make_cfg(arch, E, router_over)
    st = orc.make_state(arch, E, seed, ocfg)
    moe = build_moe(arch, E, cfg, st)
    tag = f"{arch} E={E} B={B}"
    for step in range(steps):
        batch = orc.make_batch(arch, B, seed + 17 * step)
        noise = orc.make_noise(arch, B, E, seed + 31 * step)
        collect = {}
        want, aux = orc.train_step(st, batch, noise, epoch=0, collect=collect)
        b = to_dev(batch)
        # gradient arenas are inspected after the step: Adam does not clear them
        got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=to_dev(noise))
        torch.cuda.synchronize()
        last = moe._last
        assert last["idx"].cpu().tolist() == aux["idx"].tolist(), "expert assignment must be bit-exact"
        assert last["counts"].cpu().tolist() == aux["counts"].tolist()
        masks = [(aux["idx"] == e).nonzero(as_tuple=True)[0] for e in range(E)]
        assert last["perm"].cpu().tolist() == torch.cat(masks).tolist(), "token->expert permutation must be bit-exact"
        if gold is not None:
            assert last["idx"].cpu().tolist() == gold[step]["idx"] and last["counts"].cpu().tolist() == gold[step]["counts"]
        off = 0
        H, W = orc.IMAGE_SHAPE[arch]
        for e in range(E):
            n = masks[e].numel()
            if n >= 2:
                check(f"[{tag} s{step}] fake1 expert {e}", last["img1"][off:off + n].view(n, 1, H, W), aux["fake1"][e], tol_img)
                check(f"[{tag} s{step}] fake2 expert {e}", last["img2"][off:off + n].view(n, 1, H, W), aux["fake2"][e], tol_img)
            off += n
        for k, v in want.items():
            g = float(got[k])
            scale = max(abs(v), 1e-3 if "loss" in k else 1e-6)
            log(f"[{tag} s{step}] metric {k:36s} got {g:+.6e} want {v:+.6e}")
            assert abs(g - v) <= tol_loss * scale, (k, g, v)
            if gold is not None and k in gold[step]["metrics"]:
                assert abs(g - gold[step]["metrics"][k]) <= tol_loss * max(abs(gold[step]["metrics"][k]), scale), (k, "golden")
        for key, kind in (("g", "g_grads"), ("d", "d_grads"), ("a", "a_grads")):
            arena = moe.arena(key)
            for e in range(E):
                if f"{kind}_{e}" not in collect:
                    assert float(arena.G[e].abs().max()) == 0.0, "skipped expert must receive no gradient"
                    continue
                for name, gw in collect[f"{kind}_{e}"].items():
                    # discriminator gradients of the D step were consumed by Adam(D); the arena was re-zeroed only for
                    # G and aux, so D's arena still holds the D-step gradients
                    t = tol_grad if key != "d" else tol_grad
                    check(f"[{tag} s{step}] grad {key}{e} {name}", arena.view(arena.G, name, e), gw, t)
        if "r_grads" in collect:
            for name, gw in collect["r_grads"].items():
                check(f"[{tag} s{step}] grad router {name}", moe.arena("r").view(moe.arena("r").G, name, 0), gw, 2e-3)
        # weights after the fused Adam vs the oracle's torch-Adam restatement
        for e in range(E):
            for name in ("fc2.0.weight", "conv_layers.1.weight", "conv_layers.11.bias"):
                check(f"[{tag} s{step}] weight g{e} {name}", moe.generators[e].state_dict()[name], st.gens[e][name], 1e-3)
            for name in ("fc1.0.weight_orig", "fc1.0.weight_u", "fc1.0.weight_v", "conv_layers.0.weight_u"):
                check(f"[{tag} s{step}] weight d{e} {name}", moe.discriminators[e].state_dict()[name], st.discs[e][name], 2e-3)
    return moe, st


def golden(name):
    p = os.path.join(os.path.dirname(__file__), "golden", name)
    return json.load(open(p))["case"]


def test_train_step_proton_E3_B24_golden():
    c = golden("train_step_proton_E3_B24.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_train_step_proton_E1_B8_golden():
    c = golden("train_step_proton_E1_B8.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_train_step_proton_skip_rule_golden():
    """E=8, B=10: several experts receive 0 or 1 samples and must be skipped (reference moe.py:126-135)."""
    c = golden("train_step_proton_E8_B10_skip.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_train_step_proton_entropy_and_distribution_losses_golden():
    c = golden("train_step_proton_E2_B12_ent_ed.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_generate_matches_oracle():
    """Batch inference: routing bit-exact, showers (expm1, original order) within bf16 tolerance; float64 like the
    reference's numpy output."""
    arch, E, N, seed = "proton", 4, 40, 3
    ocfg, cfg = make_cfg(arch, E)
    st = orc.make_state(arch, E, seed, ocfg)
    moe = build_moe(arch, E, cfg, st).eval()
    g = torch.Generator().manual_seed(99)
    cond = torch.randn(N, 9, generator=g)
    gumbel = -torch.empty(N, E).exponential_(generator=g).log()
    z = torch.randn(N, 10, generator=g)
    want, idx, counts = orc.moe_generate(st, cond, gumbel, z)
    got, gidx = moe.generate(cond.to(DEV), noise=z.to(DEV), gumbel=gumbel.to(DEV), out_dtype=torch.float64, return_routing=True)
    assert got.dtype == torch.float64 and tuple(got.shape) == (N, 56, 30)
    assert gidx.cpu().tolist() == idx.tolist()
    check("moe.generate showers (expm1, original order)", got, want, 3e-2)
    # the reference's helper on a single expert generator (train/utils.py:179-205)
    from expertsim.train.utils import get_predictions_from_generator_results
    res, raw = get_predictions_from_generator_results(16, N, 10, DEV, cond.to(DEV), moe.generators[1], (56, 30), input_noise=z.to(DEV))
    w_res, w_raw = orc.generate(arch, st.gens[1], z, cond, batch_size=16)
    check("get_predictions_from_generator_results transformed", torch.from_numpy(res), w_res, 3e-2)
    check("get_predictions_from_generator_results raw", torch.from_numpy(raw), w_raw, 3e-2)


def test_modules_standalone_forward():
    """The drop-in modules are usable on their own (loop.py:265,275-281 call moe.router / moe.generators[i] directly)."""
    from expertsim.models import build_model
    arch, seed, B = "proton", 5, 6
    g = torch.Generator().manual_seed(1)
    cond, z = torch.randn(B, 9, generator=g), torch.randn(B, 10, generator=g)
    img = torch.rand(B, 1, 56, 30, generator=g)
    gen = build_model("proton.generator", dict(noise_dim=10, cond_dim=9, di_strength=.1, in_strength=1e-3), DEV)
    sd = orc.make_weights(arch, "generator", seed)
    gen.load_state_dict(sd)
    check("Generator.forward", gen(z.to(DEV), cond.to(DEV)), orc.generator_forward(arch, sd, z, cond), 3e-2)
    disc = build_model("proton.discriminator", dict(cond_dim=9), DEV).eval()
    sd = orc.make_weights(arch, "discriminator", seed)
    disc.load_state_dict(sd)
    s, l = disc(img.to(DEV), cond.to(DEV))
    ws, wl = orc.discriminator_forward(arch, sd, img, cond, training=False)
    check("Discriminator.forward score", s, ws, 1e-4)
    check("Discriminator.forward latent", l, wl, 1e-4)
    aux = build_model("proton.aux_reg", dict(strength=1e-3), DEV).eval()
    sd = orc.make_weights(arch, "aux_reg", seed)
    aux.load_state_dict(sd)
    check("AuxReg.forward", aux(img.to(DEV)), orc.aux_forward(arch, sd, img, training=False), 1e-4)
    router = build_model("router_v1", dict(cond_dim=9, n_experts=5), DEV)
    sd = orc.make_weights(arch, "router", seed, n_experts=5)
    router.load_state_dict(sd)
    gates, logits = router(cond.to(DEV))
    check("RouterNetwork.forward logits", logits, orc.router_forward(sd, cond, torch.zeros(B, 5))[1], 1e-5)
    assert torch.allclose(gates.sum(1).cpu(), torch.ones(B), atol=1e-5)
