"""Step-level parity: MoEWrapper.train_step / .generate on the B200 (grouped sm_100a kernels through the C-ABI) against
the CPU oracle — and through it the golden vectors recorded from the unmodified reference — on identical weights,
inputs and injected noise.

Bars (BASELINE.json north_star): expert assignment, counts and the token->expert permutation BIT-EXACT; generated
images, losses and gradients within the bf16 tolerance written next to each check (the generator computes in bf16 with
fp32 accumulation, everything else in fp32)."""
import copy
import json
import os
import re

import pytest
import torch
import numpy as np
import torch.nn.functional as F

import oracle.expertsim_oracle as orc
from gpu_util import DEV, check as _check, log

pytestmark = pytest.mark.gpu
ABS_FLOOR = 8e-3


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


def sync_from_oracle(moe, st):
    """weights, buffers and Adam state of the oracle -> the arenas (teacher forcing between steps)."""
    for key, mods, sds, opts in (("g", moe.generators, st.gens, st.opt_g), ("d", moe.discriminators, st.discs, st.opt_d),
                                 ("a", moe.aux_regs, st.auxs, st.opt_a), ("r", [moe.router], [st.router], [st.opt_r])):
        arena = moe.arena(key)
        for e, (m, sd, opt) in enumerate(zip(mods, sds, opts)):
            m.load_state_dict(sd)
            arena.steps[e] = opt.step
            for name in opt.m:
                arena.view(arena.M, name, e).copy_(opt.m[name])
                arena.view(arena.V, name, e).copy_(opt.v[name])
    moe.mark_weights_changed()


def to_dev(d):
    return {k: v.to(DEV) for k, v in d.items()}


def run_case(arch, E, B, seed, router_over=None, steps=1, tol_img=3e-2, tol_loss=3e-2, tol_grad=None, gold=None,
             inject_images=False, tol_by_kind=None):
    """inject_images=False: the full CUDA step end to end.  Gradients are then compared at a LOOSE bound (rel. L2 0.45 and
    cosine >= 0.9): the reference's networks are non-smooth in the image (max-pool routing, ReLU kinks, GroupNorm over
    sparse maps) — in pure fp32 PyTorch a 1e-2 relative image perturbation, which is what bf16 costs, already moves the
    early-layer gradients by 10-20% (test_gradient_conditioning_of_the_reference_network below measures it).
    inject_images=True: the oracle's fp32 images replace the generated ones after the generator forward, which removes that
    amplification: discriminator / aux-regressor / loss-tail gradients must then match at fp32 tolerance (2e-3) and the
    generator's gradients at the bf16 tolerance 0.2 rel. L2 / cosine >= 0.985 (observed 0.06-0.15).  (The generator backward runs on bf16
    activations whose LeakyReLU(0.1) pre-activations carry the forward's ~1e-2 error; the ~0.5% of elements that change
    sign get a 10x different slope, which alone is several % of the gradient norm per layer.  The per-kernel tests, which
    feed both sides identical bf16 inputs, hold every backward kernel to 1e-2.)
    Steps after the first start from the ORACLE's state (weights, spectral-norm vectors, Adam moments): Adam's first
    updates are sign-like (|update| = lr whatever |g| is), so any gradient noise turns into full-size weight differences
    and a free-running comparison would only measure chaos."""
    fails = []
    tol = dict(g=0.2, d=2e-3, a=2e-3) if inject_images else dict(g=0.6, d=0.6, a=0.6)
    if tol_grad is not None:
        tol = dict(g=tol_grad, d=tol_grad, a=tol_grad)
    if tol_by_kind is not None:
        tol = dict(tol_by_kind)

    def check(*a, **k):     # report every violated bound of the case, not only the first
        try:
            _check(*a, **k)
        except AssertionError as e:
            fails.append(str(e))

    ocfg, cfg = make_cfg(arch, E, router_over)
    st = orc.make_state(arch, E, seed, ocfg)
    moe = build_moe(arch, E, cfg, st)
    tag = f"{arch} E={E} B={B}"
    lrs = dict(g=ocfg["model"]["generator"]["lr_g"], d=ocfg["model"]["discriminator"]["lr_d"], a=ocfg["model"]["aux_reg"]["lr_a"])
    for step in range(steps):
        if step > 0:
            sync_from_oracle(moe, st)
        batch = orc.make_batch(arch, B, seed + 17 * step)
        noise = orc.make_noise(arch, B, E, seed + 31 * step)
        collect = {}
        want, aux = orc.train_step(st, batch, noise, epoch=0, collect=collect)
        b = to_dev(batch)
        # gradient arenas are inspected after the step: Adam does not clear them
        nz = to_dev(noise)
        masks = [(aux["idx"] == e).nonzero(as_tuple=True)[0] for e in range(E)]
        if inject_images:
            H, W = orc.IMAGE_SHAPE[arch]
            for key in ("fake1", "fake2"):
                rows = [aux[key][e].reshape(-1, H * W) if e in aux[key] else torch.zeros(masks[e].numel(), H * W) for e in range(E)]
                nz["img1_sorted" if key == "fake1" else "img2_sorted"] = torch.cat(rows).to(DEV)
        got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=nz)
        torch.cuda.synchronize()
        last = moe._last
        assert last["idx"].cpu().tolist() == aux["idx"].tolist(), "expert assignment must be bit-exact"
        assert last["counts"].cpu().tolist() == aux["counts"].tolist()
        assert last["perm"].cpu().tolist() == torch.cat(masks).tolist(), "token->expert permutation must be bit-exact"
        if gold is not None:
            assert last["idx"].cpu().tolist() == gold[step]["idx"] and last["counts"].cpu().tolist() == gold[step]["counts"]
        off = 0
        H, W = orc.IMAGE_SHAPE[arch]
        for e in range(E):
            n = masks[e].numel()
            if n >= 2:      # with injected images the check reads what the generator produced BEFORE the injection
                i1, i2 = (last["img1_generated"], last["img2_generated"]) if inject_images else (last["img1"], last["img2"])
                check(f"[{tag} s{step}] fake1 expert {e}", i1[off:off + n].view(n, 1, H, W), aux["fake1"][e], tol_img)
                check(f"[{tag} s{step}] fake2 expert {e}", i2[off:off + n].view(n, 1, H, W), aux["fake2"][e], tol_img)
            off += n
        # Loss tolerance: 3e-2 relative plus an absolute floor of 8e-3.  The floor is the bf16 image rounding (rel. L2 1e-2,
        # checked above) seen through the discriminator: hinge scores are O(1) and move by several 1e-3 per sample, and
        # gen_loss = -mean(score) over as few as 2..8 samples is a small difference of such numbers.
        for k, v in want.items():
            g = float(got[k])
            log(f"[{tag} s{step}] metric {k:36s} got {g:+.6e} want {v:+.6e}")
            floor = ABS_FLOOR
            if k.startswith("std_intensities_experts_"):
                # std of B_e photon sums that may be nearly equal: the floor is the bf16 error of one 1680-pixel sum
                floor = 2e-3 * abs(want[k.replace("std_", "mean_")]) + ABS_FLOOR
            if not abs(g - v) <= tol_loss * abs(v) + floor:
                fails.append(f"metric {k}: got {g} want {v}")
            if gold is not None and k in gold[step]["metrics"]:
                gv = gold[step]["metrics"][k]
                if not abs(g - gv) <= tol_loss * abs(gv) + floor:
                    fails.append(f"metric {k} vs golden: got {g} want {gv}")
        for key, kind in (("g", "g_grads"), ("d", "d_grads"), ("a", "a_grads")):
            arena = moe.arena(key)
            for e in range(E):
                if f"{kind}_{e}" not in collect:
                    assert float(arena.G[e].abs().max()) == 0.0, "skipped expert must receive no gradient"
                    continue
                for name, gw in collect[f"{kind}_{e}"].items():
                    # the D arena still holds the D-step gradients (the G step computes no D weight gradients)
                    got_g = arena.view(arena.G, name, e)
                    if arch == "neutron" and ((key == "g" and name in ("fc1.0.bias", "fc2.0.bias", "conv_layers.0.bias",
                                                                        "conv_layers.5.bias", "conv_layers.9.bias"))
                                              or (key == "a" and re.fullmatch(r"feature_extractor\.conv\d\.bias", name))):
                        # a bias in front of a BatchNorm: its gradient is identically zero; autograd returns rounding noise.
                        # Both sides must be negligible next to the layer's weight gradient.
                        scale = float(collect[f"{kind}_{e}"][name.replace(".bias", ".weight")].abs().max())
                        if float(got_g.abs().max()) > 1e-2 * scale or float(gw.abs().max()) > 1e-2 * scale:
                            fails.append(f"grad {key}{e} {name}: expected ~0 (BatchNorm follows), got {float(got_g.abs().max()):.2e}, "
                                         f"oracle {float(gw.abs().max()):.2e}, weight-grad scale {scale:.2e}")
                        continue
                    if float(gw.abs().max()) < 1e-9:      # analytically-zero gradients (bias in front of a norm layer)
                        if float(got_g.abs().max()) > 1e-7:
                            fails.append(f"grad {key}{e} {name}: expected ~0, got {float(got_g.abs().max()):.2e}")
                        continue
                    if gw.numel() >= 64 or inject_images:   # e2e: a scalar "relative L2" (e.g. the 1-channel output bias) is
                        check(f"[{tag} s{step}] grad {key}{e} {name}", got_g, gw, tol[key])   # pure cancellation noise
                    if gw.numel() >= 64:
                        cos = float(F.cosine_similarity(got_g.flatten().double().cpu(), gw.flatten().double(), dim=0))
                        if cos < ((0.999 if key != "g" else 0.985) if inject_images else 0.8):
                            fails.append(f"grad {key}{e} {name}: cosine {cos:.4f}")
        if "r_grads" in collect:
            for name, gw in collect["r_grads"].items():
                # the expert-distribution term sees the photon sums of the (bf16) generated images
                check(f"[{tag} s{step}] grad router {name}", moe.arena("r").view(moe.arena("r").G, name, 0), gw,
                      2e-3 if inject_images or not (router_over or {}).get("ed_strength") else 3e-2)
        # weights after the fused Adam vs the oracle's torch-Adam restatement.  An Adam update is bounded by ~lr per element
        # and sign-like in the first steps, so the elementwise bound is 2.1*lr and the mean difference must stay well below lr
        # (the Adam kernel itself is checked to 1e-6 on identical gradients in test_kernels_gpu.py).
        for e in range(E):
            if masks[e].numel() < 2:
                continue
            gn = ("fc2.0.weight", "conv_layers.1.weight", "conv_layers.11.weight") if arch == "proton" else \
                ("fc2.0.weight", "conv_layers.0.weight", "conv_layers.13.weight")
            an = ("regressor.0.weight",) if arch == "proton" else ("feature_extractor.conv2.weight", "dense.weight")
            for key, mods, sds, names in (("g", moe.generators, st.gens, gn),
                                          ("d", moe.discriminators, st.discs, ("fc1.0.weight_orig", "conv_layers.4.weight_orig")),
                                          ("a", moe.aux_regs, st.auxs, an)):
                for name in names:
                    d = (mods[e].state_dict()[name].cpu() - sds[e][name]).abs()
                    mx, mean = float(d.max()) / lrs[key], float(d.mean()) / lrs[key]
                    log(f"[{tag} s{step}] weight {key}{e} {name}: max|dw|/lr={mx:.3f} mean|dw|/lr={mean:.4f}")
                    lim = 0.3 if (inject_images or key != "g") else 0.6
                    if mx > 2.1 or mean > lim:
                        fails.append(f"weight {key}{e} {name}: max|dw|/lr={mx:.3f} mean|dw|/lr={mean:.4f}")
            for name in ("fc1.0.weight_u", "fc1.0.weight_v", "conv_layers.0.weight_u"):
                check(f"[{tag} s{step}] buffer d{e} {name}", moe.discriminators[e].state_dict()[name], st.discs[e][name], 2e-3)
    assert not fails, "\n".join(fails)
    return moe, st


def golden(name):
    p = os.path.join(os.path.dirname(__file__), "golden", name)
    return json.load(open(p))["case"]


def test_train_step_proton_E3_B24_golden():
    c = golden("train_step_proton_E3_B24.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_train_step_proton_E3_B24_injected_images():
    """Same case with the oracle's images injected after the generator forward: tight gradient bounds."""
    c = golden("train_step_proton_E3_B24.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=2, inject_images=True)


def test_gradient_conditioning_of_the_reference_network():
    """Documents WHY end-to-end gradients carry a loose bound: in plain fp32 PyTorch (the oracle, no CUDA code involved) a
    1e-2 relative perturbation of the generated images moves dL/d(image) of the discriminator by more than 5%."""
    sd, gs = orc.make_weights("proton", "discriminator", 8), orc.make_weights("proton", "generator", 7)
    g = torch.Generator().manual_seed(1)
    z, c = torch.randn(8, 10, generator=g), torch.randn(8, 9, generator=g)
    with torch.no_grad():
        fake = orc.generator_forward("proton", gs, z, c)

    def d_img(img):
        img = img.clone().requires_grad_(True)
        out, _ = orc.discriminator_forward("proton", {k: v.clone() for k, v in sd.items()}, img, c, training=False)
        (-out.mean()).backward()
        return img.grad

    pert = fake * (1 + 1e-2 * torch.randn(fake.shape, generator=g))
    a, b = d_img(fake), d_img(pert)
    r = float((a - b).norm() / a.norm())
    log(f"fp32 PyTorch: 1e-2 image perturbation -> dL/dimg moves by relL2={r:.3f}")
    assert r > 0.05


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


def test_train_step_proton_entropy_and_distribution_losses_injected_images():
    c = golden("train_step_proton_E2_B12_ent_ed.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), inject_images=True)


def test_train_step_proton_skip_rule_injected_images():
    c = golden("train_step_proton_E8_B10_skip.json")
    run_case("proton", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), inject_images=True)


@pytest.mark.parametrize("arch", ["proton", "neutron"])
def test_train_step_E8_B1024_benched_configuration(arch):
    """The configuration bench.py measures (BASELINE configs[2] / the configs[3] per-GPU slice): 8 experts, 1024 rows — ragged
    groups of ~100-170 rows, hundreds of GEMM tiles per launch, several tiles per persistent CTA.  Routing / counts /
    permutation bit-exact; generated images 3e-2; losses 3e-2 rel + 8e-3 abs; with the oracle's images injected after the
    generator forward: D / aux / loss-tail gradients 2e-3, generator gradients 0.2 rel. L2 and cosine >= 0.985."""
    # neutron: the aux regressor's BatchNorm backward subtracts batch means over ~130 rows x 1764 pixels per channel; this
    # build accumulates those in fp64, the CPU oracle (torch fp32) does not, so at this size the ORACLE's own rounding shows
    # (observed 2.0-3.0e-3 on 15 of 200 tensors; 2e-3 holds at B = 24): the aux / D bound is 5e-3 here.
    tol = None if arch == "proton" else dict(g=0.2, d=5e-3, a=5e-3)
    run_case(arch, 8, 1024, seed=23, inject_images=True, tol_by_kind=tol)


def test_train_step_gradients_run_to_run():
    """Split-K / multi-CTA reductions accumulate with fp32 atomics (RED), so sums depend on the arrival order — the reference
    asks cuDNN for deterministic algorithms (cli.py:29).  This build bounds the effect instead of removing it.  Two runs of the
    same step from the same state:
      * routing identical; generated images equal to 1e-3 rel. L2 (observed 1.3e-4: the norm statistics of the forward are
        summed in arrival order, and an fp32-ulp difference there flips a few bf16 roundings downstream — a tenth of the 1e-2
        the bf16 arithmetic costs against fp32);
      * run 2 is fed run 1's generated images (bitwise), so everything behind the generator sees identical inputs: the fp32
        arenas (discriminator, aux regressor, router) agree to 1e-5, the bf16 generator backward to 4e-3 (observed 2.1e-3; an fp32-ulp
        difference upstream can flip a bf16 rounding, 2^-9 relative on that element), metrics to 1e-5.
    Without the injection the last-bit image differences are amplified by the networks' max-pool / ReLU decisions (measured:
    aux-regressor gradients of two free runs differ by 9e-3 at 48 samples) — the same conditioning the parity tests document."""
    arch, E, B, seed = "proton", 3, 48, 9
    ocfg, cfg = make_cfg(arch, E)
    outs, inject = [], None
    for run in range(2):
        st = orc.make_state(arch, E, seed, ocfg)
        moe = build_moe(arch, E, cfg, st)
        b, nz = to_dev(orc.make_batch(arch, B, seed)), to_dev(orc.make_noise(arch, B, E, seed))
        if inject is not None:
            nz["img1_sorted"], nz["img2_sorted"] = inject
        m = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=nz)
        torch.cuda.synchronize()
        last = moe._last
        gen_imgs = (last["img1_generated"], last["img2_generated"]) if inject is not None else (last["img1"].clone(), last["img2"].clone())
        if inject is None:
            inject = gen_imgs
        outs.append(({k: float(v) for k, v in m.items()}, last["idx"].cpu(), {k: moe.arena(k).G.clone() for k in "gdar"},
                     {k: moe.arena(k).P.clone() for k in "gdar"}, gen_imgs))
    (m0, i0, g0, p0, im0), (m1, i1, g1, p1, im1) = outs
    assert torch.equal(i0, i1)
    fails = []

    def chk(name, x, y, t):
        try:
            _check(name, x, y, t)
        except AssertionError as e:
            fails.append(str(e))

    chk("run-to-run generated images G(z1)", im1[0], im0[0], 1e-3)
    chk("run-to-run generated images G(z2)", im1[1], im0[1], 1e-3)
    worst = max(abs(m0[k] - m1[k]) / max(1.0, abs(m0[k])) for k in m0)
    log(f"run-to-run (same images): worst metric difference {worst:.3e}")
    for k, tol in (("d", 1e-5), ("a", 1e-5), ("r", 1e-5), ("g", 4e-3)):
        chk(f"run-to-run gradient arena {k} (same images)", g1[k], g0[k], tol)
        chk(f"run-to-run parameters after Adam {k} (same images)", p1[k], p0[k], 1e-4)     # first Adam step is sign-like
    assert worst <= 1e-5, f"metrics moved by {worst:.3e} between two runs of the same step"
    assert not fails, "\n".join(fails)


@pytest.mark.parametrize("cudnn", [0, 1])
def test_reference_fp32_eager_on_the_same_device(tmp_path, cudnn):
    """north_star: "checked against the reference's own PyTorch path on identical seeds, weights and synthetic inputs".  The
    UNMODIFIED reference (oracle/_ref) runs ONE fp32 eager step on cuda:0 in a process of its own (IEEE fp32, no TF32), at the
    benched size E=8 / B=1024; this build runs the same step on the same device.  Routing bit-exact; images 3e-2; losses 3e-2
    + 8e-3; with the reference's images injected: generator gradients 0.2 rel. L2 / cosine 0.985.
    Aux-regressor gradients: 2e-3 against the reference with cuDNN off (ATen's own fp32 convolutions, plain accumulation —
    the same bar as against the CPU oracle).  With cuDNN on, the REFERENCE's gradients themselves move by 1-3e-2 in the early
    layers (cuDNN's algorithm choice changes the activations in the last bits and the regressor's max-pool / ReLU decisions
    on the flat regions of the sparse images flip — the conditioning test_gradient_conditioning_of_the_reference_network
    documents); that variant is held to 6e-2 and logged."""
    import subprocess
    import sys
    import oracle.ref_shim as shim
    if shim.find_reference() is None:
        pytest.skip("no reference copy (oracle/_ref is made by oracle/make_ref.py in the build container)")
    arch, E, B, seed = "proton", 8, 1024, 23
    out = str(tmp_path / "ref_step.pt")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "ref_runner.py"), "step", "--arch", arch, "--experts", str(E),
                        "--batch", str(B), "--seed", str(seed), "--device", "cuda:0", "--cudnn", str(cudnn), "--out", out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    ocfg, cfg = make_cfg(arch, E)
    st = orc.make_state(arch, E, seed, ocfg)
    moe = build_moe(arch, E, cfg, st)
    b, nz = to_dev(orc.make_batch(arch, B, seed)), to_dev(orc.make_noise(arch, B, E, seed))
    H, W = orc.IMAGE_SHAPE[arch]
    masks = [(ref["idx"] == e).nonzero(as_tuple=True)[0] for e in range(E)]
    for key, name in (("fake1", "img1_sorted"), ("fake2", "img2_sorted")):
        nz[name] = torch.cat([ref[key][e].reshape(-1, H * W) if e in ref[key] else torch.zeros(masks[e].numel(), H * W)
                              for e in range(E)]).to(DEV)
    got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=nz)
    torch.cuda.synchronize()
    last = moe._last
    assert last["idx"].cpu().tolist() == ref["idx"].tolist() and last["counts"].cpu().tolist() == ref["counts"].tolist()
    assert last["perm"].cpu().tolist() == torch.cat(masks).tolist()
    tag = f"[same-device reference, cudnn={cudnn}]"
    fails, off = [], 0
    for e in range(E):
        n = masks[e].numel()
        if n >= 2:
            for key, t in (("fake1", last["img1_generated"]), ("fake2", last["img2_generated"])):
                try:
                    _check(f"{tag} {key} expert {e}", t[off:off + n].view(n, 1, H, W), ref[key][e], 3e-2)
                except AssertionError as ex:
                    fails.append(str(ex))
        off += n
    for k, v in ref["metrics"].items():
        g = float(got[k])
        floor = ABS_FLOOR if not k.startswith("std_intensities_experts_") else 2e-3 * abs(ref["metrics"][k.replace("std_", "mean_")]) + ABS_FLOOR
        if not abs(g - v) <= 3e-2 * abs(v) + floor:
            fails.append(f"metric {k}: got {g}, reference {v}")
    worst = {"g": 0.0, "a": 0.0}
    for key, kind, tol, cmin in (("g", "g_grads", 0.2, 0.985), ("a", "a_grads", 2e-3 if cudnn == 0 else 6e-2, 0.999 if cudnn == 0 else 0.99)):
        arena = moe.arena(key)
        for e in range(E):
            for name, gw in ref.get(f"{kind}_{e}", {}).items():
                if float(gw.abs().max()) < 1e-9:
                    continue
                got_g = arena.view(arena.G, name, e)
                worst[key] = max(worst[key], float((got_g.double().cpu() - gw.double()).norm() / gw.double().norm()))
                try:
                    _check(f"{tag} grad {key}{e} {name}", got_g, gw, tol)
                except AssertionError as ex:
                    fails.append(str(ex))
                if gw.numel() >= 64:
                    cos = float(F.cosine_similarity(got_g.flatten().double().cpu(), gw.flatten().double(), dim=0))
                    if cos < cmin:
                        fails.append(f"grad {key}{e} {name}: cosine {cos:.4f}")
    log(f"{tag} reference step on {ref['device']} took {ref['seconds']:.2f} s (torch {ref['torch']}, {ref.get('fp32_precision')}); "
        f"worst relL2: generator grads {worst['g']:.3e}, aux-regressor grads {worst['a']:.3e}")
    assert not fails, "\n".join(fails[:40])


def test_train_step_neutron_E3_B24_golden():
    """Neutron 44x44: BatchNorm batch statistics per (expert, pass), running-stat updates, Dropout(0.2) masks injected."""
    c = golden("train_step_neutron_E3_B24.json")
    moe, st = run_case("neutron", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), gold=c["steps"])


def test_train_step_neutron_E3_B24_injected_images():
    c = golden("train_step_neutron_E3_B24.json")
    moe, st = run_case("neutron", c["E"], c["B"], c["seed"], c.get("router_over"), steps=len(c["steps"]), inject_images=True)
    # BatchNorm buffers after the steps: two running-stat updates per step in the generator (G(z1), G(z2)), one in the aux net
    for e in range(c["E"]):
        gsd, asd = moe.generators[e].state_dict(), moe.aux_regs[e].state_dict()
        for name in ("fc1.1.running_mean", "fc2.1.running_var", "conv_layers.1.running_mean", "conv_layers.6.running_var",
                     "conv_layers.10.running_mean"):
            _check(f"neutron g{e} buffer {name}", gsd[name], st.gens[e][name], 2e-2)
        assert int(gsd["fc2.1.num_batches_tracked"]) == int(st.gens[e]["fc2.1.num_batches_tracked"])
        for name in ("feature_extractor.conv1_bd.0.running_mean", "feature_extractor.conv4_bd.0.running_var",
                     "feature_extractor.reduce.1.running_mean"):
            _check(f"neutron a{e} buffer {name}", asd[name], st.auxs[e][name], 2e-3)
        assert int(asd["feature_extractor.reduce.1.num_batches_tracked"]) == int(st.auxs[e]["feature_extractor.reduce.1.num_batches_tracked"])


def test_generate_neutron_matches_oracle():
    arch, E, N, seed = "neutron", 3, 30, 4
    ocfg, cfg = make_cfg(arch, E)
    st = orc.make_state(arch, E, seed, ocfg)
    moe = build_moe(arch, E, cfg, st).eval()
    g = torch.Generator().manual_seed(5)
    cond = torch.randn(N, 9, generator=g)
    gumbel = -torch.empty(N, E).exponential_(generator=g).log()
    z = torch.randn(N, 10, generator=g)
    want, idx, counts = orc.moe_generate(st, cond, gumbel, z)
    got, gidx = moe.generate(cond.to(DEV), noise=z.to(DEV), gumbel=gumbel.to(DEV), out_dtype=torch.float64, return_routing=True)
    assert tuple(got.shape) == (N, 44, 44) and gidx.cpu().tolist() == idx.tolist()
    # showers = expm1(image): the bf16 image error (1e-2, checked in log space by the train-step tests) is amplified by
    # the pixel value itself for bright pixels
    _check("neutron moe.generate showers (eval-mode BatchNorm)", got, want, 8e-2)
    # eval-mode BatchNorm normalises with the (here: arbitrary, synthetic) running statistics instead of re-centring on
    # the batch, so the bf16 error of each layer is carried forward rather than renormalised: 6e-2 in log space
    _check("neutron moe.generate log-space images", torch.log1p(got), torch.log1p(want), 6e-2)


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
    check = _check
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
    check = _check
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


def test_device_wasserstein_metric_matches_scipy():
    """§8f row 1 (evaluation metric on the device): fused expm1 + 5 channel sums and the sorted-sample Wasserstein distance
    against the reference's numpy masks + scipy.stats.wasserstein_distance on the same showers."""
    import numpy as np
    from scipy.stats import wasserstein_distance
    from expertsim.train.utils import channel_sums_device, sum_channels_parallel, ws_device
    for H, W in ((56, 30), (44, 44)):
        g = torch.Generator().manual_seed(H)
        a = torch.rand(300, H, W, generator=g) * (torch.rand(300, H, W, generator=g) < 0.05) * 5.0
        b = torch.rand(300, H, W, generator=g) * (torch.rand(300, H, W, generator=g) < 0.05) * 5.0
        ch_a = channel_sums_device(a.to(DEV).reshape(300, -1), H, W, True)
        ch_b = channel_sums_device(b.to(DEV).reshape(300, -1), H, W, True)
        want_a = np.array(list(sum_channels_parallel(np.expm1(a.numpy()).astype(np.float64))))
        want_b = np.array(list(sum_channels_parallel(np.expm1(b.numpy()).astype(np.float64))))
        _check(f"channel sums {H}x{W}", ch_a, torch.from_numpy(want_a), 1e-6)
        got = ws_device(ch_a.sort(dim=0).values, ch_b).cpu().numpy()
        for i in range(5):
            w = wasserstein_distance(want_a[:, i], want_b[:, i])
            assert abs(got[i] - w) <= 1e-6 * max(1.0, abs(w)), (i, got[i], w)


def test_evaluate_returns_reference_keys():
    arch, E = "proton", 3
    ocfg, cfg = make_cfg(arch, E)
    st = orc.make_state(arch, E, 2, ocfg)
    moe = build_moe(arch, E, cfg, st).eval()
    batch = orc.make_batch(arch, 64, 1)
    out = moe.evaluate(7, batch["cond"].to(DEV), batch["real_images"][:, 0].numpy(), batch["true_positions"], batch["std"],
                       batch["intensity"], cfg, torch.device(DEV))
    assert set(out) == {"ws_mean", "ws_std", "epoch"} | {f"ws_mean_{i}" for i in range(E)} | {f"ws_std_{i}" for i in range(E)}
    assert out["epoch"] == 7 and out["ws_mean"] > 0 and all(np.isfinite(float(v)) for v in out.values())
