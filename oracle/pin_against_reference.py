"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.json.

Runs only in the build container (needs /root/reference).  It imports the reference's own classes through the
15-line shim of SURVEY.md §8(c) (stub matplotlib/seaborn/wandb, bypass the broken expertsim/models/__init__.py),
loads the deterministic synthetic weights of ``oracle.expertsim_oracle.make_weights`` into them, injects the
same noise the oracle receives (by patching torch.randn / Tensor.exponential_ / F.dropout with queue-fed
versions), runs ``MoEWrapper.train_step`` and ``get_predictions_from_generator_results`` and stores the
REFERENCE's outputs.  It also runs the oracle on the same inputs and refuses to write a fixture the oracle does
not reproduce.  TEST INFRASTRUCTURE ONLY.

usage: python oracle/pin_against_reference.py [--out tests/golden]
"""
import argparse
import copy
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import oracle.ref_shim as shim  # noqa: E402
from oracle.ref_shim import AttrDict, NoiseInjector, to_attr  # noqa: E402,F401


def tensor_digest(t):
    t = t.detach().double().flatten()
    return {"sum": float(t.sum()), "l2": float(t.norm()), "head": [float(x) for x in t[:6]]}


def build_reference_moe(orc, arch, E, seed, cfg):
    st = orc.make_state(arch, E, seed, orc_cfg(orc, arch, E))
    moe, _ = shim.build_reference_moe(arch, E, cfg, orc.IMAGE_SHAPE[arch], state=st)
    return moe, st


def orc_cfg(orc, arch, E, **router_over):
    c = copy.deepcopy(orc.DEFAULT_CFG)
    c["model"]["architecture"] = arch
    c["model"]["n_experts"] = E
    c["model"]["router"].update(router_over)
    return c


def run_train_case(orc, arch, E, B, seed, steps, router_over=None):
    router_over = router_over or {}
    cfg_d = orc_cfg(orc, arch, E, **router_over)
    cfg = to_attr(copy.deepcopy(cfg_d))
    cfg.dataset = AttrDict(input_image_shape=list(orc.IMAGE_SHAPE[arch]))
    moe, st = build_reference_moe(orc, arch, E, seed, cfg)
    st.cfg = cfg_d
    from expertsim.train.training_setup import setup_optimizers
    g_o, d_o, a_o, r_o = setup_optimizers(moe, cfg)
    moe.train()
    case = {"arch": arch, "E": E, "B": B, "seed": seed, "router_over": router_over, "steps": []}
    for step in range(steps):
        batch = orc.make_batch(arch, B, seed + 17 * step)
        noise = orc.make_noise(arch, B, E, seed + 31 * step)
        # routing (needed to order the injected per-expert noise); asserted equal to the reference's below
        with torch.no_grad():
            gs, _ = orc.router_forward(st.router, batch["cond"], noise["gumbel"], orc.router_tau(cfg_d["model"]["router"], 0))
            idx, counts, masks = orc.route(gs, E)
        with NoiseInjector() as inj:
            inj.expo_q.append(noise["expo"])
            for e in range(E):
                mk = masks[e]
                if mk.numel() <= 1:
                    continue
                inj.randn_q += [noise["z1"][mk], noise["z2"][mk]]
                for tag, kind in (("g1", "generator"), ("g2", "generator"), ("a", "aux_reg")):
                    for site, _, p in orc.dropout_sites(arch, kind):
                        inj.drop_q.append((noise[f"drop.{tag}.{site}"][mk], p))
            ref_metrics = moe.train_step(0, batch["cond"], batch["real_images"], batch["true_positions"], batch["std"],
                                         batch["intensity"], a_o, g_o, d_o, r_o, None, torch.device("cpu"))
        ref_metrics = {k: float(v) for k, v in ref_metrics.items()}
        orc_metrics, aux = orc.train_step(st, batch, noise, epoch=0)
        # the oracle must reproduce the reference before anything is written
        for k, v in ref_metrics.items():
            assert abs(orc_metrics[k] - v) <= 2e-4 * max(1.0, abs(v)), (arch, step, k, orc_metrics[k], v)
        digests = {}
        for e in range(E):
            for tag, mod, sd in (("generators", moe.generators[e], st.gens[e]), ("discriminators", moe.discriminators[e], st.discs[e]),
                                 ("aux_regs", moe.aux_regs[e], st.auxs[e])):
                rsd = mod.state_dict()
                for k in sd:
                    if not sd[k].dtype.is_floating_point:
                        assert int(rsd[k]) == int(sd[k]), (tag, e, k, rsd[k], sd[k])
                        continue
                    err = (rsd[k] - sd[k]).norm() / rsd[k].norm().clamp_min(1e-12)
                    assert err < 5e-4, (arch, step, tag, e, k, float(err))
                    if k.endswith(("weight", "weight_orig", "weight_u", "running_var")):
                        digests[f"{tag}.{e}.{k}"] = tensor_digest(rsd[k])
        for k, v in moe.router.state_dict().items():
            assert (v - st.router[k]).norm() / v.norm() < 5e-4, k
            digests[f"router.{k}"] = tensor_digest(v)
        case["steps"].append({"metrics": ref_metrics, "idx": [int(i) for i in idx], "counts": [int(c) for c in counts],
                              "digests": digests,
                              "fake1_digest": {str(e): tensor_digest(t) for e, t in aux["fake1"].items()}})
        print(f"[pin] {arch} E={E} B={B} step {step}: reference reproduced "
              f"(gen_loss {ref_metrics['gen_loss']:.6f}, disc_loss {ref_metrics['disc_loss']:.6f}, counts {case['steps'][-1]['counts']})")
    return case


def run_module_cases(orc, arch, seed, n=4):
    """Per-module forward outputs of the reference classes (eval and train mode) on n samples."""
    cfg = to_attr(orc_cfg(orc, arch, 3))
    cfg.dataset = AttrDict(input_image_shape=list(orc.IMAGE_SHAPE[arch]))
    moe, st = build_reference_moe(orc, arch, 3, seed, cfg)
    batch = orc.make_batch(arch, n, seed)
    noise = orc.make_noise(arch, n, 3, seed)
    out = {"arch": arch, "seed": seed, "n": n}
    G, D, A = moe.generators[1], moe.discriminators[1], moe.aux_regs[1]
    sdG, sdD, sdA = st.gens[1], st.discs[1], st.auxs[1]
    with torch.no_grad():
        G.eval(); D.eval(); A.eval(); moe.router.eval()
        img = G(noise["z1"], batch["cond"])
        o_img = orc.generator_forward(arch, sdG, noise["z1"], batch["cond"], training=False)
        assert torch.allclose(img, o_img, atol=2e-5, rtol=1e-4), float((img - o_img).abs().max())
        out["gen_eval_image"] = img.flatten().tolist()
        sc, lat = D(batch["real_images"], batch["cond"])
        o_sc, o_lat = orc.discriminator_forward(arch, sdD, batch["real_images"], batch["cond"], training=False)
        assert torch.allclose(sc, o_sc, atol=1e-5, rtol=1e-4) and torch.allclose(lat, o_lat, atol=1e-5, rtol=1e-4)
        out["disc_eval_score"], out["disc_eval_latent"] = sc.flatten().tolist(), lat.flatten().tolist()
        co = A(batch["real_images"])
        o_co = orc.aux_forward(arch, sdA, batch["real_images"], training=False)
        assert torch.allclose(co, o_co, atol=1e-5, rtol=1e-4)
        out["aux_eval_coords"] = co.flatten().tolist()
        with NoiseInjector() as inj:
            inj.expo_q.append(noise["expo"])
            gates, logits = moe.router(batch["cond"], tau=1.2)
        o_g, o_l = orc.router_forward(st.router, batch["cond"], noise["gumbel"], 1.2)
        assert torch.allclose(gates, o_g, atol=1e-6, rtol=1e-5) and torch.allclose(logits, o_l, atol=1e-6, rtol=1e-5)
        out["router_gates"], out["router_logits"] = gates.flatten().tolist(), logits.flatten().tolist()
        # train-mode discriminator: one power iteration advances u, v
        D.train()
        sc_t, lat_t = D(batch["real_images"], batch["cond"])
        o_sc_t, o_lat_t = orc.discriminator_forward(arch, sdD, batch["real_images"], batch["cond"], training=True)
        assert torch.allclose(sc_t, o_sc_t, atol=1e-5, rtol=1e-4)
        assert torch.allclose(D.state_dict()["fc1.0.weight_u"], sdD["fc1.0.weight_u"], atol=1e-6)
        out["disc_train_score"] = sc_t.flatten().tolist()
        out["disc_train_fc1_u_digest"] = tensor_digest(D.state_dict()["fc1.0.weight_u"])
        # loss tails on fixed inputs
        from expertsim.models.moe import MoEWrapper
        lat2 = lat.flip(0) * 0.9 + 0.05
        sdi = MoEWrapper.sdi_gan_regularization(lat, lat2, noise["z1"], noise["z2"], batch["std"], 0.1)
        assert abs(float(sdi) - float(orc.sdi_gan_regularization(lat, lat2, noise["z1"], noise["z2"], batch["std"], 0.1))) < 1e-6
        il, s, s_std, s_mean = MoEWrapper.intensity_regularization(img, batch["intensity"], 1e-3)
        out["sdi_loss"], out["intensity_loss"] = float(sdi), float(il)
        out["photon_sums"], out["photon_std"], out["photon_mean"] = s.flatten().tolist(), float(s_std), float(s_mean)
        out["regressor_loss"] = float(A.regressor_loss(batch["true_positions"], co))
        # batch inference helper
        from expertsim.train.utils import get_predictions_from_generator_results
        res, raw = get_predictions_from_generator_results(3, n, 10, torch.device("cpu"), batch["cond"], G,
                                                          shape_images=orc.IMAGE_SHAPE[arch], input_noise=noise["z2"])
        o_res, o_raw = orc.generate(arch, sdG, noise["z2"], batch["cond"], batch_size=3)
        assert np.allclose(res, o_res.numpy(), atol=1e-4, rtol=1e-4)
        out["infer_photon_sums"] = res.sum(axis=(1, 2)).tolist()
        from expertsim.train.utils import sum_channels_parallel
        ch = np.array(list(sum_channels_parallel(res)))
        assert np.allclose(ch, orc.sum_channels(torch.from_numpy(res)).numpy(), rtol=1e-9, atol=1e-9)
        out["infer_channels"] = ch.flatten().tolist()
    print(f"[pin] {arch} module cases reproduced")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    shim.install_reference(REF)
    import oracle.expertsim_oracle as orc
    os.makedirs(args.out, exist_ok=True)
    meta = {"torch": torch.__version__, "generator": "oracle/pin_against_reference.py",
            "reference": "patrick-bedkowski/Generative-DNN-for-Physics-Simulations-CERN @ /root/reference"}
    for arch in ("proton", "neutron"):
        mods = run_module_cases(orc, arch, seed=5)
        json.dump({"meta": meta, "case": mods}, open(os.path.join(args.out, f"modules_{arch}.json"), "w"))
        tr = run_train_case(orc, arch, E=3, B=24, seed=7, steps=2)
        json.dump({"meta": meta, "case": tr}, open(os.path.join(args.out, f"train_step_{arch}_E3_B24.json"), "w"))
    # entropy + distribution router losses on, E=2, and the E=1 (router skipped) path
    tr = run_train_case(orc, "proton", E=2, B=12, seed=11, steps=1, router_over={"util_strength": 0.1, "ed_strength": 0.01})
    json.dump({"meta": meta, "case": tr}, open(os.path.join(args.out, "train_step_proton_E2_B12_ent_ed.json"), "w"))
    tr = run_train_case(orc, "proton", E=1, B=8, seed=13, steps=1)
    json.dump({"meta": meta, "case": tr}, open(os.path.join(args.out, "train_step_proton_E1_B8.json"), "w"))
    # an expert left with <=1 sample (skip path, models/moe.py:126-135): E=8 on 10 samples
    tr = run_train_case(orc, "proton", E=8, B=10, seed=19, steps=1)
    json.dump({"meta": meta, "case": tr}, open(os.path.join(args.out, "train_step_proton_E8_B10_skip.json"), "w"))
    print("[pin] all fixtures written to", args.out)


if __name__ == "__main__":
    main()
