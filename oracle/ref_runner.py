#!/usr/bin/env python
"""Run the UNMODIFIED reference (``oracle/_ref`` or ``/root/reference``) in a process of its own.
TEST / BASELINE INFRASTRUCTURE ONLY — the product package never imports this.

    python oracle/ref_runner.py bench --arch proton --experts 8 --batch 256 --steps 20 --warmup 5
        the reference's ``MoEWrapper.train_step`` (expertsim/models/moe.py:52-504) with its own torch.optim.Adam
        optimizers and its own random draws on the host cores -> one JSON line (samples/s, s/step, showers/s)
    python oracle/ref_runner.py step --arch proton --experts 8 --batch 1024 --seed 3 --device cuda:0 --out x.pt
        ONE reference step in fp32 eager mode on ``device`` with the oracle's deterministic weights / batch / injected
        noise -> metrics, routing, generated images, generator + aux-regressor gradients (what the GPU parity test of
        this build is compared with on the same device)
    python oracle/ref_runner.py train --arch proton --experts 3 --batch 256 --steps 300 --device cuda:0 --out t.json
        a training trajectory of the reference (loss curves + Wasserstein metric of the generated channel sums)

The synthetic weights / batches / noise come from ``oracle.expertsim_oracle`` (make_state / make_batch / make_noise):
the same generators the parity tests feed to the CUDA path.
"""
import argparse
import copy
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import oracle.ref_shim as shim  # noqa: E402


def oracle_cfg(orc, arch, E, **router_over):
    c = copy.deepcopy(orc.DEFAULT_CFG)
    c["model"]["architecture"], c["model"]["n_experts"] = arch, E
    c["model"]["router"].update(router_over)
    return c


def attr_cfg(orc, cfg_d, arch):
    cfg = shim.to_attr(copy.deepcopy(cfg_d))
    cfg.dataset = shim.AttrDict(input_image_shape=list(orc.IMAGE_SHAPE[arch]))
    return cfg


def strict_fp32():
    """IEEE fp32 everywhere (no TF32 in cuDNN convolutions / cuBLAS), deterministic cuDNN algorithms as the reference's
    cli.py:28-30 asks for — the comparison is against the reference's fp32 arithmetic, not against TF32 rounding."""
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    try:                                             # torch >= 2.9 precision API
        torch.backends.fp32_precision = "ieee"
        torch.backends.cuda.matmul.fp32_precision = "ieee"
        torch.backends.cudnn.fp32_precision = "ieee"
        torch.backends.cudnn.conv.fp32_precision = "ieee"
        torch.backends.cudnn.rnn.fp32_precision = "ieee"
    except Exception:                                # older torch: the legacy switches
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    return {"cudnn_conv": str(getattr(getattr(torch.backends.cudnn, "conv", None), "fp32_precision", "n/a")),
            "matmul": str(getattr(torch.backends.cuda.matmul, "fp32_precision", "n/a"))}


def queue_noise(orc, inj, arch, E, masks, noise):
    """order of the reference's draws inside train_step: Gumbel noise once, then per live expert z1, z2 and the
    dropout masks of G(z1), G(z2), aux (moe.py:76,144,535 + the nn.Dropout sites)"""
    inj.expo_q.append(noise["expo"])
    for e in range(E):
        mk = masks[e]
        if mk.numel() <= 1:
            continue
        inj.randn_q += [noise["z1"][mk], noise["z2"][mk]]
        for tag, kind in (("g1", "generator"), ("g2", "generator"), ("a", "aux_reg")):
            for site, _, p in orc.dropout_sites(arch, kind):
                inj.drop_q.append((noise[f"drop.{tag}.{site}"][mk], p))


def cmd_bench(a):
    import oracle.expertsim_oracle as orc
    shim.install_reference()
    torch.set_num_threads(a.threads or (os.cpu_count() or 1))
    cfg_d = oracle_cfg(orc, a.arch, a.experts)
    st = orc.make_state(a.arch, a.experts, 0, cfg_d, identical_experts=True)    # deepcopy semantics of MoEWrapper.__init__
    moe, (g_o, d_o, a_o, r_o) = shim.build_reference_moe(a.arch, a.experts, attr_cfg(orc, cfg_d, a.arch), orc.IMAGE_SHAPE[a.arch], st)
    moe.train()
    dev = torch.device("cpu")
    ts = []
    for i in range(a.warmup + a.steps):
        b = orc.make_batch(a.arch, a.batch, i)
        t0 = time.perf_counter()
        m = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], a_o, g_o, d_o, r_o, None, dev)
        float(m["gen_loss"])
        if i >= a.warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    # batch inference of the reference: get_predictions_from_generator_results (train/utils.py:179-205)
    from expertsim.train.utils import get_predictions_from_generator_results
    n = a.showers
    g = torch.Generator().manual_seed(0)
    cond, z = torch.randn(n, 9, generator=g), torch.randn(n, 10, generator=g)
    gen = moe.generators[0].eval()
    get_predictions_from_generator_results(64, 64, 10, dev, cond[:64], gen, shape_images=orc.IMAGE_SHAPE[a.arch], input_noise=z[:64])
    t0 = time.perf_counter()
    get_predictions_from_generator_results(256, n, 10, dev, cond, gen, shape_images=orc.IMAGE_SHAPE[a.arch], input_noise=z)
    shw = n / (time.perf_counter() - t0)
    print(json.dumps({"samples_per_s": a.batch / dt, "s_per_step": dt, "showers_per_s": shw, "threads": torch.get_num_threads(),
                      "batch": a.batch, "steps": a.steps, "warmup": a.warmup, "loss": float(m["gen_loss"]),
                      "reference_root": shim.find_reference()}), flush=True)


def cmd_step(a):
    import oracle.expertsim_oracle as orc
    shim.install_reference()
    prec = strict_fp32()
    torch.backends.cudnn.enabled = bool(a.cudnn)
    prec["cudnn_enabled"] = bool(a.cudnn)
    dev = torch.device(a.device)
    E, arch, B = a.experts, a.arch, a.batch
    cfg_d = oracle_cfg(orc, arch, E)
    st = orc.make_state(arch, E, a.seed, cfg_d)
    moe, (g_o, d_o, a_o, r_o) = shim.build_reference_moe(arch, E, attr_cfg(orc, cfg_d, arch), orc.IMAGE_SHAPE[arch], st, dev)
    moe.train()
    batch, noise = orc.make_batch(arch, B, a.seed), orc.make_noise(arch, B, E, a.seed)
    with torch.no_grad():
        gs, _ = orc.router_forward(st.router, batch["cond"], noise["gumbel"], orc.router_tau(cfg_d["model"]["router"], 0))
        idx, counts, masks = orc.route(gs, E)
    fakes = {e: [] for e in range(E)}
    hooks = [moe.generators[e].register_forward_hook(lambda m, i, o, e=e: fakes[e].append(o.detach().float().cpu())) for e in range(E)]
    bd = {k: v.to(dev) for k, v in batch.items()}
    with shim.NoiseInjector(dev) as inj:
        queue_noise(orc, inj, arch, E, masks, noise)
        t0 = time.perf_counter()
        m = moe.train_step(0, bd["cond"], bd["real_images"], bd["true_positions"], bd["std"], bd["intensity"], a_o, g_o, d_o, r_o, None, dev)
        metrics = {k: float(v) for k, v in m.items()}
        dt = time.perf_counter() - t0
    for h in hooks:
        h.remove()
    out = {"metrics": metrics, "idx": idx, "counts": counts, "seconds": dt, "device": str(dev), "torch": torch.__version__, "fp32_precision": prec,
           "fake1": {e: f[0] for e, f in fakes.items() if f}, "fake2": {e: f[1] for e, f in fakes.items() if len(f) > 1}}
    # gradients left in .grad by the step: generator and aux regressor hold exactly the generator step's gradients; the
    # discriminator's hold D-step + G-step contributions (moe.py:564) and are therefore not exported
    for key, mods in (("g_grads", moe.generators), ("a_grads", moe.aux_regs)):
        for e in range(E):
            if counts[e] > 1:
                out[f"{key}_{e}"] = {n: p.grad.detach().float().cpu() for n, p in mods[e].named_parameters() if p.grad is not None}
    out["router_after"] = {k: v.detach().cpu() for k, v in moe.router.state_dict().items()}
    out["disc_after_fc3"] = {e: moe.discriminators[e].state_dict()["fc3.weight_orig"].detach().cpu() for e in range(E)}
    torch.save(out, a.out)
    print(json.dumps({"ok": True, "seconds": dt, "gen_loss": metrics["gen_loss"], "counts": counts.tolist()}), flush=True)


def cmd_train(a):
    """loss curves + final Wasserstein metric of the reference trained for ``steps`` steps on a fixed synthetic set"""
    import oracle.expertsim_oracle as orc
    shim.install_reference()
    strict_fp32()
    dev = torch.device(a.device)
    E, arch, B = a.experts, a.arch, a.batch
    cfg_d = oracle_cfg(orc, arch, E)
    st = orc.make_state(arch, E, a.seed, cfg_d, identical_experts=True)
    moe, (g_o, d_o, a_o, r_o) = shim.build_reference_moe(arch, E, attr_cfg(orc, cfg_d, arch), orc.IMAGE_SHAPE[arch], st, dev)
    moe.train()
    torch.manual_seed(a.seed)
    pool = [{k: v.to(dev) for k, v in orc.make_batch(arch, B, 1000 + i).items()} for i in range(a.pool)]
    curve = []
    for i in range(a.steps):
        b = pool[i % a.pool]
        m = moe.train_step(i // a.pool, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], a_o, g_o, d_o, r_o, None, dev)
        if i % a.every == 0 or i == a.steps - 1:
            curve.append({"step": i, **{k: float(m[k]) for k in ("gen_loss", "disc_loss", "div_loss", "intensity_loss", "aux_reg_loss", "router_loss")}})
    # Wasserstein metric of the 5 channel sums, generated vs the pool's real showers (moe.py:644-692, train/utils.py:117-176)
    from scipy.stats import wasserstein_distance
    moe.eval()
    ws_runs = []
    with torch.no_grad():
        cond = torch.cat([b["cond"] for b in pool])
        real = torch.cat([b["real_images"] for b in pool]).reshape(-1, *orc.IMAGE_SHAPE[arch]).cpu()
        ch_real = orc.sum_channels(torch.expm1(real).double()).numpy()
        for r in range(a.ws_runs):
            g = torch.Generator(device=dev).manual_seed(77 + r)
            gates, _ = moe.router(cond)
            idx = gates.argmax(dim=1)
            z = torch.randn(cond.shape[0], 10, device=dev, generator=g)
            img = torch.zeros(cond.shape[0], *orc.IMAGE_SHAPE[arch], device=dev)
            for e in range(E):
                mk = (idx == e).nonzero(as_tuple=True)[0]
                if mk.numel():
                    img[mk] = moe.generators[e](z[mk], cond[mk]).reshape(-1, *orc.IMAGE_SHAPE[arch])
            ch = orc.sum_channels(torch.expm1(img.cpu()).double()).numpy()
            ws_runs.append(float(sum(wasserstein_distance(ch_real[:, c], ch[:, c]) for c in range(5)) / 5))
    json.dump({"impl": "reference", "arch": arch, "E": E, "B": B, "steps": a.steps, "device": str(dev), "curve": curve,
               "ws_mean": sum(ws_runs) / len(ws_runs), "ws_runs": ws_runs}, open(a.out, "w"))
    print(json.dumps({"ok": True, "ws_mean": sum(ws_runs) / len(ws_runs), "last": curve[-1]}), flush=True)


def cmd_names(a):
    """parameter order (= torch.optim.Adam's state index order) and state_dict keys of every reference module"""
    shim.install_reference()
    from expertsim.models.routers.router import RouterNetwork
    out = {}
    for arch in ("proton", "neutron"):
        G, D, A = shim.reference_classes(arch)
        for kind, m in (("generator", G(10, 9, 0.1, 1e-3)), ("discriminator", D(9)), ("aux_reg", A(1e-3))):
            out[f"{arch}.{kind}"] = {"params": [n for n, _ in m.named_parameters()], "state_dict": list(m.state_dict())}
    r = RouterNetwork(9, 5)
    out["router"] = {"params": [n for n, _ in r.named_parameters()], "state_dict": list(r.state_dict())}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["bench", "step", "train", "names"])
    ap.add_argument("--arch", default="proton")
    ap.add_argument("--experts", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--showers", type=int, default=512)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--ws-runs", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--cudnn", type=int, default=1, help="0: torch.backends.cudnn.enabled = False (native ATen convolutions)")
    a = ap.parse_args()
    {"bench": cmd_bench, "step": cmd_step, "train": cmd_train, "names": cmd_names}[a.cmd](a)


if __name__ == "__main__":
    main()
