#!/usr/bin/env python
"""Accuracy equivalence over a training TRAJECTORY (north_star: ">= 90 % of the reference's accuracy-equivalent output").

    python tools/train_equivalence.py [--arch proton] [--experts 3] [--batch 256] [--steps 300] [--out profiles/...json]

Trains the SAME system twice on the same synthetic set from the same initial weights:
  * the UNMODIFIED reference (oracle/_ref) in fp32 eager PyTorch on cuda:0 (oracle/ref_runner.py train, a process of its own),
  * this build (bf16 tensor-core generator, fp32 everything else, fused Adam),
each with its own random draws (Gumbel noise, z, dropout) — so the two runs are two samples of the same stochastic training
process, not a bit-wise replay.  Compared: the loss curves (means over the last quarter of the steps) and the reference's
evaluation metric — mean Wasserstein distance of the 5 channel sums between generated and real showers (moe.py:644-692,
train/utils.py:117-176) — of the trained generators.  The seed-to-seed spread of the reference itself is measured with a
second reference run (other seed) so the band is stated against the process's own noise.
"""
import argparse
import copy
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

KEYS = ("gen_loss", "disc_loss", "div_loss", "intensity_loss", "aux_reg_loss", "router_loss")


def run_reference(a, seed, out):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "train", "--arch", a.arch, "--experts", str(a.experts),
           "--batch", str(a.batch), "--steps", str(a.steps), "--seed", str(seed), "--device", "cuda:0", "--pool", str(a.pool),
           "--every", str(a.every), "--ws-runs", str(a.ws_runs), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"reference run failed:\n{r.stderr[-3000:]}")
    return json.load(open(out))


def run_this_build(a, seed):
    import oracle.expertsim_oracle as orc
    from expertsim.config import Config
    from expertsim.train.loop import setup_moe_system
    from expertsim.train.training_setup import setup_optimizers
    from scipy.stats import wasserstein_distance
    dev = torch.device("cuda", 0)
    E, arch, B = a.experts, a.arch, a.batch
    H, W = orc.IMAGE_SHAPE[arch]
    ocfg = copy.deepcopy(orc.DEFAULT_CFG)
    ocfg["model"]["n_experts"], ocfg["model"]["architecture"] = E, arch
    ocfg["dataset"] = {"input_image_shape": [H, W]}
    st = orc.make_state(arch, E, seed, ocfg, identical_experts=True)
    cfg = Config(ocfg)
    moe = setup_moe_system(cfg, dev)
    for e in range(E):
        moe.generators[e].load_state_dict(st.gens[e])
        moe.discriminators[e].load_state_dict(st.discs[e])
        moe.aux_regs[e].load_state_dict(st.auxs[e])
    moe.router.load_state_dict(st.router)
    g_o, d_o, a_o, r_o = setup_optimizers(moe, cfg)
    moe.train()
    torch.manual_seed(seed)
    pool = [{k: v.to(dev) for k, v in orc.make_batch(arch, B, 1000 + i).items()} for i in range(a.pool)]
    curve = []
    for i in range(a.steps):
        b = pool[i % a.pool]
        m = moe.train_step(i // a.pool, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], a_o, g_o, d_o, r_o, None, dev)
        if i % a.every == 0 or i == a.steps - 1:
            curve.append({"step": i, **{k: float(m[k]) for k in KEYS}})
    moe.eval()
    cond = torch.cat([b["cond"] for b in pool])
    real = torch.cat([b["real_images"] for b in pool]).reshape(-1, H, W).cpu()
    ch_real = orc.sum_channels(torch.expm1(real).double()).numpy()
    ws_runs = []
    for r in range(a.ws_runs):
        torch.manual_seed(77 + r)
        img = moe.generate(cond, out_dtype=torch.float64).cpu()            # routed batch inference, expm1 applied
        ch = orc.sum_channels(img).numpy()
        ws_runs.append(float(sum(wasserstein_distance(ch_real[:, c], ch[:, c]) for c in range(5)) / 5))
    return {"impl": "this build (B200, bf16 generator)", "curve": curve, "ws_mean": sum(ws_runs) / len(ws_runs), "ws_runs": ws_runs}


def tail_mean(curve, key, frac=0.25):
    n = max(1, int(len(curve) * frac))
    return sum(c[key] for c in curve[-n:]) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="proton")
    ap.add_argument("--experts", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--every", type=int, default=5)
    ap.add_argument("--ws-runs", type=int, default=3)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_equivalence.json"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    tmp = a.out + ".ref.json"
    ref = run_reference(a, a.seed, tmp)
    ref2 = run_reference(a, a.seed + 1, tmp)        # the reference's own seed-to-seed spread
    mine = run_this_build(a, a.seed)
    rows = {}
    for k in KEYS:
        r1, r2, m = tail_mean(ref["curve"], k), tail_mean(ref2["curve"], k), tail_mean(mine["curve"], k)
        rows[k] = {"reference": r1, "reference_other_seed": r2, "this_build": m,
                   "abs_diff": abs(m - r1), "reference_seed_spread": abs(r1 - r2)}
    ws = {"reference": ref["ws_mean"], "reference_other_seed": ref2["ws_mean"], "this_build": mine["ws_mean"],
          "ratio_reference_over_this_build": ref["ws_mean"] / mine["ws_mean"] if mine["ws_mean"] else None}
    out = {"config": {"arch": a.arch, "n_experts": a.experts, "batch": a.batch, "steps": a.steps, "pool_batches": a.pool,
                      "what": "same initial weights and data, independent random draws; tail means over the last quarter of the steps"},
           "loss_tail_means": rows, "wasserstein_channel_sums": ws,
           "accuracy_equivalent_pct": round(100.0 * min(1.0, ws["ratio_reference_over_this_build"]), 1) if ws["ratio_reference_over_this_build"] else None,
           "curves": {"reference": ref["curve"], "reference_other_seed": ref2["curve"], "this_build": mine["curve"]}}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "curves"}, indent=1))
    os.remove(tmp)


if __name__ == "__main__":
    main()
