#!/usr/bin/env python
"""Achieved HBM rate of the loss-tail (K4) and gating (K1) kernels at a batch that does not fit the L2 (the train step's
1024 rows are 7-20 MB per launch: latency, not bandwidth).  ALGORITHMIC bytes per row: loss reduce reads the
image (6720 B) + 596 B of latents / noise / scalars; loss grads reads the same and read-modify-writes the image gradient
(+ 13 440 B) and writes 536 B of gradients; expm1+scatter reads and writes one image; the gather moves one image; the router
reads 36 B cond + 4E B gumbel and writes ~900 B of activations.  Prints one line per kernel with GB/s and the fraction of the
measured HBM copy rate (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200"))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def measure(dev="cuda", R=65536, E=8, peak=None):
    """-> {kernel: dict(achieved GB/s, peak, frac, bytes_per_row, rows, us)}; every launch moves more than the 126 MB L2
    holds (65 536 rows x 6.7-20 KB), timed with CUDA events on the launching stream."""
    from expertsim import _lib as L
    HW = 56 * 30
    if peak is None:
        peak = 6455.6
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
    per = R // E
    grp = torch.tensor([[i * per, per, i, per] for i in range(E)], dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
    img = (torch.rand(R, HW, generator=g, device=dev) < 0.05).float() * 3
    lat1, lat2, z1, z2 = rnd(R, 64), rnd(R, 64), rnd(R, 10), rnd(R, 10)
    std, inten, coords, pos, score = torch.rand(R, 1, device=dev), torch.rand(R, 1, device=dev) * 100, rnd(R, 2), rnd(R, 2), rnd(R, 1)
    s_out, div_out = torch.zeros(R, device=dev), torch.zeros(R, device=dev)
    sums = torch.zeros(E, 8, dtype=torch.float64, device=dev)
    d_s, d_l1, d_l2, d_c = torch.zeros(R, device=dev), torch.zeros(R, 64, device=dev), torch.zeros(R, 64, device=dev), torch.zeros(R, 2, device=dev)
    d_img, losses = torch.zeros(R, HW, device=dev), torch.zeros(E, 6, device=dev)
    args = (img, HW, lat1, lat2, z1, z2, std, inten, coords, pos)
    perm = torch.randperm(R, generator=g, device=dev).to(torch.int32)
    out32 = torch.empty(R, HW, device=dev)
    gat = torch.empty(R, HW, device=dev)
    cond, gumbel = rnd(R, 9), -torch.empty(R, E, device=dev).exponential_().log()
    W = [rnd(128, 9) * .3, rnd(128) * .1, rnd(64, 128) * .1, rnd(64) * .1, rnd(32, 64) * .1, rnd(32) * .1, rnd(E, 32) * .1, rnd(E) * .1]
    nblk = (R + 255) // 256
    ro = dict(logits=torch.empty(R, E, device=dev), gates=torch.empty(R, E, device=dev), idx=torch.empty(R, dtype=torch.int64, device=dev),
              h1=torch.empty(R, 128, device=dev), h2=torch.empty(R, 64, device=dev), h3=torch.empty(R, 32, device=dev),
              hist=torch.empty(nblk, E, dtype=torch.int32, device=dev))
    # fused Adam over an arena of E x 8 Mi parameters (4 arrays x 268 MB: P, G, M, V; 28 B per parameter and step)
    n_ad = 8 << 20
    aP, aG, aM, aV = (rnd(E, n_ad) * 0.01 for _ in range(4))
    aV.abs_()
    a_steps = torch.zeros(E, dtype=torch.int32, device=dev)
    cases = {
        "es_gen_loss_reduce": (lambda: L.call("es_gen_loss_reduce", *args, score, grp, E, R, s_out, div_out, sums), HW * 4 + 596 + 24, R),
        "es_gen_loss_grads": (lambda: L.call("es_gen_loss_grads", *args, s_out, div_out, grp, E, R, sums, R, 0.1, 1e-3, 1e-3, d_s, d_l1, d_l2,
                                             d_c, d_img, losses), HW * 4 * 3 + 596 + 24 + 536, R),
        "es_expm1_scatter": (lambda: L.call("es_expm1_scatter", img, perm, R, HW, None, out32), HW * 4 * 2, R),
        "es_gather_rows (images)": (lambda: L.call("es_gather_rows", img, perm, R, HW, gat), HW * 4 * 2, R),
        "es_hinge_d": (lambda: L.call("es_hinge_d", s_out, div_out, grp, E, None, R, d_s, lat1[:, 0].contiguous(), losses[:, 0].contiguous()), 16, R),
        "es_router_fwd": (lambda: L.call("es_router_fwd", cond, R, E, *W, gumbel, 1.2, ro["logits"], ro["gates"], ro["idx"], ro["h1"], ro["h2"],
                                         ro["h3"], ro["hist"]), 36 + 4 * E + (128 + 64 + 32 + 2 * E) * 4 + 8, R),
        "es_adam_step (adam_vec4)": (lambda: L.call("es_adam_step", aP, aG, aM, aV, n_ad, n_ad, E, 1e-4, 0.9, 0.999, 1e-8, a_steps, None),
                                     28, E * n_ad),
    }
    # generator norm layers at the train step's size (2048 generator rows, proton conv2 output 55 x 29 x 128, 836 MB per
    # tensor): forward = 1 read + 1 write with the statistics from the conv epilogue; backward = 2 reads + 1 write algorithmic
    # (the two-pass kernel really reads x and dy twice)
    BF = torch.bfloat16
    Rg, P, C = 2048, 55 * 29, 128
    pg = Rg // E
    ggrp = torch.tensor([[i * pg, pg, i, pg // 2] for i in range(E)], dtype=torch.int32, device=dev)
    gx = rnd(Rg, P, C).to(BF)
    gy, gdy, gdx = torch.empty_like(gx), rnd(Rg, P, C).to(BF), torch.empty_like(gx)
    gps = torch.rand(Rg, C // 2, 2, device=dev) * 100
    gps[..., 1] += 1e5
    gst = torch.empty(Rg, 32, 2, device=dev)
    gam, bet = torch.ones(E, C, device=dev), torch.zeros(E, C, device=dev)
    dgm, dbt, dbi = torch.zeros(E, C, device=dev), torch.zeros(E, C, device=dev), torch.zeros(E, C, device=dev)
    L.call("es_gn_lrelu_fwd", gx, gam, bet, C, P, C, 32, ggrp, E, Rg, gy, gst)
    cases["es_gn_lrelu_apply_fwd (conv2 output)"] = (
        lambda: L.call("es_gn_lrelu_apply_fwd", gx, gps, gam, bet, C, 55, 29, 29, C, 32, ggrp, E, Rg, gy, gst), P * C * 2 * 2, Rg)
    cases["es_gn_lrelu_bwd (conv2 output)"] = (
        lambda: L.call("es_gn_lrelu_bwd", gdy, 55, 29, 55, 29, gx, gst, gam, bet, C, C, 32, ggrp, E, Rg, gdx, dgm, dbt, dbi), P * C * 2 * 3, Rg)
    # fc2 backward (TMA-fed dense kernels): the weight gradient reads dy2 (bf16) once and writes E x 92160 x 256 fp32; the data
    # gradient reads dy2 and the bf16 weights once
    F2, K2 = 92160, 256
    dy2, h1 = rnd(Rg, F2).to(BF), rnd(Rg, K2).to(BF)
    w2 = (rnd(E, F2, K2) * 0.05).to(BF)
    dw2 = torch.empty(E, F2, K2, device=dev)
    dh1 = torch.zeros(Rg, K2, device=dev)
    xpad = torch.empty(E * ((Rg + 63) // 64) * 64, K2, dtype=BF, device=dev)
    cases["es_dense_wgrad (fc2)"] = (lambda: L.call("es_dense_wgrad", dy2, h1, dw2, F2 * K2, F2, K2, None, ggrp, E, Rg, xpad),
                                     Rg * F2 * 2 + E * F2 * K2 * 4, 1)
    cases["es_dense_dgrad (fc2)"] = (lambda: L.call("es_dense_dgrad", dy2, w2, dh1, F2, K2, ggrp, E, Rg), Rg * F2 * 2 + E * F2 * K2 * 2, 1)
    out = {}
    for name, (fn, bytes_per_unit, units) in cases.items():
        t = timeit(fn)
        gbs = bytes_per_unit * units / t / 1e9
        out[name] = {"achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                     "algorithmic_bytes_per_unit": bytes_per_unit, "units_per_launch": units, "us_per_launch": round(t * 1e6, 1)}
    return out


if __name__ == "__main__":
    for name, r in measure().items():
        print(f"{name:26s} units={r['units_per_launch']} {r['us_per_launch']:9.1f} us  {r['algorithmic_bytes_per_unit']:6d} B/unit  "
              f"{r['achieved']:8.1f} GB/s  = {r['frac']:5.3f} of the measured HBM copy rate ({r['peak']:.0f} GB/s)")
