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
from expertsim import _lib as L  # noqa: E402

dev = "cuda"
E, R, HW = 8, 65536, 56 * 30
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


cases = {
    "es_gen_loss_reduce": (lambda: L.call("es_gen_loss_reduce", *args, score, grp, E, R, s_out, div_out, sums), HW * 4 + 596 + 24),
    "es_gen_loss_grads": (lambda: L.call("es_gen_loss_grads", *args, s_out, div_out, grp, E, R, sums, R, 0.1, 1e-3, 1e-3, d_s, d_l1, d_l2,
                                         d_c, d_img, losses), HW * 4 * 3 + 596 + 24 + 536),
    "es_expm1_scatter": (lambda: L.call("es_expm1_scatter", img, perm, R, HW, None, out32), HW * 4 * 2),
    "es_gather_rows (images)": (lambda: L.call("es_gather_rows", img, perm, R, HW, gat), HW * 4 * 2),
    "es_hinge_d": (lambda: L.call("es_hinge_d", s_out, div_out, grp, E, None, R, d_s, lat1[:, 0].contiguous(), losses[:, 0].contiguous()), 16),
    "es_router_fwd": (lambda: L.call("es_router_fwd", cond, R, E, *W, gumbel, 1.2, ro["logits"], ro["gates"], ro["idx"], ro["h1"], ro["h2"],
                                     ro["h3"], ro["hist"]), 36 + 4 * E + (128 + 64 + 32 + 2 * E) * 4 + 8),
}
for name, (fn, bytes_per_row) in cases.items():
    t = timeit(fn)
    gbs = bytes_per_row * R / t / 1e9
    print(f"{name:26s} rows={R} {t * 1e6:9.1f} us  {bytes_per_row:6d} B/row  {gbs:8.1f} GB/s  = {gbs / peak:5.3f} of the measured HBM copy rate ({peak:.0f} GB/s)")
