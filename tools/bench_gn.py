#!/usr/bin/env python
"""Times es_gn_lrelu_fwd / es_gn_lrelu_bwd on the three proton generator layers (2048 rows, 8 experts) with CUDA events.
Environment: ES_GN_LEGACY=1 (one CTA per sample) / ES_GN_SLAB_KB=<n> (cluster slab limit)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200"))
from expertsim import _lib as L  # noqa: E402

dev = "cuda"
R, E = 2048, 8
grp = torch.tensor([[i * 256, 256, i, 128] for i in range(E)], dtype=torch.int32, device=dev)
BF = torch.bfloat16


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
    return e0.elapsed_time(e1) / n * 1e3


tag = " ".join(f"{k[6:].lower()}={os.environ[k]}" for k in ("ES_GN_LEGACY", "ES_GN_SLAB_KB", "ES_GN_CARVEOUT", "ES_GN_CLUSTER_BWD") if k in os.environ) or "default"
for name, (Hs, Ws, C, Hu, Wu) in {"conv1": (35, 19, 256, 35, 19), "conv1_fanx": (35, 19, 256, 35, 38), "conv2": (55, 29, 128, 55, 29),
                                  "conv3": (55, 29, 64, 55, 29)}.items():
    P = Hs * Ws
    x = torch.randn(R, P, C, device=dev).to(BF)
    y = torch.empty_like(x)
    st = torch.empty(R, 32, 2, device=dev)
    gamma, beta = torch.ones(E, C, device=dev), torch.zeros(E, C, device=dev)
    dg, db, dbias = torch.zeros(E, C, device=dev), torch.zeros(E, C, device=dev), torch.zeros(E, C, device=dev)
    da = torch.randn(R, Hu * Wu, C, device=dev).to(BF)
    dx = torch.empty_like(x)
    tf = timeit(lambda: L.call("es_gn_lrelu_fwd", x, gamma, beta, C, P, C, 32, grp, E, R, y, st))
    tb = timeit(lambda: L.call("es_gn_lrelu_bwd", da, Hs, Ws, Hu, Wu, x, st, gamma, beta, C, C, 32, grp, E, R, dx, dg, db, dbias))
    if Wu > Ws:      # forward writing the x-upsampled layout (what conv2 reads)
        yu = torch.empty(R, Hs * Wu, C, device=dev, dtype=BF)
        tu = timeit(lambda: L.call("es_gn_lrelu_fwd_upx", x, gamma, beta, C, Hs, Ws, Wu, C, 32, grp, E, R, yu, st))
        gb_u = (x.numel() + yu.numel()) * 2 / 1e9
        print(f"[{tag}] {name:12s} fwd_upx {tu:7.1f} us ({gb_u / tu * 1e6:6.0f} GB/s)")
    ps = torch.rand(R, C // 2, 2, device=dev) * 100
    ps[..., 1] += 1e6
    ta = timeit(lambda: L.call("es_gn_lrelu_apply_fwd", x, ps, gamma, beta, C, Hs, Ws, Ws, C, 32, grp, E, R, y, st))
    print(f"[{tag}] {name:12s} apply_fwd {ta:7.1f} us ({2 * x.numel() * 2 / 1e9 / ta * 1e6:6.0f} GB/s)")
    if Wu > Ws:
        tau = timeit(lambda: L.call("es_gn_lrelu_apply_fwd", x, ps, gamma, beta, C, Hs, Ws, Wu, C, 32, grp, E, R, yu, st))
        print(f"[{tag}] {name:12s} apply_fwd_upx {tau:7.1f} us ({gb_u / tau * 1e6:6.0f} GB/s)")
    gb_f = 2 * x.numel() * 2 / 1e9
    gb_b = (2 * x.numel() + da.numel()) * 2 / 1e9
    print(f"[{tag}] {name:12s} fwd {tf:7.1f} us ({gb_f / tf * 1e6:6.0f} GB/s)   bwd {tb:7.1f} us ({gb_b / tb * 1e6:6.0f} GB/s)")
