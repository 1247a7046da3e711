"""Host logic of the upsample folding (expertsim/_nets.py: FoldedConv) checked WITHOUT a GPU: the tap tables, fold masks and
scatter geometry handed to es_igemm_taps_fwd / es_fold_*_weights are interpreted by a few lines of numpy that restate the
documented semantics of include/expertsim_b200.h (es_tap_geom, es_fold_table), and the result must equal
conv2d(interpolate(x, nearest)) of torch — forward — and its autograd data gradient, in float64.  The CUDA kernels are
tested against the same torch expressions in tests/test_kernels_gpu.py; this file pins the part that is pure table
construction (x2 folding of proton conv1 / neutron conv0+conv5, y-only folding of proton conv2 incl. its folded data
gradient on the [Hs, Wu] grid)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from expertsim._nets import FoldedConv


def nearest_map(S, U):
    """torch 'nearest': src = min(floor(dst * float32(S / U)), S - 1) — what fill_maps_uc() gives the kernels"""
    m = np.floor(np.arange(U, dtype=np.float32) * np.float32(S / U)).astype(np.int64)
    return np.minimum(m, S - 1)


def fold(w, table, dgrad):
    """es_fold_up2_weights / es_fold_weights_multi: w [N,C,KH,KW] -> [N, T*C] (forward) or [C, T*N] (data gradient)"""
    N, C, KH, KW = w.shape
    flat = w.reshape(N, C, KH * KW)
    cols = []
    for t in range(table.n_taps):
        bits = [k for k in range(KH * KW) if (table.mask[t] >> k) & 1]
        cols.append(flat[:, :, bits].sum(axis=2))              # [N, C]
    wt = np.stack(cols, axis=1)                                  # [N, T, C]
    return wt.transpose(2, 1, 0).reshape(C, -1) if dgrad else wt.reshape(N, -1)


def taps_conv(x, wk, g, out):
    """es_igemm_taps_fwd on one sample: x [Hs,Ws,C]; wk [Nout, KK]; scatters into out [Ho_full, Wo_full, Nout]"""
    ymap, xmap = nearest_map(g.Hs, g.Hu), nearest_map(g.Ws, g.Wu)
    a, b = np.meshgrid(np.arange(g.Ho), np.arange(g.Wo), indexing="ij")
    acc = np.zeros((g.Ho, g.Wo, wk.shape[0]))
    for t in range(g.n_taps):
        uy, ux = a * g.my + g.tap_dy[t], b * g.mx + g.tap_dx[t]
        ok = (uy >= 0) & (uy < g.Hu) & (ux >= 0) & (ux < g.Wu)
        src = x[ymap[np.clip(uy, 0, g.Hu - 1)], xmap[np.clip(ux, 0, g.Wu - 1)]] * ok[..., None]     # [Ho, Wo, C]
        acc += src @ wk[:, g.tap_koff[t]:g.tap_koff[t] + g.C].T
    assert ((g.Ho - 1) * g.o_my + g.o_oy < g.Ho_full) and ((g.Wo - 1) * g.o_mx + g.o_ox < g.Wo_full)
    out[g.o_oy::g.o_my, g.o_ox::g.o_mx][:g.Ho, :g.Wo] = acc


CASES = {  # Hs, Ws, C, Hu, Wu, K, pad, N, fold
    "proton conv1 (x2, k4 p1)": (6, 5, 128, 12, 10, 4, 1, 128, (True, True)),
    "neutron conv (x2, k3 p0)": (7, 6, 128, 14, 12, 3, 0, 128, (True, True)),
    "proton conv2 (5->8 rows, y only)": (10, 7, 128, 16, 11, 4, 1, 128, (True, False)),
    "proton conv2 full size": (35, 19, 128, 56, 30, 4, 1, 128, (True, False)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_folded_tables_reproduce_conv_of_upsampled_input(name):
    Hs, Ws, C, Hu, Wu, K, pad, N, fd = CASES[name]
    rng = np.random.default_rng(len(name))
    x = rng.normal(size=(Hs, Ws, C))
    w = rng.normal(size=(N, C, K, K)) / np.sqrt(C * K * K)
    f = FoldedConv(Hs, Ws, C, Hu, Wu, K, K, pad, N, fd)
    # ---- torch reference (float64)
    xt = torch.from_numpy(x).permute(2, 0, 1)[None]
    grid = (Hs, Ws) if f.dg_grid == (Hs, Ws) else f.dg_grid                    # where the folded data gradient lives
    xg = F.interpolate(xt, size=grid, mode="nearest").requires_grad_(True)     # identity for x2 folding
    up = F.interpolate(xg, size=(Hu, Wu), mode="nearest")
    assert torch.equal(up, F.interpolate(xt, size=(Hu, Wu), mode="nearest"))   # nearest upsampling is separable
    y_ref = F.conv2d(up, torch.from_numpy(w), padding=pad)
    dy = torch.from_numpy(rng.normal(size=tuple(y_ref.shape)))
    (y_ref * dy).sum().backward()
    # ---- forward: one table-conv per output class, together they tile the output exactly once
    y = np.full((f.Ho, f.Wo, N), np.nan)
    for c in f.classes:
        taps_conv(x, fold(w, c["table"], False), c["g_fwd"], y)
    assert not np.isnan(y).any(), "classes must cover every output pixel"
    np.testing.assert_allclose(y, y_ref[0].detach().permute(1, 2, 0).numpy(), rtol=0, atol=1e-10)
    executed = sum(c["Hp"] * c["Wp"] * len(c["taps"]) for c in f.classes)
    assert executed < f.Ho * f.Wo * K * K and abs(f.executed_ratio - executed / (f.Ho * f.Wo * K * K)) < 1e-12
    # ---- data gradient on the folded grid
    assert f.has_dgrad
    dyn = dy[0].permute(1, 2, 0).numpy()
    dx = np.full((grid[0], grid[1], C), np.nan)
    if f.dgrad_classes:
        for c in f.dgrad_classes:
            taps_conv(dyn, fold(w, c["table"], True), c["g"], dx)
        assert sum(c["g"].alg_flops_per_row for c in f.dgrad_classes) == 2.0 * Hu * Wu * N * K * K * C
    else:
        taps_conv(dyn, fold(w, f.table_all, True), f.g_dgrad, dx)
    assert not np.isnan(dx).any()
    np.testing.assert_allclose(dx, xg.grad[0].permute(1, 2, 0).numpy(), rtol=0, atol=1e-10)
