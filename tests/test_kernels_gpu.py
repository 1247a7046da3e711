"""Per-kernel parity tests: every C-ABI entry point against the CPU oracle / plain fp32 torch on the same inputs.

Tolerances: integer outputs (expert index, counts, permutation) are compared bit-exactly; fp32 kernels to 1e-4..1e-5
relative L2; bf16 tensor-core kernels to 1e-2 relative L2 (bf16 has 8 mantissa bits; inputs are pre-rounded to bf16 so
only accumulation order and the output rounding differ)."""
import math

import pytest
import torch
import torch.nn.functional as F

import oracle.expertsim_oracle as orc
from gpu_util import DEV, L, bf16_round, check, cuda, groups, log

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def G(seed):
    return torch.Generator().manual_seed(seed)


# ----------------------------------------------------------------------------------------------------------- K1
def run_router(B, E, seed, tau=1.2, min_rows=2):
    sd = orc.make_weights("proton", "router", seed, n_experts=E)
    g = G(seed)
    cond = torch.randn(B, 9, generator=g)
    gumbel = -torch.empty(B, E).exponential_(generator=g).log()
    d = {k: cuda(v) for k, v in sd.items()}
    nblk = (B + 255) // 256
    out = dict(logits=torch.empty(B, E, device=DEV), gates=torch.empty(B, E, device=DEV),
               idx=torch.empty(B, dtype=torch.int64, device=DEV), h1=torch.empty(B, 128, device=DEV),
               h2=torch.empty(B, 64, device=DEV), h3=torch.empty(B, 32, device=DEV),
               hist=torch.empty(nblk, E, dtype=torch.int32, device=DEV))
    L.call("es_router_fwd", cuda(cond), B, E, d["fc_layers.0.weight"], d["fc_layers.0.bias"], d["fc_layers.2.weight"],
           d["fc_layers.2.bias"], d["fc_layers.4.weight"], d["fc_layers.4.bias"], d["fc_layers.6.weight"],
           d["fc_layers.6.bias"], cuda(gumbel), tau, out["logits"], out["gates"], out["idx"], out["h1"], out["h2"],
           out["h3"], out["hist"])
    counts = torch.empty(E, dtype=torch.int32, device=DEV)
    offsets = torch.empty(E + 1, dtype=torch.int32, device=DEV)
    perm = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    gh = torch.empty(E, 4, dtype=torch.int32, device=DEV)
    gg = torch.empty(E, 4, dtype=torch.int32, device=DEV)
    scratch = torch.empty(nblk * E + E, dtype=torch.int32, device=DEV)
    L.call("es_router_partition", out["idx"], B, E, min_rows, out["hist"], counts, offsets, perm, gh, gg, scratch)
    torch.cuda.synchronize()
    return sd, cond, gumbel, out, counts, offsets, perm, gh, gg


@pytest.mark.parametrize("B,E", [(1, 1), (37, 3), (256, 3), (257, 8), (1000, 8), (5000, 5), (20000, 16)])
def test_router_fwd_partition(B, E):
    sd, cond, gumbel, out, counts, offsets, perm, gh, gg = run_router(B, E, seed=B + E)
    gates, logits = orc.router_forward(sd, cond, gumbel, 1.2)
    idx, cnt, masks = orc.route(gates, E)
    check(f"router logits B={B} E={E}", out["logits"], logits, 1e-5)
    check(f"router gates  B={B} E={E}", out["gates"], gates, 1e-5)
    assert out["idx"].cpu().tolist() == idx.tolist(), "expert assignment must be bit-exact"
    assert counts.cpu().tolist() == cnt.tolist()
    assert perm.cpu().tolist() == torch.cat(masks).tolist(), "stable token->expert permutation must be bit-exact"
    off = [0]
    for c in cnt.tolist():
        off.append(off[-1] + c)
    assert offsets.cpu().tolist() == off
    for e in range(E):
        act = cnt[e].item() if cnt[e].item() >= 2 else 0
        assert gh[e].cpu().tolist() == [off[e], act, e, act]
        assert gg[e].cpu().tolist() == [2 * off[e], 2 * act, e, act]


def test_gather_scatter_rows():
    B, W = 333, 1680
    g = G(1)
    x = torch.randn(B, W, generator=g)
    perm = torch.randperm(B, generator=g).to(torch.int32)
    out = torch.empty(B, W, device=DEV)
    L.call("es_gather_rows", cuda(x), cuda(perm), B, W, out)
    assert torch.equal(out.cpu(), x[perm.long()])
    back = torch.empty(B, W, device=DEV)
    L.call("es_scatter_rows", out, cuda(perm), B, W, back)
    assert torch.equal(back.cpu(), x)
    x9 = torch.randn(B, 9, generator=g)
    o9 = torch.empty(B, 9, device=DEV)
    L.call("es_gather_rows", cuda(x9), cuda(perm), B, 9, o9)
    assert torch.equal(o9.cpu(), x9[perm.long()])


@pytest.mark.parametrize("B,E,util", [(64, 3, 0.0), (500, 8, 0.1)])
def test_router_bwd(B, E, util):
    sd, cond, gumbel, out, *_ = run_router(B, E, seed=11 + B)
    tau, alb_s, alb_w = 1.2, 1e-5 if util == 0 else 1e-2, 0.2
    live = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    gates, _ = orc.router_forward(live, cond, gumbel, tau)
    alb = orc.adaptive_load_balancing_loss(gates.sum(0), alb_s)
    ent = -1 * orc.utilization_entropy(gates, util) if util else torch.tensor(0.0)
    extra = torch.randn(B, E, generator=G(5)) * 1e-3
    loss = alb_w * alb + ent + (gates * extra).sum()
    loss.backward()
    names = ["fc_layers.0", "fc_layers.2", "fc_layers.4", "fc_layers.6"]
    d = {k: cuda(v) for k, v in sd.items()}
    gr = {k: torch.zeros_like(v) for k, v in d.items()}
    sums = torch.empty(E, device=DEV)
    L.call("es_router_gate_sums", out["gates"], B, E, sums)
    losses = torch.zeros(2, device=DEV)
    L.call("es_router_bwd", cuda(cond), B, E, B, d["fc_layers.2.weight"], d["fc_layers.4.weight"], d["fc_layers.6.weight"],
           out["gates"], out["h1"], out["h2"], out["h3"], sums, tau, alb_s, alb_w, util, cuda(extra),
           *[gr[f"{n}.{p}"] for n in names for p in ("weight", "bias")], losses)
    check("router gate sums", sums, gates.detach().sum(0), 1e-5)
    for n in names:
        for p in ("weight", "bias"):
            check(f"router grad {n}.{p} B={B}", gr[f"{n}.{p}"], live[f"{n}.{p}"].grad, 2e-4)
    assert abs(losses[0].item() - float(alb)) <= 1e-4 * abs(float(alb))
    assert abs(losses[1].item() - float(ent)) <= 1e-4 * max(abs(float(ent)), 1e-6)


def test_router_ed_loss():
    B, E = 200, 4
    g = G(3)
    idx = torch.randint(0, E, (B,), generator=g)
    m = torch.rand(B, 1, generator=g) * 50
    gates_soft = torch.rand(B, E, generator=g).softmax(1).requires_grad_(True)
    gates = F.one_hot(idx, E).float() + (gates_soft - gates_soft.detach())
    ref = orc.expert_distribution_loss(gates, m) * 0.01
    ref.backward()
    dg = torch.empty(B, E, device=DEV)
    lo = torch.zeros(1, device=DEV)
    L.call("es_router_ed_loss", cuda(idx), cuda(m.flatten()), B, E, 0.01, dg, lo)
    check("expert-distribution loss grad", dg, gates_soft.grad, 1e-4)
    assert abs(lo.item() - float(ref)) <= 1e-4 * abs(float(ref))


# ----------------------------------------------------------------------------------------------------------- K4
def test_hinge_d():
    counts = [5, 0, 9, 1]
    grp, R = groups(counts, min_rows=2)
    g = G(2)
    real = torch.randn(R, generator=g) * 2
    fake = torch.randn(R, generator=g) * 2
    Bg = 40
    rr, ff = real.clone().requires_grad_(True), fake.clone().requires_grad_(True)
    exp_loss, off = [], 0
    for c in counts:
        if c >= 2:
            l = (F.relu(1 - rr[off:off + c]).mean() + F.relu(1 + ff[off:off + c]).mean()) * (c / Bg)
            l.backward()
            exp_loss.append(float(l))
        else:
            exp_loss.append(0.0)
        off += c
    dr, df = torch.zeros(R, device=DEV), torch.zeros(R, device=DEV)
    loss = torch.zeros(len(counts), device=DEV)
    L.call("es_hinge_d", cuda(real), cuda(fake), grp, len(counts), None, Bg, dr, df, loss)
    check("hinge D loss", loss, torch.tensor(exp_loss), 1e-5)
    check("hinge d_real", dr, rr.grad, 1e-6)
    check("hinge d_fake", df, ff.grad, 1e-6)


@pytest.mark.parametrize("arch", ["proton", "neutron"])
def test_gen_loss_tails(arch):
    H, W = orc.IMAGE_SHAPE[arch]
    HW = H * W
    counts = [6, 1, 11]
    grp, R = groups(counts, min_rows=2)
    g = G(7)
    img = (torch.rand(R, 1, H, W, generator=g) < 0.05).float() * torch.rand(R, 1, H, W, generator=g) * 5
    lat1, lat2 = torch.randn(R, 64, generator=g), torch.randn(R, 64, generator=g)
    z1, z2 = torch.randn(R, 10, generator=g), torch.randn(R, 10, generator=g)
    std, inten = torch.rand(R, 1, generator=g), torch.rand(R, 1, generator=g) * 100
    coords, pos = torch.randn(R, 2, generator=g) * 10, torch.rand(R, 2, generator=g) * 30
    score = torch.randn(R, 1, generator=g)
    Bg, di, ins, aux = 32, 0.1, 1e-3, 1e-3
    leaves = [t.clone().requires_grad_(True) for t in (img, lat1, lat2, coords, score)]
    li, l1, l2, lc, ls = leaves
    exp, off = [], 0
    for c in counts:
        if c >= 2:
            s = slice(off, off + c)
            gl = -ls[s].mean()
            dv = orc.sdi_gan_regularization(l1[s], l2[s], z1[s], z2[s], std[s], di)
            il, sums, sstd, smean = orc.intensity_regularization(li[s], inten[s], ins)
            al = orc.regressor_loss(pos[s], lc[s]) * aux
            tot = (gl + dv + il + al) * (c / Bg)
            tot.backward()
            exp.append([float(tot), float(dv), float(il), float(al), float(sstd), float(smean)])
        else:
            exp.append([0.0] * 6)
        off += c
    d_img = torch.zeros(R, HW, device=DEV)
    s_out, div_out = torch.zeros(R, device=DEV), torch.zeros(R, device=DEV)
    sums = torch.zeros(len(counts), 8, dtype=torch.float64, device=DEV)
    c_ = lambda t: cuda(t.reshape(R, -1))
    args = (c_(img), HW, c_(lat1), c_(lat2), c_(z1), c_(z2), c_(std), c_(inten), c_(coords), c_(pos))
    L.call("es_gen_loss_reduce", *args, c_(score), grp, len(counts), R, s_out, div_out, sums)
    d_s, d_l1, d_l2 = torch.zeros(R, device=DEV), torch.zeros(R, 64, device=DEV), torch.zeros(R, 64, device=DEV)
    d_c = torch.zeros(R, 2, device=DEV)
    losses = torch.zeros(len(counts), 6, device=DEV)
    L.call("es_gen_loss_grads", *args, s_out, div_out, grp, len(counts), R, sums, Bg, di, ins, aux, d_s, d_l1, d_l2, d_c, d_img, losses)
    check(f"{arch} gen losses", losses, torch.tensor(exp), 2e-5)
    check(f"{arch} d_score", d_s, ls.grad.flatten(), 1e-6)
    check(f"{arch} d_lat1", d_l1, l1.grad, 2e-5)
    check(f"{arch} d_lat2", d_l2, l2.grad, 2e-5)
    check(f"{arch} d_coords", d_c, lc.grad, 2e-5)
    check(f"{arch} d_img (intensity)", d_img, li.grad.reshape(R, HW), 2e-5)


# ----------------------------------------------------------------------------------------------------------- K2
def conv_geom(Hs, Ws, C, Hu, Wu, KH, KW, pad, N):
    return L.ESConvGeom(Hs, Ws, C, Hu, Wu, Hu + 2 * pad - KH + 1, Wu + 2 * pad - KW + 1, KH, KW, pad, N)


GEOMS = {
    "conv1_fwd": (18, 10, 512, 36, 20, 4, 4, 1, 256),
    "conv2_fwd": (35, 19, 256, 56, 30, 4, 4, 1, 128),
    "conv3_fwd": (55, 29, 128, 55, 29, 3, 3, 1, 64),
    "conv3_dgrad": (55, 29, 64, 55, 29, 3, 3, 1, 128),
    "conv2_dgrad": (55, 29, 128, 55, 29, 4, 4, 2, 256),
    "conv1_dgrad": (35, 19, 256, 35, 19, 4, 4, 2, 512),
    "neutron_conv1": (13, 13, 128, 26, 26, 3, 3, 0, 256),
    "dense_small": (1, 1, 256, 1, 1, 1, 1, 0, 1024),
}


def ref_conv(x, w, bias, geo, counts, slots):
    """fp32 reference of the grouped implicit GEMM on bf16-rounded operands; x [R,Hs,Ws,C], w [S,N,KH,KW,C]."""
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    out, off = [], 0
    for c, s in zip(counts, slots):
        if c == 0:
            continue
        xi = x[off:off + c].permute(0, 3, 1, 2)
        if (Hu, Wu) != (Hs, Ws):
            xi = F.interpolate(xi, size=(Hu, Wu), mode="nearest")
        y = F.conv2d(xi, w[s].permute(0, 3, 1, 2), bias[s] if bias is not None else None, padding=pad)
        out.append(y.permute(0, 2, 3, 1))
        off += c
    return torch.cat(out)


@pytest.mark.parametrize("name", list(GEOMS))
@pytest.mark.parametrize("impl", ["es_igemm_fwd", "es_igemm_fwd_simt"])
def test_igemm_fwd(name, impl):
    geo = GEOMS[name]
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    counts, slots = [3, 2], [1, 0]
    if name == "dense_small":
        counts, slots = [130, 77], [1, 0]
    grp, R = groups(counts, slots)
    g = G(sum(map(ord, name)))
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    w = bf16_round(torch.randn(2, N, KH, KW, C, generator=g) / math.sqrt(KH * KW * C))
    bias = torch.randn(2, N, generator=g) * 0.1
    want = ref_conv(x, w, bias, geo, counts, slots)
    y = torch.zeros(R, want.shape[1], want.shape[2], N, dtype=BF, device=DEV)
    L.call(impl, cuda(x, BF), cuda(w, BF), cuda(bias), N, y, conv_geom(*geo), grp, len(counts), R)
    torch.cuda.synchronize()
    check(f"{impl} {name}", y.float(), want, 6e-3, 3e-2)


def test_igemm_fwd_fc2_fullsize():
    """fc2: [rows,256] x [92160,256]^T with a ragged two-expert batch and an inactive group in between."""
    counts, slots = [150, 0, 45], [0, 1, 2]
    grp, R = groups(counts, slots)
    g = G(42)
    N, K = 92160, 256
    x = bf16_round(torch.randn(R, K, generator=g))
    w = bf16_round(torch.randn(3, N, K, generator=g) / 16)
    bias = torch.randn(3, N, generator=g) * 0.1
    y = torch.zeros(R, N, dtype=BF, device=DEV)
    L.call("es_igemm_fwd", cuda(x, BF), cuda(w, BF), cuda(bias), N, y, conv_geom(1, 1, K, 1, 1, 1, 1, 0, N), grp, 3, R)
    want = torch.cat([x[:150] @ w[0].T + bias[0], x[150:] @ w[2].T + bias[2]])
    check("es_igemm_fwd fc2 92160x256", y.float(), want, 6e-3, 3e-2)


RAGGED8 = ([9, 0, 13, 7, 1, 11, 15, 8], [5, 1, 0, 7, 2, 6, 3, 4])      # 64 rows in 8 ragged expert groups (one empty, one of 1 row)


@pytest.mark.parametrize("name", ["conv3_fwd", "conv3_dgrad", "conv2_dgrad", "conv2_fwd", "conv1_dgrad"])
def test_igemm_fwd_many_tiles(name):
    """The persistent path proper: 64 samples in 8 ragged expert groups = 280-800 output tiles on 148 CTAs, i.e. several tiles
    per CTA (TMEM double-buffer wrap, pipeline phase wrap across tiles, expert changes inside a CTA's tile sequence), against
    fp32 torch convolutions ON THE DEVICE (TF32 off) over the same bf16-rounded operands.  Bound: 6e-3 rel. L2."""
    geo = GEOMS[name]
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    counts, slots = RAGGED8
    grp, R = groups(counts, slots)
    g = G(11 + sum(map(ord, name)))
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    w = bf16_round(torch.randn(8, N, KH, KW, C, generator=g) / math.sqrt(KH * KW * C))
    bias = torch.randn(8, N, generator=g) * 0.1
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = ref_conv(x.to(DEV), w.to(DEV), bias.to(DEV), geo, counts, slots)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    tiles = sum((c * want.shape[1] * want.shape[2] + 127) // 128 for c in counts)
    assert tiles >= 280
    y = torch.zeros(R, want.shape[1], want.shape[2], N, dtype=BF, device=DEV)
    L.call("es_igemm_fwd", cuda(x, BF), cuda(w, BF), cuda(bias), N, y, conv_geom(*geo), grp, len(counts), R)
    torch.cuda.synchronize()
    check(f"es_igemm_fwd {name} x64 rows, 8 ragged groups, {tiles} tiles", y.float(), want, 6e-3, 3e-2)


@pytest.mark.parametrize("name", ["conv3_fwd", "conv2_fwd", "conv2_dgrad"])
def test_igemm_wgrad_many_tiles(name):
    """split-K weight gradient over 64 samples in 8 ragged groups (many K-chunks per CTA, RED epilogue), vs torch autograd on the
    device; run twice: the fp32 atomics make the result order-dependent, the run-to-run difference is bounded at 1e-5."""
    geo = GEOMS[name]
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    counts, slots = RAGGED8
    grp, R = groups(counts, slots)
    g = G(13 + sum(map(ord, name)))
    Ho, Wo = Hu + 2 * pad - KH + 1, Wu + 2 * pad - KW + 1
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    dy = bf16_round(torch.randn(R, Ho, Wo, N, generator=g))
    w = torch.zeros(8, N, KH, KW, C, requires_grad=True, device=DEV)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref_conv(x.to(DEV), w, None, geo, counts, slots).backward(dy.to(DEV))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    runs = []
    for _ in range(2):
        dw = torch.zeros(8, N, KH, KW, C, device=DEV)
        L.call("es_igemm_wgrad", cuda(x, BF), cuda(dy, BF), dw, conv_geom(*geo), grp, len(counts), R)
        torch.cuda.synchronize()
        runs.append(dw)
    check(f"es_igemm_wgrad {name} x64 rows, 8 ragged groups", runs[0], w.grad, 2e-3, 1e-2)
    check(f"es_igemm_wgrad {name} run-to-run (atomic order)", runs[1], runs[0], 1e-5)
    assert float(runs[0][1].abs().max()) == 0.0          # slot 1 belongs to the empty group


@pytest.mark.parametrize("name", ["conv1_fwd", "conv2_fwd", "conv3_fwd", "neutron_conv1", "conv2_dgrad"])
@pytest.mark.parametrize("impl", ["es_igemm_wgrad", "es_igemm_wgrad_simt"])
def test_igemm_wgrad(name, impl):
    geo = GEOMS[name]
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    counts, slots = [2, 3], [1, 0]
    grp, R = groups(counts, slots)
    g = G(7 + sum(map(ord, name)))
    Ho, Wo = Hu + 2 * pad - KH + 1, Wu + 2 * pad - KW + 1
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    dy = bf16_round(torch.randn(R, Ho, Wo, N, generator=g))
    w = torch.zeros(2, N, KH, KW, C, requires_grad=True)
    out = ref_conv(x, w, None, geo, counts, slots)
    out.backward(dy)
    dw = torch.zeros(2, N, KH, KW, C, device=DEV)
    L.call(impl, cuda(x, BF), cuda(dy, BF), dw, conv_geom(*geo), grp, len(counts), R)
    torch.cuda.synchronize()
    check(f"{impl} {name}", dw, w.grad, 2e-3, 1e-2)


def test_dense_dgrad_wgrad():
    counts, slots = [140, 0, 60], [2, 1, 0]
    grp, R = groups(counts, slots)
    g = G(9)
    N, K = 92160, 256
    dy = bf16_round(torch.randn(R, N, generator=g) * 0.1)
    w = bf16_round(torch.randn(3, N, K, generator=g) / 16)
    x = bf16_round(torch.randn(R, K, generator=g))
    dx = torch.zeros(R, K, device=DEV)
    L.call("es_dense_dgrad", cuda(dy, BF), cuda(w, BF), dx, N, K, grp, 3, R)
    want = torch.cat([dy[:140] @ w[2], dy[140:] @ w[0]])
    check("es_dense_dgrad fc2", dx, want, 3e-3, 1e-2)
    row_map = torch.randperm(N, generator=g).to(torch.int32)
    dw = torch.zeros(3, N, K, device=DEV)
    xpad = torch.full((3 * ((R + 63) // 64) * 64, K), float("nan"), dtype=BF, device=DEV)      # scratch may hold anything
    L.call("es_dense_wgrad", cuda(dy, BF), cuda(x, BF), dw, N * K, N, K, cuda(row_map), grp, 3, R, xpad)
    want_w = torch.zeros(3, N, K)
    want_w[2][row_map.long()] = dy[:140].T @ x[:140]
    want_w[0][row_map.long()] = dy[140:].T @ x[140:]
    check("es_dense_wgrad fc2 (slot 2)", dw[2], want_w[2], 3e-3, 1e-2)
    check("es_dense_wgrad fc2 (slot 0)", dw[0], want_w[0], 3e-3, 1e-2)
    assert float(dw[1].abs().max()) == 0.0


def test_pack_unpack():
    g = G(4)
    S, N, C, KH, KW = 2, 16, 64, 3, 2
    w = torch.randn(S, N, C, KH, KW, generator=g)
    wf = torch.zeros(S, N, KH, KW, C, dtype=BF, device=DEV)
    wd = torch.zeros(S, C, KH, KW, N, dtype=BF, device=DEV)
    L.call("es_pack_conv_weight", cuda(w), N * C * KH * KW, S, N, C, KH, KW, wf, wd)
    assert torch.equal(wf.float().cpu(), bf16_round(w.permute(0, 1, 3, 4, 2)))
    assert torch.equal(wd.float().cpu(), bf16_round(w.flip(3, 4).permute(0, 2, 3, 4, 1)))
    dwp = torch.randn(S, N, KH, KW, C, generator=g)
    dwr = torch.zeros(S, N, C, KH, KW, device=DEV)
    L.call("es_unpack_conv_wgrad", cuda(dwp), S, N, C, KH, KW, dwr, N * C * KH * KW)
    assert torch.equal(dwr.cpu(), dwp.permute(0, 1, 4, 2, 3))
    Nd, K = 40, 24
    wd2 = torch.randn(S, Nd, K, generator=g)
    rm = torch.randperm(Nd, generator=g).to(torch.int32)
    wp = torch.zeros(S, Nd, K, dtype=BF, device=DEV)
    L.call("es_pack_dense_weight", cuda(wd2), Nd * K, S, Nd, K, cuda(rm), wp)
    assert torch.equal(wp.float().cpu(), bf16_round(wd2[:, rm.long()]))
    v = torch.randn(S, Nd, generator=g)
    o = torch.zeros(S, Nd, device=DEV)
    L.call("es_permute_features", cuda(v), Nd, cuda(rm), S, Nd, o, Nd, 0)
    assert torch.equal(o.cpu(), v[:, rm.long()])
    o2 = torch.zeros(S, Nd, device=DEV)
    L.call("es_permute_features", o, Nd, cuda(rm), S, Nd, o2, Nd, 1)
    assert torch.equal(o2.cpu(), v)


def test_gen_fc1_fwd_bwd():
    counts = [5, 0, 7]
    E = 3
    gh, B = groups(counts)
    gg, _ = groups(counts, two_pass=True)
    R = 2 * B
    g = G(12)
    z1, z2, cond = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g), torch.randn(B, 9, generator=g)
    W = torch.randn(E, 256, 19, generator=g) / 4
    b, ga, be = torch.randn(E, 256, generator=g) * .1, 1 + .1 * torch.randn(E, 256, generator=g), .1 * torch.randn(E, 256, generator=g)
    x0 = torch.zeros(R, 19, device=DEV)
    lin = torch.zeros(R, 256, device=DEV)
    h = torch.zeros(R, 256, dtype=BF, device=DEV)
    L.call("es_gen_fc1_fwd", cuda(z1), cuda(z2), cuda(cond), cuda(W), cuda(b), cuda(ga), cuda(be), 256 * 19, 256, gg, E, R, 1, x0, lin, h)
    leaves = [t.clone().requires_grad_(True) for t in (W, b, ga, be)]
    lW, lb, lg, lbe = leaves
    rows, off = [], 0
    dh = torch.randn(R, 256, generator=g)
    want_h = torch.zeros(R, 256)
    for e, c in enumerate(counts):
        for p, z in enumerate((z1, z2)):
            xin = torch.cat((z[off:off + c], cond[off:off + c]), 1)
            y = orc.lrelu(F.layer_norm(F.linear(xin, lW[e], lb[e]), (256,), lg[e], lbe[e], 1e-5))
            r0 = 2 * off + p * c
            want_h[r0:r0 + c] = y.detach()
            (y * dh[r0:r0 + c]).sum().backward()
        off += c
    check("gen fc1 fwd", h.float(), want_h, 4e-3)
    grads = [torch.zeros_like(cuda(t)) for t in (W, b, ga, be)]
    L.call("es_gen_fc1_bwd", cuda(dh), x0, lin, cuda(ga), cuda(be), 256 * 19, 256, gg, E, R, *grads)
    for nme, gt, lf in zip(("dW", "db", "dgamma", "dbeta"), grads, leaves):
        check(f"gen fc1 bwd {nme}", gt, lf.grad, 1e-4)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("P,C,groups_n,Hs,Ws,Hu,Wu", [(665, 256, 32, 35, 19, 56, 30), (665, 256, 32, 35, 19, 35, 30), (1595, 128, 32, 55, 29, 55, 29),
                                                    (1595, 64, 32, 55, 29, 55, 29)])
def test_gn_lrelu_fwd_bwd(P, C, groups_n, Hs, Ws, Hu, Wu):
    counts, slots = [2, 0, 3], [0, 1, 2]
    grp, R = groups(counts, slots)
    g = G(C)
    x = bf16_round(torch.randn(R, C, Hs, Ws, generator=g) * 2 + 0.5)
    ga, be = 1 + .2 * torch.randn(3, C, generator=g), .2 * torch.randn(3, C, generator=g)
    y = torch.zeros(R, P, C, dtype=BF, device=DEV)
    stats = torch.zeros(R, groups_n, 2, device=DEV)
    L.call("es_gn_lrelu_fwd", cuda(nhwc(x), BF), cuda(ga), cuda(be), C, P, C, groups_n, grp, 3, R, y, stats)
    lx = x.clone().requires_grad_(True)
    lg, lb = ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    dy_up = bf16_round(torch.randn(R, C, Hu, Wu, generator=g))
    outs, off = [], 0
    for c, s in zip(counts, slots):
        if c:
            o = orc.lrelu(F.group_norm(lx[off:off + c], groups_n, lg[s], lb[s], 1e-5))
            outs.append(o)
            up = F.interpolate(o, size=(Hu, Wu), mode="nearest") if (Hu, Wu) != (Hs, Ws) else o
            (up * dy_up[off:off + c]).sum().backward()
        off += c
    check(f"gn_lrelu fwd C={C}", y.float().reshape(R, Hs, Ws, C), nhwc(torch.cat(outs).detach()), 5e-3)
    dx = torch.zeros(R, P, C, dtype=BF, device=DEV)
    dga, dbe, dbias = torch.zeros(3, C, device=DEV), torch.zeros(3, C, device=DEV), torch.zeros(3, C, device=DEV)
    L.call("es_gn_lrelu_bwd", cuda(nhwc(dy_up), BF), Hs, Ws, Hu, Wu, cuda(nhwc(x), BF), stats, cuda(ga), cuda(be), C, C, groups_n,
           grp, 3, R, dx, dga, dbe, dbias)
    check(f"gn_lrelu bwd dx C={C}", dx.float().reshape(R, Hs, Ws, C), nhwc(lx.grad), 8e-3)
    check(f"gn_lrelu bwd dgamma C={C}", dga, lg.grad, 2e-3)
    check(f"gn_lrelu bwd dbeta C={C}", dbe, lb.grad, 2e-3)
    want_db = torch.zeros(3, C)
    off = 0
    for c, s in zip(counts, slots):
        want_db[s] = lx.grad[off:off + c].sum((0, 2, 3))
        off += c
    check(f"gn_lrelu bwd dbias C={C}", dbias, want_db, 2e-2)


def test_ln_lrelu_fwd_bwd():
    Hs, Ws, C, Hu, Wu = 18, 10, 512, 36, 20
    F_ = Hs * Ws * C
    counts, slots = [3, 2], [1, 0]
    grp, R = groups(counts, slots)
    g = G(21)
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g) * 1.5 + 0.3)  # packed (NHWC) feature order
    ga, be = 1 + .2 * torch.randn(2, F_, generator=g), .2 * torch.randn(2, F_, generator=g)
    y = torch.zeros(R, F_, dtype=BF, device=DEV)
    stats = torch.zeros(R, 2, device=DEV)
    L.call("es_ln_lrelu_fwd", cuda(x, BF), cuda(ga), cuda(be), F_, F_, grp, 2, R, y, stats)
    lx = x.clone().requires_grad_(True)
    lg, lb = ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    dy_up = bf16_round(torch.randn(R, Hu, Wu, C, generator=g))
    outs, off = [], 0
    for c, s in zip(counts, slots):
        o = orc.lrelu(F.layer_norm(lx[off:off + c].reshape(c, F_), (F_,), lg[s], lb[s], 1e-5))
        outs.append(o)
        up = F.interpolate(o.reshape(c, Hs, Ws, C).permute(0, 3, 1, 2), size=(Hu, Wu), mode="nearest")
        (up * dy_up[off:off + c].permute(0, 3, 1, 2)).sum().backward()
        off += c
    check("ln_lrelu fwd", y.float(), torch.cat(outs).detach(), 5e-3)
    dx = torch.zeros(R, F_, dtype=BF, device=DEV)
    L.call("es_ln_lrelu_bwd", cuda(dy_up, BF), Hs, Ws, Hu, Wu, C, cuda(x, BF), stats, cuda(ga), cuda(be), F_, grp, 2, R, dx)
    check("ln_lrelu bwd dx", dx.float(), lx.grad.reshape(R, F_), 8e-3)
    rm = torch.randperm(F_, generator=g).to(torch.int32)
    dga, dbe, dbl = torch.zeros(2, F_, device=DEV), torch.zeros(2, F_, device=DEV), torch.zeros(2, F_, device=DEV)
    L.call("es_ln_affine_bwd", cuda(dy_up, BF), Hs, Ws, Hu, Wu, C, cuda(x, BF), dx, stats, cuda(ga), cuda(be), F_, grp, 2, R,
           cuda(rm), F_, dga, dbe, dbl)
    inv = torch.empty(F_, dtype=torch.long)
    inv[rm.long()] = torch.arange(F_)
    check("ln affine dgamma", dga[:, rm.long()], lg.grad, 3e-3)
    check("ln affine dbeta", dbe[:, rm.long()], lb.grad, 3e-3)
    want = torch.stack([lx.grad[3:].reshape(2, F_).sum(0), lx.grad[:3].reshape(3, F_).sum(0)])
    check("ln affine dbias", dbl[:, rm.long()], want, 2e-2)


def test_gen_out_fwd_bwd():
    counts = [3, 0, 2]
    E = 3
    gg, B = groups(counts, two_pass=True)
    R = 2 * B
    Hs, Ws, C = 55, 29, 64
    g = G(31)
    x = bf16_round(torch.randn(R, C, Hs, Ws, generator=g))
    w = torch.randn(E, 1, C, 2, 2, generator=g) / 16
    b = torch.randn(E, generator=g) * .1
    img1, img2 = torch.zeros(B, 56 * 30, device=DEV), torch.zeros(B, 56 * 30, device=DEV)
    L.call("es_gen_out_fwd", cuda(nhwc(x), BF), cuda(w), cuda(b), C * 4, 1, Hs, Ws, C, 2, 2, 1, gg, E, R, 1, img1, img2)
    lx, lw, lb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    d1, d2 = torch.randn(B, 1, 56, 30, generator=g), torch.randn(B, 1, 56, 30, generator=g)
    w1, w2 = torch.zeros(B, 1, 56, 30), torch.zeros(B, 1, 56, 30)
    off = 0
    for e, c in enumerate(counts):
        for p, (dst, dd) in enumerate(((w1, d1), (w2, d2))):
            r0 = 2 * off + p * c
            o = F.relu(F.conv2d(lx[r0:r0 + c], lw[e], lb[e:e + 1], padding=1))
            dst[off:off + c] = o.detach()
            (o * dd[off:off + c]).sum().backward()
        off += c
    check("gen out conv fwd img1", img1, w1.reshape(B, -1), 1e-4)
    check("gen out conv fwd img2", img2, w2.reshape(B, -1), 1e-4)
    dx = torch.zeros(R, Hs * Ws, C, dtype=BF, device=DEV)
    dw, db = torch.zeros(E, 1, C, 2, 2, device=DEV), torch.zeros(E, device=DEV)
    L.call("es_gen_out_bwd", cuda(nhwc(x), BF), cuda(w), C * 4, 1, Hs, Ws, C, 2, 2, 1, img1, img2, cuda(d1.reshape(B, -1)),
           cuda(d2.reshape(B, -1)), gg, E, R, 1, dx, dw, db)
    check("gen out conv bwd dx", dx.float().reshape(R, Hs, Ws, C), nhwc(lx.grad), 5e-3)
    check("gen out conv bwd dw", dw, lw.grad, 1e-4)
    check("gen out conv bwd db", db, lb.grad, 1e-4)


# ----------------------------------------------------------------------------------------------------------- K3
CONVS = {  # Ci,Hi,Wi,Co,KH,KW,stride,pad
    "D.conv1": (1, 56, 30, 32, 3, 3, 1, 0), "D.conv2": (32, 27, 14, 16, 3, 3, 1, 0),
    "A.conv1": (1, 56, 30, 32, 5, 5, 2, 1), "A.res1.conv1": (32, 26, 13, 32, 5, 5, 2, 2),
    "A.res1.conv2": (32, 13, 7, 32, 5, 5, 1, 2), "A.res1.down": (32, 26, 13, 32, 1, 1, 2, 0),
    "A.res2.conv1": (32, 12, 6, 64, 5, 5, 2, 2), "A.res2.conv2": (64, 6, 3, 64, 5, 5, 1, 2),
    "nA.conv4": (128, 3, 17, 256, 3, 3, 1, 0),
}


@pytest.mark.parametrize("name", list(CONVS))
def test_conv2d_fwd_bwd(name):
    Ci, Hi, Wi, Co, KH, KW, st, pad = CONVS[name]
    Ho, Wo = (Hi + 2 * pad - KH) // st + 1, (Wi + 2 * pad - KW) // st + 1
    geo = L.ESConv2d(Ci, Hi, Wi, Co, Ho, Wo, KH, KW, st, pad)
    counts, slots = [7, 0, 12], [2, 1, 0]
    grp, R = groups(counts, slots)
    g = G(len(name))
    x = torch.randn(R, Ci, Hi, Wi, generator=g)
    w = torch.randn(3, Co, Ci, KH, KW, generator=g) / math.sqrt(Ci * KH * KW)
    b = torch.randn(3, Co, generator=g) * .1
    dy = torch.randn(R, Co, Ho, Wo, generator=g)
    lx, lw, lb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    outs, off = [], 0
    for c, s in zip(counts, slots):
        if c:
            o = F.conv2d(lx[off:off + c], lw[s], lb[s], stride=st, padding=pad)
            outs.append(o)
            (o * dy[off:off + c]).sum().backward()
        off += c
    y = torch.zeros(R, Co, Ho, Wo, device=DEV)
    L.call("es_conv2d_fwd", cuda(x), cuda(w), cuda(b), Co * Ci * KH * KW, Co, geo, grp, 3, R, y)
    check(f"conv2d fwd {name}", y, torch.cat(outs).detach(), 1e-5)
    dx = torch.ones(R, Ci, Hi, Wi, device=DEV)
    L.call("es_conv2d_bwd_data", cuda(dy), cuda(w), Co * Ci * KH * KW, geo, grp, 3, R, dx, 0)
    check(f"conv2d bwd data {name}", dx, lx.grad, 1e-5)
    L.call("es_conv2d_bwd_data", cuda(dy), cuda(w), Co * Ci * KH * KW, geo, grp, 3, R, dx, 1)
    check(f"conv2d bwd data (accumulate) {name}", dx, 2 * lx.grad, 1e-5)
    dw, db = torch.zeros(3, Co, Ci, KH, KW, device=DEV), torch.zeros(3, Co, device=DEV)
    L.call("es_conv2d_bwd_weight", cuda(x), cuda(dy), geo, grp, 3, R, dw, db, Co * Ci * KH * KW, Co)
    check(f"conv2d bwd weight {name}", dw, lw.grad, 2e-5)
    check(f"conv2d bwd bias {name}", db, lb.grad, 2e-5)


@pytest.mark.parametrize("arch", ["proton", "neutron"])
def test_disc_fused_trunk(arch):
    """Fused discriminator trunk (stem + stage 2, forward and backward) against torch autograd on sparse images (flat
    regions make every pooling window a tie: the arg-max rule — first maximum in scan order — is what is tested)."""
    H, W, pk = (56, 30, (2, 1)) if arch == "proton" else (44, 44, (2, 2))
    counts, slots = [5, 0, 9, 1], [2, 1, 0, 3]
    grp, R = groups(counts, slots)
    S = 4
    g = G(7 if arch == "proton" else 8)
    img = torch.log1p((torch.rand(R, 1, H, W, generator=g) < 0.03).float() * torch.empty(R, 1, H, W).exponential_(generator=g) * 20)
    cond = torch.randn(R, 9, generator=g)
    w0 = torch.randn(S, 32, 1, 3, 3, generator=g) / 3
    b0 = torch.randn(S, 32, generator=g) * .1
    g0, be0 = 1 + .2 * torch.randn(S, 32, generator=g), .1 * torch.randn(S, 32, generator=g)
    w1 = torch.randn(S, 16, 32, 3, 3, generator=g) / math.sqrt(288)
    b1 = torch.randn(S, 16, generator=g) * .1
    g1, be1 = 1 + .2 * torch.randn(S, 16, generator=g), .1 * torch.randn(S, 16, generator=g)
    H1, W1 = (H - 2) // 2, (W - 2) // 2
    Hp, Wp = (H1 - 2) // pk[0], (W1 - 2) // pk[1]
    flat = 16 * Hp * Wp
    ldf = flat + 9
    dfc = torch.randn(R, ldf, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (img, w0, b0, g0, be0, w1, b1, g1, be1)]
    limg, lw0, lb0, lg0, lbe0, lw1, lb1, lg1, lbe1 = leaves
    ref_p1, ref_fc, off = [], [], 0
    for c, s in zip(counts, slots):
        if c:
            x = F.conv2d(limg[off:off + c], lw0[s], lb0[s])
            x = F.max_pool2d(F.leaky_relu(F.group_norm(x, 8, lg0[s], lbe0[s]), 0.1), 2)
            ref_p1.append(x.detach())
            y = F.conv2d(x, lw1[s], lb1[s])
            y = F.max_pool2d(F.leaky_relu(F.group_norm(y, 8, lg1[s], lbe1[s]), 0.1), pk).flatten(1)
            ref_fc.append(y.detach())
            (y * dfc[off:off + c, :flat]).sum().backward()
        off += c
    d = lambda t: cuda(t.reshape(t.shape[0], -1))
    p1, st1 = torch.zeros(R, 32, H1, W1, device=DEV), torch.zeros(R, 8, 2, device=DEV)
    L.call("es_disc_stem_fwd", d(img), d(w0), 288, d(b0), 32, d(g0), d(be0), 32, H, W, grp, len(counts), R, p1, st1)
    check(f"disc stem fwd {arch}", p1, torch.cat(ref_p1), 1e-5)
    y2, st2 = torch.zeros(R, 16, (H1 - 2) * (W1 - 2), device=DEV), torch.zeros(R, 8, 2, device=DEV)
    fcin = torch.zeros(R, ldf, device=DEV)
    L.call("es_disc_stage2_fwd", p1, d(w1), 16 * 288, d(b1), 16, d(g1), d(be1), 16, cuda(cond), H1, W1, pk[1], grp,
           len(counts), R, y2, st2, fcin, ldf)
    check(f"disc stage2 fwd {arch}", fcin[:, :flat], torch.cat(ref_fc), 2e-5)
    assert torch.equal(fcin[:, flat:].cpu(), cond)
    for want_w in (True, False):
        dp1 = torch.zeros(R, 32, H1, W1, device=DEV)
        dw1, db1 = torch.zeros(S, 16 * 288, device=DEV), torch.zeros(S, 16, device=DEV)
        dg1, dbe1 = torch.zeros(S, 16, device=DEV), torch.zeros(S, 16, device=DEV)
        L.call("es_disc_stage2_bwd", cuda(dfc), ldf, y2, st2, p1, d(w1), 16 * 288, d(g1), d(be1), 16, H1, W1, pk[1], grp,
               len(counts), R, dp1, dw1 if want_w else None, 16 * 288, db1 if want_w else None, 16,
               dg1 if want_w else None, dbe1 if want_w else None)
        d_img = torch.zeros(R, H * W, device=DEV)
        dw0, db0 = torch.zeros(S, 288, device=DEV), torch.zeros(S, 32, device=DEV)
        dg0, dbe0 = torch.zeros(S, 32, device=DEV), torch.zeros(S, 32, device=DEV)
        L.call("es_disc_stem_bwd", dp1, d(img), d(w0), 288, d(b0), 32, d(g0), d(be0), 32, st1, H, W, grp, len(counts), R,
               d_img, dw0 if want_w else None, 288, db0 if want_w else None, dg0 if want_w else None,
               dbe0 if want_w else None)
        check(f"disc trunk d_img {arch} w={want_w}", d_img, limg.grad.reshape(R, -1), 2e-4)
        if want_w:
            for nm, got, ref in (("dw1", dw1, lw1.grad), ("db1", db1, lb1.grad), ("dgamma1", dg1, lg1.grad),
                                 ("dbeta1", dbe1, lbe1.grad), ("dw0", dw0, lw0.grad), ("db0", db0, lb0.grad),
                                 ("dgamma0", dg0, lg0.grad), ("dbeta0", dbe0, lbe0.grad)):
                check(f"disc trunk {nm} {arch}", got, ref.reshape(S, -1), 2e-4)
            # weight gradients only (discriminator step: no image gradient requested)
            dw0b = torch.zeros(S, 288, device=DEV)
            L.call("es_disc_stem_bwd", dp1, d(img), d(w0), 288, d(b0), 32, d(g0), d(be0), 32, st1, H, W, grp, len(counts),
                   R, None, dw0b, 288, torch.zeros(S, 32, device=DEV), torch.zeros(S, 32, device=DEV),
                   torch.zeros(S, 32, device=DEV))
            check(f"disc stem dw0 (no d_img) {arch}", dw0b, lw0.grad.reshape(S, -1), 2e-4)


@pytest.mark.parametrize("C,H,W,groups_n,act", [(32, 54, 28, 8, 2), (16, 25, 12, 8, 2), (32, 27, 14, 8, 1), (64, 6, 3, 32, 0)])
def test_groupnorm_fwd_bwd(C, H, W, groups_n, act):
    counts, slots = [4, 5], [1, 0]
    grp, R = groups(counts, slots)
    g = G(C + H)
    x = torch.randn(R, C, H, W, generator=g) * 2 + .5
    ga, be = 1 + .2 * torch.randn(2, C, generator=g), .2 * torch.randn(2, C, generator=g)
    dy = torch.randn(R, C, H, W, generator=g)
    lx, lg, lb = x.clone().requires_grad_(True), ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    f = {0: lambda t: t, 1: F.relu, 2: orc.lrelu}[act]
    outs, off = [], 0
    for c, s in zip(counts, slots):
        o = f(F.group_norm(lx[off:off + c], groups_n, lg[s], lb[s], 1e-5))
        outs.append(o)
        (o * dy[off:off + c]).sum().backward()
        off += c
    y, stats = torch.zeros(R, C, H, W, device=DEV), torch.zeros(R, groups_n, 2, device=DEV)
    L.call("es_groupnorm_fwd", cuda(x), cuda(ga), cuda(be), C, C, H * W, groups_n, act, grp, 2, R, y, stats)
    check(f"groupnorm fwd C={C} act={act}", y, torch.cat(outs).detach(), 1e-5)
    dx, dga, dbe = torch.zeros_like(y), torch.zeros(2, C, device=DEV), torch.zeros(2, C, device=DEV)
    L.call("es_groupnorm_bwd", cuda(dy), cuda(x), stats, cuda(ga), cuda(be), C, C, H * W, groups_n, act, grp, 2, R, dx, dga, dbe)
    check(f"groupnorm bwd dx C={C}", dx, lx.grad, 5e-5)
    check(f"groupnorm bwd dgamma C={C}", dga, lg.grad, 5e-5)
    check(f"groupnorm bwd dbeta C={C}", dbe, lb.grad, 5e-5)


@pytest.mark.parametrize("F_,act", [(128, 2), (64, 2)])
def test_layernorm_fwd_bwd(F_, act):
    counts, slots = [9, 0, 20], [2, 1, 0]
    grp, R = groups(counts, slots)
    g = G(F_)
    x = torch.randn(R, F_, generator=g) * 2 + .3
    ga, be = 1 + .2 * torch.randn(3, F_, generator=g), .2 * torch.randn(3, F_, generator=g)
    dy = torch.randn(R, F_, generator=g)
    lx, lg, lb = x.clone().requires_grad_(True), ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    outs, off = [], 0
    for c, s in zip(counts, slots):
        if c:
            o = orc.lrelu(F.layer_norm(lx[off:off + c], (F_,), lg[s], lb[s], 1e-5))
            outs.append(o)
            (o * dy[off:off + c]).sum().backward()
        off += c
    y, stats = torch.zeros(R, F_, device=DEV), torch.zeros(R, 2, device=DEV)
    L.call("es_layernorm_fwd", cuda(x), cuda(ga), cuda(be), F_, F_, act, grp, 3, R, y, stats)
    check(f"layernorm fwd F={F_}", y, torch.cat(outs).detach(), 1e-5)
    dx, dga, dbe = torch.zeros_like(y), torch.zeros(3, F_, device=DEV), torch.zeros(3, F_, device=DEV)
    L.call("es_layernorm_bwd", cuda(dy), cuda(x), stats, cuda(ga), cuda(be), F_, F_, act, grp, 3, R, dx, dga, dbe)
    check(f"layernorm bwd dx F={F_}", dx, lx.grad, 5e-5)
    check(f"layernorm bwd dgamma F={F_}", dga, lg.grad, 5e-5)
    check(f"layernorm bwd dbeta F={F_}", dbe, lb.grad, 5e-5)


@pytest.mark.parametrize("C,H,W,k,s", [(32, 54, 28, (2, 2), (2, 2)), (16, 25, 12, (2, 1), (2, 1)), (32, 27, 14, (2, 2), (1, 1)), (64, 6, 3, (2, 2), (1, 1))])
def test_maxpool(C, H, W, k, s):
    R = 5
    g = G(H)
    x = torch.randn(R, C, H, W, generator=g).requires_grad_(True)
    o = F.max_pool2d(x, k, s)
    dy = torch.randn(o.shape, generator=g)
    (o * dy).sum().backward()
    y = torch.zeros(o.shape, device=DEV)
    idx = torch.zeros(o.shape, dtype=torch.uint8, device=DEV)
    L.call("es_maxpool_fwd", cuda(x.detach()), C, H, W, k[0], k[1], s[0], s[1], R, y, idx)
    assert torch.equal(y.cpu(), o.detach())
    dx = torch.zeros(R, C, H, W, device=DEV)
    L.call("es_maxpool_bwd", cuda(dy), idx, C, H, W, k[0], k[1], s[0], s[1], R, dx)
    check(f"maxpool bwd {H}x{W}", dx, x.grad, 1e-6)


@pytest.mark.parametrize("I,O", [(2313, 128), (128, 64), (64, 1), (64, 2)])
def test_linear(I, O):
    counts, slots = [70, 0, 33], [0, 1, 2]
    grp, R = groups(counts, slots)
    g = G(I)
    x = torch.randn(R, I, generator=g)
    w, b = torch.randn(3, O, I, generator=g) / math.sqrt(I), torch.randn(3, O, generator=g) * .1
    dy = torch.randn(R, O, generator=g)
    lx, lw, lb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    outs, off = [], 0
    for c, s in zip(counts, slots):
        if c:
            o = F.linear(lx[off:off + c], lw[s], lb[s])
            outs.append(o)
            (o * dy[off:off + c]).sum().backward()
        off += c
    y = torch.zeros(R, O, device=DEV)
    L.call("es_linear_fwd", cuda(x), I, cuda(w), cuda(b), O * I, O, I, O, grp, 3, R, y)
    check(f"linear fwd {I}->{O}", y, torch.cat(outs).detach(), 1e-5)
    dx = torch.zeros(R, I, device=DEV)
    L.call("es_linear_bwd_data", cuda(dy), cuda(w), O * I, I, O, grp, 3, R, dx, I)
    check(f"linear bwd data {I}->{O}", dx, lx.grad, 1e-5)
    dw, db = torch.zeros(3, O, I, device=DEV), torch.zeros(3, O, device=DEV)
    L.call("es_linear_bwd_weight", cuda(x), I, cuda(dy), I, O, grp, 3, R, dw, db, O * I, O)
    check(f"linear bwd weight {I}->{O}", dw, lw.grad, 2e-5)
    check(f"linear bwd bias {I}->{O}", db, lb.grad, 2e-5)


@pytest.mark.parametrize("multi_cta", [False, True])
@pytest.mark.parametrize("O,I", [(32, 9), (16, 288), (128, 2313), (64, 128), (1, 64)])
def test_spectral_norm(O, I, multi_cta):
    S = 3
    sc_f = torch.empty(S, I + O + 2, device=DEV) if multi_cta else None
    sc_b = torch.empty(S, device=DEV) if multi_cta else None
    g = G(O * I)
    w = torch.randn(S, O, I, generator=g) / math.sqrt(I)
    u = F.normalize(torch.randn(S, O, generator=g), dim=1)
    v = F.normalize(torch.randn(S, I, generator=g), dim=1)
    grp, _ = groups([4, 0, 5])  # slot 1 inactive: its u, v must not advance
    du, dv = cuda(u.clone()), cuda(v.clone())
    wsn = torch.zeros(S, O, I, device=DEV)
    sig, uu, vu = torch.zeros(S, device=DEV), torch.zeros(S, O, device=DEV), torch.zeros(S, I, device=DEV)
    L.call("es_spectral_norm_fwd", cuda(w), du, dv, O * I, O, I, S, O, I, 1, grp, wsn, O * I, sig, uu, vu, sc_f)
    dwsn = torch.randn(S, O, I, generator=g)
    dwo = torch.zeros(S, O, I, device=DEV)
    L.call("es_spectral_norm_bwd", cuda(dwsn), wsn, uu, vu, sig, O * I, S, O, I, dwo, O * I, grp, sc_b)
    for s in (0, 2):
        sd = {"l.weight_orig": w[s].clone().requires_grad_(True), "l.weight_u": u[s].clone(), "l.weight_v": v[s].clone()}
        ws = orc.spectral_norm_weight(sd, "l", True)
        (ws * dwsn[s]).sum().backward()
        check(f"spectral norm W/sigma {O}x{I} slot {s}", wsn[s], ws.detach(), 2e-5)
        check(f"spectral norm u {O}x{I}", du[s], sd["l.weight_u"], 2e-5)
        check(f"spectral norm v {O}x{I}", dv[s], sd["l.weight_v"], 2e-5)
        check(f"spectral norm bwd {O}x{I}", dwo[s], sd["l.weight_orig"].grad, 1e-4)
    assert torch.equal(du[1].cpu(), u[1]) and torch.equal(dv[1].cpu(), v[1])


def test_elementwise_and_adam():
    g = G(77)
    n = 10007
    a, b = torch.randn(n, generator=g), torch.randn(n, generator=g)
    y = torch.zeros(n, device=DEV)
    L.call("es_add_relu_fwd", cuda(a), cuda(b), n, y)
    assert torch.equal(y.cpu(), F.relu(a + b))
    dx = torch.zeros(n, device=DEV)
    L.call("es_relu_bwd", cuda(b), y, n, dx)
    assert torch.equal(dx.cpu(), torch.where(F.relu(a + b) > 0, b, torch.zeros(())))
    x = torch.randn(6, 64, 5, 2, generator=g)
    o = torch.zeros(6, 64, device=DEV)
    L.call("es_gap_fwd", cuda(x), 64, 10, 6, o)
    check("gap fwd", o, x.mean((2, 3)), 1e-6)
    # Adam: 3 slots, slot 1 inactive, two steps
    S, N = 3, 5000
    p = torch.randn(S, N, generator=g)
    st = [orc.AdamState({"p": p[s].clone()}, 1e-3) for s in range(S)]
    pp = [{"p": p[s].clone()} for s in range(S)]
    dp, dm, dv = cuda(p.clone()), torch.zeros(S, N, device=DEV), torch.zeros(S, N, device=DEV)
    steps = torch.zeros(S, dtype=torch.int32, device=DEV)
    grp, _ = groups([3, 0, 2])
    for it in range(2):
        gr = torch.randn(S, N, generator=g)
        L.call("es_adam_step", dp, cuda(gr), dm, dv, N, N, S, 1e-3, 0.9, 0.999, 1e-8, steps, grp)
        for s in (0, 2):
            st[s].apply(pp[s], {"p": gr[s]})
    assert steps.cpu().tolist() == [2, 0, 2]
    for s in (0, 2):
        check(f"adam slot {s}", dp[s], pp[s]["p"], 1e-6)
    assert torch.equal(dp[1].cpu(), p[1])


def test_adam_pipelined_rectangles_equal_one_launch():
    """MoEWrapper._adam_pipelined through the real es_adam_step: row blocks + both column remainders, and column blocks of
    single rows, against ONE launch over the arena — bit-identical parameters, moments and counters (skipped slot untouched)"""
    from types import SimpleNamespace
    from expertsim.models.moe import MoEWrapper
    g = G(31)
    E, n = 5, 1 << 16
    P0, G0 = torch.randn(E, n, generator=g), torch.randn(E, n, generator=g)
    M0, V0 = torch.randn(E, n, generator=g) * .1, torch.rand(E, n, generator=g) * .1
    grp, _ = groups([4, 0, 2, 0, 9])

    def arena():
        return SimpleNamespace(P=cuda(P0.clone()), G=cuda(G0.clone()), M=cuda(M0.clone()), V=cuda(V0.clone()),
                               steps=torch.tensor([3, 0, 7, 7, 1], dtype=torch.int32, device=DEV), n=n, E=E, version=0)
    one = arena()
    MoEWrapper._adam(one, 1e-3, grp)
    red = SimpleNamespace(join=lambda: None)
    for rects in ([(0, 2, 4096, 60000, None), (2, 4, 4096, 60000, None), (4, 5, 4096, 60000, None)],
                  [(e, e + 1, c, min(c + 24576, n), None) for e in range(E) for c in range(0, n, 24576)]):
        pip = arena()
        MoEWrapper._adam_pipelined(SimpleNamespace(), pip, 1e-3, grp, rects, red)
        for k in ("P", "M", "V", "steps"):
            assert torch.equal(getattr(pip, k), getattr(one, k)), k
    assert one.steps.cpu().tolist() == [4, 0, 8, 7, 2] and torch.equal(one.P[1].cpu(), P0[1])


def test_expm1_scatter():
    g = G(5)
    R, HW = 9, 1680
    img = torch.rand(R, HW, generator=g) * 3
    perm = torch.randperm(R, generator=g).to(torch.int32)
    o64 = torch.zeros(R, HW, dtype=torch.float64, device=DEV)
    L.call("es_expm1_scatter", cuda(img), cuda(perm), R, HW, o64, None)
    want = torch.zeros(R, HW, dtype=torch.float64)
    want[perm.long()] = torch.expm1(img).double()
    check("expm1 + scatter (f64)", o64, want, 1e-6)


# ------------------------------------------------------------------------------------- x2-upsample folding (tap tables)
@pytest.mark.parametrize("Hs,Ws,C,KH,KW,pad,N", [(18, 10, 512, 4, 4, 1, 256), (13, 13, 128, 3, 3, 0, 256), (24, 24, 256, 3, 3, 0, 128),
                                                  (5, 4, 128, 4, 4, 1, 64)])
def test_up2_folded_conv_fwd_dgrad_wgrad(Hs, Ws, C, KH, KW, pad, N):
    """conv(upsample_x2_nearest(x)) computed as four phase convs with pre-summed taps, its data gradient as ONE table-conv over
    dy on the low-resolution grid and its weight gradient per phase + unfold — against fp32 torch autograd on the same
    (bf16-rounded) x and dy.  Weights are folded in fp32 and rounded once, so the bound is the bf16 tolerance."""
    from expertsim._nets import Up2Conv
    counts, slots = [3, 0, 2], [2, 0, 1]
    E = 3
    grp, R = groups(counts, slots)
    g = G(Hs * 1000 + C + KH)
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    w = torch.randn(E, N, C, KH, KW, generator=g) / math.sqrt(KH * KW * C)
    bias = torch.randn(E, N, generator=g) * 0.1
    u = Up2Conv(Hs, Ws, C, KH, KW, pad, N)
    u.alloc(E, DEV)
    dw_ref = cuda(w)
    u.fold(dw_ref, N * C * KH * KW, E)
    y = torch.zeros(R, u.Ho * u.Wo, N, dtype=BF, device=DEV)
    u.forward(cuda(x, BF), cuda(bias), N, y, grp, E, R)
    dy = bf16_round(torch.randn(R, u.Ho, u.Wo, N, generator=g))
    want_y, want_dx, want_dw, off = [], [], torch.zeros_like(w), 0
    for c, s in zip(counts, slots):
        if c == 0:
            continue
        xi = x[off:off + c].permute(0, 3, 1, 2).clone().requires_grad_(True)
        wi = w[s].clone().requires_grad_(True)
        yi = F.conv2d(F.interpolate(xi, scale_factor=2, mode="nearest"), wi, bias[s], padding=pad)
        (yi * dy[off:off + c].permute(0, 3, 1, 2)).sum().backward()
        want_y.append(yi.detach().permute(0, 2, 3, 1))
        want_dx.append(xi.grad.permute(0, 2, 3, 1))
        want_dw[s] = wi.grad
        off += c
    check(f"up2 folded fwd {Hs}x{Ws}x{C} k{KH} N{N}", y.float().view(R, u.Ho, u.Wo, N), torch.cat(want_y), 8e-3, 4e-2)
    dx = torch.zeros(R, Hs * Ws, C, dtype=BF, device=DEV)
    u.dgrad(cuda(dy, BF), dx, grp, E, R)
    check(f"up2 folded dgrad {Hs}x{Ws}x{C}", dx.float().view(R, Hs, Ws, C), torch.cat(want_dx), 8e-3, 4e-2)
    dw = torch.zeros(E, N, C, KH, KW, device=DEV)
    u.wgrad(cuda(x, BF), cuda(dy, BF), dw, N * C * KH * KW, grp, E, R)
    for s in (1, 2):
        check(f"up2 folded wgrad slot {s} {Hs}x{Ws}x{C}", dw[s], want_dw[s], 5e-3, 2e-2)
    assert float(dw[0].abs().max()) == 0.0


def test_y_folded_conv2_fwd_wgrad():
    """proton conv2: 35x19 -> 56x30 nearest (rows repeat with period 8 <- 5 source rows), k4/p1.  Folded along y only:
    8 row classes with 2..3 distinct source rows instead of 4; x keeps the nearest map.  Forward and weight gradient against
    fp32 torch on bf16-rounded x / dy."""
    from expertsim._nets import FoldedConv
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = 35, 19, 256, 56, 30, 4, 4, 1, 128
    counts, slots, E = [2, 0, 3], [2, 0, 1], 3
    grp, R = groups(counts, slots)
    g = G(4242)
    x = bf16_round(torch.randn(R, Hs, Ws, C, generator=g))
    w = torch.randn(E, N, C, KH, KW, generator=g) / math.sqrt(KH * KW * C)
    bias = torch.randn(E, N, generator=g) * 0.1
    f = FoldedConv(Hs, Ws, C, Hu, Wu, KH, KW, pad, N, (True, False))
    assert len(f.classes) == 8 and abs(f.executed_ratio - 0.718) < 1e-3
    assert f.has_dgrad and len(f.dgrad_classes) == 5 and f.dg_grid == (Hs, Wu)
    f.alloc(E, DEV)
    f.fold(cuda(w), N * C * KH * KW, E)
    y = torch.zeros(R, f.Ho * f.Wo, N, dtype=BF, device=DEV)
    f.forward(cuda(x, BF), cuda(bias), N, y, grp, E, R)
    dy = bf16_round(torch.randn(R, f.Ho, f.Wo, N, generator=g))
    want_y, want_dw, want_dx, off = [], torch.zeros_like(w), [], 0
    for c, s in zip(counts, slots):
        if c == 0:
            continue
        xi = x[off:off + c].permute(0, 3, 1, 2)
        wi = w[s].clone().requires_grad_(True)
        # nearest upsampling is separable: x first (the grid the folded data gradient lives on), then y
        xx = F.interpolate(xi, size=(Hs, Wu), mode="nearest").requires_grad_(True)
        yi = F.conv2d(F.interpolate(xx, size=(Hu, Wu), mode="nearest"), wi, bias[s], padding=pad)
        assert torch.equal(F.interpolate(xx, size=(Hu, Wu), mode="nearest"), F.interpolate(xi, size=(Hu, Wu), mode="nearest"))
        (yi * dy[off:off + c].permute(0, 3, 1, 2)).sum().backward()
        want_y.append(yi.detach().permute(0, 2, 3, 1))
        want_dw[s] = wi.grad
        want_dx.append(xx.grad.permute(0, 2, 3, 1))
        off += c
    check("y-folded conv2 fwd", y.float().view(R, f.Ho, f.Wo, N), torch.cat(want_y), 8e-3, 4e-2)
    dw = torch.zeros(E, N, C, KH, KW, device=DEV)
    f.wgrad(cuda(x, BF), cuda(dy, BF), dw, N * C * KH * KW, grp, E, R)
    for s in (1, 2):
        check(f"y-folded conv2 wgrad slot {s}", dw[s], want_dw[s], 5e-3, 2e-2)
    assert float(dw[0].abs().max()) == 0.0
    dx = torch.zeros(R, Hs * Wu, C, dtype=BF, device=DEV)
    f.dgrad(cuda(dy, BF), dx, grp, E, R)
    check("y-folded conv2 dgrad", dx.float().view(R, Hs, Wu, C), torch.cat(want_dx), 1e-2, 5e-2)


def test_gn_lrelu_fwd_upx_and_conv2_on_the_upsampled_source():
    """proton conv2 with the x half of its upsample materialised by the GroupNorm kernel (es_gn_lrelu_fwd_upx) — the layout
    the generator uses: norm+LeakyReLU output stored [35, 30] (nearest along x, bit-equal to torch's map), then the y-folded
    conv reads its source directly (TMA-fed variant): forward, weight gradient and data gradient against fp32 torch."""
    from expertsim._nets import FoldedConv
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = 35, 19, 256, 56, 30, 4, 4, 1, 128
    counts, slots, E = [2, 0, 3], [2, 0, 1], 3
    grp, R = groups(counts, slots)
    g = G(777)
    pre = bf16_round(torch.randn(R, Hs * Ws, C, generator=g))
    gamma, beta = 1 + 0.1 * torch.randn(E, C, generator=g), 0.1 * torch.randn(E, C, generator=g)
    y0, st0 = torch.zeros(R, Hs * Ws, C, dtype=BF, device=DEV), torch.zeros(R, 32, 2, device=DEV)
    L.call("es_gn_lrelu_fwd", cuda(pre, BF), cuda(gamma), cuda(beta), C, Hs * Ws, C, 32, grp, E, R, y0, st0)
    yu, st1 = torch.zeros(R, Hs * Wu, C, dtype=BF, device=DEV), torch.zeros(R, 32, 2, device=DEV)
    L.call("es_gn_lrelu_fwd_upx", cuda(pre, BF), cuda(gamma), cuda(beta), C, Hs, Ws, Wu, C, 32, grp, E, R, yu, st1)
    torch.cuda.synchronize()
    # the column map is torch's nearest rule, exactly: every output column is a bit-copy of the first column with the same source
    yu4 = yu.float().view(R, Hs, Wu, C)
    src = torch.clamp((torch.arange(Wu) * (Ws / Wu)).floor().long(), max=Ws - 1)
    first = torch.tensor([int((src == s_).nonzero()[0]) for s_ in range(Ws)])
    rep = yu4[:, :, first.to(DEV), :]                                  # one representative column per source column
    want_u = F.interpolate(rep.permute(0, 3, 1, 2), size=(Hs, Wu), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(yu4, want_u)
    # and the values are the plain kernel's (the statistics are summed in arrival order: not bit-equal between two launches)
    check("gn_lrelu_fwd_upx values", rep, y0.float().view(R, Hs, Ws, C), 1e-3)
    check("gn_lrelu_fwd_upx stats", st1, st0, 1e-5)
    # the conv on the upsampled source
    x = yu.float().view(R, Hs, Wu, C).cpu()                         # already bf16 values
    w = torch.randn(E, N, C, KH, KW, generator=g) / math.sqrt(KH * KW * C)
    bias = torch.randn(E, N, generator=g) * 0.1
    f = FoldedConv(Hs, Wu, C, Hu, Wu, KH, KW, pad, N, (True, False))
    assert len(f.classes) == 8 and f.has_dgrad and len(f.dgrad_classes) == 5 and f.dg_grid == (Hs, Wu)
    assert all(c["g_fwd"].Wu == c["g_fwd"].Ws for c in f.classes)      # the conv reads its source directly
    f.alloc(E, DEV)
    f.fold(cuda(w), N * C * KH * KW, E)
    y = torch.zeros(R, f.Ho * f.Wo, N, dtype=BF, device=DEV)
    f.forward(yu, cuda(bias), N, y, grp, E, R)
    dy = bf16_round(torch.randn(R, f.Ho, f.Wo, N, generator=g))
    want_y, want_dw, want_dx, off = [], torch.zeros_like(w), [], 0
    for c, s_ in zip(counts, slots):
        if c == 0:
            continue
        xx = x[off:off + c].permute(0, 3, 1, 2).clone().requires_grad_(True)
        wi = w[s_].clone().requires_grad_(True)
        yi = F.conv2d(F.interpolate(xx, size=(Hu, Wu), mode="nearest"), wi, bias[s_], padding=pad)
        (yi * dy[off:off + c].permute(0, 3, 1, 2)).sum().backward()
        want_y.append(yi.detach().permute(0, 2, 3, 1))
        want_dw[s_] = wi.grad
        want_dx.append(xx.grad.permute(0, 2, 3, 1))
        off += c
    check("conv2 on the x-upsampled source: fwd", y.float().view(R, f.Ho, f.Wo, N), torch.cat(want_y), 8e-3, 4e-2)
    dw = torch.zeros(E, N, C, KH, KW, device=DEV)
    f.wgrad(yu, cuda(dy, BF), dw, N * C * KH * KW, grp, E, R)
    for s_ in (1, 2):
        check(f"conv2 on the x-upsampled source: wgrad slot {s_}", dw[s_], want_dw[s_], 5e-3, 2e-2)
    dx = torch.zeros(R, Hs * Wu, C, dtype=BF, device=DEV)
    f.dgrad(cuda(dy, BF), dx, grp, E, R)
    check("conv2 on the x-upsampled source: dgrad", dx.float().view(R, Hs, Wu, C), torch.cat(want_dx), 1e-2, 5e-2)


@pytest.mark.parametrize("layer", ["conv1_folded_upx", "conv2_yfolded", "conv3_plain"])
def test_groupnorm_statistics_fused_into_the_conv_epilogue(layer):
    """Norm fusion (generator.py:27-40: conv -> GroupNorm(32) -> LeakyReLU): es_igemm_*_fwd_sums accumulates per-(row, channel
    pair) sum / sum of squares of the stored bf16 outputs in the GEMM epilogue, es_gn_lrelu_apply_fwd normalises in one
    streaming pass.  Checked against (a) the sums recomputed from the stored conv output in fp64, (b) the two-pass
    es_gn_lrelu_fwd on the same conv output (statistics 1e-4, activations to bf16 rounding).  Ragged groups, an empty expert,
    and sample boundaries inside a warp of the epilogue (P is not a multiple of 32)."""
    import ctypes

    from expertsim._nets import FoldedConv, Up2Conv, conv_geom
    geo = {"conv1_folded_upx": (18, 10, 512, 36, 20, 4, 4, 1, 256), "conv2_yfolded": (35, 30, 256, 56, 30, 4, 4, 1, 128),
           "conv3_plain": (55, 29, 128, 55, 29, 3, 3, 1, 64)}[layer]
    Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
    counts, slots, E = [3, 1, 2], [2, 0, 1], 3
    grp, R = groups(counts, slots, min_rows=2)       # the middle expert holds ONE row: skipped (moe.py:126-135), its row stays
    live = torch.tensor([True] * 3 + [False] + [True] * 2, device=DEV)
    g = G(4242)
    x = cuda(bf16_round(torch.randn(R, Hs * Ws, C, generator=g)), BF)
    w = torch.randn(E, N, C, KH, KW, generator=g) / math.sqrt(KH * KW * C)
    bias = cuda(torch.randn(E, N, generator=g) * 0.5)            # a real mean offset: exercises E[x^2] - mean^2
    gm = conv_geom(*geo)
    Ho, Wo = gm.Ho, gm.Wo
    y = torch.zeros(R, Ho * Wo, N, dtype=BF, device=DEV)
    ps = torch.zeros(R, N // 2, 2, device=DEV)
    flag = (ctypes.c_int32 * 1)()
    if layer == "conv3_plain":
        wp, wd = torch.empty(E, N, KH, KW, C, dtype=BF, device=DEV), torch.empty(E, C, KH, KW, N, dtype=BF, device=DEV)
        L.call("es_pack_conv_weight", cuda(w), N * C * KH * KW, E, N, C, KH, KW, wp, wd)
        L.call("es_igemm_fwd_sums", x, wp, bias, N, y, gm, grp, E, R, ps, flag)
        fused = flag[0] == 1
    else:
        f = Up2Conv(Hs, Ws, C, KH, KW, pad, N) if layer == "conv1_folded_upx" else FoldedConv(Hs, Ws, C, Hu, Wu, KH, KW, pad, N, (True, False))
        f.alloc(E, DEV)
        f.fold(cuda(w), N * C * KH * KW, E)
        fused = f.forward(x, bias, N, y, grp, E, R, ps)
    torch.cuda.synchronize()
    assert fused, "the TMA-fed pair variants take these geometries and accumulate the sums"
    yf = y.double().view(R, Ho * Wo, N // 2, 2)
    want = torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1)          # [R, N/2, 2]
    check("epilogue channel-pair sums", ps[live], want[live].float(), 2e-5, 2e-4)
    assert float(ps[~live].abs().max()) == 0.0
    gamma, beta = cuda(1 + 0.1 * torch.randn(E, N, generator=g)), cuda(0.1 * torch.randn(E, N, generator=g))
    a0, st0 = torch.zeros(R, Ho * Wo, N, dtype=BF, device=DEV), torch.zeros(R, 32, 2, device=DEV)
    L.call("es_gn_lrelu_fwd", y, gamma, beta, N, Ho * Wo, N, 32, grp, E, R, a0, st0)
    a0.fill_(float("nan"))
    L.call("es_gn_lrelu_fwd", y, gamma, beta, N, Ho * Wo, N, 32, grp, E, R, a0, st0)
    a1, st1 = torch.full((R, Ho * Wo, N), float("nan"), dtype=BF, device=DEV), torch.zeros(R, 32, 2, device=DEV)
    L.call("es_gn_lrelu_apply_fwd", y, ps, gamma, beta, N, Ho, Wo, Wo, N, 32, grp, E, R, a1, st1)
    check("fused statistics vs two-pass", st1[live], st0[live], 1e-4, 1e-4)
    check("fused activations vs two-pass", a1.float(), a0.float(), 2e-3, 2e-2)
    # rows of a skipped expert are stored as zeros by both kernels: the strip kernels of the next conv's weight gradient read
    # activation strips past a group's end against zero-filled dy columns, and 0 x NaN would poison the sum
    assert float(a0[~live].float().abs().max()) == 0.0 and float(a1[~live].float().abs().max()) == 0.0
    if layer == "conv1_folded_upx":       # the layout proton conv2 reads: nearest-upsampled along x, 19 -> 30 columns
        wu = 30
        au, st2 = torch.full((R, Ho * wu, N), float("nan"), dtype=BF, device=DEV), torch.zeros(R, 32, 2, device=DEV)
        L.call("es_gn_lrelu_apply_fwd", y, ps, gamma, beta, N, Ho, Wo, wu, N, 32, grp, E, R, au, st2)
        src = torch.clamp((torch.arange(wu) * (Wo / wu)).floor().long(), max=Wo - 1).to(DEV)
        assert torch.equal(au.view(R, Ho, wu, N), a1.view(R, Ho, Wo, N)[:, :, src, :])


# ----------------------------------------------------------------------------------------------------------- preprocessing
def test_preprocess_kernels_match_golden_and_oracle():
    """arg-max coordinates (bit-exact) and per-condition-group pixel std (1e-5) through the host layer."""
    import os

    import numpy as np

    from expertsim.utils import preprocess as pp
    from oracle import preprocess_oracle as po
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_small.npz"))
    data, cond = torch.from_numpy(gold["data"]), torch.from_numpy(gold["cond"])
    pos = pp.max_coordinates(cuda(data))
    assert torch.equal(pos.cpu().long(), torch.from_numpy(gold["positions"]))
    assert torch.equal(pp.max_coordinates(cuda(data), as_float=True).cpu(), torch.from_numpy(gold["positions"]).float())
    std = pp.condition_group_std(cuda(cond), cuda(data))
    check("condition-group std (golden)", std, torch.from_numpy(gold["std"]), 1e-5)
    # larger seeded case, both detector shapes, ragged groups (1 .. ~40 members), ties and an all-equal image
    for H, W, seed in ((56, 30, 1), (44, 44, 2)):
        g = G(seed)
        n, ng = 3000, 211
        rows = torch.randn(ng, 9, generator=g)
        member = torch.randint(0, ng, (n,), generator=g)
        c = rows[member]
        img = torch.log1p((torch.rand(n, H, W, generator=g) < 0.03).float() * torch.ceil(torch.empty(n, H, W).exponential_(generator=g) * 20))
        img[5] = 0.0
        img[6, 3, 3] = img[6, 40, 20] = 9.0
        want_pos = po.max_coordinates(img.numpy())
        got_pos = pp.max_coordinates(cuda(img))
        assert torch.equal(got_pos.cpu().long(), torch.from_numpy(want_pos)), "arg-max coordinates must be bit-exact"
        want = po.condition_group_std(c.numpy(), img.numpy())
        got, gid, sums = pp.condition_group_std(cuda(c), cuda(img), return_groups=True)
        check(f"condition-group std {H}x{W}", got, torch.from_numpy(want), 1e-5)
        assert int(gid.max()) + 1 == len(torch.unique(member)) and float(got.max()) == 1.0
    with pytest.raises(RuntimeError):
        pp.max_coordinates(data)      # host tensors are refused: no CPU fallback


# ----------------------------------------------------------------------------------------------------------- BatchNorm (neutron)
@pytest.mark.parametrize("Hs,Ws,C,Hu,Wu", [(24, 24, 256, 24, 24), (46, 46, 128, 46, 46), (45, 45, 64, 45, 45), (13, 13, 128, 20, 22)])
def test_bn2d_nhwc_fwd_bwd(Hs, Ws, C, Hu, Wu):
    """BatchNorm2d + Dropout(0.2) + LeakyReLU on bf16 NHWC rows (neutron generator, fast path), two-pass batch of 3 experts:
    (1) injected keep-masks against torch autograd (per (expert, pass) batch statistics); (2) hashed dropout: the pattern
    the forward drew is exactly the one both backward kernels re-evaluate (backward with the mask recovered from the
    forward output == backward with the hash)."""
    counts, slots, E, p = [3, 0, 4], [2, 0, 1], 3, 0.2
    gg, B = groups(counts, slots, two_pass=True)
    R, P = 2 * B, Hs * Ws
    g = G(Hs + C)
    x = bf16_round(torch.randn(R, C, Hs, Ws, generator=g) * 1.5 + 0.3)
    da = bf16_round(torch.randn(R, C, Hu, Wu, generator=g))
    gamma, beta = 1 + .2 * torch.randn(E, C, generator=g), .1 * torch.randn(E, C, generator=g)
    keep = (torch.rand(R, C, Hs, Ws, generator=g) >= p).float()
    n_rows = torch.zeros(2 * E)
    lx, lg, lb = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    want_y, off = torch.zeros(R, C, Hs, Ws), 0
    for c, s in zip(counts, slots):
        for ps in range(2):
            if c:
                r0 = 2 * off + ps * c
                n_rows[2 * s + ps] = c
                bn = F.batch_norm(lx[r0:r0 + c], None, None, lg[s], lb[s], training=True, eps=1e-5)
                y = F.leaky_relu(bn * keep[r0:r0 + c] / (1 - p), 0.1)
                want_y[r0:r0 + c] = y.detach()
                up = F.interpolate(y, size=(Hu, Wu), mode="nearest") if (Hu, Wu) != (Hs, Ws) else y
                (up * da[r0:r0 + c]).sum().backward()
        off += c
    xd, dad = cuda(nhwc(x), BF), cuda(nhwc(da), BF)
    n_sg = cuda(n_rows * P)
    sums = torch.zeros(2 * E, C, 2, dtype=torch.float64, device=DEV)
    L.call("es_bn_stats_nhwc", xd, Hs, Ws, C, 0, gg, len(counts), R, 1, sums)
    stats = torch.zeros(2 * E, C, 2, device=DEV)
    act = torch.zeros(E, 4, dtype=torch.int32)
    for c, s in zip(counts, slots):
        act[s] = torch.tensor([0, c, s, c])
    L.call("es_bn_finalize", sums, n_sg, C, 2, 1, 0.1, None, None, None, 0, None, 0, cuda(act), E, stats)
    gd, bd = cuda(gamma), cuda(beta)
    y = torch.zeros(R, P, C, dtype=BF, device=DEV)
    L.call("es_bn_apply_fwd_nhwc", xd, Hs, Ws, C, 0, stats, gd, bd, C, None, cuda(keep.reshape(R, -1)), 0, p, gg, len(counts), R, 1, y)
    check(f"bn2d fwd C={C}", y.float().view(R, Hs, Ws, C), nhwc(want_y), 8e-3)

    def backward(mask, seed):
        s2 = torch.zeros(2 * E, C, 2, dtype=torch.float64, device=DEV)
        L.call("es_bn_bwd_reduce_nhwc", dad, Hs, Ws, Hu, Wu, C, 0, xd, stats, gd, bd, C, None, mask, seed, p, gg, len(counts), R, 1, s2)
        dx = torch.zeros(R, P, C, dtype=BF, device=DEV)
        L.call("es_bn_bwd_apply_nhwc", dad, Hs, Ws, Hu, Wu, C, 0, xd, stats, s2, n_sg, gd, bd, C, None, mask, seed, p, gg,
               len(counts), R, 1, dx)
        dg, db = torch.zeros(E, C, device=DEV), torch.zeros(E, C, device=DEV)
        L.call("es_bn_affine_grads", s2, C, 2, 1.0, None, cuda(act), E, dg, db, C)
        return dx, dg, db

    dx, dg, db = backward(cuda(keep.reshape(R, -1)), 0)
    check(f"bn2d bwd dx C={C}", dx.float().view(R, Hs, Ws, C), nhwc(lx.grad), 1e-2)
    check(f"bn2d bwd dgamma C={C}", dg, lg.grad, 5e-3)
    check(f"bn2d bwd dbeta C={C}", db, lb.grad, 5e-3)
    # hashed dropout: recover the forward's pattern from its output (a dropped element is exactly 0) and replay it as a mask
    seed = 0x1234567 + C
    yh = torch.zeros(R, P, C, dtype=BF, device=DEV)
    L.call("es_bn_apply_fwd_nhwc", xd, Hs, Ws, C, 0, stats, gd, bd, C, None, None, seed, p, gg, len(counts), R, 1, yh)
    rows_on = torch.cat([torch.arange(2 * o, 2 * o + 2 * c) for o, c in zip([0, 3, 3], counts) if c])
    kept = (yh.float().view(R, Hs, Ws, C)[rows_on] != 0).float()
    frac = 1 - kept.mean().item()
    assert abs(frac - p) < 0.01, f"dropout rate {frac:.4f}"
    mask_nchw = torch.zeros(R, C, Hs, Ws, device=DEV)
    mask_nchw[rows_on] = kept.permute(0, 3, 1, 2)
    dxa, dga, dba = backward(None, seed)
    dxb, dgb, dbb = backward(mask_nchw.reshape(R, -1).contiguous(), 0)
    # (the fp64 partial sums meet in atomics, so the two runs may differ in the last bit — a wrong keep decision would not)
    mism = (dxa != dxb).float().mean().item()
    assert mism < 1e-3, f"forward and backward must draw the same dropout pattern ({mism:.2e} of dx differs)"
    check(f"bn2d hashed-vs-replayed dx C={C}", dxa.float(), dxb.float(), 1e-4)
    check(f"bn2d hashed-vs-replayed dgamma C={C}", dga, dgb, 1e-5)
    check(f"bn2d hashed-vs-replayed dbeta C={C}", dba, dbb, 1e-5)
