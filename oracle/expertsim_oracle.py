"""CPU oracle for the ExpertSim MoE-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch on the CPU (fp32, autograd for
gradients), the algorithm of the reference's training step and batch inference.
It is the checker the CUDA path is compared with.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import it; the product package never does.

Parity pin: the reference ships NO tests or golden vectors,
so this restatement is pinned against the reference itself, imported in the build
container through ``oracle/pin_against_reference.py``, which runs the unmodified
reference classes (``/root/reference/expertsim``) and this file on identical
weights, inputs and injected noise and commits the results to ``tests/golden/``.

Every function cites the reference file:line (relative to /root/reference) that it
follows.  Noise is always an INPUT (never drawn here) and is indexed by ORIGINAL
sample index: the noise row used for the k-th sample routed to expert e is
``z[mask_e[k]]``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from itertools import combinations
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # every LeakyReLU on the path: routers/router.py:13-17, proton/generator.py:16-40
NORM_EPS = 1e-5
SN_EPS = 1e-12

IMAGE_SHAPE = {"proton": (56, 30), "neutron": (44, 44)}


# --------------------------------------------------------------------------------------
# state_dict layouts (measured on the reference modules)
# --------------------------------------------------------------------------------------
def _lin(name, out_f, in_f, bias=True):
    r = [(f"{name}.weight", (out_f, in_f), "w", in_f)]
    if bias:
        r.append((f"{name}.bias", (out_f,), "b", in_f))
    return r


def _conv(name, co, ci, kh, kw, bias=True):
    r = [(f"{name}.weight", (co, ci, kh, kw), "w", ci * kh * kw)]
    if bias:
        r.append((f"{name}.bias", (co,), "b", ci * kh * kw))
    return r


def _norm(name, c):
    return [(f"{name}.weight", (c,), "gamma", 0), (f"{name}.bias", (c,), "beta", 0)]


def _bn_spec(name, c):
    return _norm(name, c) + [
        (f"{name}.running_mean", (c,), "rm", 0),
        (f"{name}.running_var", (c,), "rv", 0),
        (f"{name}.num_batches_tracked", (), "nbt", 0),
    ]


def _sn_lin(name, out_f, in_f):
    return [
        (f"{name}.bias", (out_f,), "b", in_f),
        (f"{name}.weight_orig", (out_f, in_f), "w", in_f),
        (f"{name}.weight_u", (out_f,), "unit", 0),
        (f"{name}.weight_v", (in_f,), "unit", 0),
    ]


def _sn_conv(name, co, ci, kh, kw):
    return [
        (f"{name}.bias", (co,), "b", ci * kh * kw),
        (f"{name}.weight_orig", (co, ci, kh, kw), "w", ci * kh * kw),
        (f"{name}.weight_u", (co,), "unit", 0),
        (f"{name}.weight_v", (ci * kh * kw,), "unit", 0),
    ]


def param_spec(arch: str, kind: str, n_experts: int = 3, cond_dim: int = 9, noise_dim: int = 10):
    """(name, shape, role, fan_in) in state_dict order for one reference module."""
    s: list = []
    if kind == "router":  # routers/router.py:11-19
        s += _lin("fc_layers.0", 128, cond_dim) + _lin("fc_layers.2", 64, 128)
        s += _lin("fc_layers.4", 32, 64) + _lin("fc_layers.6", n_experts, 32)
    elif (arch, kind) == ("proton", "generator"):  # proton/generator.py:13-44
        s += _lin("fc1.0", 256, noise_dim + cond_dim) + _norm("fc1.1", 256)
        s += _lin("fc2.0", 512 * 18 * 10, 256) + _norm("fc2.1", 512 * 18 * 10)
        s += _conv("conv_layers.1", 256, 512, 4, 4) + _norm("conv_layers.2", 256)
        s += _conv("conv_layers.5", 128, 256, 4, 4) + _norm("conv_layers.6", 128)
        s += _conv("conv_layers.8", 64, 128, 3, 3) + _norm("conv_layers.9", 64)
        s += _conv("conv_layers.11", 1, 64, 2, 2)
    elif (arch, kind) == ("neutron", "generator"):  # neutron/generator.py:11-40
        s += _lin("fc1.0", 256, noise_dim + cond_dim) + _bn_spec("fc1.1", 256)
        s += _lin("fc2.0", 128 * 13 * 13, 256) + _bn_spec("fc2.1", 128 * 13 * 13)
        s += _conv("conv_layers.0", 256, 128, 3, 3) + _bn_spec("conv_layers.1", 256)
        s += _conv("conv_layers.5", 128, 256, 3, 3) + _bn_spec("conv_layers.6", 128)
        s += _conv("conv_layers.9", 64, 128, 2, 2) + _bn_spec("conv_layers.10", 64)
        s += _conv("conv_layers.13", 1, 64, 2, 2)
    elif kind == "discriminator":  # proton/discriminator.py:121-146, neutron/discriminator.py:11-39
        flat = 16 * 12 * 12 if arch == "proton" else 9 * 12 * 12
        s += _sn_conv("conv_layers.0", 32, 1, 3, 3) + _norm("conv_layers.1", 32)
        s += _sn_conv("conv_layers.4", 16, 32, 3, 3) + _norm("conv_layers.5", 16)
        s += _sn_lin("fc1.0", 128, flat + cond_dim) + _norm("fc1.1", 128)
        s += _sn_lin("fc2.0", 64, 128) + _norm("fc2.1", 64)
        s += _sn_lin("fc3", 1, 64)
    elif (arch, kind) == ("proton", "aux_reg"):  # proton/aux_reg.py:11-131
        fe = "feature_extractor"
        s += _conv(f"{fe}.conv1.0", 32, 1, 5, 5) + _norm(f"{fe}.conv1.1", 32)
        for blk, ci, co in (("res1", 32, 32), ("res2", 32, 64)):
            s += _conv(f"{fe}.{blk}.conv1.0", co, ci, 5, 5) + _norm(f"{fe}.{blk}.conv1.1", co)
            s += _conv(f"{fe}.{blk}.conv2.0", co, co, 5, 5) + _norm(f"{fe}.{blk}.conv2.1", co)
            s += _conv(f"{fe}.{blk}.downsample.0", co, ci, 1, 1) + _norm(f"{fe}.{blk}.downsample.1", co)
        s += _lin("regressor.0", 128, 64) + _norm("regressor.1", 128)
        s += _lin("regressor.4", 64, 128) + _norm("regressor.5", 64)
        s += _lin("regressor.8", 2, 64)
    elif (arch, kind) == ("neutron", "aux_reg"):  # neutron/aux_reg.py:8-80
        fe = "feature_extractor"
        for i, (ci, co) in enumerate(((1, 32), (32, 64), (64, 128), (128, 256)), start=1):
            s += _conv(f"{fe}.conv{i}", co, ci, 3, 3) + _bn_spec(f"{fe}.conv{i}_bd.0", co)
        s += _conv(f"{fe}.reduce.0", 64, 256, 1, 1, bias=False) + _bn_spec(f"{fe}.reduce.1", 64)
        s += _lin("dense", 2, 64)
    else:
        raise ValueError((arch, kind))
    return s


def make_weights(arch: str, kind: str, seed: int, n_experts: int = 3, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights in the reference's state_dict layout.

    Magnitudes follow torch's default init (uniform ±1/sqrt(fan_in)) so activations are in the
    range the reference sees; norm affines are perturbed so they are exercised.  A given
    (arch, kind, seed) always yields the same tensors (CPU generator), which is what lets the
    golden fixtures store outputs only.
    """
    g = torch.Generator().manual_seed(0x5EED0000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape, role, fan_in in param_spec(arch, kind, n_experts):
        if role in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif role == "gamma":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif role == "beta":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif role == "unit":
            t = torch.randn(shape, generator=g, dtype=torch.float64)
            t = t / t.norm().clamp_min(SN_EPS)
        elif role == "rm":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif role == "rv":
            t = 1.0 + 0.2 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif role == "nbt":
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        else:
            raise ValueError(role)
        sd[name] = t.to(dtype)
    return sd


def perturb_expert(sd: Dict[str, torch.Tensor], expert: int, seed: int) -> "OrderedDict[str, torch.Tensor]":
    """Expert e's weights = base * (1 + 1e-2 N(0,1)) so experts differ."""
    if expert == 0:
        return OrderedDict((k, v.clone()) for k, v in sd.items())
    g = torch.Generator().manual_seed(0xE0000 + 977 * seed + expert)
    out = OrderedDict()
    for k, v in sd.items():
        if v.dtype.is_floating_point and not k.endswith(("weight_u", "weight_v", "running_var")):
            out[k] = v * (1.0 + 1e-2 * torch.randn(v.shape, generator=g, dtype=v.dtype))
        else:
            out[k] = v.clone()
    return out


def make_batch(arch: str, B: int, seed: int):
    """Synthetic batch of SURVEY.md §8(d): cond, real_images(log1p photons), std, intensity, positions."""
    H, W = IMAGE_SHAPE[arch]
    g = torch.Generator().manual_seed(1234 + seed)
    p, vmax = (0.011, 765.0) if arch == "proton" else (0.039, 591.0)
    cond = torch.randn(B, 9, generator=g)
    hit = torch.rand(B, 1, H, W, generator=g) < p
    amp = torch.ceil(-20.0 * torch.log(torch.rand(B, 1, H, W, generator=g).clamp_min(1e-12))).clamp(max=vmax)
    photons = torch.where(hit, amp, torch.zeros(()))
    # guarantee >=1 photon per image (the reference filters on photon sum >= 1, data_filtering.ipynb:493)
    photons[:, 0, H // 2, W // 2] += 1.0
    real = torch.log1p(photons)
    intensity = torch.expm1(real).sum(dim=(2, 3))  # [B,1]
    std = torch.rand(B, 1, generator=g)
    flat = real.view(B, -1).argmax(dim=1)
    pos = torch.stack((flat // W, flat % W), dim=1).float()  # (row, col) = (max_x, max_y), train/utils.py:81-82
    return {"cond": cond, "real_images": real, "std": std, "intensity": intensity, "true_positions": pos}


def dropout_sites(arch: str, kind: str):
    """(site name, per-sample shape, p) for every nn.Dropout in module order."""
    if (arch, kind) == ("proton", "aux_reg"):  # proton/aux_reg.py:25,29
        return [("regressor.3", (128,), 0.3), ("regressor.7", (64,), 0.3)]
    if (arch, kind) == ("neutron", "generator"):  # neutron/generator.py:14,20,27,32,36
        return [("fc1.2", (256,), 0.2), ("fc2.2", (21632,), 0.2), ("conv_layers.2", (256, 24, 24), 0.2),
                ("conv_layers.7", (128, 46, 46), 0.2), ("conv_layers.11", (64, 45, 45), 0.2)]
    if (arch, kind) == ("neutron", "aux_reg"):  # neutron/aux_reg.py:17,25,33,41
        return [("conv1_bd.2", (32, 42, 42), 0.2), ("conv2_bd.2", (64, 19, 19), 0.2),
                ("conv3_bd.2", (128, 7, 17), 0.2), ("conv4_bd.2", (256, 1, 15), 0.2)]
    return []


def make_noise(arch: str, B: int, E: int, seed: int, noise_dim: int = 10):
    """All stochastic inputs of one train step, indexed by original sample index."""
    g = torch.Generator().manual_seed(0xA015E + seed)
    expo = torch.empty(B, E).exponential_(generator=g)
    n = {"expo": expo, "gumbel": -expo.log(),  # torch/nn/functional.py:2218-2224
         "z1": torch.randn(B, noise_dim, generator=g), "z2": torch.randn(B, noise_dim, generator=g)}
    for tag, kind in (("g1", "generator"), ("g2", "generator"), ("a", "aux_reg")):
        for site, shape, p in dropout_sites(arch, kind):
            n[f"drop.{tag}.{site}"] = (torch.rand((B,) + shape, generator=g) >= p).float()
    return n


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def gumbel_softmax_from_noise(logits, gumbel, tau):
    """F.gumbel_softmax(hard=False) with the Gumbel draw injected (torch/nn/functional.py:2218-2236)."""
    return ((logits + gumbel) / tau).softmax(dim=-1)


def router_forward(sd, cond, gumbel, tau=1.0):
    """RouterNetwork.forward (routers/router.py:21-26)."""
    h = cond
    for i in (0, 2, 4):
        h = lrelu(F.linear(h, sd[f"fc_layers.{i}.weight"], sd[f"fc_layers.{i}.bias"]))
    logits = F.linear(h, sd["fc_layers.6.weight"], sd["fc_layers.6.bias"])
    return gumbel_softmax_from_noise(logits, gumbel, tau), logits


def _dropout(x, mask, p, training):
    if not training or mask is None:
        return x
    return x * mask.view_as(x) / (1.0 - p)


def _bn(sd, name, x, training, update_running=True, momentum=0.1):
    """BatchNorm1d/2d (neutron/generator.py:13-35, neutron/aux_reg.py:15-47).  Train: biased batch variance
    normalises, running stats get the unbiased variance; eval: running stats."""
    w, b = sd[f"{name}.weight"], sd[f"{name}.bias"]
    rm, rv = sd[f"{name}.running_mean"], sd[f"{name}.running_var"]
    if training and update_running:
        sd[f"{name}.num_batches_tracked"] += 1
        return F.batch_norm(x, rm, rv, w, b, True, momentum, NORM_EPS)
    if training:
        return F.batch_norm(x, None, None, w, b, True, momentum, NORM_EPS)
    return F.batch_norm(x, rm, rv, w, b, False, momentum, NORM_EPS)


def generator_forward(arch, sd, noise, cond, training=True, drop=None, collect=None):
    """Generator.forward (proton/generator.py:46-52) / GeneratorNeutron.forward (neutron/generator.py:42-49).

    drop: dict site -> keep-mask [B_e,...] (neutron, train mode).  collect: optional dict that receives
    intermediate activations (used by the per-kernel parity tests)."""
    x = torch.cat((noise, cond), dim=1)
    keep = (lambda k, v: collect.__setitem__(k, v)) if collect is not None else (lambda k, v: None)
    if arch == "proton":
        x = F.linear(x, sd["fc1.0.weight"], sd["fc1.0.bias"])
        x = lrelu(F.layer_norm(x, (256,), sd["fc1.1.weight"], sd["fc1.1.bias"], NORM_EPS))
        keep("fc1", x)
        x = F.linear(x, sd["fc2.0.weight"], sd["fc2.0.bias"])
        keep("fc2_lin", x)
        x = lrelu(F.layer_norm(x, (92160,), sd["fc2.1.weight"], sd["fc2.1.bias"], NORM_EPS))
        keep("fc2", x)
        x = x.view(-1, 512, 18, 10)
        x = F.interpolate(x, scale_factor=(2, 2), mode="nearest")
        x = F.conv2d(x, sd["conv_layers.1.weight"], sd["conv_layers.1.bias"], padding=1)
        keep("conv1_lin", x)
        x = lrelu(F.group_norm(x, 32, sd["conv_layers.2.weight"], sd["conv_layers.2.bias"], NORM_EPS))
        keep("conv1", x)
        x = F.interpolate(x, size=(56, 30), mode="nearest")
        x = F.conv2d(x, sd["conv_layers.5.weight"], sd["conv_layers.5.bias"], padding=1)
        keep("conv2_lin", x)
        x = lrelu(F.group_norm(x, 32, sd["conv_layers.6.weight"], sd["conv_layers.6.bias"], NORM_EPS))
        keep("conv2", x)
        x = F.conv2d(x, sd["conv_layers.8.weight"], sd["conv_layers.8.bias"], padding=1)
        keep("conv3_lin", x)
        x = lrelu(F.group_norm(x, 32, sd["conv_layers.9.weight"], sd["conv_layers.9.bias"], NORM_EPS))
        keep("conv3", x)
        x = F.relu(F.conv2d(x, sd["conv_layers.11.weight"], sd["conv_layers.11.bias"], padding=1))
        return x
    drop = drop or {}
    x = F.linear(x, sd["fc1.0.weight"], sd["fc1.0.bias"])
    x = lrelu(_dropout(_bn(sd, "fc1.1", x, training), drop.get("fc1.2"), 0.2, training))
    x = F.linear(x, sd["fc2.0.weight"], sd["fc2.0.bias"])
    x = lrelu(_dropout(_bn(sd, "fc2.1", x, training), drop.get("fc2.2"), 0.2, training))
    x = x.view(-1, 128, 13, 13)
    x = F.interpolate(x, scale_factor=(2, 2), mode="nearest")
    x = F.conv2d(x, sd["conv_layers.0.weight"], sd["conv_layers.0.bias"])
    x = lrelu(_dropout(_bn(sd, "conv_layers.1", x, training), drop.get("conv_layers.2"), 0.2, training))
    x = F.interpolate(x, scale_factor=(2, 2), mode="nearest")
    x = F.conv2d(x, sd["conv_layers.5.weight"], sd["conv_layers.5.bias"])
    x = lrelu(_dropout(_bn(sd, "conv_layers.6", x, training), drop.get("conv_layers.7"), 0.2, training))
    x = F.conv2d(x, sd["conv_layers.9.weight"], sd["conv_layers.9.bias"])
    x = lrelu(_dropout(_bn(sd, "conv_layers.10", x, training), drop.get("conv_layers.11"), 0.2, training))
    x = F.relu(F.conv2d(x, sd["conv_layers.13.weight"], sd["conv_layers.13.bias"]))
    return x


def spectral_norm_weight(sd, prefix, training):
    """Hook-based torch.nn.utils.spectral_norm, 1 power iteration, dim 0, eps 1e-12
    (torch/nn/utils/spectral_norm.py:62-113).  In training mode u and v in ``sd`` are advanced IN PLACE."""
    w = sd[f"{prefix}.weight_orig"]
    u, v = sd[f"{prefix}.weight_u"], sd[f"{prefix}.weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            nv = torch.mv(wm.t(), u)
            v.copy_(nv / nv.norm().clamp_min(SN_EPS))
            nu = torch.mv(wm, v)
            u.copy_(nu / nu.norm().clamp_min(SN_EPS))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def discriminator_forward(arch, sd, img, cond, training=True, collect=None):
    """Discriminator.forward (proton/discriminator.py:148-155) / DiscriminatorNeutron.forward
    (neutron/discriminator.py:41-48).  Returns (score [B,1], latent [B,64])."""
    keep = (lambda k, v: collect.__setitem__(k, v)) if collect is not None else (lambda k, v: None)
    w = spectral_norm_weight(sd, "conv_layers.0", training)
    x = F.conv2d(img, w, sd["conv_layers.0.bias"])
    x = lrelu(F.group_norm(x, 8, sd["conv_layers.1.weight"], sd["conv_layers.1.bias"], NORM_EPS))
    x = F.max_pool2d(x, (2, 2))
    keep("pool1", x)
    w = spectral_norm_weight(sd, "conv_layers.4", training)
    x = F.conv2d(x, w, sd["conv_layers.4.bias"])
    x = lrelu(F.group_norm(x, 8, sd["conv_layers.5.weight"], sd["conv_layers.5.bias"], NORM_EPS))
    x = F.max_pool2d(x, (2, 1) if arch == "proton" else (2, 2))
    keep("pool2", x)
    x = torch.cat((x.reshape(x.size(0), -1), cond), dim=1)
    w = spectral_norm_weight(sd, "fc1.0", training)
    x = lrelu(F.layer_norm(F.linear(x, w, sd["fc1.0.bias"]), (128,), sd["fc1.1.weight"], sd["fc1.1.bias"], NORM_EPS))
    keep("fc1", x)
    w = spectral_norm_weight(sd, "fc2.0", training)
    latent = lrelu(F.layer_norm(F.linear(x, w, sd["fc2.0.bias"]), (64,), sd["fc2.1.weight"], sd["fc2.1.bias"], NORM_EPS))
    w = spectral_norm_weight(sd, "fc3", training)
    out = F.linear(latent, w, sd["fc3.bias"])
    return out, latent


def _gn_groups(C, groups=32):
    """Norm2d (proton/aux_reg.py:48-53)."""
    g = min(groups, C)
    while C % g != 0 and g > 1:
        g -= 1
    return g


def _res_block(sd, p, x, stride):
    """ResidualBlock.forward (proton/aux_reg.py:99-131); kernel 5, padding 2."""
    co = sd[f"{p}.conv1.0.weight"].shape[0]
    g = _gn_groups(co)
    out = F.conv2d(x, sd[f"{p}.conv1.0.weight"], sd[f"{p}.conv1.0.bias"], stride=stride, padding=2)
    out = F.relu(F.group_norm(out, g, sd[f"{p}.conv1.1.weight"], sd[f"{p}.conv1.1.bias"], NORM_EPS))
    out = F.conv2d(out, sd[f"{p}.conv2.0.weight"], sd[f"{p}.conv2.0.bias"], padding=2)
    out = F.group_norm(out, g, sd[f"{p}.conv2.1.weight"], sd[f"{p}.conv2.1.bias"], NORM_EPS)
    idn = F.conv2d(x, sd[f"{p}.downsample.0.weight"], sd[f"{p}.downsample.0.bias"], stride=stride)
    idn = F.group_norm(idn, g, sd[f"{p}.downsample.1.weight"], sd[f"{p}.downsample.1.bias"], NORM_EPS)
    return F.relu(out + idn)


def aux_forward(arch, sd, img, training=True, drop=None, collect=None):
    """AuxReg.forward (proton/aux_reg.py:33-40,84-96) / AuxRegNeutron.forward (neutron/aux_reg.py:51-59,76-80)."""
    drop = drop or {}
    keep = (lambda k, v: collect.__setitem__(k, v)) if collect is not None else (lambda k, v: None)
    fe = "feature_extractor"
    if arch == "proton":
        x = F.conv2d(img, sd[f"{fe}.conv1.0.weight"], sd[f"{fe}.conv1.0.bias"], stride=2, padding=1)
        x = F.relu(F.group_norm(x, 8, sd[f"{fe}.conv1.1.weight"], sd[f"{fe}.conv1.1.bias"], NORM_EPS))
        x = F.max_pool2d(x, 2, stride=1)
        keep("pool1", x)
        x = _res_block(sd, f"{fe}.res1", x, 2)
        keep("res1", x)
        x = F.max_pool2d(x, 2, stride=1)
        x = _res_block(sd, f"{fe}.res2", x, 2)
        x = F.max_pool2d(x, 2, stride=1)
        feat = x.mean((2, 3))
        keep("feat", feat)
        x = F.linear(feat, sd["regressor.0.weight"], sd["regressor.0.bias"])
        x = lrelu(F.layer_norm(x, (128,), sd["regressor.1.weight"], sd["regressor.1.bias"], NORM_EPS))
        x = _dropout(x, drop.get("regressor.3"), 0.3, training)
        x = F.linear(x, sd["regressor.4.weight"], sd["regressor.4.bias"])
        x = lrelu(F.layer_norm(x, (64,), sd["regressor.5.weight"], sd["regressor.5.bias"], NORM_EPS))
        x = _dropout(x, drop.get("regressor.7"), 0.3, training)
        return F.linear(x, sd["regressor.8.weight"], sd["regressor.8.bias"])
    x = img
    pools = {1: (2, 2), 2: (2, 1), 3: (2, 1)}
    for i in (1, 2, 3, 4):
        x = F.conv2d(x, sd[f"{fe}.conv{i}.weight"], sd[f"{fe}.conv{i}.bias"])
        x = lrelu(_bn(sd, f"{fe}.conv{i}_bd.0", x, training))
        x = _dropout(x, drop.get(f"conv{i}_bd.2"), 0.2, training)
        if i in pools:
            x = F.max_pool2d(x, pools[i])
    x = F.conv2d(x, sd[f"{fe}.reduce.0.weight"])
    x = lrelu(_bn(sd, f"{fe}.reduce.1", x, training))
    feat = x.mean((2, 3))
    return F.linear(feat, sd["dense.weight"], sd["dense.bias"])


# --------------------------------------------------------------------------------------
# loss tails
# --------------------------------------------------------------------------------------
def regressor_loss(real_coords, fake_coords):
    """AuxReg.regressor_loss (proton/aux_reg.py:42-45; neutron/aux_reg.py:70-74)."""
    d = fake_coords - real_coords
    return torch.mean(d + F.softplus(-2.0 * d) - math.log(2.0))


def sdi_gan_regularization(lat1, lat2, z1, z2, std, di_strength):
    """MoEWrapper.sdi_gan_regularization (models/moe.py:573-588), including its [B_e,1]/[B_e] broadcast."""
    a = torch.mean(torch.abs(lat1 - lat2), dim=1)
    n = torch.mean(torch.abs(z1 - z2), dim=1)
    div = a / (n + 1e-5)
    div_loss = std / (div + 1e-5)  # std [B_e,1], div [B_e]  -> [B_e,B_e]
    return torch.mean(std) * torch.mean(div_loss) * di_strength


def intensity_regularization(img, intensity, in_strength):
    """MoEWrapper.intensity_regularization (models/moe.py:590-642)."""
    s = torch.sum(torch.exp(img) - 1, dim=[2, 3])  # [B_e,1]
    std_i, mean_i = s.std(), s.mean()
    mae = F.l1_loss(s, intensity.view(-1, 1)) * in_strength
    return mae, s, std_i, mean_i


def expert_distribution_loss(gates, feats, lambda_reg=0.1):
    """calculate_expert_distribution_loss (train/utils.py:372-395)."""
    pd = torch.cdist(feats, feats, p=2)
    gs = gates @ gates.T
    return lambda_reg * torch.sum(gs * pd) / gs.size(0)


def utilization_entropy(gates_soft, strength):
    """calculate_expert_utilization_entropy (train/utils.py:398-419)."""
    p = gates_soft.mean(dim=0)
    return -(p * torch.log(p + 1e-9)).sum(dim=-1) * strength


def adaptive_load_balancing_loss(routing_scores, alb_strength, eps=1e-6):
    """calculate_adaptive_load_balancing_loss (train/utils.py:623-642)."""
    return torch.exp(1.0 / (routing_scores + eps)).mean() * alb_strength


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; training_setup.py:20-40)
# --------------------------------------------------------------------------------------
class AdamState:
    def __init__(self, params: Dict[str, torch.Tensor], lr: float):
        self.lr, self.step = lr, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    @torch.no_grad()
    def apply(self, params, grads, b1=0.9, b2=0.999, eps=1e-8):
        """Single-tensor torch Adam (torch/optim/adam.py _single_tensor_adam, non-capturable)."""
        if all(g is None for g in grads.values()):
            return
        self.step += 1
        bc1 = 1 - b1 ** self.step
        bc2_sqrt = math.sqrt(1 - b2 ** self.step)
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].lerp_(g, 1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / bc2_sqrt).add_(eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


# --------------------------------------------------------------------------------------
# the training step
# --------------------------------------------------------------------------------------
def trainable(sd):
    return {k: v for k, v in sd.items()
            if v.dtype.is_floating_point and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))}


class OracleState:
    """Weights + Adam state of one MoE system (MoEWrapper.__init__, models/moe.py:24-50 +
    setup_optimizers, train/training_setup.py:12-41)."""

    def __init__(self, arch, gens, discs, auxs, router, cfg):
        self.arch, self.cfg, self.E = arch, cfg, len(gens)
        self.gens, self.discs, self.auxs, self.router = gens, discs, auxs, router
        m = cfg["model"]
        self.opt_g = [AdamState(trainable(s), m["generator"]["lr_g"]) for s in gens]
        self.opt_d = [AdamState(trainable(s), m["discriminator"]["lr_d"]) for s in discs]
        self.opt_a = [AdamState(trainable(s), m["aux_reg"]["lr_a"]) for s in auxs]
        self.opt_r = AdamState(trainable(router), m["router"]["lr_r"])


def make_state(arch, E, seed, cfg, identical_experts=False):
    g0 = make_weights(arch, "generator", seed)
    d0 = make_weights(arch, "discriminator", seed + 1)
    a0 = make_weights(arch, "aux_reg", seed + 2)
    r = make_weights(arch, "router", seed + 3, n_experts=E)
    pick = (lambda sd, e: perturb_expert(sd, 0, seed)) if identical_experts else (lambda sd, e: perturb_expert(sd, e, seed))
    return OracleState(arch, [pick(g0, e) for e in range(E)], [pick(d0, e) for e in range(E)],
                       [pick(a0, e) for e in range(E)], r, cfg)


DEFAULT_CFG = {  # expertsim/config/default.yaml:8-35 with diff_strength=1e-6
    "model": {
        "architecture": "proton", "n_experts": 3, "noise_dim": 10, "cond_dim": 9,
        "generator": {"lr_g": 1e-4, "di_strength": 1e-1, "in_strength": 1e-3},
        "discriminator": {"lr_d": 1e-5},
        "aux_reg": {"lr_a": 1e-4, "strength": 1e-3},
        "router": {"version": "router_v1", "lr_r": 1e-4, "ed_strength": 0, "gan_strength": 1e-1,
                   "diff_strength": 1e-6, "util_strength": 0, "alb_strength": 1e-5,
                   "stop_router_training_epoch": 40, "alpha": 60, "min_weight": 0.2,
                   "tau_start": 1.2, "tau_min": 0.8, "tau_decay": 0.985},
    },
}


def router_tau(rcfg, epoch):
    """models/moe.py:62-74."""
    return max(rcfg["tau_min"], rcfg["tau_start"] * (rcfg["tau_decay"] ** epoch))


def route(gates_soft, E):
    """models/moe.py:97-103,123: argmax (first max), bincount, ascending per-expert index lists."""
    idx = gates_soft.argmax(dim=1)
    counts = torch.bincount(idx, minlength=E)
    masks = [(idx == i).nonzero(as_tuple=True)[0] for i in range(E)]
    return idx, counts, masks


def _with_grad(sd):
    out = {}
    for k, v in sd.items():
        if k in trainable(sd):
            out[k] = v.detach().clone().requires_grad_(True)
        else:
            out[k] = v  # buffers are shared so in-place updates (u, v, running stats) persist
    return out


def _grads(live, names):
    return {k: live[k].grad for k in names}


def train_step(st: OracleState, batch, noise, epoch=0, collect=None):
    """MoEWrapper.train_step (models/moe.py:52-504) with injected noise.  Mutates ``st`` (weights, buffers,
    Adam state) and returns (metrics dict of python floats, aux dict with idx/counts/images)."""
    cfg, E, arch = st.cfg["model"], st.E, st.arch
    rc = cfg["router"]
    cond, real = batch["cond"], batch["real_images"]
    pos, std, inten = batch["true_positions"], batch["std"], batch["intensity"]
    B = cond.size(0)
    tau = router_tau(rc, epoch)

    r_live = _with_grad(st.router)
    gates_soft, logits = router_forward(r_live, cond, noise["gumbel"], tau)
    idx, counts, masks = route(gates_soft, E)
    counts_adj = counts.to(real.dtype) / B
    gates = F.one_hot(idx, num_classes=E).float() + (gates_soft - gates_soft.detach())

    gen_losses, disc_losses = [], []
    div_l, aux_l, int_l = [0.0] * E, [0.0] * E, [0.0] * E
    mean_int, std_int = [], []
    mean_int_batch = torch.zeros(B, 1)
    aux_out = {"idx": idx.clone(), "counts": counts.clone(), "fake1": {}, "fake2": {}, "logits": logits.detach().clone(),
               "gates_soft": gates_soft.detach().clone()}

    def drops(tag, kind, mask):
        return {site: noise[f"drop.{tag}.{site}"][mask] for site, _, _ in dropout_sites(arch, kind)}

    for i in range(E):
        mask = masks[i]
        B_e = mask.numel()
        if B_e <= 1:  # models/moe.py:126-135
            gen_losses.append(torch.tensor(0.0))
            disc_losses.append(torch.tensor(0.0))
            mean_int.append(torch.tensor(0.0))
            std_int.append(torch.tensor(0.0))
            continue
        g_live, d_live, a_live = _with_grad(st.gens[i]), _with_grad(st.discs[i]), _with_grad(st.auxs[i])
        c_e, z1, z2 = cond[mask], noise["z1"][mask], noise["z2"][mask]
        fake = generator_forward(arch, g_live, z1, c_e, True, drops("g1", "generator", mask))  # moe.py:143-145
        w = float(counts_adj[i])

        # ---- discriminator_train_step (models/moe.py:506-527)
        real_out, _ = discriminator_forward(arch, d_live, real[mask], c_e, True)
        fake_out, _ = discriminator_forward(arch, d_live, fake.detach(), c_e, True)
        d_loss = (F.relu(1.0 - real_out).mean() + F.relu(1.0 + fake_out).mean()) * w
        d_loss.backward()
        d_names = list(trainable(st.discs[i]).keys())
        if collect is not None:
            collect[f"d_grads_{i}"] = {k: d_live[k].grad.clone() for k in d_names}
        st.opt_d[i].apply({k: st.discs[i][k] for k in d_names}, _grads(d_live, d_names))
        disc_losses.append(d_loss.detach())

        # ---- generator_train_step (models/moe.py:529-571): D now has its UPDATED weights
        d_live2 = _with_grad(st.discs[i])
        fake2 = generator_forward(arch, g_live, z2, c_e, True, drops("g2", "generator", mask))
        out1, lat1 = discriminator_forward(arch, d_live2, fake, c_e, True)
        out2, lat2 = discriminator_forward(arch, d_live2, fake2, c_e, True)
        g_loss = -out1.mean()
        div = sdi_gan_regularization(lat1, lat2, z1, z2, std[mask], cfg["generator"]["di_strength"])
        il, sums, s_std, s_mean = intensity_regularization(fake, inten[mask], cfg["generator"]["in_strength"])
        coords = aux_forward(arch, a_live, fake, True, drops("a", "aux_reg", mask))
        al = regressor_loss(pos[mask], coords) * cfg["aux_reg"]["strength"]
        g_loss = (g_loss + div + il + al) * w
        if collect is not None:
            fake.retain_grad()
            fake2.retain_grad()
        g_loss.backward()
        g_names, a_names = list(trainable(st.gens[i]).keys()), list(trainable(st.auxs[i]).keys())
        if collect is not None:
            collect[f"g_grads_{i}"] = {k: g_live[k].grad.clone() for k in g_names}
            collect[f"a_grads_{i}"] = {k: a_live[k].grad.clone() for k in a_names}
            collect[f"dfake1_{i}"], collect[f"dfake2_{i}"] = fake.grad.clone(), fake2.grad.clone()
            collect[f"lat1_{i}"], collect[f"lat2_{i}"] = lat1.detach().clone(), lat2.detach().clone()
            collect[f"coords_{i}"] = coords.detach().clone()
        st.opt_g[i].apply({k: st.gens[i][k] for k in g_names}, _grads(g_live, g_names))
        st.opt_a[i].apply({k: st.auxs[i][k] for k in a_names}, _grads(a_live, a_names))

        mean_int_batch[mask] = sums.detach()
        mean_int.append(s_mean.detach())
        std_int.append(s_std.detach())
        gen_losses.append(g_loss.detach())
        div_l[i], int_l[i], aux_l[i] = float(div.detach()), float(il.detach()), float(al.detach())
        aux_out["fake1"][i], aux_out["fake2"][i] = fake.detach(), fake2.detach()

    zero = torch.tensor(0.0)
    if E > 1:  # models/moe.py:213-442
        gan = torch.stack(gen_losses).mean() * rc["gan_strength"]
        ent = -1 * utilization_entropy(gates_soft, rc["util_strength"]) if rc["util_strength"] != 0 else zero
        ed = (expert_distribution_loss(gates, mean_int_batch) if rc["ed_strength"] != 0 else zero) * rc["ed_strength"]
        if rc["diff_strength"] != 0:
            di = sum(F.l1_loss(mean_int[i].unsqueeze(0), mean_int[j].unsqueeze(0))
                     for i, j in combinations(range(E), 2)) * rc["diff_strength"]
        else:
            di = zero
        diff = -di * rc["diff_strength"]
        alb = adaptive_load_balancing_loss(gates_soft.sum(dim=0), rc["alb_strength"]) if rc["alb_strength"] != 0 else zero
        alpha = min(max((epoch - 0) / (rc["alpha"] - 0), 0.0), 1.0)
        dec_w = rc["min_weight"] + (1.0 - rc["min_weight"]) * alpha
        r_loss = ed + gan + diff + ent + dec_w * alb
        if epoch < rc["stop_router_training_epoch"]:
            r_loss.backward()
            r_names = list(trainable(st.router).keys())
            if collect is not None:
                collect["r_grads"] = {k: r_live[k].grad.clone() for k in r_names}
            st.opt_r.apply({k: st.router[k] for k in r_names}, _grads(r_live, r_names))
        else:
            r_loss = zero
    else:
        gan = r_loss = ed = diff = ent = alb = zero

    f = lambda t: float(t.detach()) if isinstance(t, torch.Tensor) else float(t)
    metrics = {
        "gen_loss": f(torch.stack(gen_losses).mean()), "disc_loss": f(torch.stack(disc_losses).mean()),
        "div_loss": sum(div_l) / E, "intensity_loss": sum(int_l) / E, "aux_reg_loss": sum(aux_l) / E,
        "router_loss": f(r_loss), "expert_distribution_loss": f(ed), "differentiation_loss": f(diff),
        "expert_entropy_loss": f(ent), "adaptive_load_balancing_loss": f(alb), "gan_loss": f(gan),
    }
    for i in range(E):
        metrics.update({f"gen_loss_{i}": f(gen_losses[i]), f"disc_loss_{i}": f(disc_losses[i]),
                        f"div_loss_experts_{i}": div_l[i], f"intensity_loss_experts_{i}": int_l[i],
                        f"aux_reg_loss_experts_{i}": aux_l[i], f"std_intensities_experts_{i}": f(std_int[i]),
                        f"mean_intensities_experts_{i}": f(mean_int[i]),
                        f"n_choosen_experts_mean_epoch_{i}": float(counts[i])})
    return metrics, aux_out


# --------------------------------------------------------------------------------------
# batch inference
# --------------------------------------------------------------------------------------
@torch.no_grad()
def generate(arch, sd, noise, cond, batch_size=64):
    """get_predictions_from_generator_results (train/utils.py:179-205): eval-mode generator, expm1, float64."""
    H, W = IMAGE_SHAPE[arch]
    n = cond.shape[0]
    out = torch.zeros(n, H, W, dtype=torch.float64)
    raw = torch.zeros(n, H, W, dtype=torch.float64)
    for s in range(0, n, batch_size):
        r = generator_forward(arch, sd, noise[s:s + batch_size], cond[s:s + batch_size], training=False)
        raw[s:s + batch_size] = r.reshape(-1, H, W).double()
        out[s:s + batch_size] = torch.expm1(r).reshape(-1, H, W).double()  # np.expm1 on fp32 then stored to f64
    return out, raw


@torch.no_grad()
def moe_generate(st: OracleState, cond, gumbel, z):
    """Routing of MoEWrapper.evaluate (models/moe.py:650-653; tau=1, gumbel still sampled) followed by
    per-expert generation; returns showers in ORIGINAL sample order."""
    gates, _ = router_forward(st.router, cond, gumbel, 1.0)
    idx, counts, masks = route(gates, st.E)
    H, W = IMAGE_SHAPE[st.arch]
    out = torch.zeros(cond.shape[0], H, W, dtype=torch.float64)
    for e, m in enumerate(masks):
        if m.numel():
            out[m] = generate(st.arch, st.gens[e], z[m], cond[m])[0]
    return out, idx, counts


# --------------------------------------------------------------------------------------
# evaluation metric helpers (train/utils.py:18-78) — "next" row §8(f)1, restated for completeness
# --------------------------------------------------------------------------------------
def channel_masks(H, W):
    ii, jj = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    chk = ((ii % 2) != (jj % 2)).double()  # pattern[[0,1],[1,0]]
    m5 = 1.0 - chk
    top, left = ii < H // 2, jj < W // 2
    m1 = chk * (~top & left)
    m2 = chk * (~top & ~left)
    m3 = chk * (top & left)
    m4 = chk * (top & ~left)
    return m1, m2, m3, m4, m5


def sum_channels(data):
    """sum_channels_parallel (train/utils.py:63-78): data [N,H,W] -> [N,5]."""
    ms = channel_masks(data.shape[1], data.shape[2])
    return torch.stack([(data.double() * m).sum(dim=(1, 2)) for m in ms], dim=1)
