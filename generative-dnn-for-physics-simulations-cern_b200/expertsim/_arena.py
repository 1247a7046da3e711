"""Parameter arenas: all experts' parameters of one network kind live in ONE flat fp32 tensor [E, n] (plus matching
gradient / Adam-moment tensors), so that every grouped kernel addresses expert e's copy as ``base + e * n`` and a single
fused Adam launch updates every expert.  The ``nn.Parameter`` objects the reference's API exposes
(``moe.generators[i].parameters()``, ``state_dict()`` with the reference's key names — SURVEY.md §8b) are VIEWS into
the arena, so checkpoints and externally-built optimizers keep working.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
from torch import nn

# (name, shape, kind, fan_in);  kind: w/b = weight/bias (uniform +-1/sqrt(fan_in), torch's default Linear/Conv init),
# g/z = norm scale (ones) / shift (zeros), u = spectral-norm vector buffer, rm/rv/nbt = BatchNorm buffers
Spec = List[Tuple[str, tuple, str, int]]


def _lin(n, o, i, bias=True):
    return [(f"{n}.weight", (o, i), "w", i)] + ([(f"{n}.bias", (o,), "b", i)] if bias else [])


def _conv(n, co, ci, kh, kw, bias=True):
    f = ci * kh * kw
    return [(f"{n}.weight", (co, ci, kh, kw), "w", f)] + ([(f"{n}.bias", (co,), "b", f)] if bias else [])


def _aff(n, c):
    return [(f"{n}.weight", (c,), "g", 0), (f"{n}.bias", (c,), "z", 0)]


def _bn(n, c):
    return _aff(n, c) + [(f"{n}.running_mean", (c,), "rm", 0), (f"{n}.running_var", (c,), "rv", 0),
                         (f"{n}.num_batches_tracked", (), "nbt", 0)]


def _sn(n, shape):
    o, f = shape[0], int(math.prod(shape[1:]))
    return [(f"{n}.bias", (o,), "b", f), (f"{n}.weight_orig", shape, "w", f), (f"{n}.weight_u", (o,), "u", 0),
            (f"{n}.weight_v", (f,), "u", 0)]


def spec_for(arch: str, kind: str, n_experts: int = 3, cond_dim: int = 9, noise_dim: int = 10) -> Spec:
    """state_dict layout of the reference modules (measured)."""
    s: Spec = []
    if kind == "router":
        s += _lin("fc_layers.0", 128, cond_dim) + _lin("fc_layers.2", 64, 128) + _lin("fc_layers.4", 32, 64)
        s += _lin("fc_layers.6", n_experts, 32)
    elif (arch, kind) == ("proton", "generator"):
        s += _lin("fc1.0", 256, noise_dim + cond_dim) + _aff("fc1.1", 256)
        s += _lin("fc2.0", 92160, 256) + _aff("fc2.1", 92160)
        s += _conv("conv_layers.1", 256, 512, 4, 4) + _aff("conv_layers.2", 256)
        s += _conv("conv_layers.5", 128, 256, 4, 4) + _aff("conv_layers.6", 128)
        s += _conv("conv_layers.8", 64, 128, 3, 3) + _aff("conv_layers.9", 64)
        s += _conv("conv_layers.11", 1, 64, 2, 2)
    elif (arch, kind) == ("neutron", "generator"):
        s += _lin("fc1.0", 256, noise_dim + cond_dim) + _bn("fc1.1", 256)
        s += _lin("fc2.0", 21632, 256) + _bn("fc2.1", 21632)
        s += _conv("conv_layers.0", 256, 128, 3, 3) + _bn("conv_layers.1", 256)
        s += _conv("conv_layers.5", 128, 256, 3, 3) + _bn("conv_layers.6", 128)
        s += _conv("conv_layers.9", 64, 128, 2, 2) + _bn("conv_layers.10", 64)
        s += _conv("conv_layers.13", 1, 64, 2, 2)
    elif kind == "discriminator":
        flat = 2304 if arch == "proton" else 1296
        s += _sn("conv_layers.0", (32, 1, 3, 3)) + _aff("conv_layers.1", 32)
        s += _sn("conv_layers.4", (16, 32, 3, 3)) + _aff("conv_layers.5", 16)
        s += _sn("fc1.0", (128, flat + cond_dim)) + _aff("fc1.1", 128)
        s += _sn("fc2.0", (64, 128)) + _aff("fc2.1", 64)
        s += _sn("fc3", (1, 64))
    elif (arch, kind) == ("proton", "aux_reg"):
        fe = "feature_extractor"
        s += _conv(f"{fe}.conv1.0", 32, 1, 5, 5) + _aff(f"{fe}.conv1.1", 32)
        for blk, ci, co in (("res1", 32, 32), ("res2", 32, 64)):
            s += _conv(f"{fe}.{blk}.conv1.0", co, ci, 5, 5) + _aff(f"{fe}.{blk}.conv1.1", co)
            s += _conv(f"{fe}.{blk}.conv2.0", co, co, 5, 5) + _aff(f"{fe}.{blk}.conv2.1", co)
            s += _conv(f"{fe}.{blk}.downsample.0", co, ci, 1, 1) + _aff(f"{fe}.{blk}.downsample.1", co)
        s += _lin("regressor.0", 128, 64) + _aff("regressor.1", 128) + _lin("regressor.4", 64, 128)
        s += _aff("regressor.5", 64) + _lin("regressor.8", 2, 64)
    elif (arch, kind) == ("neutron", "aux_reg"):
        fe = "feature_extractor"
        for i, (ci, co) in enumerate(((1, 32), (32, 64), (64, 128), (128, 256)), start=1):
            s += _conv(f"{fe}.conv{i}", co, ci, 3, 3) + _bn(f"{fe}.conv{i}_bd.0", co)
        s += _conv(f"{fe}.reduce.0", 64, 256, 1, 1, bias=False) + _bn(f"{fe}.reduce.1", 64) + _lin("dense", 2, 64)
    else:
        raise ValueError(f"unknown network {arch}.{kind}")
    return s


TRAINABLE = ("w", "b", "g", "z")


def _round4(n):
    return (n + 3) & ~3


class Holder(nn.Module):
    """Anonymous container so dotted reference names ('fc1.0.weight') map onto real sub-modules; integer indexing
    works as on the reference's nn.Sequential blocks (``gen.fc2[0].weight``)."""

    def __getitem__(self, i):
        return self._modules[str(i)]


def register_by_spec(module: nn.Module, spec: Spec):
    """Create the nested parameter/buffer tree described by ``spec`` on ``module`` (reference key names)."""
    for name, shape, kind, fan_in in spec:
        *path, leaf = name.split(".")
        m = module
        for p in path:
            if p not in m._modules:
                m.add_module(p, Holder())
            m = m._modules[p]
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = torch.empty(shape).uniform_(-bound, bound)
        elif kind == "g":
            t = torch.ones(shape)
        elif kind in ("z", "rm"):
            t = torch.zeros(shape)
        elif kind == "rv":
            t = torch.ones(shape)
        elif kind == "u":
            t = torch.nn.functional.normalize(torch.randn(shape), dim=0, eps=1e-12)
        elif kind == "nbt":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        if kind in TRAINABLE:
            m.register_parameter(leaf, nn.Parameter(t))
        else:
            m.register_buffer(leaf, t)


def get_by_name(module: nn.Module, name: str):
    *path, leaf = name.split(".")
    m = module
    for p in path:
        m = m._modules[p]
    return m, leaf


class Arena:
    """Flat storage for the parameters (and float buffers) of E copies of one network kind."""

    def __init__(self, spec: Spec, n_slots: int, device):
        self.spec, self.E, self.device = spec, n_slots, torch.device(device)
        self.off: Dict[str, int] = OrderedDict()
        self.boff: Dict[str, int] = OrderedDict()
        self.shape: Dict[str, tuple] = {}
        self.ioff: Dict[str, int] = OrderedDict()     # int64 buffers (BatchNorm num_batches_tracked)
        n = nb = 0
        for name, shape, kind, _ in spec:
            self.shape[name] = shape
            numel = int(math.prod(shape)) if shape else 1
            if kind in TRAINABLE:
                self.off[name] = n
                n += _round4(numel)
            elif kind != "nbt":
                self.boff[name] = nb
                nb += _round4(numel)
            else:
                self.ioff[name] = len(self.ioff)
        self.n, self.nb, self.ni = n, max(nb, 4), max(len(self.ioff), 1)
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=self.device)
        self.P, self.G = z(n_slots, n), z(n_slots, n)
        self.M, self.V = z(n_slots, n), z(n_slots, n)
        self.steps = z(n_slots, dt=torch.int32)
        self.Bf = z(n_slots, self.nb)
        self.Bi = z(n_slots, self.ni, dt=torch.int64)
        self.modules: List[nn.Module] = []
        self.version = 0      # bumped whenever parameters change (optimizer step, load_state_dict)
        self.engine = None    # compute engine bound to this arena (created by the owner)

    # ---- raw addresses for the C-ABI (slot 0; slot e = + e * stride floats)
    def addr(self, name):
        return self.P.data_ptr() + 4 * self.off[name]

    def gaddr(self, name):
        return self.G.data_ptr() + 4 * self.off[name]

    def baddr(self, name):
        return self.Bf.data_ptr() + 4 * self.boff[name]

    def iaddr(self, name):
        return self.Bi.data_ptr() + 8 * self.ioff[name]

    def numel(self, name):
        return int(math.prod(self.shape[name])) if self.shape[name] else 1

    def view(self, store, name, slot):
        o = self.off[name] if store is not self.Bf else self.boff[name]
        return store[slot, o:o + self.numel(name)].view(self.shape[name])

    def adopt(self, module: nn.Module, slot: int):
        """Copy ``module``'s tensors into slot ``slot`` and re-point its Parameters / buffers at arena views."""
        with torch.no_grad():
            for name, shape, kind, _ in self.spec:
                holder, leaf = get_by_name(module, name)
                if kind in TRAINABLE:
                    p = holder._parameters[leaf]
                    v = self.view(self.P, name, slot)
                    v.copy_(p.data.to(self.device))
                    p.data = v
                    p.grad = None
                elif kind == "nbt":
                    v = self.Bi[slot, self.ioff[name]]
                    v.copy_(holder._buffers[leaf].to(self.device))
                    holder._buffers[leaf] = v
                else:
                    v = self.view(self.Bf, name, slot)
                    v.copy_(holder._buffers[leaf].to(self.device))
                    holder._buffers[leaf] = v
        while len(self.modules) <= slot:
            self.modules.append(None)
        self.modules[slot] = module
        module._arena, module._slot = self, slot

    def owns(self, module: nn.Module) -> bool:
        """True if the module's first parameter still aliases this arena (a later .to()/.cuda() would break that)."""
        name = next(iter(self.off))
        holder, leaf = get_by_name(module, name)
        return holder._parameters[leaf].data_ptr() == self.view(self.P, name, module._slot).data_ptr()

    def expose_grads(self, slots=None):
        """Point every Parameter's .grad at its slice of the gradient arena (for externally-built optimizers)."""
        for e, m in enumerate(self.modules):
            if m is None or (slots is not None and e not in slots):
                continue
            for name in self.off:
                holder, leaf = get_by_name(m, name)
                holder._parameters[leaf].grad = self.view(self.G, name, e)
