"""Helpers shared by the GPU parity tests (they call the CUDA path through the C-ABI binding)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from expertsim import _lib as L  # noqa: E402

DEV = "cuda"
LOG = os.path.join(ROOT, "gpurun_out", "parity.log")


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as f:
        f.write(msg + "\n")
    print(msg)


def groups(counts, slots=None, two_pass=False, min_rows=0):
    """Device group table for consecutive groups with the given row counts."""
    rows, off = [], 0
    for i, c in enumerate(counts):
        act = c if c >= min_rows else 0
        s = slots[i] if slots is not None else i
        if two_pass:
            rows.append([2 * off, 2 * act, s, act])
        else:
            rows.append([off, act, s, act])
        off += c
    return torch.tensor(rows, dtype=torch.int32, device=DEV), off


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max())


def check(name, got, want, rtol_l2, atol_max=None):
    """relative L2 error bound (+ optional max-abs bound scaled by max|want|)."""
    r = rel_err(got, want)
    m = max_err(got, want)
    scale = float(want.detach().abs().max())
    log(f"{name:58s} relL2={r:.3e} maxabs={m:.3e} (max|ref|={scale:.3e})")
    assert torch.isfinite(got.detach().float()).all(), f"{name}: non-finite output"
    assert r <= rtol_l2, f"{name}: relative L2 error {r:.3e} > {rtol_l2:.1e}"
    if atol_max is not None:
        assert m <= atol_max * max(scale, 1e-30), f"{name}: max abs error {m:.3e} > {atol_max:.1e} * {scale:.3e}"


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def cuda(t, dtype=None):
    t = t.to(DEV)
    return t.to(dtype).contiguous() if dtype is not None else t.contiguous()
