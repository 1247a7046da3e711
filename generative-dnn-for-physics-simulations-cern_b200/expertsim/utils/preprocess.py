"""Offline preprocessing of the training set on the device: the two notebook steps that produce
inputs of the training step.

* ``max_coordinates``  — notebooks/calculate_and_analysis_of_max_coordinates.ipynb cell 6 looping
  ``get_max_value_image_coordinates`` (expertsim/train/utils.py:81-82) over the showers -> ``positions`` (row, col).
* ``condition_group_std`` — notebooks/calculating_diversity_for_data.ipynb cells 16-23:
  ``data_all.groupby(CONDITIONAL_COLS).transform(np.std).sum(axis=1) / max`` -> the per-sample ``std`` column that weights
  the SDI diversity loss.

The arithmetic runs in csrc/preprocess.cu through the C-ABI; torch only finds the groups of identical conditioning rows
(index plumbing: unique / stable argsort / bincount) and owns the memory.  No CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib as L


def _images(images: torch.Tensor):
    if not images.is_cuda:
        raise RuntimeError("preprocessing runs on the device: move the images to CUDA first (there is no CPU fallback)")
    n = images.shape[0]
    H, W = images.shape[-2], images.shape[-1]
    return images.reshape(n, H * W).to(torch.float32).contiguous(), n, H, W


def max_coordinates(images: torch.Tensor, as_float: bool = False) -> torch.Tensor:
    """images [N,H,W] (or [N,1,H,W]) -> [N,2] (row, col) of the first maximum in row-major order; int32, or float32 like the
    ``positions`` targets of the loader (train/utils.py:81-82, data_transformations.py:194-195)."""
    img, n, H, W = _images(images)
    out_i = None if as_float else torch.empty(n, 2, dtype=torch.int32, device=img.device)
    out_f = torch.empty(n, 2, device=img.device) if as_float else None
    L.call("es_argmax_coords", img, n, H, W, out_i, out_f)
    return out_f if as_float else out_i


def condition_group_std(cond: torch.Tensor, images: torch.Tensor, return_groups: bool = False):
    """cond [N,K] conditioning rows, images [N,H,W] -> std [N] in (0, 1]: showers with identical conditioning rows form a
    group; per group and pixel the population standard deviation over the group; summed over pixels; divided by the
    largest sum (calculating_diversity_for_data.ipynb cells 16-23)."""
    img, n, H, W = _images(images)
    cond = cond.to(img.device)
    _, gid = torch.unique(cond.reshape(n, -1), dim=0, return_inverse=True)
    G = int(gid.max().item()) + 1
    order = torch.argsort(gid, stable=True).to(torch.int32).contiguous()
    seg = torch.zeros(G + 1, dtype=torch.int32, device=img.device)
    seg[1:] = torch.bincount(gid, minlength=G).cumsum(0).to(torch.int32)
    gid32 = gid.to(torch.int32).contiguous()
    sums = torch.empty(G, dtype=torch.float64, device=img.device)
    out = torch.empty(n, device=img.device)
    L.call("es_group_pixel_std", img, n, H * W, order, seg, G, gid32, sums, out)
    return (out, gid32, sums) if return_groups else out
