"""Batch-inference helper, evaluation metric and the router-loss helper functions.  Drop-in for the live parts of
expertsim/train/utils.py of the reference; plotting helpers (matplotlib/seaborn) are out of scope."""
from typing import List

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------- batch inference
def get_predictions_from_generator_results(batch_size, num_samples, noise_dim, device, y_test, generator,
                                           shape_images=(56, 30), input_noise=None):
    """Reference train/utils.py:179-205: generate ``num_samples`` showers from one expert generator (eval mode), return
    (expm1-transformed, raw) float64 numpy arrays [N,H,W].  The reference copies every batch of 64 to the host and
    applies np.expm1 there; here the generator runs in large chunks, expm1 + float64 widening happen on the device
    (es_expm1_scatter) and each chunk is copied to the host once."""
    from .. import _lib as L
    H, W = shape_images
    res = np.zeros((num_samples, H, W))
    raw = np.zeros((num_samples, H, W))
    generator.eval()
    chunk = max(int(batch_size), 8192)
    for s in range(0, num_samples, chunk):
        e = min(s + chunk, num_samples)
        z = input_noise[s:e] if input_noise is not None else torch.randn(e - s, noise_dim, device=device)
        with torch.no_grad():
            img = generator(z.to(device), y_test[s:e].to(device)).reshape(e - s, H * W).contiguous()
        o64 = torch.empty(e - s, H, W, dtype=torch.float64, device=img.device)
        L.call("es_expm1_scatter", img, None, e - s, H * W, o64, None)
        res[s:e] = o64.cpu().numpy()
        raw[s:e] = img.view(e - s, H, W).double().cpu().numpy()
    return res, raw


# ---------------------------------------------------------------------------------------------------- evaluation metric
def get_channel_masks(input_array: np.ndarray):
    """Five masks (4 checkerboard quadrants + the complementary checkerboard); reference train/utils.py:18-60."""
    n, m = input_array.shape
    ii, jj = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    chk = ((ii % 2) != (jj % 2)).astype(input_array.dtype)
    mask5 = 1 - chk
    top, left = ii < n // 2, jj < m // 2
    return chk * (~top & left), chk * (~top & ~left), chk * (top & left), chk * (top & ~left), mask5


def sum_channels_parallel(data: np.ndarray):
    """[x,N,M] images -> iterable of 5 channel sums per image (reference train/utils.py:63-78)."""
    masks = get_channel_masks(data[0])
    return zip(*[(data * m).sum(axis=1).sum(axis=1) for m in masks])


def channel_sums_device(img: torch.Tensor, H: int, W: int, apply_expm1: bool = True) -> torch.Tensor:
    """[N, H*W] fp32 CUDA images -> [N, 5] fp64 channel sums on the device (expm1 fused; same masks as get_channel_masks)."""
    from .. import _lib as L
    n = img.shape[0]
    out = torch.empty(n, 5, dtype=torch.float64, device=img.device)
    L.call("es_channel_sums", img.reshape(n, H * W).contiguous(), H, W, n, int(apply_expm1), out)
    return out


def ws_device(ch_a_sorted: torch.Tensor, ch_b: torch.Tensor) -> torch.Tensor:
    """1-D Wasserstein distance per channel between two equally sized samples ([n,5] fp64; ``ch_a_sorted`` already sorted per
    column): for equal sizes scipy.stats.wasserstein_distance is mean |sort(a) - sort(b)|.  -> [5] fp64 on the device."""
    from .. import _lib as L
    assert ch_a_sorted.shape == ch_b.shape
    out = torch.empty(ch_b.shape[1], dtype=torch.float64, device=ch_b.device)
    L.call("es_w1_sorted", ch_a_sorted.contiguous(), ch_b.sort(dim=0).values.contiguous(), ch_b.shape[0], ch_b.shape[1], out)
    return out


def calculate_joint_ws_across_experts(n_calc, x_tests: List, y_tests: List, generators: List, ch_org, ch_org_expert,
                                      noise_dim, device, batch_size=64, n_experts=3, shape_images=(56, 30)):
    """Wasserstein distance between real and generated channel sums, overall and per expert, ``n_calc`` repetitions
    (reference train/utils.py:117-176)."""
    from scipy.stats import wasserstein_distance
    if len(x_tests) != len(y_tests) or len(x_tests) != len(generators):
        raise ValueError("Length of data is not the same")
    ws = np.zeros((n_calc, 5))
    ws_exp = np.zeros((n_calc, n_experts, 5))
    for j in range(n_calc):
        ch_all, ch_exp = [], []
        for gi, gen in enumerate(generators):
            n = x_tests[gi].shape[0]
            if n == 0:
                ch_exp.append(np.zeros((0, 5)))
                continue
            res, _ = get_predictions_from_generator_results(batch_size, n, noise_dim, device, y_tests[gi], gen, shape_images)
            ch = np.array(list(sum_channels_parallel(res)))
            ch_exp.append(ch)
            ch_all.extend(ch)
        ch_all = np.array(ch_all)
        for i in range(5):
            ws[j][i] = wasserstein_distance(ch_org[:, i], ch_all[:, i])
            for e in range(len(generators)):
                if ch_exp[e].shape[0] == 0 or ch_org_expert[e].shape[0] == 0:
                    continue
                ws_exp[j][e][i] = wasserstein_distance(ch_org_expert[e][:, i], ch_exp[e][:, i])
    runs = ws.mean(axis=1)
    runs_exp = ws_exp.mean(axis=2)
    return runs.mean(), runs.std(), runs_exp.mean(axis=0), runs_exp.std(axis=0)


# ---------------------------------------------------------------------------------------------------- router losses
# Functional forms of the helpers at reference train/utils.py:372-419,623-642.  The training step evaluates these
# inside the router kernels (es_router_bwd / es_router_ed_loss); the functions are kept for callers that use them
# stand-alone on tensors.
def calculate_expert_distribution_loss(gating_probs, features, lambda_reg=0.1):
    pd = torch.cdist(features, features, p=2)
    gs = torch.mm(gating_probs, gating_probs.T)
    return lambda_reg * torch.sum(gs * pd) / gs.size(0)


def calculate_entropy(p):
    return -(p * torch.log(p + 1e-9)).sum(dim=-1)


def calculate_expert_utilization_entropy(gating_probs, ENT_STRENGTH=0.1):
    return calculate_entropy(gating_probs.mean(dim=0)) * ENT_STRENGTH


def calculate_adaptive_load_balancing_loss(routing_scores, alb_strength=1e-2, eps=1e-6):
    return torch.exp(1.0 / (routing_scores + eps)).mean() * alb_strength
