"""Batches for the training loop.  The reference reads proprietary GEANT4 pickles through pandas
(expertsim/utils/data_transformations.py:23-309; not shipped).  The 6-tuple batch contract of its loaders —
``(x, x_2, cond, std, intensity, positions)`` (data_transformations.py:269-271) — is kept; data here is either
synthetic with the measured statistics of the real set or tensors the caller provides."""
import math

import torch

IMAGE_SHAPE = {"proton": (56, 30), "neutron": (44, 44)}


def synthetic_showers(arch: str, n: int, seed: int = 0, device="cpu"):
    """cond ~ N(0,1) (the reference standardises the conditionals), log1p photon images with the measured pixel
    sparsity (1.1% proton / 3.9% neutron non-zero, exponential amplitudes clipped to the data maximum), photon-sum
    ``intensity``, min-max scaled ``std`` in [0,1], arg-max pixel ``positions`` (row, col)."""
    H, W = IMAGE_SHAPE[arch]
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    p, vmax = (0.011, 765.0) if arch == "proton" else (0.039, 591.0)
    cond = torch.randn(n, 9, generator=g, device=device)
    hit = torch.rand(n, H, W, generator=g, device=device) < p
    amp = torch.ceil(-20.0 * torch.log(torch.rand(n, H, W, generator=g, device=device).clamp_min(1e-12))).clamp(max=vmax)
    photons = torch.where(hit, amp, torch.zeros((), device=device))
    photons[:, H // 2, W // 2] += 1.0   # the reference keeps only showers with >= 1 photon
    x = torch.log1p(photons)
    intensity = photons.sum(dim=(1, 2)).unsqueeze(1)
    std = torch.rand(n, 1, generator=g, device=device)
    flat = x.view(n, -1).argmax(dim=1)
    pos = torch.stack((flat // W, flat % W), dim=1).float()
    return {"x": x, "cond": cond, "std": std, "intensity": intensity, "positions": pos}


class DeviceLoader:
    """Iterates 6-tuple batches from tensors that already live on one device (optionally a fresh permutation per
    epoch); under data parallelism rank r of N sees rows r::N of every global batch."""

    def __init__(self, data, batch_size, shuffle=True, rank=0, world=1, drop_last=True, seed=0):
        self.d, self.bs, self.shuffle, self.rank, self.world, self.drop_last = data, batch_size, shuffle, rank, world, drop_last
        self.n = data["x"].shape[0]
        self.epoch, self.seed = 0, seed

    def __len__(self):
        gb = self.bs * self.world
        return self.n // gb if self.drop_last else math.ceil(self.n / gb)

    def __iter__(self):
        dev = self.d["x"].device
        if self.shuffle:
            g = torch.Generator(device=dev).manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g, device=dev)
        else:
            order = torch.arange(self.n, device=dev)
        self.epoch += 1
        gb = self.bs * self.world
        for i in range(len(self)):
            ix = order[i * gb:(i + 1) * gb][self.rank::self.world]
            d = self.d
            yield d["x"][ix], d["x"][ix], d["cond"][ix], d["std"][ix], d["intensity"][ix], d["positions"][ix]


def get_train_test_data_loaders(cfg, device="cuda", rank=0, world=1):
    """Counterpart of data_transformations.get_train_test_data_loaders (:260-309).  With ``dataset.synthetic_samples``
    set it builds device-resident synthetic loaders; reading the reference's pickles is out of scope (SURVEY.md §2
    row 15) and raises."""
    n = cfg.dataset.get("synthetic_samples")
    if not n:
        raise NotImplementedError("only dataset.synthetic_samples=N is supported: the reference's GEANT4 pickles are not "
                                  "shipped and its pandas pipeline is outside the hot path")
    arch = cfg.model.architecture
    data = synthetic_showers(arch, int(n), 0, device)
    n_test = int(int(n) * float(cfg.dataset.test_size))
    tr = {k: v[n_test:] for k, v in data.items()}
    te = {k: v[:n_test].cpu() for k, v in data.items()}
    return (DeviceLoader(tr, cfg.train.batch_size, True, rank, world),
            DeviceLoader(te, max(n_test, 1), False, 0, 1, drop_last=False))
