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

    @classmethod
    def from_arrays(cls, x, cond, std, intensity, positions, batch_size, device="cuda", x_2=None, **kw):
        """Loader over the arrays ``transform_data_for_training`` returns (data_transformations.py:118-258: x_train,
        y_train, std_train, intensity_train, positions_train — numpy or torch, any float dtype), moved to ``device`` ONCE.
        The reference wraps the same arrays in a TensorDataset of 6 tensors (data_transformations.py:269-271); batches
        of this loader carry the same 6 positions ``(x, x_2, cond, std, intensity, positions)``."""
        t = lambda a, w=None: _as_f32(a, w).to(device)
        d = {"x": t(x), "cond": t(cond), "std": t(std, 1), "intensity": t(intensity, 1), "positions": t(positions)}
        n = d["x"].shape[0]
        for k, v in d.items():
            if v.shape[0] != n:
                raise ValueError(f"{k} holds {v.shape[0]} rows, x holds {n}")
        if x_2 is not None:
            d["x_2"] = t(x_2)
        return cls(d, batch_size, **kw)

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
            x = d["x"][ix]
            yield x, (d["x_2"][ix] if "x_2" in d else x), d["cond"][ix], d["std"][ix], d["intensity"][ix], d["positions"][ix]


def _as_f32(a, width=None):
    """numpy / torch / pandas values -> contiguous float32 CPU tensor; ``width`` reshapes 1-D columns to [n, width]"""
    if hasattr(a, "to_numpy"):
        a = a.to_numpy()
    t = torch.as_tensor(a)
    t = t.to(torch.float32)
    if width is not None and t.dim() == 1:
        t = t.reshape(-1, width)
    return t.contiguous()


class PinnedHostLoader:
    """Batches from HOST arrays through pinned staging buffers: batch i+1 is copied host->device on a copy stream while
    the step of batch i runs (two staging slots, one event per slot).  For sets that do not fit the device or must stay
    on the host; same 6-tuple contract and rank sharding as DeviceLoader.  Replaces the reference's DataLoader with
    ``pin_memory=True, prefetch_factor=4`` worker processes (data_transformations.py:273-279)."""

    def __init__(self, x, cond, std, intensity, positions, batch_size, device="cuda", shuffle=False, rank=0, world=1, seed=0):
        self.h = {"x": _as_f32(x), "cond": _as_f32(cond), "std": _as_f32(std, 1), "intensity": _as_f32(intensity, 1),
                  "positions": _as_f32(positions)}
        self.n, self.bs, self.dev = self.h["x"].shape[0], batch_size, torch.device(device)
        self.shuffle, self.rank, self.world, self.seed, self.epoch = shuffle, rank, world, seed, 0
        pin = self.dev.type == "cuda"
        self.stage = [{k: torch.empty((batch_size,) + tuple(v.shape[1:]), pin_memory=pin) for k, v in self.h.items()} for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev) if pin else None

    def __len__(self):
        return self.n // (self.bs * self.world)

    def _fill(self, slot, ix):
        for k, v in self.h.items():
            torch.index_select(v, 0, ix, out=self.stage[slot][k])

    def __iter__(self):
        order = torch.randperm(self.n, generator=torch.Generator().manual_seed(self.seed + self.epoch)) if self.shuffle \
            else torch.arange(self.n)
        self.epoch += 1
        gb = self.bs * self.world
        nb = len(self)
        sel = lambda i: order[i * gb:(i + 1) * gb][self.rank::self.world]
        if self.copy_stream is None:          # host-only use (CPU tests of the contract)
            for i in range(nb):
                self._fill(0, sel(i))
                d = {k: v.clone() for k, v in self.stage[0].items()}
                yield d["x"], d["x"], d["cond"], d["std"], d["intensity"], d["positions"]
            return
        free = [torch.cuda.Event(), torch.cuda.Event()]     # slot's previous device copy has been issued and finished
        pending = None

        def launch(i):
            slot = i % 2
            free[slot].synchronize()                         # the staging buffer is no longer being read by a copy
            self._fill(slot, sel(i))
            with torch.cuda.stream(self.copy_stream):
                d = {k: v.to(self.dev, non_blocking=True) for k, v in self.stage[slot].items()}
                free[slot].record(self.copy_stream)
                ready = self.copy_stream.record_event()
            return d, ready

        if nb:
            pending = launch(0)
        for i in range(nb):
            d, ready = pending
            pending = launch(i + 1) if i + 1 < nb else None   # overlaps with the consumer's step of batch i
            torch.cuda.current_stream(self.dev).wait_event(ready)
            for v in d.values():
                v.record_stream(torch.cuda.current_stream(self.dev))
            yield d["x"], d["x"], d["cond"], d["std"], d["intensity"], d["positions"]


def get_train_test_data_loaders(cfg, device="cuda", rank=0, world=1):
    """Counterpart of data_transformations.get_train_test_data_loaders (:260-309).  With ``dataset.synthetic_samples``
    set it builds device-resident synthetic loaders; reading the reference's pickles is out of scope (SURVEY.md §2
    row 15) and raises."""
    n = cfg.dataset.get("synthetic_samples")
    if not n:
        raise NotImplementedError("dataset.synthetic_samples=N builds synthetic loaders; real data enters through "
                                  "loaders_from_arrays(...) with the arrays of the reference's transform_data_for_training "
                                  "(its GEANT4 pickles and pandas pipeline are not shipped: SURVEY.md §2 row 15)")
    arch = cfg.model.architecture
    data = synthetic_showers(arch, int(n), 0, device)
    n_test = int(int(n) * float(cfg.dataset.test_size))
    tr = {k: v[n_test:] for k, v in data.items()}
    te = {k: v[:n_test].cpu() for k, v in data.items()}
    return (DeviceLoader(tr, cfg.train.batch_size, True, rank, world),
            DeviceLoader(te, max(n_test, 1), False, 0, 1, drop_last=False))


def loaders_from_arrays(cfg, x_train, x_test, y_train, y_test, std_train, std_test, intensity_train, intensity_test,
                        positions_train, positions_test, device="cuda", rank=0, world=1, resident=True):
    """(train_loader, test_loader) from the split arrays ``transform_data_for_training`` returns
    (data_transformations.py:260-309 builds TensorDataset/DataLoader from exactly these).  ``resident`` keeps the training
    set on the device (DeviceLoader); otherwise batches stream through pinned double-buffered staging (PinnedHostLoader)."""
    bs = cfg.train.batch_size
    if resident:
        tr = DeviceLoader.from_arrays(x_train, y_train, std_train, intensity_train, positions_train, bs, device,
                                      shuffle=False, rank=rank, world=world)
    else:
        tr = PinnedHostLoader(x_train, y_train, std_train, intensity_train, positions_train, bs, device, False, rank, world)
    te = DeviceLoader.from_arrays(x_test, y_test, std_test, intensity_test, positions_test, bs, "cpu", shuffle=False,
                                  drop_last=False)
    return tr, te
