"""Import shim for the UNMODIFIED reference package (TEST / BASELINE INFRASTRUCTURE ONLY — never on the product path).

The reference (patrick-bedkowski/Generative-DNN-for-Physics-Simulations-CERN) is pure Python + eager PyTorch.  Its
package is also called ``expertsim``, so it cannot live in one interpreter together with this repo's drop-in package:
everything that runs the reference does so in a process of its own (``oracle/ref_runner.py``, ``bench.py --impl
reference``, ``oracle/pin_against_reference.py``) and calls ``install_reference()`` BEFORE importing ``expertsim``.

Where the reference comes from:
  * ``oracle/_ref/``   — a verbatim copy made by ``oracle/make_ref.py`` (git-ignored, shipped to the GPU box by gpurun),
  * ``/root/reference`` — the read-only checkout in the build container.
The shim of SURVEY.md §8(c): stub the plotting / logging imports the hot path never calls (matplotlib, seaborn,
wandb) and bypass ``expertsim/models/__init__.py`` (it imports modules that do not exist in the checkout).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def find_reference():
    """-> directory that holds the reference's ``expertsim/`` package, or None."""
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "expertsim", "models", "moe.py")):
            return root
    return None


def install_reference(root=None):
    """Make ``import expertsim`` resolve to the reference.  Raises if this repo's package is already imported."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError("no reference checkout: run oracle/make_ref.py in the build container")
    mod = sys.modules.get("expertsim")
    if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(os.path.abspath(root)):
        raise RuntimeError("this repo's expertsim package is already imported; run the reference in a process of its own")
    # drop this repo's product package from the search path: the reference must win the name
    sys.path[:] = [p for p in sys.path if not os.path.isfile(os.path.join(p, "expertsim", "_lib.py"))]
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "wandb"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.rcParams = {}
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, root)
    import expertsim  # noqa: F401  (the reference package)
    pkg = types.ModuleType("expertsim.models")
    pkg.__path__ = [os.path.join(root, "expertsim", "models")]
    sys.modules["expertsim.models"] = pkg
    return root


class AttrDict(dict):
    """dict with attribute access — stands in for the OmegaConf object the reference's cli builds"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    return AttrDict({k: to_attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def reference_classes(arch):
    """(Generator, Discriminator, AuxReg) classes of the reference for ``arch`` (after install_reference())."""
    if arch == "proton":
        from expertsim.models.proton.generator import Generator as G
        from expertsim.models.proton.discriminator import Discriminator as D
        from expertsim.models.proton.aux_reg import AuxReg as A
    else:
        from expertsim.models.neutron.generator import GeneratorNeutron as G
        from expertsim.models.neutron.discriminator import DiscriminatorNeutron as D
        from expertsim.models.neutron.aux_reg import AuxRegNeutron as A
    return G, D, A


def build_reference_moe(arch, E, cfg, image_shape, state=None, device="cpu"):
    """The reference's own MoEWrapper (expertsim/models/moe.py:14-50) + its optimizers
    (expertsim/train/training_setup.py:12-41).  ``state`` (oracle ``State``) loads given weights, else torch's init."""
    from expertsim.models.moe import MoEWrapper
    from expertsim.models.routers.router import RouterNetwork
    from expertsim.train.training_setup import setup_optimizers
    G, D, A = reference_classes(arch)
    m = cfg.model
    moe = MoEWrapper(G(m.noise_dim, m.cond_dim, m.generator.di_strength, m.generator.in_strength), D(m.cond_dim),
                     A(m.aux_reg.strength), RouterNetwork(m.cond_dim, E), E, cfg, image_shape=tuple(image_shape))
    if state is not None:
        for e in range(E):
            moe.generators[e].load_state_dict(state.gens[e])
            moe.discriminators[e].load_state_dict(state.discs[e])
            moe.aux_regs[e].load_state_dict(state.auxs[e])
        moe.router.load_state_dict(state.router)
    moe.to(device)
    opts = setup_optimizers(moe, cfg)
    return moe, opts


class NoiseInjector:
    """Feeds pre-drawn noise to the reference's torch.randn / Tensor.exponential_ / F.dropout call sites, so that the
    unmodified reference and the build under test consume identical random draws (moved to ``device`` when popped)."""

    def __init__(self, device="cpu"):
        self.randn_q, self.expo_q, self.drop_q = [], [], []
        self.device = device
        self._orig = {}

    def __enter__(self):
        import torch
        import torch.nn.functional as F
        self._orig = {"randn": torch.randn, "expo": torch.Tensor.exponential_, "drop": F.dropout}
        inj = self

        def randn(*size, **kw):
            t = inj.randn_q.pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(t.shape) == shape, (t.shape, shape)
            return t.clone().to(kw.get("device") or inj.device)

        def exponential_(self_t, *a, **kw):
            t = inj.expo_q.pop(0)
            assert t.shape == self_t.shape, (t.shape, self_t.shape)
            return self_t.copy_(t)

        def dropout(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            m, pm = inj.drop_q.pop(0)
            assert abs(pm - p) < 1e-12 and m.numel() == x.numel(), (pm, p, m.shape, x.shape)
            return x * m.to(x.device).view_as(x) / (1.0 - p)

        torch.randn = randn
        torch.Tensor.exponential_ = exponential_
        F.dropout = dropout
        return self

    def __exit__(self, *a):
        import torch
        import torch.nn.functional as F
        torch.randn = self._orig["randn"]
        torch.Tensor.exponential_ = self._orig["expo"]
        F.dropout = self._orig["drop"]
        if a[0] is None:
            assert not self.randn_q and not self.expo_q and not self.drop_q, "unused injected noise"
