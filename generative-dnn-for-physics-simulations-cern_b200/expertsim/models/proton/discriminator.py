"""Proton discriminator.  Drop-in for Discriminator (expertsim/models/proton/discriminator.py:116-155 of the reference):
spectral-normalised conv/linear stack, hinge scores.  The reference's unused DiscriminatorUnified / GroupedLinear
prototypes (same file, :8-113) are not reproduced (dead code)."""
import torch

from ..._nets import engine_for
from .._base import ArenaModule, one_group


class Discriminator(ArenaModule):
    ARCH, KIND = "proton", "discriminator"

    def __init__(self, cond_dim, **kwargs):
        super().__init__()
        self.name = "Discriminator-5-hinge-spectralnorm"
        self._init_params(dict(cond_dim=cond_dim), cond_dim=cond_dim)

    @torch.no_grad()
    def forward(self, img, cond):
        arena = self._home()
        eng = engine_for(arena, self.ARCH, self.KIND)
        R = img.shape[0]
        grp = one_group(R, self._slot, img.device, arena.E)
        sn = eng.spectral(grp, self.training)   # one power iteration per forward in training mode, as the hook does
        score, latent, _ = eng.forward(img.float().reshape(R, -1).contiguous(), cond.float().contiguous(), grp, R, sn)
        return score, latent
