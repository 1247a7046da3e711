"""Neutron auxiliary coordinate regressor.  Drop-in for AuxRegNeutron (expertsim/models/neutron/aux_reg.py:8-80 of
the reference)."""
import math

import torch
import torch.nn.functional as F

from .._base import ArenaModule, one_group


class AuxRegNeutron(ArenaModule):
    ARCH, KIND = "neutron", "aux_reg"

    def __init__(self, strength, output_dim=2, **kwargs):
        super().__init__()
        self.name = "aux-architecture-neutron"
        self.strength = strength
        if output_dim != 2:
            raise ValueError("the regressor predicts (row, col) of the brightest pixel: output_dim must be 2")
        self._init_params(dict(strength=strength, output_dim=output_dim))

    @torch.no_grad()
    def forward(self, x):
        from ..._nets import engine_for
        arena = self._home()
        eng = engine_for(arena, self.ARCH, self.KIND)
        R = x.shape[0]
        grp = one_group(R, self._slot, x.device, arena.E)
        coords, _ = eng.forward(x.float().reshape(R, -1).contiguous(), grp, R, self.training, None)
        return coords

    @staticmethod
    def regressor_loss(real_coords, fake_coords):
        """mean log-cosh-style loss (reference neutron/aux_reg.py:70-74)."""
        diff = fake_coords - real_coords
        return torch.mean(diff + F.softplus(-2.0 * diff) - math.log(2.0))
