"""Proton auxiliary coordinate regressor.  Drop-in for AuxReg (expertsim/models/proton/aux_reg.py:11-45 of the
reference); FeatureExtractor / ResidualBlock (:55-131) are executed by AuxEngineProton."""
import math

import torch
import torch.nn.functional as F

from ..._nets import engine_for
from .._base import ArenaModule, one_group


class AuxReg(ArenaModule):
    ARCH, KIND = "proton", "aux_reg"

    def __init__(self, strength, output_dim=2, **kwargs):
        super().__init__()
        self.name = "regressor_v3_changed_loss_log_cosh"
        self.strength = strength
        if output_dim != 2:
            raise ValueError("the regressor predicts (row, col) of the brightest pixel: output_dim must be 2")
        self._init_params(dict(strength=strength, output_dim=output_dim))

    @torch.no_grad()
    def forward(self, x):
        arena = self._home()
        eng = engine_for(arena, self.ARCH, self.KIND)
        R = x.shape[0]
        grp = one_group(R, self._slot, x.device, arena.E)
        masks = None
        if self.training:
            masks = ((torch.rand(R, 128, device=x.device) >= 0.3).float(), (torch.rand(R, 64, device=x.device) >= 0.3).float())
        coords, _ = eng.forward(x.float().reshape(R, -1).contiguous(), grp, R, self.training, masks)
        return coords

    @staticmethod
    def regressor_loss(real_coords, fake_coords):
        """mean log-cosh-style loss (reference proton/aux_reg.py:42-45); torch ops, used only outside the fused step."""
        diff = fake_coords - real_coords
        return torch.mean(diff + F.softplus(-2.0 * diff) - math.log(2.0))
