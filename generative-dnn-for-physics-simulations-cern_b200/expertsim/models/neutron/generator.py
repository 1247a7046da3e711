"""Neutron ZDC generator (44x44).  Drop-in for GeneratorNeutron (expertsim/models/neutron/generator.py:5-49 of the
reference): same constructor and state_dict keys (BatchNorm running statistics included)."""
import torch

from .._base import ArenaModule, one_group


class GeneratorNeutron(ArenaModule):
    ARCH, KIND = "neutron", "generator"
    IMAGE_SHAPE = (44, 44)

    def __init__(self, noise_dim, cond_dim, di_strength, in_strength, **kwargs):
        super().__init__()
        self.name = "Generator-neutron"
        self.di_strength = di_strength
        self.in_strength = in_strength
        if (noise_dim, cond_dim) != (10, 9):
            raise ValueError("the sm_100a generator head is specialised for noise_dim=10, cond_dim=9")
        self._init_params(dict(noise_dim=noise_dim, cond_dim=cond_dim, di_strength=di_strength, in_strength=in_strength),
                          noise_dim=noise_dim, cond_dim=cond_dim)

    @torch.no_grad()
    def forward(self, noise, cond):
        from ..._nets import engine_for
        arena = self._home()
        eng = engine_for(arena, self.ARCH, self.KIND)
        R = noise.shape[0]
        grp = one_group(R, self._slot, noise.device, arena.E)
        img, _, _ = eng.forward(noise.float().contiguous(), None, cond.float().contiguous(), grp, R, False, keep=False,
                                training=self.training)
        return img.view(R, 1, *self.IMAGE_SHAPE)
