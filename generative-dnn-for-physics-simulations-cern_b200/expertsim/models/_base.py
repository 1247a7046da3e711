"""Common machinery of the drop-in network modules: parameters under the reference's state_dict names that alias an
arena slot, plus stand-alone (no-grad) forward passes through the CUDA engines."""
from __future__ import annotations

import torch
from torch import nn

from .._arena import Arena, register_by_spec, spec_for


class ArenaModule(nn.Module):
    """A network whose parameters live in an Arena slot.  Until a MoEWrapper adopts it into a shared arena it owns a
    private single-slot arena, created lazily on the (CUDA) device of its parameters."""
    ARCH = "proton"
    KIND = "generator"

    def _init_params(self, ctor_kwargs, **spec_kw):
        self._ctor_kwargs = dict(ctor_kwargs)
        self._spec = spec_for(self.ARCH, self.KIND, **spec_kw)
        register_by_spec(self, self._spec)
        self._arena, self._slot = None, 0
        self.register_load_state_dict_post_hook(lambda m, k: m._bump())

    def _bump(self):
        if self._arena is not None:
            self._arena.version += 1

    def _home(self) -> Arena:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("expertsim (B200) modules compute on CUDA only: move the module to a CUDA device "
                               "(there is no CPU fallback)")
        if self._arena is None or self._arena.device != dev or not self._arena.owns(self):
            Arena(self._spec, 1, dev).adopt(self, 0)
        return self._arena

    def __deepcopy__(self, memo):
        # a copy must not alias the source's arena slot (MoEWrapper.__init__ deep-copies one module E times)
        new = self.__class__(**self._ctor_kwargs)
        new.to(next(self.parameters()).device)
        new.load_state_dict(self.state_dict())
        new.train(self.training)
        return new


def one_group(rows: int, slot: int, device, n_slots: int = 1):
    """Group table with one entry per arena slot in which only ``slot`` owns rows (a module used on its own)."""
    t = torch.zeros(n_slots, 4, dtype=torch.int32)
    t[:, 2] = torch.arange(n_slots)
    t[slot] = torch.tensor([0, rows, slot, rows], dtype=torch.int32)
    return t.to(device)
