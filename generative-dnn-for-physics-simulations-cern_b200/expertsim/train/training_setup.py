"""Optimizer setup and checkpoint loading.  Drop-in for expertsim/train/training_setup.py of the reference.

``setup_optimizers`` keeps the reference's signature and return order (training_setup.py:12-41: generator list,
discriminator list, aux-reg list, router) but returns ``ArenaAdam`` handles: the Adam state (exp_avg, exp_avg_sq, step)
lives next to the parameters in the expert arenas and ONE fused multi-tensor kernel (es_adam_step) updates every expert
of a network kind inside ``MoEWrapper.train_step``.  Each handle still looks like a torch optimizer (``param_groups``,
``zero_grad``, ``state_dict``/``load_state_dict``) so hooks and checkpointing code written for the reference work."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch

from .._arena import Arena


def count_model_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def _torch_adam_defaults(lr):
    """param_group defaults of torch.optim.Adam in the running torch version (so a state_dict written here loads into a
    stock torch optimizer and vice versa)"""
    d = dict(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=float(lr)).defaults)
    d["lr"] = float(lr)
    return d


class ArenaAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) for one arena slot (one expert of one network kind).

    ``state_dict()`` / ``load_state_dict()`` speak torch.optim.Adam's format (``{'state': {i: {'step', 'exp_avg',
    'exp_avg_sq'}}, 'param_groups': [{..., 'params': [0..n-1]}]}``, parameters in ``module.parameters()`` order — the
    reference's order, since the modules register the reference's names in the reference's order), so the
    ``*_optim_*_epoch_N.pth`` files of the reference (train/training_utils.py:316-380) and of this build interchange in
    both directions.  The moments themselves live in the arena (``M``, ``V``, ``steps``)."""

    def __init__(self, module, lr: float):
        self.module = module
        self.defaults = _torch_adam_defaults(lr)
        self.param_groups = [dict(self.defaults, params=list(module.parameters()))]

    @property
    def _arena(self) -> Arena:
        return self.module._arena

    def zero_grad(self, set_to_none: bool = True):
        a = self._arena
        if a is not None:
            a.G[self.module._slot].zero_()

    def step(self, closure=None):
        """Stand-alone update of this slot from its gradient slice (train_step does this for all slots at once)."""
        from .. import _lib as L
        a, s = self._arena, self.module._slot
        g = self.param_groups[0]
        L.call("es_adam_step", a.P[s], a.G[s], a.M[s], a.V[s], a.n, a.n, 1, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
               a.steps[s:s + 1], None)
        a.version += 1

    def _names(self):
        return [n for n, _ in self.module.named_parameters()]

    def state_dict(self):
        a, s = self._arena, self.module._slot
        names = self._names()
        step = int(a.steps[s])
        state = {}
        if step > 0:        # torch keeps no state for parameters that never stepped
            for i, n in enumerate(names):
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": a.view(a.M, n, s).clone(),
                            "exp_avg_sq": a.view(a.V, n, s).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(names)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        a, s = self._arena, self.module._slot
        names = self._names()
        if "state" in sd:                                   # torch.optim.Adam format (the reference's files, and ours)
            groups = sd["param_groups"]
            order = [i for g in groups for i in g["params"]]
            if len(order) != len(names):
                raise ValueError(f"optimizer state holds {len(order)} parameters, the module has {len(names)}")
            steps = set()
            a.M[s].zero_()
            a.V[s].zero_()
            for pos, key in enumerate(order):
                st = sd["state"].get(key)
                if st is None:
                    continue
                m, v = a.view(a.M, names[pos], s), a.view(a.V, names[pos], s)
                if tuple(st["exp_avg"].shape) != tuple(m.shape):
                    raise ValueError(f"{names[pos]}: optimizer state has shape {tuple(st['exp_avg'].shape)}, expected {tuple(m.shape)}")
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError(f"parameters of one module carry different step counts {sorted(steps)}: not representable "
                                 "by the fused per-expert Adam")
            a.steps[s] = steps.pop() if steps else 0
            g0 = {k: v for k, v in groups[0].items() if k != "params"}
        else:                                               # flat layout written by round-1 builds of this package
            a.steps[s] = int(sd["step"])
            a.M[s].copy_(sd["exp_avg"])
            a.V[s].copy_(sd["exp_avg_sq"])
            g0 = {k: v for k, v in sd["param_groups"][0].items() if k != "params"}
        self.param_groups[0].update(g0)


def setup_optimizers(wrapper, cfg) -> Tuple[List, List, List, ArenaAdam]:
    """-> (generator_optimizers, discriminator_optimizers, aux_reg_optimizers, router_optimizer)."""
    gen_optims = [ArenaAdam(g, cfg.model.generator.lr_g) for g in wrapper.generators]
    disc_optims = [ArenaAdam(d, cfg.model.discriminator.lr_d) for d in wrapper.discriminators]
    aux_reg_optims = [ArenaAdam(a, cfg.model.aux_reg.lr_a) for a in wrapper.aux_regs]
    router_optim = ArenaAdam(wrapper.router, cfg.model.router.lr_r)
    return gen_optims, disc_optims, aux_reg_optims, router_optim


def print_model_info(wrapper):
    print("=== MoE System Information ===")
    print(f"Number of experts: {wrapper.n_experts}")
    total = 0
    for kind, mods in (("Generator", wrapper.generators), ("Discriminator", wrapper.discriminators),
                       ("Aux Regressor", wrapper.aux_regs)):
        for i, m in enumerate(mods):
            n = count_model_parameters(m)
            print(f"{kind} {i}: {n:,} parameters")
            total += n
    n = count_model_parameters(wrapper.router)
    print(f"Router: {n:,} parameters")
    print(f"Total parameters: {total + n:,}")


def load_checkpoint_weights(checkpoint_dir, epoch, wrapper, gen_optims=None, disc_optims=None, aux_reg_optims=None,
                            router_optim=None, device=None):
    """Load the files written by ``save_models_and_architectures`` (train/training_utils.py) for ``epoch``.  The
    reference pickles whole module objects, which cannot be un-pickled into other classes; this build stores
    ``state_dict``s under the same file names, and also accepts a pickled reference module (its state_dict is taken)."""
    def _sd(path):
        obj = torch.load(path, map_location=device or "cpu", weights_only=False)
        return obj.state_dict() if hasattr(obj, "state_dict") and not isinstance(obj, dict) else obj

    for i in range(wrapper.n_experts):
        for prefix, mods, opts in (("gen", wrapper.generators, gen_optims), ("disc", wrapper.discriminators, disc_optims),
                                   ("aux_reg", wrapper.aux_regs, aux_reg_optims)):
            mods[i].load_state_dict(_sd(os.path.join(checkpoint_dir, f"{prefix}_{i}_epoch_{epoch}.pth")))
            p = os.path.join(checkpoint_dir, f"{prefix}_optim_{i}_epoch_{epoch}.pth")
            if opts is not None and os.path.exists(p):
                opts[i].load_state_dict(_sd(p))
    wrapper.router.load_state_dict(_sd(os.path.join(checkpoint_dir, f"router_network_epoch_{epoch}.pth")))
    p = os.path.join(checkpoint_dir, f"router_optim_epoch_{epoch}.pth")
    if router_optim is not None and os.path.exists(p):
        router_optim.load_state_dict(_sd(p))
    wrapper.mark_weights_changed()
    return wrapper
