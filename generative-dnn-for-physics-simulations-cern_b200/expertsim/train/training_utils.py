"""Checkpoint writer.  Counterpart of ``save_models_and_architectures`` (expertsim/train/training_utils.py:316-380 of
the reference), same file names; state_dicts are stored instead of pickled module objects so the files load into both
the reference's modules (``load_state_dict``) and this build's."""
import os

import torch


def save_models_and_architectures(filepath_models, n_experts, aux_regs, aux_reg_optimizers, generators,
                                  generator_optimizers, discriminators, discriminator_optimizers, router_network,
                                  router_optimizer, epoch):
    os.makedirs(filepath_models, exist_ok=True)
    cpu = lambda sd: {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
    for i in range(n_experts):
        for prefix, mods, opts in (("gen", generators, generator_optimizers), ("disc", discriminators, discriminator_optimizers),
                                   ("aux_reg", aux_regs, aux_reg_optimizers)):
            torch.save(cpu(mods[i].state_dict()), os.path.join(filepath_models, f"{prefix}_{i}_epoch_{epoch}.pth"))
            if opts is not None:
                torch.save(cpu(opts[i].state_dict()), os.path.join(filepath_models, f"{prefix}_optim_{i}_epoch_{epoch}.pth"))
    torch.save(cpu(router_network.state_dict()), os.path.join(filepath_models, f"router_network_epoch_{epoch}.pth"))
    if router_optimizer is not None:
        torch.save(cpu(router_optimizer.state_dict()), os.path.join(filepath_models, f"router_optim_epoch_{epoch}.pth"))
