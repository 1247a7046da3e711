"""Model registry.  Drop-in for expertsim/models/__init__.py:11-28 of the reference: the same registry keys and
``build_model(name, model_specs, device)``.  The reference registers two classes that do not exist in its own tree
(``proton.generator_unified`` -> GeneratorUnified and ``router_attention`` -> AttentionRouterNetwork make its
``import expertsim.models`` raise AttributeError); those keys are not offered here."""
from .neutron.aux_reg import AuxRegNeutron
from .neutron.discriminator import DiscriminatorNeutron
from .neutron.generator import GeneratorNeutron
from .proton.aux_reg import AuxReg
from .proton.discriminator import Discriminator
from .proton.generator import Generator
from .routers.router import RouterNetwork

MODEL_REGISTRY = {
    "proton.generator": Generator,
    "proton.discriminator": Discriminator,
    "proton.aux_reg": AuxReg,
    "neutron.generator": GeneratorNeutron,
    "neutron.discriminator": DiscriminatorNeutron,
    "neutron.aux_reg": AuxRegNeutron,
    "router_v1": RouterNetwork,
}


def build_model(name, model_specs, device):
    """``MODEL_REGISTRY[name](**model_specs).to(device)`` (reference models/__init__.py:25-28)."""
    if name not in MODEL_REGISTRY:
        raise KeyError(f"unknown model '{name}'; available: {sorted(MODEL_REGISTRY)}")
    return MODEL_REGISTRY[name](**dict(model_specs)).to(device)
