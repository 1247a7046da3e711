"""Neutron discriminator.  Drop-in for DiscriminatorNeutron (expertsim/models/neutron/discriminator.py:6-48 of the
reference): the proton stack with 44x44 inputs and a (2,2) second pooling window."""
from ..proton.discriminator import Discriminator


class DiscriminatorNeutron(Discriminator):
    ARCH, KIND = "neutron", "discriminator"
