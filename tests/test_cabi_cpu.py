"""The C-ABI library must load without a GPU and export every symbol include/expertsim_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from expertsim import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_built_in_tree():
    assert os.path.exists(L.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    assert L.LIB_PATH.startswith(ROOT), "the library must live in-tree so it travels to the GPU box"


def test_every_declared_symbol_is_exported():
    lib = L.load()
    protos = L.prototypes()
    assert len(protos) >= 50
    src = open(L.HEADER_PATH).read()
    declared = set(re.findall(r"\b(es_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", src, flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_header_cites_the_reference_interfaces():
    src = open(L.HEADER_PATH).read()
    for cite in ("routers/router.py", "models/moe.py", "proton/generator.py", "proton/discriminator.py", "proton/aux_reg.py",
                 "train/utils.py", "training_setup.py"):
        assert cite in src, f"header does not cite {cite}"


def test_version_error_string_and_device_probe():
    lib = L.load()
    assert lib.es_version() >= 100
    assert isinstance(L.last_error(), str)
    import torch
    if not torch.cuda.is_available():
        assert L.device_ok() is False      # no silent CPU path: the probe says so and the product refuses to compute


def test_binding_rejects_host_tensors():
    import torch
    with pytest.raises(RuntimeError, match="device pointers only"):
        L.call("es_axpy", 1.0, torch.zeros(4), 4, torch.zeros(4))


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(L.ESGroup) == 16 and ctypes.sizeof(L.ESConvGeom) == 44 and ctypes.sizeof(L.ESConv2d) == 40
