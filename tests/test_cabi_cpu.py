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


def test_forward_gemm_variant_plan():
    """es_igemm_fwd_plan (host-only): which kernel variant a conv takes.  3 = TMA-fed CTA pair with tap-row strips (N tile
    <= 128, equal rows of consecutive taps), 2 = TMA-fed CTA pair, one im2col box per tap (the conv reads its source
    directly, N tile >= 64, not the dense product), 1 = strip (no upsample, N <= 128, 2..4 taps per row, nx * BN <= 384, where
    variant 2 does not apply), 0 = single-CTA kernel with the cp.async gather (upsample inside the conv, fc2)."""
    import ctypes

    from expertsim import _lib as L
    lib = L.load()

    def plan(Hs, Ws, C, Hu, Wu, K, pad, N, rows=2048):
        g = L.ESConvGeom(Hs, Ws, C, Hu, Wu, Hu + 2 * pad - K + 1, Wu + 2 * pad - K + 1, K, K, pad, N)
        out = (ctypes.c_int32 * 8)()
        assert lib.es_igemm_fwd_plan(ctypes.addressof(g), rows, ctypes.addressof(out)) == 0, L.last_error()
        return list(out)

    # proton conv3 forward (3x3, 128 -> 64) and its data gradient (64 -> 128): 3 strips of 3 taps on the padded pitch 29 + 2
    assert plan(55, 29, 128, 55, 29, 3, 1, 64) == [3, 64, 3, 3, 31, 4, 6, (55 * 31 + 127) // 128]
    assert plan(55, 29, 64, 55, 29, 3, 1, 128) == [3, 128, 3, 3, 31, 3, 3, (55 * 31 + 127) // 128]
    # conv2's data gradient geometry (4x4, pad 2, N = 256): one im2col box per tap; neutron conv9 (2x2, no padding): strips of 2
    assert plan(55, 29, 128, 55, 29, 4, 2, 256)[:3] == [2, 256, 16]
    assert plan(46, 46, 128, 46, 46, 2, 0, 64)[:5] == [3, 64, 2, 2, 46]
    # an upsample in front of the conv or the dense 1x1 product (fc2) keep the single-CTA gather kernel
    assert plan(35, 19, 256, 56, 30, 4, 1, 128)[:2] == [0, 128]
    assert plan(1, 1, 256, 1, 1, 1, 0, 92160)[:3] == [0, 256, 0]
    # N = 32: too narrow for a CTA pair -> the strip variant (3 tap rows of 3), or 4 taps per row while nx * BN <= 384
    assert plan(20, 20, 64, 20, 20, 3, 1, 32)[:5] == [1, 32, 3, 3, 22]
    assert plan(20, 20, 64, 20, 20, 4, 1, 32)[:4] == [1, 32, 4, 4]
