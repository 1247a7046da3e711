"""Checks against the UNMODIFIED reference itself (``oracle/_ref`` — made by ``oracle/make_ref.py`` — or /root/reference),
run in a process of its own because its package is also called ``expertsim``.  Skipped where no reference checkout exists."""
import json
import os
import subprocess
import sys

import pytest

import oracle.ref_shim as shim
from expertsim.models import build_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(shim.find_reference() is None, reason="no reference checkout (oracle/_ref or /root/reference)")


def run_ref(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), *args], capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_ref
def test_parameter_and_state_dict_order_equal_the_reference_modules():
    """torch.optim.Adam indexes its state by ``module.parameters()`` position: optimizer checkpoints interchange with the
    reference only if this build's modules enumerate their parameters in the reference's order."""
    ref = run_ref("names")
    kw = {"generator": dict(noise_dim=10, cond_dim=9, di_strength=0.1, in_strength=1e-3), "discriminator": dict(cond_dim=9),
          "aux_reg": dict(strength=1e-3)}
    for arch in ("proton", "neutron"):
        for kind in ("generator", "discriminator", "aux_reg"):
            m = build_model(f"{arch}.{kind}", kw[kind], "cpu")
            assert [n for n, _ in m.named_parameters()] == ref[f"{arch}.{kind}"]["params"], (arch, kind)
            assert list(m.state_dict()) == ref[f"{arch}.{kind}"]["state_dict"], (arch, kind)
    r = build_model("router_v1", dict(cond_dim=9, n_experts=5), "cpu")
    assert [n for n, _ in r.named_parameters()] == ref["router"]["params"]


@needs_ref
def test_reference_runner_reproduces_the_golden_step(golden_dir):
    """the runner that feeds ``bench.py --impl reference`` and the same-device GPU parity test drives the reference
    correctly: its injected-noise step equals the committed golden fixture (which the pin script wrote from the reference)"""
    import torch
    out = os.path.join(ROOT, "gpurun_out", "ref_step_cpu.pt")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    run_ref("step", "--arch", "proton", "--experts", "3", "--batch", "24", "--seed", "7", "--device", "cpu", "--out", out)
    got = torch.load(out, weights_only=False)
    want = json.load(open(os.path.join(golden_dir, "train_step_proton_E3_B24.json")))["case"]["steps"][0]
    assert got["idx"].tolist() == want["idx"] and got["counts"].tolist() == want["counts"]
    for k, v in want["metrics"].items():
        assert abs(got["metrics"][k] - v) <= 1e-6 * max(1.0, abs(v)), k
    assert set(got["fake1"]) == {0, 1, 2} and tuple(got["fake1"][0].shape)[-2:] == (56, 30)
