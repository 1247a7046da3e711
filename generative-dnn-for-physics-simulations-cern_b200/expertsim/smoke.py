"""One small invocation of the hot path on cuda:0, checked against the CPU oracle (driver smoke test).
The oracle import lives here and nowhere else in the package: this module is test infrastructure, not product path."""
import copy
import os
import sys

import torch


def run_smoke():
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    import oracle.expertsim_oracle as orc
    from . import _lib as L
    from .config import Config
    from .train.loop import setup_moe_system

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    torch.cuda.set_device(0)
    if not L.device_ok():
        raise RuntimeError("libexpertsim_b200.so has no sm_100a code for this device")
    arch, E, B, seed = "proton", 3, 24, 7
    ocfg = copy.deepcopy(orc.DEFAULT_CFG)
    ocfg["dataset"] = {"input_image_shape": [56, 30]}
    st = orc.make_state(arch, E, seed, ocfg)
    moe = setup_moe_system(Config(ocfg), torch.device("cuda:0"))
    for e in range(E):
        moe.generators[e].load_state_dict(st.gens[e])
        moe.discriminators[e].load_state_dict(st.discs[e])
        moe.aux_regs[e].load_state_dict(st.auxs[e])
    moe.router.load_state_dict(st.router)
    moe.train()
    batch, noise = orc.make_batch(arch, B, seed), orc.make_noise(arch, B, E, seed)
    want, aux = orc.train_step(st, batch, noise, epoch=0)
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    b = dev(batch)
    n0 = L.n_calls
    got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=dev(noise))
    torch.cuda.synchronize()
    assert moe._last["idx"].cpu().tolist() == aux["idx"].tolist(), "routing differs from the oracle"
    worst = 0.0
    for k, v in want.items():
        g = float(got[k])
        err = abs(g - v) / (abs(v) + 0.25)   # 3e-2 relative + 8e-3 absolute (bf16 images seen through O(1) hinge scores)
        worst = max(worst, err)
        assert err <= 3e-2, f"{k}: got {g}, oracle {v}"
    print(f"smoke ok: routing bit-exact, {len(want)} metrics within 3e-2 rel + 8e-3 abs of the oracle (worst {worst:.2e}), "
          f"{L.n_calls - n0} C-ABI kernel calls, gen_loss={float(got['gen_loss']):.6f}")
