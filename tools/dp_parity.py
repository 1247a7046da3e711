#!/usr/bin/env python
"""Data-parallel parity on N GPUs (launch: python -m torch.distributed.run --nproc-per-node N tools/dp_parity.py).

Every rank holds rows r::N of ONE global batch (and of the injected noise), runs MoEWrapper.train_step with data
parallelism enabled, and the all-reduced gradients / losses are compared with the CPU oracle's step on the WHOLE batch:
the sharded step must equal the single-device global-batch step.  Images are injected from the oracle so
the comparison runs at fp32 / bf16-kernel tolerance (see tests/test_step_gpu.py for why)."""
import copy
import os
import re
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle.expertsim_oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from expertsim.config import Config
    from expertsim.train.loop import setup_moe_system
    arch = sys.argv[1] if len(sys.argv) > 1 else "proton"
    E, seed = 3, 7
    B = 24 if 24 % world == 0 else 8 * world
    H, W = orc.IMAGE_SHAPE[arch]
    ocfg = copy.deepcopy(orc.DEFAULT_CFG)
    ocfg["model"]["n_experts"] = E
    ocfg["model"]["architecture"] = arch
    ocfg["dataset"] = {"input_image_shape": [H, W]}
    st = orc.make_state(arch, E, seed, ocfg)
    moe = setup_moe_system(Config(ocfg), dev)
    for e in range(E):
        moe.generators[e].load_state_dict(st.gens[e])
        moe.discriminators[e].load_state_dict(st.discs[e])
        moe.aux_regs[e].load_state_dict(st.auxs[e])
    moe.router.load_state_dict(st.router)
    moe.train()
    moe.enable_data_parallel()
    batch, noise = orc.make_batch(arch, B, seed), orc.make_noise(arch, B, E, seed)
    collect = {}
    want, aux = orc.train_step(st, batch, noise, epoch=0, collect=collect)     # global batch, every rank (deterministic)
    mine = torch.arange(B)[rank::world]
    sh = lambda d: {k: v[mine].to(dev) for k, v in d.items()}
    nz = sh(noise)
    masks = [(aux["idx"] == e).nonzero(as_tuple=True)[0] for e in range(E)]
    for key, name in (("fake1", "img1_sorted"), ("fake2", "img2_sorted")):
        rows = []
        for e in range(E):
            sel = (masks[e] % world) == rank
            src = aux[key][e].reshape(-1, H * W) if e in aux[key] else torch.zeros(masks[e].numel(), H * W)
            rows.append(src[sel])
        nz[name] = torch.cat(rows).to(dev)
    b = sh(batch)
    got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=nz)
    torch.cuda.synchronize()
    fails = []
    assert moe._last["idx"].cpu().tolist() == aux["idx"][mine].tolist(), "routing of the shard differs"
    for k, v in want.items():
        g = float(got[k])
        if abs(g - v) > 3e-2 * abs(v) + 8e-3:
            fails.append(f"metric {k}: {g} vs {v}")
    tol = dict(g=0.2, d=2e-3, a=2e-3)
    worst = dict(g=0.0, d=0.0, a=0.0)
    for key, kind in (("g", "g_grads"), ("d", "d_grads"), ("a", "a_grads")):
        arena = moe.arena(key)
        for e in range(E):
            for name, gw in collect.get(f"{kind}_{e}", {}).items():
                if float(gw.abs().max()) < 1e-9:
                    continue
                if arch == "neutron" and ((key == "g" and name in ("fc1.0.bias", "fc2.0.bias", "conv_layers.0.bias", "conv_layers.5.bias",
                                                                   "conv_layers.9.bias"))
                                          or (key == "a" and re.fullmatch(r"feature_extractor\.conv\d\.bias", name))):
                    continue          # a bias in front of a BatchNorm: identically-zero gradient, autograd returns rounding noise
                gg = arena.view(arena.G, name, e).double().cpu()
                r = float((gg - gw.double()).norm() / gw.double().norm())
                worst[key] = max(worst[key], r)
                if r > tol[key]:
                    fails.append(f"grad {key}{e} {name}: relL2 {r:.3e}")
    # replicas must stay bit-identical after the step
    chk = torch.stack([moe.arena(k).P.double().sum() for k in "gdar"])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        fails.append("replicas diverged after the step")
    ok = torch.tensor([0 if fails else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dp parity {arch} world={world} B={B}: worst relL2 g={worst['g']:.3e} d={worst['d']:.3e} a={worst['a']:.3e}; "
              f"gen_loss {float(got['gen_loss']):.6f} vs oracle {want['gen_loss']:.6f}")
    for f in fails:
        print(f"[rank {rank}] FAIL {f}")
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)


if __name__ == "__main__":
    main()
