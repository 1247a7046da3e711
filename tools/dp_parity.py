#!/usr/bin/env python
"""Data-parallel parity: the sharded step must equal the single-device global-batch step.

    python -m torch.distributed.run --nproc-per-node N tools/dp_parity.py [proton|neutron] [--unbalanced] [--one-gpu] [--experts=E]

Every rank holds rows r::N of ONE global batch (and of the injected noise), runs ``MoEWrapper.train_step`` with data
parallelism enabled, and the all-reduced gradients / losses are compared with the CPU oracle's step on the WHOLE batch.
Images are injected from the oracle so the comparison runs at fp32 / bf16-kernel tolerance (tests/test_step_gpu.py says
why).  ``--unbalanced`` re-orders the global batch so that rank N-1 holds NO row of expert 0 although the expert is alive
globally — the regime in which per-expert state updates must be gated on the global count.  After the step the replicas
must be BIT-identical: parameters, Adam moments, step counters, float and integer buffers of all four arenas.

``--one-gpu`` (or a box with fewer GPUs than ranks): all ranks share cuda:0 and the collectives run over gloo — the same
host logic and kernels, no NCCL; this is how the driver's single-GPU test box exercises the N>1 path.
``run_dp_parity`` is also called by ``bench.py --gpus N`` before its timed region (the ``dp_parity`` key of the line).
"""
import copy
import os
import re
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle.expertsim_oracle as orc  # noqa: E402   (the checker)


def unbalanced_order(idx, world, expert=0):
    """permutation of the global batch that puts every sample of ``expert`` on positions p with p % world != world-1,
    so the last rank (rows world-1::world) holds none of them"""
    B = idx.numel()
    mine = [int(i) for i in (idx == expert).nonzero(as_tuple=True)[0]]
    rest = [i for i in range(B) if i not in set(mine)]
    slots_ok = [p for p in range(B) if p % world != world - 1]
    assert len(mine) <= len(slots_ok)
    order = [None] * B
    for p, i in zip(slots_ok, mine):
        order[p] = i
    free = [p for p in range(B) if order[p] is None]
    for p, i in zip(free, rest):
        order[p] = i
    return torch.tensor(order)


def replica_checksums(moe):
    """two position-weighted integer checksums per arena tensor: equal on every rank iff the replicas are bit-identical"""
    out, names = [], []
    for k in "gdar":
        a = moe.arena(k)
        for tn, t in (("P", a.P), ("M", a.M), ("V", a.V), ("steps", a.steps), ("Bf", a.Bf), ("Bi", a.Bi)):
            names += [f"{k}.{tn}", f"{k}.{tn}"]
            v = t.contiguous().view(-1)
            bits = v.view(torch.int32) if v.dtype in (torch.float32, torch.int32) else v.to(torch.int64)
            bits = bits.to(torch.int64)
            w = (torch.arange(bits.numel(), device=bits.device) % 65521) + 1
            out += [bits.sum(), (bits * w).sum()]
    return torch.stack(out), names


def run_dp_parity(arch, dev, unbalanced=False, E=3, seed=7, group=None):
    """-> dict(ok, worst relative-L2 gradient errors g/d/a, replicas_identical, empty_rank_expert, fails[...]).
    Collective: every rank of ``group`` must call it."""
    from expertsim.config import Config
    from expertsim.train.loop import setup_moe_system
    rank, world = dist.get_rank(group), dist.get_world_size(group)
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
    moe.enable_data_parallel(group)
    batch, noise = orc.make_batch(arch, B, seed), orc.make_noise(arch, B, E, seed)
    if unbalanced:
        with torch.no_grad():
            gs, _ = orc.router_forward(st.router, batch["cond"], noise["gumbel"], orc.router_tau(ocfg["model"]["router"], 0))
        order = unbalanced_order(gs.argmax(dim=1), world)
        batch = {k: v[order] for k, v in batch.items()}
        noise = {k: v[order] for k, v in noise.items()}
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
    local_counts = [int(((masks[e] % world) == rank).sum()) for e in range(E)]
    b = sh(batch)
    got = moe.train_step(0, b["cond"], b["real_images"], b["true_positions"], b["std"], b["intensity"], noise=nz)
    torch.cuda.synchronize()
    fails = []
    if moe._last["idx"].cpu().tolist() != aux["idx"][mine].tolist():
        fails.append("routing of the shard differs")
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
    # the optimizer state advanced for every globally-live expert on EVERY rank (also where the rank held none of its rows)
    live = [int(c) >= 2 for c in aux["counts"]]
    for k in "gda":
        steps = moe.arena(k).steps.cpu().tolist()
        if steps != [1 if lv else 0 for lv in live]:
            fails.append(f"Adam step counters of arena {k}: {steps}, live experts {live}")
    # the generator's Adam touched every element of every live expert exactly once (first step from zero moments:
    # exp_avg == (1 - beta1) * gradient; twice would give 0.19 g, a missed rectangle 0) — the pipelined form included
    ag = moe.arena("g")
    c1 = torch.tensor(1.0) - torch.tensor(0.9)
    for e in range(E):
        m_want = ag.G[e] * float(c1) if live[e] else torch.zeros_like(ag.G[e])
        if not torch.allclose(ag.M[e], m_want, rtol=1e-5, atol=1e-30):
            bad = int((~torch.isclose(ag.M[e], m_want, rtol=1e-5, atol=1e-30)).sum())
            fails.append(f"generator exp_avg of expert {e} is not one Adam step of the reduced gradient ({bad} elements)")
    # replicas must stay BIT-identical after the step: parameters, moments, step counters, buffers
    chk, chk_names = replica_checksums(moe)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    identical = bool(torch.equal(lo, hi))
    if not identical:
        which = sorted({n for n, a_, b_ in zip(chk_names, lo.tolist(), hi.tolist()) if a_ != b_})
        fails.append(f"replicas diverged after the step: {which}")
    ok = torch.tensor([0 if fails else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    w = torch.tensor([worst["g"], worst["d"], worst["a"]], dtype=torch.float64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    empty = torch.tensor([1 if (unbalanced and local_counts[0] == 0 and live[0]) else 0], device=dev)
    dist.all_reduce(empty, op=dist.ReduceOp.MAX, group=group)
    return {"ok": int(ok) == 1, "g": float(w[0]), "d": float(w[1]), "a": float(w[2]), "replicas_identical": identical,
            "world": world, "global_batch": B, "unbalanced": bool(unbalanced), "rank_without_rows_of_a_live_expert": bool(int(empty)),
            "gen_loss": float(got["gen_loss"]), "oracle_gen_loss": want["gen_loss"], "fails": fails, "local_counts": local_counts,
            "pipelined_rects": int(getattr(moe, "n_pipelined_rects", 0))}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = {a for a in sys.argv[1:] if a.startswith("--")}
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    one_gpu = "--one-gpu" in flags or torch.cuda.device_count() < world
    dev = torch.device("cuda", 0 if one_gpu else local)
    torch.cuda.set_device(dev)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    arch = args[0] if args else "proton"
    n_exp = next((int(f.split("=", 1)[1]) for f in flags if f.startswith("--experts=")), 3)
    res = run_dp_parity(arch, dev, unbalanced="--unbalanced" in flags, E=n_exp)
    if rank == 0:
        print(f"dp parity {arch} world={world} B={res['global_batch']} backend={'gloo, ranks share cuda:0' if one_gpu else 'nccl'}"
              f"{' UNBALANCED (a rank holds no row of live expert 0: ' + str(res['rank_without_rows_of_a_live_expert']) + ')' if res['unbalanced'] else ''}: "
              f"worst relL2 g={res['g']:.3e} d={res['d']:.3e} a={res['a']:.3e}; replicas bit-identical: {res['replicas_identical']}; "
              f"gen_loss {res['gen_loss']:.6f} vs oracle {res['oracle_gen_loss']:.6f}; pipelined Adam rectangles: {res['pipelined_rects']}")
    for f in res["fails"]:
        print(f"[rank {rank}] FAIL {f}")
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
