#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/experiments/chunked_nccl_check.py
BucketedGradReducer.reduce_chunked over NCCL (coalesced row groups on the communication stream, one event per rectangle)
against a plain all-reduce of the same column range: bit-equal on every rank (two-rank sums commute), events usable from
the compute stream.  Also times the chunked against the one-message form for fc2's bucket shape [8, 23.7 M]."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200"))
from expertsim._reduce import BucketedGradReducer  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_MIN_NCHANNELS", "64")
    os.environ.setdefault("NCCL_BUFFSIZE", str(16 << 20))
    dist.init_process_group("nccl", device_id=dev)
    red = BucketedGradReducer(dist)
    ok = True
    for E, n, lo, hi, chunks in ((8, 1 << 20, 4096, (1 << 20) - 8192, 4), (3, 1 << 20, 0, 1 << 20, 4), (8, 26571844, 5632, 23875072, 4)):
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        G = torch.randn(E, n, device=dev, generator=g)
        want = G.clone()
        part = G[:, lo:hi].clone()
        dist.all_reduce(part)
        want[:, lo:hi] = part
        red.begin()
        rects = red.reduce_chunked(G, lo, hi, chunks)
        main_s = torch.cuda.current_stream()
        seen = []
        for e0, e1, c0, c1, ev in rects:
            main_s.wait_event(ev)
            seen.append(bool(torch.equal(G[e0:e1, c0:c1], want[e0:e1, c0:c1])))      # compute stream, behind the event only
        red.join()
        same = bool(torch.equal(G, want))
        ok = ok and same and all(seen)
        if rank == 0:
            print(f"E={E} n={n} [{lo},{hi}) chunks={len(rects)}: equal {same}, per-rectangle behind its event {seen}")
    # timing: one message group vs 4 / 8 rectangles, fc2's bucket of the proton generator
    G = torch.randn(8, 26571844, device=dev)
    lo, hi = 5632, 23875072
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, fn in (("one group", lambda: red.reduce(G, lo, hi)), ("4 rectangles", lambda: red.reduce_chunked(G, lo, hi, 4)),
                      ("8 rectangles", lambda: red.reduce_chunked(G, lo, hi, 8))):
        for it in range(6):
            if it == 1:
                torch.cuda.synchronize()
                dist.barrier()
                e0.record()
            red.begin()
            fn()
            red.join()
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"all-reduce of [8, {hi - lo}] fp32 ({8 * (hi - lo) * 4 / 1e6:.0f} MB), {label}: {e0.elapsed_time(e1) / 5:.3f} ms")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
