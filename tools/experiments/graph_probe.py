#!/usr/bin/env python
"""Probe: can MoEWrapper.train_step (≈400 launches on 4 streams per step) be captured into ONE CUDA graph, and what does a
replay cost against the eager step?  python tools/experiments/graph_probe.py [--arch proton]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="proton")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sysm = bench.System(a.arch, 8, 1024, dev, 0, 1, 4)
    for i in range(3):
        sysm.step(i)
    torch.cuda.synchronize()

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, host / n * 1e3

    ms, host = timed(lambda i: sysm.step(i), a.steps)
    print(f"eager : {ms:.3f} ms/step on the device, host issue {host:.3f} ms/step")
    static = [t.clone() for t in sysm.batch(0)]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sysm.step(0, host=tuple(static))          # warm-up on the capture stream (side streams get created here)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g, stream=s):
            out = sysm.step(0, host=tuple(static))
    except Exception as e:
        print("capture FAILED:", repr(e)[:2000])
        return
    torch.cuda.synchronize()
    loss0 = None
    for i in range(3):
        for dst, src in zip(static, sysm.batch(i)):
            dst.copy_(src)
        g.replay()
        torch.cuda.synchronize()
        print("replay", i, "gen_loss", float(out["gen_loss"]), "disc_loss", float(out["disc_loss"]))
    ms, host = timed(lambda i: g.replay(), a.steps)
    print(f"graph : {ms:.3f} ms/step on the device, host issue {host:.3f} ms/step")
    ms, host = timed(lambda i: sysm.step(i), a.steps)
    print(f"eager again: {ms:.3f} ms/step, gen_loss {float(sysm.step(0)['gen_loss'])}")


if __name__ == "__main__":
    main()
