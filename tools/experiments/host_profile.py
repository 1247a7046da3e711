#!/usr/bin/env python
"""Where does the HOST time of one eager train step go (≈ 400 C-ABI launches + allocations)?  cProfile over 10 steps."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sysm = bench.System(sys.argv[1] if len(sys.argv) > 1 else "proton", 8, 1024, dev, 0, 1, 4)
for i in range(3):
    sysm.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    sysm.step(i)
host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print(f"host issue time per step (no profiler): {host * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    sysm.step(i)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
