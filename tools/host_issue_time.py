#!/usr/bin/env python
"""How long does the HOST need to issue one train step (no sync) vs. how long the GPU needs to run it?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")]
import torch
import bench
from expertsim.train.loop import setup_moe_system
from expertsim.utils.data import synthetic_showers
dev = torch.device("cuda:0")
cfg = bench.make_cfg("proton", 8)
moe = setup_moe_system(cfg, dev); moe.train()
d = synthetic_showers("proton", 1024, 0, dev)
def step():
    return moe.train_step(0, d["cond"], d["x"], d["positions"], d["std"], d["intensity"])
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host issue {1e3*(t1-t0):.1f} ms, until GPU done {1e3*(t2-t0):.1f} ms")
