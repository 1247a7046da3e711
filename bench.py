#!/usr/bin/env python
"""bench.py — train samples/s (and generated showers/s) of the ExpertSim MoE-GAN hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W              # this build (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the box's host cores

Workload (BASELINE.json configs[2], the MoE config the metric is quoted on; the same per-GPU work at every N = weak
scaling): proton ZDC 56x30, 8 experts, top-1 Gumbel router, SDI diversity + photon-sum + aux coordinate-regressor
losses, adaptive load-balancing router loss, batch 1024 per GPU, synthetic showers, random-init weights.  A "step" is
one full ``MoEWrapper.train_step`` (route, 2 G fwd, 4 D fwd, aux fwd, all backward passes, 3E+1 fused Adam updates,
NCCL gradient all-reduce when N>1).

One JSON line is printed by rank 0 (see the keys in ``main``).  ``value`` is timed with inputs resident in HBM;
``e2e`` feeds every step from pinned HOST memory and reads the step's loss back.  ``roofline`` is the tcgen05
implicit-GEMM family (dominant: >85% of step time), algorithmic FLOPs / CUDA-event time measured inside the timed region.
``cpu_baseline`` is the CPU oracle (a PyTorch restatement of the reference's algorithm; the reference itself is Python
+ PyTorch and /root/reference does not exist on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "generative-dnn-for-physics-simulations-cern_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# algorithmic forward FLOPs per sample (hook-counted on the reference modules)
F_G = {"proton": 4_744_939_008, "neutron": 1_732_593_152}
F_D = {"proton": 4_244_352, "neutron": 4_693_632}
F_A = {"proton": 15_745_920, "neutron": 41_210_368}
F_R = 23_296
TRAIN_FLOPS = {a: 6 * F_G[a] + 12 * F_D[a] + 3 * F_A[a] + 3 * F_R for a in F_G}   # as executed by the reference


def ncu_traffic():
    """DRAM bytes (read + write) per launch of the dominant kernel family, from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py); None if no capture is committed."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "gbs": d["hbm_gbs"], "src": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_cfg(arch, E):
    from expertsim.config import load_config
    shape = "[56,30]" if arch == "proton" else "[44,44]"
    return load_config(None, [f"model.architecture={arch}", f"model.n_experts={E}", f"dataset.zdc_type={arch}",
                              f"dataset.input_image_shape={shape}"])


def oracle_cfg(arch, E):
    import copy
    import oracle.expertsim_oracle as orc
    c = copy.deepcopy(orc.DEFAULT_CFG)
    c["model"]["architecture"], c["model"]["n_experts"] = arch, E
    return c


# ------------------------------------------------------------------------------------------------ CPU reference leg
def cpu_train_samples_per_s(arch, E, B, steps, warmup, threads):
    """The reference's algorithm (oracle port, plain PyTorch fp32 on the host cores) on a bounded sample of the workload:
    the same MoE step at batch ``B``; returns (samples/s, seconds per step)."""
    import oracle.expertsim_oracle as orc
    torch.set_num_threads(threads)
    st = orc.make_state(arch, E, 0, oracle_cfg(arch, E), identical_experts=True)
    ts = []
    for i in range(warmup + steps):
        batch, noise = orc.make_batch(arch, B, i), orc.make_noise(arch, B, E, i)
        t0 = time.perf_counter()
        orc.train_step(st, batch, noise, epoch=0)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return B / dt, dt


def cpu_showers_per_s(arch, n, threads):
    import oracle.expertsim_oracle as orc
    torch.set_num_threads(threads)
    sd = orc.make_weights(arch, "generator", 0)
    g = torch.Generator().manual_seed(0)
    z, c = torch.randn(n, 10, generator=g), torch.randn(n, 9, generator=g)
    orc.generate(arch, sd, z[:64], c[:64], batch_size=64)
    t0 = time.perf_counter()
    orc.generate(arch, sd, z, c, batch_size=256)
    return n / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_batch
    v, dt = cpu_train_samples_per_s(args.arch, args.experts, B, args.steps, min(args.warmup, 1), threads)
    shw = cpu_showers_per_s(args.arch, 512, threads)
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": round(v, 3), "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_cfg(args, note=f"bounded sample: the same MoE step at batch {B} on the host CPU"),
            "showers_per_sec": round(shw, 2),
            "cpu_baseline": {"value": round(v, 3), "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} MoE train steps at batch {B} (E={args.experts}, {args.arch}), fp32 PyTorch "
                                       f"restatement of the reference on {threads} threads; showers/s on 512 showers"},
            "e2e": {"value": round(v, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_cfg(args, note=None):
    c = {"workload": f"BASELINE configs[2]: {args.arch} ZDC MoE-GAN train step, E={args.experts} experts top-1, SDI+photon-sum+aux "
                     f"losses + ALB router loss, batch {args.batch}/GPU (global {args.batch * args.gpus}), dp{args.gpus}",
         "arch": args.arch, "n_experts": args.experts, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
         "parallelism": f"dp{args.gpus}",
         "l2": "per-step working set (weights 0.5 GB bf16 + activations > 4 GB) exceeds the 126 MB L2; inputs cycle over a pool"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------ B200 leg
def igemm_flops(name, a):
    """algorithmic FLOPs of one grouped GEMM launch from its logged scalar arguments."""
    ints = [x for x in a if isinstance(x, int)]
    if name in ("es_igemm_taps_fwd", "es_igemm_taps_wgrad"):
        # phase-folded x2-upsample conv: ALGORITHMIC = the un-folded direct convolution; executed MACs are
        # n_taps*C per output instead of KH*KW*C
        g = next(x for x in a if hasattr(x, "n_taps"))
        return g.alg_flops_per_row * ints[-1]
    if name in ("es_igemm_fwd", "es_igemm_wgrad"):
        g = next(x for x in a if hasattr(x, "Ho"))
        rows = ints[-1]
        return 2.0 * rows * g.Ho * g.Wo * g.N * g.KH * g.KW * g.C
    if name == "es_dense_dgrad":       # (N, K, n_groups, total_rows)
        return 2.0 * ints[-1] * ints[0] * ints[1]
    if name == "es_dense_wgrad":       # (..., N, K, n_groups, total_rows) — leading ints may be raw addresses / strides
        return 2.0 * ints[-1] * ints[-4] * ints[-3]
    return 0.0


def run_b200(args):
    import torch.distributed as dist
    from expertsim import _lib as L
    from expertsim.train.loop import setup_moe_system
    from expertsim.train.training_setup import setup_optimizers
    from expertsim.utils.data import synthetic_showers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner out of stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    if not L.device_ok():
        raise SystemExit("libexpertsim_b200.so carries sm_100a code only; no usable device")

    arch, E, B = args.arch, args.experts, args.batch
    cfg = make_cfg(arch, E)
    torch.manual_seed(0)
    moe = setup_moe_system(cfg, dev)
    with torch.no_grad():   # experts differ (deepcopy would make them identical): seeded 1% perturbation per expert
        gp = torch.Generator(device=dev).manual_seed(1)
        for k in "gda":
            P = moe.arena(k).P
            P.mul_(1.0 + 1e-2 * torch.randn(P.shape, generator=gp, device=dev))
    moe.mark_weights_changed()
    if world > 1:
        moe.enable_data_parallel()
        moe.overlap_grad_allreduce = False if args.no_overlap_allreduce else (args.overlap_mode == "eager" or "deferred")
    g_opt, d_opt, a_opt, r_opt = setup_optimizers(moe, cfg)
    moe.train()

    pool_n = args.pool
    data = synthetic_showers(arch, B * pool_n, seed=rank, device=dev)

    def batch(i):
        s = (i % pool_n) * B
        return (data["cond"][s:s + B], data["x"][s:s + B], data["positions"][s:s + B], data["std"][s:s + B],
                data["intensity"][s:s + B])

    def step(i):
        c, x, pos, sd, it = batch(i)
        return moe.train_step(0, c, x, pos, sd, it, a_opt, g_opt, d_opt, r_opt, None, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for i in range(args.warmup):
        step(i)
    if args.ncu_step:
        # profiling aid: `ncu --profile-from-start off ... bench.py --ncu-step` captures exactly one train step (and one
        # inference batch); nothing measured under the profiler is ever reported as a bench value
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(args.warmup)
        if args.ncu_step > 1:
            moe.eval()
            moe.generate(torch.randn(args.infer_batch, 9, device=dev), chunk=args.infer_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- timed region: K steps, inputs resident in HBM
    names = {"es_igemm_fwd", "es_igemm_wgrad", "es_igemm_taps_fwd", "es_igemm_taps_wgrad", "es_dense_dgrad", "es_dense_wgrad"}
    L.profile = {"names": names, "log": []}
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    n0 = L.n_calls
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        m = step(args.warmup + i)
    e1.record()
    barrier()
    launches = L.n_calls - n0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    plog, L.profile = L.profile["log"], None
    loss = float(m["gen_loss"])
    if not (loss == loss):
        raise SystemExit("non-finite loss in the timed region")
    value = args.steps * B * world / (ms / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 grouped implicit GEMM), per launch, inside the timed region
    if args.per_launch and rank == 0:      # tuning aid: every timed GEMM launch of the last step, to stderr
        per = len(plog) // args.steps
        for name, a, s, e in plog[-per:]:
            g = next((x for x in a if hasattr(x, "Ho")), None)
            if g is not None and hasattr(g, "n_taps"):
                geo = f"Hs{g.Hs} C{g.C} taps{g.n_taps} N{g.N} Ho{g.Ho} (folded)"
            elif g is not None:
                geo = f"Hs{g.Hs} C{g.C} {g.KH}x{g.KW} N{g.N} Ho{g.Ho}"
            else:
                geo = str([x for x in a if isinstance(x, int)][:3])
            t = s.elapsed_time(e)
            print(f"  {name:16s} {geo:34s} {t:8.3f} ms {igemm_flops(name, a) / t / 1e9:8.1f} TFLOP/s", file=sys.stderr)
    fam = {}
    for name, a, s, e in plog:
        key = name
        f = fam.setdefault(key, [0.0, 0.0, 0])
        f[0] += igemm_flops(name, a)
        f[1] += s.elapsed_time(e)
        f[2] += 1
    tc = [fam[k] for k in ("es_igemm_fwd", "es_igemm_wgrad", "es_igemm_taps_fwd", "es_igemm_taps_wgrad") if k in fam]
    tc_flops, tc_ms, tc_n = (sum(x[i] for x in tc) for i in range(3))
    pk = peaks()
    roof = {"bound": "tensor", "kernel": "igemm_persist (grouped bf16 tcgen05 implicit GEMM family: igemm_fwd_kernel + igemm_wgrad_kernel, incl. the x2-upsample-folded tap-table launches; algorithmic = un-folded direct-conv FLOPs)",
            "achieved": round(tc_flops / (tc_ms * 1e-3) / 1e12, 2) if tc_ms else None, "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": round(tc_flops / (tc_ms * 1e-3) / 1e12 / pk["tflops"], 4) if tc_ms else None,
            "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch"), "traffic_source": (ncu_traffic() or {}).get("source"),
            "algorithmic_flops_per_launch": round(tc_flops / max(tc_n, 1)),
            "peak_source": f"{pk['src']} sustained bf16 (MEASURED_PEAKS.json)", "launches": tc_n,
            "share_of_step": round(tc_ms / (e0.elapsed_time(e1)), 4),
            "families": {k: {"tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 2) if v[1] else None, "ms_per_step": round(v[1] / args.steps, 3),
                             "launches_per_step": v[2] // args.steps} for k, v in fam.items()}}
    step_tflops = value / world * TRAIN_FLOPS[arch] / 1e12
    roof["step_algorithmic_tflops_per_gpu"] = round(step_tflops, 2)
    roof["step_frac_of_peak"] = round(step_tflops / pk["tflops"], 4)

    # ---- e2e: the same step fed from pinned host memory, loss read back every step
    host = [tuple(t.cpu().pin_memory() for t in batch(i)) for i in range(min(pool_n, 8))]
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_step(i):
        c, x, pos, sd, it = (t.to(dev, non_blocking=True) for t in host[i % len(host)])
        mm = moe.train_step(0, c, x, pos, sd, it, a_opt, g_opt, d_opt, r_opt, None, dev)
        return float(mm["gen_loss"])     # device -> host read of the step's loss

    for i in range(2):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": round(args.steps * B * world / (ms_e2e / 1e3), 2), "unit": "samples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)}

    # ---- batch inference: generated showers/s (router -> partition -> 8 expert generators -> expm1), device resident
    moe.eval()
    n_inf = args.infer_batch
    cond_inf = torch.randn(n_inf, 9, device=dev)
    for _ in range(2):
        moe.generate(cond_inf, chunk=n_inf)
    barrier()
    e0.record()
    for _ in range(args.infer_iters):
        out = moe.generate(cond_inf, chunk=n_inf)
    e1.record()
    barrier()
    ms_inf = max_over_ranks(e0.elapsed_time(e1))
    showers = args.infer_iters * n_inf * world / (ms_inf / 1e3)
    host_cond = cond_inf.cpu().pin_memory()
    out_host = torch.empty(n_inf, *moe.image_shape, pin_memory=True)
    barrier()
    e0.record()
    for _ in range(args.infer_iters):
        o = moe.generate(host_cond, chunk=n_inf)
        out_host.copy_(o, non_blocking=True)
    e1.record()
    barrier()
    ms_inf2 = max_over_ranks(e0.elapsed_time(e1))
    inference = {"showers_per_sec": round(showers, 1), "batch_per_gpu": n_inf, "iters": args.infer_iters,
                 "e2e_showers_per_sec": round(args.infer_iters * n_inf * world / (ms_inf2 / 1e3), 1),
                 "algorithmic_tflops_per_gpu": round(showers / world * (F_G[arch] + F_R) / 1e12, 2),
                 "frac_of_peak": round(showers / world * (F_G[arch] + F_R) / 1e12 / pk["tflops"], 4)}
    moe.train()

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_train_samples_per_s(arch, E, args.cpu_batch, 4, 1, threads)
        cpu = {"value": round(v, 3), "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"4 timed MoE train steps (after 1 warm-up) at batch {args.cpu_batch}, E={E}, {arch}: fp32 PyTorch "
                         f"restatement of the reference (oracle/) on {threads} host threads, {dt:.1f} s/step"}
    line = {"metric": "train_samples_per_sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_cfg(args),
            "showers_per_sec": inference["showers_per_sec"], "inference": inference, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "loss_check": round(loss, 6)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="proton", choices=["proton", "neutron"])
    ap.add_argument("--experts", type=int, default=8)
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU per step")
    ap.add_argument("--pool", type=int, default=16, help="distinct resident input batches cycled through")
    ap.add_argument("--infer-batch", type=int, default=8192)
    ap.add_argument("--infer-iters", type=int, default=5)
    ap.add_argument("--cpu-batch", type=int, default=128, help="batch of the bounded CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-mode", default="deferred", choices=["deferred", "eager"])
    ap.add_argument("--no-overlap-allreduce", action="store_true",
                    help="A/B switch: one whole-arena gradient all-reduce after backward instead of the overlapped layer buckets")
    ap.add_argument("--per-launch", action="store_true", help="print every timed GEMM launch of the last step to stderr")
    ap.add_argument("--ncu-step", type=int, default=0, help="1: cudaProfilerStart/Stop around one train step; 2: + one inference batch")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
