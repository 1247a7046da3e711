#!/usr/bin/env python
"""bench.py — train samples/s (and generated showers/s) of the ExpertSim MoE-GAN hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W              # this build (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the box's host cores

Workload (BASELINE.json configs[2], the MoE config the metric is quoted on; the same per-GPU work at every N = weak
scaling): proton ZDC 56x30, 8 experts, top-1 Gumbel router, SDI diversity + photon-sum + aux coordinate-regressor
losses, adaptive load-balancing router loss, batch 1024 per GPU, synthetic showers, random-init weights.  A "step" is
one full ``MoEWrapper.train_step`` (route, 2 G fwd, 4 D fwd, aux fwd, all backward passes, 3E+1 fused Adam updates,
NCCL gradient all-reduce when N>1).

One JSON line is printed by rank 0 (see the keys in ``main``).  ``value`` is timed with inputs resident in HBM and NO
per-launch instrumentation; ``e2e`` feeds every step from pinned HOST memory and reads the step's loss back.
``roofline`` is the tcgen05 implicit-GEMM family (dominant: ~55% of step time), measured in a SECOND pass with CUDA
events around every GEMM launch: algorithmic (un-folded direct-conv) FLOPs, EXECUTED FLOPs (after upsample folding),
and the tensor-pipe utilisation of the committed ncu capture.  ``roofline_hbm`` holds the achieved GB/s of the loss-tail /
gating / Adam kernels at a size that does not fit the L2.  ``extra`` carries the other BASELINE configs (neutron ZDC
44x44 = configs[3] per-GPU slice, single expert = configs[1]); at N>1 ``dp_parity`` is the on-device proof that the
sharded step equals the global-batch step (proton + neutron/SyncBN, balanced and with a rank holding no row of a live
expert) and ``comm.exposed_ms`` the step time the collectives add.  ``cpu_baseline`` / ``--impl reference`` time the
UNMODIFIED reference (``oracle/_ref``, a verbatim copy made by oracle/make_ref.py; the oracle port only if that copy is
absent) on the box's host cores, on a bounded sample of the same workload.
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
def reference_on_cpu(arch, E, B, steps, warmup, threads, showers=512):
    """The UNMODIFIED reference (oracle/_ref or /root/reference) on the host cores, in a process of its own (its package is
    also called ``expertsim``): MoEWrapper.train_step with its own torch.optim.Adam optimizers and random draws, batch
    ``B``; get_predictions_from_generator_results for showers/s.  -> dict, or None if no reference copy is present."""
    import oracle.ref_shim as shim
    if shim.find_reference() is None:
        return None
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "bench", "--arch", arch, "--experts", str(E), "--batch", str(B),
           "--steps", str(steps), "--warmup", str(warmup), "--threads", str(threads), "--showers", str(showers)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"reference run failed:\n{r.stderr[-2000:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def port_on_cpu(arch, E, B, steps, warmup, threads):
    """Fallback when no reference copy travelled: the oracle port (plain PyTorch fp32 restatement) on the host cores."""
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
    sd = orc.make_weights(arch, "generator", 0)
    g = torch.Generator().manual_seed(0)
    z, c = torch.randn(512, 10, generator=g), torch.randn(512, 9, generator=g)
    orc.generate(arch, sd, z[:64], c[:64], batch_size=64)
    t0 = time.perf_counter()
    orc.generate(arch, sd, z, c, batch_size=256)
    return {"samples_per_s": B / dt, "s_per_step": dt, "showers_per_s": 512 / (time.perf_counter() - t0), "threads": threads}


def cpu_leg(arch, E, B, steps, warmup, threads):
    r = reference_on_cpu(arch, E, B, steps, warmup, threads)
    if r is not None:
        return r, "reference", "the unmodified reference (oracle/_ref: MoEWrapper.train_step, torch.optim.Adam, eager fp32 PyTorch)"
    return port_on_cpu(arch, E, B, steps, warmup, threads), "port", "fp32 PyTorch restatement of the reference (oracle/; no reference copy present)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_batch
    r, kind, what = cpu_leg(args.arch, args.experts, B, args.steps, args.warmup, threads)
    v = r["samples_per_s"]
    sample = (f"{args.steps} timed MoE train steps (after {args.warmup} warm-up) of {what} at batch {B} — a bounded sample of the "
              f"batch-{args.batch} workload — E={args.experts}, {args.arch}, {threads} host threads, {r['s_per_step']:.2f} s/step; "
              f"showers/s on 512 showers")
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": round(v, 3), "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["s_per_step"] * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_cfg(args), "showers_per_sec": round(r["showers_per_s"], 2),
            "cpu_baseline": {"value": round(v, 3), "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
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
GEMM_CALLS = {"es_igemm_fwd", "es_igemm_wgrad", "es_igemm_taps_fwd", "es_igemm_taps_wgrad", "es_dense_dgrad", "es_dense_wgrad",
              "es_igemm_fwd_sums", "es_igemm_taps_fwd_sums"}      # *_sums: the same GEMMs with the GroupNorm sums in the epilogue


def igemm_flops(name, a):
    """(algorithmic, executed) FLOPs of one grouped GEMM launch from its logged scalar arguments.  ALGORITHMIC = the un-folded
    direct convolution; EXECUTED = the MACs the tensor cores really perform: n_taps*C per output of a
    folded class instead of KH*KW*C (equal for plain convolutions and the dense layers)."""
    ints = [x for x in a if isinstance(x, int)]
    name = name.replace("_sums", "")
    if name in ("es_igemm_taps_fwd", "es_igemm_taps_wgrad"):
        g = next(x for x in a if hasattr(x, "n_taps"))
        rows = ints[-1]
        return g.alg_flops_per_row * rows, 2.0 * rows * g.Ho * g.Wo * g.N * g.n_taps * g.C
    if name in ("es_igemm_fwd", "es_igemm_wgrad"):
        g = next(x for x in a if hasattr(x, "Ho"))
        f = 2.0 * ints[-1] * g.Ho * g.Wo * g.N * g.KH * g.KW * g.C
        return f, f
    if name == "es_dense_dgrad":       # (N, K, n_groups, total_rows)
        f = 2.0 * ints[-1] * ints[0] * ints[1]
        return f, f
    if name == "es_dense_wgrad":       # (..., N, K, n_groups, total_rows) — leading ints may be raw addresses / strides
        f = 2.0 * ints[-1] * ints[-4] * ints[-3]
        return f, f
    return 0.0, 0.0


class System:
    """one MoE system + optimizers + a pool of resident synthetic batches"""

    def __init__(self, arch, E, B, dev, rank, world, pool_n, dp=True, overlap="deferred", graph=False):
        from expertsim.train.loop import setup_moe_system
        from expertsim.train.training_setup import setup_optimizers
        from expertsim.utils.data import synthetic_showers
        self.arch, self.E, self.B, self.dev, self.world, self.pool_n = arch, E, B, dev, world, pool_n
        cfg = make_cfg(arch, E)
        torch.manual_seed(0)
        self.moe = moe = setup_moe_system(cfg, dev)
        with torch.no_grad():   # experts differ (deepcopy would make them identical): seeded 1% perturbation per expert
            gp = torch.Generator(device=dev).manual_seed(1)
            for k in "gda":
                P = moe.arena(k).P
                P.mul_(1.0 + 1e-2 * torch.randn(P.shape, generator=gp, device=dev))
        moe.mark_weights_changed()
        if world > 1 and dp:
            moe.enable_data_parallel()
            moe.overlap_grad_allreduce = overlap
        self.opts = setup_optimizers(moe, cfg)
        if graph:
            moe.enable_cuda_graph()
        moe.train()
        self.data = synthetic_showers(arch, B * pool_n, seed=rank, device=dev)

    def batch(self, i):
        s, B, d = (i % self.pool_n) * self.B, self.B, self.data
        return d["cond"][s:s + B], d["x"][s:s + B], d["positions"][s:s + B], d["std"][s:s + B], d["intensity"][s:s + B]

    def step(self, i, host=None):
        g_o, d_o, a_o, r_o = self.opts
        c, x, pos, sd, it = host if host is not None else self.batch(i)
        return self.moe.train_step(0, c, x, pos, sd, it, a_o, g_o, d_o, r_o, None, self.dev)


def run_b200(args):
    import torch.distributed as dist
    from expertsim import _lib as L

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
        # the gradient buckets run when the SMs are otherwise idle (after backward): two GPUs talk over point-to-point
        # NVLink rings, which need more channels and larger staging buffers than NCCL's defaults to fill the 18 links
        # (measured at N = 2: exposed communication 1.93 -> 1.71 ms per step).  N > 2 keeps NCCL's own choice (NVLS /
        # tree over the NVSwitch) as in the round-1 4- and 8-GPU runs: the setting was never measured there.
        if world == 2:
            os.environ.setdefault("NCCL_MIN_NCHANNELS", "64")
            os.environ.setdefault("NCCL_BUFFSIZE", str(16 << 20))
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    if not L.device_ok():
        raise SystemExit("libexpertsim_b200.so carries sm_100a code only; no usable device")
    arch, E, B = args.arch, args.experts, args.batch
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        """n calls of fn(i) between barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks"""
        barrier()
        e0.record()
        out = None
        for i in range(n):
            out = fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), out

    # ---- N>1: the sharded step equals the global-batch step, checked ON THE DEVICE before anything is timed (small batch,
    # CPU oracle as the checker): proton and neutron (SyncBN), balanced and with a rank that holds no row of a live expert
    if args.pipeline_adam is not None:
        from expertsim.models.moe import MoEWrapper
        MoEWrapper.pipeline_adam = bool(args.pipeline_adam)
    dp_parity = None
    if world > 1 and not args.no_dp_parity:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from dp_parity import run_dp_parity
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
        dp_parity, bad = {}, []
        for a_ in ("proton", "neutron"):
            for unb in (False, True):
                r = run_dp_parity(a_, dev, unbalanced=unb)
                key = f"{a_}{'_unbalanced' if unb else ''}"
                dp_parity[key] = {k: (float(f"{r[k]:.3e}") if isinstance(r[k], float) else r[k]) for k in
                                  ("g", "d", "a", "replicas_identical", "rank_without_rows_of_a_live_expert", "pipelined_rects", "ok")}
                if not r["ok"]:
                    bad.append((key, r["fails"][:4]))
        dp_parity.update(g=max(v["g"] for v in dp_parity.values()), d=max(v["d"] for v in dp_parity.values()),
                         a=max(v["a"] for v in dp_parity.values()),
                         replicas_identical=all(v["replicas_identical"] for v in dp_parity.values()),
                         reducer="deferred", global_batch=24 if 24 % world == 0 else 8 * world, n_experts=3,
                         bound={"g": 0.2, "d": 2e-3, "a": 2e-3, "what": "relative L2 of the all-reduced gradients vs the CPU oracle's "
                                "global-batch step (images injected); replicas bit-identical (P, M, V, steps, buffers)"})
        if bad:
            raise SystemExit(f"data-parallel parity broken: {bad}")
        torch.cuda.empty_cache()

    overlap = False if args.no_overlap_allreduce else args.overlap_mode
    sysm = System(arch, E, B, dev, rank, world, args.pool, overlap=overlap, graph=args.cuda_graph)
    moe = sysm.moe
    for i in range(args.warmup):
        sysm.step(i)
    if args.ncu_step:
        # profiling aid: `ncu --profile-from-start off ... bench.py --ncu-step` captures exactly one train step (and one
        # inference batch); nothing measured under the profiler is ever reported as a bench value
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sysm.step(args.warmup)
        if args.ncu_step > 1:
            moe.eval()
            moe.generate(torch.randn(args.infer_batch, 9, device=dev), chunk=args.infer_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- headline timed region: K steps, inputs resident in HBM, no instrumentation inside
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = L.n_calls
    ms, m = timed(lambda i: sysm.step(args.warmup + i), args.steps)
    launches = L.n_calls - n0
    clk = clocks.stop() if rank == 0 else None
    loss = float(m["gen_loss"])
    if not (loss == loss):
        raise SystemExit("non-finite loss in the timed region")
    value = args.steps * B * world / (ms / 1e3)

    # ---- roofline pass (second pass, NOT the headline): CUDA events around every GEMM launch of n_prof steps
    n_prof = max(1, min(args.steps, 5))
    L.profile = {"names": GEMM_CALLS, "log": []}
    ms_prof, _ = timed(lambda i: sysm.step(args.warmup + args.steps + i), n_prof)
    plog, L.profile = L.profile["log"], None
    if args.per_launch and rank == 0:      # tuning aid: every timed GEMM launch of the last step, to stderr
        per = len(plog) // n_prof
        for name, a, s_, e_ in plog[-per:]:
            g = next((x for x in a if hasattr(x, "Ho")), None)
            if g is not None and hasattr(g, "n_taps"):
                geo = f"Hs{g.Hs} C{g.C} taps{g.n_taps} N{g.N} Ho{g.Ho} (folded)"
            elif g is not None:
                geo = f"Hs{g.Hs} C{g.C} {g.KH}x{g.KW} N{g.N} Ho{g.Ho}"
            else:
                geo = str([x for x in a if isinstance(x, int)][:3])
            t = s_.elapsed_time(e_)
            fa, fe = igemm_flops(name, a)
            print(f"  {name:16s} {geo:34s} {t:8.3f} ms {fa / t / 1e9:8.1f} TFLOP/s algorithmic {fe / t / 1e9:8.1f} executed", file=sys.stderr)
    fam = {}
    for name, a, s_, e_ in plog:
        f = fam.setdefault(name.replace("_sums", ""), [0.0, 0.0, 0.0, 0])
        fa, fe = igemm_flops(name, a)
        f[0] += fa
        f[1] += fe
        f[2] += s_.elapsed_time(e_)
        f[3] += 1
    tc = [fam[k] for k in ("es_igemm_fwd", "es_igemm_wgrad", "es_igemm_taps_fwd", "es_igemm_taps_wgrad") if k in fam]
    tc_alg, tc_exe, tc_ms, tc_n = (sum(x[i] for x in tc) for i in range(4))
    tr = ncu_traffic() or {}
    tfl = lambda f, t: round(f / (t * 1e-3) / 1e12, 2) if t else None
    roof = {"bound": "tensor",
            "kernel": "igemm_persist (grouped bf16 tcgen05 implicit GEMM family: igemm_tma_pair_kernel / igemm_tma_strip_kernel / "
                      "igemm_fwd_kernel (fc2) + igemm_wgrad_tma_kernel / igemm_wgrad_strip_kernel, incl. the upsample-folded tap-table launches)",
            "achieved": tfl(tc_alg, tc_ms), "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": round(tc_alg / (tc_ms * 1e-3) / 1e12 / pk["tflops"], 4) if tc_ms else None,
            "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
            "algorithmic_flops_per_launch": round(tc_alg / max(tc_n, 1)),
            "executed_flops_per_launch": round(tc_exe / max(tc_n, 1)),
            "executed_achieved": tfl(tc_exe, tc_ms),
            "executed_frac": round(tc_exe / (tc_ms * 1e-3) / 1e12 / pk["tflops"], 4) if tc_ms else None,
            "tensor_pipe_active_pct": tr.get("tensor_pipe_active_pct_time_weighted"),
            "note": "frac = ALGORITHMIC (un-folded direct-conv) FLOPs / time / peak: upsample folding removes 2.56x (conv1) / 1.39x (conv2) "
                    "of the MACs, so it is not tensor-pipe utilisation; executed_frac = FLOPs the tensor cores really perform / time / peak; "
                    "tensor_pipe_active_pct = sm__pipe_tensor_cycles_active, time-weighted over the committed ncu capture",
            "peak_source": f"{pk['src']} sustained bf16 (MEASURED_PEAKS.json)", "launches": tc_n, "timed_steps": n_prof,
            "timing": "CUDA events around each launch on the launching stream, second pass after the headline timed region",
            "share_of_step": round(tc_ms / ms_prof, 4) if ms_prof else None,
            "families": {k: {"tflops": tfl(v[0], v[2]), "executed_tflops": tfl(v[1], v[2]), "ms_per_step": round(v[2] / n_prof, 3),
                             "launches_per_step": v[3] // n_prof} for k, v in fam.items()}}
    step_tflops = value / world * TRAIN_FLOPS[arch] / 1e12
    roof["step_algorithmic_tflops_per_gpu"] = round(step_tflops, 2)
    roof["step_frac_of_peak"] = round(step_tflops / pk["tflops"], 4)

    # ---- e2e: the same step fed from pinned host memory, loss read back every step
    host = [tuple(t.cpu().pin_memory() for t in sysm.batch(i)) for i in range(min(args.pool, 8))]
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_step(i):
        hb = tuple(t.to(dev, non_blocking=True) for t in host[i % len(host)])
        return float(sysm.step(i, host=hb)["gen_loss"])     # device -> host read of the step's loss

    for i in range(2):
        e2e_step(i)
    ms_e2e, _ = timed(e2e_step, args.steps)
    e2e = {"value": round(args.steps * B * world / (ms_e2e / 1e3), 2), "unit": "samples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)}

    # ---- N>1: what the collectives cost — the same steps with every data-parallel collective turned into a no-op
    comm = None
    if world > 1:
        moe.set_collectives_enabled(False)
        for i in range(2):
            sysm.step(i)
        ms_nc, _ = timed(lambda i: sysm.step(i), max(3, min(args.steps, 10)))
        moe.set_collectives_enabled(True)
        per_step, per_nc = ms / args.steps, ms_nc / max(3, min(args.steps, 10))
        gb = sum(moe.arena(k).G.numel() for k in "gdar") * 4
        red = getattr(moe, "_reducer", None)
        sent = gb - moe.arena("g").G.numel() * 4 + red.bytes_sent if red is not None and red.bytes_sent else gb
        comm = {"exposed_ms": round(per_step - per_nc, 3), "ms_per_step_without_collectives": round(per_nc, 3),
                "allreduce_bytes_per_step": sent, "gradient_bytes_per_step_fp32": gb, "mode": str(overlap),
                "fc2_bucket_dtype": "bf16" if red is not None and red.compress_min_cols is not None else "fp32",
                "adam_pipelined_rects": int(getattr(moe, "n_pipelined_rects", 0)),
                "what": "step time minus the time of the same step with all collectives as no-ops (gradient buckets, per-expert loss "
                        "sums, counts, SyncBN statistics); the exposed part is dominated by fc2's gradient bucket (88 % of the "
                        "generator's gradient bytes), produced last; adam_pipelined_rects > 0: the optimizer pass over that bucket "
                        "runs rectangle by rectangle behind its chunked all-reduce"}

    # ---- batch inference: generated showers/s (router -> partition -> 8 expert generators -> expm1), device resident
    inference = bench_inference(moe, args, dev, world, pk, arch, timed)
    moe.train()

    # ---- HBM-bound kernels (loss tails, gating, Adam) at a size the L2 does not hold; measured here, not taken from a file
    roof_hbm = None
    if not args.no_hbm_kernels:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from bench_hbm_kernels import measure
        roof_hbm = measure(dev, peak=pk["gbs"])
        for v in roof_hbm.values():
            v["peak_source"] = f"{pk['src']} HBM copy rate (MEASURED_PEAKS.json)"
        barrier()

    # ---- the other BASELINE configs, per-GPU slice, short runs: neutron ZDC 44x44 (configs[3]) and one expert (configs[1])
    extra = {}
    if not args.no_extra:
        del sysm, moe
        torch.cuda.empty_cache()
        for key, (a_, e_) in (("neutron", ("neutron", E)), ("e1", ("proton", 1))):
            if (a_, e_) == (arch, E):
                continue
            try:
                extra[key] = short_bench(a_, e_, B, dev, rank, world, args, pk, timed, overlap)
            except Exception as ex:      # a side line must never cost the headline line (the error is reported in its place)
                extra[key] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r, kind, what = cpu_leg(arch, E, args.cpu_batch, 3, 1, threads)
        cpu = {"value": round(r["samples_per_s"], 3), "unit": "samples/s", "cores": threads, "kind": kind,
               "showers_per_sec": round(r["showers_per_s"], 2),
               "sample": f"3 timed MoE train steps (after 1 warm-up) of {what} at batch {args.cpu_batch} — a bounded sample of the "
                         f"batch-{B} workload — E={E}, {arch}, {threads} host threads, {r['s_per_step']:.2f} s/step"}
    line = {"metric": "train_samples_per_sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_cfg(args),
            "showers_per_sec": inference["showers_per_sec"], "inference": inference, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "clocks": clk, "loss_check": round(loss, 6),
            "extra": extra}
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
    if comm is not None:
        line["comm"] = comm
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_inference(moe, args, dev, world, pk, arch, timed):
    moe.eval()
    n_inf = args.infer_batch
    cond_inf = torch.randn(n_inf, 9, device=dev)
    for _ in range(2):
        moe.generate(cond_inf, chunk=n_inf)
    ms_inf, _ = timed(lambda i: moe.generate(cond_inf, chunk=n_inf), args.infer_iters)
    showers = args.infer_iters * n_inf * world / (ms_inf / 1e3)
    host_cond = cond_inf.cpu().pin_memory()
    out_host = torch.empty(n_inf, *moe.image_shape, pin_memory=True)

    def e2e_gen(i):
        o = moe.generate(host_cond, chunk=n_inf)
        out_host.copy_(o, non_blocking=True)

    ms_inf2, _ = timed(e2e_gen, args.infer_iters)
    return {"showers_per_sec": round(showers, 1), "batch_per_gpu": n_inf, "iters": args.infer_iters,
            "e2e_showers_per_sec": round(args.infer_iters * n_inf * world / (ms_inf2 / 1e3), 1),
            "algorithmic_tflops_per_gpu": round(showers / world * (F_G[arch] + F_R) / 1e12, 2),
            "frac_of_peak": round(showers / world * (F_G[arch] + F_R) / 1e12 / pk["tflops"], 4)}


def short_bench(arch, E, B, dev, rank, world, args, pk, timed, overlap):
    """train samples/s + showers/s of another BASELINE configuration (weak-scaling slice of batch B per GPU)"""
    sysm = System(arch, E, B, dev, rank, world, min(args.pool, 8), overlap=overlap, graph=args.cuda_graph)
    n = max(3, min(args.steps, 10))
    for i in range(3):
        sysm.step(i)
    ms, m = timed(lambda i: sysm.step(3 + i), n)
    v = n * B * world / (ms / 1e3)
    inf = bench_inference(sysm.moe, args, dev, world, pk, arch, timed)
    out = {"workload": f"{arch} ZDC MoE-GAN train step, E={E}, batch {B}/GPU (global {B * world}), dp{world}",
           "train_samples_per_sec": round(v, 2), "ms_per_step": round(ms / n, 3), "steps": n, "warmup": 3,
           "showers_per_sec": inf["showers_per_sec"], "loss_finite": bool(float(m["gen_loss"]) == float(m["gen_loss"])),
           "step_frac_of_peak": round(v / world * TRAIN_FLOPS[arch] / 1e12 / pk["tflops"], 4)}
    del sysm
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--cpu-batch", type=int, default=256, help="batch of the bounded CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="MoEWrapper.enable_cuda_graph(): replay the step as one CUDA graph")
    ap.add_argument("--no-extra", action="store_true", help="skip the neutron / single-expert lines of `extra`")
    ap.add_argument("--no-hbm-kernels", action="store_true", help="skip the roofline_hbm micro-measurements")
    ap.add_argument("--no-dp-parity", action="store_true", help="N>1: skip the on-device data-parallel parity check")
    ap.add_argument("--overlap-mode", default="deferred", choices=["deferred", "eager"])
    ap.add_argument("--pipeline-adam", type=int, default=None, choices=[0, 1],
                    help="A/B switch (N>1): optimizer pass over fc2's gradient bucket pipelined behind its chunked all-reduce")
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
