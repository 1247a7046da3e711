"""MoE wrapper: the training step and batch inference of the Mixture-of-Experts conditional GAN.

Drop-in for ``MoEWrapper`` (expertsim/models/moe.py:14-699 of the reference): same constructor, same
``train_step`` / ``evaluate`` signatures, same metric keys, ``.generators/.discriminators/.aux_regs`` module lists and
``.router``.  The arithmetic is NOT PyTorch: every numeric step is a grouped sm_100a kernel called through the C-ABI
(include/expertsim_b200.h).  Where the reference loops ``for i in range(n_experts)`` over eager modules with two
device->host syncs per expert (moe.py:121,522,562), this step is sync-free: the router kernel emits a stable
token->expert permutation and device-side group tables, every layer runs ONCE for all experts (ragged groups), and
the only thing that depends on the routing outcome lives on the device.

Phase order inside a step (experts are independent, so the reference's per-expert D-step/G-step sequence can be
regrouped across experts without changing any result; SURVEY.md §7 "serial semantics"):
    route -> G(z1),G(z2) [one two-pass launch per layer] -> D(real),D(fake1) -> hinge -> D backward -> Adam(D)
          -> D'(fake1),D'(fake2) with the UPDATED D -> aux(fake1) -> loss tails -> D'/aux backward to the images
          -> G backward -> Adam(G), Adam(aux) -> router loss backward -> Adam(router)
Spectral-norm power iterations advance once per D forward (4x per step), exactly like the reference's hook.

Data parallelism (new; the reference is single-device): ``enable_data_parallel()`` makes every loss normaliser a
GLOBAL-batch quantity (per-expert partial sums are all-reduced before the gradient kernels run) and sum-all-reduces
the gradient arenas over NCCL before each fused Adam, which reproduces the single-GPU global-batch step.
"""
from __future__ import annotations

import copy
import os
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .._arena import Arena
from .._nets import ZP, engine_for, zeros as _zeros

IMAGE_SHAPE = {"proton": (56, 30), "neutron": (44, 44)}


def _f(v, default=0.0):
    return default if v is None else float(v)


class MoEWrapper(nn.Module):
    # data-parallel steps: the optimizer pass over the widest gradient bucket is pipelined behind its chunked all-reduce
    # (_adam_pipelined; measured at N = 2: 31.79 -> 31.32 ms per step, exposed communication 1.58 -> 0.83 ms);
    # ES_DP_PIPELINE_ADAM=0 falls back to join-then-one-launch
    pipeline_adam = os.environ.get("ES_DP_PIPELINE_ADAM", "1") == "1"
    PIPELINE_CHUNKS = int(os.environ.get("ES_DP_PIPELINE_CHUNKS", "4"))
    PIPELINE_MIN_FLOATS = 8 << 20           # below 32 MB the bucket is not worth the extra launches

    def __init__(self, generator, discriminator, aux_reg, router, n_experts, cfg, image_shape):
        super().__init__()
        self.cfg = cfg
        self.n_experts = int(n_experts)
        self.noise_dim = cfg.model.noise_dim
        self.image_shape = tuple(image_shape)
        self.arch = generator.ARCH
        if self.image_shape != IMAGE_SHAPE[self.arch]:
            raise ValueError(f"{self.arch} generator produces {IMAGE_SHAPE[self.arch]} images, config says {self.image_shape}")
        # identical initial weights in every expert, as the reference's deepcopy (moe.py:29-31)
        self.generators = nn.ModuleList([copy.deepcopy(generator) for _ in range(self.n_experts)])
        self.discriminators = nn.ModuleList([copy.deepcopy(discriminator) for _ in range(self.n_experts)])
        self.aux_regs = nn.ModuleList([copy.deepcopy(aux_reg) for _ in range(self.n_experts)])
        self.router = router
        self.g_steps = [0] * self.n_experts   # never advanced by the reference either (moe.py:37-38)
        self.d_steps = [0] * self.n_experts
        self._dp = None
        self._arenas: Dict[str, Arena] = {}
        self._bind()
        ids = {id(p) for g in self.generators for p in g.parameters()}
        assert len(ids) == sum(1 for g in self.generators for _ in g.parameters()), "experts must not share parameters"

    # ------------------------------------------------------------------------------------------------ storage
    def _bind(self):
        """(Re)build the shared arenas: expert e of each network kind becomes slot e of one flat [E, n] tensor."""
        dev = next(self.router.parameters()).device
        old, self._arenas = self._arenas, {}
        for key, mods in (("g", self.generators), ("d", self.discriminators), ("a", self.aux_regs), ("r", [self.router])):
            arena = Arena(mods[0]._spec, len(mods), dev)
            for e, m in enumerate(mods):
                arena.adopt(m, e)
            if key in old:      # a re-bind after .to()/.cuda(): the optimizer state moves with the parameters, as torch's does
                arena.M.copy_(old[key].M)
                arena.V.copy_(old[key].V)
                arena.steps.copy_(old[key].steps)
            self._arenas[key] = arena

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        # .to()/.cuda()/.float() re-create the parameter storages: re-adopt them so experts stay slots of one arena
        if self._arenas and not all(a.owns(a.modules[0]) for a in self._arenas.values()):
            self._bind()
        return self

    def mark_weights_changed(self):
        """Call after writing to parameters outside load_state_dict / the fused optimizers (bf16 copies are re-packed)."""
        for a in self._arenas.values():
            a.version += 1

    def arena(self, key: str) -> Arena:
        return self._arenas[key]

    def _ensure_bound(self):
        if any(not arena.owns(arena.modules[0]) for arena in self._arenas.values()):
            self._bind()

    def _engines(self):
        self._ensure_bound()
        a = self._arenas
        return (engine_for(a["g"], self.arch, "generator"), engine_for(a["d"], self.arch, "discriminator"),
                engine_for(a["a"], self.arch, "aux_reg"))

    # ------------------------------------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None):
        """Shard the batch over the ranks of ``process_group`` (default: the world).  Each rank calls train_step with
        ITS rows; gradients and loss normalisers are reduced so the update equals the single-device global-batch one."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dp = (dist, process_group, dist.get_world_size(process_group))
        from .._reduce import BucketedGradReducer
        self._reducer = BucketedGradReducer(dist, process_group)      # gradient buckets ride a communicator of their own
        # identical replicas: broadcast rank 0's parameters and buffers
        for arena in self._arenas.values():
            dist.broadcast(arena.P, 0, group=process_group)
            dist.broadcast(arena.Bf, 0, group=process_group)
            dist.broadcast(arena.Bi, 0, group=process_group)
            arena.version += 1
        return self

    @property
    def world_size(self):
        return self._dp[2] if self._dp else 1

    def set_collectives_enabled(self, on: bool):
        """Measurement aid only (bench.py's ``comm.exposed_ms``): with ``on=False`` every data-parallel collective of the
        step becomes a no-op, so the replicas diverge — never use it for training."""
        self._comm_disabled = not on
        if getattr(self, "_reducer", None) is not None:
            self._reducer.disabled = not on

    def _allreduce(self, t):
        if self._dp and not getattr(self, "_comm_disabled", False):
            dist, pg, _ = self._dp
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
        return t

    # ------------------------------------------------------------------------------------------------ helpers
    def _route(self, cond, gumbel, tau, min_rows):
        """Router forward + stable partition.  Returns the router activations and the device-side tables."""
        E, B, dev = self.n_experts, cond.shape[0], cond.device
        r = self.router.raw_forward(cond, gumbel, tau)
        nblk = (B + 255) // 256
        r["counts"] = torch.empty(E, dtype=torch.int32, device=dev)
        r["offsets"] = torch.empty(E + 1, dtype=torch.int32, device=dev)
        r["perm"] = torch.empty(B, dtype=torch.int32, device=dev)
        r["grp_half"] = torch.empty(E, 4, dtype=torch.int32, device=dev)
        r["grp_gen"] = torch.empty(E, 4, dtype=torch.int32, device=dev)
        scratch = torch.empty(nblk * E + E, dtype=torch.int32, device=dev)
        L.call("es_router_partition", r["idx"], B, E, min_rows, r["hist"], r["counts"], r["offsets"], r["perm"],
               r["grp_half"], r["grp_gen"], scratch)
        return r

    def _side_streams(self, dev):
        """three side streams per device; with ``overlap_streams = False`` everything runs on the caller's stream"""
        if not getattr(self, "overlap_streams", True):
            cur = torch.cuda.current_stream()
            return cur, cur, cur
        key = (dev.index if dev.index is not None else torch.cuda.current_device())
        if getattr(self, "_streams", None) is None or self._streams[0] != key:
            self._streams = (key, torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        return self._streams[1], self._streams[2], self._streams[3]

    @staticmethod
    def _gather(x, perm, width):
        B = perm.shape[0]
        out = torch.empty(B, width, device=x.device)
        L.call("es_gather_rows", x, perm, B, width, out)
        return out

    @staticmethod
    def _adam(arena: Arena, lr, grp):
        L.call("es_adam_step", arena.P, arena.G, arena.M, arena.V, arena.n, arena.n, arena.E, float(lr), 0.9, 0.999, 1e-8,
               arena.steps, grp)
        arena.version += 1

    def _adam_pipelined(self, arena: Arena, lr, grp, rects, red):
        """The fused Adam of a data-parallel step, run rectangle by rectangle behind the chunked all-reduce of the widest
        gradient bucket (fc2's weight: 88 % of the generator's bytes, produced LAST by backward, so its collective cannot
        hide behind backward kernels): the optimizer pass over rectangle k — HBM-bound, 28 B per parameter — runs while
        rectangle k + 1 is still being summed over NVLink.  ``rects`` = BucketedGradReducer.reduce_chunked's list; the
        columns outside the bucket follow (right of it as soon as the first event has fired — their buckets were queued
        earlier on the same communication stream — left of it after join()).  Elementwise identical to ONE es_adam_step
        over the arena: every call advances its own copy of the pre-step counters; the real ones take the maximum."""
        n, E = arena.n, arena.E
        lo, hi = min(r[2] for r in rects), max(r[3] for r in rects)
        regions = [(0, E, hi, n, rects[0][4])] if hi < n else []
        regions += list(rects)
        regions += [(0, E, 0, lo, "join")] if lo > 0 else [(0, 0, 0, 0, "join")]
        live = [r for r in regions if r[1] > r[0] and r[3] > r[2]]
        tmp = arena.steps.unsqueeze(0).repeat(len(live), 1)        # pre-step counters, one private copy per call
        main, k = (torch.cuda.current_stream(arena.P.device) if arena.P.is_cuda else None), 0
        gp = grp.data_ptr() if grp is not None else None
        for e0, e1, c0, c1, ev in regions:
            if ev == "join":
                red.join()
            elif ev is not None and main is not None:
                main.wait_event(ev)
            if e1 <= e0 or c1 <= c0:
                continue
            k += 1
            steps = tmp[k - 1]
            off = 4 * (e0 * n + c0)
            L.call("es_adam_step", arena.P.data_ptr() + off, arena.G.data_ptr() + off, arena.M.data_ptr() + off,
                   arena.V.data_ptr() + off, c1 - c0, n, e1 - e0, float(lr), 0.9, 0.999, 1e-8, steps.data_ptr() + 4 * e0,
                   None if gp is None else gp + 16 * e0)
        # the regions tile the arena, so every slot was advanced in at least one copy (live experts: t + 1, skipped: t)
        arena.steps.copy_(tmp.amax(0))
        arena.version += 1

    @staticmethod
    def _lr(opt, default):
        """Learning rate of one network kind, read from the optimizer handles.  ONE fused launch updates every expert
        with torch.optim.Adam's defaults (as training_setup.py:12-41 builds them), so the handles of a kind must agree:
        per-expert learning rates / non-default betas or eps are refused loudly instead of being ignored."""
        if opt is None:
            return default
        opts = list(opt) if isinstance(opt, (list, tuple)) else [opt]
        lr = float(opts[0].param_groups[0]["lr"])
        for o in opts:
            for g in o.param_groups:
                if float(g["lr"]) != lr:
                    raise NotImplementedError(f"per-expert learning rates ({g['lr']} vs {lr}) are not supported by the fused Adam")
                if tuple(g.get("betas", (0.9, 0.999))) != (0.9, 0.999) or float(g.get("eps", 1e-8)) != 1e-8 or \
                        g.get("weight_decay", 0) or g.get("amsgrad", False):
                    raise NotImplementedError("the fused Adam implements torch.optim.Adam defaults only (betas=(0.9, 0.999), "
                                              "eps=1e-8, no weight decay / amsgrad), as the reference's setup_optimizers")
        return lr

    def _tau(self, epoch):
        rc = self.cfg.model.router
        return max(float(rc.tau_min), float(rc.tau_start) * (float(rc.tau_decay) ** epoch))  # moe.py:62-74

    # ------------------------------------------------------------------------------------------------ CUDA graph
    def enable_cuda_graph(self, on: bool = True):
        """Replay the whole training step — ~400 kernel launches on four streams — as ONE CUDA graph.  The step is static
        (no host sync, device-side group tables), so a capture is valid for as long as its host-side scalars are: the
        epoch (router temperature, ALB weight), the learning rates, the batch size.  Per key the first call runs eagerly
        (which also creates streams, attributes, tensor maps), the second captures, every later one copies the batch into
        the graph's static inputs and replays: the host issues 1 launch instead of ~400 (20.9 -> 0.02 ms of host time
        per step, 35.1 -> 34.1 ms on the device; tools/experiments/graph_probe.py).  Steps with injected noise (the parity
        harness) and data-parallel steps always run eagerly."""
        self._graphs = {} if on else None
        return self

    def _graph_key(self, epoch, B, opts):
        lrs = tuple(None if o is None else self._lr(o, 0.0) for o in opts)
        return (int(epoch), int(B), bool(self.training), lrs)

    def train_step(self, epoch, cond, real_images, true_positions, std, intensity, aux_reg_optimizers=None,
                   generator_optimizers=None, discriminator_optimizers=None, router_optimizer=None, ema_helper=None,
                   device=None, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Public entry (signature of the reference, moe.py:52-56): eager step, or CUDA-graph replay once
        ``enable_cuda_graph()`` was called (see there)."""
        ZP.active = False          # (a step that raised half-way must not leave the scratch pool switched on)
        graphs = getattr(self, "_graphs", None)
        opts = (aux_reg_optimizers, generator_optimizers, discriminator_optimizers, router_optimizer)
        args = (cond, real_images, true_positions, std, intensity)
        # single-device steps only: capturing the step WITH its NCCL collectives (two communicators, a communication stream)
        # was tried on 2 x B200 and hung in the first replay, so data-parallel steps stay eager
        if graphs is None or noise or self.world_size > 1:
            return self._train_step_impl(epoch, *args, *opts, ema_helper, device, noise)
        key = self._graph_key(epoch, cond.shape[0], opts)
        ent = graphs.get(key)
        if ent is None:                                   # first step with these scalars: eager (also the warm-up)
            for k in [k for k in graphs if k != key]:     # scalars moved on (new epoch / lr): drop the old captures
                del graphs[k]
            graphs[key] = "warm"
            return self._train_step_impl(epoch, *args, *opts, ema_helper, device, None)
        dev = self._arenas["g"].device
        if ent == "warm":                                 # second step: capture (records, does not execute) ...
            static = tuple(torch.empty(tuple(t.shape), dtype=torch.float32, device=dev) for t in args)
            for d_, s_ in zip(static, args):
                d_.copy_(s_, non_blocking=True)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.graph(g, stream=side):
                out = self._train_step_impl(epoch, *static, *opts, ema_helper, device, None)
            ent = graphs[key] = (g, static, out)
        else:
            for d_, s_ in zip(ent[1], args):
                d_.copy_(s_.reshape(d_.shape), non_blocking=True)
        ent[0].replay()                                   # ... and replay: this IS the step
        return ent[2]

    def _train_step_impl(self, epoch, cond, real_images, true_positions, std, intensity, aux_reg_optimizers=None,
                         generator_optimizers=None, discriminator_optimizers=None, router_optimizer=None, ema_helper=None,
                         device=None, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """One optimisation step on a batch (reference moe.py:52-504).  ``noise`` (optional) injects every random draw,
        indexed by ORIGINAL sample: 'gumbel' [B,E], 'z1','z2' [B,noise_dim] and, per network, dropout keep-masks — the
        parity harness uses it; by default the draws come from torch's CUDA generator.  Learning rates are read from the
        optimizers (``param_groups[0]['lr']``); the update itself is the fused multi-tensor Adam over the arenas.
        Returns the reference's metric dict; values are 0-dim DEVICE tensors (no host sync inside the step)."""
        self._ensure_bound()
        if self._arenas["g"].device.type != "cuda":
            raise RuntimeError("expertsim (B200) computes on CUDA only; there is no CPU fallback")
        cfgm = self.cfg.model
        rc = cfgm.router
        E, B = self.n_experts, cond.shape[0]
        H, W = self.image_shape
        HW = H * W
        dev = cond.device
        world = self.world_size
        Bg = B * world
        a_g, a_d, a_a, a_r = (self._arenas[k] for k in "gdar")
        f32 = lambda t, w: t.to(dev, torch.float32).reshape(B, w).contiguous()
        cond = f32(cond, 9)
        noise = noise or {}
        gumbel = noise["gumbel"].to(dev).float().contiguous() if "gumbel" in noise else \
            -torch.empty(B, E, device=dev).exponential_().log()
        tau = self._tau(epoch)

        # ---- K1: gating + stable partition (moe.py:76-77,97-103,123-126); the skip rule B_e<=1 is min_rows=2.
        # The router (4 CTAs, latency-bound) runs on a side stream under the re-packing of the generator's bf16 weight
        # copies, which `_engines()` enqueues on the main stream after an optimizer step.
        main = torch.cuda.current_stream()
        s1, s2, s3 = self._side_streams(dev)
        ZP.begin(dev)                   # ONE memset for every zero-initialised scratch buffer of the step
        ev_in = main.record_event()
        with torch.cuda.stream(s1):
            s1.wait_event(ev_in)
            r = self._route(cond, gumbel, tau, 2 if world == 1 else 1)
            ev_route = s1.record_event()
        gen, disc, aux = self._engines()
        main.wait_event(ev_route)
        perm, gh, gg = r["perm"], r["grp_half"], r["grp_gen"]
        counts_g = r["counts"].to(torch.float32)
        lv_h, lv_g = gh, gg     # LIVENESS tables: gate everything that must advance identically on every replica
        if world > 1:   # skip rule and every mean use the GLOBAL per-expert count
            self._allreduce(counts_g)
            alive = (counts_g >= 2).to(torch.int32)
            gh[:, 1] *= alive
            gh[:, 3] *= alive
            gg[:, 1] *= alive
            gg[:, 3] *= alive
            # A rank may hold NO rows of an expert that is alive globally (unbalanced routing is the normal regime of
            # an MoE).  The row-walking kernels see the local tables (rows = 0: nothing to do); the per-expert STATE
            # updates — fused Adam + step counters, spectral-norm power iterations, BatchNorm running statistics and
            # the affine gradients built from all-reduced sums — are gated on the GLOBAL count instead, otherwise the
            # replicas would diverge for good.
            lv_h, lv_g = gh.clone(), gg.clone()
            lv_h[:, 1], lv_h[:, 3], lv_g[:, 1], lv_g[:, 3] = alive, alive, alive, alive

        # ---- expert-sorted views of the batch (cond[mask], real[mask], ...; moe.py:143,150,165-168)
        cond_s = self._gather(cond, perm, 9)
        real_s = self._gather(f32(real_images, HW), perm, HW)
        pos_s = self._gather(f32(true_positions, 2), perm, 2)
        std_s = self._gather(f32(std, 1), perm, 1)
        int_s = self._gather(f32(intensity, 1), perm, 1)
        if "z1" in noise:
            z1 = self._gather(f32(noise["z1"], 10), perm, 10)
            z2 = self._gather(f32(noise["z2"], 10), perm, 10)
        else:   # i.i.d. rows: drawing directly in sorted order is the same distribution as randn(B_e, 10) per expert
            z1, z2 = torch.randn(B, 10, device=dev), torch.randn(B, 10, device=dev)
        drop = self._dropout_masks(noise, r, B, dev)
        # data-parallel context of the BatchNorm layers (neutron): SyncBN over the global per-expert rows
        dp = None
        if world > 1:
            dp = {"allreduce": self._allreduce, "world": world, "rows_global": counts_g * (counts_g >= 2).to(torch.float32),
                  "live_gen": lv_g, "live_half": lv_h}

        # Independent chains of small kernels run on side streams (they are latency-bound, not throughput-bound):
        #   D(real) forward            || generator forward
        #   D'(fake2) fwd / bwd, aux regressor fwd / bwd   || D'(fake1) fwd / bwd
        # Spectral-norm power iterations keep their reference order a -> b -> c -> d through events.
        ev0 = main.record_event()
        with torch.cuda.stream(s1):
            s1.wait_event(ev0)
            sn_a = disc.spectral(lv_h, self.training)
            ev_a = s1.record_event()
            s_real, _, sv_real = disc.forward(real_s, cond_s, gh, B, sn_a)
            ev_real = s1.record_event()

        # ---- G(z1) and G(z2): one two-pass batch of 2B rows (moe.py:143-145,535-538)
        img1, img2, sg = gen.forward(z1, z2, cond_s, gg, 2 * B, True, training=self.training, drop=drop.get("g"), dp=dp)
        if "img1_sorted" in noise:
            # parity harness only: replace the generated images (expert-sorted rows) by the oracle's fp32 images, so that
            # everything downstream of the generator can be compared at fp32 tolerance.  The networks' gradients are
            # discontinuous in the image (max-pool routing, ReLU/LeakyReLU kinks, GroupNorm over sparse maps): in pure
            # fp32 PyTorch a 1e-2 relative image perturbation already moves dL/d(image) by ~18% (tests/test_step_gpu.py).
            generated = (img1.clone(), img2.clone())      # what the generator really produced, for the image check
            img1.copy_(noise["img1_sorted"].to(dev).reshape(B, HW))
            img2.copy_(noise["img2_sorted"].to(dev).reshape(B, HW))

        # ---- auxiliary regressor (moe.py:557-559): forward AND backward need nothing but G(z1) — the gradient of the
        # regression loss depends only on the regressor's own output — so the whole chain runs on a side stream under the
        # discriminator step and the D' passes of the generator step; it joins before the generator backward.
        gcfg, stren_a = cfgm.generator, float(cfgm.aux_reg.strength)
        ev_g = main.record_event()
        with torch.cuda.stream(s2):
            s2.wait_event(ev_g)
            coords, sv_a = aux.forward(img1, gh, B, self.training, drop.get("a"), dp=dp)
            ev_ax = s2.record_event()
            d_coords = torch.empty(B, 2, device=dev)
            L.call("es_aux_loss_grad", coords, pos_s, gh, E, B, Bg, stren_a, d_coords)
            d_img1_aux = _zeros(B, HW)
            a_a.G.zero_()
            aux.backward(sv_a, d_coords, d_img1_aux, accumulate=False, wgrad_stream=s3 if s3 is not s2 else None)
            ev_ba = s2.record_event()       # d_img1_aux is complete (the conv weight gradients may still run on s3)
            if s3 is not s2:
                s2.wait_stream(s3)
            self._allreduce(a_a.G)          # (data parallel) rides the side stream too
            ev_aw = s2.record_event()       # the auxiliary regressor's gradient arena is complete

        # ---- discriminator step (moe.py:506-527)
        main.wait_event(ev_a)
        sn_b = disc.spectral(lv_h, self.training)
        s_fake, _, sv_fake = disc.forward(img1, cond_s, gh, B, sn_b)
        main.wait_event(ev_real)
        d_real, d_fake = _zeros(B), _zeros(B)
        loss_d = torch.zeros(E, device=dev)
        L.call("es_hinge_d", s_real, s_fake, gh, E, counts_g, Bg, d_real, d_fake, loss_d)
        a_d.G.zero_()
        ev_h = main.record_event()
        with torch.cuda.stream(s1):     # the two passes only meet in atomic accumulations into the gradient arena
            s1.wait_event(ev_h)
            disc.backward(sv_real, sn_a, d_real, None, want_w=True)
            ev_br = s1.record_event()
        disc.backward(sv_fake, sn_b, d_fake, None, want_w=True)
        main.wait_event(ev_br)
        self._allreduce(a_d.G)
        self._adam(a_d, self._lr(discriminator_optimizers, cfgm.discriminator.lr_d), lv_h)
        del sv_real, sv_fake

        # ---- generator step (moe.py:529-571): D carries its UPDATED weights
        sn_c = disc.spectral(lv_h, self.training)
        ev_c = main.record_event()
        with torch.cuda.stream(s1):
            s1.wait_event(ev_c)
            sn_d = disc.spectral(lv_h, self.training)
            _, lat2, sv2 = disc.forward(img2, cond_s, gh, B, sn_d)
            ev_f2 = s1.record_event()
        score1, lat1, sv1 = disc.forward(img1, cond_s, gh, B, sn_c)
        main.wait_event(ev_f2)
        main.wait_event(ev_ax)
        sums = _zeros(E, 8, dtype=torch.float64)
        s_out, div_out = _zeros(B), _zeros(B)
        L.call("es_gen_loss_reduce", img1, HW, lat1, lat2, z1, z2, std_s, int_s, coords, pos_s, score1, gh, E, B,
               s_out, div_out, sums)
        self._allreduce(sums)
        d_score1, d_coords2 = _zeros(B), _zeros(B, 2)   # d_coords2: same values as d_coords
        d_lat1, d_lat2 = _zeros(B, 64), _zeros(B, 64)
        d_img1, d_img2 = _zeros(B, HW), _zeros(B, HW)
        d_score2 = _zeros(B)
        losses = torch.zeros(E, 6, device=dev)
        L.call("es_gen_loss_grads", img1, HW, lat1, lat2, z1, z2, std_s, int_s, coords, pos_s, s_out, div_out, gh, E, B,
               sums, Bg, float(gcfg.di_strength), float(gcfg.in_strength), stren_a, d_score1, d_lat1, d_lat2, d_coords2,
               d_img1, losses)
        ev_l = main.record_event()
        with torch.cuda.stream(s1):
            s1.wait_event(ev_l)
            disc.backward(sv2, sn_d, d_score2, d_lat2, want_w=False, d_img=d_img2, accumulate=False)
            ev_b2 = s1.record_event()
        disc.backward(sv1, sn_c, d_score1, d_lat1, want_w=False, d_img=d_img1, accumulate=True)
        main.wait_event(ev_b2)
        main.wait_event(ev_ba)
        L.call("es_axpy", 1.0, d_img1_aux, B * HW, d_img1)
        a_g.G.zero_()
        if world > 1 and getattr(self, "overlap_grad_allreduce", "deferred"):
            # layer buckets of the generator gradient are all-reduced on the communication stream as backward produces them
            # ("deferred": the conv buckets wait until the persistent tensor-core kernels — one CTA per SM, static tile
            # striding, so an SM lent to NCCL stretches the whole launch — are all enqueued, and then overlap the
            # LayerNorm / fc2 / fc1 part of the backward; True: every bucket starts as soon as it is complete)
            red = self._reducer
            red.begin()
            mode, pending, heavy = getattr(self, "overlap_grad_allreduce", "deferred"), [], [True]
            pipe = self.pipeline_adam and red.compress_min_cols is None

            def on_ready(lo, hi):
                if lo is None:
                    heavy[0] = False
                    if pending:             # adjacent column ranges: one bucket
                        red.reduce(a_g.G, min(b[0] for b in pending), max(b[1] for b in pending))
                    pending.clear()
                elif mode == "deferred" and heavy[0]:
                    pending.append((lo, hi))
                elif pipe and not rects and lo == a_g.off["fc2.0.weight"] and (hi - lo) * E >= self.PIPELINE_MIN_FLOATS:
                    rects.extend(red.reduce_chunked(a_g.G, lo, hi, self.PIPELINE_CHUNKS))
                else:
                    red.reduce(a_g.G, lo, hi)

            rects = []
            gen.backward(sg, d_img1, d_img2, on_grads_ready=on_ready)
            assert red.n_reduced == a_g.G.numel(), "gradient buckets must cover the arena exactly once"
            if not rects:
                red.join()
        else:
            rects = []
            gen.backward(sg, d_img1, d_img2)
            self._allreduce(a_g.G)
        del sg, sv1, sv2, sv_a

        def finish_generator():
            if rects:
                self._adam_pipelined(a_g, self._lr(generator_optimizers, gcfg.lr_g), lv_h, rects, red)
            else:
                self._adam(a_g, self._lr(generator_optimizers, gcfg.lr_g), lv_h)
            if ema_helper is not None and getattr(ema_helper, "enabled", False):
                ema_helper.update(self, live=lv_h[:, 1])      # only the experts that took an optimizer step (loop.py:392-400)

        self.n_pipelined_rects = len(rects)
        if not rects:
            finish_generator()
        # (pipelined: the widest bucket is still on the wire — the auxiliary regressor's optimizer and the router block
        # below do not depend on it and run first; the generator's optimizer follows them)
        main.wait_event(ev_aw)
        self._adam(a_a, self._lr(aux_reg_optimizers, cfgm.aux_reg.lr_a), lv_h)

        # ---- router loss (moe.py:213-449) and metrics
        zero = torch.zeros((), device=dev)
        gen_losses, mean_int = losses[:, 0], losses[:, 5]
        if world > 1:
            self._allreduce(loss_d)
        if E > 1:
            gan = gen_losses.mean() * float(rc.gan_strength)
            gate_sums = torch.empty(E, device=dev)
            L.call("es_router_gate_sums", r["gates"], B, E, gate_sums)
            self._allreduce(gate_sums)
            extra, ed = None, zero
            if _f(rc.ed_strength) != 0:
                if world > 1:
                    raise NotImplementedError("ed_strength couples all sample pairs of the global batch; not sharded")
                m = torch.zeros(B, 1, device=dev)
                L.call("es_scatter_rows", s_out, perm, B, 1, m)       # moe.py:197-198
                extra, edl = torch.zeros(B, E, device=dev), torch.zeros(1, device=dev)
                L.call("es_router_ed_loss", r["idx"], m, B, E, float(rc.ed_strength), extra, edl)
                ed = edl[0]
            ds = _f(rc.diff_strength)
            if ds != 0:   # -sum_{i<j} |mean_i - mean_j| * diff_strength^2 on detached scalars (moe.py:395-405)
                diff = -(mean_int[:, None] - mean_int[None, :]).abs().triu(1).sum() * (ds * ds)
            else:
                diff = zero
            alpha = min(max(epoch / float(rc.alpha), 0.0), 1.0)
            dec_w = float(rc.min_weight) + (1.0 - float(rc.min_weight)) * alpha            # moe.py:412-422
            lr_out = torch.zeros(2, device=dev)
            a_r.G.zero_()
            w = lambda n: a_r.addr(f"fc_layers.{n}.weight")
            gw = [a_r.gaddr(f"fc_layers.{n}.{p}") for n in (0, 2, 4, 6) for p in ("weight", "bias")]
            L.call("es_router_bwd", cond, B, E, Bg, w(2), w(4), w(6), r["gates"], r["h1"], r["h2"], r["h3"], gate_sums, tau,
                   _f(rc.alb_strength), dec_w, _f(rc.util_strength), extra, *gw, lr_out)
            alb, ent = lr_out[0], lr_out[1]
            stop = rc.stop_router_training_epoch
            if stop is None or epoch < stop:
                router_loss = ed + gan + diff + ent + dec_w * alb
                self._allreduce(a_r.G)
                self._adam(a_r, self._lr(router_optimizer, rc.lr_r), None)
            else:
                router_loss = zero
        else:
            gan = router_loss = ed = diff = ent = alb = zero

        if rects:
            finish_generator()
        m = {"gen_loss": gen_losses.mean(), "disc_loss": loss_d.mean(), "div_loss": losses[:, 1].sum() / E,
             "intensity_loss": losses[:, 2].sum() / E, "aux_reg_loss": losses[:, 3].sum() / E, "router_loss": router_loss,
             "expert_distribution_loss": ed, "differentiation_loss": diff, "expert_entropy_loss": ent,
             "adaptive_load_balancing_loss": alb, "gan_loss": gan}
        for i in range(E):
            m.update({f"gen_loss_{i}": losses[i, 0], f"disc_loss_{i}": loss_d[i], f"div_loss_experts_{i}": losses[i, 1],
                      f"intensity_loss_experts_{i}": losses[i, 2], f"aux_reg_loss_experts_{i}": losses[i, 3],
                      f"std_intensities_experts_{i}": losses[i, 4], f"mean_intensities_experts_{i}": losses[i, 5],
                      f"n_choosen_experts_mean_epoch_{i}": counts_g[i]})
        ZP.end()
        self._last = {"idx": r["idx"], "counts": r["counts"], "perm": perm, "img1": img1, "img2": img2, "gates": r["gates"],
                      "logits": r["logits"]}
        if "img1_sorted" in noise:
            self._last["img1_generated"], self._last["img2_generated"] = generated
        return m

    def _dropout_masks(self, noise, r, B, dev):
        """keep-masks of every nn.Dropout on the path when the parity harness injects them ('drop.<net>.<site>' indexed
        by ORIGINAL sample), re-ordered to the kernels' row order; otherwise None = counter-hash dropout inside the
        kernels (neutron) / torch-drawn masks for the two small proton aux-regressor sites.
        Proton: aux head (proton/aux_reg.py:25,29; p=0.3).  Neutron: generator (neutron/generator.py:14-36) and aux
        feature extractor (neutron/aux_reg.py:17-41), p=0.2."""
        if not self.training:
            return {}
        perm = r["perm"]
        g = lambda k: self._gather(noise[k].to(dev).float().reshape(B, -1).contiguous(), perm, noise[k][0].numel())
        if self.arch == "proton":
            if "drop.a.regressor.3" in noise:
                return {"a": (g("drop.a.regressor.3"), g("drop.a.regressor.7"))}
            return {"a": ((torch.rand(B, 128, device=dev) >= 0.3).float(), (torch.rand(B, 64, device=dev) >= 0.3).float())}
        if not any(k.startswith("drop.") for k in noise):
            return {}
        # parity path (host sync is fine here): the generator's two-pass batch holds, per expert, its z1 rows then its z2 rows
        off = r["offsets"].cpu().tolist()
        segs = []
        for e in range(self.n_experts):
            segs += [(0, off[e], off[e + 1]), (1, off[e], off[e + 1])]
        out = {"g": {}, "a": {}}
        for k in noise:
            if k.startswith("drop.g1."):
                site = k[len("drop.g1."):]
                m1, m2 = g(k), g("drop.g2." + site)
                out["g"][site] = torch.cat([(m2 if ps else m1)[lo:hi] for ps, lo, hi in segs]).contiguous()
            elif k.startswith("drop.a."):
                out["a"][k[len("drop.a."):]] = g(k)
        return out

    # ------------------------------------------------------------------------------------------------ inference
    @torch.no_grad()
    def generate(self, cond, noise=None, gumbel=None, out_dtype=torch.float32, to_host=False, chunk=16384,
                 return_routing=False):
        """Batch inference (reference moe.py:650-653 routing + train/utils.py:179-205 generation): router with Gumbel
        noise at tau=1 -> arg-max expert -> every expert's generator in eval mode on its samples -> expm1, returned in
        the ORIGINAL sample order as [N,H,W] (float32, or float64 like the reference's numpy result)."""
        ZP.active = False
        gen, _, _ = self._engines()
        E, N = self.n_experts, cond.shape[0]
        H, W = self.image_shape
        dev = next(self.router.parameters()).device
        outs, idxs = [], []
        for s in range(0, N, chunk):
            c = cond[s:s + chunk].to(dev, torch.float32, non_blocking=True).contiguous()
            B = c.shape[0]
            gmb = gumbel[s:s + chunk].to(dev).float().contiguous() if gumbel is not None else \
                -torch.empty(B, E, device=dev).exponential_().log()
            r = self._route(c, gmb, 1.0, 1)
            z = noise[s:s + chunk].to(dev).float().contiguous() if noise is not None else torch.randn(B, 10, device=dev)
            cs = self._gather(c, r["perm"], 9)
            zs = self._gather(z, r["perm"], 10)
            img, _, _ = gen.forward(zs, None, cs, r["grp_half"], B, False, keep=False, training=False)   # eval: running BN stats
            o64 = torch.empty(B, H, W, dtype=torch.float64, device=dev) if out_dtype == torch.float64 else None
            o32 = torch.empty(B, H, W, device=dev) if o64 is None else None
            L.call("es_expm1_scatter", img, r["perm"], B, H * W, o64, o32)
            o = o64 if o64 is not None else o32
            outs.append(o.cpu() if to_host else o)
            idxs.append(r["idx"])
        out = outs[0] if len(outs) == 1 else torch.cat(outs)
        return (out, torch.cat(idxs)) if return_routing else out

    @torch.no_grad()
    def evaluate(self, epoch, y_test, x_test, true_positions, std, intensity, cfg, device):
        """Wasserstein metric of the reference (moe.py:644-692 with train/utils.py:117-176): route the test conditionals
        once (Gumbel noise at tau=1, as ``self.router(y_test)`` does), then ``min(epoch//5+1, 5)`` times generate every
        sample with its expert's generator and compare the 5 channel sums of generated vs real showers, overall and per
        expert.  Everything up to the per-channel distances stays on the device (grouped generation, expm1 + channel sums
        fused, sort, mean |a-b| for the equally sized samples); the reference copies every image to the host, widens it to
        float64 and calls scipy.  One host sync (expert offsets) and one readback of the distances per call."""
        from ..train.utils import channel_sums_device, ws_device
        gen, _, _ = self._engines()
        E = self.n_experts
        H, W = self.image_shape
        dev = next(self.router.parameters()).device
        cond = y_test.to(dev, torch.float32).contiguous()
        N = cond.shape[0]
        real = torch.as_tensor(np.asarray(x_test.cpu() if isinstance(x_test, torch.Tensor) else x_test)).to(dev, torch.float32)
        real = real.reshape(N, H * W).contiguous()
        gumbel = -torch.empty(N, E, device=dev).exponential_().log()
        r = self._route(cond, gumbel, 1.0, 1)
        off = r["offsets"].cpu().tolist()
        cond_s = self._gather(cond, r["perm"], 9)
        ch_org = channel_sums_device(self._gather(real, r["perm"], H * W), H, W, True)
        org_all = ch_org.sort(dim=0).values
        org_e = [ch_org[off[e]:off[e + 1]].sort(dim=0).values for e in range(E)]
        n_calc = min(epoch // 5 + 1, 5)
        ws = torch.zeros(n_calc, 5, dtype=torch.float64, device=dev)
        ws_exp = torch.zeros(n_calc, E, 5, dtype=torch.float64, device=dev)
        for j in range(n_calc):
            z = torch.randn(N, 10, device=dev)
            img, _, _ = gen.forward(z, None, cond_s, r["grp_half"], N, False, keep=False, training=False)
            ch = channel_sums_device(img, H, W, True)
            ws[j] = ws_device(org_all, ch)
            for e in range(E):
                if off[e + 1] > off[e]:
                    ws_exp[j, e] = ws_device(org_e[e], ch[off[e]:off[e + 1]])
        ws, ws_exp = ws.cpu().numpy(), ws_exp.cpu().numpy()
        runs, runs_exp = ws.mean(axis=1), ws_exp.mean(axis=2)
        log = {"ws_mean": runs.mean(), "ws_std": runs.std(), "epoch": epoch}
        for i in range(E):
            log[f"ws_mean_{i}"], log[f"ws_std_{i}"] = runs_exp.mean(axis=0)[i], runs_exp.std(axis=0)[i]
        return log

    def get_expert_assignment_counts(self, expert_assignments: torch.Tensor) -> torch.Tensor:
        return torch.bincount(expert_assignments, minlength=self.n_experts).float() / expert_assignments.size(0)
