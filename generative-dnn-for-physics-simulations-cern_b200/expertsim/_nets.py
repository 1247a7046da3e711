"""Host-side sequencing of the grouped sm_100a kernels for the three expert networks (generator, discriminator,
auxiliary regressor): explicit forward and hand-written backward, all experts in one launch per layer.

Nothing here computes on the host: every numeric step is a call into libexpertsim_b200.so (see include/expertsim_b200.h);
torch only allocates device buffers.  Shapes follow the reference modules:
  generator      expertsim/models/proton/generator.py:13-52      (reference repo)
  discriminator  expertsim/models/proton/discriminator.py:121-155, expertsim/models/neutron/discriminator.py:11-48
  aux regressor  expertsim/models/proton/aux_reg.py:11-131
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from ._arena import Arena

BF = torch.bfloat16
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def empty(*shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=_dev())


_ITEMSIZE = {torch.float32: 4, torch.float64: 8, torch.int32: 4, torch.int64: 8, torch.bfloat16: 2, torch.uint8: 1, torch.float16: 2}


class ZeroPool:
    """Zero-initialised scratch of one training step from ONE memset.  The step needs ~90 zeroed buffers (accumulators of
    atomics, gradient staging); as separate ``torch.zeros`` calls they were 88 fill launches per step (r01 launch list).
    ``begin()`` clears one persistent byte buffer sized from the previous step's demand, ``take()`` carves 256-byte aligned
    views, anything that does not fit (first step, a larger batch) falls back to ``torch.zeros``.  Two buffers alternate
    and are never freed, so a view stays untouched until the step after the next begins — kernels on the side streams of a
    step have all joined by then.  Tensors that ESCAPE the step (metrics, generated images) are never taken from here."""

    def __init__(self):
        self.bufs, self.flip, self.need = [None, None], 0, 0
        self.cur, self.used, self.extra, self.active = None, 0, 0, False

    def begin(self, dev):
        self.flip ^= 1
        self.cur = None
        if self.need:
            buf = self.bufs[self.flip]
            if buf is None or buf.numel() < self.need or buf.device != dev:
                buf = self.bufs[self.flip] = torch.empty(self.need, dtype=torch.uint8, device=dev)
            buf[:self.need].zero_()
            self.cur = buf
        self.used, self.extra, self.active = 0, 0, True

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = (n * _ITEMSIZE[dtype] + 255) & ~255
        if self.cur is not None and self.used + nbytes <= self.cur.numel():
            v = self.cur[self.used:self.used + nbytes]
            self.used += nbytes
            return v.view(dtype)[:n].view(tuple(shape))
        self.extra += nbytes
        return torch.zeros(tuple(shape), dtype=dtype, device=_dev())

    def end(self):
        self.need = max(self.need, self.used + self.extra)
        self.active, self.cur = False, None


ZP = ZeroPool()


def zeros(*shape, dtype=torch.float32, escape=False):
    """zero-filled device tensor; inside a training step (ZP.active) a view of the step's pre-cleared scratch unless the
    tensor outlives the step (``escape``)"""
    if ZP.active and not escape:
        return ZP.take(shape, dtype)
    return torch.zeros(shape, dtype=dtype, device=_dev())


def conv_geom(Hs, Ws, C, Hu, Wu, KH, KW, pad, N):
    return L.ESConvGeom(Hs, Ws, C, Hu, Wu, Hu + 2 * pad - KH + 1, Wu + 2 * pad - KW + 1, KH, KW, pad, N)


def conv2d(Ci, Hi, Wi, Co, KH, KW, stride, pad):
    return L.ESConv2d(Ci, Hi, Wi, Co, (Hi + 2 * pad - KH) // stride + 1, (Wi + 2 * pad - KW) // stride + 1, KH, KW, stride, pad)


# =====================================================================================================================
# nearest upsample folded into the following conv (see es_igemm_taps_fwd in include/expertsim_b200.h)
# =====================================================================================================================
def _axis_classes(S, U, K, pad, O, fold):
    """Output classes of one axis of conv_K(pad)(upsample_nearest S->U).  -> (per_o, mult, grid, [(p, count, [(off, [k...])])])
    fold: outputs o = p + per_o*a (U/S = per_o/per_s in lowest terms) read source a*per_s + floor((p+k-pad)*S/U); taps k with
    the same offset are pre-summed.  not fold: one class, taps k-pad on the upsampled grid (nearest map applied by the kernel)."""
    from math import gcd
    if not fold or U == S:
        return 1, 1, U, [(0, O, [(k - pad, [k]) for k in range(K)])]
    g = gcd(U, S)
    per_o, per_s = U // g, S // g
    classes = []
    for p_ in range(min(per_o, O)):
        offs = {}
        for k in range(K):
            offs.setdefault((p_ + k - pad) * S // U, []).append(k)      # python floor division
        classes.append((p_, (O - p_ + per_o - 1) // per_o, sorted(offs.items())))
    return per_o, per_s, S, classes


_FUSED_FLAG = (ctypes.c_int32 * 1)()     # host-side answer of the es_igemm_*_fwd_sums calls


class FoldedConv:
    """Tap tables of a stride-1 conv (KHxKW, pad) behind a nearest upsample [Hs,Ws,C] -> [Hu,Wu] -> [Ho,Wo,N], with the
    upsample folded into the conv along the axes named in ``fold``: outputs of one class read the same pattern of distinct
    source pixels, so the taps that land on the same source pixel are pre-summed.
      x2 on both axes, k4/p1: 9+6+6+4 = 25 folded taps per 4 outputs instead of 64 (2.56x fewer MACs); k3/p0: 16 vs 36.
      35->56 (= 5->8) along y only, k4/p1: 8 row classes with 2..3 distinct source rows instead of 4 (1.39x fewer MACs).
    Forward = one table-conv per class (strided output); weight gradient = one table-GEMM per class on a contiguous copy of
    the class's dy pixels, un-folded afterwards; for exact x2 folding the data gradient is ONE table-conv over dy that writes
    the low-resolution gradient directly (``has_dgrad``)."""

    def __init__(self, Hs, Ws, C, Hu, Wu, KH, KW, pad, N, fold=(True, True)):
        self.Hs, self.Ws, self.C, self.KH, self.KW, self.pad, self.N = Hs, Ws, C, KH, KW, pad, N
        self.Ho, self.Wo = Hu + 2 * pad - KH + 1, Wu + 2 * pad - KW + 1
        oy_per, my, gy, ycls = _axis_classes(Hs, Hu, KH, pad, self.Ho, fold[0])
        ox_per, mx, gx, xcls = _axis_classes(Ws, Wu, KW, pad, self.Wo, fold[1])
        self.classes = []       # dict(py, px, Hp, Wp, taps [(dy, dx, mask)], table, g_fwd, g_wg)
        for py, Hp, ytaps in ycls:
            for px, Wp, xtaps in xcls:
                taps = [(dy, dx, sum(1 << (ky * KW + kx) for ky in kys for kx in kxs)) for dy, kys in ytaps for dx, kxs in xtaps]
                assert len(taps) <= 32 and (len(taps) * C) % 128 == 0
                tab = L.ESFoldTable()
                tab.n_taps = len(taps)
                geos = []
                for _ in range(2):
                    g = L.ESTapGeom()
                    g.Hs, g.Ws, g.C, g.Hu, g.Wu, g.Ho, g.Wo, g.my, g.mx = Hs, Ws, C, gy, gx, Hp, Wp, my, mx
                    g.n_taps = len(taps)
                    for i, (dy, dx, mask) in enumerate(taps):
                        g.tap_dy[i], g.tap_dx[i], g.tap_koff[i] = dy, dx, i * C
                        tab.mask[i] = mask
                    g.KK, g.N = len(taps) * C, N
                    g.o_my, g.o_oy, g.o_mx, g.o_ox, g.Ho_full, g.Wo_full = oy_per, py, ox_per, px, self.Ho, self.Wo
                    g.alg_flops_per_row = 2.0 * Hp * Wp * N * KH * KW * C     # un-folded direct-conv FLOPs of this class's pixels
                    geos.append(g)
                self.classes.append(dict(py=py, px=px, Hp=Hp, Wp=Wp, taps=taps, table=tab, g_fwd=geos[0], g_wg=geos[1]))
        self.oy_per, self.ox_per = oy_per, ox_per
        self.executed_ratio = sum(c["Hp"] * c["Wp"] * len(c["taps"]) for c in self.classes) / (self.Ho * self.Wo * KH * KW)
        # combined data gradient: exact x2 on both axes (source row of tap = a + dy  <=>  dy pixel = 2*(s - dy) + py)
        self.has_dgrad = fold[0] and fold[1] and Hu == 2 * Hs and Wu == 2 * Ws and sum(len(c["taps"]) for c in self.classes) <= 32
        if self.has_dgrad:
            g, tab, t = L.ESTapGeom(), L.ESFoldTable(), 0
            g.Hs, g.Ws, g.C, g.Hu, g.Wu, g.Ho, g.Wo, g.my, g.mx = self.Ho, self.Wo, N, self.Ho, self.Wo, Hs, Ws, 2, 2
            for c in self.classes:
                for dy, dx, mask in c["taps"]:
                    g.tap_dy[t], g.tap_dx[t], g.tap_koff[t] = c["py"] - 2 * dy, c["px"] - 2 * dx, t * N
                    tab.mask[t] = mask
                    t += 1
            g.n_taps = tab.n_taps = t
            g.KK, g.N = t * N, C
            g.o_my, g.o_oy, g.o_mx, g.o_ox, g.Ho_full, g.Wo_full = 1, 0, 1, 0, Hs, Ws
            g.alg_flops_per_row = 2.0 * self.Ho * self.Wo * N * KH * KW * C
            self.g_dgrad, self.table_all = g, tab
            self.dg_grid = (Hs, Ws)
        # y-only folding (35 -> 56 rows = 5 -> 8): the data gradient is folded along y too.  Source row s = per_s*j + cls
        # collects the up-rows u with floor(u*Hs/Hu) = cls of every period; up-row u meets dy rows u + pad - ky, so the
        # (u, ky) pairs with the same dy-row offset are pre-summed (5 or 4 distinct dy rows instead of 2 x 4 or 1 x 4 —
        # 1.39x fewer MACs) and the result lands on the [Hs, Wu] grid: y at source resolution, x still upsampled (its
        # fan-in stays in the norm backward).  One table-conv per source-row class.
        self.dgrad_classes = []
        if not self.has_dgrad and fold[0] and not fold[1] and Hu != Hs:
            from math import gcd
            per_o, per_s = Hu // gcd(Hu, Hs), Hs // gcd(Hu, Hs)
            for cls in range(per_s):
                us = [u for u in range(per_o) if u * Hs // Hu == cls]
                offs = {}
                for u in us:
                    for ky in range(KH):
                        offs.setdefault(u + pad - ky, []).append(ky)
                taps = [(o, pad - kx, sum(1 << (ky * KW + kx) for ky in kys)) for o, kys in sorted(offs.items()) for kx in range(KW)]
                assert len(taps) <= 32 and (len(taps) * N) % 128 == 0
                Hp = (Hs - cls + per_s - 1) // per_s
                g, tab = L.ESTapGeom(), L.ESFoldTable()
                g.Hs, g.Ws, g.C, g.Hu, g.Wu, g.Ho, g.Wo, g.my, g.mx = self.Ho, self.Wo, N, self.Ho, self.Wo, Hp, Wu, per_o, 1
                for t, (dy, dx, mask) in enumerate(taps):
                    g.tap_dy[t], g.tap_dx[t], g.tap_koff[t] = dy, dx, t * N
                    tab.mask[t] = mask
                g.n_taps = tab.n_taps = len(taps)
                g.KK, g.N = len(taps) * N, C
                g.o_my, g.o_oy, g.o_mx, g.o_ox, g.Ho_full, g.Wo_full = per_s, cls, 1, 0, Hs, Wu
                n_up = sum(1 for j in range(Hp) for u in us if per_o * j + u < Hu)
                g.alg_flops_per_row = 2.0 * n_up * Wu * N * KH * KW * C
                self.dgrad_classes.append(dict(cls=cls, taps=taps, table=tab, g=g))
            self.has_dgrad = True
            self.dg_grid = (Hs, Wu)

    def alloc(self, E, dev):
        for c in self.classes:
            T = len(c["taps"])
            c["w_f"] = torch.empty(E, self.N, T * self.C, dtype=BF, device=dev)
            c["dw_f"] = torch.empty(E, self.N, T * self.C, device=dev)
        for c in self.dgrad_classes:
            c["w_d"] = torch.empty(E, self.C, len(c["taps"]) * self.N, dtype=BF, device=dev)
        if self.has_dgrad and not self.dgrad_classes:
            self.w_d = torch.empty(E, self.C, self.table_all.n_taps * self.N, dtype=BF, device=dev)

    def fold(self, w_addr, slot_stride, E):
        """fp32 master weights -> every folded bf16 copy: one pass over the weights per output layout"""
        import ctypes
        jobs = [([c["table"] for c in self.classes], [c["w_f"] for c in self.classes], 0)]
        if self.dgrad_classes:
            jobs.append(([c["table"] for c in self.dgrad_classes], [c["w_d"] for c in self.dgrad_classes], 1))
        elif self.has_dgrad:
            jobs.append(([self.table_all], [self.w_d], 1))
        for tables, outs, dg in jobs:
            tarr = (L.ESFoldTable * len(tables))(*tables)
            oarr = (ctypes.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
            L.call("es_fold_weights_multi", w_addr, slot_stride, E, self.N, self.C, self.KH, self.KW, tarr, len(tables), oarr, dg)

    def forward(self, x, bias_addr, bias_stride, y, grp, E, R, pair_sums=None):
        """-> True when ``pair_sums`` (zeroed [R, N/2, 2]) received the GroupNorm sums of the whole map (every class fused)."""
        if pair_sums is None:
            for c in self.classes:
                L.call("es_igemm_taps_fwd", x, c["w_f"], bias_addr, bias_stride, y, c["g_fwd"], grp, E, R)
            return False
        fused, flag = True, _FUSED_FLAG
        for c in self.classes:
            L.call("es_igemm_taps_fwd_sums", x, c["w_f"], bias_addr, bias_stride, y, c["g_fwd"], grp, E, R, pair_sums, flag)
            fused = fused and flag[0] == 1
        return fused

    def wgrad(self, x, dy, dw_addr, slot_stride, grp, E, R):
        """dy [R, Ho*Wo, N] -> parameter gradient in the reference layout (accumulated at dw_addr)."""
        import ctypes
        for c in self.classes:
            c["dw_f"].zero_()
            dyp = empty(R, c["Hp"] * c["Wp"], self.N, dtype=BF)
            L.call("es_pick_pixels", dy, self.Ho, self.Wo, self.N, self.oy_per, c["py"], self.ox_per, c["px"], c["Hp"], c["Wp"], R, dyp)
            L.call("es_igemm_taps_wgrad", x, dyp, c["dw_f"], c["g_wg"], grp, E, R)
        # every class's folded gradient -> the reference layout, one read-modify-write of the gradient arena
        tarr = (L.ESFoldTable * len(self.classes))(*[c["table"] for c in self.classes])
        iarr = (ctypes.c_void_p * len(self.classes))(*[c["dw_f"].data_ptr() for c in self.classes])
        L.call("es_unfold_wgrad_multi", iarr, E, self.N, self.C, self.KH, self.KW, tarr, len(self.classes), dw_addr, slot_stride)

    def dgrad(self, dy, dx, grp, E, R):
        """dy [R, Ho*Wo, N] -> dx [R, dg_grid, C]: on the LOW-resolution grid for exact x2 folding (the upsample's backward
        is folded in), on the [Hs, Wu] grid for y-only folding."""
        for c in self.dgrad_classes:
            L.call("es_igemm_taps_fwd", dy, c["w_d"], None, 0, dx, c["g"], grp, E, R)
        if not self.dgrad_classes:
            L.call("es_igemm_taps_fwd", dy, self.w_d, None, 0, dx, self.g_dgrad, grp, E, R)


def Up2Conv(Hs, Ws, C, KH, KW, pad, N):
    """conv behind an exact x2 nearest upsample, folded on both axes"""
    return FoldedConv(Hs, Ws, C, 2 * Hs, 2 * Ws, KH, KW, pad, N, (True, True))


# =====================================================================================================================
# generator (proton): bf16 NHWC activations, tcgen05 implicit GEMMs
# =====================================================================================================================
class GenEngineProton:
    H, W = 56, 30
    F2 = 92160                                   # 512 * 18 * 10
    # (name, fwd geometry (Hs,Ws,C,Hu,Wu,KH,KW,pad,N), norm name, groups)
    CONVS = (("conv_layers.1", (18, 10, 512, 36, 20, 4, 4, 1, 256), "conv_layers.2", 32),
             ("conv_layers.5", (35, 19, 256, 56, 30, 4, 4, 1, 128), "conv_layers.6", 32),
             ("conv_layers.8", (55, 29, 128, 55, 29, 3, 3, 1, 64), "conv_layers.9", 32))

    def __init__(self, arena: Arena):
        self.a = arena
        E, dev = arena.E, arena.device
        # packed feature f' = (y*10+x)*512 + c  <-  reference feature f = c*180 + y*10 + x   (NCHW view(-1,512,18,10))
        fp = torch.arange(self.F2, device=dev)
        self.row_map = ((fp % 512) * 180 + fp // 512).to(torch.int32).contiguous()
        self.w_fc2 = torch.empty(E, self.F2, 256, dtype=BF, device=dev)
        self.b_fc2 = torch.empty(E, self.F2, device=dev)
        self.g_fc2 = torch.empty(E, self.F2, device=dev)
        self.z_fc2 = torch.empty(E, self.F2, device=dev)
        self.w_fwd, self.w_dg, self.dw_p, self.up2 = {}, {}, {}, {}
        self.upx = {}       # conv name -> (Hs, Ws, Wu): its input arrives x-upsampled from the previous norm kernel
        import os
        self.gn_fused = os.environ.get("ES_GN_FUSED", "1") != "0"      # A/B switch of the conv-epilogue GroupNorm statistics
        for name, (Hs, Ws, C, Hu, Wu, KH, KW, pad, N), _, _ in self.CONVS:
            if (Hu, Wu) == (2 * Hs, 2 * Ws):      # exact x2 nearest upsample in front of the conv: fold it away
                self.up2[name] = Up2Conv(Hs, Ws, C, KH, KW, pad, N)
            elif (Hu, Wu) != (Hs, Ws) and os.environ.get("ES_CONV2_UPX", "1") != "0":
                # 35x19 -> 56x30.  y: rows repeat with period 8 (5 source rows) -> folded.  x (19 -> 30, not a rational
                # phase pattern worth folding): the previous GroupNorm kernel stores its activation nearest-upsampled along
                # x (es_gn_lrelu_fwd_upx: +58 % bytes on one tensor), so the conv reads a [35, 30] source DIRECTLY — which
                # is what makes its forward classes and its weight gradient TMA-fed (es_igemm_fwd_plan variant 2).
                self.up2[name] = FoldedConv(Hs, Wu, C, Hu, Wu, KH, KW, pad, N, (True, False))
                self.upx[name] = (Hs, Ws, Wu)
            elif (Hu, Wu) != (Hs, Ws):            # A/B switch: nearest map along x inside the conv's gather
                self.up2[name] = FoldedConv(Hs, Ws, C, Hu, Wu, KH, KW, pad, N, (True, False))
            if name in self.up2:
                self.up2[name].alloc(E, dev)
                if self.up2[name].has_dgrad:
                    continue
                self.w_dg[name] = torch.empty(E, C, KH, KW, N, dtype=BF, device=dev)     # data gradient stays un-folded
                continue
            self.w_fwd[name] = torch.empty(E, N, KH, KW, C, dtype=BF, device=dev)
            self.w_dg[name] = torch.empty(E, C, KH, KW, N, dtype=BF, device=dev)
            self.dw_p[name] = torch.empty(E, N, KH, KW, C, device=dev)

    def repack(self):
        """fp32 master weights (reference layout) -> bf16 kernel layouts; called after every optimizer step."""
        a, E = self.a, self.a.E
        L.call("es_pack_dense_weight", a.addr("fc2.0.weight"), a.n, E, self.F2, 256, self.row_map, self.w_fc2)
        for src, dst in (("fc2.0.bias", self.b_fc2), ("fc2.1.weight", self.g_fc2), ("fc2.1.bias", self.z_fc2)):
            L.call("es_permute_features", a.addr(src), a.n, self.row_map, E, self.F2, dst, self.F2, 0)
        for name, (Hs, Ws, C, Hu, Wu, KH, KW, pad, N), _, _ in self.CONVS:
            if name in self.up2:
                self.up2[name].fold(a.addr(name + ".weight"), a.n, E)
                if not self.up2[name].has_dgrad:
                    L.call("es_pack_conv_weight", a.addr(name + ".weight"), a.n, E, N, C, KH, KW, None, self.w_dg[name])
            else:
                L.call("es_pack_conv_weight", a.addr(name + ".weight"), a.n, E, N, C, KH, KW, self.w_fwd[name], self.w_dg[name])

    def forward(self, z1, z2, cond, grp, R, two_pass, keep=True, training=True, drop=None, dp=None):
        """z1,z2 [B,10], cond [B,9] in expert-sorted order; grp = generator group table; R = total rows.
        Returns (img1 [B,HW], img2 or None, saved).  The proton generator has no train/eval difference (LayerNorm /
        GroupNorm only, no dropout), so ``training`` and ``drop`` are accepted for interface symmetry and unused."""
        a, E = self.a, self.a.E
        B = R // 2 if two_pass else R
        s = {"R": R, "two_pass": two_pass, "grp": grp}
        s["x0"], s["lin1"], s["h1"] = empty(R, 19), empty(R, 256), empty(R, 256, dtype=BF)
        L.call("es_gen_fc1_fwd", z1, z2, cond, a.addr("fc1.0.weight"), a.addr("fc1.0.bias"), a.addr("fc1.1.weight"),
               a.addr("fc1.1.bias"), a.n, a.n, grp, E, R, int(two_pass), s["x0"], s["lin1"], s["h1"])
        s["y2"] = empty(R, self.F2, dtype=BF)
        L.call("es_igemm_fwd", s["h1"], self.w_fc2, self.b_fc2, self.F2, s["y2"], conv_geom(1, 1, 256, 1, 1, 1, 1, 0, self.F2), grp, E, R)
        act, s["st2"] = empty(R, self.F2, dtype=BF), empty(R, 2)
        L.call("es_ln_lrelu_fwd", s["y2"], self.g_fc2, self.z_fc2, self.F2, self.F2, grp, E, R, act, s["st2"])
        s["a2"] = act
        for i, (name, geo, norm, groups) in enumerate(self.CONVS):
            Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
            g = conv_geom(*geo)
            P = g.Ho * g.Wo
            y = empty(R, P, N, dtype=BF)
            # norm fusion: the conv's epilogue accumulates the GroupNorm sums, the norm is then one streaming pass
            ps = zeros(R, N // 2, 2) if self.gn_fused else None
            if name in self.up2:
                fused = self.up2[name].forward(act, a.addr(name + ".bias"), a.n, y, grp, E, R, ps)
            elif ps is not None:
                L.call("es_igemm_fwd_sums", act, self.w_fwd[name], a.addr(name + ".bias"), a.n, y, g, grp, E, R, ps, _FUSED_FLAG)
                fused = _FUSED_FLAG[0] == 1
            else:
                L.call("es_igemm_fwd", act, self.w_fwd[name], a.addr(name + ".bias"), a.n, y, g, grp, E, R)
                fused = False
            st = empty(R, groups, 2)
            nxt_name = self.CONVS[i + 1][0] if i + 1 < len(self.CONVS) else None
            wu_ = self.upx[nxt_name][2] if nxt_name in self.upx else g.Wo      # the next conv may want its input upsampled along x
            if fused:
                nxt = empty(R, g.Ho * wu_, N, dtype=BF)
                L.call("es_gn_lrelu_apply_fwd", y, ps, a.addr(norm + ".weight"), a.addr(norm + ".bias"), a.n, g.Ho, g.Wo, wu_, N,
                       groups, grp, E, R, nxt, st)
            elif nxt_name in self.upx:
                nxt = empty(R, g.Ho * wu_, N, dtype=BF)
                L.call("es_gn_lrelu_fwd_upx", y, a.addr(norm + ".weight"), a.addr(norm + ".bias"), a.n, g.Ho, g.Wo, wu_, N, groups,
                       grp, E, R, nxt, st)
            else:
                nxt = empty(R, P, N, dtype=BF)
                L.call("es_gn_lrelu_fwd", y, a.addr(norm + ".weight"), a.addr(norm + ".bias"), a.n, P, N, groups, grp, E, R, nxt, st)
            s[f"y{i + 3}"], s[f"st{i + 3}"], s[f"a{i + 3}"] = y, st, nxt
            act = nxt
        img1 = zeros(B, self.H * self.W, escape=True)
        img2 = zeros(B, self.H * self.W, escape=True) if two_pass else None
        L.call("es_gen_out_fwd", act, a.addr("conv_layers.11.weight"), a.addr("conv_layers.11.bias"), a.n, a.n, 55, 29, 64, 2, 2, 1,
               grp, E, R, int(two_pass), img1, img2)
        s["img1"], s["img2"] = img1, img2
        return img1, img2, (s if keep else None)

    def backward(self, s, dimg1, dimg2, on_grads_ready=None):
        """Accumulates every generator parameter gradient into the arena's G tensor (reference layouts).
        ``on_grads_ready(lo, hi)`` is called as soon as the last kernel writing the arena columns [lo, hi) has been
        enqueued (layer buckets from the output conv back to fc1; they tile [0, n) exactly) — the data-parallel step
        starts their all-reduce there."""
        a, E, R, grp = self.a, self.a.E, s["R"], s["grp"]
        hi_col = [a.n]

        def ready(first_name):
            if on_grads_ready is not None:
                lo = a.off[first_name] if first_name else 0
                on_grads_ready(lo, hi_col[0])
                hi_col[0] = lo
        for t in self.dw_p.values():
            t.zero_()
        da = empty(R, 55 * 29, 64, dtype=BF)
        L.call("es_gen_out_bwd", s["a5"], a.addr("conv_layers.11.weight"), a.n, a.n, 55, 29, 64, 2, 2, 1, s["img1"], s["img2"],
               dimg1, dimg2, grp, E, R, int(s["two_pass"]), da, a.gaddr("conv_layers.11.weight"), a.gaddr("conv_layers.11.bias"))
        up = (55, 29)   # spatial size of `da` (gradient w.r.t. the input the next-later layer consumed)
        for i in (2, 1, 0):
            name, geo, norm, groups = self.CONVS[i]
            Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
            g = conv_geom(*geo)
            P = g.Ho * g.Wo
            # norm + LeakyReLU backward; `da` lives on the (possibly upsampled) grid `up`, the layer output on (Ho,Wo)
            dy = empty(R, P, N, dtype=BF)
            L.call("es_gn_lrelu_bwd", da, g.Ho, g.Wo, up[0], up[1], s[f"y{i + 3}"], s[f"st{i + 3}"], a.addr(norm + ".weight"),
                   a.addr(norm + ".bias"), a.n, N, groups, grp, E, R, dy, a.gaddr(norm + ".weight"), a.gaddr(norm + ".bias"),
                   a.gaddr(name + ".bias"))
            if name in self.up2:
                u = self.up2[name]
                u.wgrad(s[f"a{i + 2}"], dy, a.gaddr(name + ".weight"), a.n, grp, E, R)
                ready(name + ".weight")
                if u.has_dgrad:
                    # folded upsample: the data gradient comes out on the folded grid (x2: low resolution, no fan-in left for
                    # the norm backward; y-only: source rows x upsampled columns)
                    da = empty(R, u.dg_grid[0] * u.dg_grid[1], C, dtype=BF)
                    u.dgrad(dy, da, grp, E, R)
                    up = u.dg_grid
                    continue
            else:
                # weight gradient (packed fp32), unpacked into the reference layout
                L.call("es_igemm_wgrad", s[f"a{i + 2}"], dy, self.dw_p[name], g, grp, E, R)
                L.call("es_unpack_conv_wgrad", self.dw_p[name], E, N, C, KH, KW, a.gaddr(name + ".weight"), a.n)
                ready(name + ".weight")
            # data gradient on the (upsampled) input grid
            da = empty(R, Hu * Wu, C, dtype=BF)
            L.call("es_igemm_fwd", dy, self.w_dg[name], None, 0, da, conv_geom(g.Ho, g.Wo, N, g.Ho, g.Wo, KH, KW, KH - 1 - pad, C), grp, E, R)
            up = (Hu, Wu)
        if on_grads_ready is not None:
            on_grads_ready(None, None)      # marker: the persistent tensor-core kernels of this backward are all enqueued
        dy2 = empty(R, self.F2, dtype=BF)
        L.call("es_ln_lrelu_bwd", da, 18, 10, up[0], up[1], 512, s["y2"], s["st2"], self.g_fc2, self.z_fc2, self.F2, grp, E, R, dy2)
        L.call("es_ln_affine_bwd", da, 18, 10, up[0], up[1], 512, s["y2"], dy2, s["st2"], self.g_fc2, self.z_fc2, self.F2, grp, E, R,
               self.row_map, a.n, a.gaddr("fc2.1.weight"), a.gaddr("fc2.1.bias"), a.gaddr("fc2.0.bias"))
        xpad = empty(E * ((R + 63) // 64) * 64, 256, dtype=BF)      # scratch of the TMA-fed kernel: zero-padded per-group h1
        L.call("es_dense_wgrad", dy2, s["h1"], a.gaddr("fc2.0.weight"), a.n, self.F2, 256, self.row_map, grp, E, R, xpad)
        ready("fc2.0.weight")           # 88 % of the generator's gradient bytes
        dh1 = zeros(R, 256)
        L.call("es_dense_dgrad", dy2, self.w_fc2, dh1, self.F2, 256, grp, E, R)
        L.call("es_gen_fc1_bwd", dh1, s["x0"], s["lin1"], a.addr("fc1.1.weight"), a.addr("fc1.1.bias"), a.n, a.n, grp, E, R,
               a.gaddr("fc1.0.weight"), a.gaddr("fc1.0.bias"), a.gaddr("fc1.1.weight"), a.gaddr("fc1.1.bias"))
        ready(None)


# =====================================================================================================================
# discriminator (proton / neutron): fp32 NCHW
# =====================================================================================================================
class DiscEngine:
    SN = ("conv_layers.0", "conv_layers.4", "fc1.0", "fc2.0", "fc3")

    def __init__(self, arena: Arena, arch: str):
        self.a, self.arch = arena, arch
        self.H, self.W = (56, 30) if arch == "proton" else (44, 44)
        self.c0 = conv2d(1, self.H, self.W, 32, 3, 3, 1, 0)
        self.H1, self.W1 = self.c0.Ho // 2, self.c0.Wo // 2
        self.c4 = conv2d(32, self.H1, self.W1, 16, 3, 3, 1, 0)
        self.pool2 = (2, 1) if arch == "proton" else (2, 2)
        self.H2, self.W2 = self.c4.Ho // self.pool2[0], self.c4.Wo // self.pool2[1]
        self.flat = 16 * self.H2 * self.W2
        self.dims = {"conv_layers.0": (32, 9), "conv_layers.4": (16, 288), "fc1.0": (128, self.flat + 9), "fc2.0": (64, 128),
                     "fc3": (1, 64)}

    def spectral(self, grp, training: bool):
        """One spectral-norm evaluation of all five layers (a forward pre-hook firing in the reference): returns the
        normalised weights and what backward needs.  In training mode u, v advance in place (one power iteration)."""
        a, E = self.a, self.a.E
        out = {}
        for name in self.SN:
            O, I = self.dims[name]
            wsn, sig = empty(E, O * I), empty(E)
            uu, vu = empty(E, O), empty(E, I)
            L.call("es_spectral_norm_fwd", a.addr(name + ".weight_orig"), a.baddr(name + ".weight_u"), a.baddr(name + ".weight_v"),
                   a.n, a.nb, a.nb, E, O, I, int(training), grp if training else None, wsn, O * I, sig, uu, vu,
                   empty(E, I + O + 2))
            out[name] = (wsn, sig, uu, vu)
        return out

    def forward(self, img, cond, grp, R, sn):
        """img [R, H*W], cond [R, 9] (sorted order) -> score [R,1], latent [R,64], saved tensors."""
        a, E, n = self.a, self.a.E, self.a.n
        s = {"img": img, "R": R, "grp": grp}
        # fused trunk (csrc/disc_fused.cu): stem = conv0 + GN + LReLU + pool, stage 2 = conv4 + GN + LReLU + pool + concat
        p1, s["st1"] = empty(R, 32, self.H1, self.W1), empty(R, 8, 2)
        L.call("es_disc_stem_fwd", img, sn["conv_layers.0"][0], 32 * 9, a.addr("conv_layers.0.bias"), n,
               a.addr("conv_layers.1.weight"), a.addr("conv_layers.1.bias"), n, self.H, self.W, grp, E, R, p1, s["st1"])
        y2, s["st2"] = empty(R, 16, self.c4.Ho * self.c4.Wo), empty(R, 8, 2)
        fcin = empty(R, self.flat + 9)
        L.call("es_disc_stage2_fwd", p1, sn["conv_layers.4"][0], 16 * 288, a.addr("conv_layers.4.bias"), n,
               a.addr("conv_layers.5.weight"), a.addr("conv_layers.5.bias"), n, cond, self.H1, self.W1, self.pool2[1], grp, E, R,
               y2, s["st2"], fcin, self.flat + 9)
        l1 = empty(R, 128)
        L.call("es_linear_fwd", fcin, self.flat + 9, sn["fc1.0"][0], a.addr("fc1.0.bias"), 128 * (self.flat + 9), n, self.flat + 9, 128, grp, E, R, l1)
        f1, s["s1"] = empty(R, 128), empty(R, 2)
        L.call("es_layernorm_fwd", l1, a.addr("fc1.1.weight"), a.addr("fc1.1.bias"), n, 128, ACT_LRELU, grp, E, R, f1, s["s1"])
        l2 = empty(R, 64)
        L.call("es_linear_fwd", f1, 128, sn["fc2.0"][0], a.addr("fc2.0.bias"), 64 * 128, n, 128, 64, grp, E, R, l2)
        lat, s["s2"] = empty(R, 64), empty(R, 2)
        L.call("es_layernorm_fwd", l2, a.addr("fc2.1.weight"), a.addr("fc2.1.bias"), n, 64, ACT_LRELU, grp, E, R, lat, s["s2"])
        score = zeros(R, 1)
        L.call("es_linear_fwd", lat, 64, sn["fc3"][0], a.addr("fc3.bias"), 64, n, 64, 1, grp, E, R, score)
        s.update(p1=p1, y2=y2, fcin=fcin, l1=l1, f1=f1, l2=l2, lat=lat)
        return score, lat, s

    def backward(self, s, sn, d_score, d_latent, want_w: bool, d_img=None, accumulate=False):
        """want_w: accumulate parameter gradients into the arena (discriminator step).  d_img: output buffer for the
        image gradient (generator step).  d_latent may be None."""
        a, E, n, R, grp = self.a, self.a.E, self.a.n, s["R"], s["grp"]
        dsn = {k: zeros(E, self.dims[k][0] * self.dims[k][1]) for k in self.SN} if want_w else None

        def lin_w(x, ldx, dy, name):
            if want_w:
                O, I = self.dims[name]
                L.call("es_linear_bwd_weight", x, ldx, dy, I, O, grp, E, R, dsn[name], a.gaddr(name + ".bias"), O * I, n)

        d_lat = zeros(R, 64)
        L.call("es_linear_bwd_data", d_score, sn["fc3"][0], 64, 64, 1, grp, E, R, d_lat, 64)
        lin_w(s["lat"], 64, d_score, "fc3")
        if d_latent is not None:
            L.call("es_axpy", 1.0, d_latent, R * 64, d_lat)
        dl2 = zeros(R, 64)
        L.call("es_layernorm_bwd", d_lat, s["l2"], s["s2"], a.addr("fc2.1.weight"), a.addr("fc2.1.bias"), n, 64, ACT_LRELU, grp, E, R, dl2,
               a.gaddr("fc2.1.weight") if want_w else None, a.gaddr("fc2.1.bias") if want_w else None)
        df1 = zeros(R, 128)
        L.call("es_linear_bwd_data", dl2, sn["fc2.0"][0], 64 * 128, 128, 64, grp, E, R, df1, 128)
        lin_w(s["f1"], 128, dl2, "fc2.0")
        dl1 = zeros(R, 128)
        L.call("es_layernorm_bwd", df1, s["l1"], s["s1"], a.addr("fc1.1.weight"), a.addr("fc1.1.bias"), n, 128, ACT_LRELU, grp, E, R, dl1,
               a.gaddr("fc1.1.weight") if want_w else None, a.gaddr("fc1.1.bias") if want_w else None)
        I1 = self.flat + 9
        dfc = zeros(R, I1)
        L.call("es_linear_bwd_data", dl1, sn["fc1.0"][0], 128 * I1, I1, 128, grp, E, R, dfc, I1)
        lin_w(s["fcin"], I1, dl1, "fc1.0")
        dp1 = empty(R, 32, self.H1, self.W1)
        gw = (lambda name: a.gaddr(name)) if want_w else (lambda name: None)
        L.call("es_disc_stage2_bwd", dfc, I1, s["y2"], s["st2"], s["p1"], sn["conv_layers.4"][0], 16 * 288,
               a.addr("conv_layers.5.weight"), a.addr("conv_layers.5.bias"), n, self.H1, self.W1, self.pool2[1], grp, E, R, dp1,
               dsn["conv_layers.4"] if want_w else None, 16 * 288, gw("conv_layers.4.bias"), n, gw("conv_layers.5.weight"),
               gw("conv_layers.5.bias"))
        if d_img is not None and not accumulate:
            d_img.zero_()           # the stem's image gradient is accumulated with atomics (8 GroupNorm groups per sample)
        L.call("es_disc_stem_bwd", dp1, s["img"], sn["conv_layers.0"][0], 32 * 9, a.addr("conv_layers.0.bias"), n,
               a.addr("conv_layers.1.weight"), a.addr("conv_layers.1.bias"), n, s["st1"], self.H, self.W, grp, E, R, d_img,
               dsn["conv_layers.0"] if want_w else None, 32 * 9, gw("conv_layers.0.bias"), gw("conv_layers.1.weight"),
               gw("conv_layers.1.bias"))
        if want_w:
            for name in self.SN:
                O, I = self.dims[name]
                wsn, sig, uu, vu = sn[name]
                L.call("es_spectral_norm_bwd", dsn[name], wsn, uu, vu, sig, O * I, E, O, I, a.gaddr(name + ".weight_orig"), n, grp, empty(E))


def _on_side_stream(stream, keep, fn, *tensors):
    """Run ``fn()`` (kernel launches that only PRODUCE parameter gradients) on ``stream`` behind the work enqueued so far on
    the current stream; the tensors it reads are parked in ``keep`` until the caller has joined the stream (the caching
    allocator would otherwise hand their memory to later allocations of the current stream)."""
    if stream is None:
        fn()
        return
    ev = torch.cuda.current_stream().record_event()
    keep.extend(tensors)
    with torch.cuda.stream(stream):
        stream.wait_event(ev)
        fn()


# =====================================================================================================================
# auxiliary regressor (proton): fp32 NCHW, residual feature extractor + MLP head
# =====================================================================================================================
class AuxEngineProton:
    FE = "feature_extractor"

    def __init__(self, arena: Arena):
        self.a = arena
        self._wstream, self._keep = None, []
        self.c1 = conv2d(1, 56, 30, 32, 5, 5, 2, 1)                     # -> [32,27,14]
        # after MaxPool(k2,s1): [32,26,13]
        self.blocks = []
        H, W, Ci = 26, 13, 32
        for blk, Co in (("res1", 32), ("res2", 64)):
            ca = conv2d(Ci, H, W, Co, 5, 5, 2, 2)
            cb = conv2d(Co, ca.Ho, ca.Wo, Co, 5, 5, 1, 2)
            cd = conv2d(Ci, H, W, Co, 1, 1, 2, 0)
            self.blocks.append((blk, ca, cb, cd, Co))
            H, W, Ci = ca.Ho - 1, ca.Wo - 1, Co                         # MaxPool(k2,s1)
        self.Hf, self.Wf = H, W                                         # [64,5,2]

    # -- small wrappers ------------------------------------------------------------------------------------------
    def _conv(self, x, name, g, grp, R):
        a = self.a
        y = empty(R, g.Co, g.Ho, g.Wo)
        L.call("es_conv2d_fwd", x, a.addr(name + ".weight"), a.addr(name + ".bias"), a.n, a.n, g, grp, a.E, R, y)
        return y

    def _gn(self, x, name, C, HW, groups, act, grp, R):
        a = self.a
        y, st = empty(*x.shape), empty(R, groups, 2)
        L.call("es_groupnorm_fwd", x, a.addr(name + ".weight"), a.addr(name + ".bias"), a.n, C, HW, groups, act, grp, a.E, R, y, st)
        return y, st

    def _pool(self, x, C, H, W, R):
        y, idx = empty(R, C, H - 1, W - 1), empty(R, C, H - 1, W - 1, dtype=torch.uint8)
        L.call("es_maxpool_fwd", x, C, H, W, 2, 2, 1, 1, R, y, idx)
        return y, idx

    def _conv_bwd(self, x, dy, name, g, grp, R, dx, accumulate, want_dx=True):
        a = self.a
        # the weight gradient feeds nothing downstream in this backward: it runs beside the data-gradient chain
        _on_side_stream(self._wstream, self._keep, lambda: L.call(
            "es_conv2d_bwd_weight", x, dy, g, grp, a.E, R, a.gaddr(name + ".weight"), a.gaddr(name + ".bias"), a.n, a.n), x, dy)
        if want_dx:
            L.call("es_conv2d_bwd_data", dy, a.addr(name + ".weight"), a.n, g, grp, a.E, R, dx, int(accumulate))

    def _gn_bwd(self, dy, x, st, name, C, HW, groups, act, grp, R):
        a = self.a
        dx = zeros(*x.shape)
        L.call("es_groupnorm_bwd", dy, x, st, a.addr(name + ".weight"), a.addr(name + ".bias"), a.n, C, HW, groups, act, grp, a.E, R,
               dx, a.gaddr(name + ".weight"), a.gaddr(name + ".bias"))
        return dx

    # -- forward / backward -----------------------------------------------------------------------------------------
    def forward(self, img, grp, R, training, masks=None, dp=None):
        """img [R, 1680] -> coords [R,2].  masks = (keep128 [R,128], keep64 [R,64]) in training mode."""
        a, E, fe = self.a, self.a.E, self.FE
        s = {"img": img, "R": R, "grp": grp, "training": training, "masks": masks}
        c = self._conv(img, f"{fe}.conv1.0", self.c1, grp, R)
        n1, st = self._gn(c, f"{fe}.conv1.1", 32, 27 * 14, 8, ACT_RELU, grp, R)
        x, idx = self._pool(n1, 32, 27, 14, R)
        s.update(c1=c, st1=st, i1=idx)
        for blk, ca, cb, cd, Co in self.blocks:
            p = f"{fe}.{blk}"
            HW = ca.Ho * ca.Wo
            ya = self._conv(x, f"{p}.conv1.0", ca, grp, R)
            na, sta = self._gn(ya, f"{p}.conv1.1", Co, HW, 32, ACT_RELU, grp, R)
            yb = self._conv(na, f"{p}.conv2.0", cb, grp, R)
            nb, stb = self._gn(yb, f"{p}.conv2.1", Co, HW, 32, ACT_NONE, grp, R)
            yd = self._conv(x, f"{p}.downsample.0", cd, grp, R)
            nd, std_ = self._gn(yd, f"{p}.downsample.1", Co, HW, 32, ACT_NONE, grp, R)
            r = empty(R, Co, ca.Ho, ca.Wo)
            L.call("es_add_relu_fwd", nb, nd, r.numel(), r)
            xo, idx = self._pool(r, Co, ca.Ho, ca.Wo, R)
            s[blk] = dict(x=x, ya=ya, sta=sta, na=na, yb=yb, stb=stb, yd=yd, std=std_, r=r, idx=idx)
            x = xo
        feat = empty(R, 64)
        L.call("es_gap_fwd", x, 64, self.Hf * self.Wf, R, feat)
        n = a.n
        m1 = empty(R, 128)
        L.call("es_linear_fwd", feat, 64, a.addr("regressor.0.weight"), a.addr("regressor.0.bias"), n, n, 64, 128, grp, E, R, m1)
        h1, s["s1"] = empty(R, 128), empty(R, 2)
        L.call("es_layernorm_fwd", m1, a.addr("regressor.1.weight"), a.addr("regressor.1.bias"), n, 128, ACT_LRELU, grp, E, R, h1, s["s1"])
        h1d = h1
        if training:
            h1d = empty(R, 128)
            L.call("es_dropout", h1, masks[0], 0.3, R * 128, h1d)
        m2 = empty(R, 64)
        L.call("es_linear_fwd", h1d, 128, a.addr("regressor.4.weight"), a.addr("regressor.4.bias"), n, n, 128, 64, grp, E, R, m2)
        h2, s["s2"] = empty(R, 64), empty(R, 2)
        L.call("es_layernorm_fwd", m2, a.addr("regressor.5.weight"), a.addr("regressor.5.bias"), n, 64, ACT_LRELU, grp, E, R, h2, s["s2"])
        h2d = h2
        if training:
            h2d = empty(R, 64)
            L.call("es_dropout", h2, masks[1], 0.3, R * 64, h2d)
        coords = zeros(R, 2)
        L.call("es_linear_fwd", h2d, 64, a.addr("regressor.8.weight"), a.addr("regressor.8.bias"), n, n, 64, 2, grp, E, R, coords)
        s.update(feat=feat, m1=m1, h1d=h1d, m2=m2, h2d=h2d)
        return coords, s

    def backward(self, s, d_coords, d_img, accumulate=True, wgrad_stream=None):
        """Parameter gradients -> arena G; image gradient accumulated into d_img [R,1680].  With ``wgrad_stream`` the
        convolution weight gradients are enqueued on that stream (the caller joins it before using the gradient arena and
        drops ``s["keep"]`` afterwards); the current stream then carries only the chain that ends in d_img."""
        a, E, n, R, grp, fe = self.a, self.a.E, self.a.n, s["R"], s["grp"], self.FE
        self._wstream, self._keep = wgrad_stream, s.setdefault("keep", [])
        masks, training = s["masks"], s["training"]

        def lin_bwd(x, ldx, dy, name, I, O, need_dx=True):
            L.call("es_linear_bwd_weight", x, ldx, dy, I, O, grp, E, R, a.gaddr(name + ".weight"), a.gaddr(name + ".bias"), n, n)
            if not need_dx:
                return None
            dx = zeros(R, I)
            L.call("es_linear_bwd_data", dy, a.addr(name + ".weight"), n, I, O, grp, E, R, dx, I)
            return dx

        d = lin_bwd(s["h2d"], 64, d_coords, "regressor.8", 64, 2)
        if training:
            L.call("es_dropout", d, masks[1], 0.3, R * 64, d)
        dm2 = zeros(R, 64)
        L.call("es_layernorm_bwd", d, s["m2"], s["s2"], a.addr("regressor.5.weight"), a.addr("regressor.5.bias"), n, 64, ACT_LRELU, grp, E, R,
               dm2, a.gaddr("regressor.5.weight"), a.gaddr("regressor.5.bias"))
        d = lin_bwd(s["h1d"], 128, dm2, "regressor.4", 128, 64)
        if training:
            L.call("es_dropout", d, masks[0], 0.3, R * 128, d)
        dm1 = zeros(R, 128)
        L.call("es_layernorm_bwd", d, s["m1"], s["s1"], a.addr("regressor.1.weight"), a.addr("regressor.1.bias"), n, 128, ACT_LRELU, grp, E, R,
               dm1, a.gaddr("regressor.1.weight"), a.gaddr("regressor.1.bias"))
        dfeat = lin_bwd(s["feat"], 64, dm1, "regressor.0", 64, 128)
        dx = empty(R, 64, self.Hf, self.Wf)
        L.call("es_gap_bwd", dfeat, 64, self.Hf * self.Wf, R, dx)
        for blk, ca, cb, cd, Co in reversed(self.blocks):
            p, b = f"{fe}.{blk}", s[blk]
            HW = ca.Ho * ca.Wo
            dr = empty(R, Co, ca.Ho, ca.Wo)
            L.call("es_maxpool_bwd", dx, b["idx"], Co, ca.Ho, ca.Wo, 2, 2, 1, 1, R, dr)
            dsum = empty(*dr.shape)
            L.call("es_relu_bwd", dr, b["r"], dr.numel(), dsum)
            dyb = self._gn_bwd(dsum, b["yb"], b["stb"], f"{p}.conv2.1", Co, HW, 32, ACT_NONE, grp, R)
            dna = zeros(R, Co, ca.Ho, ca.Wo)
            self._conv_bwd(b["na"], dyb, f"{p}.conv2.0", cb, grp, R, dna, False)
            dya = self._gn_bwd(dna, b["ya"], b["sta"], f"{p}.conv1.1", Co, HW, 32, ACT_RELU, grp, R)
            dxin = zeros(*b["x"].shape)
            self._conv_bwd(b["x"], dya, f"{p}.conv1.0", ca, grp, R, dxin, False)
            dyd = self._gn_bwd(dsum, b["yd"], b["std"], f"{p}.downsample.1", Co, HW, 32, ACT_NONE, grp, R)
            self._conv_bwd(b["x"], dyd, f"{p}.downsample.0", cd, grp, R, dxin, True)
            dx = dxin
        dn1 = empty(R, 32, 27, 14)
        L.call("es_maxpool_bwd", dx, s["i1"], 32, 27, 14, 2, 2, 1, 1, R, dn1)
        dc1 = self._gn_bwd(dn1, s["c1"], s["st1"], f"{fe}.conv1.1", 32, 27 * 14, 8, ACT_RELU, grp, R)
        self._conv_bwd(s["img"], dc1, f"{fe}.conv1.0", self.c1, grp, R, d_img, accumulate)
        self._wstream, self._keep = None, []        # the parked tensors now live (only) in s["keep"]


# =====================================================================================================================
# neutron: BatchNorm (+ Dropout) networks.  Batch statistics are per (expert, pass) "stat group" and only fp64 partial
# sums cross kernels, so the optional ``allreduce`` callback (data parallelism) turns them into SyncBN.
# =====================================================================================================================
def _noop(t):
    return t


def _seed():
    """fresh 62-bit seed from torch's CPU generator (no device sync; honours torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class GenEngineNeutron:
    H, W = 44, 44
    SP = (13, 13, 128)                           # fc2 output viewed as [128, 13, 13]
    F2 = 21632
    P_DROP = 0.2
    # (conv name, fwd geometry (Hs,Ws,C,Hu,Wu,KH,KW,pad,N), BatchNorm name, dropout site)
    CONVS = (("conv_layers.0", (13, 13, 128, 26, 26, 3, 3, 0, 256), "conv_layers.1", "conv_layers.2"),
             ("conv_layers.5", (24, 24, 256, 48, 48, 3, 3, 0, 128), "conv_layers.6", "conv_layers.7"),
             ("conv_layers.9", (46, 46, 128, 46, 46, 2, 2, 0, 64), "conv_layers.10", "conv_layers.11"))

    def __init__(self, arena: Arena):
        self.a = arena
        E, dev = arena.E, arena.device
        fp = torch.arange(self.F2, device=dev)
        # packed feature f' = (y*13+x)*128 + c  <-  reference feature f = c*169 + y*13 + x   (view(-1,128,13,13))
        self.row_map = ((fp % 128) * 169 + fp // 128).to(torch.int32).contiguous()
        self.w_fc2 = torch.empty(E, self.F2, 256, dtype=BF, device=dev)
        self.b_fc2 = torch.empty(E, self.F2, device=dev)
        self.w_fwd, self.w_dg, self.dw_p, self.up2 = {}, {}, {}, {}
        for name, (Hs, Ws, C, Hu, Wu, KH, KW, pad, N), _, _ in self.CONVS:
            if (Hu, Wu) == (2 * Hs, 2 * Ws):      # exact x2 nearest upsample in front of the conv: fold it away
                self.up2[name] = Up2Conv(Hs, Ws, C, KH, KW, pad, N)
                self.up2[name].alloc(E, dev)
                continue
            self.w_fwd[name] = torch.empty(E, N, KH, KW, C, dtype=BF, device=dev)
            self.w_dg[name] = torch.empty(E, C, KH, KW, N, dtype=BF, device=dev)
            self.dw_p[name] = torch.empty(E, N, KH, KW, C, device=dev)

    def repack(self):
        a, E = self.a, self.a.E
        L.call("es_pack_dense_weight", a.addr("fc2.0.weight"), a.n, E, self.F2, 256, self.row_map, self.w_fc2)
        L.call("es_permute_features", a.addr("fc2.0.bias"), a.n, self.row_map, E, self.F2, self.b_fc2, self.F2, 0)
        for name, (Hs, Ws, C, Hu, Wu, KH, KW, pad, N), _, _ in self.CONVS:
            if name in self.up2:
                self.up2[name].fold(a.addr(name + ".weight"), a.n, E)
            else:
                L.call("es_pack_conv_weight", a.addr(name + ".weight"), a.n, E, N, C, KH, KW, self.w_fwd[name], self.w_dg[name])

    # ---- one BatchNorm (+dropout +LeakyReLU) layer on bf16 NHWC rows
    def _bn_fwd(self, x, bn, geo, feat, chmap, site, ctx):
        a, E = self.a, self.a.E
        Hs, Ws, C = geo
        CS = Hs * Ws * C if feat else C
        passes = 2 if ctx["two_pass"] else 1
        stats = empty(2 * E, CS, 2)
        sums = None
        if ctx["training"]:
            sums = zeros(2 * E, CS, 2, dtype=torch.float64)
            L.call("es_bn_stats_nhwc", x, Hs, Ws, C, int(feat), ctx["grp"], E, ctx["R"], int(ctx["two_pass"]), sums)
            ctx["allreduce"](sums)
        n_sg = (ctx["n_rows"] * (1 if feat else Hs * Ws)).contiguous()
        L.call("es_bn_finalize", sums, n_sg, CS, passes, int(ctx["training"]), 0.1, chmap, a.baddr(bn + ".running_mean"),
               a.baddr(bn + ".running_var"), a.nb, a.iaddr(bn + ".num_batches_tracked"), a.ni,
               ctx["grp_slots"] if ctx["training"] else None, E, stats)
        mask, seed = ctx["drop"](site)
        p = self.P_DROP if ctx["training"] else 0.0
        y = empty(*x.shape, dtype=BF)
        L.call("es_bn_apply_fwd_nhwc", x, Hs, Ws, C, int(feat), stats, a.addr(bn + ".weight"), a.addr(bn + ".bias"), a.n, chmap,
               mask, seed, p, ctx["grp"], E, ctx["R"], int(ctx["two_pass"]), y)
        return y, dict(stats=stats, n_sg=n_sg, mask=mask, seed=seed, p=p)

    def _bn_bwd(self, da, up, x, sv, bn, geo, feat, chmap, ctx):
        """da: gradient on the consumer's (upsampled) grid ``up``; returns dx (bf16) w.r.t. the pre-BatchNorm tensor."""
        a, E = self.a, self.a.E
        Hs, Ws, C = geo
        CS = Hs * Ws * C if feat else C
        passes = 2 if ctx["two_pass"] else 1
        sums2 = zeros(2 * E, CS, 2, dtype=torch.float64)
        args = (da, Hs, Ws, up[0], up[1], C, int(feat), x, sv["stats"])
        tail = (a.addr(bn + ".weight"), a.addr(bn + ".bias"), a.n, chmap, sv["mask"], sv["seed"], sv["p"], ctx["grp"], E, ctx["R"],
                int(ctx["two_pass"]))
        L.call("es_bn_bwd_reduce_nhwc", *args, *tail, sums2)
        ctx["allreduce"](sums2)
        L.call("es_bn_affine_grads", sums2, CS, passes, 1.0 / ctx["world"], chmap, ctx["grp_slots"], E, a.gaddr(bn + ".weight"),
               a.gaddr(bn + ".bias"), a.n)
        dx = empty(*x.shape, dtype=BF)
        L.call("es_bn_bwd_apply_nhwc", *args, sums2, sv["n_sg"], *tail, dx)
        return dx

    def _ctx(self, grp, R, two_pass, training, drop, dp):
        """drop: None (hash-based dropout, fresh seed per site) or dict site -> keep mask in the batch's row order."""
        E = self.a.E
        g = grp.view(E, 4)
        rows_pass = (g[:, 3] if two_pass else g[:, 1]).to(torch.float32)
        n_rows = dp["rows_global"] if dp and dp.get("rows_global") is not None else rows_pass
        sites = {}

        def dropf(site):
            if site not in sites:
                sites[site] = (drop[site], 0) if drop is not None else (None, _seed())
            return sites[site]

        # grp_slots gates the per-expert STATE updates (running statistics, affine gradients from all-reduced sums): under
        # data parallelism that is the liveness table built from the GLOBAL counts, not this rank's rows
        return dict(grp=grp, grp_slots=(dp or {}).get("live_gen", grp), R=R, two_pass=two_pass, training=training, drop=dropf,
                    n_rows=n_rows.repeat_interleave(2).contiguous(), allreduce=(dp or {}).get("allreduce", _noop),
                    world=(dp or {}).get("world", 1))

    def forward(self, z1, z2, cond, grp, R, two_pass, keep=True, training=True, drop=None, dp=None):
        a, E = self.a, self.a.E
        B = R // 2 if two_pass else R
        ctx = self._ctx(grp, R, two_pass, training, drop, dp)
        s = {"R": R, "two_pass": two_pass, "grp": grp, "ctx": ctx}
        s["x0"], lin1 = empty(R, 19), empty(R, 256, dtype=BF)
        L.call("es_gen_fc1_fwd", z1, z2, cond, a.addr("fc1.0.weight"), a.addr("fc1.0.bias"), None, None, a.n, a.n, grp, E, R,
               int(two_pass), s["x0"], None, lin1)
        s["lin1"] = lin1
        s["h1"], s["bn1"] = self._bn_fwd(lin1, "fc1.1", (1, 1, 256), False, None, "fc1.2", ctx)
        y2 = empty(R, self.F2, dtype=BF)
        L.call("es_igemm_fwd", s["h1"], self.w_fc2, self.b_fc2, self.F2, y2, conv_geom(1, 1, 256, 1, 1, 1, 1, 0, self.F2), grp, E, R)
        s["y2"] = y2
        act, s["bn2"] = self._bn_fwd(y2, "fc2.1", self.SP, True, self.row_map, "fc2.2", ctx)
        s["a2"] = act
        for i, (name, geo, bn, site) in enumerate(self.CONVS):
            g = conv_geom(*geo)
            y = empty(R, g.Ho * g.Wo, g.N, dtype=BF)
            if name in self.up2:
                self.up2[name].forward(act, a.addr(name + ".bias"), a.n, y, grp, E, R)
            else:
                L.call("es_igemm_fwd", act, self.w_fwd[name], a.addr(name + ".bias"), a.n, y, g, grp, E, R)
            act, s[f"bn{i + 3}"] = self._bn_fwd(y, bn, (g.Ho, g.Wo, g.N), False, None, site, ctx)
            s[f"y{i + 3}"], s[f"a{i + 3}"] = y, act
        img1 = zeros(B, self.H * self.W, escape=True)
        img2 = zeros(B, self.H * self.W, escape=True) if two_pass else None
        L.call("es_gen_out_fwd", act, a.addr("conv_layers.13.weight"), a.addr("conv_layers.13.bias"), a.n, a.n, 45, 45, 64, 2, 2, 0,
               grp, E, R, int(two_pass), img1, img2)
        s["img1"], s["img2"] = img1, img2
        return img1, img2, (s if keep else None)

    def backward(self, s, dimg1, dimg2, on_grads_ready=None):
        """``on_grads_ready(lo, hi)``: see GenEngineProton.backward."""
        a, E, R, grp, ctx = self.a, self.a.E, s["R"], s["grp"], s["ctx"]
        hi_col = [a.n]

        def ready(first_name):
            if on_grads_ready is not None:
                lo = a.off[first_name] if first_name else 0
                on_grads_ready(lo, hi_col[0])
                hi_col[0] = lo
        for t in self.dw_p.values():
            t.zero_()
        da = empty(R, 45 * 45, 64, dtype=BF)
        L.call("es_gen_out_bwd", s["a5"], a.addr("conv_layers.13.weight"), a.n, a.n, 45, 45, 64, 2, 2, 0, s["img1"], s["img2"],
               dimg1, dimg2, grp, E, R, int(s["two_pass"]), da, a.gaddr("conv_layers.13.weight"), a.gaddr("conv_layers.13.bias"))
        up = (45, 45)
        for i in (2, 1, 0):
            name, geo, bn, _ = self.CONVS[i]
            Hs, Ws, C, Hu, Wu, KH, KW, pad, N = geo
            g = conv_geom(*geo)
            dy = self._bn_bwd(da, up, s[f"y{i + 3}"], s[f"bn{i + 3}"], bn, (g.Ho, g.Wo, N), False, None, ctx)
            # the conv bias sits in front of a BatchNorm: its gradient is identically zero and is left at zero
            if name in self.up2:
                u = self.up2[name]
                u.wgrad(s[f"a{i + 2}"], dy, a.gaddr(name + ".weight"), a.n, grp, E, R)
                ready(name + ".weight")
                da = empty(R, Hs * Ws, C, dtype=BF)
                u.dgrad(dy, da, grp, E, R)
                up = (Hs, Ws)
                continue
            L.call("es_igemm_wgrad", s[f"a{i + 2}"], dy, self.dw_p[name], g, grp, E, R)
            L.call("es_unpack_conv_wgrad", self.dw_p[name], E, N, C, KH, KW, a.gaddr(name + ".weight"), a.n)
            ready(name + ".weight")
            da = empty(R, Hu * Wu, C, dtype=BF)
            L.call("es_igemm_fwd", dy, self.w_dg[name], None, 0, da, conv_geom(g.Ho, g.Wo, N, g.Ho, g.Wo, KH, KW, KH - 1 - pad, C), grp, E, R)
            up = (Hu, Wu)
        if on_grads_ready is not None:
            on_grads_ready(None, None)      # marker: the persistent tensor-core kernels of this backward are all enqueued
        dy2 = self._bn_bwd(da, up, s["y2"], s["bn2"], "fc2.1", self.SP, True, self.row_map, ctx)
        xpad = empty(E * ((R + 63) // 64) * 64, 256, dtype=BF)      # scratch of the TMA-fed kernel: zero-padded per-group h1
        L.call("es_dense_wgrad", dy2, s["h1"], a.gaddr("fc2.0.weight"), a.n, self.F2, 256, self.row_map, grp, E, R, xpad)
        ready("fc2.0.weight")           # 88 % of the generator's gradient bytes
        dh1 = zeros(R, 256)
        L.call("es_dense_dgrad", dy2, self.w_fc2, dh1, self.F2, 256, grp, E, R)
        dlin = self._bn_bwd(dh1.to(BF), (1, 1), s["lin1"], s["bn1"], "fc1.1", (1, 1, 256), False, None, ctx)
        L.call("es_gen_fc1_bwd", dlin.float(), s["x0"], None, None, None, a.n, a.n, grp, E, R, a.gaddr("fc1.0.weight"),
               a.gaddr("fc1.0.bias"), None, None)
        ready(None)


class AuxEngineNeutron:
    """AuxRegNeutron (reference neutron/aux_reg.py:8-80): 4 x [Conv k3 -> BatchNorm2d -> LeakyReLU -> Dropout(.2) (-> MaxPool)],
    1x1 reduce (no bias) -> BatchNorm2d -> LeakyReLU, global average pool, Linear(64, 2).  fp32 NCHW."""
    FE = "feature_extractor"
    P_DROP = 0.2

    def __init__(self, arena: Arena):
        self.a = arena
        fe = self.FE
        self.layers = []            # (conv name, conv geometry, bn name, dropout site or None, pool (kh,kw) or None)
        H, W, Ci = 44, 44, 1
        for i, (Co, pool) in enumerate(((32, (2, 2)), (64, (2, 1)), (128, (2, 1)), (256, None)), start=1):
            c = conv2d(Ci, H, W, Co, 3, 3, 1, 0)
            self.layers.append((f"{fe}.conv{i}", c, f"{fe}.conv{i}_bd.0", f"conv{i}_bd.2", pool))
            H, W, Ci = c.Ho, c.Wo, Co
            if pool:
                H, W = H // pool[0], W // pool[1]
        c = conv2d(Ci, H, W, 64, 1, 1, 1, 0)
        self.layers.append((f"{fe}.reduce.0", c, f"{fe}.reduce.1", None, None))
        self.Hf, self.Wf = c.Ho, c.Wo

    def forward(self, img, grp, R, training, masks=None, dp=None):
        """masks: None (hash dropout) or dict site -> keep mask [R, C, H, W]."""
        a, E, n = self.a, self.a.E, self.a.n
        dp = dp or {}
        allreduce, world = dp.get("allreduce", _noop), dp.get("world", 1)
        g = grp.view(E, 4)
        n_rows = dp["rows_global"] if dp.get("rows_global") is not None else g[:, 1].to(torch.float32)
        n_rows = n_rows.repeat_interleave(2).contiguous()
        live = dp.get("live_half", grp)      # gates running statistics / affine gradients (global liveness under DP)
        s = {"R": R, "grp": grp, "layers": [], "allreduce": allreduce, "world": world, "live": live}
        x = img
        for name, c, bn, site, pool in self.layers:
            has_bias = (name + ".bias") in a.off
            y = empty(R, c.Co, c.Ho, c.Wo)
            L.call("es_conv2d_fwd", x, a.addr(name + ".weight"), a.addr(name + ".bias") if has_bias else None, n, n, c, grp, E, R, y)
            P = c.Ho * c.Wo
            stats, sums = empty(2 * E, c.Co, 2), None
            if training:
                sums = zeros(2 * E, c.Co, 2, dtype=torch.float64)
                L.call("es_bn2d_stats", y, c.Co, P, grp, E, R, sums)
                allreduce(sums)
            n_sg = (n_rows * P).contiguous()
            L.call("es_bn_finalize", sums, n_sg, c.Co, 1, int(training), 0.1, None, a.baddr(bn + ".running_mean"),
                   a.baddr(bn + ".running_var"), a.nb, a.iaddr(bn + ".num_batches_tracked"), a.ni, live if training else None, E, stats)
            p = self.P_DROP if (training and site) else 0.0
            mask, seed = (masks[site], 0) if (masks is not None and site) else (None, _seed() if p else 0)
            act = empty(R, c.Co, c.Ho, c.Wo)
            L.call("es_bn2d_apply_fwd", y, c.Co, P, stats, a.addr(bn + ".weight"), a.addr(bn + ".bias"), n, mask, seed, p, grp, E, R, act)
            rec = dict(x=x, y=y, stats=stats, n_sg=n_sg, mask=mask, seed=seed, p=p, idx=None)
            if pool:
                kh, kw = pool
                po = empty(R, c.Co, c.Ho // kh, c.Wo // kw)
                rec["idx"] = empty(R, c.Co, c.Ho // kh, c.Wo // kw, dtype=torch.uint8)
                L.call("es_maxpool_fwd", act, c.Co, c.Ho, c.Wo, kh, kw, kh, kw, R, po, rec["idx"])
                act = po
            s["layers"].append(rec)
            x = act
        feat = empty(R, 64)
        L.call("es_gap_fwd", x, 64, self.Hf * self.Wf, R, feat)
        coords = zeros(R, 2)
        L.call("es_linear_fwd", feat, 64, a.addr("dense.weight"), a.addr("dense.bias"), n, n, 64, 2, grp, E, R, coords)
        s["feat"] = feat
        return coords, s

    def backward(self, s, d_coords, d_img, accumulate=True, wgrad_stream=None):
        """``wgrad_stream``: see AuxEngineProton.backward."""
        a, E, n, R, grp = self.a, self.a.E, self.a.n, s["R"], s["grp"]
        keep = s.setdefault("keep", [])
        L.call("es_linear_bwd_weight", s["feat"], 64, d_coords, 64, 2, grp, E, R, a.gaddr("dense.weight"), a.gaddr("dense.bias"), n, n)
        dfeat = zeros(R, 64)
        L.call("es_linear_bwd_data", d_coords, a.addr("dense.weight"), n, 64, 2, grp, E, R, dfeat, 64)
        d = empty(R, 64, self.Hf, self.Wf)
        L.call("es_gap_bwd", dfeat, 64, self.Hf * self.Wf, R, d)
        for li in range(len(self.layers) - 1, -1, -1):
            name, c, bn, site, pool = self.layers[li]
            rec = s["layers"][li]
            P = c.Ho * c.Wo
            if pool:
                kh, kw = pool
                dact = empty(R, c.Co, c.Ho, c.Wo)
                L.call("es_maxpool_bwd", d, rec["idx"], c.Co, c.Ho, c.Wo, kh, kw, kh, kw, R, dact)
                d = dact
            sums2 = zeros(2 * E, c.Co, 2, dtype=torch.float64)
            common = (a.addr(bn + ".weight"), a.addr(bn + ".bias"), n, rec["mask"], rec["seed"], rec["p"], grp, E, R)
            L.call("es_bn2d_bwd_reduce", d, rec["y"], c.Co, P, rec["stats"], *common, sums2)
            s["allreduce"](sums2)
            L.call("es_bn_affine_grads", sums2, c.Co, 1, 1.0 / s["world"], None, s["live"], E, a.gaddr(bn + ".weight"), a.gaddr(bn + ".bias"), n)
            dy = zeros(R, c.Co, c.Ho, c.Wo)
            L.call("es_bn2d_bwd_apply", d, rec["y"], c.Co, P, rec["stats"], sums2, rec["n_sg"], *common, dy)
            has_bias = (name + ".bias") in a.off
            # (a conv bias in front of a BatchNorm has an identically-zero gradient; db is still accumulated: it is free)
            _on_side_stream(wgrad_stream, keep, lambda rec=rec, dy=dy, c=c, name=name, has_bias=has_bias: L.call(
                "es_conv2d_bwd_weight", rec["x"], dy, c, grp, E, R, a.gaddr(name + ".weight"),
                a.gaddr(name + ".bias") if has_bias else None, n, n), rec["x"], dy)
            if li == 0:
                L.call("es_conv2d_bwd_data", dy, a.addr(name + ".weight"), n, c, grp, E, R, d_img, int(accumulate))
            else:
                d = zeros(*rec["x"].shape)
                L.call("es_conv2d_bwd_data", dy, a.addr(name + ".weight"), n, c, grp, E, R, d, 0)


# =====================================================================================================================
def engine_for(arena: Arena, arch: str, kind: str):
    """The compute engine bound to ``arena`` (created on first use).  Generator engines keep bf16 kernel-layout
    copies of the weights and are re-packed whenever ``arena.version`` moved (optimizer step, load_state_dict)."""
    if arena.device.type != "cuda":
        raise RuntimeError("expertsim (B200) computes on CUDA only; there is no CPU fallback")
    if arena.engine is None:
        if kind == "discriminator":
            arena.engine = DiscEngine(arena, arch)
        elif (arch, kind) == ("proton", "generator"):
            arena.engine = GenEngineProton(arena)
        elif (arch, kind) == ("proton", "aux_reg"):
            arena.engine = AuxEngineProton(arena)
        elif (arch, kind) == ("neutron", "generator"):
            arena.engine = GenEngineNeutron(arena)
        elif (arch, kind) == ("neutron", "aux_reg"):
            arena.engine = AuxEngineNeutron(arena)
        else:
            raise NotImplementedError(f"no sm_100a engine for {arch}.{kind}")
    eng = arena.engine
    if hasattr(eng, "repack") and getattr(eng, "packed_version", -1) != arena.version:
        eng.repack()
        eng.packed_version = arena.version
    return eng
