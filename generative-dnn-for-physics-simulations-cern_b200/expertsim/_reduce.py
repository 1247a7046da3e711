"""Bucketed gradient all-reduce of the data-parallel step, overlapped with backward.

The gradient arena of a network kind is ONE tensor ``G [E, n]`` (expert e's gradients at ``G[e]``); a *bucket* is a column
range ``[lo, hi)`` of it — the parameters of one or more consecutive layers, for every expert.  Backward produces the
buckets from the last layer to the first; as soon as a bucket's last kernel has been enqueued the engine calls
``reduce(G, lo, hi)``: the sum-all-reduce of that range is enqueued on a COMMUNICATION stream behind an event recorded on
the compute stream, so NCCL moves it over NVLink while the remaining backward kernels run.  ``join()`` makes the compute
stream wait for all outstanding buckets (called once, before the fused Adam).

The reducer owns a process group of its own: ProcessGroupNCCL serialises the collectives of one communicator on one
internal stream, so a 94 MB gradient bucket queued on the default group would sit in front of the small latency-critical
all-reduces the step issues on the compute path (per-expert loss sums, SyncBN statistics).

OPTIONAL (``ES_DP_COMPRESS=1``): buckets at least ``compress_min_cols`` wide (8 Mi columns per expert row: only fc2's 23.6
M-parameter weight gradient — 755 of the 868 MB a proton step all-reduces, produced LAST by backward and therefore the part
that stays exposed) travel as bf16: the column range is cast to one contiguous bf16 buffer, summed, and cast back in place,
all on the communication stream; every rank still receives the same bits.  Measured at N = 2 on B200: 32.35 ms/step against
32.17 with fp32 messages — over NVLink 5 the two cast passes (1.1 GB of HBM traffic each) cost what the halved message
saves, so fp32 stays the default.

On CPU tensors (gloo; the host-logic tests) the same calls run synchronously.
"""
from __future__ import annotations

import os

import torch


class BucketedGradReducer:
    def __init__(self, dist, parent_group=None):
        self.dist = dist
        ranks = dist.get_process_group_ranks(parent_group) if parent_group is not None else list(range(dist.get_world_size()))
        self.group = dist.new_group(ranks=ranks)          # collective: every rank of the parent group constructs one
        self._comm = None
        self.n_reduced = 0                                # floats handed to all_reduce since the last join()
        self.buckets = []                                 # (lo, hi) of the last step, for tests / logging
        self.disabled = False                             # measurement aid (bench.py's comm.exposed_ms): account, do not send
        # buckets at least this wide (columns per expert row) are all-reduced as bf16; None = never
        self.compress_min_cols = (8 << 20) if os.environ.get("ES_DP_COMPRESS", "0") == "1" else None
        self.bytes_sent = 0                               # payload bytes handed to all_reduce since begin()

    def _stream(self, device):
        if self._comm is None or self._comm.device != device:
            self._comm = torch.cuda.Stream(device=device)
        return self._comm

    def begin(self):
        self.n_reduced, self.buckets, self.bytes_sent = 0, [], 0

    def reduce(self, G: torch.Tensor, lo: int, hi: int):
        """SUM-all-reduce ``G[:, lo:hi]`` (one contiguous message per expert row)."""
        if hi <= lo:
            return
        self.buckets.append((lo, hi))
        self.n_reduced += (hi - lo) * G.shape[0]
        if self.disabled:
            return
        if not G.is_cuda:
            self._all_reduce_rows(G, lo, hi)
            return
        ev = torch.cuda.current_stream(G.device).record_event()
        comm = self._stream(G.device)
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            self._all_reduce_rows(G, lo, hi)

    def reduce_chunked(self, G: torch.Tensor, lo: int, hi: int, n_chunks: int = 4):
        """``reduce(G, lo, hi)`` as ``n_chunks`` consecutive collectives over row blocks ``[e0, e1)`` (or, when the arena
        has fewer rows than chunks, over column blocks of every row).  Returns ``[(e0, e1, c0, c1, event)]`` in issue order:
        ``event`` (None on CPU tensors / when disabled) fires on the communication stream when that rectangle of G holds
        the sum, so the caller can run the optimizer over rectangle k while rectangle k + 1 is still on the wire
        (MoEWrapper._adam_pipelined).  Never compressed: every rectangle is summed in place."""
        if hi <= lo:
            return []
        E = G.shape[0]
        n_chunks = max(1, int(n_chunks))
        if E >= n_chunks:
            per = -(-E // n_chunks)
            rects = [(e0, min(e0 + per, E), lo, hi) for e0 in range(0, E, per)]
        else:
            per = -(-(hi - lo) // (n_chunks // E or 1))
            per = -(-per // 1024) * 1024                  # column blocks stay 4 KB aligned (vectorised Adam, NCCL)
            rects = [(e, e + 1, c0, min(c0 + per, hi)) for e in range(E) for c0 in range(lo, hi, per)]
        self.buckets.append((lo, hi))
        self.n_reduced += (hi - lo) * E
        if self.disabled:
            return [r + (None,) for r in rects]
        if not G.is_cuda:
            for e0, e1, c0, c1 in rects:
                self._all_reduce_rect(G, e0, e1, c0, c1)
            return [r + (None,) for r in rects]
        ev = torch.cuda.current_stream(G.device).record_event()
        comm = self._stream(G.device)
        out = []
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            for e0, e1, c0, c1 in rects:
                self._all_reduce_rect(G, e0, e1, c0, c1)
                out.append((e0, e1, c0, c1, comm.record_event()))
        return out

    def _all_reduce_rect(self, G, e0, e1, c0, c1):
        self.bytes_sent += (c1 - c0) * (e1 - e0) * G.element_size()
        rows = [G[e, c0:c1] for e in range(e0, e1)]
        if len(rows) > 1 and G.is_cuda and self.dist.get_backend(self.group) == "nccl":
            with self.dist._coalescing_manager(group=self.group, device=G.device, async_ops=False):
                for t in rows:
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            for t in rows:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def _all_reduce_rows(self, G, lo, hi):
        """one message per expert row, all rows of the bucket in ONE grouped NCCL launch (ncclGroupStart/End)"""
        if self.compress_min_cols is not None and hi - lo >= self.compress_min_cols:
            buf = G[:, lo:hi].to(torch.bfloat16)          # contiguous [E, hi - lo] copy: one message
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
            G[:, lo:hi].copy_(buf)
            self.bytes_sent += buf.numel() * 2
            return
        self.bytes_sent += (hi - lo) * G.shape[0] * G.element_size()
        if lo == 0 and hi == G.shape[1]:
            self.dist.all_reduce(G, op=self.dist.ReduceOp.SUM, group=self.group)
            return
        rows = [G[e, lo:hi] for e in range(G.shape[0])]
        if G.is_cuda and self.dist.get_backend(self.group) == "nccl":
            with self.dist._coalescing_manager(group=self.group, device=G.device, async_ops=False):
                for t in rows:
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            for t in rows:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def join(self):
        """The current stream waits for every bucket enqueued so far."""
        if self._comm is not None:
            torch.cuda.current_stream(self._comm.device).wait_stream(self._comm)
