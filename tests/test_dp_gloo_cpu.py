"""Data-parallel contract on CPU (gloo, world_size 2): the step shards by samples and exchanges (a) per-expert partial sums
before the gradient kernels and (b) gradient sums before Adam (models/moe.py).  No CUDA here — these tests pin
the HOST logic: parameter broadcast, sum-all-reduce plumbing, and that all-reduced per-expert partial sums reproduce the
global-batch loss of the oracle exactly (the normalisers are global counts, so SUM — not mean — is the right reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle.expertsim_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run2(fn):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), fn, out), nprocs=2, join=True)
    return dict(out)


def _broadcast_case(rank, world):
    from expertsim.config import load_config
    from expertsim.train.loop import setup_moe_system
    torch.manual_seed(100 + rank)                      # different initial weights per rank ...
    cfg = load_config(None, ["model.n_experts=2"])
    moe = setup_moe_system(cfg, torch.device("cpu"))
    moe.enable_data_parallel()                         # ... made identical by the rank-0 broadcast
    assert moe.world_size == 2
    t = torch.full((3,), float(rank + 1))
    moe._allreduce(t)
    return (float(moe.arena("g").P.double().sum()), float(moe.arena("d").Bf.double().sum()), t.tolist())


def test_enable_data_parallel_broadcasts_and_sums():
    out = run2(_broadcast_case)
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert out[0][2] == out[1][2] == [3.0, 3.0, 3.0]


def _loss_case(rank, world):
    """Per-expert loss tails from sharded partial sums == the oracle on the global batch."""
    E, B, seed = 3, 24, 5
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, E, (B,), generator=g)
    lat1, lat2 = torch.randn(B, 64, generator=g), torch.randn(B, 64, generator=g)
    z1, z2 = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
    std, score = torch.rand(B, 1, generator=g), torch.randn(B, 1, generator=g)
    s, inten = torch.rand(B, 1, generator=g) * 50, torch.rand(B, 1, generator=g) * 50
    mine = torch.arange(B)[rank::world]
    sums = torch.zeros(E, 6, dtype=torch.float64)      # rows, sum std, sum 1/(div+eps), sum |s-I|, sum score, (unused)
    for b in mine.tolist():
        e = int(idx[b])
        a = (lat1[b] - lat2[b]).abs().mean()
        n = (z1[b] - z2[b]).abs().mean()
        div = a / (n + 1e-5)
        sums[e] += torch.tensor([1.0, float(std[b]), float(1.0 / (div + 1e-5)), float((s[b] - inten[b]).abs()), float(score[b]), 0.0],
                                dtype=torch.float64)
    dist.all_reduce(sums)                              # the ONLY exchange before the gradient kernels
    got = []
    for e in range(E):
        n = sums[e, 0]
        mstd = sums[e, 1] / n
        got.append(float((-sums[e, 4] / n + mstd * mstd * (sums[e, 2] / n) * 0.1 + sums[e, 3] / n * 1e-3) * n / B))
    want = []
    for e in range(E):
        m = (idx == e).nonzero(as_tuple=True)[0]
        l = -score[m].mean() + orc.sdi_gan_regularization(lat1[m], lat2[m], z1[m], z2[m], std[m], 0.1) \
            + (s[m] - inten[m]).abs().mean() * 1e-3
        want.append(float(l * m.numel() / B))
    return got, want


def test_sharded_partial_sums_reproduce_the_global_batch_loss():
    out = run2(_loss_case)
    for rank in (0, 1):
        got, want = out[rank]
        for a, b in zip(got, want):
            assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (a, b)
    assert out[0][0] == out[1][0]                       # every rank holds the same global value


def _loader_case(rank, world):
    from expertsim.utils.data import DeviceLoader, synthetic_showers
    data = synthetic_showers("proton", 32, seed=0)
    rows = [b[2] for b in DeviceLoader(data, 4, shuffle=True, rank=rank, world=world, seed=9)]
    gathered = [torch.zeros(len(rows), 4, 9) for _ in range(world)]
    dist.all_gather(gathered, torch.stack(rows))
    return torch.cat([g.reshape(-1, 9) for g in gathered]).sum().item(), data["cond"].sum().item()


def test_sharded_loader_covers_the_dataset_once():
    out = run2(_loader_case)
    assert abs(out[0][0] - out[0][1]) < 1e-3


def _bucket_case(rank, world):
    """Layer buckets of a gradient arena, reduced back to front on the reducer's own group == one whole-arena all-reduce."""
    from expertsim._reduce import BucketedGradReducer
    red = BucketedGradReducer(dist)
    g = torch.Generator().manual_seed(7 + rank)
    G = torch.randn(3, 100, generator=g)
    whole = G.clone()
    dist.all_reduce(whole)
    red.begin()
    for lo, hi in ((60, 100), (20, 60), (20, 20), (0, 20)):
        red.reduce(G, lo, hi)
    red.join()
    covered = red.n_reduced == G.numel() and red.buckets == [(60, 100), (20, 60), (0, 20)]
    G2 = torch.randn(3, 100, generator=g)
    whole2 = G2.clone()
    dist.all_reduce(whole2)
    red.begin()
    red.reduce(G2, 0, 100)
    return bool(torch.equal(G, whole)), covered, bool(torch.equal(G2, whole2))


def _chunked_case(rank, world):
    """reduce_chunked == the plain all-reduce of the same column range, for row blocks (E >= chunks) and column blocks
    (E < chunks); the rectangles tile the bucket exactly once, in issue order"""
    from expertsim._reduce import BucketedGradReducer
    red = BucketedGradReducer(dist)
    ok = []
    for E, cols, lo, hi, chunks in ((8, 5000, 1000, 4600, 4), (3, 9000, 500, 8200, 4), (1, 9000, 0, 9000, 4), (5, 300, 10, 290, 2)):
        g = torch.Generator().manual_seed(3 + rank)
        G = torch.randn(E, cols, generator=g)
        want = G.clone()
        part = G[:, lo:hi].clone()
        dist.all_reduce(part)
        want[:, lo:hi] = part
        red.begin()
        rects = red.reduce_chunked(G, lo, hi, chunks)
        cover = torch.zeros(E, cols, dtype=torch.int32)
        for e0, e1, c0, c1, ev in rects:
            cover[e0:e1, c0:c1] += 1
            assert ev is None
        tiled = bool((cover[:, lo:hi] == 1).all()) and int(cover.sum()) == E * (hi - lo)
        ok.append(bool(torch.equal(G, want)) and tiled and red.n_reduced == E * (hi - lo) and red.buckets == [(lo, hi)]
                  and red.bytes_sent == 4 * E * (hi - lo) and len(rects) >= min(chunks, E))
    return ok


def test_chunked_bucket_equals_plain_allreduce():
    out = run2(_chunked_case)
    assert out[0] == [True] * 4 and out[1] == [True] * 4


def _compressed_bucket_case(rank, world):
    """A bucket at least ``compress_min_cols`` wide travels as bf16: result = the bf16 sum of the bf16-rounded rows, identical
    on every rank; narrower buckets stay fp32 and exact."""
    from expertsim._reduce import BucketedGradReducer
    red = BucketedGradReducer(dist)
    red.compress_min_cols = 64
    g = torch.Generator().manual_seed(11 + rank)
    G = torch.randn(2, 100, generator=g)
    mine = G.clone()
    parts = [torch.zeros_like(G) for _ in range(world)]
    dist.all_gather(parts, mine)
    red.begin()
    red.reduce(G, 30, 100)       # 70 columns: compressed
    red.reduce(G, 0, 30)         # 30 columns: fp32
    red.join()
    want_c = sum(p[:, 30:].to(torch.bfloat16).float() for p in parts).to(torch.bfloat16).float()
    want_f = sum(p[:, :30] for p in parts)
    allg = [torch.zeros_like(G) for _ in range(world)]
    dist.all_gather(allg, G)
    same = all(torch.equal(allg[0], a) for a in allg)
    return (bool(torch.allclose(G[:, 30:], want_c, rtol=2 ** -7, atol=1e-6)), bool(torch.equal(G[:, :30], want_f)), same,
            red.bytes_sent == 2 * 70 * 2 + 2 * 30 * 4)


def test_large_gradient_buckets_travel_as_bf16():
    out = run2(_compressed_bucket_case)
    assert out[0] == (True, True, True, True) and out[1] == (True, True, True, True)


def test_bucketed_gradient_reducer_equals_whole_arena_allreduce():
    out = run2(_bucket_case)
    assert out[0] == (True, True, True) and out[1] == (True, True, True)


def test_generator_backward_buckets_tile_the_arena():
    """The hook points of both generator engines hand out column ranges that tile [0, n) exactly once, back to front
    (host logic only: the ranges come from the arena's parameter offsets)."""
    from expertsim._arena import Arena, spec_for
    for arch, firsts in (("proton", ["conv_layers.8.weight", "conv_layers.5.weight", "conv_layers.1.weight", "fc2.0.weight", None]),
                         ("neutron", ["conv_layers.9.weight", "conv_layers.5.weight", "conv_layers.0.weight", "fc2.0.weight", None])):
        a = Arena(spec_for(arch, "generator"), 1, "cpu")
        hi, got = a.n, []
        for f in firsts:
            lo = a.off[f] if f else 0
            got.append((lo, hi))
            hi = lo
        assert got[-1][0] == 0 and all(x[0] == y[1] for x, y in zip(got, got[1:])) and sum(h - l for l, h in got) == a.n
        # every parameter of the layers behind a bucket's first name lies inside that bucket or a later-produced one
        order = list(a.off)
        for (lo, hi2), f in zip(got, firsts):
            if f:
                assert all(a.off[n] >= lo for n in order[order.index(f):])
