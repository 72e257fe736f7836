"""CPU: host-side logic of the keyframe-batched data-parallel mapping step (segs_slam_b200/mapper.py)
with the gloo backend, world_size 2 — view partitioning, flat gradient bucket, one all-reduce, and
replica consistency.  The per-view render is replaced by a small differentiable stand-in (the real
one needs the GPU kernels; tests/test_mapper_gpu.py covers it)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from segs_slam_b200 import mapper


def test_partition_views_is_disjoint_and_complete():
    for n, g in [(64, 8), (64, 4), (10, 3), (1, 2), (0, 2)]:
        parts = [mapper.partition_views(n, g, r) for r in range(g)]
        flat = sorted(v for p in parts for v in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        mapper.partition_views(4, 2, 2)


def test_grad_bucket_layout():
    a, b = torch.zeros(5, 3, requires_grad=True), torch.zeros(7, requires_grad=True)
    bk = mapper.GradBucket([a, b])
    assert bk.flat.numel() == 22
    bk.accumulate([torch.ones(5, 3), None])
    bk.accumulate([torch.ones(5, 3), torch.full((7,), 2.0)])
    assert bk.flat[:15].eq(2).all() and bk.flat[15:].eq(2).all()
    bk.all_reduce_mean(4)
    bk.scatter_to_grads()
    assert a.grad.data_ptr() == bk.flat.data_ptr() and torch.allclose(a.grad, torch.full((5, 3), 0.5))


def _make_problem():
    torch.manual_seed(0)
    params = [torch.randn(40, 3, requires_grad=True), torch.randn(16, requires_grad=True)]
    targets = [torch.randn(40) for _ in range(8)]

    def render_loss(v):   # differentiable stand-in for prefilter -> decode -> rasterize -> L1
        img = (params[0] * (1.0 + 0.1 * v)).sum(dim=1) * torch.tanh(params[1]).sum()
        return (img - targets[v]).abs().mean()

    return params, render_loss


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, render_loss = _make_problem()
    opt = torch.optim.Adam(params, lr=1e-2)
    losses = []
    for _ in range(3):
        loss, _ = mapper.mapping_step(params, render_loss, 8, opt)
        losses.append(float(loss))
    out[rank] = ([p.detach().clone() for p in params], losses)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_step_matches_single_process():
    params, render_loss = _make_problem()
    opt = torch.optim.Adam(params, lr=1e-2)
    ref_losses = [float(mapper.mapping_step(params, render_loss, 8, opt)[0]) for _ in range(3)]

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (p0, l0), (p1, l1) = out[0], out[1]
    for a, b in zip(p0, p1):
        assert torch.equal(a, b), "replicas diverged"                 # same bucket, same step on both ranks
    for a, b in zip(p0, params):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)    # == sequential accumulation (sum order differs)
    assert l0 == l1
    assert all(abs(x - y) < 1e-5 for x, y in zip(l0, ref_losses))


def _exchange_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank holds the gradients / loss / statistics of ITS views only (rank r: views r, r + 2, ...)
    views = mapper.partition_views(6, world, rank)
    grad = torch.zeros(10)
    loss = torch.zeros(())
    delta, stats = torch.zeros(4), torch.full((4,), 100.0)
    for v in views:
        grad += torch.arange(10, dtype=torch.float32) * (v + 1)
        loss += 0.5 * (v + 1)
        delta += torch.tensor([1.0, float(v), 0.0, 2.0])
    total = mapper.exchange_step(grad, loss, delta, stats)
    out[rank] = (grad.clone(), float(total), stats.clone(), float(loss))
    dist.destroy_process_group()


def test_exchange_step_sums_gradients_loss_and_statistics_over_ranks():
    """FusedMapper's rank exchange (gradient bucket, loss, densification-statistics delta) on gloo, world size 2."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (g0, l0, s0, own0), (g1, l1, s1, own1) = out[0], out[1]
    assert torch.equal(g0, g1) and torch.equal(g0, torch.arange(10, dtype=torch.float32) * 21)     # 1 + ... + 6
    assert l0 == l1 == 0.5 * 21 and own0 != own1                    # the local accumulators are left alone
    assert torch.equal(s0, s1) and torch.equal(s0, torch.tensor([106.0, 115.0, 100.0, 112.0]))
    # single process: no collective, statistics still accumulate
    stats = torch.zeros(2)
    total = mapper.exchange_step(torch.ones(3), torch.tensor(2.0), torch.tensor([1.0, 2.0]), stats)
    assert float(total) == 2.0 and torch.equal(stats, torch.tensor([1.0, 2.0]))


def _fused_exchange_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # FusedMapper's layout: gradients | loss | statistics delta in ONE allocation -> one all-reduce
    bk = mapper.GradBucket([torch.zeros(6, 2), torch.zeros(3)], tail=1 + 4)
    loss, delta = bk.tail[:1].view(()), bk.tail[1:]
    stats = torch.zeros(4)
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    try:
        bk.flat += float(rank + 1)
        loss += 0.25 * (rank + 1)
        delta += torch.tensor([1.0, 0.0, float(rank), 2.0])
        total = mapper.exchange_step(bk.flat, loss, delta, stats)
    finally:
        dist.all_reduce = real
    out[rank] = (bk.flat.clone(), float(total), stats.clone(), calls)
    dist.destroy_process_group()


def test_exchange_step_is_one_collective_when_everything_shares_an_allocation():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fused_exchange_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in (0, 1):
        g, l, s, calls = out[r]
        assert calls == [15 + 1 + 4], calls                       # ONE all-reduce over gradients + loss + delta
        assert torch.equal(g, torch.full((15,), 3.0)) and l == 0.75 and torch.equal(s, torch.tensor([2.0, 0.0, 1.0, 4.0]))


def _densify_worker(rank, world, port, out):
    """Two replicas with identical parameters accumulate the densification statistics of THEIR views, exchange the delta,
    and take the densification decision with a generator seeded identically on every rank: the anchor sets must come out
    identical (the decision routine here is the restated oracle — the CUDA one needs a GPU, tests/test_densify_gpu.py)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    sys.path.insert(0, os.path.join(root, "tests"))
    import densify_cases as dc
    import densify_oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = 400
    st, _g = dc.make_state(A, 9)
    # split the statistics into per-rank contributions (views seen by rank 0 / rank 1)
    stat_names = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")
    full = torch.cat([st[k].reshape(-1) for k in stat_names])
    g = torch.Generator().manual_seed(3)
    share = torch.rand(full.numel(), generator=g)
    mine = torch.where((share < 0.5) == (rank == 0), full, torch.zeros_like(full))      # disjoint, sums to `full` exactly
    bk = mapper.GradBucket([torch.zeros(4)], tail=1 + full.numel())
    loss, delta = bk.tail[:1].view(()), bk.tail[1:]
    delta.copy_(mine)
    stats = torch.zeros_like(full)
    mapper.exchange_step(bk.flat, loss, delta, stats)
    sizes = [st[k].numel() for k in stat_names]
    for k, chunk in zip(stat_names, torch.split(stats, sizes)):
        st[k] = chunk.view_as(st[k]).clone()
    gen = torch.Generator().manual_seed(20260101)                   # the shared densification seed
    rands = [torch.rand(A * 10, generator=gen) for _ in range(dc.MODEL["update_depth"])]
    res = densify_oracle.adjust_anchor(st, rands, 100, 0.8, 0.0002, 0.005, **dc.MODEL)
    out[rank] = {k: res[k].clone() for k in ("_anchor", "_anchor_feat", "_scaling", "offset_denom", "anchor_demon")}
    dist.destroy_process_group()


def test_replicas_take_identical_densification_decisions():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_densify_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["_anchor"].size(0) != 400                              # something was grown / pruned
    for k in a:
        assert torch.equal(a[k], b[k]), k
