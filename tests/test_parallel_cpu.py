"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: gradient all-reduce for scene-batch DP, the
collective wrapper and the log-sum-exp merge used by view sharding, and the view partition."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgcdet_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _grads(rank, world):
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    parallel.allreduce_gradients(params)
    return [p.grad.clone() for p in params]


def test_allreduce_gradients_averages():
    out = _run('_grads')
    for r in range(2):
        assert torch.equal(out[r][0], torch.full((3, 4), 1.5))   # mean of 1, 2
        assert torch.equal(out[r][1], torch.full((5,), 2.5))     # mean of 2, 3


def _lse(rank, world):
    """Each rank owns half of the views; merged partial softmax == softmax over all views."""
    g = torch.Generator().manual_seed(3)
    V, Q, Hh, C = 10, 7, 8, 16
    scores = torch.randn(V, Q, Hh, generator=g) * 3
    vis = torch.rand(V, Q, generator=g) < 0.5
    vis[:, 0] = False                       # voxel seen by nobody
    vis[:5, 1] = False                      # voxel seen only by rank 1's views
    slots = torch.randn(V, Q, C, generator=g)
    views = parallel.shard_views(V, world, rank)
    sc = scores[views.start:views.stop].masked_fill(~vis[views.start:views.stop].unsqueeze(-1), -3.0e38)
    coll = parallel.Collective()
    m = coll.reduce([sc.max(0).values], 'max')                                   # exchange 2a
    e = torch.exp(sc - m) * vis[views.start:views.stop].unsqueeze(-1)
    s = coll.reduce([e.sum(0)], 'sum')                                           # exchange 2b
    o = coll.reduce([torch.einsum('vqh,vqc->qhc', e, slots[views.start:views.stop])], 'sum')
    t = o / s.clamp(min=1e-30).unsqueeze(-1)
    # reference: masked softmax over all views
    full = scores.masked_fill(~vis.unsqueeze(-1), float('-inf'))
    alpha = torch.softmax(full, dim=0).nan_to_num(0.0)
    t_ref = torch.einsum('vqh,vqc->qhc', alpha, slots)
    # and the rescale formulation
    m_parts = [scores[v.start:v.stop].masked_fill(~vis[v.start:v.stop].unsqueeze(-1), -3.0e38).max(0).values
               for v in (parallel.shard_views(V, world, r) for r in range(world))]
    return torch.allclose(t, t_ref, atol=1e-5), float((t - t_ref).abs().max()), [float(x.max()) for x in m_parts]


def test_partial_softmax_merge_over_gloo():
    out = _run('_lse')
    assert all(o[0] for o in out), out


def test_merge_partial_softmax_rescale_form():
    g = torch.Generator().manual_seed(1)
    sc = torch.randn(3, 6, 5, 8, generator=g) * 4          # 3 shards x 6 views each
    sl = torch.randn(3, 6, 5, 16, generator=g)
    m_p = [x.max(0).values for x in sc]
    s_p = [torch.exp(x - m).sum(0) for x, m in zip(sc, m_p)]
    o_p = [torch.einsum('vqh,vqc->qhc', torch.exp(x - m), y) for x, m, y in zip(sc, m_p, sl)]
    m, s, o = parallel.merge_partial_softmax(m_p, s_p, o_p)
    alpha = torch.softmax(sc.reshape(18, 5, 8), dim=0)
    ref = torch.einsum('vqh,vqc->qhc', alpha, sl.reshape(18, 5, 16))
    torch.testing.assert_close(o / s.unsqueeze(-1), ref, rtol=1e-4, atol=1e-5)


def test_shard_views_partition():
    for V in (40, 100, 7):
        for W in (1, 2, 4, 8):
            parts = [parallel.shard_views(V, W, r) for r in range(W)]
            assert sum(len(p) for p in parts) == V
            assert parts[0].start == 0 and parts[-1].stop == V
            assert all(parts[i].stop == parts[i + 1].start for i in range(W - 1))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
