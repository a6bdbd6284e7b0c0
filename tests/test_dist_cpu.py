"""The N > 1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200diff.dist import allreduce_mean_, gather_samples, rank_seed, shard_plan


def test_shard_plan_matches_reference_formula():
    # scripts/sample_uncond.py:182-183: bspp = min(batch_size, ceil(n / P)); folds = amortize(n, bspp * P)
    assert shard_plan(50000, 256, 8) == (256, [2048] * 24 + [848])
    assert shard_plan(64, 256, 8) == (8, [64])
    assert shard_plan(100, 16, 1) == (16, [16] * 6 + [4])
    assert shard_plan(10, 256, 4) == (3, [10])
    assert rank_seed(2022, 3) == 2025


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        bspp, folds = shard_plan(10, 4, world)
        g = torch.Generator().manual_seed(rank_seed(2022, rank))
        mine = torch.randn(bspp, 3, 2, 2, generator=g)             # "samples" of this rank, no traffic while sampling
        allx = gather_samples(mine, keep=folds[0])
        grads = [torch.full((5,), float(rank + 1)), torch.full((2, 3), float(10 * (rank + 1)))]
        allreduce_mean_(grads, bucket_bytes=16)
        ret[rank] = (allx.clone(), [t.clone() for t in grads], mine.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_and_allreduce_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert torch.equal(r0[0], r1[0])                                   # every rank sees the same gathered batch
    assert torch.equal(r0[0], torch.cat([r0[2], r1[2]], dim=0)[:8])   # rank order, truncated to the fold size
    assert not torch.equal(r0[2], r1[2])                               # rank-specific seeds
    for t0, t1, want in zip(r0[1], r1[1], (1.5, 15.0)):
        assert torch.allclose(t0, torch.full_like(t0, want)) and torch.equal(t0, t1)


# ------------------------------------------------------------------------------------------------------
# gradient buckets all-reduced while the backward still runs (models/backward.py: _completion_layout, _OverlappedAllReduce)
# ------------------------------------------------------------------------------------------------------
def _fake_net():
    import torch.nn as nn
    torch.manual_seed(0)
    emb = nn.Linear(4, 4)                                   # embedding path: complete last
    blocks = [nn.Conv2d(2, 3, 3) for _ in range(4)]          # forward order b0..b3; backward completes b3 first
    projs = [nn.Linear(4, 3) for _ in range(4)]              # per-block embedding projections (interleaved in registration)
    params, tape = list(emb.parameters()), []
    tape.append(dict(kind='embed', lin1=emb, linears=projs))
    for b, pr in zip(blocks, projs):
        params += list(b.parameters()) + list(pr.parameters())
        tape.append(dict(kind='res', conv1=b, emb_linear=pr))
    return params, tape, blocks, projs, emb


def test_completion_layout_orders_gradients_by_backward_completion():
    from models.backward import _completion_layout
    params, tape, blocks, projs, emb = _fake_net()
    order, _ = _completion_layout(tape, params)
    ids = [id(p) for p in order]
    # the projections' weights lie side by side, then their biases: _embed_bwd writes each group with one launch
    want = [id(p) for b in reversed(blocks) for p in b.parameters()] + [id(p) for p in emb.parameters()] + \
        [id(pr.weight) for pr in projs] + [id(pr.bias) for pr in projs]
    assert ids == want and len(order) == len(params)


def _overlap_worker(rank, world, port, ret):
    import types
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from models.backward import _Grads, _OverlappedAllReduce
        params, tape, blocks, projs, emb = _fake_net()
        eng = types.SimpleNamespace(device=torch.device('cpu'))
        G = _Grads(eng, params, tape)
        ar = _OverlappedAllReduce(G, tape, bucket_bytes=4 * 40)      # ~ one conv per bucket
        sent_after = []
        for e in reversed(tape):
            if e['kind'] == 'embed':
                continue
            for p in e['conv1'].parameters():                        # "this record's backward": gradient = rank + 1
                G(p).fill_(float(rank + 1))
            ar.after_record(e)
            sent_after.append(ar.sent)
        for p in list(emb.parameters()) + [q for pr in projs for q in pr.parameters()]:
            G(p).fill_(float(10 * (rank + 1)))
        ar.finish()
        ret[rank] = (G.flat.clone(), sent_after, ar.n_buckets, [G.offset[id(p)] for p in params])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_overlapped_bucket_allreduce_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_overlap_worker, args=(world, port, ret), nprocs=world, join=True)
        (f0, sent0, nb0, off0), (f1, sent1, nb1, off1) = ret[0], ret[1]
    assert torch.equal(f0, f1) and sent0 == sent1 and nb0 == nb1 and off0 == off1
    n_conv = 4 * (3 * 2 * 9 + 3)
    assert torch.allclose(f0[:n_conv], torch.full((n_conv,), 1.5)) and torch.allclose(f0[n_conv:], torch.full_like(f0[n_conv:], 15.0))
    assert nb0 >= 3 and sent0[0] > 0 and sent0 == sorted(sent0) and sent0[-1] == n_conv   # buckets left DURING the walk
