"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of key frames over
ranks, the single all-gather of inter-video support rows and the ring selection."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, V, P, D, n_support, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from hvrnet_b200 import intervideo
        from hvrnet_b200.ops import Split
        # row r of global key frame g carries the value g*16 + r (exact in bf16) in hi and its negative in lo
        rows = []
        for v in range(V):
            g = rank * V + v
            rows.append(torch.arange(P, dtype=torch.float32).view(P, 1).expand(P, D) + 16.0 * g)
        z = torch.cat(rows, 0)
        zl = Split(z.to(torch.bfloat16), (-z).to(torch.bfloat16))
        sup = intervideo.gather_support(zl, P, n_support)
        ok = len(sup) == V
        for v in range(V):
            g = rank * V + v
            idx = intervideo.support_indices(g, world * V, n_support)
            exp = torch.cat([(torch.arange(P, dtype=torch.float32) + 16.0 * i).view(P, 1).expand(P, D) for i in idx], 0)
            ok = ok and torch.equal(sup[v].hi, exp.to(torch.bfloat16)) and torch.equal(sup[v].lo, (-exp).to(torch.bfloat16))
        pool = intervideo.all_gather_rows(zl)
        ok = ok and pool.hi.shape == (world * V * P, D) and float(pool.hi[(world * V - 1) * P, 0]) == 16.0 * (world * V - 1)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('V,n_support', [(3, 4), (1, 4), (2, 1)])
def test_support_exchange_world2(V, n_support):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, 4, 8, n_support, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=150)
    assert [p.exitcode for p in procs] == [0] * world
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {0: True, 1: True}


def _worker_similarity(rank, world, port, V, P, D, C, n_support, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from hvrnet_b200 import intervideo
        from hvrnet_b200.ops import Split
        from oracle import ref_torch as R
        G = world * V
        gen = torch.Generator().manual_seed(7)
        desc_all = torch.randn(G, C, generator=gen) * 3          # every rank can rebuild the global truth
        z_all = torch.cat([(torch.arange(P, dtype=torch.float32) + 16.0 * g).view(P, 1).expand(P, D) for g in range(G)], 0)
        lo_, hi_ = rank * V * P, (rank + 1) * V * P
        zl = Split(z_all[lo_:hi_].to(torch.bfloat16), (-z_all[lo_:hi_]).to(torch.bfloat16))
        calls = []

        def selector(d, g0, n_local, k):                          # CPU stand-in for hvr_support_select
            calls.append((tuple(d.shape), g0, n_local, k, bool(torch.equal(d, desc_all))))
            return [R.select_support_by_similarity(d, g0 + i, k)[0] for i in range(n_local)]
        sup = intervideo.gather_support(zl, P, n_support, desc_local=desc_all[rank * V:(rank + 1) * V], selector=selector)
        ok = len(sup) == V and calls == [((G, C), rank * V, V, min(n_support, G - 1), True)]
        for v in range(V):
            idx = R.select_support_by_similarity(desc_all, rank * V + v, n_support)[0]
            exp = torch.cat([z_all[i * P:(i + 1) * P] for i in idx], 0)
            ok = ok and torch.equal(sup[v].hi, exp.to(torch.bfloat16)) and torch.equal(sup[v].lo, (-exp).to(torch.bfloat16))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('V,C,n_support', [(3, 256, 2), (2, 1500, 4)])
def test_similarity_support_exchange_world2(V, C, n_support):
    """Next row N4 at world_size 2: the fp32 descriptors ride in the one all-gather as carrier rows (C = 1500
    needs two rows of D = 1024 bf16 per video), every rank sees the same [G, C] table bit for bit, and the
    selected rows are those of the oracle's selection."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_similarity, args=(r, world, port, V, 4, 1024, C, n_support, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=150)
    assert [p.exitcode for p in procs] == [0] * world
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {0: True, 1: True}


def test_ring_rule_and_sharding():
    from hvrnet_b200 import intervideo as iv
    assert iv.support_indices(0, 256, 4) == [1, 2, 3, 4]
    assert iv.support_indices(254, 256, 4) == [255, 0, 1, 2]          # (g*32+b+1..+4) mod 256, SURVEY.md 8d
    assert iv.support_indices(0, 1, 4) == []                          # no other video: degenerates to forward_test
    assert iv.support_indices(1, 3, 4) == [2, 0]                      # never includes itself
    cover = []
    for r in range(8):
        lo, hi = iv.shard_range(555, 8, r)
        cover += list(range(lo, hi))
        assert 69 <= hi - lo <= 70
    assert cover == list(range(555))


def test_single_process_is_identity():
    from hvrnet_b200 import intervideo as iv
    from hvrnet_b200.ops import Split
    z = Split(torch.randn(6, 4).to(torch.bfloat16), torch.randn(6, 4).to(torch.bfloat16))
    assert iv.all_gather_rows(z) is z
    sup = iv.gather_support(z, 2, 4)
    assert [tuple(s.hi.shape) for s in sup] == [(4, 4)] * 3
    assert torch.equal(sup[2].hi, torch.cat([z.hi[0:2], z.hi[2:4]]))


def test_shard_videos_matches_reference_get_indices():
    """Multi-GPU partition of real videos: tests/golden/ref_shard_golden.json holds what the REFERENCE's own
    VIDSeqDataset.get_indices (imagenet_vid_sequence.py:117-158, cut out of the file and run on a bare object,
    tests/golden/make_shard_golden.py) hands to each rank for 20 seeded (video lengths, world size) cases,
    including ranks that receive nothing; intervideo.shard_videos reproduces every one."""
    import json
    from hvrnet_b200 import intervideo as iv
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_shard_golden.json')))
    assert len(cases) == 20 and any([] in c['videos'] for c in cases)
    for c in cases:
        got = iv.shard_videos(c['seg_lens'], c['world'])
        assert got == c['videos']
        assert [sum(c['seg_lens'][v] for v in r) for r in got] == c['frames']
