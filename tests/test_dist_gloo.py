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


def _worker(rank, world, port, V, P, D, C, n_support, q):
    """One rank of the inter-video exchange on the CPU: intervideo.pack_exchange -> ONE all-gather
    (intervideo.exchange) -> the receive buffer every rank addresses with hvr_support_index's formula."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from hvrnet_b200 import intervideo, window
        from hvrnet_b200.ops import Split
        G = world * V
        gen = torch.Generator().manual_seed(7)
        desc_all = torch.randn(G, C, generator=gen) * 3 if C else None   # every rank can rebuild the global truth
        counts_all = torch.randint(1, P + 1, (G,), generator=gen, dtype=torch.int32)
        # row j of global key frame g carries the value g*16 + j (exact in bf16) in hi and its negative in lo
        z_all = torch.cat([(torch.arange(P, dtype=torch.float32) + 16.0 * g).view(P, 1).expand(P, D) for g in range(G)], 0)
        lo_, hi_ = rank * V * P, (rank + 1) * V * P
        z = Split(z_all[lo_:hi_].to(torch.bfloat16).contiguous(), (-z_all[lo_:hi_]).to(torch.bfloat16).contiguous())
        send, rpr, n_desc = intervideo.pack_exchange(z, counts_all[rank * V:(rank + 1) * V].contiguous(),
                                                     desc_all[rank * V:(rank + 1) * V] if C else None)
        recv, work = intervideo.exchange(send)
        ok = work is None and tuple(recv.shape) == (world, 2, rpr, D) and rpr == V * P + 1 + n_desc
        flat = recv.view(world * 2 * rpr, D)
        # the addressing hvr_support_index uses: key frame g, row j -> hi row (g // V) * 2*rpr + (g % V) * P + j, lo row + rpr
        sel = window.ring_selection(rank, world, V, n_support, 'cpu')
        for v in range(V):
            want = intervideo.support_indices(rank * V + v, G, n_support)
            ok = ok and sel[v].tolist() == want + [-1] * (n_support - len(want))
            for g in want:
                r0 = (g // V) * 2 * rpr + (g % V) * P
                ok = ok and torch.equal(flat[r0:r0 + P], z_all[g * P:(g + 1) * P].to(torch.bfloat16))
                ok = ok and torch.equal(flat[r0 + rpr:r0 + rpr + P], (-z_all[g * P:(g + 1) * P]).to(torch.bfloat16))
        # counts carrier: pool_counts[(g // V) * counts_rank_stride + g % V] with the int32 view starting at row V*P
        ci = flat.view(torch.int32)[V * P:].reshape(-1)
        stride = 2 * rpr * (D // 2)
        ok = ok and [int(ci[(g // V) * stride + g % V]) for g in range(G)] == counts_all.tolist()
        if C:
            ok = ok and torch.equal(intervideo.unpack_descriptors(recv, V, P, n_desc, C), desc_all)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('V,C,n_support', [(3, 0, 4), (1, 0, 4), (2, 0, 1), (3, 256, 2), (2, 1500, 4)])
def test_support_exchange_world2(V, C, n_support):
    """The single collective of the path at world_size 2 (gloo): key rows (split pair, bit-exact), per-key-frame
    proposal counts and - C > 0 - fp32 video descriptors (C = 1500 needs two carrier rows of D = 1024 bf16 per
    video) travel in ONE all_gather_into_tensor; every rank finds every other rank's rows, counts and descriptors
    at the addresses the device-side index kernel computes; ring selection = intervideo.support_indices."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    D = 1024 if C else 8
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, 4, D, C, n_support, q)) for r in range(world)]
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
    """Without a process group the exchange is the send buffer itself (world 1: configs[3], one GPU)."""
    from hvrnet_b200 import intervideo as iv, window
    from hvrnet_b200.ops import Split
    z = Split(torch.randn(6, 8).to(torch.bfloat16), torch.randn(6, 8).to(torch.bfloat16))
    send, rpr, n_desc = iv.pack_exchange(z, torch.tensor([2, 1, 2], dtype=torch.int32))
    recv, work = iv.exchange(send)
    assert work is None and recv.shape == (1, 2, 7, 8) and recv.data_ptr() == send.data_ptr() and n_desc == 0
    assert torch.equal(recv[0, 0, :6], z.hi) and torch.equal(recv[0, 1, :6], z.lo)
    assert recv[0, 0, 6].view(torch.int32)[:3].tolist() == [2, 1, 2]
    assert window.ring_selection(0, 1, 3, 4, 'cpu').tolist() == [[1, 2, -1, -1], [2, 0, -1, -1], [0, 1, -1, -1]]


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
