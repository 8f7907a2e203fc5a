"""Inter-video support exchange: the one collective of the path.

At inference the shipped reference has no inter-video step (forward_test pools one window,
hrnmp_bbox_head.py:800-909); the stage exists in the training forward (:740-795).  BASELINE.json
configs 4-5 define it for inference: stage 4 of key frame g attends, in
addition to its own window, to the post-fc_new_4 key-frame rows Z of `n_support` other videos,
chosen by ring order over the global list of key frames:  (g+1 .. g+n_support) mod G  - or, with
video descriptors (next row N4, hnmb_rcnn.py:76-101), the n_support most similar other videos.

Sharding: each rank owns V key frames (global index g = rank*V + v), computes stages 1-3
locally, and ONE all-gather of Z (split-bf16 pair = 4 bytes / element, bit-exact transport of
what a single GPU would hold) gives every rank the pool [G*P, D]; stage 4 then runs locally.
The functions below are device-agnostic torch.distributed code (NCCL on the GPUs; the gloo
CPU tests drive the same code path with world_size 2).
"""
import torch
import torch.distributed as dist

from .ops import Split


def support_indices(g, total, n_support):
    """Global key-frame indices whose rows support key frame g (ring order, self excluded)."""
    n = min(n_support, total - 1)
    return [(g + 1 + i) % total for i in range(n)]


def shard_range(n_items, world, rank):
    """Contiguous, count-balanced shard [lo, hi) of n_items for `rank` (key frames of bench.py's synthetic
    videos, which all have the same length; real videos go through shard_videos below)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_videos(seg_lens, world):
    """Whole videos to ranks as the reference's distributed test does (VIDSeqDataset.get_indices,
    imagenet_vid_sequence.py:117-158): videos stay in order; a rank takes videos while its frame count stays
    within ceil(total_frames / world), the video that would exceed it opens the next rank, and the last rank
    takes everything that is left.  Returns, per rank, the list of video indices (pinned against the
    reference's own method by tests/golden/ref_shard_golden.json)."""
    avg = -(-int(sum(seg_lens)) // world)
    out = [[] for _ in range(world)]
    rank, used = 0, 0
    for v, n in enumerate(seg_lens):
        if used + n > avg and rank != world - 1:
            rank, used = rank + 1, 0
        out[rank].append(v)
        used += n
    return out


def all_gather_rows(z, group=None):
    """z: Split [V*P, D] (same V*P on every rank) -> Split [world*V*P, D], rank-major.
    One collective: hi and lo travel as one [2, V*P, D] bf16 tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return z
    world = dist.get_world_size(group)
    send = torch.stack([z.hi, z.lo]).contiguous()
    rows = z.hi.shape[0]
    recv = torch.empty((world * 2,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)       # rank-major: [hi_0, lo_0, hi_1, lo_1, ...]
    recv = recv.view(world, 2, rows, -1)
    return Split(recv[:, 0].reshape(world * rows, -1), recv[:, 1].reshape(world * rows, -1))


def _desc_rows(desc, D):
    """fp32 descriptors [V, C] -> bf16 carrier rows [V*r, D] holding their bytes (r rows per video), so
    that they travel inside the one all-gather of the support rows."""
    V, C = desc.shape
    r = -(-(C * 4) // (D * 2))
    rows = torch.zeros((V, r * D * 2), dtype=torch.uint8, device=desc.device)
    rows[:, :C * 4] = desc.contiguous().float().view(torch.uint8).view(V, C * 4)
    return rows.view(torch.bfloat16).view(V * r, D), r


def _desc_from_rows(rows, V, r, C):
    """Inverse of _desc_rows: carrier rows [V*r, D] (bf16) -> fp32 [V, C]."""
    return rows.contiguous().view(torch.uint8).view(V, -1)[:, :C * 4].contiguous().view(torch.float32).view(V, C)


def select_by_similarity(desc_all, g0, n_local, n_support):
    """Product selector: hvr_support_select on the device, one small D2H read of the indices."""
    from . import ops
    idx = ops.support_select(desc_all, g0, n_local, n_support).cpu().tolist()
    return [[i for i in row if i >= 0] for row in idx]


def gather_support(z_local, rows_per_key, n_support, group=None, async_stream=None, desc_local=None,
                   selector=select_by_similarity):
    """Exchange + selection.  z_local: Split [V*P, D] of this rank's V key frames.
    Returns a list of V Splits [n_sel*P, D]: the support rows of each local key frame.
    desc_local None: ring order.  desc_local fp32 [V, C] (ops.video_descriptor of each local video): the
    n_support most similar other videos; the descriptors ride in the same all-gather as extra rows and
    `selector(desc_all [G,C], g0, V, n_support)` returns the chosen global indices per local key frame
    (the gloo CPU tests pass a CPU selector; the default runs the CUDA kernel)."""
    P = rows_per_key
    V = z_local.hi.shape[0] // P
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    G = world * V
    chosen = None
    if desc_local is None:
        pool = all_gather_rows(z_local, group)
    else:
        D, C = z_local.hi.shape[1], desc_local.shape[1]
        carrier, r = _desc_rows(desc_local, D)
        sent = Split(torch.cat([z_local.hi, carrier], 0), torch.cat([z_local.lo, torch.zeros_like(carrier)], 0))
        got = all_gather_rows(sent, group)                     # still ONE collective
        per = V * P + V * r
        hi, lo = got.hi.view(world, per, D), got.lo.view(world, per, D)
        pool = Split(hi[:, :V * P].reshape(G * P, D), lo[:, :V * P].reshape(G * P, D))
        desc_all = _desc_from_rows(hi[:, V * P:].reshape(G * r, D), G, r, C)
        chosen = selector(desc_all, rank * V, V, min(n_support, G - 1)) if G > 1 else [[] for _ in range(V)]
    out = []
    for v in range(V):
        idx = support_indices(rank * V + v, G, n_support) if chosen is None else chosen[v]
        if not idx:
            out.append(Split(pool.hi[:0], pool.lo[:0]))
            continue
        out.append(Split(torch.cat([pool.hi[i * P:(i + 1) * P] for i in idx], 0),
                         torch.cat([pool.lo[i * P:(i + 1) * P] for i in idx], 0)))
    return out
