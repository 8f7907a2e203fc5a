"""Inter-video support exchange: the one collective of the path.

At inference the shipped reference has no inter-video step (forward_test pools one window,
hrnmp_bbox_head.py:800-909); the stage exists in the training forward (:740-795).  BASELINE.json
configs 4-5 define it for inference: stage 4 of key frame g attends, in
addition to its own window, to the post-fc_new_4 key-frame rows Z of `n_support` other videos,
chosen by ring order over the global list of key frames:  (g+1 .. g+n_support) mod G.

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
    """Contiguous shard [lo, hi) of n_items for `rank` (videos are sliced contiguously per rank,
    as imagenet_vid_sequence.py:117-158 does for the reference's distributed test)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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


def gather_support(z_local, rows_per_key, n_support, group=None, async_stream=None):
    """Exchange + ring selection.  z_local: Split [V*P, D] of this rank's V key frames.
    Returns a list of V Splits [n_sel*P, D]: the support rows of each local key frame."""
    P = rows_per_key
    V = z_local.hi.shape[0] // P
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    pool = all_gather_rows(z_local, group)
    G = world * V
    out = []
    for v in range(V):
        idx = support_indices(rank * V + v, G, n_support)
        if not idx:
            out.append(Split(pool.hi[:0], pool.lo[:0]))
            continue
        out.append(Split(torch.cat([pool.hi[i * P:(i + 1) * P] for i in idx], 0),
                         torch.cat([pool.lo[i * P:(i + 1) * P] for i in idx], 0)))
    return out
