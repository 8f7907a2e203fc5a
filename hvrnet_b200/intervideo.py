"""Inter-video support exchange: the one collective of the path.

At inference the shipped reference has no inter-video step (forward_test pools one window,
hrnmp_bbox_head.py:800-909); the stage exists in the training forward (:740-795).  BASELINE.json
configs 4-5 define it for inference: stage 4 of key frame g attends, in
addition to its own window, to the post-fc_new_4 key-frame rows Z of `n_support` other videos,
chosen by ring order over the global list of key frames:  (g+1 .. g+n_support) mod G  - or, with
video descriptors (next row N4, hnmb_rcnn.py:76-101), the n_support most similar other videos.

Sharding: each rank owns V key frames (global index g = rank*V + v), computes stages 1-3
locally, and ONE all-gather (split-bf16 pair = 4 bytes / element, bit-exact transport of what a
single GPU would hold) hands every rank the post-fc_new_4 key rows of all G key frames, their
proposal counts (frames may yield fewer than max_num proposals: the receivers mask the support
keys) and, for similarity selection, the video descriptors; stage 4 then runs locally
(window.inter_stage_a/b/c; the support rows are gathered out of the receive buffer as it lies by
hvr_support_index + hvr_gather_rows_split).  The functions below are device-agnostic
torch.distributed plumbing - layout of the send buffer, the collective, ring rule, sharding -
(NCCL on the GPUs; the gloo CPU tests drive the same code with world_size 2).
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


def pack_exchange(z_key, key_counts, desc=None):
    """Send buffer of the ONE all-gather: bf16 [2, rpr, D].  Block 0 (hi): the V*P post-fc_new_4 key rows' hi parts,
    then one carrier row holding the V per-key-frame proposal counts (int32 bits; the receivers mask the support
    keys with them), then - similarity selection only - r carrier rows per video holding its fp32 descriptor.
    Block 1 (lo): the lo parts of the key rows; its carrier rows are zero.  Split-bf16 travels bit-exactly."""
    VP, D = z_key.hi.shape
    V = key_counts.shape[0]
    assert V * 4 <= D * 2, 'counts carrier row too small'
    n_desc = 0
    carrier = None
    if desc is not None:
        carrier, r = _desc_rows(desc, D)
        n_desc = V * r
    rpr = VP + 1 + n_desc
    send = torch.zeros((2, rpr, D), dtype=torch.bfloat16, device=z_key.hi.device)
    send[0, :VP].copy_(z_key.hi)                # contiguous blocks: device-to-device memcpy nodes, no kernel
    send[1, :VP].copy_(z_key.lo)
    send[0, VP].view(torch.int32)[:V].copy_(key_counts)
    if carrier is not None:
        send[0, VP + 1:].copy_(carrier)
    return send, rpr, n_desc


def exchange(send, group=None, async_op=False):
    """The one collective of the path: all_gather_into_tensor of the send buffer -> recv bf16 [world, 2, rpr, D]
    (rank-major).  Returns (recv, work) - work is None unless async_op."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return send.unsqueeze(0), None
    world = dist.get_world_size(group)
    recv = torch.empty((world * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=async_op)   # dim-0 concatenation, rank-major
    return recv.view((world,) + tuple(send.shape)), (work if async_op else None)


def unpack_descriptors(recv, V, P, n_desc_rows, C):
    """Descriptor carrier rows of every rank -> fp32 [world*V, C]."""
    world, _, rpr, D = recv.shape
    rows = recv[:, 0, V * P + 1:V * P + 1 + n_desc_rows].contiguous().view(world * n_desc_rows, D)
    return _desc_from_rows(rows, world * V, n_desc_rows // V, C)


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
