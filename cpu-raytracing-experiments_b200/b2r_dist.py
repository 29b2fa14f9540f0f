"""Multi-GPU plumbing: one process per GPU, sample buckets partitioned over the ranks.

Partition (BASELINE north_star): scene + BVH replicated; rank g of G owns the median-of-means buckets
{b : b % G == g} and renders exactly the sample indices acc with acc % K in that set (RNG streams are a pure function of
(acc, pixel, bounce), Renderer.hpp:117,255,362, so no rank needs another's state). Nothing is exchanged while rendering.

The frame is put together by TEAM MODE (open_team + Renderer.RenderTeam, b2r_team_* in include/b2r.h): every rank resolves its slab of
tiles, pulling each bucket's sums from its owner's HBM over NVLink, into rank 0's framebuffer; the ranks hand frames over through
release/acquire flags in peer-mapped memory, so there is no host barrier and no NCCL call per frame. torch.distributed / NCCL carries
the one-time exchange of CUDA IPC handles (and, in bench.py, the barriers around the timed region and the max-over-ranks of the time).
Two older combines are kept and tested: open_peers + RenderPeers (rank 0 resolves everything from peer memory between two barriers the
caller provides) and combine_buckets (out-of-place NCCL all-reduce of the owner-only [K][3][npix] bucket arrays — every bucket has one
owner and is zero elsewhere, so adding is exact — then Render(dev_buckets=...); gloo on CPU for the tests). All three give a frame that
is bit-identical to a single-GPU render of the same samples.
"""
import numpy as np


def owned_buckets(rank, world, K):
    if K % world:
        raise ValueError(f"bucket count {K} must be a multiple of the world size {world}")
    return [b for b in range(K) if b % world == rank]


def shard_kwargs(rank, world, K):
    """bucket_first / bucket_stride for b2r_config."""
    owned_buckets(rank, world, K)
    return dict(bucket_first=rank, bucket_stride=world) if world > 1 else dict(bucket_first=0, bucket_stride=0)


def owns_sample(acc, rank, world, K):
    return world <= 1 or (acc % K) % world == rank


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def buckets_tensor(renderer, device):
    import torch
    ptr, nbytes = renderer.device_buckets()
    return torch.as_tensor(_DevArray(ptr, nbytes // 4), device=device)


def combine_buckets(local_buckets, group=None):
    """All ranks end with every bucket: out-of-place all-reduce(sum) of the owner-only bucket arrays (torch tensor, any device)."""
    import torch.distributed as dist
    out = local_buckets.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    if out.is_cuda:
        # The Renderer resolves on the library's own stream (or the one given at construction), which is not ordered against
        # torch's current stream: the reduced buckets must have landed before Render(dev_buckets=...) is enqueued.
        import torch
        torch.cuda.current_stream(out.device).synchronize()
    return out


def open_peers(renderer, group=None):
    """Fused alternative to combine_buckets: exchange CUDA IPC handles of the bucket arrays once (all_gather_object) and map the
    peers' arrays, so Renderer.RenderPeers() pulls each bucket from its owner over NVLink inside the resolve kernel. Per frame the
    only cross-rank operation left is a barrier that orders the resolve after every rank's last bounce."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    handles = [None] * world
    mine = renderer.ipc_export_buckets()
    if world > 1:
        dist.all_gather_object(handles, mine, group=group)
    else:
        handles = [mine]
    renderer.ipc_open_peers(handles, rank)
    return world


def open_team(renderer, group=None):
    """Team mode (b2r_team_*): exchange the three CUDA IPC handles of every rank once (bucket array, framebuffer, hand-shake flags),
    map the peers', and barrier ONCE so that nobody signals into a block that is not mapped yet. After this Renderer.RenderTeam() needs no
    host-side synchronisation between ranks: the ranks hand frames over through release/acquire flags in peer memory."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = renderer.team_export()
    handles = [None] * world
    if world > 1:
        dist.all_gather_object(handles, mine, group=group)
    else:
        handles = [mine]
    renderer.team_open(handles, rank)
    if world > 1:
        dist.barrier(group=group)
    return world
