"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes, bucket ownership + the one combine collective.

The renderer itself needs a GPU; here each rank's "rendered" bucket sums come from the oracle restricted to the samples the
rank owns (oracle use is confined to tests). What is under test is b2r_dist: the partition is a partition, every sample has
exactly one owner, and all-reduce(sum) of owner-only buckets reproduces the single-process buckets bit-for-bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b2r_dist
import oracle_py
import scenes

W, H, K, MB, SAMPLES = 64, 48, 8, 6, 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]


def _rank_buckets(rank, world):
    sc = scenes.default_scene()
    o = oracle_py.Oracle(W, H, max_bounces=MB, K=K); o.set_scene(sc)
    for acc in range(1, SAMPLES + 1):  # Renderer::Accumulate pre-increments: first sample index is 1 (Q1)
        if b2r_dist.owns_sample(acc, rank, world, K):
            o.set_accumulations(acc - 1); o.accumulate(1, threads=1)
    return o.buckets()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.from_numpy(_rank_buckets(rank, world))
    owned = b2r_dist.owned_buckets(rank, world, K)
    others = [k for k in range(K) if k not in owned]
    assert not local[others].any() and local[owned].any()
    combined = b2r_dist.combine_buckets(local)
    assert torch.equal(local, torch.from_numpy(_rank_buckets(rank, world)))  # out-of-place: the local accumulator is untouched
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), combined.numpy())
    dist.barrier(); dist.destroy_process_group()


def test_partition_is_a_partition():
    for world in (1, 2, 4, 8):
        owners = [b2r_dist.owned_buckets(r, world, 8) for r in range(world)]
        assert sorted(sum(owners, [])) == list(range(8))
        for acc in range(1, 100):
            assert sum(b2r_dist.owns_sample(acc, r, world, 8) for r in range(world)) == 1
        kw = b2r_dist.shard_kwargs(world - 1, world, 8)
        assert kw == (dict(bucket_first=world - 1, bucket_stride=world) if world > 1 else dict(bucket_first=0, bucket_stride=0))
    with pytest.raises(ValueError):
        b2r_dist.owned_buckets(0, 3, 8)


def test_two_rank_combine_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    single = _rank_buckets(0, 1)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.tobytes() == single.tobytes()
