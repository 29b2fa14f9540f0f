"""Two-GPU test (-m gpu, skipped on a single-GPU box): two processes, NCCL, sample buckets split between them. Both frame
combines — NCCL all-reduce then resolve, and the fused resolve that reads the peer's buckets over NVLink — must reproduce the
single-GPU frame bit-for-bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
W, H, K, MB, SAMPLES = 256, 144, 8, 8, 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import b2r, b2r_dist, scenes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    r = b2r.Renderer(scenes.default_scene(), W, H, max_bounces=MB, buckets=K, device=rank, **b2r_dist.shard_kwargs(rank, world, K))
    r.Accumulate(SAMPLES); r.sync()
    combined = b2r_dist.combine_buckets(b2r_dist.buckets_tensor(r, torch.device(f"cuda:{rank}")))
    a = np.zeros((H, W, 4), np.float32); assert r.Render(out=a, dev_buckets=combined.data_ptr())
    b2r_dist.open_peers(r)
    dist.barrier()
    b = np.zeros((H, W, 4), np.float32); assert r.RenderPeers(out=b)
    dist.barrier()
    np.save(os.path.join(out_dir, f"nccl{rank}.npy"), a); np.save(os.path.join(out_dir, f"p2p{rank}.npy"), b)
    r.ipc_close(); r.close()
    dist.destroy_process_group()


def test_two_gpu_frame_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import b2r, scenes
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    single = b2r.Renderer(scenes.default_scene(), W, H, max_bounces=MB, buckets=K); single.Accumulate(SAMPLES); assert single.Render()
    for name in ("nccl0", "nccl1", "p2p0", "p2p1"):
        assert np.load(tmp_path / f"{name}.npy").tobytes() == single.framebuffer.tobytes(), name
    single.close()
