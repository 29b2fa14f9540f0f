"""Two-GPU test (-m gpu, skipped on a single-GPU box): two processes, NCCL, sample buckets split between them. All three frame
combines — NCCL all-reduce then resolve, rank 0's resolve reading the peer's buckets over NVLink, and team mode (every rank resolves its
slab into rank 0's framebuffer, device-side hand-shakes, no host barrier per frame) — must reproduce the single-GPU frame bit-for-bit."""
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
    r.ipc_close()
    # team mode: every rank resolves its slab into rank 0's framebuffer; no host barrier between the frames. Three frames back to back
    # (reset, more samples each time) so that the "peers have read my buckets" and "rank 0 has copied the frame" hand-shakes are exercised;
    # the last one leaves through rank 0's copy stream.
    b2r_dist.open_team(r)
    frames = []
    for k, n in enumerate((SAMPLES, 2 * SAMPLES, SAMPLES)):
        r.ResetAccumulator(); r.Accumulate(n)
        f = np.zeros((H, W, 4), np.float32)
        if k == 2:
            import torch as _t
            pinned = _t.empty((H, W, 4), dtype=_t.float32, pin_memory=True).numpy()
            assert r.RenderTeam(out=pinned, use_async=True)
            if rank == 0:
                r.WaitFrame(); f = pinned.copy()
        else:
            assert r.RenderTeam(out=f)
        frames.append(f)
    # progressive: samples keep accumulating across resolves without a reset (the next Accumulate waits for the peers' reads on the device)
    r.Accumulate(SAMPLES); g = np.zeros((H, W, 4), np.float32); assert r.RenderTeam(out=g); frames.append(g)
    assert r.team_error() == 0
    if rank == 0:
        np.save(os.path.join(out_dir, "team.npy"), np.stack(frames))
    dist.barrier()
    r.team_close(); r.close()
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
    team = np.load(tmp_path / "team.npy")
    assert team[0].tobytes() == single.framebuffer.tobytes() and team[2].tobytes() == single.framebuffer.tobytes()
    single.ResetAccumulator(); single.Accumulate(2 * SAMPLES); assert single.Render()
    assert team[1].tobytes() == single.framebuffer.tobytes()
    single.ResetAccumulator(); single.Accumulate(SAMPLES); single.Accumulate(SAMPLES); assert single.Render()
    assert team[3].tobytes() == single.framebuffer.tobytes()
    single.close()
