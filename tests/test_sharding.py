"""N > 1 host logic on the CPU: frame sharding + the optional result gather, world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hmd_ego_pose_b200 import sharding


def test_shard_bounds_cover_frames_exactly():
    for n in (0, 1, 7, 16, 64, 257):
        for w in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_pack_roundtrip():
    g = torch.Generator().manual_seed(0)
    b, d = 3, 100
    det = [torch.rand(b, d, 4, generator=g), torch.rand(b, d, generator=g),
           torch.randint(-1, 3, (b, d), generator=g, dtype=torch.int32), torch.rand(b, d, 3, generator=g),
           torch.rand(b, d, 3, generator=g), torch.rand(b, d, 63, generator=g),
           torch.randint(-1, 12276, (b, d), generator=g, dtype=torch.int32)]
    back = sharding.unpack_detections(sharding.pack_detections(det))
    for a, c in zip(det, back):
        assert torch.equal(a, c.to(a.dtype))


def _worker(rank, world, port, frames):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(frames, world, rank)
        # each rank "detects" on its shard: the payload encodes the global frame number
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].expand(hi - lo, 100, 76).contiguous()
        out = sharding.gather_detections(local, frames, dst=0)
        if rank == 0:
            assert out.shape == (frames, 100, 76)
            assert torch.equal(out[:, 0, 0], torch.arange(frames, dtype=torch.float32))
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frames", [16, 5])
def test_gather_world_size_2_gloo(frames):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, frames), nprocs=2, join=True)
