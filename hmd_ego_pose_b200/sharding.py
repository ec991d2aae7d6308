"""Frame sharding across the GPUs of one node: frames are independent units, so a batch is split
into contiguous shards, every rank runs its own libhmdpose handle on its shard, and there is NO collective on the
data path.  The only (optional) communication is a gather of the fixed-size detection tensors
(``[B_local, D, 75]`` floats, ~30 KB / frame) to rank 0 -- NCCL over NVLink on GPUs, gloo in the CPU tests.
The reference has no multi-GPU path to mirror (its DataParallel wrapper is dead code: pytorch-sandbox/train.py:123-127).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank: sizes differ by at most one frame, earlier ranks get the extra ones."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(max(num_frames, 0), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(num_frames: int, world_size: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_bounds(num_frames, world_size, r) for r in range(world_size))]


def pack_detections(det: Sequence[torch.Tensor]) -> torch.Tensor:
    """[boxes (b,D,4), scores (b,D), labels (b,D), rotation (b,D,3), translation (b,D,3), hand (b,D,63), idx (b,D)]
    -> one float32 tensor (b, D, 76) so that the gather is a single collective."""
    boxes, scores, labels, rot, trans, hand, idx = det[:7]
    return torch.cat([boxes.float(), scores.float()[..., None], labels.float()[..., None], rot.float(), trans.float(),
                      hand.float(), idx.float()[..., None]], dim=-1)


def unpack_detections(packed: torch.Tensor) -> List[torch.Tensor]:
    return [packed[..., 0:4], packed[..., 4], packed[..., 5].to(torch.int32), packed[..., 6:9], packed[..., 9:12],
            packed[..., 12:75], packed[..., 75].to(torch.int32)]


def gather_detections(local: torch.Tensor, num_frames: int, dst: int = 0,
                      group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Gather the per-rank packed detections (ragged in dim 0) to ``dst`` in frame order.  Returns the
    (num_frames, D, 76) tensor on ``dst`` and None elsewhere.  Uses all_gather on equal-size padded shards so it
    works on both NCCL and gloo."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(num_frames, world)
    pad_to = max(sizes) if sizes else 0
    buf = local.new_zeros((pad_to,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)
