"""Frame-batch sharding across GPUs (one process per GPU).

The warp of one frame is independent of every other frame (the lensmap is replicated —
each rank rebuilds it deterministically from the same scripts), so a batch shards with NO
collective on the data path.  The only exchange is the reference topology's final step:
finished frames travel to rank 0 (the one display).  The product path for that is the C ABI's
blinky_shard_warp_gather (include/blinky_b200.h: NCCL send/recv or NVLink peer memory, overlapped with
the warp); gather_frames below is the same exchange with torch.distributed point-to-point ops — used
by the CPU tests (gloo) and as the cross-check of the C path in bench.py.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def frames_for_rank(total_frames: int, rank: int, world: int) -> range:
    """contiguous block of global frame ids owned by `rank` (sizes differ by at most one) — the C ABI's
    blinky_shard_range, the one partition rule every gather path (NCCL, peer copy, peer store, gloo) uses"""
    from . import shard_range

    return shard_range(total_frames, rank, world)


def gather_frames(local: torch.Tensor, rank: int, world: int, total_frames: Optional[int] = None,
                  dst: int = 0) -> Optional[torch.Tensor]:
    """Collects every rank's finished frames ([n_r, H, W] uint8) on `dst` in global frame
    order; returns the [total, H, W] tensor there and None elsewhere."""
    if world == 1:
        return local
    if total_frames is None:
        total_frames = local.shape[0] * world
    if rank == dst:
        out = torch.empty((total_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        ops = []
        for r in range(world):
            fr = frames_for_rank(total_frames, r, world)
            if len(fr) == 0:
                continue
            if r == dst:
                out[fr.start:fr.stop].copy_(local)
            else:
                ops.append(dist.P2POp(dist.irecv, out[fr.start:fr.stop], r))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if local.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
            req.wait()
    return None
