"""N>1 path on CPU: world_size-2 (and 3) gloo runs of the frame sharding + gather-to-rank-0
logic that bench.py uses with NCCL on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blinky_b200.sharding import frames_for_rank, gather_frames

    mine = frames_for_rank(total, rank, world)
    H, W = 6, 8
    # each "finished frame" is filled with its global frame id, as a stand-in for the warp output
    local = torch.stack([torch.full((H, W), f, dtype=torch.uint8) for f in mine]) if len(mine) else torch.empty((0, H, W), dtype=torch.uint8)
    got = gather_frames(local, rank, world, total_frames=total)
    if rank == 0:
        ok = got.shape == (total, H, W) and all(bool((got[f] == f).all()) for f in range(total))
        ret.put(ok)
    else:
        ret.put(got is None)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 8), (2, 5), (3, 7)])
def test_gather_to_rank0_gloo(world, total):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = [ret.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_shard_range_c_entry_rejects_bad_arguments(bb):
    import ctypes

    lib = bb.load_library()
    f, c = ctypes.c_int(), ctypes.c_int()
    assert lib.blinky_shard_range(8, 0, 2, ctypes.byref(f), ctypes.byref(c)) == 0 and (f.value, c.value) == (0, 4)
    assert lib.blinky_shard_range(8, 2, 2, ctypes.byref(f), ctypes.byref(c)) == bb.E_INVALID
    assert lib.blinky_shard_range(8, 0, 0, ctypes.byref(f), ctypes.byref(c)) == bb.E_INVALID
    assert lib.blinky_shard_range(-1, 0, 1, ctypes.byref(f), ctypes.byref(c)) == bb.E_INVALID
    # the collective entries need a GPU context and an initialised group
    with bb.Fisheye(device=None) as fe:
        with pytest.raises(bb.BlinkyError) as ei:
            fe.shard_init(0, 1, b"\0" * 128)
        assert ei.value.code == bb.E_NODEVICE


def test_frames_for_rank_partitions_exactly():
    from blinky_b200.sharding import frames_for_rank

    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 8, 64, 65):
            seen = []
            for r in range(world):
                fr = frames_for_rank(total, r, world)
                seen.extend(fr)
                assert len(fr) in (total // world, total // world + 1)
            assert seen == list(range(total))
