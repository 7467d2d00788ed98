"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: shard, all-gather, restore order."""

import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seaduck_b200 import parallel


def _worker(rank, world, n, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = parallel.shard_indices(n, world, rank, seed=3)
    per = -(-n // world)
    # each rank "advects" its own slice: here the result is a known function of the particle id
    state = {
        "lon": torch.tensor(idx * 0.5, dtype=torch.float64),
        "lat": torch.tensor(idx * -0.25, dtype=torch.float64),
        "dep": torch.tensor(-1.0 * idx, dtype=torch.float64),
        "face": torch.tensor(idx % 13, dtype=torch.int32),
        "iy": torch.tensor(idx % 90, dtype=torch.int32),
        "ix": torch.tensor((idx * 7) % 90, dtype=torch.int32),
        "izl": torch.tensor(idx % 50, dtype=torch.int32),
    }
    g = parallel.all_gather_snapshot(parallel.pack_snapshot(state, n_pad=per))
    # the overlapped gatherer takes the collective route on CPU and must deliver the same thing
    sg = parallel.SnapshotGather(parallel.SNAPSHOT_ROWS, per, torch.device("cpu"))
    assert sg.mode == "collective"
    sg.post(parallel.pack_snapshot(state, n_pad=per), slot=1)
    sg.wait()
    assert torch.equal(sg.result(1), g)
    out = parallel.unshard(g, n, world, seed=3)
    ids = np.arange(n)
    ok = (
        np.array_equal(out["lon"], ids * 0.5)
        and np.array_equal(out["dep"], -1.0 * ids)
        and np.array_equal(out["face"], ids % 13)
        and np.array_equal(out["ix"], (ids * 7) % 90)
    )
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_gather_restores_order():
    world, n = 2, 1001  # odd on purpose: the last slice is shorter (ragged)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, n, port, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_shards_are_disjoint_and_complete():
    n, world = 1000, 8
    parts = [parallel.shard_indices(n, world, r, seed=1) for r in range(world)]
    allidx = np.concatenate(parts)
    assert len(allidx) == n and len(set(allidx.tolist())) == n
    assert parallel.shard_indices(0, 2, 0).size == 0  # empty input
