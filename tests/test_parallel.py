"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: shard, all-gather, restore order."""

import os

import numpy as np
import pytest
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
    # the product packs on the device and ships with sdk_pack_ship; here the host restatement of the record and a
    # gloo all-gather stand in for them: what is under test is the slicing and its inverse
    mine = torch.from_numpy(parallel.pack_records_host({k: v.numpy() for k, v in state.items()}, n_pad=per).view(np.uint8).copy())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = parallel.unshard([p.numpy() for p in parts], n, world, seed=3)
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


def test_record_layout_round_trip():
    """32-byte record: three doubles and a cell word with 6 / 20 / 20 / 12 bits (include/seaduck_b200.h)."""
    rng = np.random.default_rng(1)
    n = 257
    state = {
        "lon": rng.uniform(-180, 180, n), "lat": rng.uniform(-90, 90, n), "dep": rng.uniform(-6000, 0, n),
        "face": rng.integers(0, 13, n), "iy": rng.integers(0, 17280, n), "ix": rng.integers(0, 1 << 20, n),
        "izl": rng.integers(0, 4096, n),
    }
    rec = parallel.pack_records_host(state, n_pad=260)
    assert rec.dtype.itemsize == 32 and rec.shape == (260,)
    back = parallel.unpack_records(rec.view(np.uint8))
    for k, v in state.items():
        np.testing.assert_array_equal(back[k][:n], v, err_msg=k)
    assert (back["lon"][n:] == 0).all() and (back["ix"][n:] == 0).all()
    # missing members read as zero (2-D runs have no depth / level, single-face grids no face)
    rec2 = parallel.unpack_records(parallel.pack_records_host(dict(state, dep=None, izl=None, face=None)))
    assert (rec2["dep"] == 0).all() and (rec2["izl"] == 0).all() and (rec2["face"] == 0).all()


def test_grids_that_do_not_fit_the_record_are_refused():
    from seaduck_b200.topology import Topology

    parallel.check_grid_fits_record(Topology.from_shape((13, 4320, 4320)), n_zl=90)
    parallel.check_grid_fits_record(Topology.from_shape((8640, 17280)))
    with pytest.raises(ValueError):
        parallel.check_grid_fits_record(Topology.from_shape((10, 10)), n_zl=5000)


def test_unshard_with_sender_orders():
    """A rank may ship its records in the (cell) order of its device state; its permutation undoes that."""
    n, world, seed = 103, 3, 9
    rng = np.random.default_rng(0)
    truth = {"lon": rng.uniform(-180, 180, n), "lat": rng.uniform(-90, 90, n), "dep": -rng.uniform(0, 5000, n),
             "face": rng.integers(0, 13, n), "iy": rng.integers(0, 90, n), "ix": rng.integers(0, 90, n), "izl": rng.integers(0, 50, n)}
    per = -(-n // world)
    blocks, orders = [], []
    for r in range(world):
        idx = parallel.shard_indices(n, world, r, seed)
        order = rng.permutation(len(idx)) if r != 1 else None  # rank 1 ships in slice order
        sel = idx if order is None else idx[order]
        blocks.append(parallel.pack_records_host({k: v[sel] for k, v in truth.items()}, n_pad=per))
        orders.append(order)
    out = parallel.unshard(blocks, n, world, seed, orders=orders)
    for k, v in truth.items():
        np.testing.assert_array_equal(out[k], v, err_msg=k)
