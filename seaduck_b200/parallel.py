"""Particle sharding over the GPUs of one node.

Particles never interact, fields are read-only during a leg: every rank holds a replica of the
velocity field and a disjoint slice of the particles.  Work per particle (number of crossings)
depends on where it is, so slices are taken from a fixed random permutation rather than
contiguously.  The only collective is an all-gather of the output snapshot of a stop; the original
particle order is restored with the inverse permutation.
"""

from __future__ import annotations

import numpy as np

SNAPSHOT_FIELDS = ("lon", "lat", "dep", "face", "iy", "ix", "izl")


def permutation(n, seed=0):
    return np.random.default_rng(seed).permutation(n)


def shard_indices(n, world, rank, seed=0):
    """Indices (into the user's particle arrays) owned by ``rank``: equal slices of a permutation."""
    perm = permutation(n, seed)
    per = -(-n // world)
    return perm[rank * per : min((rank + 1) * per, n)]


def pack_snapshot(state, n_pad=None):
    """One contiguous float64 tensor [7, n] from a particle state dict (ints are exact in fp64)."""
    import torch

    n = state["lon"].shape[0]
    n_pad = n if n_pad is None else n_pad
    out = torch.zeros((len(SNAPSHOT_FIELDS), n_pad), dtype=torch.float64, device=state["lon"].device)
    for k, name in enumerate(SNAPSHOT_FIELDS):
        v = state.get(name)
        if v is not None:
            out[k, :n] = v.to(torch.float64)
    return out


def all_gather_snapshot(packed, group=None):
    """All-gather the packed snapshots of every rank -> [world, 7, n_pad]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rows, cols = packed.shape
    out = torch.empty((world * rows, cols), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return out.view(world, rows, cols)


def unshard(gathered, n, world, seed=0):
    """Undo :func:`shard_indices`: dict of arrays of length ``n`` in the user's particle order."""
    perm = permutation(n, seed)
    per = -(-n // world)
    g = gathered.cpu().numpy() if hasattr(gathered, "cpu") else np.asarray(gathered)
    out = {}
    for k, name in enumerate(SNAPSHOT_FIELDS):
        full = np.empty(n, np.float64)
        for r in range(world):
            idx = perm[r * per : min((r + 1) * per, n)]
            full[idx] = g[r, k, : len(idx)]
        out[name] = full if name in ("lon", "lat", "dep") else full.astype(np.int64)
    return out
