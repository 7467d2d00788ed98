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


SNAPSHOT_ROWS = 4  # lon, lat, dep, one word of indices: 32 bytes per particle
_SHIFT = {"face": 34, "iy": 21, "ix": 8, "izl": 0}
_BITS = {"face": 6, "iy": 13, "ix": 13, "izl": 8}


def pack_snapshot(state, n_pad=None):
    """One contiguous float64 tensor [4, n_pad] from a particle state dict: positions as they are, the
    four indices bit-packed into one 64-bit word (face 6 bits, iy 13, ix 13, izl 8) -- one
    ``sdk_pack_snapshot`` launch on the GPU, plain torch ops for host tensors."""
    import torch

    n = state["lon"].shape[0]
    n_pad = n if n_pad is None else n_pad
    dev = state["lon"].device
    out = torch.empty((SNAPSHOT_ROWS, n_pad), dtype=torch.float64, device=dev)
    if dev.type == "cuda":
        from . import _lib

        ptr = lambda k: None if state.get(k) is None else state[k].data_ptr()  # noqa: E731
        _lib.check(
            _lib.lib().sdk_pack_snapshot(n, n_pad, ptr("lon"), ptr("lat"), ptr("dep"), ptr("face"), ptr("iy"), ptr("ix"),
                                         ptr("izl"), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
            "sdk_pack_snapshot",
        )
        return out
    out.zero_()
    for k, name in enumerate(("lon", "lat", "dep")):
        if state.get(name) is not None:
            out[k, :n] = state[name].to(torch.float64)
    word = torch.zeros(n_pad, dtype=torch.int64)
    for name in ("face", "iy", "ix", "izl"):
        if state.get(name) is not None:
            word[:n] |= (state[name].to(torch.int64) & ((1 << _BITS[name]) - 1)) << _SHIFT[name]
    out[3] = word.view(torch.float64)
    return out


def unpack_snapshot(packed):
    """Inverse of :func:`pack_snapshot` for one rank's [4, n] block -> dict of numpy arrays."""
    g = packed.cpu().numpy() if hasattr(packed, "cpu") else np.asarray(packed)
    g = np.ascontiguousarray(g)
    out = {"lon": g[0], "lat": g[1], "dep": g[2]}
    word = g[3].view(np.int64)
    for name in ("face", "iy", "ix", "izl"):
        out[name] = ((word >> _SHIFT[name]) & ((1 << _BITS[name]) - 1)).astype(np.int32)
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
    out = {name: np.empty(n, np.float64 if name in ("lon", "lat", "dep") else np.int64) for name in SNAPSHOT_FIELDS}
    for r in range(world):
        idx = perm[r * per : min((r + 1) * per, n)]
        block = unpack_snapshot(g[r])
        for name in SNAPSHOT_FIELDS:
            out[name][idx] = block[name][: len(idx)]
    return out


class SnapshotGather:
    """All-gather of the packed output snapshots, overlapped with the next leg.

    The only exchange of the path ships *finished* positions, so nothing has to wait
    for it.  ``post`` starts the all-gather of one snapshot and returns; the advection kernel of the
    next leg is launched right behind it.  Because that kernel is persistent and would otherwise
    own every SM, callers launch it with ``reserve_sms`` (``Particle.tuning``) so that the NCCL
    kernel finds free SMs and the two really run side by side.  Receive buffers are double-buffered
    (``slot``); ``wait()`` drains everything and meets the other ranks.

    ``use_peer=True`` selects copy-engine writes into the other ranks' receive buffers through CUDA
    IPC mappings (``sdk_peer_copy_async``) instead of NCCL.  On the round-1 boxes those copies ran at
    25 GB/s (staged, not NVLink) against ~215 GB/s for NCCL, so it is off by default.
    """

    def __init__(self, rows, cols, device, group=None, slots=2, use_peer=False):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows, self.cols, self.device = rows, cols, torch.device(device)
        self.recv = [torch.zeros((self.world, rows, cols), dtype=torch.float64, device=self.device) for _ in range(slots)]
        self.work = [None] * slots
        self.keep = [None] * slots
        self.peers = None
        self.stream = None
        self.mode = "local"
        if self.world > 1:
            self.mode = "collective"
            if use_peer and self.device.type == "cuda":
                try:
                    self._open_peers()
                    self.mode = "peer"
                except Exception as exc:  # noqa: BLE001  (no P2P / IPC: NCCL does the job)
                    self.peer_error = repr(exc)
                flag = torch.tensor([1 if self.mode == "peer" else 0], device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # every rank takes the same route
                if int(flag.item()) == 0:
                    self.mode, self.peers = "collective", None
            if self.mode == "peer":
                self.stream = torch.cuda.Stream(self.device)

    def _open_peers(self):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        from . import _lib

        me = self.device.index if self.device.index is not None else torch.cuda.current_device()
        mine = [reduce_tensor(buf) for buf in self.recv]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (me, mine), group=self.group)
        peers = []
        for r, (dev_r, handles) in enumerate(everyone):
            if r == self.rank:
                peers.append(self.recv)
                continue
            _lib.check(_lib.lib().sdk_peer_enable(dev_r), "sdk_peer_enable")
            peers.append([fn(*args) for fn, args in handles])
        self.peers = peers
        self.me = me

    def post(self, packed, slot=0):
        """Start shipping ``packed`` ([rows, cols] float64) into slot ``slot`` of every rank."""
        import torch
        import torch.distributed as dist

        if self.world == 1:
            self.recv[slot][0].copy_(packed)
            return
        if self.mode == "collective":
            if self.work[slot] is not None:
                self.work[slot].wait()  # the gather that used this slot two posts ago
            out = self.recv[slot].view(self.world * self.rows, self.cols)
            packed = packed.contiguous()
            self.keep[slot] = packed
            self.work[slot] = dist.all_gather_into_tensor(out, packed, group=self.group, async_op=True)
            return
        from . import _lib

        L = _lib.lib()
        ready = torch.cuda.Event()
        ready.record()
        self.stream.wait_event(ready)
        packed.record_stream(self.stream)
        nbytes = packed.numel() * packed.element_size()
        for k in range(self.world):
            r = (self.rank + k) % self.world  # own copy first, then round the ring
            dst = self.peers[r][slot][self.rank]
            _lib.check(
                L.sdk_peer_copy_async(dst.data_ptr(), dst.device.index, packed.data_ptr(), self.me, nbytes, self.stream.cuda_stream),
                "sdk_peer_copy_async",
            )

    def wait(self):
        """Everything posted so far has landed everywhere."""
        import torch
        import torch.distributed as dist

        for k, w in enumerate(self.work):
            if w is not None:
                w.wait()
                self.work[k] = None
        if self.device.type == "cuda":
            if self.stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
            torch.cuda.current_stream(self.device).synchronize()
        self.keep = [None] * len(self.keep)
        if self.world > 1:
            dist.barrier(group=self.group)

    def result(self, slot=0):
        return self.recv[slot]
