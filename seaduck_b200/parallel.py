"""Particles sharded over the GPUs of one node.

Particles never interact and the fields are read-only during a leg, so every rank (one process per
GPU) holds a replica of the velocity field and a disjoint slice of the particles.  Work per particle
(its number of crossings) depends on where it is, so slices are taken from a fixed random permutation
rather than contiguously.  The only exchange is that of the *output*: what ``to_list_of_time`` returns
per stop (reference ``lagrangian.py:804``) is one 32-byte record per particle (``sdk_snapshot_record``).

Two ways to assemble it:

``gather="all"``    every rank ends up with every record.  ``sdk_pack_ship`` packs the records of this rank
                    and stores them straight into the receive buffers of all ranks (peer mappings of symmetric
                    memory over NVLink, one ``multimem.st`` per 16 bytes where the NVSwitch can replicate) from
                    inside one kernel; flags in the same memory say when a block has landed and when a slot
                    may be overwritten.  No collective library call on the data path.
``gather="owner"``  every rank keeps the records of its own slice in HBM; :meth:`ShardedSnapshot.assemble`
                    collects them on one rank when somebody asks.  For runs whose output volume per stop
                    (32 B x 10^8 particles) would otherwise dwarf the compute of a stop.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SNAPSHOT_FIELDS = ("lon", "lat", "dep", "face", "iy", "ix", "izl")
RECORD_BYTES = 32
_HEADER_BYTES = 4096  # signal words at the start of the symmetric buffer, records behind them


def permutation(n, seed=0):
    return np.random.default_rng(seed).permutation(n)


def shard_indices(n, world, rank, seed=0):
    """Indices (into the user's particle arrays) owned by ``rank``: equal slices of a permutation."""
    perm = permutation(n, seed)
    per = -(-n // world)
    return perm[rank * per : min((rank + 1) * per, n)]


def check_grid_fits_record(tp, n_zl=0):
    """The cell word of a record has 6 / 20 / 20 / 12 bits for face / iy / ix / izl."""
    nface = tp.h_shape[0] if tp.has_face else 1
    ny, nx = tp.h_shape[-2:]
    limits = dict(face=nface, iy=ny, ix=nx, izl=n_zl + 1)
    for name, size in limits.items():
        if size > (1 << _lib.CELL_BITS[name]):
            raise ValueError(f"{name} extent {size} does not fit the {_lib.CELL_BITS[name]} bits of the output record")


# ---- records on the host -----------------------------------------------------------------------------------
def pack_records_host(state, n_pad=None):
    """numpy restatement of ``sdk_pack_records`` for host arrays (used by the multi-process CPU tests of the
    sharding logic; the product packs on the device)."""
    n = len(state["lon"])
    n_pad = n if n_pad is None else n_pad
    rec = np.zeros(n_pad, np.dtype(_lib.SNAPSHOT_RECORD))
    for name in ("lon", "lat", "dep"):
        if state.get(name) is not None:
            rec[name][:n] = np.asarray(state[name], np.float64)
    word = np.zeros(n, np.uint64)
    for name in ("face", "iy", "ix", "izl"):
        if state.get(name) is not None:
            v = np.asarray(state[name]).astype(np.uint64) & np.uint64((1 << _lib.CELL_BITS[name]) - 1)
            word |= v << np.uint64(_lib.CELL_SHIFT[name])
    rec["cell"][:n] = word
    return rec


def unpack_records(rec):
    """Structured record array (or raw bytes / uint8 tensor of them) -> dict of numpy arrays."""
    if hasattr(rec, "cpu"):
        rec = rec.cpu().numpy()
    rec = np.ascontiguousarray(rec)
    if rec.dtype != np.dtype(_lib.SNAPSHOT_RECORD):
        rec = rec.reshape(-1).view(np.dtype(_lib.SNAPSHOT_RECORD))
    out = {"lon": rec["lon"], "lat": rec["lat"], "dep": rec["dep"]}
    word = rec["cell"]
    for name in ("face", "iy", "ix", "izl"):
        out[name] = ((word >> np.uint64(_lib.CELL_SHIFT[name])) & np.uint64((1 << _lib.CELL_BITS[name]) - 1)).astype(np.int64)
    return out


def unshard(blocks, n, world, seed=0, orders=None):
    """Undo :func:`shard_indices`: ``blocks`` = per-rank record arrays ([world, n_pad] or a list) -> dict of arrays
    of length ``n`` in the user's particle order.  ``orders[r]`` (optional): record k of rank r's block belongs to
    particle ``orders[r][k]`` of that rank's slice (a rank that ships its records in the cell order of its device state)."""
    perm = permutation(n, seed)
    per = -(-n // world)
    out = {name: np.empty(n, np.float64 if name in ("lon", "lat", "dep") else np.int64) for name in SNAPSHOT_FIELDS}
    for r in range(world):
        idx = perm[r * per : min((r + 1) * per, n)]
        if orders is not None and orders[r] is not None:
            idx = idx[np.asarray(orders[r])[: len(idx)]]
        block = unpack_records(blocks[r])
        for name in SNAPSHOT_FIELDS:
            out[name][idx] = block[name][: len(idx)]
    return out


# ---- records on the device -----------------------------------------------------------------------------------
def pack_records(cparticles, n, out=None, out_index=None, device=None):
    """``sdk_pack_records`` into a uint8 tensor [n, 32] (allocated when ``out`` is None)."""
    import torch

    if out is None:
        out = torch.empty((n, RECORD_BYTES), dtype=torch.uint8, device=device)
    _lib.check(
        _lib.lib().sdk_pack_records(C.byref(cparticles), None if out_index is None else out_index.data_ptr(), out.data_ptr(),
                                    _lib.current_stream()),
        "sdk_pack_records",
    )
    return out


class RecordExchange:
    """Receive buffers and signal words of the record exchange in symmetric memory, and the calls that use them.

    ``n_pad`` records per rank and slot; ``slots`` receive slots (a slot is reused once every rank has released it).
    Needs an initialised NCCL process group (for the rendezvous of the mappings only) and one GPU per rank.
    """

    def __init__(self, n_pad, device, group=None, slots=2, multicast=True):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"at most {_lib.MAX_RANKS} ranks")
        self.n_pad, self.slots, self.device = int(n_pad), int(slots), torch.device(device)
        nbytes = _HEADER_BYTES + self.slots * self.world * self.n_pad * RECORD_BYTES
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        x = _lib.Exchange()
        x.world, x.rank, x.n_slots, x.n_pad = self.world, self.rank, self.slots, self.n_pad
        ptrs = list(self.handle.buffer_ptrs)
        for r in range(self.world):
            x.signal[r] = ptrs[r]
            x.recv[r] = ptrs[r] + _HEADER_BYTES
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0) if multicast else 0
        x.recv_multicast = (mc + _HEADER_BYTES) if mc else None
        x.ticket = self.ticket.data_ptr()
        self.x = x
        self.mode = "multimem.st through the NVSwitch (NVLS)" if mc else "stores through peer mappings"
        self.step = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # every buffer is zeroed before anybody writes into it

    def ship(self, cparticles, out_index=None):
        """Pack this rank's records and store them into every rank's receive buffer (one kernel on the
        current stream); ``out_index`` (int32 device tensor) undoes a cell sort.  Returns the step number to wait for."""
        self.step += 1
        _lib.check(
            _lib.lib().sdk_pack_ship(C.byref(self.x), C.byref(cparticles), None if out_index is None else out_index.data_ptr(),
                                     self.step, _lib.current_stream()),
            "sdk_pack_ship",
        )
        return self.step

    def next_step(self):
        """Step number for a kernel that ships on its own (``sdk_advect`` with ``sdk_advect_params.ship``)."""
        self.step += 1
        return self.step

    def wait(self, step, release=False):
        """The current stream waits until every rank's records of ``step`` are here; ``release`` also frees the slot."""
        _lib.check(_lib.lib().sdk_exchange_wait(C.byref(self.x), step, 1, int(release), _lib.current_stream()), "sdk_exchange_wait")

    def release(self, step):
        _lib.check(_lib.lib().sdk_exchange_wait(C.byref(self.x), step, 0, 1, _lib.current_stream()), "sdk_exchange_wait")

    def slot_view(self, step):
        """uint8 view [world, n_pad, 32] of the slot that holds ``step``."""
        slot = step % self.slots
        per = self.world * self.n_pad * RECORD_BYTES
        return self.buf[_HEADER_BYTES + slot * per : _HEADER_BYTES + (slot + 1) * per].view(self.world, self.n_pad, RECORD_BYTES)

    def check(self):
        """Raise if a wait timed out on the device (a rank died or fell far behind)."""
        words = self.buf[: _lib.SIGNAL_WORDS * 8].view(self.buf.dtype).cpu().numpy().view(np.uint64)
        if words[_lib.SIGNAL_ERROR]:
            raise _lib.SdkError("record exchange: a wait for another rank timed out on the device")


class ShardedSnapshot:
    """Output of one stop of a sharded run.  ``gather="all"``: ``blocks`` is a device tensor [world, n_pad, 32] with
    every rank's records; ``gather="owner"``: [1, n_pad, 32] with this rank's own."""

    def __init__(self, owner, t, blocks, everyone):
        self.owner, self.t, self.blocks, self.everyone = owner, float(t), blocks, everyone

    def assemble(self, dst=0):
        """dict of numpy arrays of all ``N`` particles in the caller's order (on every rank for ``gather="all"``,
        on rank ``dst`` -- ``None`` elsewhere -- for ``gather="owner"``)."""
        import torch
        import torch.distributed as dist

        o = self.owner
        if self.everyone or o.world == 1:
            return unshard(self.blocks.cpu().numpy(), o.n_total, o.world, o.seed, orders=o.orders if self.everyone else None)
        mine = self.blocks[0].contiguous()
        parts = [torch.empty_like(mine) for _ in range(o.world)] if o.rank == dst else None
        dist.gather(mine, parts, dst=dst, group=o.group)
        if o.rank != dst:
            return None
        return unshard([p.cpu().numpy() for p in parts], o.n_total, o.world, o.seed)


class ShardedParticle:
    """``Particle`` over all the GPUs of a node: every rank calls this with the SAME full coordinate arrays and its
    own ``OceData`` replica; each advects its slice.  ``to_list_of_time`` returns, per stop, a
    :class:`ShardedSnapshot` whose ``assemble()`` gives positions and cells of ALL particles in the caller's order
    (reference ``lagrangian.py:804`` returns one full snapshot per stop).  Other keyword arguments go to ``Particle``.
    """

    def __init__(self, x, y, z=None, t=None, data=None, group=None, gather="all", seed=0, slots=2, **kw):
        import torch.distributed as dist

        from .lagrangian import Particle

        if gather not in ("all", "owner"):
            raise ValueError("gather must be 'all' or 'owner'")
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.gather, self.seed = gather, seed
        x = np.asarray(x, float)
        self.n_total = len(x)
        self.idx = shard_indices(self.n_total, self.world, self.rank, seed)
        self.n_pad = -(-self.n_total // self.world)
        pick = lambda a: None if a is None else (np.asarray(a, float)[self.idx] if np.ndim(a) else a)  # noqa: E731
        self.local = Particle(x=x[self.idx], y=pick(y), z=pick(z), t=pick(t), data=data, **kw)
        self.local.lazy = True
        check_grid_fits_record(data.tp, len(data.Zl) if data.readiness["Zl"] else 0)
        self.exchange = None
        self.orders = None
        if self.world > 1 and gather == "all":
            self.exchange = RecordExchange(self.n_pad, data._torch_device(), group=group, slots=slots)
            # The advection kernel ships the records itself, in the (cell) order of its device state: stores of lanes
            # that finish together are neighbours in the receive buffer.  Which particle a record belongs to is told
            # once, here: every rank's permutation goes to every rank.
            self.local.ship_to(self.exchange, order="state")
            self.orders = self._exchange_orders(group)

    def _exchange_orders(self, group):
        """numpy permutations of all ranks (None for a rank whose state is in slice order)."""
        import torch
        import torch.distributed as dist

        pt = self.local
        dev = pt.ocedata._torch_device()
        mine = torch.full((self.n_pad,), -1, dtype=torch.int32, device=dev)
        if pt._perm is not None:
            mine[: pt.N] = pt._perm
        everyone = torch.empty((self.world, self.n_pad), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        host = everyone.cpu().numpy()
        return [None if row[0] < 0 else row for row in host]

    def to_list_of_time(self, normal_stops, update_stops="default", return_in_between=True):
        import torch

        pt = self.local
        dev = pt.ocedata._torch_device()
        side = torch.cuda.Stream(dev)
        snaps = []
        # one block for the records of all stops (at most one per stop in the two lists)
        od = pt.ocedata
        if isinstance(update_stops, str):
            n_update = len(od.time_midp) if (od.readiness["time"] and "time" in od[pt.uname].dims) else 0
        else:
            n_update = len(list(update_stops))
        n_max = len(list(normal_stops)) + n_update
        blocks_per = self.world if self.exchange is not None else 1
        pool = torch.zeros((n_max, blocks_per, self.n_pad, RECORD_BYTES), dtype=torch.uint8, device=dev)

        def on_stop(t_stop):
            # runs right after the leg to t_stop was queued
            cp = pt._cparticles()
            if len(snaps) >= pool.shape[0]:
                raise RuntimeError("more stops than expected: pass the update stops explicitly")
            if self.exchange is None:
                rec = pool[len(snaps)]
                pack_records(cp, pt.N, out=rec[0], out_index=pt._out_index())
                snaps.append(ShardedSnapshot(self, t_stop, rec, everyone=False))
                return
            step = pt.last_ship_step  # shipped by the leg that just ended (sdk_advect with the fused exchange)
            done = torch.cuda.Event()
            done.record()
            side.wait_event(done)
            with torch.cuda.stream(side):
                # off the critical path: wait for everybody's block, keep a copy, free the slot
                self.exchange.wait(step)
                blocks = pool[len(snaps)]
                blocks.copy_(self.exchange.slot_view(step))
                self.exchange.release(step)
            snaps.append(ShardedSnapshot(self, t_stop, blocks, everyone=True))

        stops, _ = pt.to_list_of_time(normal_stops, update_stops, return_in_between, _on_stop=on_stop)
        torch.cuda.current_stream(dev).wait_stream(side)
        pt.sync()
        if self.exchange is not None:
            self.exchange.check()
        return stops, snaps
