"""Crossing-log post-processing for the Lagrangian budget, on the device.

The reference (``seaduck/lagrangian_budget.py``) dumps the crossing log of a ``Particle(save_raw=True)``
to zarr (``particle2xarray`` :297, ``dump_to_zarr`` :348) and later turns every record into the index of
the cell wall it sits on, the fraction of the cell traverse and the residence time
(``find_ind_frac_tres`` :245).  Here the log is still in HBM as 128-byte records, and the same quantities
come out of one pass of ``sdk_log_budget`` over it.
"""

from __future__ import annotations

import numpy as np

from . import _lib


def first_last(offset):
    """Index of the first and last record of every particle (reference ``first_last_neither`` :114)."""
    offset = np.asarray(offset)
    return offset[:-1].copy(), offset[1:] - 1


def find_ind_frac_tres_device(particle, chunk=-1):
    """``(ind1, ind2, frac, tres)`` as torch tensors on the device for one chunk of the particle's log."""
    import torch  # noqa: PLC0415

    if not particle.save_raw:
        raise AttributeError("This is not a particle_rawlist object")
    if not particle._raw:
        raise ValueError("the particle has no crossing log yet")
    if not particle._has_dep:
        raise NotImplementedError("the budget post-processing needs particles with a depth (so does the reference)")
    ch = particle._raw[chunk]
    records, offset = ch["records"], ch["offset"]
    n_rec = records.shape[0]
    od = particle.ocedata
    dev = records.device
    rows = 5 if od.tp.has_face else 4
    ind1 = torch.empty((rows, n_rec), dtype=torch.int16, device=dev)
    ind2 = torch.empty((rows, n_rec), dtype=torch.int16, device=dev)
    frac = torch.empty(n_rec, dtype=torch.float64, device=dev)
    tres = torch.empty(n_rec, dtype=torch.float64, device=dev)
    work = torch.empty(n_rec, dtype=torch.float64, device=dev)
    _lib.check(
        _lib.lib().sdk_log_budget(
            od.grid_handle(), records.data_ptr(), offset.data_ptr(), particle.N, n_rec, ind1.data_ptr(), ind2.data_ptr(),
            frac.data_ptr(), tres.data_ptr(), work.data_ptr(), _lib.current_stream(),
        ),
        "sdk_log_budget",
    )
    return ind1, ind2, frac, tres


def find_ind_frac_tres(particle, oce=None, chunk=-1):
    """``ind1, ind2, frac, tres, last, first`` like the reference's ``find_ind_frac_tres(neo, oce)``
    (:245, ``by_type=True``, no regions), computed from the particle's device-resident crossing log."""
    del oce  # the particle knows its OceData
    ind1, ind2, frac, tres = find_ind_frac_tres_device(particle, chunk)
    first, last = first_last(particle._raw[chunk]["offset"].cpu().numpy())
    return ind1.cpu().numpy(), ind2.cpu().numpy(), frac.cpu().numpy(), tres.cpu().numpy(), last, first


# ---------------------------------------------------------------------------------------------------------
# the budget itself (reference calculate_budget :603)
# ---------------------------------------------------------------------------------------------------------


def _field_on_device(a, device):
    import torch  # noqa: PLC0415

    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float64)).to(device)


def prefetch_scalar(ds_slc, scalar_names, device=None):
    """name -> fp64 CUDA tensor for the budget terms of one model time step (reference ``prefetch_scalar`` :398)."""
    import torch  # noqa: PLC0415

    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    return {var: _field_on_device(ds_slc[var], device) for var in scalar_names}


def prefetch_vector(ds_slc, xname="sxprime", yname="syprime", zname="szprime", same_size=True, device=None):
    """The wall concentrations in x, y, z stacked into one ``(3, ...)`` CUDA tensor (reference ``prefetch_vector``
    :548).  With ``same_size=False`` the three arrays may differ in shape and are padded to the largest."""
    import torch  # noqa: PLC0415

    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    parts = [_field_on_device(ds_slc[k], device) for k in (xname, yname, zname)]
    if same_size:
        return torch.stack(parts)  # raises if they are not
    shape = tuple(max(p.shape[d] for p in parts) for d in range(parts[0].ndim))
    out = torch.empty((3,) + shape, dtype=torch.float64, device=device)
    for k, p in enumerate(parts):
        out[(k,) + tuple(slice(0, m) for m in p.shape)] = p
    return out


def calculate_budget_device(particle, data_slice, rhs_list, prefetch_vector_kwarg=None, lhs_name="lhs", dir_of_time=-1, chunk=-1):
    """``calculate_budget`` with everything staying in HBM: ``(contr, trc_conc, first, last)`` where ``contr`` maps
    every name of ``rhs_list``, ``"error"`` and ``lhs_name`` to a CUDA tensor of ``n_rec - 1`` values."""
    import torch  # noqa: PLC0415

    rhs_list = list(rhs_list)
    if not 1 <= len(rhs_list) <= _lib.MAX_BUDGET_TERMS:
        raise ValueError(f"between 1 and {_lib.MAX_BUDGET_TERMS} right-hand-side terms are supported, got {len(rhs_list)}")
    if dir_of_time not in (1, -1):
        raise ValueError("dir_of_time is -1 (backward) or 1 (forward)")
    ind1, ind2, frac, tres = find_ind_frac_tres_device(particle, chunk)
    ch = particle._raw[chunk]
    records, offset = ch["records"], ch["offset"]
    n_rec = records.shape[0]
    dev = records.device
    terms = prefetch_scalar(data_slice, rhs_list + [lhs_name], dev)
    walls = prefetch_vector(data_slice, device=dev, **(prefetch_vector_kwarg or {}))
    has_face = bool(particle.ocedata.tp.has_face)
    want_ndim = 4 if has_face else 3
    f = _lib.BudgetFields()
    f.n_terms, f.dir_of_time, f.rows = len(rhs_list), int(dir_of_time), 5 if has_face else 4

    def shape4(t, what):
        if t.ndim != want_ndim:
            raise ValueError(f"{what} has {t.ndim} dimensions, the grid needs {want_ndim}")
        sh = tuple(t.shape)
        return sh if has_face else (sh[0], 1, sh[1], sh[2])

    cell = shape4(terms[lhs_name], lhs_name)
    for k, name in enumerate(rhs_list):
        if shape4(terms[name], name) != cell:
            raise ValueError(f"{name} has shape {tuple(terms[name].shape)}, {lhs_name} has {tuple(terms[lhs_name].shape)}")
        f.term[k] = terms[name].data_ptr()
    f.lhs, f.walls = terms[lhs_name].data_ptr(), walls.data_ptr()
    f.cell_shape[:] = cell
    f.wall_shape[:] = shape4(walls[0], "the wall concentration")
    conc = torch.empty(n_rec, dtype=torch.float64, device=dev)
    m = max(n_rec - 1, 0)
    contr = torch.empty((len(rhs_list) + 2, m), dtype=torch.float64, device=dev)
    work = torch.empty(max(n_rec, 1), dtype=torch.uint8, device=dev)
    _lib.check(
        _lib.lib().sdk_log_calculate_budget(
            records.data_ptr(), offset.data_ptr(), particle.N, n_rec, ind1.data_ptr(), ind2.data_ptr(), frac.data_ptr(),
            tres.data_ptr(), f, conc.data_ptr(), contr.data_ptr(), work.data_ptr(), _lib.current_stream(),
        ),
        "sdk_log_calculate_budget",
    )
    out = {name: contr[k] for k, name in enumerate(rhs_list)}
    out["error"] = contr[len(rhs_list)]
    out[lhs_name] = contr[len(rhs_list) + 1]
    first = offset[:-1]
    last = offset[1:] - 1
    return out, conc, first, last


def calculate_budget(particle_array, data_slice, rhs_list, prefetch_vector_kwarg=dict(), lhs_name="lhs", dir_of_time=-1):  # noqa: B006
    """``contr_dic, trc_conc, first, last`` like the reference's ``calculate_budget`` (:603).

    ``particle_array`` is the :class:`Particle` whose crossing log is still on the device (what the reference
    would have dumped to zarr and read back); ``data_slice`` maps the names in ``rhs_list``, ``lhs_name`` and
    the wall concentrations (``prefetch_vector_kwarg``: ``xname, yname, zname, same_size``) to arrays of one
    model time step.  numpy arrays come back."""
    contr, conc, first, last = calculate_budget_device(particle_array, data_slice, rhs_list, prefetch_vector_kwarg, lhs_name, dir_of_time)
    return {k: v.cpu().numpy() for k, v in contr.items()}, conc.cpu().numpy(), first.cpu().numpy(), last.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------
# wall values around every record and the particle / data consistency check (reference :406, :470)
# ---------------------------------------------------------------------------------------------------------

_DIRECTIONS = np.array([np.pi / 2, -np.pi / 2, np.pi, 0])  # up, down, left, right (reference topology.py:30)


def _wall_values_device(particle, walls, conv=None, scalar=True, chunk=-1, want_velocities=False):
    """``c_list`` (6, n) [, ``u_list`` (6, n), Lagrangian convergence (n), Eulerian convergence (n)] as CUDA tensors."""
    import torch  # noqa: PLC0415

    if not particle.save_raw:
        raise AttributeError("This is not a particle_rawlist object")
    if not particle._raw:
        raise ValueError("the particle has no crossing log yet")
    od = particle.ocedata
    if not od.tp.has_face:
        raise NotImplementedError("wall values are implemented for grids with faces (the branch the reference tests)")
    ch = particle._raw[chunk]
    records = ch["records"]
    n_rec = records.shape[0]
    dev = records.device
    arrays = [_field_on_device(a, dev) for a in walls]
    f = _lib.WallFields()
    for t, name in zip(arrays, ("x", "y", "z")):
        if t.ndim != 4:
            raise ValueError(f"the {name}-wall field has {t.ndim} dimensions, (Z, face, Y, X) expected")
        getattr(f, f"shape_{name}")[:] = tuple(t.shape)
    f.x_walls, f.y_walls, f.z_walls = (t.data_ptr() for t in arrays)
    conv_t = None
    if conv is not None:
        conv_t = _field_on_device(conv, dev)
        if conv_t.ndim != 4:
            raise ValueError("the convergence field must be (Z, face, Y, X)")
        f.conv = conv_t.data_ptr()
        f.shape_c[:] = tuple(conv_t.shape)
    rot = np.pi - _DIRECTIONS[:, None] + _DIRECTIONS[None, :]  # topology.py:248, [edge][new_edge]
    f.rot_cos[:] = np.cos(rot).reshape(-1).tolist()
    f.rot_sin[:] = np.sin(rot).reshape(-1).tolist()
    f.scalar = int(bool(scalar))
    c_list = torch.empty((6, n_rec), dtype=torch.float64, device=dev)
    u_list = torch.empty((6, n_rec), dtype=torch.float64, device=dev) if want_velocities else None
    lag = torch.empty(n_rec, dtype=torch.float64, device=dev) if want_velocities else None
    eul = torch.empty(n_rec, dtype=torch.float64, device=dev) if (want_velocities and conv_t is not None) else None
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    _lib.check(
        _lib.lib().sdk_log_wall_values(od.grid_handle(), records.data_ptr(), n_rec, f, c_list.data_ptr(), ptr(u_list), ptr(lag), ptr(eul), _lib.current_stream()),
        "sdk_log_wall_values",
    )
    return c_list, u_list, lag, eul


def read_wall_list(neo, tp=None, prefetch=None, scalar=True):
    """The values of the three wall fields ``prefetch = (x walls, y walls, z walls)`` on the six walls of the cell
    of every record of the particle's crossing log: ``(6, n)`` = left, right, y-bottom, y-top, deep, shallow.
    Reference ``read_wall_list(neo, tp, prefetch, scalar)`` (:406); ``neo`` is the :class:`Particle` whose log is on
    the device."""
    del tp  # the particle's OceData knows its topology
    if prefetch is None or len(prefetch) != 3:
        raise ValueError("prefetch must be the three wall arrays (x, y, z)")
    return _wall_values_device(neo, prefetch, scalar=scalar)[0].cpu().numpy()


def check_particle_data_compat(xrpt, xrslc, tp=None, use_tracer_name=None, wall_names=("sx", "sy", "sz"), conv_name="divus",
                               debug=False, allclose_kwarg={}):  # noqa: B006
    """Can the Lagrangian budget be closed with this model output?  The convergence of the wall fluxes seen by the
    particles must equal the Eulerian convergence ``xrslc[conv_name]`` at their cells.  Reference
    ``check_particle_data_compat`` (:470); returns ``(ok, extra)`` with ``extra = (u_list, c_list, lagrangian_conv,
    eulerian_conv)`` when ``debug`` else ``None``."""
    del tp
    if isinstance(use_tracer_name, str):
        wall_names = tuple(use_tracer_name + i for i in ["x", "y", "z"])
        conv_name = "divu" + use_tracer_name
    elif use_tracer_name is not None:
        raise ValueError("use_tracer_name has to be a string.")
    if not xrpt._has_dep:
        raise NotImplementedError("This functionality only support 3D simulation at the moment.")
    c_list, u_list, lag, eul = _wall_values_device(xrpt, [xrslc[k] for k in wall_names], conv=xrslc[conv_name], want_velocities=True)
    lag_h, eul_h = lag.cpu().numpy(), eul.cpu().numpy()
    extra = (u_list.cpu().numpy(), c_list.cpu().numpy(), lag_h, eul_h) if debug else None
    return bool(np.allclose(lag_h, eul_h, **allclose_kwarg)), extra
