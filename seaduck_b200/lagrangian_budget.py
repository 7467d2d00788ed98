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
