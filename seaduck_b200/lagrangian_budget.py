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
