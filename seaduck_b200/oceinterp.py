"""``OceInterp``: one call for Eulerian or Lagrangian interpolation (same signature, return values and exceptions
as the reference's ``seaduck/oceinterp.py:14``).

The call is a thin front end: it pairs every requested variable with a kernel, turns the time argument into seconds,
and then either locates the points once and interpolates (``Position``), or carries particles through the requested
times (``Particle.to_list_of_time``) and evaluates every request on every snapshot.  Requests of the form
``"__particle.<attr>"`` read an attribute of the snapshots instead of interpolating; ``"__particle.raw"`` hands the
snapshots themselves back.
"""

import warnings

import numpy as np

from .eulerian import Position
from .kernel_weight import KnW
from .lagrangian import Particle, uknw, vknw
from .ocedata import OceData
from .utils import convert_time

lagrange_token = "__particle."


def _requests(var_list, kernel_list, kernel_kwarg):
    """``(variables, kernels)`` as two parallel lists.  A dict maps variables to kernels; without kernels every scalar
    gets ``KnW(**kernel_kwarg)`` and every (u, v) pair the particle kernels (or that same kernel twice when
    ``kernel_kwarg`` asks for something specific)."""
    if isinstance(var_list, dict):
        return list(var_list), list(var_list.values())
    if isinstance(var_list, (str, tuple)):
        var_list = [var_list]
    if not isinstance(var_list, list):
        raise ValueError("var_list type not recognized.")
    if kernel_list is not None:
        return var_list, kernel_list
    shared = KnW(**kernel_kwarg)
    pair = (shared, shared) if kernel_kwarg else (uknw, vknw)
    kernels = []
    for var in var_list:
        if not isinstance(var, (str, tuple)):
            raise ValueError("members of var_list need to be made up of string or tuples")
        kernels.append(shared if isinstance(var, str) else pair)
    return var_list, kernels


def _seconds(t):
    """Time argument -> seconds since 1970: floats pass, datetime64 / strings (scalars or arrays) are converted."""
    if isinstance(t, np.ndarray):
        stamped = np.issubdtype(t.dtype, np.datetime64) or np.issubdtype(t.dtype, str)
        return np.array([convert_time(one) for one in t]) if stamped else t
    if isinstance(t, (np.datetime64, str)):
        return convert_time(t)
    if isinstance(t, float):
        return t
    raise ValueError("time format not supported")


def _at_points(od, variables, kernels, x, y, z, t):
    """Eulerian branch (reference ``oceinterp.py:107-112``).  Large batches of host coordinates go through the library's
    pipelined host entry (``interp.interpolate_host``: upload, locate, interpolate and read-back of different chunks
    overlap); anything else through ``Position.from_latlon`` + ``Position.interpolate``.  Same values either way."""
    from . import interp  # noqa: PLC0415

    if isinstance(x, np.ndarray) and isinstance(y, np.ndarray) and x.ndim == 1 and len(x) >= interp.HOST_PATH_THRESHOLD:
        # (a coordinate the dataset has no axis for is never looked at: do not broadcast a scalar to N values for it)
        if not od.readiness["time"] and np.ndim(t) == 0:
            t = None
        if not (od.readiness["Z"] or od.readiness["Zl"]) and np.ndim(z) == 0:
            z = None
        x, y, z, t = Position()._coordinate_arrays(x, y, z, t)
        return interp.interpolate_host(od, x, y, z, t, variables, kernels)
    return Position().from_latlon(x=x, y=y, z=z, t=t, data=od).interpolate(variables, kernels)


def _along_trajectories(od, variables, kernels, x, y, z, t, lagrange_kwarg, update_stops, return_in_between):
    if np.ndim(t) == 0 or len(t) < 2:
        raise ValueError("There needs to be at least two time steps to run the lagrangian Particle")
    pt = Particle(x=x, y=y, z=z, t=np.ones_like(x) * t[0], data=od, **lagrange_kwarg)
    stops, snapshots = pt.to_list_of_time(t[1:], update_stops=update_stops, return_in_between=return_in_between)

    def answer(var, knw):
        if var == lagrange_token + "raw":
            return snapshots
        if isinstance(var, str) and lagrange_token in var:
            attr = var[len(lagrange_token):]
            return [getattr(snap, attr) for snap in snapshots]
        return [snap.interpolate(var, knw) for snap in snapshots]

    return stops, [answer(var, knw) for var, knw in zip(variables, kernels)]


def OceInterp(
    od,
    var_list,
    x,
    y,
    z,
    t,
    kernel_list=None,
    lagrangian=False,
    lagrange_kwarg={},
    update_stops="default",
    return_in_between=True,
    return_pt_time=True,
    kernel_kwarg={},
):
    """Interpolate ``var_list`` at points, or along trajectories when ``lagrangian=True``."""
    if not isinstance(od, OceData):
        od = OceData(od)
    variables, kernels = _requests(var_list, kernel_list, kernel_kwarg)
    t = _seconds(t)
    if not lagrangian:
        if any(lagrange_token in var for var in variables):
            raise AttributeError("__particle variables is only available for Lagrangian Particles")
        return _at_points(od, variables, kernels, x, y, z, t)
    stops, results = _along_trajectories(od, variables, kernels, x, y, z, t, lagrange_kwarg, update_stops, return_in_between)
    if return_pt_time:
        return stops, results
    if return_in_between:
        warnings.warn("Some of the returns is not on the times you specified.")
    return results
