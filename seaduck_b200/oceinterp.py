"""``OceInterp``: one call for Eulerian or Lagrangian interpolation (reference ``seaduck/oceinterp.py:14``)."""

import warnings

import numpy as np

from .eulerian import Position
from .kernel_weight import KnW
from .lagrangian import Particle, uknw, vknw
from .ocedata import OceData
from .utils import convert_time

lagrange_token = "__particle."


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
    if isinstance(var_list, dict):
        kernel_list = list(var_list.values())
        var_list = list(var_list.keys())
    elif isinstance(var_list, (str, tuple)):
        var_list = [var_list]
    elif not isinstance(var_list, list):
        raise ValueError("var_list type not recognized.")
    if kernel_list is None:
        kernel_list = []
        the_kernel = KnW(**kernel_kwarg)
        for i in var_list:
            if isinstance(i, str):
                kernel_list.append(the_kernel)
            elif isinstance(i, tuple):
                kernel_list.append((the_kernel, the_kernel) if kernel_kwarg != {} else (uknw, vknw))
            else:
                raise ValueError("members of var_list need to be made up of string or tuples")
    if isinstance(t, np.ndarray):
        if np.issubdtype(t.dtype, np.datetime64) or np.issubdtype(t.dtype, str):
            t = np.array([convert_time(some_t) for some_t in t])
    elif isinstance(t, (np.datetime64, str)):
        t = convert_time(t)
    elif isinstance(t, float):
        pass
    else:
        raise ValueError("time format not supported")
    if not lagrangian:
        pt = Position()
        pt.from_latlon(x=x, y=y, z=z, t=t, data=od)
        for var in var_list:
            if lagrange_token in var:
                raise AttributeError("__particle variables is only available for Lagrangian Particles")
        return pt.interpolate(var_list, kernel_list)
    try:
        assert len(t) > 1
    except AssertionError as exc:
        raise ValueError("There needs to be at least two time steps to run the lagrangian Particle") from exc
    t_start = t[0]
    t_nec = t[1:]
    pt = Particle(x=x, y=y, z=z, t=np.ones_like(x) * t_start, data=od, **lagrange_kwarg)
    stops, raw = pt.to_list_of_time(t_nec, update_stops=update_stops, return_in_between=return_in_between)
    to_return = []
    for i, var in enumerate(var_list):
        if var == lagrange_token + "raw":
            to_return.append(raw)
        elif lagrange_token in var:
            to_return.append([getattr(snap, var[len(lagrange_token):]) for snap in raw])
        else:
            to_return.append([snap.interpolate(var, kernel_list[i]) for snap in raw])
    if return_pt_time:
        return stops, to_return
    if return_in_between:
        warnings.warn("Some of the returns is not on the times you specified.")
    return to_return
