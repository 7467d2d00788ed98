"""Small host-side helpers with the reference's names (``seaduck/utils.py``)."""

import numpy as np

from .eulerian import to_180, weight_f_node  # noqa: F401

deg2m = 6271e3 * np.pi / 180  # sic, reference utils.py:306


def local_to_latlon(u, v, cs, sn):
    """Convert a local vector to east/north components (reference ``utils.py:206``)."""
    uu = u * cs - v * sn
    vv = u * sn + v * cs
    return uu, vv


def convert_time(time):
    """Seconds after 1970-01-01 for a string or ``np.datetime64`` (reference ``utils.py:688``)."""
    t0 = np.datetime64("1970-01-01")
    one_sec = np.timedelta64(1, "s")
    if isinstance(time, str):
        return (np.datetime64(time) - t0) / one_sec
    if isinstance(time, np.datetime64):
        return (time - t0) / one_sec
    raise ValueError(f"Unsupported time format {type(time)}")


def easy_3d_cube(lon, lat, dep, tim, print_total_number=False):
    """4-D coordinates of a regular cube of points (reference ``utils.py:704``)."""
    east, west, Nlon = lon
    south, north, Nlat = lat
    shallow, deep, Ndep = dep
    t_in_sec = convert_time(tim)
    x = np.linspace(east, west, Nlon)
    y = np.linspace(south, north, Nlat)
    x, y = np.meshgrid(x, y)
    x = x.ravel()
    y = y.ravel()
    levels = np.linspace(shallow, deep, Ndep)
    x, z = np.meshgrid(x, levels)
    y, z = np.meshgrid(y, levels)
    x = x.ravel()
    y = y.ravel()
    z = z.ravel()
    t = np.ones_like(x) * t_in_sec
    if print_total_number:
        print(f"A total of {len(x)} positions are defined.")
    return x, y, z, t


def _general_len(thing):
    try:
        return len(thing)
    except TypeError:
        return 1
