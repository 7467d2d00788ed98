"""Small host-side helpers under the reference's names (``seaduck/utils.py``)."""

import numpy as np

from .eulerian import to_180, weight_f_node  # noqa: F401

deg2m = 6271e3 * np.pi / 180  # sic, reference utils.py:306

_EPOCH = np.datetime64("1970-01-01")


def local_to_latlon(u, v, cs, sn):
    """Rotate grid-aligned vector components to east/north ones (reference ``utils.py:206``)."""
    return u * cs - v * sn, u * sn + v * cs


def convert_time(time):
    """Seconds since 1970-01-01 of a date string or an ``np.datetime64`` (reference ``utils.py:688``)."""
    if isinstance(time, str):
        time = np.datetime64(time)
    elif not isinstance(time, np.datetime64):
        raise ValueError(f"Unsupported time format {type(time)}")
    return (time - _EPOCH) / np.timedelta64(1, "s")


def easy_3d_cube(lon, lat, dep, tim, print_total_number=False):
    """Coordinates of a regular lon x lat x depth box of points at one time (reference ``utils.py:704``).

    Every argument but ``tim`` is ``(first, last, count)``.  Order of the result: level by level, within a level
    row by row (latitude), longitude fastest -- the order the reference's nested meshgrids produce.
    """
    axes = [np.linspace(first, last, count) for first, last, count in (lon, lat, dep)]
    n_lon, n_lat, n_dep = (len(a) for a in axes)
    x = np.tile(axes[0], n_lat * n_dep)
    y = np.tile(np.repeat(axes[1], n_lon), n_dep)
    z = np.repeat(axes[2], n_lon * n_lat)
    t = np.full(x.shape, convert_time(tim), dtype=x.dtype)
    if print_total_number:
        print(f"A total of {x.size} positions are defined.")
    return x, y, z, t


def _general_len(thing):
    """``len`` that counts a scalar as one element."""
    return len(thing) if hasattr(thing, "__len__") else 1
