"""ctypes binding of ``libseaduck_b200.so`` (C ABI declared in ``include/seaduck_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / :func:`build`.  There is no CPU
fallback: if the shared object is missing or cannot be loaded every compute entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.environ.get("SEADUCK_B200_LIB") or os.path.join(_HERE, "libseaduck_b200.so")
SOURCES = ["advect.cu", "locate.cu", "interp.cu", "budget.cu", "stencil.cu", "selftest.cu", "exchange.cu", "sort.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-fmad=false",  # reference arithmetic is numpy: no fused multiply-add (index parity)
    "-std=c++17",
    "-shared",
    "-Xcompiler",
    "-fPIC",
]

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p


def _fields(spec):
    out = []
    for typ, names in spec:
        for n in names.split():
            out.append((n, typ))
    return out


class GridDesc(C.Structure):
    _fields_ = _fields(
        [
            (C.c_int32, "typ has_face nface ny nx gny gnx hmode n_zl n_z n_zc reserved"),
            (vp, "nbr_face nbr_edge XC YC XG YG dX dY CS SN Zl dZl Z dZ corner_tab lon lat dlon dlat coslat32"),
            (C.c_int32, "dlon_is_f32 reserved2"),
        ]
    )


class Velocity(C.Structure):
    _fields_ = _fields(
        [
            (vp, "U V W Vol"),
            (
                C.c_int32,
                "dtype vol_dtype has_time it0 nt has_z nzU nzW nyU nxU nyV nxV u_stag v_stag vol_has_z transport",
            ),
        ]
    )


PARTICLE_F64 = "lon lat dep t".split()
PARTICLE_I32 = "face iy ix izl it".split()
PARTICLE_F64B = "rx ry rzl u v w du dv dw px py".split()
PARTICLE_F32 = "dx dy".split()
PARTICLE_F64C = "bzl dzl vol".split()
PARTICLE_OUT = "status niter ncross active".split()


class Particles(C.Structure):
    _fields_ = (
        [("n", C.c_int64)]
        + [(n, vp) for n in PARTICLE_F64 + PARTICLE_I32 + PARTICLE_F64B + PARTICLE_F32 + PARTICLE_F64C + PARTICLE_OUT]
    )


class AdvectParams(C.Structure):
    _fields_ = [
        ("tol", C.c_double),
        ("trim_tol", C.c_double),
        ("max_iteration", C.c_int32),
        ("kick_back", C.c_int32),
        ("has_dep", C.c_int32),
        ("refill_threshold", C.c_int32),
        ("blocks_per_sm", C.c_int32),
        ("threads_per_block", C.c_int32),
        ("reserve_sms", C.c_int32),
        ("keep_status", C.c_int32),
        ("stats", vp),
        ("ship", vp),
        ("ship_index", vp),
        ("ship_step", C.c_uint64),
    ]


LOG_I32 = "fc it izl iy ix vs".split()
LOG_F64 = "rzl rx ry tt uu vv ww du dv dw xx yy zz".split()


class Log(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("offset", vp), ("count", vp), ("records", vp), ("emit_start", C.c_int32), ("continue_mode", C.c_int32)]


# numpy view of sdk_log_record (128 bytes)
LOG_RECORD = [(n, "<f8") for n in LOG_F64] + [(n, "<i4") for n in LOG_I32]


LOCATED = (
    "face iy ix rx ry cs sn dx dy bx by px py iz rz dz bz izl_near rzl_near dzl_near bzl_near "
    "izl rzl dzl bzl it rt dt bt iz_lin rz_lin dz_lin bz_lin it_lin rt_lin dt_lin bt_lin status"
).split()


class Located(C.Structure):
    _fields_ = [(n, vp) for n in LOCATED]


class TrackIO(C.Structure):
    _fields_ = [("n", C.c_int64)] + [
        (n, vp)
        for n in "lon lat dep t out_records out_lon out_lat out_dep out_face out_iy out_ix out_izl out_ncross out_status out_stats".split()
    ] + [("sort", C.c_int32), ("reserved", C.c_int32)]


MAX_RANKS = 16
SIGNAL_ARRIVED, SIGNAL_CONSUMED, SIGNAL_ERROR, SIGNAL_WORDS = 0, MAX_RANKS, 2 * MAX_RANKS, 2 * MAX_RANKS + 1

# numpy view of sdk_snapshot_record (32 bytes) and the layout of its cell word
SNAPSHOT_RECORD = [("lon", "<f8"), ("lat", "<f8"), ("dep", "<f8"), ("cell", "<u8")]
CELL_SHIFT = {"face": 52, "iy": 32, "ix": 12, "izl": 0}
CELL_BITS = {"face": 6, "iy": 20, "ix": 20, "izl": 12}


class Exchange(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("n_slots", C.c_int32), ("reserved", C.c_int32), ("n_pad", C.c_int64),
        ("recv", vp * MAX_RANKS), ("recv_multicast", vp), ("signal", vp * MAX_RANKS), ("ticket", vp),
    ]


MAX_BUDGET_TERMS = 8


class BudgetFields(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32), ("dir_of_time", C.c_int32), ("rows", C.c_int32), ("reserved", C.c_int32),
        ("cell_shape", C.c_int32 * 4), ("wall_shape", C.c_int32 * 4),
        ("term", C.c_void_p * MAX_BUDGET_TERMS), ("lhs", C.c_void_p), ("walls", C.c_void_p),
    ]


class WallFields(C.Structure):
    _fields_ = [
        ("x_walls", C.c_void_p), ("y_walls", C.c_void_p), ("z_walls", C.c_void_p), ("conv", C.c_void_p),
        ("shape_x", C.c_int32 * 4), ("shape_y", C.c_int32 * 4), ("shape_z", C.c_int32 * 4), ("shape_c", C.c_int32 * 4),
        ("rot_cos", C.c_double * 16), ("rot_sin", C.c_double * 16),
        ("scalar", C.c_int32), ("reserved", C.c_int32),
    ]


class SdkError(RuntimeError):
    pass


_lib = None


def build(verbose=False, force=False):
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU): one object per source
    file, compiled side by side and only when the file or a header changed, then one link."""
    from concurrent.futures import ThreadPoolExecutor

    src_dir = os.path.join(_HERE, "csrc")
    obj_dir = os.path.join(_ROOT, "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(_ROOT, "include", "seaduck_b200.h"))
    headers.append(os.path.abspath(__file__))  # the flags live here
    newest_header = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    jobs = []
    for name in SOURCES:
        src = os.path.join(src_dir, name)
        obj = os.path.join(obj_dir, name[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append(["nvcc"] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for rc, cmd in zip(pool.map(lambda c: subprocess.call(c, cwd=_ROOT), jobs), jobs):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, cmd)
    objs = [os.path.join(obj_dir, name[:-3] + ".o") for name in SOURCES]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_PATH] + objs, cwd=_ROOT)
    return SO_PATH


def is_stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    deps.append(os.path.join(_ROOT, "include", "seaduck_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    """Load the library; raise (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SdkError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "seaduck_b200 has no CPU fallback."
            )
        L = C.CDLL(SO_PATH)
        L.sdk_last_error.restype = C.c_char_p
        L.sdk_particles_bytes.restype = C.c_int64
        L.sdk_particles_bytes.argtypes = [C.c_int64]
        L.sdk_grid_create.argtypes = [C.POINTER(GridDesc), C.POINTER(vp)]
        L.sdk_grid_destroy.argtypes = [vp]
        L.sdk_advect.argtypes = [vp, C.POINTER(Velocity), C.POINTER(Particles), C.c_double, C.POINTER(AdvectParams), C.POINTER(Log), C.c_int, vp]
        L.sdk_advect_host.argtypes = [vp, C.POINTER(Velocity), C.POINTER(Particles), C.c_double, C.POINTER(AdvectParams), vp, vp]
        L.sdk_get_u_du.argtypes = [vp, C.POINTER(Velocity), C.POINTER(Particles), C.c_int, vp]
        L.sdk_advect_profile.argtypes = [vp, C.POINTER(Velocity), C.POINTER(Particles), C.POINTER(AdvectParams), C.POINTER(Log)]
        L.sdk_grid_build_locator.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, vp]
        L.sdk_locate.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int32, C.POINTER(Located), vp]
        L.sdk_time_index.argtypes = [vp, vp, C.c_int32, vp, C.c_int64, vp]
        L.sdk_track_scratch_bytes.restype = C.c_int64
        L.sdk_track_scratch_bytes.argtypes = [C.c_int64]
        L.sdk_track_host.argtypes = [vp, C.POINTER(Velocity), C.POINTER(TrackIO), vp, C.c_int32, C.c_double, C.POINTER(AdvectParams), vp, vp]
        L.sdk_finish_stop.argtypes = [vp, vp, vp, C.c_int32, C.c_double, vp, vp, C.c_int64, vp]
        L.sdk_log_calculate_budget.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, vp, vp, vp, C.POINTER(BudgetFields), vp, vp, vp, vp]
        L.sdk_log_wall_values.argtypes = [vp, vp, C.c_int64, C.POINTER(WallFields), vp, vp, vp, vp, vp]
        L.sdk_gather_plane.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp]
        L.sdk_wall_stencil.argtypes = [C.c_int, vp, vp, vp, C.c_int64, C.c_int64, C.c_int64, vp]
        L.sdk_upwind3_z.argtypes = [vp, vp, vp, C.c_int64, C.c_int64, vp]
        L.sdk_log_budget.argtypes = [vp, vp, vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp]
        L.sdk_pack_records.argtypes = [C.POINTER(Particles), vp, vp, vp]
        L.sdk_cell_sort_scratch_bytes.restype = C.c_int64
        L.sdk_cell_sort_scratch_bytes.argtypes = [C.c_int64]
        L.sdk_cell_sort.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
        L.sdk_particles_gather.argtypes = [C.POINTER(Particles), C.POINTER(Particles), vp, vp]
        L.sdk_pack_ship.argtypes = [C.POINTER(Exchange), C.POINTER(Particles), vp, C.c_uint64, vp]
        L.sdk_exchange_wait.argtypes = [C.POINTER(Exchange), C.c_uint64, C.c_int, C.c_int, vp]
        L.sdk_selftest_math.argtypes = [C.c_int, C.c_int64, vp, vp, vp, vp, vp]
        L.sdk_selftest_cell_move.argtypes = [vp, C.c_int64, vp, vp, vp]
        # (interp.py narrows the structure pointers of these; typed here so that no caller can reach them untyped:
        # ctypes would pass 64-bit pointers as 32-bit ints)
        L.sdk_build_masks.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp]
        L.sdk_struct_size.argtypes = [C.c_int]
        L.sdk_struct_size.restype = C.c_int64
        L.sdk_interp_scalar.argtypes = [vp] * 6
        L.sdk_interp_scalar_multi.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp]
        L.sdk_interp_vector.argtypes = [vp] * 6 + [C.c_int, vp, vp, vp]
        L.sdk_interp_host_scratch_bytes.restype = C.c_int64
        L.sdk_interp_host_scratch_bytes.argtypes = [C.c_int64, C.c_int32]
        L.sdk_interp_host.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
        L.sdk_status_or.argtypes = [vp, C.c_int64, vp, vp]
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().sdk_last_error()
        raise SdkError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device (or pinned host) pointer of a torch tensor, ``None`` -> NULL."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream():
    import torch  # noqa: PLC0415

    return torch.cuda.current_stream().cuda_stream
