// Device-side helpers shared by the seaduck_b200 kernels (sm_100a).
//
// Numerics follow the reference operation by operation (fp64, no fused multiply-add: this file is
// compiled with -fmad=false) because the exit wall of a cell is an argmin over log() based times
// (reference seaduck/utils.py:859-896) and index trajectories have to match bit for bit.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#include "../../include/seaduck_b200.h"

namespace sdk {

// Grid description as the kernels see it.  Passed as a __grid_constant__ kernel parameter, so the
// face tables sit in the constant bank.
struct GridDev {
    int typ, has_face, nface, ny, nx, gny, gnx, hmode, n_zl, n_z, n_zc;
    signed char nbr_face[SDK_MAX_FACES * 4];
    signed char nbr_edge[SDK_MAX_FACES * 4];
    const float *XC, *YC, *XG, *YG, *dX, *dY, *CS, *SN, *Zl, *dZl, *Z, *dZ;
    const int32_t *corner_tab;
};

struct VelDev {
    const void *U, *V, *W, *Vol;
    int dtype, vol_dtype, has_time, it0, nt, has_z, nzU, nzW, nyU, nxU, nyV, nxV;
    int u_stag, v_stag, vol_has_z, transport, has_w;
    int64_t planeU, planeV, planeC;  // elements per (time, level) slab of U, V and of cell-centred fields
};

__device__ __forceinline__ double py_mod360(double a) {
    // python / numpy float "%" with a positive modulus (reference utils.py:201)
    double m = (fabs(a) < 360.0) ? a : fmod(a, 360.0);
    if (m != 0.0) {
        if (m < 0.0) m += 360.0;
    } else {
        m = 0.0;
    }
    return m;
}

// reference to_180, utils.py:199:  x % 360 % 360, then minus 360 in the upper half.
// The first "%" leaves m in [0, 360]; m == 360.0 happens when x is a tiny negative number (a
// particle sitting on a wall, 360 - 1e-14 rounds to 360) and is what the second "%" is there for:
// fmod(360, 360) = +0.  For m in [360, 720) fmod(m, 360) == m - 360 exactly, so no fmod is needed.
__device__ __forceinline__ double to_180(double x) {
    x = py_mod360(x);
    if (x >= 360.0) x = x - 360.0;
    double q = (x >= 180.0) ? 1.0 : 0.0;  // == x // 180 for x in [0, 360)
    return x + (-1.0) * q * 360.0;
}

// numpy.nan_to_num: nan -> 0, +-inf -> +-DBL_MAX.  One integer test on the exponent in the common
// (finite) case.
__device__ __forceinline__ double nan_to_num(double x) {
    if ((__double2hiint(x) & 0x7ff00000) == 0x7ff00000) {
        if (x != x) return 0.0;
        return x > 0 ? DBL_MAX : -DBL_MAX;
    }
    return x;
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0) - (x < 0)); }

__device__ __forceinline__ int wrap_index(int i, int n) { return i < 0 ? i + n : i; }  // numpy style

template <typename T>
__device__ __forceinline__ double ldf(const void *base, int64_t idx) {
    return (double)__ldg(reinterpret_cast<const T *>(base) + idx);
}

// One C-grid step across cell walls (reference _llc_ind_tend topology.py:149, _box_ind_tend :100,
// _x_per_ind_tend :124).  dir: 0 up, 1 down, 2 left, 3 right.  Returns SDK_ST_* bits (0 = fine)
// and, when a face edge was crossed, the edge of the new face we came in through (else -1).
__device__ __forceinline__ int cell_move(const GridDev &g, int dir, int &f, int &iy, int &ix,
                                         int &came_in) {
    came_in = -1;
    int ny_ = iy + (dir == 0) - (dir == 1);
    int nx_ = ix + (dir == 3) - (dir == 2);
    const int iymax = g.ny - 1, ixmax = g.nx - 1;
    const bool off = ny_ < 0 || ny_ > iymax || nx_ < 0 || nx_ > ixmax;
    if (!off) {
        iy = ny_;
        ix = nx_;
        return 0;
    }
    if (g.typ == SDK_TYP_LLC) {
        const int nf = g.nbr_face[f * 4 + dir];
        if (nf < 0) return SDK_ST_EDGE;  // reference keeps the old index and warns
        const int ne = g.nbr_edge[f * 4 + dir];
        const int s = dir >= 2 ? iy : ix;                  // coordinate along the crossed edge
        const bool flip = ((dir & 1) == 0) == ((ne & 1) == 0);
        const bool dest_x = ne <= 1;
        const int s_new = flip ? (dest_x ? ixmax : iymax) - s : s;
        f = nf;
        iy = ne == 0 ? iymax : (ne == 1 ? 0 : s_new);
        ix = ne == 2 ? 0 : (ne == 3 ? ixmax : s_new);
        came_in = ne;
        return 0;
    }
    if (ny_ < 0 || ny_ > iymax) return SDK_ST_OUT;
    if (g.typ == SDK_TYP_XPERIODIC) {
        iy = ny_;
        ix = nx_ > ixmax ? nx_ - ixmax : ixmax + nx_ + 1;  // sic: reference wraps to 1, not 0
        return 0;
    }
    return SDK_ST_OUT;
}

// Flat XG indices of the four corners LL, LR, UR, UL of cell (f, iy, ix)
// (reference find_px_py utils.py:495).  Cells on the top row / right column of a face whose XG has
// no extra row/column come from the host-built table.
__device__ __forceinline__ int corner_indices(const GridDev &g, int f, int iy, int ix, int64_t c[4]) {
    const int64_t base = (int64_t)f * g.gny * g.gnx;
    c[0] = base + (int64_t)iy * g.gnx + ix;
    if (ix + 1 <= g.gnx - 1 && iy + 1 <= g.gny - 1) {
        c[1] = c[0] + 1;
        c[2] = c[0] + g.gnx + 1;
        c[3] = c[0] + g.gnx;
        return 0;
    }
    const int slot = (iy == g.ny - 1) ? ix : g.nx + iy;
    const int32_t *t = g.corner_tab + ((int64_t)f * (g.nx + g.ny) + slot) * 3;
    const int32_t a = __ldg(t), b = __ldg(t + 1), d = __ldg(t + 2);
    if (a < 0 || b < 0 || d < 0) {
        c[1] = c[2] = c[3] = c[0];
        return SDK_ST_BAD_CELL;
    }
    c[1] = a;
    c[2] = b;
    c[3] = d;
    return 0;
}

// reference find_rx_ry_oceanparcel, utils.py:532
__device__ __forceinline__ void inverse_bilinear(double x, double y, const double pxi[4],
                                                 const double py[4], double &rx, double &ry) {
    const double x0 = pxi[0];
    double px[4];
    x = to_180(x - x0);
#pragma unroll
    for (int j = 0; j < 4; ++j) px[j] = to_180(pxi[j] - x0);
    const double a0 = px[0], a1 = -px[0] + px[1], a2 = -px[0] + px[3];
    const double a3 = ((px[0] - px[1]) + px[2]) - px[3];
    const double b0 = py[0], b1 = -py[0] + py[1], b2 = -py[0] + py[3];
    const double b3 = ((py[0] - py[1]) + py[2]) - py[3];
    const double aa = a3 * b2 - a2 * b3;
    const double bb = a3 * b0 - a0 * b3 + a1 * b2 - a2 * b1 + x * b3 - y * a3;
    const double cc = a1 * b0 - a0 * b1 + x * b1 - y * a1;
    const double det2 = bb * bb - 4 * aa * cc;
    const bool order1 = fabs(aa) < 1e-12;
    const bool order2 = !order1 && det2 >= 0;
    double r = -(cc / bb);
    if (order2) r = (-bb + sqrt(det2)) / (2 * aa);
    const double den = a1 + a3 * r;
    double q;
    if (fabs(den) < 1e-12)
        q = ((y - py[0]) / (py[1] - py[0]) + (y - py[3]) / (py[2] - py[3])) * 0.5;
    else
        q = (x - a0 - a2 * r) / den;
    rx = q - 0.5;
    ry = r - 0.5;
}

// Exit times of one axis (reference _stationary_time utils.py:821, _uleftright_from_udu :852 and the
// closed-wall rule of _time2wall :864-869)
//
// The reference evaluates both logarithms and then overwrites the time of a wall the flow cannot
// reach with -sign(tf).  Only reachable walls matter, and usually just one wall per axis is
// reachable, so the (expensive) logarithm is evaluated once with the arguments of that wall:
// tl = log(1 - k(1/2 + x0))/du == log(1 + k(-1/2 - x0))/du bit for bit.
__device__ __forceinline__ void wall_times(double u, double du, double x0, double sg, double &tl,
                                           double &tr) {
    const double ul = u - (x0 + 0.5) * du;
    const double ur = u + (0.5 - x0) * du;
    const bool open_l = !(-ul * sg <= 1e-12);
    const bool open_r = !(ur * sg <= 1e-12);
    tl = -sg;
    tr = -sg;
    if (!(open_l || open_r)) return;
    const double dl = -0.5 - x0, dr = 0.5 - x0;
    if (fabs(du) < 1e-12) {
        if (open_l) tl = dl / u;
        if (open_r) tr = dr / u;
        return;
    }
    const double k = du / u;
    const double t1 = log(1 + k * (open_l ? dl : dr)) / du;
    if (open_l) tl = t1;
    else tr = t1;
    if (open_l && open_r) tr = log(1 + k * dr) / du;  // diverging flow inside the cell: rare
}

// reference _increment, utils.py:777
__device__ __forceinline__ double increment(double t, double u, double du) {
    if (fabs(du) < 1e-12) return u * t;
    return u / du * (exp(du * t) - 1);
}

}  // namespace sdk
