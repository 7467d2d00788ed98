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
#include "fastmath.cuh"

// branch hint: keeps the rarely taken side out of the fall-through path
#define SDK_UNLIKELY(x) __builtin_expect(!!(x), 0)

namespace sdk {

// Grid description as the kernels see it.  Passed as a __grid_constant__ kernel parameter, so the
// face tables sit in the constant bank.
struct GridDev {
    int typ, has_face, nface, ny, nx, gny, gnx, hmode, n_zl, n_z, n_zc;
    signed char nbr_face[SDK_MAX_FACES * 4];
    signed char nbr_edge[SDK_MAX_FACES * 4];
    const float *XC, *YC, *XG, *YG, *dX, *dY, *CS, *SN, *Zl, *dZl, *Z, *dZ;
    // (XG, YG) of every corner point side by side, built by sdk_grid_create: the four corner reads of a crossing touch two
    // 32-byte sectors per row pair instead of four (oceanparcel scheme only, else NULL)
    const float2 *XYG;
    const int32_t *corner_tab;
    // schemes without corner points (SDK_H_RECTILINEAR / SDK_H_LOCAL_CARTESIAN), see sdk_grid_desc
    const float *lon1, *lat1, *coslat32;
    const double *dlon, *dlat;
    int dlon_is_f32;
    // +1 / -1: the vertical table is strictly ascending / descending (looked up by bisection), 0: it is not (linear scan)
    int z_order, zl_order;
};

struct VelDev {
    const void *U, *V, *W, *Vol;
    int dtype, vol_dtype, has_time, it0, nt, has_z, nzU, nzW, nyU, nxU, nyV, nxV;
    int u_stag, v_stag, vol_has_z, transport, has_w;
    int64_t planeU, planeV, planeC;  // elements per (time, level) slab of U, V and of cell-centred fields
};

// fmod is ~70 instructions; it is only needed for |a| >= 360, so keep one copy out of line
static __device__ __noinline__ double fmod360(double a) { return fmod(a, 360.0); }

__device__ __forceinline__ double py_mod360(double a) {
    // python / numpy float "%" with a positive modulus (reference utils.py:201)
    double m = (fabs(a) < 360.0) ? a : fmod360(a);
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

// to_180 for |x| < 720 (differences of two longitudes, one of which may have drifted past the date
// line) in a dozen instructions and without the fmod branch; anything else sets `bad` and the
// caller redoes the stage with to_180.  Step by step equal to the reference's x % 360 % 360 followed
// by x + (-1) * (x // 180) * 360:
//   fmod(x, 360) == x -+ 360 exactly for 360 <= |x| < 720 (Sterbenz);
//   first "%": a negative remainder gets + 360 (rounded), -0.0 becomes +0.0 (x + 0.0 does both);
//   the second "%" only ever turns exactly 360.0 into 0.0, which the last line does as well
//   (360 - 360); the last line adds -360 in the upper half and -0.0 otherwise.
__device__ __forceinline__ double to_180_fast(double x, bool &bad) {
    const double ax = fabs(x);
    bad |= !(ax < 720.0);
    x = x - copysign(ax >= 360.0 ? 360.0 : 0.0, x);
    const double m = x + (x < 0.0 ? 360.0 : 0.0);
    return m + (m >= 180.0 ? -360.0 : -0.0);
}

// to_180 for |x| < 360 -- the difference of two longitudes that both lie within 180 degrees of a third one (the
// corners of a cell, unwrapped around the particle) -- the same steps as to_180_fast without the fmod stage, which
// is the identity there.
__device__ __forceinline__ double to_180_near(double x) {
    const double m = x + (x < 0.0 ? 360.0 : 0.0);
    return m + (m >= 180.0 ? -360.0 : -0.0);
}

// numpy.nan_to_num: nan -> 0, +-inf -> +-DBL_MAX.  Fields and quotients are finite almost always: the
// callers test a whole group of values with integer operations on the exponents and only then come here
// (out of line, so that the step's hot path stays one contiguous run of code).
static __device__ __noinline__ double nan_to_num_cold(double x) {
    if (x != x) return 0.0;
    if (x == INFINITY) return DBL_MAX;
    if (x == -INFINITY) return -DBL_MAX;
    return x;
}

__device__ __forceinline__ bool not_finite(double x) { return (__double2hiint(x) & 0x7ff00000) == 0x7ff00000; }

// inline form for kernels where NaN is ordinary data (land points of a field)
__device__ __forceinline__ double nan_to_num(double x) {
    if (not_finite(x)) {
        if (x != x) return 0.0;
        return x > 0 ? DBL_MAX : -DBL_MAX;
    }
    return x;
}

// a / b, bit for bit, for numerators that are often exactly zero (a particle that has just crossed
// a wall of a lat-lon aligned cell sits exactly on the wall coordinate): CUDA's division takes a
// ~60 instruction slow path for a zero numerator, 1 / b does not.  0 * (1 / b) has the sign and
// the NaN behaviour of 0 / b for every b.
__device__ __forceinline__ double div_maybe_zero(double a, double b) {
    const bool z = a == 0.0;
    const double q = (z ? 1.0 : a) / b;
    return z ? a * q : q;
}

// nan_to_num of four / six values at once: one test in the (finite) common case
__device__ __forceinline__ void nan_to_num4(double &a, double &b, double &c, double &d) {
    if (SDK_UNLIKELY(not_finite(a) | not_finite(b) | not_finite(c) | not_finite(d))) {
        a = nan_to_num_cold(a);
        b = nan_to_num_cold(b);
        c = nan_to_num_cold(c);
        d = nan_to_num_cold(d);
    }
}

__device__ __forceinline__ void nan_to_num6(double &a, double &b, double &c, double &d, double &e, double &f) {
    if (SDK_UNLIKELY(not_finite(a) | not_finite(b) | not_finite(c) | not_finite(d) | not_finite(e) | not_finite(f))) {
        a = nan_to_num_cold(a);
        b = nan_to_num_cold(b);
        c = nan_to_num_cold(c);
        d = nan_to_num_cold(d);
        e = nan_to_num_cold(e);
        f = nan_to_num_cold(f);
    }
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
// the step off the face / off the array (rare): out of line; (face, iy, ix, came_in + 1 | status << 8)
static __device__ __noinline__ int4 cell_move_off_face(const GridDev &g, int dir, int f, int iy, int ix) {
    const int iymax = g.ny - 1, ixmax = g.nx - 1;
    const int ny_ = iy + (dir == 0) - (dir == 1);
    const int nx_ = ix + (dir == 3) - (dir == 2);
    int came_in = -1, st = 0;
    if (g.typ == SDK_TYP_LLC) {
        const int nf = g.nbr_face[f * 4 + dir];
        if (nf < 0) {
            st = SDK_ST_EDGE;  // reference keeps the old index and warns
        } else {
            const int ne = g.nbr_edge[f * 4 + dir];
            const int s = dir >= 2 ? iy : ix;  // coordinate along the crossed edge
            const bool flip = ((dir & 1) == 0) == ((ne & 1) == 0);
            const bool dest_x = ne <= 1;
            const int s_new = flip ? (dest_x ? ixmax : iymax) - s : s;
            f = nf;
            iy = ne == 0 ? iymax : (ne == 1 ? 0 : s_new);
            ix = ne == 2 ? 0 : (ne == 3 ? ixmax : s_new);
            came_in = ne;
        }
    } else if (ny_ < 0 || ny_ > iymax) {
        st = SDK_ST_OUT;
    } else if (g.typ == SDK_TYP_XPERIODIC) {
        iy = ny_;
        ix = nx_ > ixmax ? nx_ - ixmax : ixmax + nx_ + 1;  // sic: reference wraps to 1, not 0
    } else {
        st = SDK_ST_OUT;
    }
    return make_int4(f, iy, ix, (came_in + 1) | (st << 8));
}

__device__ __forceinline__ int cell_move(const GridDev &g, int dir, int &f, int &iy, int &ix,
                                         int &came_in) {
    came_in = -1;
    const int ny_ = iy + (dir == 0) - (dir == 1);
    const int nx_ = ix + (dir == 3) - (dir == 2);
    // one unsigned comparison per coordinate: negative values wrap to huge ones
    const bool off = ((unsigned)ny_ >= (unsigned)g.ny) | ((unsigned)nx_ >= (unsigned)g.nx);
    if (!SDK_UNLIKELY(off)) {
        iy = ny_;
        ix = nx_;
        return 0;
    }
    const int4 r = cell_move_off_face(g, dir, f, iy, ix);
    f = r.x;
    iy = r.y;
    ix = r.z;
    came_in = (r.w & 0xff) - 1;
    return r.w >> 8;
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

// out-of-line copies of plain-operator fallbacks: the hot path only carries the calls
static __device__ __noinline__ double div_cold(double a, double b) { return a / b; }
static __device__ __noinline__ double to_180_cold(double x) { return to_180(x); }
// the root of the quadratic for curved cells (find_rx_ry_oceanparcel utils.py:568)
static __device__ __noinline__ double curved_root_cold(double bb, double det2, double aa) { return (-bb + sqrt(det2)) / (2 * aa); }

// reference find_rx_ry_oceanparcel, utils.py:532 (plain operators; also the slow path of
// inverse_bilinear_fast)
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
    double r = -div_maybe_zero(cc, bb);
    if (order2) r = (-bb + sqrt(det2)) / (2 * aa);
    const double den = a1 + a3 * r;
    double q;
    if (fabs(den) < 1e-12)
        q = ((y - py[0]) / (py[1] - py[0]) + (y - py[3]) / (py[2] - py[3])) * 0.5;
    else
        q = div_maybe_zero(x - a0 - a2 * r, den);
    rx = q - 0.5;
    ry = r - 0.5;
}

// (everything by value: an out-of-line call must not force the caller's arrays into local memory)
static __device__ __noinline__ double2 inverse_bilinear_slow(double x, double y, double px0, double px1, double px2,
                                                             double px3, double py0, double py1, double py2, double py3) {
    const double px[4] = {px0, px1, px2, px3}, py[4] = {py0, py1, py2, py3};
    double2 r;
    inverse_bilinear(x, y, px, py, r.x, r.y);
    return r;
}

// Same arithmetic with the branch-free helpers of fastmath.cuh; `bad` asks for inverse_bilinear_slow.
__device__ __forceinline__ void inverse_bilinear_fast(double x, double y, const double pxi[4], const double py[4],
                                                      double &rx, double &ry, bool &bad) {
    // (pxi[j] = x + to_180(XG - x) lie within 180 degrees of x: every difference below is smaller than 360 in
    // magnitude unless a coordinate is not finite, which the divisions further down flag)
    const double x0 = pxi[0];
    double px[4];
    x = to_180_near(x - x0);
    px[0] = 0.0;  // to_180(x0 - x0)
#pragma unroll
    for (int j = 1; j < 4; ++j) px[j] = to_180_near(pxi[j] - x0);
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
    bool b1_ = false, b2_ = false, b3_ = false;
    double r = -fdiv(cc, bb, b1_);
    if (SDK_UNLIKELY(order2)) r = curved_root_cold(bb, det2, aa);  // curved cells only (polar cap)
    const double den = a1 + a3 * r;
    // Cells whose x axis runs along a meridian (the rotated LLC faces) take the mean of two ratios in y
    // instead.  Both forms go through the same two divisions with the operands selected: no branch, the
    // second quotient is simply not used by the ordinary form.
    const bool mer = fabs(den) < 1e-12;
    const double q1 = fdiv(mer ? y - py[0] : x - a0 - a2 * r, mer ? py[1] - py[0] : den, b2_);
    const double q2 = fdiv(y - py[3], py[2] - py[3], b3_);
    const double q = mer ? (q1 + q2) * 0.5 : q1;
    bad |= (!order2 & b1_) | b2_ | (mer & b3_);
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

// out-of-line copies of the plain-operator fallbacks: the hot path only carries the calls
static __device__ __noinline__ double2 wall_times_cold(double u, double du, double x0, double sg) {
    double tl, tr;
    wall_times(u, du, x0, sg, tl, tr);
    return make_double2(tl, tr);
}

// One copy of the per-axis code, called for x, y and z: the particle step has to stay inside the
// SM's instruction cache (32 KB); with everything inlined three times the loop body was 43 KB and
// instruction fetch the top stall reason (profiles/README.md, r01d).
//
// Exit times of one axis (_stationary_time utils.py:821 + the closed-wall rule of _time2wall :864-869), branch-free for the
// usual case; operands the fast helpers cannot take fall back to wall_times above.
// (x = time to the left wall, y = to the right wall, -sign(tf) = never; double2 comes back in
// registers, a struct of two doubles would go through the stack)
typedef double2 AxisTimes;
static __device__ __noinline__ AxisTimes axis_wall_times(double u, double du, double x0, double sg) {
    double tl, tr;
    const double ul = u - (x0 + 0.5) * du;
    const double ur = u + (0.5 - x0) * du;
    const bool open_l = !(-ul * sg <= 1e-12);
    const bool open_r = !(ur * sg <= 1e-12);
    const bool lin = fabs(du) < 1e-12;
    const double dl = -0.5 - x0, dr = 0.5 - x0;
    bool bk = false, bl = false, bt = false;
    const double k = fdiv(du, u, bk);
    const double inv = rcp_refined(lin ? u : du);
    const double den = lin ? u : du;
    // left wall if it is open, else the right one
    const double d1 = open_l ? dl : dr;
    // A stagnation point between the particle and an open wall (diverging flow, the particle on the other
    // side) makes 1 + k d <= 0: the reference takes the logarithm anyway and gets NaN (-inf for 0), which
    // its argmin reads as "never".  Return the NaN without evaluating anything.
    const double arg1 = lin ? 1.0 : 1 + k * d1;
    const bool never1 = arg1 <= 0.0;
    const double lg1 = flog(never1 ? 1.0 : arg1, bl);
    double t1 = div_with(lin ? d1 : lg1, den, inv, bt);
    t1 = never1 ? __longlong_as_double(0x7ff8000000000000LL) : t1;
    tl = open_l ? t1 : -sg;
    tr = (open_r && !open_l) ? t1 : -sg;
    bool redo = (open_l || open_r) && (bt || (!lin && (bk || bl)));
    if (SDK_UNLIKELY(open_l & open_r)) {  // diverging flow inside the cell: the right wall needs its own logarithm
        const double arg2 = lin ? 1.0 : 1 + k * dr;
        const bool never2 = arg2 <= 0.0;
        const double lg2 = flog(never2 ? 1.0 : arg2, bl);
        tr = div_with(lin ? dr : lg2, den, inv, bt);
        tr = never2 ? __longlong_as_double(0x7ff8000000000000LL) : tr;
        redo = bt || (!lin && (bk || bl));
    }
    if (SDK_UNLIKELY(redo)) return wall_times_cold(u, du, x0, sg);
    return make_double2(tl, tr);
}

// True when none of this axis' walls can be the first event: both are closed, or the time the reference
// would compute for the one reachable wall is certainly later than `tcap` (a directed time that some
// other candidate already achieves), so the logarithm behind it never needs to be evaluated.
//
// Rigorous bound: between the particle and the wall the speed never exceeds umax = max(|u|, |u_wall|)
// (it varies linearly), so the exact travel time is at least d / umax.  The reference's value is
// log(fl(1 + fl(fl(du / u) d))) / du; the roundings inside the logarithm shift its argument u_wall / u by at
// most 1.1e-16 (1 + 2 |du d| / |u_wall| + |du| / |u_wall|) relatively, i.e. the time by less than
// (5e-16 (|u| + |u_wall|) + 2e-16 |du|) / (|u_wall| |du|); what the logarithm and the division add is
// relative (< 1e-15) and sits in the 1e-9 margin.  The reciprocal is a float rounded up: it only has to be an
// upper bound.  In the linear branch (|du| < 1e-12) the reference divides d by u: no absolute term.
__device__ __forceinline__ bool axis_is_late(double u, double du, double x0, double sg, double tcap) {
    const double ul = u - (x0 + 0.5) * du;
    const double ur = u + (0.5 - x0) * du;
    const bool open_l = !(-ul * sg <= 1e-12);
    const bool open_r = !(ur * sg <= 1e-12);
    if (open_l && open_r) return false;  // diverging flow: two candidate walls, leave it to the exact code
    if (!(open_l || open_r)) return true;
    const double uw = open_l ? ul : ur;
    if (!(u * uw > 0)) return false;  // a stagnation point on the way (logarithm of a negative number)
    const double d = fabs(open_l ? -0.5 - x0 : 0.5 - x0);
    const double au = fabs(u), aw = fabs(uw), adu = fabs(du);
    const bool lin = adu < 1e-12;
    const double umax = lin ? au : fmax(au, aw);
    const double rc = (double)__frcp_ru((float)(aw * adu)) * 1.001;  // inf when the product underflows: never "late"
    const double slack = lin ? 0.0 : (5e-16 * (au + aw) + 2e-16 * adu) * umax * rc;
    return d - slack > tcap * umax * (1 + 1e-9);
}

// The same for the usual case only (exactly one reachable wall), to be inlined three times so that
// ptxas can interleave the axes; `redo` asks for axis_wall_times.
__device__ __forceinline__ void wall_times_fast_inline(double u, double du, double x0, double sg, double &tl, double &tr,
                                                       bool &redo) {
    const double ul = u - (x0 + 0.5) * du;
    const double ur = u + (0.5 - x0) * du;
    const bool open_l = !(-ul * sg <= 1e-12);
    const bool open_r = !(ur * sg <= 1e-12);
    const bool lin = fabs(du) < 1e-12;
    const double dsel = open_l ? -0.5 - x0 : 0.5 - x0;
    bool bk = false, bl = false, bt = false;
    const double k = fdiv(du, u, bk);
    const double lg = flog(lin ? 1.0 : 1 + k * dsel, bl);
    const double t1 = fdiv(lin ? dsel : lg, lin ? u : du, bt);
    tl = open_l ? t1 : -sg;
    tr = (open_r && !open_l) ? t1 : -sg;
    redo = redo || (open_l && open_r) || ((open_l || open_r) && (bt || (!lin && (bk || bl))));
}

// reference _increment, utils.py:777
__device__ __forceinline__ double increment(double t, double u, double du) {
    if (fabs(du) < 1e-12) return u * t;
    return u / du * (exp(du * t) - 1);
}

__device__ __forceinline__ double increment_fast_inline(double t, double u, double du, bool &redo) {
    const bool lin = fabs(du) < 1e-12;
    bool bq = false, be = false;
    const double q = fdiv(u, du, bq);
    const double e = fexp(du * t, be);
    redo |= !lin & (bq | be);
    return lin ? u * t : q * (e - 1);
}

static __device__ __noinline__ double increment_cold(double t, double u, double du) { return increment(t, u, du); }

static __device__ __noinline__ double axis_increment(double t, double u, double du) {
    bool redo = false;
    double r = increment_fast_inline(t, u, du, redo);
    if (SDK_UNLIKELY(redo)) r = increment_cold(t, u, du);
    return r;
}

}  // namespace sdk
