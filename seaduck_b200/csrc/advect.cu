// K1: advect particles to a stop time, one analytic in-cell step after the other.
//
// Replaces, for one call of Particle.to_next_stop (reference seaduck/lagrangian.py:741), the whole
// per-iteration numpy pipeline: trim :400 -> analytical_step :561 (time-to-wall, which wall first,
// move) -> lon/lat/dep sync :539 -> cross_cell_wall :706 (re-index across walls and LLC faces,
// re-read corners, invert the bilinear map) -> get_vol :202 -> get_u_du :218 (six wall-velocity
// gathers) -- fused in ONE kernel.
//
// Scheduling: persistent grid (a multiple of the SM count).  Each lane owns one particle in
// registers and steps it until its stop time.  Particles need very different numbers of steps, so
// warps are kept full by refilling idle lanes from a global work queue (warp-aggregated atomicAdd:
// one atomic per refill, ranks by ballot/popc) as soon as `refill_threshold` lanes are idle.
#include <cuda_runtime.h>

#include "advect_args.cuh"
#include "exchange_device.cuh"

namespace sdk {

// Per-particle state of a lane.  The hot half (position, velocity, cell) lives in registers.  The
// cold half -- read or written once or twice per step -- can live in shared memory
// (SDK_ADVECT_COLD_SMEM=1): one column per thread, [field][thread] so that a warp's access is
// conflict free.  That takes the kernel from 128 to 96 / 80 registers (5 / 6 instead of 4 CTAs per
// SM) without the compiler choosing what to spill; the members are references, so the rest of the
// code reads the same either way.  Measured (profiles/README.md, r01i): the extra warps only win
// back what the shared-memory traffic costs (1.63 ms vs 1.59 ms per launch), because the kernel is
// bound by instruction supply, not by latency -- off by default.
#ifndef SDK_ADVECT_COLD_SMEM
#define SDK_ADVECT_COLD_SMEM 0
#endif
// Skip the wall-time logarithm of an axis that provably cannot give the first event (axis_is_late,
// sdk_device.cuh): 0 never, 1 test z, 2 test y and z, 3 test x, y and z.  Parity stays green, but the
// tests cost more than the skipped logarithms save (1.93 / 2.00 / 2.00 ms against 1.61 ms per launch,
// profiles/README.md r01k): lanes of a warp rarely agree on skipping and the extra code pushes the loop
// further past the instruction cache -- off.
#ifndef SDK_ADVECT_DEFER_STORE
#define SDK_ADVECT_DEFER_STORE 1
#endif
// Profiles (bit P) whose steps issue the loads of the next cell as soon as the wall is known (see step()).  Parity
// stays green, but the values in flight cost registers: the surface profile spills 13 doubles instead of 5 and
// configs[3] goes from 3.70 to 4.35 ms per launch (3.72 with 128 registers and four CTAs per SM), profile 1 from 1.10 to
// 1.14 ms (profiles/README.md, round 2, third session) -- off.
#ifndef SDK_ADVECT_HOIST
#define SDK_ADVECT_HOIST 0
#endif
#ifndef SDK_ADVECT_PRUNE
#define SDK_ADVECT_PRUNE 0
#endif
constexpr int kColdDoubles = 12;  // px[4], py[4], bzl, dzl, vol, pid
[[maybe_unused]] constexpr int kColdInts = 7;  // dx, dy (float bits), it, status, niter, ncross, nrec

#if SDK_ADVECT_COLD_SMEM
struct ColdArr {
    double *b;
    int stride;
    __device__ __forceinline__ double &operator[](int j) const { return b[j * stride]; }
};

struct Lane {
    double lon, lat, dep, t;
    double rx, ry, rzl;
    double u, v, w, du, dv, dw;
    int f, iy, ix, izl;
    ColdArr px, py;
    double &bzl, &dzl, &vol;
    float &dx, &dy;
    int &it, &status, &niter, &ncross, &nrec;
    long long &pid;
    // cd / ci: this thread's column of the cold doubles / ints, consecutive fields `stride` apart
    __device__ __forceinline__ Lane(double *cd, int *ci, int stride)
        : px{cd, stride}, py{cd + 4 * stride, stride}, bzl(cd[8 * stride]), dzl(cd[9 * stride]), vol(cd[10 * stride]),
          dx(reinterpret_cast<float &>(ci[0])), dy(reinterpret_cast<float &>(ci[stride])), it(ci[2 * stride]),
          status(ci[3 * stride]), niter(ci[4 * stride]), ncross(ci[5 * stride]), nrec(ci[6 * stride]),
          pid(reinterpret_cast<long long &>(cd[11 * stride])) {}
};
#define SDK_LANE(name, cd, ci, stride) Lane name(cd, ci, stride)
#else
struct Lane {
    double lon, lat, dep, t;
    double rx, ry, rzl;
    double u, v, w, du, dv, dw;
    double px[4], py[4];
    double bzl, dzl, vol;
    float dx, dy;
    int f, iy, ix, izl, it;
    int status, niter, ncross, nrec;
    long long pid;
};
#define SDK_LANE(name, cd, ci, stride) Lane name
#endif

// bytes of dynamic shared memory of a CTA: the two vertical tables + the cold columns
__host__ __device__ inline size_t advect_smem_bytes(int n_zl, int n_z, int threads) {
    size_t b = (sizeof(float) * (size_t)(n_zl + n_z) + 15) & ~(size_t)15;
#if SDK_ADVECT_COLD_SMEM
    b += (size_t)threads * (kColdDoubles * sizeof(double) + kColdInts * sizeof(int));
#endif
    return b;
}

// What is known about a launch at compile time.  Profile 0 reads every switch from the arguments.  Profile 1
// is the production case of the north star -- LLC faces, particles with a depth, vertical velocity, volume
// transports, level-dependent volumes -- with those switches folded: fewer instructions per step and, what
// matters more for a loop that overflows the instruction cache, less code in the loop body.
#ifndef SDK_ADVECT_PROFILES
#define SDK_ADVECT_PROFILES 1
#endif
// Profile 2 is profile 0 for the horizontal schemes without corner points (local_cartesian, rectilinear:
// reference lagrangian.py:647-673, :693-704 and rel2latlon utils.py:214): px / py carry bx, cs, dx / by, sn, dy.
// Profile 3 is the other production case, the surface run of the LLC4320 notebook (BASELINE.json configs[3]): LLC faces,
// volume transports, no depth, no vertical velocity, fields and volumes without a vertical axis.
#define SDK_KNOWN(P, fixed_value, runtime_expr) ((P) == 1 ? (fixed_value) : (runtime_expr))
#define SDK_KNOWN3(P, value_1, value_3, runtime_expr) ((P) == 1 ? (value_1) : ((P) == 3 ? (value_3) : (runtime_expr)))

// sic: lagrangian.py:22, utils.py:306 (6271 km)
#define SDK_DEG2M (6271e3 * 3.141592653589793 / 180)

template <int P = 0>
__device__ __forceinline__ void load_lane(const sdk_particles &p, int64_t i, int has_dep, Lane &s) {
    s.pid = i;
    s.lon = p.lon[i];
    s.lat = p.lat[i];
    s.t = p.t[i];
    s.rx = p.rx[i];
    s.ry = p.ry[i];
    s.u = p.u[i];
    s.v = p.v[i];
    s.du = p.du[i];
    s.dv = p.dv[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s.px[j] = p.px[j * p.n + i];
        s.py[j] = p.py[j * p.n + i];
    }
    s.f = SDK_KNOWN3(P, true, true, p.face != nullptr) ? p.face[i] : 0;
    s.iy = p.iy[i];
    s.ix = p.ix[i];
    s.it = p.it ? p.it[i] : 0;
    // (with transports everything is scaled by the cell volume, dx and dy are never read)
    s.dx = SDK_KNOWN3(P, false, false, p.dx != nullptr) ? p.dx[i] : 1.0f;
    s.dy = SDK_KNOWN3(P, false, false, p.dy != nullptr) ? p.dy[i] : 1.0f;
    s.vol = SDK_KNOWN3(P, true, true, p.vol != nullptr) ? p.vol[i] : 1.0;
    if (SDK_KNOWN3(P, true, false, has_dep != 0)) {
        s.dep = p.dep[i];
        s.rzl = p.rzl[i];
        s.w = p.w[i];
        s.dw = p.dw[i];
        s.bzl = p.bzl[i];
        s.dzl = p.dzl[i];
        s.izl = p.izl[i];
    } else {
        s.dep = p.dep ? p.dep[i] : 0.0;
        s.rzl = 0.0;
        s.w = 0.0;
        s.dw = 0.0;
        s.bzl = 0.0;
        s.dzl = 1.0;
        s.izl = 0;
    }
    s.status = 0;
    s.niter = 0;
    s.ncross = 0;
    s.nrec = 0;
}

template <int P = 0>
__device__ __forceinline__ void store_lane(const sdk_particles &p, int has_dep, const Lane &s) {
    const long long i = s.pid;
    p.lon[i] = s.lon;
    p.lat[i] = s.lat;
    p.t[i] = s.t;
    p.rx[i] = s.rx;
    p.ry[i] = s.ry;
    p.u[i] = s.u;
    p.v[i] = s.v;
    p.du[i] = s.du;
    p.dv[i] = s.dv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        p.px[j * p.n + i] = s.px[j];
        p.py[j * p.n + i] = s.py[j];
    }
    if (SDK_KNOWN3(P, true, true, p.face != nullptr)) p.face[i] = s.f;
    p.iy[i] = s.iy;
    p.ix[i] = s.ix;
    if (SDK_KNOWN3(P, true, true, p.vol != nullptr)) p.vol[i] = s.vol;
    if (SDK_KNOWN3(P, true, false, has_dep != 0)) {
        p.dep[i] = s.dep;
        p.rzl[i] = s.rzl;
        p.w[i] = s.w;
        p.dw[i] = s.dw;
        p.bzl[i] = s.bzl;
        p.dzl[i] = s.dzl;
        p.izl[i] = s.izl;
    }
}

__device__ __forceinline__ void note(const AdvectArgs &a, Lane &s, int stamp) {
    const sdk_log &L = a.log;
    if (L.offset) {
        const int64_t r = L.offset[s.pid] + s.nrec;
        if (r < L.capacity) {
            // one 128-byte record = eight 16-byte stores to consecutive addresses
            double2 *dst = reinterpret_cast<double2 *>(L.records + r);
            dst[0] = make_double2(s.rzl, s.rx);
            dst[1] = make_double2(s.ry, s.t);
            dst[2] = make_double2(s.u, s.v);
            dst[3] = make_double2(s.w, s.du);
            dst[4] = make_double2(s.dv, s.dw);
            dst[5] = make_double2(s.lon, s.lat);
            int4 *idst = reinterpret_cast<int4 *>(dst + 6);
            idst[0] = make_int4(__double2loint(s.dep), __double2hiint(s.dep), s.f, s.it);
            idst[1] = make_int4(s.izl, s.iy, s.ix, stamp);
        }
    }
    s.nrec += 1;
}

// reference Particle.trim, lagrangian.py:400 (one axis)
__device__ __forceinline__ void trim_axis(double &x, double &u, double du, double lo, double hi) {
    if (x >= hi) {
        const double d = hi - x;
        x += d;
        u += du * d;
    }
    if (x <= lo) {
        const double d = lo - x;
        x += d;
        u += du * d;
    }
}

// the wall values of a cell as they come out of memory (velocity_loads), before the weights are applied (velocity_finish)
struct VelRaw {
    double ul, ur, vl, vr, w0, w1, vol;
};

// First half of read_velocity: where the six wall values and the volume of cell (f, iy, ix) at level izl and time index it
// live, and the loads.  Needs nothing but the cell, so a step can issue it as soon as it knows which wall the particle
// leaves through, long before the relative coordinates in the new cell exist.
template <typename T, int P = 0>
__device__ __forceinline__ void velocity_loads(const GridDev &g, const VelDev &v, int f, int iy, int ix, int izl, int it, int &status, VelRaw &r) {
    const bool k_face = SDK_KNOWN3(P, true, true, g.has_face != 0);
    const bool k_w = SDK_KNOWN3(P, true, false, v.has_w != 0);
    const bool k_tr = SDK_KNOWN3(P, true, true, v.transport != 0);
    const bool k_z = SDK_KNOWN3(P, true, false, v.has_z != 0);
    const bool k_volz = SDK_KNOWN3(P, true, false, v.vol_has_z != 0);
    // ---- where do the right and top wall values live?
    int rf = f, ry_ = iy, rx_ = ix, r_in;
    int tf = f, ty_ = iy, tx_ = ix, t_in;
    bool r_is_v = false, t_is_u = false;
    if (k_face) {
        int st = cell_move(g, 3, rf, ry_, rx_, r_in);
        if (st) {
            status |= SDK_ST_BAD_CELL;
            rf = f, ry_ = iy, rx_ = ix;
        } else if (r_in >= 0) {
            if (r_in == 1) r_is_v = true;
            else if (r_in != 2) status |= SDK_ST_BAD_CELL;
        }
        st = cell_move(g, 0, tf, ty_, tx_, t_in);
        if (st) {
            status |= SDK_ST_BAD_CELL;
            tf = f, ty_ = iy, tx_ = ix;
        } else if (t_in >= 0) {
            if (t_in == 2) t_is_u = true;
            else if (t_in != 1) status |= SDK_ST_BAD_CELL;
        }
    } else {
        // scalar kernels on a single face: naive neighbour, topology walk only off the array
        // (an index of -1, -1 from a closed box reads the last element, like numpy does)
        const int xlim = v.u_stag ? g.gnx : g.nx, ylim = v.v_stag ? g.gny : g.ny;
        rx_ = ix + 1;
        if (rx_ > xlim - 1) {
            if (g.typ == SDK_TYP_XPERIODIC) rx_ = rx_ - (g.nx - 1);
            else { ry_ = v.nyU - 1; rx_ = v.nxU - 1; }
        }
        ty_ = iy + 1;
        if (ty_ > ylim - 1) { ty_ = v.nyV - 1; tx_ = v.nxV - 1; }
    }
    int tlev = v.has_time ? wrap_index(it - v.it0, v.nt) : 0;
    if (tlev < 0 || tlev >= v.nt) {
        // the particle's time level is not in the staged slab (the reference raises IndexError here:
        // Particle.update_uvw_array was not called after crossing time_midp); stay inside the array
        status |= SDK_ST_OOB_T;
        tlev = tlev < 0 ? 0 : v.nt - 1;
    }
    const int64_t lev_uv = k_z ? (int64_t)(tlev * v.nzU + wrap_index(izl - 1, v.nzU)) : (int64_t)tlev;
    const int c_ul = (f * v.nyU + iy) * v.nxU + ix;
    const int c_vl = (f * v.nyV + iy) * v.nxV + ix;
    const int c_ur = r_is_v ? (rf * v.nyV + ry_) * v.nxV + rx_ : (rf * v.nyU + ry_) * v.nxU + rx_;
    const int c_vr = t_is_u ? (tf * v.nyU + ty_) * v.nxU + tx_ : (tf * v.nyV + ty_) * v.nxV + tx_;
    // issue every load before using any of them
    r.ul = ldf<T>(v.U, lev_uv * v.planeU + c_ul);
    r.vl = ldf<T>(v.V, lev_uv * v.planeV + c_vl);
    r.ur = ldf<T>(r_is_v ? v.V : v.U, lev_uv * (r_is_v ? v.planeV : v.planeU) + c_ur);
    r.vr = ldf<T>(t_is_u ? v.U : v.V, lev_uv * (t_is_u ? v.planeU : v.planeV) + c_vr);
    const int cell = (f * g.ny + iy) * g.nx + ix;
    r.w0 = 0.0, r.w1 = 0.0, r.vol = 1.0;
    if (k_w) {
        const int64_t wl = (int64_t)tlev * v.nzW;
        r.w0 = ldf<T>(v.W, (wl + wrap_index(izl, v.nzW)) * v.planeC + cell);
        r.w1 = ldf<T>(v.W, (wl + wrap_index(izl - 1, v.nzW)) * v.planeC + cell);
    }
    if (k_tr) {
        const int64_t idx = (k_volz ? (int64_t)wrap_index(izl - 1, g.n_z > 0 ? g.n_z : 1) : 0) * v.planeC + cell;
        r.vol = v.vol_dtype == SDK_F64 ? ldf<double>(v.Vol, idx) : ldf<float>(v.Vol, idx);
    }
}

// Second half: the two-node kernels in the particle's relative coordinates, and the division by the cell's scale.
template <typename T, int P = 0>
__device__ __forceinline__ void velocity_finish(const GridDev &g, const VelDev &v, Lane &s, const VelRaw &r) {
    const bool k_face = SDK_KNOWN3(P, true, true, g.has_face != 0);
    const bool k_w = SDK_KNOWN3(P, true, false, v.has_w != 0);
    const bool k_tr = SDK_KNOWN3(P, true, true, v.transport != 0);
    if (k_tr) s.vol = r.vol;
    double ul_ = r.ul, ur_ = r.ur, vl_ = r.vl, vr_ = r.vr, w0_ = r.w0, w1_ = r.w1;
    nan_to_num6(ul_, ur_, vl_, vr_, w0_, w1_);
    const bool stag_x = k_face || v.u_stag, stag_y = k_face || v.v_stag;
    const double rxs = stag_x ? s.rx + 0.5 : s.rx;
    const double rys = stag_y ? s.ry + 0.5 : s.ry;
    // Lagrange weights of the two-node kernels (reference kernel_weight.py:211-229)
    // (r - 1) / (0 - 1) is an exact negation
    const double wx0 = -(rxs - 1.0), wx1 = rxs;
    const double wy0 = -(rys - 1.0), wy1 = rys;
    double u = ul_ * wx0 + ur_ * wx1;
    double du = ul_ * -1.0 + ur_;
    double vv = vl_ * wy0 + vr_ * wy1;
    double dv = vl_ * -1.0 + vr_;
    // (schemes without corners keep dx, dy as float64 in px[2], py[2]: the rectilinear ones are not float32 values)
    const double sx = k_tr ? s.vol : (P == 2 ? s.px[2] : (double)s.dx);
    const double sy = k_tr ? s.vol : (P == 2 ? s.py[2] : (double)s.dy);
    const double sz = k_tr ? s.vol : s.dzl;
    double w = 0.0, dw = 0.0;
    if (k_w) {
        w = w0_ * (1 - s.rzl) + w1_ * s.rzl;
        dw = w0_ * -1.0 + w1_;
    }
    // six IEEE divisions by at most three denominators (one, the cell volume, for transports): the
    // refined reciprocal is shared, each quotient costs three more operations (fastmath.cuh)
    bool redo = false;
    const double yx = rcp_refined(sx);
    const double yy = k_tr ? yx : rcp_refined(sy);
    const double yz = k_tr ? yx : rcp_refined(sz);
    double qu = div_with(u, sx, yx, redo), qdu = div_with(du, sx, yx, redo);
    double qv = div_with(vv, sy, yy, redo), qdv = div_with(dv, sy, yy, redo);
    double qw = 0.0, qdw = 0.0;
    if (k_w) {
        qw = div_with(w, sz, yz, redo);
        qdw = div_with(dw, sz, yz, redo);
    }
    if (SDK_UNLIKELY(redo)) {  // zero / denormal / non-finite scale: plain operators (out of line)
        qu = div_cold(u, sx), qdu = div_cold(du, sx), qv = div_cold(vv, sy), qdv = div_cold(dv, sy);
        if (k_w) qw = div_cold(w, sz), qdw = div_cold(dw, sz);
        // (a quotient the fast path vouches for is finite: nan_to_num has something to do only here)
        nan_to_num6(qu, qv, qdu, qdv, qw, qdw);
    }
    s.u = qu;
    s.v = qv;
    s.du = qdu;
    s.dv = qdv;
    if (k_w) {
        s.w = qw;
        s.dw = qdw;
    }
}

// Gather the wall values and form u, du, v, dv, w, dw (reference get_u_du lagrangian.py:218 with the
// kernels of lagrangian.py:25-53; on LLC grids the right / top wall may live on the neighbouring
// face and be the other velocity component, reference eulerian.py:887 + topology.py:618-647).
// Index arithmetic: 32-bit within a horizontal plane, one 64-bit multiply-add for the level/time.
template <typename T, int P = 0>
__device__ __forceinline__ void read_velocity(const GridDev &g, const VelDev &v, int has_dep, Lane &s) {
    (void)has_dep;
    VelRaw r;
    velocity_loads<T, P>(g, v, s.f, s.iy, s.ix, s.izl, s.it, s.status, r);
    velocity_finish<T, P>(g, v, s, r);
}

// One iteration of the reference loop (lagrangian.py:766-791) for one particle.
template <typename T, bool LOG, int P>
__device__ __forceinline__ void step(const AdvectArgs &a, const float *sZl, const float *sdZl, Lane &s,
                                     bool &live) {
    const GridDev &g = a.g;
    const int has_dep = SDK_KNOWN3(P, 1, 0, a.has_dep);
    // trim, lagrangian.py:400
    const double lo = -0.5 + a.trim_tol, hi = 0.5 - a.trim_tol;
    trim_axis(s.rx, s.u, s.du, lo, hi);
    trim_axis(s.ry, s.v, s.dv, lo, hi);
    if (has_dep) trim_axis(s.rzl, s.w, s.dw, -0.0 + a.trim_tol, 1.0 - a.trim_tol);
    // analytical_step, lagrangian.py:561: exit times, one axis at a time (one copy of the code)
    const double tf = a.t_stop - s.t;
    const double sg = sgn(tf);
    const double z0 = has_dep ? s.rzl - 0.5 : 0.0;
    double ts[7];
#if SDK_ADVECT_INLINE_AXES
    {
        // three independent straight-line copies that ptxas interleaves; rare cases out of line
        bool redo_x = false, redo_y = false, redo_z = false;
        wall_times_fast_inline(s.u, s.du, s.rx, sg, ts[0], ts[1], redo_x);
        wall_times_fast_inline(s.v, s.dv, s.ry, sg, ts[2], ts[3], redo_y);
        if (has_dep) {
            wall_times_fast_inline(s.w, s.dw, z0, sg, ts[4], ts[5], redo_z);
        } else {
            ts[4] = -sg;
            ts[5] = -sg;
        }
        if (redo_x) {
            const AxisTimes t = axis_wall_times(s.u, s.du, s.rx, sg);
            ts[0] = t.x, ts[1] = t.y;
        }
        if (redo_y) {
            const AxisTimes t = axis_wall_times(s.v, s.dv, s.ry, sg);
            ts[2] = t.x, ts[3] = t.y;
        }
        if (redo_z) {
            const AxisTimes t = axis_wall_times(s.w, s.dw, z0, sg);
            ts[4] = t.x, ts[5] = t.y;
        }
    }
#else
    {
        // An axis whose wall is provably reached later than an event already found keeps "never" (-sign(tf)),
        // which the argmin below treats like any later time: its logarithm is not evaluated (axis_is_late).
        double cap = tf * sg;  // directed time of the earliest event so far; the stop itself to begin with
        auto earliest = [&](const AxisTimes &t) {
            const double a = t.x * sg, b = t.y * sg;
            if (a >= 0) cap = fmin(cap, a);
            if (b >= 0) cap = fmin(cap, b);
        };
        ts[0] = ts[1] = ts[2] = ts[3] = ts[4] = ts[5] = -sg;
        if (SDK_ADVECT_PRUNE < 3 || !axis_is_late(s.u, s.du, s.rx, sg, cap)) {
            const AxisTimes tx = axis_wall_times(s.u, s.du, s.rx, sg);
            ts[0] = tx.x, ts[1] = tx.y;
            earliest(tx);
        }
        if (SDK_ADVECT_PRUNE < 2 || !axis_is_late(s.v, s.dv, s.ry, sg, cap)) {
            const AxisTimes ty = axis_wall_times(s.v, s.dv, s.ry, sg);
            ts[2] = ty.x, ts[3] = ty.y;
            earliest(ty);
        }
        if (has_dep && (SDK_ADVECT_PRUNE < 1 || !axis_is_late(s.w, s.dw, z0, sg, cap))) {
            const AxisTimes tz = axis_wall_times(s.w, s.dw, z0, sg);
            ts[4] = tz.x, ts[5] = tz.y;
        }
    }
#endif
    ts[6] = tf;
    // _which_early, utils.py:875: first minimum of the directed times, nan / negative = never
    int tend = 0;
    double best = INFINITY;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if (P == 3 && (k == 4 || k == 5)) continue;  // no vertical axis: "never" (a live particle has sg != 0)
        double d = ts[k] * sg;
        if (isnan(d) || d < 0) d = INFINITY;
        if (d < best) {
            best = d;
            tend = k;
        }
    }
    double te = ts[0];
#pragma unroll
    for (int k = 1; k < 7; ++k)
        if (!(P == 3 && (k == 4 || k == 5)) && tend == k) te = ts[k];
    // Which wall is known: so is the cell the particle is about to enter (_cross_cell_wall_index, lagrangian.py:598), and
    // with it every address the rest of the step reads from -- corners, wall values, volume.  With HOIST the loads go
    // out here, and their latency passes behind the move, the logarithm-free but long inverse bilinear map and the
    // vertical lookup instead of being waited for twice (corners, then velocities) at the end of the step.
    constexpr bool HOIST = P != 2 && ((SDK_ADVECT_HOIST >> P) & 1);
    int nf = s.f, niy = s.iy, nix = s.ix, nizl = s.izl, st_move = 0, st_early = 0;
    bool leaves = false;  // out of a closed domain / through the deepest level: nothing to read
    int64_t c[4];
    float gx[4], gy[4];
    VelRaw raw;
    if (HOIST) {
        if (tend <= 3) {
            const int dir = tend == 0 ? 2 : (tend == 1 ? 3 : (tend == 2 ? 1 : 0));
            int came_in;
            st_move = cell_move(g, dir, nf, niy, nix, came_in);
            leaves = (st_move & SDK_ST_OUT) != 0;
        } else if (P != 3 && tend == 4) {
            nizl += 1;
            leaves = nizl > g.n_zl - 1;
        } else if (P != 3 && tend == 5) {
            if (!a.kick_back || s.izl != 1) nizl -= 1;
        }
        if (!leaves) {
            st_early |= corner_indices(g, nf, niy, nix, c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 q = __ldg(g.XYG + c[j]);
                gx[j] = q.x;
                gy[j] = q.y;
            }
            velocity_loads<T, P>(g, a.v, nf, niy, nix, nizl, s.it, st_early, raw);
        }
    }
    // _move_within_cell, lagrangian.py:528
    s.t += te;
    {
#if SDK_ADVECT_INLINE_AXES
        bool redo_x = false, redo_y = false, redo_z = false;
        double ix_ = increment_fast_inline(te, s.u, s.du, redo_x);
        double iy_ = increment_fast_inline(te, s.v, s.dv, redo_y);
        double iz_ = has_dep ? increment_fast_inline(te, s.w, s.dw, redo_z) : 0.0;
        if (redo_x) ix_ = axis_increment(te, s.u, s.du);
        if (redo_y) iy_ = axis_increment(te, s.v, s.dv);
        if (redo_z) iz_ = axis_increment(te, s.w, s.dw);
#else
        const double ix_ = axis_increment(te, s.u, s.du);
        const double iy_ = axis_increment(te, s.v, s.dv);
        const double iz_ = has_dep ? axis_increment(te, s.w, s.dw) : 0.0;
#endif
        s.u = s.u + s.du * ix_;
        s.v = s.v + s.dv * iy_;
        s.rx = ix_ + s.rx;
        s.ry = iy_ + s.ry;
        if (has_dep) {
            s.w = s.w + s.dw * iz_;
            s.rzl = (iz_ + z0) + 0.5;
        }
    }
    // _sync_latlondep_before_cross, lagrangian.py:539
    if (has_dep) s.dep = s.bzl + s.dzl * s.rzl;
    if (P == 2) {
        // rel2latlon, utils.py:214 (no px / py on the particle: lagrangian.py:549).  numba compiles it, so the
        // cosine of the float32 latitude is taken in double precision.
        const double bx = s.px[0], cs = s.px[1], dxm = s.px[2], by = s.py[0], sn = s.py[1], dym = s.py[2];
        const double temp_x = s.rx * dxm / SDK_DEG2M;
        const double temp_y = s.ry * dym / SDK_DEG2M;
        const double dlon = (temp_x * cs - temp_y * sn) / cos(by * 3.141592653589793 / 180);
        const double dlat = temp_x * sn + temp_y * cs;
        s.lon = dlon + bx;
        s.lat = dlat + by;
    } else {
        const double w0 = (0.5 - s.rx) * (0.5 - s.ry), w1 = (0.5 + s.rx) * (0.5 - s.ry);
        const double w2 = (0.5 + s.rx) * (0.5 + s.ry), w3 = (0.5 - s.rx) * (0.5 + s.ry);
        s.lon = ((w0 * s.px[0] + w1 * s.px[1]) + w2 * s.px[2]) + w3 * s.px[3];
        s.lat = ((w0 * s.py[0] + w1 * s.py[1]) + w2 * s.py[2]) + w3 * s.py[3];
    }
    if (LOG) note(a, s, tend);
    s.ncross += tend < 6;
    // _cross_cell_wall_index, lagrangian.py:598
    if (HOIST) {
        s.status |= st_move | st_early;
        if (SDK_UNLIKELY(leaves)) {
            if (tend == 4) {
                s.izl = g.n_zl - 1;
                s.status |= SDK_ST_OOB_LEVEL;
            }
            live = false;
            return;
        }
        if (P != 3 && tend == 5 && a.kick_back && s.izl == 1) s.dep = s.bzl + s.dzl / 2;
        s.f = nf, s.iy = niy, s.ix = nix, s.izl = nizl;
    } else if (tend <= 3) {
        const int dir = tend == 0 ? 2 : (tend == 1 ? 3 : (tend == 2 ? 1 : 0));
        int came_in;
        const int st = cell_move(g, dir, s.f, s.iy, s.ix, came_in);
        s.status |= st;
        if (st & SDK_ST_OUT) {
            live = false;  // nothing sensible to read outside a closed domain
            return;
        }
    } else if (P != 3 && tend == 4) {
        s.izl += 1;
        if (SDK_UNLIKELY(s.izl > g.n_zl - 1)) {
            // through the deepest staged level (a non-zero vertical velocity at the bottom): the reference
            // raises IndexError at Zl[izl_lin], lagrangian.py:675; freeze the particle on that wall
            s.izl = g.n_zl - 1;
            s.status |= SDK_ST_OOB_LEVEL;
            live = false;
            return;
        }
    } else if (P != 3 && tend == 5) {
        if (!a.kick_back || s.izl != 1) s.izl -= 1;
        else s.dep = s.bzl + s.dzl / 2;
    }
    // _cross_cell_wall_read, lagrangian.py:640: corners (oceanparcel) or centre + metric (other schemes), vertical levels
    float l_cs = 1.f, l_sn = 0.f, l_dx = 1.f, l_dy = 1.f, l_cos = 1.f, l_bx = 0.f, l_by = 0.f;
    double l_dlon = 1.0, l_dlat = 1.0;
    if (P == 2) {
        if (g.hmode == SDK_H_LOCAL_CARTESIAN) {  // :647-659
            const int64_t h = ((int64_t)s.f * g.ny + wrap_index(s.iy, g.ny)) * g.nx + wrap_index(s.ix, g.nx);
            l_bx = __ldg(g.XC + h), l_by = __ldg(g.YC + h);
            l_cs = __ldg(g.CS + h), l_sn = __ldg(g.SN + h);
            l_dx = __ldg(g.dX + h), l_dy = __ldg(g.dY + h);
            l_cos = __ldg(g.coslat32 + h);
        } else {  // rectilinear, :667-672
            const int ix = wrap_index(s.ix, g.nx), iy = wrap_index(s.iy, g.ny);
            l_bx = __ldg(g.lon1 + ix), l_by = __ldg(g.lat1 + iy);
            l_cos = __ldg(g.coslat32 + iy);
            l_dlon = __ldg(g.dlon + ix), l_dlat = __ldg(g.dlat + iy);
        }
    } else if (!HOIST) {
        s.status |= corner_indices(g, s.f, s.iy, s.ix, c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 q = __ldg(g.XYG + c[j]);
            gx[j] = q.x;
            gy[j] = q.y;
        }
    }
    if (has_dep) {
        s.bzl = (double)sZl[wrap_index(s.izl, g.n_zl)];
        s.dzl = (double)sdZl[wrap_index(s.izl - 1, g.n_z)];
        // _cross_cell_wall_rel, lagrangian.py:681
        s.rzl = div_maybe_zero(s.dep - s.bzl, s.dzl);
    }
    if (P == 2) {
        double dxm, dym;
        if (g.hmode == SDK_H_LOCAL_CARTESIAN) {
            dxm = (double)l_dx, dym = (double)l_dy;
        } else {
            // dlon[ix] * np.cos(by * np.pi / 180): float64 spacing times the float32 cosine (a float32 product
            // when the dataset's lon is float32 itself)
            dxm = g.dlon_is_f32 ? (double)((float)l_dlon * l_cos) : l_dlon * (double)l_cos;
            dym = l_dlat;
        }
        const double bx = (double)l_bx, by = (double)l_by, cs = (double)l_cs, sn = (double)l_sn, cb = (double)l_cos;
        s.px[0] = bx, s.px[1] = cs, s.px[2] = dxm;
        s.py[0] = by, s.py[1] = sn, s.py[2] = dym;
        // _cross_cell_wall_rel, lagrangian.py:693-703 (plain numpy there: by is the float32 array just read,
        // so its cosine is numpy's float32 one -- the host-made table)
        const double dlon = to_180(s.lon - bx);
        const double dlat = to_180(s.lat - by);
        s.rx = (dlon * cb * cs + dlat * sn) * SDK_DEG2M / dxm;
        s.ry = (dlat * cs - dlon * sn * cb) * SDK_DEG2M / dym;
    } else {
        bool redo = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s.px[j] = s.lon + to_180_fast((double)gx[j] - s.lon, redo);  // eulerian.py:618
            s.py[j] = (double)gy[j];
        }
        if (SDK_UNLIKELY(redo)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s.px[j] = s.lon + to_180_cold((double)gx[j] - s.lon);
        }
        redo = false;
        const double qx[4] = {s.px[0], s.px[1], s.px[2], s.px[3]}, qy[4] = {s.py[0], s.py[1], s.py[2], s.py[3]};
        inverse_bilinear_fast(s.lon, s.lat, qx, qy, s.rx, s.ry, redo);
        if (SDK_UNLIKELY(redo)) {
            const double2 r = inverse_bilinear_slow(s.lon, s.lat, s.px[0], s.px[1], s.px[2], s.px[3], s.py[0], s.py[1], s.py[2], s.py[3]);
            s.rx = r.x, s.ry = r.y;
        }
    }
    // get_vol + get_u_du
    if (HOIST) velocity_finish<T, P>(g, a.v, s, raw);
    else read_velocity<T, P>(g, a.v, has_dep, s);
    s.niter += 1;
    live = fabs(a.t_stop - s.t) > a.tol;
    if (LOG && live && !(a.log.continue_mode & 1)) note(a, s, 7);
}

// Tuning knobs (profiles/README.md, round 1 iterations d-g).  The particle step is ~2000 SASS
// instructions; 16 resident warps walking through it independently overflow the SM's 32 KB
// instruction cache and saturate the GPC-level instruction cache behind it.  LOCKSTEP makes the
// warps of one 512-thread CTA start every step together (one barrier per step) so that they share
// the fetched lines (hit rate 79 % -> 95 %, but the barrier wait costs what the misses did);
// INLINE_AXES picks between three interleaved copies of the per-axis code and one out-of-line copy
// (smaller loop body).  Measured on the bench workload: all four combinations within 10 %, the
// defaults below are the fastest.
#ifndef SDK_ADVECT_LOCKSTEP
#define SDK_ADVECT_LOCKSTEP 0
#endif
// compiled for CTAs of up to 512 threads (128 registers either way); launched as 4 x 128 threads per
// SM by default, as 1 x 512 when SMs are to be left free for a concurrent collective (reserve_sms)
#ifndef SDK_ADVECT_THREADS
#define SDK_ADVECT_THREADS 512
#endif
#ifndef SDK_ADVECT_MIN_BLOCKS
#define SDK_ADVECT_MIN_BLOCKS 1
#endif
#ifndef SDK_ADVECT_DEFAULT_THREADS
#define SDK_ADVECT_DEFAULT_THREADS 128
#endif
#ifndef SDK_ADVECT_INLINE_AXES
#define SDK_ADVECT_INLINE_AXES 0
#endif

// one out-of-line copy of the record store of the fused exchange: it runs once per particle, the step loop must not
// carry its code (the kernel is bound by the instruction lines it has to fetch, profiles/README.md)
static __device__ __noinline__ void ship_lane_cold(const AdvectArgs &a, long long base, long long pid, double lon, double lat, double dep,
                                                   int f, int iy, int ix, int izl) {
    const unsigned long long cell = SDK_CELL_PACK((uint64_t)(uint32_t)f, (uint64_t)(uint32_t)iy, (uint64_t)(uint32_t)ix, (uint64_t)(uint32_t)izl);
    const long long at = base + (a.ship_index ? (long long)a.ship_index[pid] : pid);
    ship_record(a.x, at, make_double2(lon, lat), make_double2(dep, __longlong_as_double((long long)cell)));
}

// SHIP: the write-back of a finished particle also stores its 32-byte output record into every rank's receive
// buffer (sdk_advect_params.ship): the exchange of a stop's output rides on the kernel that produces it.
// Launch bounds per profile.  The 3-D profiles need their 128 registers (4 warps per scheduler; with 96 or 80 the spills
// and the instruction cache cost more than the extra warps bring, profiles/README.md).  The surface profile carries
// neither the vertical axis nor its code: it fits 96 registers with five doubles spilled and gains from the fifth CTA
// per SM (configs[3]: 4.05 -> 3.78 ms per launch).
#ifndef SDK_ADVECT_P3_BLOCKS
#define SDK_ADVECT_P3_BLOCKS 5
#endif
template <int P>
struct Bounds {
    static constexpr int threads = SDK_ADVECT_THREADS, blocks = SDK_ADVECT_MIN_BLOCKS;
};
template <>
struct Bounds<3> {
    static constexpr int threads = SDK_ADVECT_DEFAULT_THREADS, blocks = SDK_ADVECT_P3_BLOCKS;
};

template <typename T, bool LOG, int P, bool SHIP = false>
__global__ void __launch_bounds__(Bounds<P>::threads, Bounds<P>::blocks)
advect_kernel(const __grid_constant__ AdvectArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sZl = smem;
    float *sdZl = smem + a.g.n_zl;
    for (int k = threadIdx.x; k < a.g.n_zl; k += blockDim.x) sZl[k] = a.g.Zl[k];
    for (int k = threadIdx.x; k < a.g.n_z; k += blockDim.x) sdZl[k] = a.g.dZl[k];
    __shared__ int ship_ok;
    if (SHIP && threadIdx.x == 0) ship_ok = ship_slot_free(a.x, a.ship_step);
    __syncthreads();
    const long long ship_at = (SHIP && ship_ok) ? ship_base(a.x, a.ship_step) : -1;  // -1: the slot never came free
    double *cold_d = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + advect_smem_bytes(a.g.n_zl, a.g.n_z, 0)) + threadIdx.x;
    int *cold_i = reinterpret_cast<int *>(cold_d - threadIdx.x + kColdDoubles * blockDim.x) + threadIdx.x;
    (void)cold_i;

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t n = a.p.n;
    SDK_LANE(s, cold_d, cold_i, (int)blockDim.x);
    bool have = false, live = false;
    bool queue_empty = false;
    // per-lane totals for the call statistics (reduced per warp at the end: a few thousand atomics)
    unsigned acc_live = 0, acc_cross = 0, acc_max = 0, acc_flag = 0;
#if SDK_ADVECT_LOCKSTEP
    unsigned iter = 0;
#endif
#if SDK_ADVECT_DEFER_STORE
    // A lane that has finished its particle keeps it until the warp next goes to the queue: write-back and
    // refill then run once for all those lanes together instead of once per finishing lane (on the bench
    // workload: 93 k passes with 10.8 lanes instead of 278 k passes with 3.6), and that code stays out of the
    // instruction cache in between.
    for (;;) {
        const bool working = have && live && s.niter < a.max_iteration;
        const unsigned idle = __ballot_sync(full, !working);
        const unsigned pending = __ballot_sync(full, have && !working);
        if (idle == full || (__popc(idle) >= a.refill_threshold && (pending != 0 || !queue_empty))) {
            if (have && !working) {
                // done (or out of iterations): write back and free the lane
                if (live) s.status |= SDK_ST_MAX_ITER;
                if (a.commit) store_lane<P>(a.p, a.has_dep, s);
                if (SHIP && ship_at >= 0) ship_lane_cold(a, ship_at, s.pid, s.lon, s.lat, s.dep, s.f, s.iy, s.ix, s.izl);
                if (a.p.status) a.p.status[s.pid] = s.status;
                if (a.p.niter) a.p.niter[s.pid] = s.niter;
                if (a.p.ncross) a.p.ncross[s.pid] = s.ncross;
                if (LOG && a.log.count) a.log.count[s.pid] = s.nrec;
                acc_cross += s.ncross;
                acc_max += live;
                acc_flag += s.status != 0;
                have = false;
                live = false;
            }
            if (!queue_empty) {
                const int cnt = __popc(idle);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.queue, (unsigned long long)cnt);
                base = __shfl_sync(full, base, 0);
                if ((int64_t)base + cnt >= n) queue_empty = true;
                if (!have) {
                    const int64_t pid = (int64_t)base + __popc(idle & ((1u << lane) - 1));
                    if (pid < n) {
                        load_lane<P>(a.p, pid, a.has_dep, s);
                        if (a.keep_status) s.status = a.p.status[pid];
                        have = true;
                        if (LOG && a.log.emit_start) note(a, s, 15);
                        live = fabs(a.t_stop - s.t) > a.tol;
                        if (a.p.active && !a.p.active[pid]) live = false;
                        if (LOG && live && (a.log.continue_mode & 2)) note(a, s, 7);
                        acc_live += live;
                    }
                }
            }
        }
#if SDK_ADVECT_LOCKSTEP
        // the warps of a CTA start every SDK_ADVECT_LOCKSTEP-th step together, so that they walk through the
        // same instruction-cache lines at the same time (the exit test has to be block-uniform, so it is
        // only made there; a warp without work idles in between)
        if ((iter++ % SDK_ADVECT_LOCKSTEP) == 0 && !__syncthreads_or(have)) break;
#else
        if (!__any_sync(full, have)) break;
#endif
        if (have && live && s.niter < a.max_iteration) step<T, LOG, P>(a, sZl, sdZl, s, live);
    }
#else
    for (;;) {
        // ---- refill idle lanes from the global queue (warp-aggregated)
        const unsigned idle = __ballot_sync(full, !have);
        if (!queue_empty && (__popc(idle) >= a.refill_threshold || idle == full)) {
            const int cnt = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.queue, (unsigned long long)cnt);
            base = __shfl_sync(full, base, 0);
            if ((int64_t)base + cnt >= n) queue_empty = true;
            if (!have) {
                const int64_t pid = (int64_t)base + __popc(idle & ((1u << lane) - 1));
                if (pid < n) {
                    load_lane<P>(a.p, pid, a.has_dep, s);
                    have = true;
                    if (LOG && a.log.emit_start) note(a, s, 15);
                    live = fabs(a.t_stop - s.t) > a.tol;
                    if (a.p.active && !a.p.active[pid]) live = false;
                    if (LOG && live && (a.log.continue_mode & 2)) note(a, s, 7);
                    acc_live += live;
                }
            }
        }
#if SDK_ADVECT_LOCKSTEP
        // meet the other warps of the CTA every SDK_ADVECT_LOCKSTEP-th iteration (the exit test has
        // to be block-uniform, so it is only made there; a warp without work idles in between)
        if ((iter++ % SDK_ADVECT_LOCKSTEP) == 0 && !__syncthreads_or(have)) break;
#else
        if (!__any_sync(full, have)) break;
#endif
        if (have) {
            if (live && s.niter < a.max_iteration) {
                step<T, LOG, P>(a, sZl, sdZl, s, live);
            } else {
                // done (or out of iterations): write back and free the lane
                if (live) s.status |= SDK_ST_MAX_ITER;
                if (a.commit) store_lane<P>(a.p, a.has_dep, s);
                if (a.p.status) a.p.status[s.pid] = s.status;
                if (a.p.niter) a.p.niter[s.pid] = s.niter;
                if (a.p.ncross) a.p.ncross[s.pid] = s.ncross;
                if (LOG && a.log.count) a.log.count[s.pid] = s.nrec;
                acc_cross += s.ncross;
                acc_max += live;
                acc_flag += s.status != 0;
                have = false;
            }
        }
    }
#endif
    if (a.stats) {
        acc_live = __reduce_add_sync(full, acc_live);
        acc_cross = __reduce_add_sync(full, acc_cross);
        acc_max = __reduce_add_sync(full, acc_max);
        acc_flag = __reduce_add_sync(full, acc_flag);
        if (lane == 0) {
            if (acc_live) atomicAdd(a.stats + 0, (unsigned long long)acc_live);
            if (acc_cross) atomicAdd(a.stats + 1, (unsigned long long)acc_cross);
            if (acc_max) atomicAdd(a.stats + 2, (unsigned long long)acc_max);
            if (acc_flag) atomicAdd(a.stats + 3, (unsigned long long)acc_flag);
        }
    }
    if (SHIP) ship_finish(a.x, a.ship_step);
}

// get_u_du on its own (after Particle.update_uvw_array swapped the velocity slab)
template <typename T, int P>
__global__ void __launch_bounds__(128) get_u_du_kernel(const __grid_constant__ AdvectArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.p.n) return;
    extern __shared__ __align__(16) float smem[];
    double *cold_d = reinterpret_cast<double *>(smem) + threadIdx.x;
    int *cold_i = reinterpret_cast<int *>(cold_d - threadIdx.x + kColdDoubles * blockDim.x) + threadIdx.x;
    (void)cold_i;
    SDK_LANE(s, cold_d, cold_i, (int)blockDim.x);
    load_lane<P>(a.p, i, a.has_dep, s);
    read_velocity<T, P>(a.g, a.v, a.has_dep, s);
    a.p.u[i] = s.u;
    a.p.v[i] = s.v;
    a.p.du[i] = s.du;
    a.p.dv[i] = s.dv;
    if (a.has_dep) {
        // no vertical velocity field (wname=None): w = dw = 0 like the reference's zero arrays
        a.p.w[i] = a.v.has_w ? s.w : 0.0;
        a.p.dw[i] = a.v.has_w ? s.dw : 0.0;
    }
    if (a.p.vol && a.v.transport) a.p.vol[i] = s.vol;
}

}  // namespace sdk

// ------------------------------------------------------------------------------------------------
// host side launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------------
namespace sdk {

int advect_max_threads() { return SDK_ADVECT_THREADS; }
int advect_default_threads() { return SDK_ADVECT_DEFAULT_THREADS; }

// which compile-time profile a launch takes (0 = every switch from the arguments)
int advect_profile(const AdvectArgs &args, int threads) {
    if (args.g.hmode != SDK_H_OCEANPARCEL) return 2;
    const bool llc = SDK_ADVECT_PROFILES && args.g.has_face && args.g.typ == SDK_TYP_LLC && args.v.transport && args.p.face && args.p.vol;
    // profile 1: what the folded switches of the specialised kernel assume
    if (llc && args.has_dep && args.v.has_w && args.v.has_z && args.v.vol_has_z) return 1;
    // profile 3: the surface case (no depth, no vertical velocity, 2-D transports and volumes; fp64 fields, no log)
    if (llc && !args.has_dep && !args.v.has_w && !args.v.has_z && !args.v.vol_has_z && args.v.dtype != SDK_F32 && !args.has_log &&
        threads <= Bounds<3>::threads)
        return 3;
    return 0;
}

// CTAs of SDK_ADVECT_DEFAULT_THREADS threads per SM a profile is compiled for
int advect_default_blocks_per_sm(int profile) {
    return profile == 3 ? Bounds<3>::blocks : SDK_ADVECT_MIN_BLOCKS * SDK_ADVECT_THREADS / SDK_ADVECT_DEFAULT_THREADS;
}

cudaError_t launch_advect(const AdvectArgs &args, int blocks, int threads, cudaStream_t stream) {
    const size_t smem = advect_smem_bytes(args.g.n_zl, args.g.n_z, threads);
    const bool f32 = args.v.dtype == SDK_F32;
    const int profile = advect_profile(args, threads);
    const bool p1 = profile == 1, p3 = profile == 3;
    void (*kern)(AdvectArgs);
    if (args.ship) {
        // (launch_advect's caller has checked: no log with the fused exchange)
        if (args.g.hmode != SDK_H_OCEANPARCEL) kern = f32 ? advect_kernel<float, false, 2, true> : advect_kernel<double, false, 2, true>;
        else if (f32) kern = p1 ? advect_kernel<float, false, 1, true> : advect_kernel<float, false, 0, true>;
        else kern = p1 ? advect_kernel<double, false, 1, true> : (p3 ? advect_kernel<double, false, 3, true> : advect_kernel<double, false, 0, true>);
    } else if (args.g.hmode != SDK_H_OCEANPARCEL) {
        if (args.has_log) kern = f32 ? advect_kernel<float, true, 2> : advect_kernel<double, true, 2>;
        else kern = f32 ? advect_kernel<float, false, 2> : advect_kernel<double, false, 2>;
    } else if (args.has_log) {
        if (f32) kern = p1 ? advect_kernel<float, true, 1> : advect_kernel<float, true, 0>;
        else kern = p1 ? advect_kernel<double, true, 1> : advect_kernel<double, true, 0>;
    } else {
        if (f32) kern = p1 ? advect_kernel<float, false, 1> : advect_kernel<float, false, 0>;
        else kern = p1 ? advect_kernel<double, false, 1> : (p3 ? advect_kernel<double, false, 3> : advect_kernel<double, false, 0>);
    }
    if (smem > 48 * 1024) {  // CTAs of more than ~384 threads need more than the default 48 KB of dynamic shared memory
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    kern<<<blocks, threads, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_get_u_du(const AdvectArgs &args, cudaStream_t stream) {
    const int threads = 128;
    const int blocks = (int)((args.p.n + threads - 1) / threads);
    if (blocks == 0) return cudaSuccess;
    const size_t smem = advect_smem_bytes(0, 0, threads);
    const bool local = args.g.hmode != SDK_H_OCEANPARCEL;
    if (args.v.dtype == SDK_F32) {
        if (local) get_u_du_kernel<float, 2><<<blocks, threads, smem, stream>>>(args);
        else get_u_du_kernel<float, 0><<<blocks, threads, smem, stream>>>(args);
    } else {
        if (local) get_u_du_kernel<double, 2><<<blocks, threads, smem, stream>>>(args);
        else get_u_du_kernel<double, 0><<<blocks, threads, smem, stream>>>(args);
    }
    return cudaGetLastError();
}

}  // namespace sdk
