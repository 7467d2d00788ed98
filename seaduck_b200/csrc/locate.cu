// K2: lat/lon/depth/time -> grid indices + rel-coords.
//
// Replaces Position.from_latlon (reference seaduck/eulerian.py:138) -> OceData._find_rel_h
// (ocedata.py:422) -> find_rel_h_oceanparcel (utils.py:614): nearest cell centre in 3-D chord
// distance (reference: scipy cKDTree on cartesian XC/YC, utils.py:225,309), the four corner points
// (find_px_py utils.py:495) and the inverse bilinear map (find_rx_ry_oceanparcel utils.py:532);
// and the 1-D searches find_rel_nearest / find_rel_z / find_rel_time (utils.py:321-454).
//
// Instead of a KD-tree the device uses a lat/lon bucket grid that stores one nearby cell per bucket
// and a greedy walk over the 8 neighbours (crossing LLC faces through the constant-memory face
// tables) that ends on the cell whose centre is closest to the query.
#include <cuda_runtime.h>

#include "locate_args.cuh"

namespace sdk {

__device__ __forceinline__ void unit_vec(double lon, double lat, double &x, double &y, double &z) {
    const double d2r = 0.017453292519943295;
    double sl, cl, sp, cp;
    sincos(lon * d2r, &sl, &cl);
    sincos(lat * d2r, &sp, &cp);
    x = cp * cl;
    y = cp * sl;
    z = sp;
}

__device__ __forceinline__ int bucket_of(const BucketGrid &b, double lon, double lat) {
    double dl = lon - b.lon0;
    dl = dl - 360.0 * floor(dl / 360.0);  // [0, 360)
    int bx = (int)(dl * b.inv_dlon);
    int by = (int)((lat - b.lat0) * b.inv_dlat);
    if (b.periodic) {
        bx = bx % b.nlon;
    } else {
        // non-global grid: points west of lon0 may have wrapped to ~360
        if (dl > b.span_lon + 0.5 * (360.0 - b.span_lon)) bx = 0;
        bx = min(max(bx, 0), b.nlon - 1);
    }
    by = min(max(by, 0), b.nlat - 1);
    return by * b.nlon + bx;
}

__global__ void bucket_scatter_kernel(GridDev g, BucketGrid b) {
    const int64_t ncell = (int64_t)g.nface * g.ny * g.nx;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell;
         c += (int64_t)gridDim.x * blockDim.x) {
        const float lon = g.XC[c], lat = g.YC[c];
        if (!(lon == lon) || !(lat == lat)) continue;
        atomicMin(&b.table[bucket_of(b, (double)lon, (double)lat)], (int)c);
    }
}

// one dilation pass: empty buckets copy a filled neighbour; *changed counts buckets still empty
__global__ void bucket_dilate_kernel(BucketGrid b, const int *src, int *dst, int *remaining) {
    const int nb = b.nlat * b.nlon;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
        int v = src[i];
        if (v == INT_MAX) {
            const int by = i / b.nlon, bx = i % b.nlon;
            for (int dy = -1; dy <= 1 && v == INT_MAX; ++dy) {
                const int yy = by + dy;
                if (yy < 0 || yy >= b.nlat) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = bx + dx;
                    if (b.periodic) xx = (xx + b.nlon) % b.nlon;
                    else if (xx < 0 || xx >= b.nlon) continue;
                    const int w = src[yy * b.nlon + xx];
                    if (w != INT_MAX) {
                        v = w;
                        break;
                    }
                }
            }
            if (v == INT_MAX) atomicAdd(remaining, 1);
        }
        dst[i] = v;
    }
}

// rot in quarter turns when leaving through edge e and arriving through edge ne
// (reference topology.py:470: rot = pi - DIRECTIONS[e] + DIRECTIONS[ne])
__device__ __forceinline__ int remap_dir(int dir, int e, int ne) {
    // (the small tables are packed into immediates, two bits per entry: indexed local arrays would live on the stack)
    const unsigned dq = 0x2Du;  // {1, 3, 2, 0}: quarter turns of the edge directions up, down, left, right
    const int q = (2 - (int)(dq >> (2 * e) & 3u) + (int)(dq >> (2 * ne) & 3u) + 8) & 3;
    if (q == 1) return (int)(0x1Eu >> (2 * dir) & 3u);  // {2, 3, 1, 0}
    if (q == 3) return (int)(0x4Bu >> (2 * dir) & 3u);  // {3, 2, 0, 1}
    return dir;
}

// One index step of the SEARCH: cell_move, except on a zonally periodic grid, where the reference's own step across the
// seam lands on column 1 or nx - 1 (_x_per_ind_tend, topology.py:124 -- kept as it is for particles that cross a wall):
// the nearest-centre search needs the cells that really lie next to each other there, columns nx - 1 and 0.
__device__ __forceinline__ int search_move(const GridDev &g, int dir, int &f, int &iy, int &ix, int &came_in) {
    if (g.typ == SDK_TYP_XPERIODIC && dir >= 2) {
        came_in = -1;
        ix = dir == 3 ? (ix + 1 < g.nx ? ix + 1 : 0) : (ix > 0 ? ix - 1 : g.nx - 1);
        return 0;
    }
    return cell_move(g, dir, f, iy, ix, came_in);
}

// neighbour of a cell after up to two moves (second one re-labelled after a face change)
__device__ __forceinline__ bool neighbour(const GridDev &g, int f, int iy, int ix, int d1, int d2,
                                          int &nf, int &ny, int &nx) {
    nf = f;
    ny = iy;
    nx = ix;
    int came;
    if (search_move(g, d1, nf, ny, nx, came)) return false;
    if (d2 >= 0) {
        if (came >= 0) d2 = remap_dir(d2, d1, came);
        if (search_move(g, d2, nf, ny, nx, came)) return false;
    }
    return true;
}

// The centre table is HALOED: [nface][ny + 2][nx + 2][3], the ring around a face holding the centres of the cells
// across its edges (filled once per grid through cell_move, so across LLC faces, the periodic seam, ...).  The 8
// neighbours of ANY cell are then index offsets, and a point near a face edge costs what an interior one does.
__device__ __forceinline__ int64_t halo_index(const GridDev &g, int f, int iy, int ix) {
    return ((int64_t)f * (g.ny + 2) + (iy + 1)) * (g.nx + 2) + (ix + 1);
}

// slot of a halo position in the per-face rim table (flat index of the cell that lives there, -1 = none):
// bottom row (ix = -1 .. nx), top row, left column (iy = 0 .. ny - 1), right column
__device__ __forceinline__ int64_t rim_slot(const GridDev &g, int f, int iy, int ix) {
    const int row = g.nx + 2;
    const int64_t base = (int64_t)f * (2 * row + 2 * g.ny);
    if (iy < 0) return base + ix + 1;
    if (iy >= g.ny) return base + row + ix + 1;
    return base + 2 * row + (ix < 0 ? 0 : g.ny) + iy;
}

__device__ __forceinline__ double dist2_to_cell(const GridDev &g, const double *cxyz, int f, int iy, int ix, double qx,
                                                double qy, double qz) {
    double x, y, z;
    if (cxyz) {
        const int64_t h = halo_index(g, f, iy, ix);
        x = __ldg(cxyz + 3 * h), y = __ldg(cxyz + 3 * h + 1), z = __ldg(cxyz + 3 * h + 2);
    } else {
        const int64_t c = ((int64_t)f * g.ny + iy) * g.nx + ix;
        const float lon = __ldg(g.XC + c), lat = __ldg(g.YC + c);
        if (!(lon == lon) || !(lat == lat)) return INFINITY;
        unit_vec((double)lon, (double)lat, x, y, z);
    }
    const double dx = x - qx, dy = y - qy, dz = z - qz;
    return dx * dx + dy * dy + dz * dz;
}

// the table behind dist2_to_cell: the same unit_vec of the same float32 centre, evaluated once per cell, and for the
// halo ring the cell one (edge) or two (corner: vertical move first, as the scan of the 8 neighbours orders them) cell
// moves away.  A halo position without a cell (closed edge, missing face) gets a far-away FINITE point: never the
// nearest, and not mistaken for a NaN centre (+inf), which makes the walk look further out.
__global__ void cell_xyz_kernel(GridDev g, double *cxyz, int *rim) {
    const int H = g.ny + 2, W = g.nx + 2;
    const int64_t ntab = (int64_t)g.nface * H * W;
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < ntab; h += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(h / ((int64_t)H * W));
        const int rem = (int)(h - (int64_t)f * H * W);
        const int iy = rem / W - 1, ix = rem % W - 1;
        int nf = f, ny = iy, nx = ix;
        bool there = true;
        const bool off_y = iy < 0 || iy >= g.ny, off_x = ix < 0 || ix >= g.nx;
        if (off_y || off_x) {
            const int sy = min(max(iy, 0), g.ny - 1), sx = min(max(ix, 0), g.nx - 1);
            const int dv = iy < 0 ? 1 : 0, dh = ix < 0 ? 2 : 3;
            there = off_y ? neighbour(g, f, sy, sx, dv, off_x ? dh : -1, nf, ny, nx) : neighbour(g, f, sy, sx, dh, -1, nf, ny, nx);
            rim[rim_slot(g, f, iy, ix)] = there ? (int)(((int64_t)nf * g.ny + ny) * g.nx + nx) : -1;
        }
        double x = 1e150, y = 0.0, z = 0.0;
        if (there) {
            const int64_t c = ((int64_t)nf * g.ny + ny) * g.nx + nx;
            const float lon = g.XC[c], lat = g.YC[c];
            x = y = z = INFINITY;
            if (lon == lon && lat == lat) unit_vec((double)lon, (double)lat, x, y, z);
        }
        cxyz[3 * h] = x, cxyz[3 * h + 1] = y, cxyz[3 * h + 2] = z;
    }
}

// The cell (dy, dx) index steps away from (f, iy, ix), |dy|, |dx| <= 3: vertical steps first, then horizontal ones,
// every remaining step re-labelled after a face change (reference ind_moves, topology.py:436).  false = no such cell.
__device__ __forceinline__ bool cell_at_offset(const GridDev &g, int f, int iy, int ix, int dy, int dx, int &nf, int &ny, int &nx) {
    // the moves live in one word, two bits each (no indexed local array: that would put the whole kernel on the stack)
    unsigned moves = 0;
    int m = 0;
    for (int k = 0; k < abs(dy); ++k) moves |= (dy > 0 ? 0u : 1u) << (2 * m++);
    for (int k = 0; k < abs(dx); ++k) moves |= (dx > 0 ? 3u : 2u) << (2 * m++);
    nf = f, ny = iy, nx = ix;
    for (int k = 0; k < m; ++k) {
        int came;
        const int d = (int)(moves >> (2 * k) & 3u);
        if (search_move(g, d, nf, ny, nx, came)) return false;
        if (came >= 0) {
            for (int j = k + 1; j < m; ++j) {
                const unsigned old = moves >> (2 * j) & 3u;
                moves = (moves & ~(3u << (2 * j))) | ((unsigned)remap_dir((int)old, d, came) << (2 * j));
            }
        }
    }
    return true;
}

// np.argmin(np.abs(arr - x)) (find_ind, utils.py:273): the first of the nearest entries.  `order` = +1 / -1 for a strictly
// ascending / descending table (GridDev::z_order), which is then bisected: the distance falls, then rises, so the minimum
// is one of the two entries around x -- the lower index on a tie, as argmin has it; before the table it is entry 0, past
// its end the last entry.  Anything unusual (unsorted table, NaN, a value further past the end than the table is long,
// where rounding could make distant entries tie) takes the scan.
static __device__ __noinline__ int nearest_index_scan(const float *arr, int n, double x) {
    int best = 0;
    double bestv = INFINITY;
    for (int k = 0; k < n; ++k) {
        const double a = fabs((double)arr[k] - x);
        if (a < bestv) {
            bestv = a;
            best = k;
        }
    }
    return best;
}

__device__ __forceinline__ int nearest_index(const float *arr, int n, double x, int order) {
    if (order != 0 && x == x) {
        const double s = (double)order, sx = s * x;
        const double first = s * (double)arr[0], last = s * (double)arr[n - 1];
        if (sx < first) return 0;
        if (sx < last) {
            int lo = 0, hi = n - 1;  // s * arr[lo] <= sx < s * arr[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s * (double)arr[mid] <= sx) lo = mid;
                else hi = mid;
            }
            return fabs((double)arr[hi] - x) < fabs((double)arr[lo] - x) ? hi : lo;
        }
        if (sx - last <= last - first) return n - 1;
    }
    return nearest_index_scan(arr, n, x);
}

// reference find_rel (utils.py:321) with find_ind (utils.py:273) for a 1-D table in shared/global
__device__ __forceinline__ void find_rel_1d(const float *arr, const float *darr, int n, int n_d,
                                            double x, int ascending, bool above, bool dx_right,
                                            int &idx, double &r, double &d, double &b, int &status, int order = 0) {
    int best = nearest_index(arr, n, x, order);
    if (above && (double)arr[best] > x) best -= ascending;
    if (best < 0 || best > n - 1) status |= SDK_ST_OOB;  // reference raises "Value out of bound."
    best = min(max(best, 0), n - 1);
    const double bx = (double)arr[best];
    const double dx = x - bx;
    int off = (int)(!dx_right) + (int)(ascending > 0) - 1;
    if (!above) off = (int)(!dx_right) + (int)(ascending * dx > 0) - 1;
    int j = best + off;
    double dval;
    if (darr) {
        dval = (double)darr[wrap_index(j, n_d)];
    } else {
        // darray = |diff(array)| with the last value repeated (utils.py:361-364), float32 arithmetic
        j = wrap_index(j, n);
        const int jj = j < n - 1 ? j : n - 2;
        dval = (double)fabsf(arr[jj + 1] - arr[jj]);
    }
    idx = best;
    r = dx / dval;
    d = dval;
    b = bx;
}

// find_rel_periodic (utils.py:392 -> find_rel :321 with peri=360, find_ind :273) on a float32 axis: nearest node in
// the periodic distance (first minimum, like np.argmin), spacing |diff(axis)| in float32 with the last value
// repeated, taken to the right of the node when the point lies to its right, else to the left (numpy index
// wrap-around for -1 included).
__device__ __forceinline__ void find_rel_periodic(const float *arr, int n, double x, int &idx, double &r, double &d,
                                                  double &b) {
    int best = 0;
    double bestv = INFINITY;
    for (int k = 0; k < n; ++k) {
        const double a = fabs(to_180((double)__ldg(arr + k) - x));
        if (a < bestv) {
            bestv = a;
            best = k;
        }
    }
    const double bx = (double)__ldg(arr + best);
    const double dx = to_180(x - bx);
    int j = wrap_index(best + (int)(dx > 0) - 1, n);
    const int jj = j < n - 1 ? j : n - 2;
    const double dval = n > 1 ? (double)fabsf(__ldg(arr + jj + 1) - __ldg(arr + jj)) : 1.0;
    idx = best;
    r = dx / dval;
    d = dval;
    b = bx;
}

// same for float64 tables (time axes)
__device__ __forceinline__ void find_rel_1d_f64(const double *arr, int n, double x, bool above, int &idx,
                                                double &r, double &d, double &b, int &status) {
    int best = 0;
    double bestv = INFINITY;
    for (int k = 0; k < n; ++k) {
        const double a = fabs(arr[k] - x);
        if (a < bestv) {
            bestv = a;
            best = k;
        }
    }
    if (above && arr[best] > x) best -= 1;
    if (best < 0) status |= SDK_ST_OOB;
    best = min(max(best, 0), n - 1);
    const double bx = arr[best];
    const double dx = x - bx;
    int off = 0;
    if (!above) off = (int)(dx > 0) - 1;
    int j = wrap_index(best + off, n);
    const int jj = j < n - 1 ? j : n - 2;
    const double dval = n > 1 ? fabs(arr[jj + 1] - arr[jj]) : 1.0;
    idx = best;
    r = dx / dval;
    d = dval;
    b = bx;
}

// Which cells need the second ring: with a = centre -> right neighbour and b = centre -> upper neighbour, the Voronoi
// neighbours of a lattice point all lie among the 8 index neighbours when the basis is (nearly) reduced,
// |a.b| <= min(|a|^2, |b|^2) / 2; real ocean grids are orthogonal (a.b ~ 0), a sheared or kinked patch is flagged cell by
// cell (threshold 0.35, and any cell with a missing or NaN neighbour).
__global__ void cell_wide_kernel(GridDev g, const double *cxyz, unsigned char *wide) {
    const int64_t ncell = (int64_t)g.nface * g.ny * g.nx;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(c / ((int64_t)g.ny * g.nx));
        const int rem = (int)(c - (int64_t)f * g.ny * g.nx);
        const int iy = rem / g.nx, ix = rem - iy * g.nx;
        double v[2][3];
        bool flag = false;
        for (int k = 0; k < 2; ++k) {
            int nf, ny, nx;
            // right / up neighbour, or left / down where there is none (the sign does not matter)
            if (!neighbour(g, f, iy, ix, k == 0 ? 3 : 0, -1, nf, ny, nx) && !neighbour(g, f, iy, ix, k == 0 ? 2 : 1, -1, nf, ny, nx)) {
                flag = true;
                v[k][0] = v[k][1] = v[k][2] = 0.0;
                continue;
            }
            const int64_t n = halo_index(g, nf, ny, nx), h = halo_index(g, f, iy, ix);
            for (int j = 0; j < 3; ++j) v[k][j] = cxyz[3 * n + j] - cxyz[3 * h + j];
        }
        const double ab = v[0][0] * v[1][0] + v[0][1] * v[1][1] + v[0][2] * v[1][2];
        const double aa = v[0][0] * v[0][0] + v[0][1] * v[0][1] + v[0][2] * v[0][2];
        const double bb = v[1][0] * v[1][0] + v[1][1] * v[1][1] + v[1][2] * v[1][2];
        const double lim = 0.35 * fmin(aa, bb);
        flag |= !(fabs(ab) <= lim) || !(lim > 0.0);  // (also true for NaN / inf operands)
        wide[c] = flag ? 1 : 0;
    }
}

// Second ring around a cell where the walk over the 8 neighbours has stopped -- on a sheared or strongly stretched grid
// the Voronoi neighbours of a cell are not all among its 8 index neighbours, and the descent could stop one cell short
// -- and a third ring where a centre nearby is NaN (land without coordinates: the nearest valid centre may lie beyond
// the hole).  Returns the closest cell found (d2 >= best: none closer).
struct Nearest {
    double d2;
    int f, iy, ix;
    bool hole;  // scan_neighbours_cold: one of the cells looked at has no centre (NaN coordinates)
};

static __device__ __noinline__ Nearest confirm_rings(const GridDev &g, const double *cxyz, int f, int iy, int ix, double qx, double qy,
                                                     double qz, double best, bool hole) {
    Nearest r = {best, f, iy, ix, false};
#ifndef SDK_LOCATE_INTERIOR
#define SDK_LOCATE_INTERIOR 1
#endif
#ifndef SDK_LOCATE_MAXRING
#define SDK_LOCATE_MAXRING 3
#endif
    const bool inside = SDK_LOCATE_INTERIOR && cxyz && iy >= 2 && iy < g.ny - 2 && ix >= 2 && ix < g.nx - 2;
    if (inside) {
        // away from the face edges the second ring is plain index arithmetic on the centre table
        const int64_t c0 = halo_index(g, f, iy, ix);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const int dy = k < 5 ? -2 : (k < 10 ? 2 : (k < 13 ? k - 11 : k - 14));
            const int dx = k < 5 ? k - 2 : (k < 10 ? k - 7 : (k < 13 ? -2 : 2));
            const int64_t c = c0 + (int64_t)dy * (g.nx + 2) + dx;
            const double ex = __ldg(cxyz + 3 * c) - qx, ey = __ldg(cxyz + 3 * c + 1) - qy, ez = __ldg(cxyz + 3 * c + 2) - qz;
            const double dd = ex * ex + ey * ey + ez * ez;
            hole |= !(dd < INFINITY);
            if (dd < r.d2) r = {dd, f, iy + dy, ix + dx, false};
        }
    }
    for (int ring = inside ? 3 : 2; ring <= ((hole && SDK_LOCATE_MAXRING >= 3) ? 3 : (SDK_LOCATE_MAXRING >= 2 ? 2 : 0)); ++ring) {
        for (int dy = -ring; dy <= ring; ++dy) {
            for (int dx = -ring; dx <= ring; ++dx) {
                if (abs(dy) != ring && abs(dx) != ring) continue;
                int nf, ny, nx;
                if (!cell_at_offset(g, f, iy, ix, dy, dx, nf, ny, nx)) continue;
                const double dd = dist2_to_cell(g, cxyz, nf, ny, nx, qx, qy, qz);
                hole |= !(dd < INFINITY);
                if (dd < r.d2) r = {dd, nf, ny, nx, false};
            }
        }
    }
    return r;
}

// The 8 neighbours of a cell next to a face edge (or of any cell when there is no centre table): one or two cell moves
// each, across faces where need be.  Out of line and rolled (the two moves of neighbour k are packed two bits each):
// only a few per cent of the cells come here, and the walk's loop should not carry eight inlined copies of cell_move.
static __device__ __noinline__ Nearest scan_neighbours_cold(const GridDev &g, const double *cxyz, int f, int iy, int ix, double qx, double qy,
                                                            double qz, double best) {
    Nearest r = {best, f, iy, ix, false};
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const int d1 = (int)(0x50E4u >> (2 * k) & 3u);                       // {0, 1, 2, 3, 0, 0, 1, 1}
        const int d2 = k < 4 ? -1 : (int)(0xEEu >> (2 * (k - 4)) & 3u);      // {-, -, -, -, 2, 3, 2, 3}
        int nf, ny, nx;
        if (!neighbour(g, f, iy, ix, d1, d2, nf, ny, nx)) continue;
        const double dd = dist2_to_cell(g, cxyz, nf, ny, nx, qx, qy, qz);
        const bool hole = r.hole || !(dd < INFINITY);
        if (dd < r.d2) r = {dd, nf, ny, nx, false};
        r.hole = hole;
    }
    return r;
}

__global__ void __launch_bounds__(128) locate_kernel(const __grid_constant__ LocateArgs a) {
    const GridDev &g = a.g;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const sdk_located &o = a.out;
    int status = 0;
    if (a.lon && a.lat && g.hmode == SDK_H_RECTILINEAR) {
        // find_rel_h_rectilinear, utils.py:602: two periodic 1-D lookups; dx = cos(by) * deg2m * spacing in float64
        const double lon = a.lon[i], lat = a.lat[i];
        int ix, iy;
        double rx, ry, ddx, ddy, bx, by;
        find_rel_periodic(g.lon1, g.nx, lon, ix, rx, ddx, bx);
        find_rel_periodic(g.lat1, g.ny, lat, iy, ry, ddy, by);
        const double deg2m = 6271e3 * 3.141592653589793 / 180;
        const double dxm = cos(by * 3.141592653589793 / 180) * deg2m * ddx;
        const double dym = deg2m * ddy;
        o.iy[i] = iy;
        o.ix[i] = ix;
        o.rx[i] = rx;
        o.ry[i] = ry;
        if (o.bx) o.bx[i] = (float)bx;
        if (o.by) o.by[i] = (float)by;
        if (o.dx) o.dx[i] = (float)dxm;
        if (o.dy) o.dy[i] = (float)dym;
        if (o.cs) o.cs[i] = 1.f;
        if (o.sn) o.sn[i] = 0.f;
        if (o.px) {
            o.px[0 * a.n + i] = bx, o.px[1 * a.n + i] = 1.0, o.px[2 * a.n + i] = dxm, o.px[3 * a.n + i] = 0.0;
            o.py[0 * a.n + i] = by, o.py[1 * a.n + i] = 0.0, o.py[2 * a.n + i] = dym, o.py[3 * a.n + i] = 0.0;
        }
    } else if (a.lon && a.lat) {
        const double lon = a.lon[i], lat = a.lat[i];
        double qx, qy, qz;
        unit_vec(lon, lat, qx, qy, qz);
        int c0 = __ldg(a.b.table + bucket_of(a.b, lon, lat));
        int f = c0 / (g.ny * g.nx);
        int rem = c0 - f * (g.ny * g.nx);
        int iy = rem / g.nx, ix = rem - iy * g.nx;
        const double *cxyz = a.b.cxyz;
        double best = dist2_to_cell(g, cxyz, f, iy, ix, qx, qy, qz);
        // Greedy walk to the nearest centre over the 8 neighbours; where it stops, the 16 cells of the second ring
        // are looked at as well (on a sheared or strongly stretched grid the Voronoi neighbours of a cell are not all
        // among its 8 index neighbours, and the descent could stop one cell short) and the walk resumes from a closer
        // one if there is any.
        for (int iter = 0; iter < a.max_walk; ++iter) {
            int bf = f, by = iy, bx = ix;
            double bd = best;
            bool near_hole = false;
            if (cxyz) {
                // the 8 neighbours are index offsets into the (haloed) centre table -- in the order of the general scan
                // (up, down, left, right, up-left, up-right, down-left, down-right: ties go to the first)
                const int64_t c0 = halo_index(g, f, iy, ix);
                const int row = g.nx + 2;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int dy = (k == 0 || k == 4 || k == 5) ? 1 : ((k == 1 || k == 6 || k == 7) ? -1 : 0);
                    const int dx = (k == 3 || k == 5 || k == 7) ? 1 : ((k == 2 || k == 4 || k == 6) ? -1 : 0);
                    const int64_t c = c0 + dy * row + dx;
                    const double ex = __ldg(cxyz + 3 * c) - qx, ey = __ldg(cxyz + 3 * c + 1) - qy, ez = __ldg(cxyz + 3 * c + 2) - qz;
                    const double dd = ex * ex + ey * ey + ez * ez;
                    near_hole |= !(dd < INFINITY);
                    if (dd < bd) {
                        bd = dd;
                        by = iy + dy;
                        bx = ix + dx;
                    }
                }
                if (SDK_UNLIKELY(((unsigned)by >= (unsigned)g.ny) | ((unsigned)bx >= (unsigned)g.nx))) {
                    // the closest neighbour lives across a face edge: the rim table says which cell it is
                    const int c = __ldg(a.b.rim + rim_slot(g, f, by, bx));
                    if (c < 0) break;  // (no cell there: its far-away stand-in can only win against a seed without a centre)
                    bf = c / (g.ny * g.nx);
                    const int r = c - bf * (g.ny * g.nx);
                    by = r / g.nx;
                    bx = r - by * g.nx;
                }
            } else {
                const Nearest nb = scan_neighbours_cold(g, cxyz, f, iy, ix, qx, qy, qz, best);
                bd = nb.d2, bf = nb.f, by = nb.iy, bx = nb.ix, near_hole = nb.hole;
            }
            if (bd >= best) {
                // the walk has stopped: where the grid is sheared (or has holes) look further out once
                if (a.b.wide && !near_hole && !__ldg(a.b.wide + ((int64_t)f * g.ny + iy) * g.nx + ix)) break;
                const Nearest far = confirm_rings(g, cxyz, f, iy, ix, qx, qy, qz, best, near_hole);
                if (far.d2 >= best) break;
                bd = far.d2, bf = far.f, by = far.iy, bx = far.ix;
            }
            best = bd;
            f = bf;
            iy = by;
            ix = bx;
        }
        if (o.face) o.face[i] = f;
        o.iy[i] = iy;
        o.ix[i] = ix;
        const int64_t c = ((int64_t)f * g.ny + iy) * g.nx + ix;
        const float cbx = __ldg(g.XC + c), cby = __ldg(g.YC + c);
        const float cdx = g.dX ? __ldg(g.dX + c) : 0.f, cdy = g.dY ? __ldg(g.dY + c) : 0.f;
        const float ccs = g.CS ? __ldg(g.CS + c) : 1.f, csn = g.SN ? __ldg(g.SN + c) : 0.f;
        if (o.bx) o.bx[i] = cbx;
        if (o.by) o.by[i] = cby;
        if (o.dx) o.dx[i] = cdx;
        if (o.dy) o.dy[i] = cdy;
        if (o.cs) o.cs[i] = ccs;
        if (o.sn) o.sn[i] = csn;
        if (g.hmode == SDK_H_LOCAL_CARTESIAN) {
            // find_rel_h_naive, utils.py:577 -> find_rx_ry_naive :522 (numba: the cosine of the float32 latitude in
            // double precision)
            const double bx = (double)cbx, by = (double)cby, cs = (double)ccs, sn = (double)csn;
            const double deg2m = 6271e3 * 3.141592653589793 / 180;
            const double dlon = to_180(lon - bx), dlat = to_180(lat - by);
            const double cb = cos(by * 3.141592653589793 / 180);
            o.rx[i] = (dlon * cb * cs + dlat * sn) * deg2m / (double)cdx;
            o.ry[i] = (dlat * cs - dlon * sn * cb) * deg2m / (double)cdy;
            if (o.px) {
                o.px[0 * a.n + i] = bx, o.px[1 * a.n + i] = cs, o.px[2 * a.n + i] = (double)cdx, o.px[3 * a.n + i] = 0.0;
                o.py[0 * a.n + i] = by, o.py[1 * a.n + i] = sn, o.py[2 * a.n + i] = (double)cdy, o.py[3 * a.n + i] = 0.0;
            }
        } else {
            // corners + inverse bilinear map
            int64_t cc[4];
            status |= corner_indices(g, f, iy, ix, cc);
            double px[4], py[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 q = __ldg(g.XYG + cc[j]);
                px[j] = lon + to_180((double)q.x - lon);
                py[j] = (double)q.y;
                if (o.px) {
                    o.px[j * a.n + i] = px[j];
                    o.py[j * a.n + i] = py[j];
                }
            }
            double rx, ry;
            inverse_bilinear(lon, lat, px, py, rx, ry);
            o.rx[i] = rx;
            o.ry[i] = ry;
        }
    }
    if (a.dep) {
        const double z = a.dep[i];
        int k;
        double r, d, b;
        if (g.n_zc > 1 && o.iz) {  // _find_rel_v: nearest Z
            find_rel_1d(g.Z, nullptr, g.n_zc, 0, z, 1, false, true, k, r, d, b, status, g.z_order);
            o.iz[i] = k; o.rz[i] = r; o.dz[i] = d; o.bz[i] = b;
        }
        if (g.n_zc > 1 && o.iz_lin) {  // _find_rel_v_lin: find_rel_z(z, Z, dZ) -> ascending=-1, dx_right=False
            int st = 0;
            find_rel_1d(g.Z, g.dZ, g.n_zc, g.n_zc, z, -1, true, false, k, r, d, b, st, g.z_order);
            if (st) status |= SDK_ST_OOB_Z;
            o.iz_lin[i] = k; o.rz_lin[i] = r; o.dz_lin[i] = d; o.bz_lin[i] = b;
        }
        if (g.n_zl > 1) {
            if (o.izl_near) {  // _find_rel_vl: nearest Zl
                find_rel_1d(g.Zl, nullptr, g.n_zl, 0, z, 1, false, true, k, r, d, b, status, g.zl_order);
                o.izl_near[i] = k; o.rzl_near[i] = r; o.dzl_near[i] = d; o.bzl_near[i] = b;
            }
            // _find_rel_vl_lin: find_rel_z(z, Zl, dZl, dz_above_z=False) -> ascending=-1, dx_right=True
            if (o.izl) {
                find_rel_1d(g.Zl, g.dZl, g.n_zl, g.n_z, z, -1, true, true, k, r, d, b, status, g.zl_order);
                o.izl[i] = k; o.rzl[i] = r; o.dzl[i] = d; o.bzl[i] = b;
            }
        }
    }
    if (a.t && a.ts && a.n_t > 1) {  // _find_rel_t: nearest time level
        int k;
        double r, d, b;
        if (o.it) {
            find_rel_1d_f64(a.ts, a.n_t, a.t[i], false, k, r, d, b, status);
            o.it[i] = k;
            if (o.rt) { o.rt[i] = r; o.dt[i] = d; o.bt[i] = b; }
        }
        if (o.it_lin) {  // _find_rel_t_lin: find_rel_time -> find_rel(t, ts), above
            int st = 0;
            find_rel_1d_f64(a.ts, a.n_t, a.t[i], true, k, r, d, b, st);
            if (st) status |= SDK_ST_OOB_T;
            o.it_lin[i] = k; o.rt_lin[i] = r; o.dt_lin[i] = d; o.bt_lin[i] = b;
        }
    }
    if (o.status) o.status[i] = status;
}

// sort key of a point that has not been located yet: (vertical level, locator bucket), see launch_point_sort (sort.cu)
__global__ void __launch_bounds__(256) point_key_kernel(GridDev g, BucketGrid b, int64_t n, const double *lon, const double *lat,
                                                        const double *dep, uint64_t *key, int32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t level = 0;
    if (dep && g.n_zl > 1) {
        int k, st = 0;
        double r, d, bb;
        find_rel_1d(g.Zl, g.dZl, g.n_zl, g.n_z, dep[i], -1, true, true, k, r, d, bb, st, g.zl_order);
        level = (uint64_t)k;
    }
    const double x = lon[i], y = lat[i];
    const uint64_t bucket = (x == x && y == y) ? (uint64_t)bucket_of(b, x, y) : 0;
    key[i] = level * ((uint64_t)b.nlat * (uint64_t)b.nlon) + bucket;
    idx[i] = (int32_t)i;
}

cudaError_t launch_point_keys(const GridDev &g, const BucketGrid &b, int64_t n, const double *lon, const double *lat, const double *dep,
                              uint64_t *key, int32_t *idx, cudaStream_t s) {
    point_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, b, n, lon, lat, dep, key, idx);
    return cudaGetLastError();
}

// it = 0 before time_midp[0], else find_rel(t, time_midp)[0] + 1 (reference lagrangian.py:796-802)
__global__ void time_index_kernel(const double *t, const double *midp, int n_midp, int32_t *it, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = t[i];
    if (x < midp[0]) {
        it[i] = 0;
        return;
    }
    int k, st = 0;
    double r, d, b;
    find_rel_1d_f64(midp, n_midp, x, true, k, r, d, b, st);
    it[i] = k + 1;
}

// t := t_stop and it := index(t_stop) for everybody, but only if the advect call had live particles
__global__ void finish_stop_kernel(double *t, int32_t *it, const double *midp, int n_midp, double t_stop,
                                   const unsigned long long *stats, unsigned long long *total, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool any_live = stats[0] > 0;
    if (i < n && any_live) {
        t[i] = t_stop;
        if (it && midp) {
            int idx = 0;
            if (!(t_stop < midp[0])) {
                int k, st = 0;
                double r, d, b;
                find_rel_1d_f64(midp, n_midp, t_stop, true, k, r, d, b, st);
                idx = k + 1;
            }
            it[i] = idx;
        }
    }
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x < 4) {
        // the last block adds the call's counters to the running totals; it reads stats only after
        // having used stats[0] itself, and nobody else writes them during this kernel
        total[threadIdx.x] += stats[threadIdx.x];
    }
}

__global__ void fill_int_kernel(int *p, int v, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

// Setup (once per grid): scatter cells into buckets, then grow the filled region until no bucket is
// empty.  Host-synchronous; not part of any timed path.
cudaError_t build_buckets(const GridDev &g, BucketGrid &b, cudaStream_t s) {
    const int nb = b.nlat * b.nlon;
    int *tmp = nullptr, *remaining = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&b.table, sizeof(int) * (size_t)nb)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&tmp, sizeof(int) * (size_t)nb)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&remaining, sizeof(int))) != cudaSuccess) return e;
    const int threads = 256;
    const int blocks = min((nb + threads - 1) / threads, 4096);
    fill_int_kernel<<<blocks, threads, 0, s>>>(b.table, INT_MAX, nb);
    const int64_t ncell = (int64_t)g.nface * g.ny * g.nx;
    int64_t sblocks = (ncell + threads - 1) / threads;
    if (sblocks > 65535) sblocks = 65535;
    bucket_scatter_kernel<<<(int)sblocks, threads, 0, s>>>(g, b);
    int *src = b.table, *dst = tmp;
    for (int pass = 0; pass < b.nlat + b.nlon; ++pass) {
        int left = 0;
        cudaMemsetAsync(remaining, 0, sizeof(int), s);
        bucket_dilate_kernel<<<blocks, threads, 0, s>>>(b, src, dst, remaining);
        if ((e = cudaMemcpyAsync(&left, remaining, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
        int *t = src;
        src = dst;
        dst = t;
        if (left == 0) break;
    }
    if (e == cudaSuccess && src != b.table)
        e = cudaMemcpyAsync(b.table, src, sizeof(int) * (size_t)nb, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) {
        // unit vectors of the cell centres with a halo ring per face (24 bytes per cell); a grid too large for it keeps the
        // sincos path and the general neighbour scan
        double *cxyz = nullptr;
        unsigned char *wide = nullptr;
        int *rim = nullptr;
        b.cxyz = nullptr;
        b.wide = nullptr;
        b.rim = nullptr;
        const size_t ntab = (size_t)g.nface * (g.ny + 2) * (g.nx + 2);
        const size_t nrim = (size_t)g.nface * (2 * (g.nx + 2) + 2 * g.ny);
        if (cudaMalloc(&rim, sizeof(int) * nrim) != cudaSuccess) {
            cudaGetLastError();
        } else if (cudaMalloc(&cxyz, sizeof(double) * 3 * ntab) == cudaSuccess) {
            cell_xyz_kernel<<<(int)sblocks, threads, 0, s>>>(g, cxyz, rim);
            b.cxyz = cxyz;
            b.rim = rim;
            if (cudaMalloc(&wide, (size_t)ncell) == cudaSuccess) {
                cell_wide_kernel<<<(int)sblocks, threads, 0, s>>>(g, cxyz, wide);
                b.wide = wide;
            } else {
                cudaGetLastError();
            }
        } else {
            cudaGetLastError();
            cudaFree(rim);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp);
    cudaFree(remaining);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e;
}

// (XG, YG) pairs for the corner reads of K1 / K2 (GridDev::XYG)
__global__ void interleave_corners_kernel(const float *xg, const float *yg, float2 *out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = make_float2(xg[i], yg[i]);
}

cudaError_t launch_interleave_corners(const float *xg, const float *yg, float2 *out, int64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    interleave_corners_kernel<<<(unsigned)blocks, 256, 0, s>>>(xg, yg, out, n);
    return cudaGetLastError();
}

cudaError_t launch_locate(const LocateArgs &a, cudaStream_t s) {
    if (a.n == 0) return cudaSuccess;
    const int threads = 128;
    locate_kernel<<<(unsigned)((a.n + threads - 1) / threads), threads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_finish_stop(double *t, int32_t *it, const double *midp, int n_midp, double t_stop,
                               const unsigned long long *stats, unsigned long long *total, int64_t n, cudaStream_t s) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    finish_stop_kernel<<<blocks ? blocks : 1, threads, 0, s>>>(t, it, midp, n_midp, t_stop, stats, total, n);
    return cudaGetLastError();
}

cudaError_t launch_time_index(const double *t, const double *midp, int n_midp, int32_t *it, int64_t n,
                              cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int threads = 128;
    time_index_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(t, midp, n_midp, it, n);
    return cudaGetLastError();
}

// OR of the status words of a batch of points (what sdk_locate flagged anywhere): warp reduction, one atomic per warp that saw a bit
__global__ void __launch_bounds__(256) status_or_kernel(const int32_t *status, int64_t n, unsigned int *flags) {
    unsigned int mine = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mine |= (unsigned int)status[i];
    mine = __reduce_or_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicOr(flags, mine);
}

cudaError_t launch_status_or(const int32_t *status, int64_t n, unsigned int *flags, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    status_or_kernel<<<(unsigned)blocks, 256, 0, s>>>(status, n, flags);
    return cudaGetLastError();
}

}  // namespace sdk
