// K3: masked kernel-weight interpolation at points, K4: wall masks from the centre mask.
//
// Replaces Position.interpolate (reference seaduck/eulerian.py:1268) and the stages it is made of:
// index "fattening" (_fatten_h :386, fatten :520), vector component swaps across faces
// (_transform_vector_and_register :887, topology.py:689), the gathers (_read_data_and_register
// :1064), the mask lookup (_mask_value_and_register :959, get_masks.py:169), the choice of the
// largest all-wet stencil and its Lagrange weights (kernel_weight.py:481-584, :84-428, KnW.get_weight
// :772) and the final dot product -- fused: one thread per point, nothing materialised.
//
// Neighbour indices: naive offsets inside a face; cells closer than `reach` to a face edge read a
// host-built "rim" table, so the device never walks the topology here.
#include <cuda_runtime.h>

#include "interp_args.cuh"

namespace sdk {

__device__ __forceinline__ int rim_slot(int ny, int nx, int r, int dense, int iy, int ix) {
    if (dense) return iy * nx + ix;
    if (iy >= ny - r) return (iy - (ny - r)) * nx + ix;
    if (iy < r) return r * nx + iy * nx + ix;
    if (ix < r) return 2 * r * nx + (iy - r) * r + ix;
    if (ix >= nx - r) return 2 * r * nx + (ny - 2 * r) * r + (iy - r) * r + (ix - (nx - r));
    return -1;
}

// Everything per point lives in registers: the loops over nodes are fully unrolled up to the
// compile-time bound MAXM (9 covers the default 9-point cross, 3 x 3 squares and the particle
// kernels; 32 is the general case), with `m < n_nodes` guards that cost a predicate each.
template <int MAXM>
struct NodeSet {
    int32_t packed[MAXM];  // SDK_PACK(face, iy, ix), SDK_NODE_RAISES or SDK_NODE_LAST
    uint8_t code[MAXM];    // bit0 other component, bit1 negative
};

template <int MAXM>
__device__ __forceinline__ void find_nodes(const GridDev &g, const sdk_kernel_desc &K, int f, int iy, int ix,
                                           NodeSet<MAXM> &ns) {
    const int slot = rim_slot(g.ny, g.nx, K.reach, K.rim_dense, iy, ix);
    if (slot < 0 || K.rim == nullptr) {
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            ns.packed[m] = SDK_PACK(f, iy + K.dy[m], ix + K.dx[m]);
            ns.code[m] = 0;
        }
        return;
    }
    const int per_face = K.rim_dense ? g.ny * g.nx : 2 * K.reach * g.nx + 2 * K.reach * (g.ny - 2 * K.reach);
    const int64_t base = ((int64_t)f * per_face + slot) * K.n_nodes;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
        const bool on = m < K.n_nodes;
        ns.packed[m] = on ? __ldg(K.rim + base + m) : SDK_NODE_RAISES;
        ns.code[m] = (on && K.rim_code) ? __ldg(K.rim_code + base + m) : 0;
    }
}

// SDK_NODE_LAST: the reference walked off a closed box and indexes [-1, -1], numpy's last element
__device__ __forceinline__ int64_t plane_index(int packed, int nface, int fny, int fnx) {
    if (packed == SDK_NODE_LAST) return (int64_t)nface * fny * fnx - 1;
    const int f = packed >> 26 & 0x1f, y = packed >> 13 & 0x1fff, x = packed & 0x1fff;
    return ((int64_t)f * fny + y) * fnx + x;
}

__device__ __forceinline__ double field_value(const GridDev &g, const sdk_field &F, int packed, int kz, int kt) {
    int64_t lev = 0;
    if (F.has_time) lev = wrap_index(kt - F.it0, F.nt);
    if (F.has_z) lev = lev * F.nz + wrap_index(kz, F.nz);
    const int64_t idx = lev * ((int64_t)g.nface * F.ny * F.nx) + plane_index(packed, g.nface, F.ny, F.nx);
    const double v = F.dtype == SDK_F64 ? ldf<double>(F.data, idx) : ldf<float>(F.data, idx);
    return nan_to_num(v);
}

__device__ __forceinline__ bool is_wet(const GridDev &g, const sdk_field &F, int packed, int kz) {
    if (packed == SDK_NODE_RAISES) return false;
    int z = F.has_z ? kz : 0;
    if (z < 0) z += F.mask_nz;
    if (z < 0 || z >= F.mask_nz) return false;
    if (packed != SDK_NODE_LAST) {
        const int y = packed >> 13 & 0x1fff, x = packed & 0x1fff;
        if (y >= F.mask_ny || x >= F.mask_nx) return false;
    }
    const int64_t idx = (int64_t)z * ((int64_t)g.nface * F.mask_ny * F.mask_nx) + plane_index(packed, g.nface, F.mask_ny, F.mask_nx);
    return __ldg(F.mask + idx) != 0;
}

// Lagrange weight of node m in one inheritance level: two polynomials A(rx), C(ry) (Horner on the
// host-expanded coefficients; every lane of a warp that picked the same level reads the same
// coefficients) combined as the node's mode says.  Evaluated on the fly inside the dot product: the
// kernel is bound by the latency of its gathers, so registers (occupancy) matter more than ALU work.
__device__ __forceinline__ double node_weight(const sdk_kernel_desc &K, int level, int m, double rx, double ry) {
    const int nc = K.n_coef;
    const double *ca = K.coef + ((int64_t)level * K.n_nodes + m) * 2 * nc, *cc = ca + nc;
    const int mode = K.level[level].mode[m];
    double a = 0.0, c = 0.0;
    if (mode != SDK_W_C) {
        a = __ldg(ca + nc - 1);
        for (int k = nc - 2; k >= 0; --k) a = a * rx + __ldg(ca + k);
    }
    if (mode != SDK_W_A) {
        c = __ldg(cc + nc - 1);
        for (int k = nc - 2; k >= 0; --k) c = c * ry + __ldg(cc + k);
    }
    return mode == SDK_W_A ? a : (mode == SDK_W_C ? c : (mode == SDK_W_A_PLUS_C ? a + c : a * c));
}

// One interpolated value.  P = field of the kernel's own component, O = the other component
// (vectors on multi-face grids), both may be the same.
template <int MAXM>
__device__ __forceinline__ double interp_value(const GridDev &g, const sdk_kernel_desc &K, const sdk_field &P,
                                               const sdk_field &O, const NodeSet<MAXM> &ns, double rxs, double rys, int kz0,
                                               double rz, int kt0, double rt) {
    const int nz = K.vmode == SDK_MODE_NEAREST ? 1 : 2;
    const int nt = K.tmode == SDK_MODE_NEAREST ? 1 : 2;
    const bool use_mask = !K.ignore_mask && P.mask != nullptr;
    double total = 0.0;
    // Per vertical level: the wet flags of all nodes and (speculatively) their values are requested
    // together, so a point costs one round trip to memory instead of two; the values of nodes that
    // the chosen (smaller) stencil does not use are simply dropped.  Masks do not depend on time.
    int level_z[2] = {0, 0};
    for (int jt = 0; jt < nt; ++jt) {
        const int kt = kt0 + jt;
        const double tw = K.tmode == SDK_MODE_LINEAR ? (jt == 0 ? 1 - rt : rt) : (K.tmode == SDK_MODE_DERIV ? (jt == 0 ? -1.0 : 1.0) : 1.0);
        double zsum[2] = {0.0, 0.0};
        bool znan[2] = {false, false};
        for (int jz = 0; jz < nz; ++jz) {
            const int kz = kz0 - jz;
            double val[MAXM];
            unsigned wet = 0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                val[m] = 0.0;
                if (m < K.n_nodes) {
                    const bool other = ns.code[m] & 1;
                    if (use_mask && jt == 0 && is_wet(g, other ? O : P, ns.packed[m], kz)) wet |= 1u << m;
                    if (ns.packed[m] == SDK_NODE_RAISES) {
                        val[m] = NAN;
                    } else {
                        const double v = field_value(g, other ? O : P, ns.packed[m], kz, kt);
                        val[m] = (ns.code[m] & 2) ? -v : v;
                    }
                }
            }
            if (use_mask && jt == 0) {
                // the largest stencil whose nodes are all wet (reference find_which_points_for_each_kernel :481)
                int level = -1;
                for (int l = K.n_levels - 1; l >= 0; --l)
                    if ((wet & K.level[l].node_mask) == K.level[l].node_mask) level = l;
                level_z[jz] = level;
            }
            const int level = level_z[jz];
            if (level < 0) {
                znan[jz] = true;  // weight[:, 0] = nan (reference :557-558)
                continue;
            }
            const uint32_t mask = K.level[level].node_mask;
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m)
                if (m < K.n_nodes && (mask >> m & 1u)) acc += val[m] * node_weight(K, level, m, rxs, rys);
            zsum[jz] = acc;
        }
        double zw0 = 1.0, zw1 = 0.0;
        if (K.vmode == SDK_MODE_LINEAR) {
            zw0 = 1 - rz;
            zw1 = rz;
            if (K.bottom_no_flux && znan[0]) {
                // ghost point below the bottom (reference kernel_weight.py:860-869)
                znan[0] = false;
                zsum[0] = 0.0;
                if (rz < 0.5) {
                    zsum[1] = 0.0;
                    znan[1] = false;
                }
                zw1 = 1.0;
            }
        } else if (K.vmode == SDK_MODE_DERIV) {
            zw0 = -1.0;
            zw1 = 1.0;
        }
        double part = znan[0] ? NAN : zsum[0] * zw0;
        if (nz == 2) part += znan[1] ? NAN : zsum[1] * zw1;
        total += part * tw;
    }
    return total;
}

#ifndef SDK_INTERP_MIN_BLOCKS
#define SDK_INTERP_MIN_BLOCKS 8
#endif

template <int MAXM>
__global__ void __launch_bounds__(128, MAXM <= 9 ? SDK_INTERP_MIN_BLOCKS : 1) interp_scalar_kernel(const __grid_constant__ InterpArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const int f = p.face ? p.face[i] : 0, iy = p.iy[i], ix = p.ix[i];
    NodeSet<MAXM> ns;
    find_nodes<MAXM>(a.g, a.ku, f, iy, ix, ns);
    const double rxs = a.fu.xstag ? p.rx[i] + 0.5 : p.rx[i];
    const double rys = a.fu.ystag ? p.ry[i] + 0.5 : p.ry[i];
    const double r = interp_value<MAXM>(a.g, a.ku, a.fu, a.fu, ns, rxs, rys, p.kz ? p.kz[i] : 0, p.rz ? p.rz[i] : 0.0,
                                        p.kt ? p.kt[i] : 0, p.rt ? p.rt[i] : 0.0);
    a.out_u[p.out_index ? (int64_t)p.out_index[i] : i] = r;
}

template <int MAXM>
__global__ void __launch_bounds__(128, MAXM <= 9 ? SDK_INTERP_MIN_BLOCKS : 1) interp_vector_kernel(const __grid_constant__ InterpArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const int f = p.face ? p.face[i] : 0, iy = p.iy[i], ix = p.ix[i];
    const int kz = p.kz ? p.kz[i] : 0, kt = p.kt ? p.kt[i] : 0;
    const double rz = p.rz ? p.rz[i] : 0.0, rt = p.rt ? p.rt[i] : 0.0;
    NodeSet<MAXM> ns;
    find_nodes<MAXM>(a.g, a.ku, f, iy, ix, ns);
    double u = interp_value<MAXM>(a.g, a.ku, a.fu, a.fv, ns, p.rx[i] + 0.5, p.ry[i], kz, rz, kt, rt);
    find_nodes<MAXM>(a.g, a.kv, f, iy, ix, ns);
    double v = interp_value<MAXM>(a.g, a.kv, a.fv, a.fu, ns, p.rx[i], p.ry[i] + 0.5, kz, rz, kt, rt);
    if (a.vec_transform) {
        const double cs = (double)p.cs[i], sn = (double)p.sn[i];
        const double uu = u * cs - v * sn, vv = u * sn + v * cs;
        u = uu;
        v = vv;
    }
    const int64_t o = p.out_index ? (int64_t)p.out_index[i] : i;
    a.out_u[o] = u;
    a.out_v[o] = v;
}

// Several scalar variables at once (T and S with one KnW, say): node indices, wet flags, the choice of the stencil and
// the Lagrange weights depend on the points and the kernel only -- they are worked out once, as a weight per node and
// vertical level, and every field then costs its gathers and a dot product.  TWO: vertical or temporal kernels with two
// nodes (otherwise the second set of weights does not exist and takes no registers).
template <int MAXM, bool TWO>
__global__ void __launch_bounds__(128, MAXM <= 9 ? (TWO ? 4 : 6) : 1) interp_scalar_multi_kernel(const __grid_constant__ InterpMultiArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const sdk_kernel_desc &K = a.k;
    const sdk_field &F0 = a.f[0];
    const int f = p.face ? p.face[i] : 0, iy = p.iy[i], ix = p.ix[i];
    NodeSet<MAXM> ns;
    find_nodes<MAXM>(a.g, K, f, iy, ix, ns);
    const double rxs = F0.xstag ? p.rx[i] + 0.5 : p.rx[i];
    const double rys = F0.ystag ? p.ry[i] + 0.5 : p.ry[i];
    const int kz0 = p.kz ? p.kz[i] : 0, kt0 = p.kt ? p.kt[i] : 0;
    const double rz = p.rz ? p.rz[i] : 0.0, rt = p.rt ? p.rt[i] : 0.0;
    const int nz = (TWO && K.vmode != SDK_MODE_NEAREST) ? 2 : 1;
    const int nt = (TWO && K.tmode != SDK_MODE_NEAREST) ? 2 : 1;
    const bool use_mask = !K.ignore_mask && F0.mask != nullptr;
    constexpr int NZ = TWO ? 2 : 1;
    double w[NZ][MAXM];
    uint32_t use[NZ];  // nodes of the stencil chosen for this vertical level
    bool znan[NZ];
#pragma unroll
    for (int jz = 0; jz < NZ; ++jz) {
        znan[jz] = false;
        int level = 0;
        if (jz < nz && use_mask) {
            unsigned wet = 0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m)
                if (m < K.n_nodes && is_wet(a.g, F0, ns.packed[m], kz0 - jz)) wet |= 1u << m;
            level = -1;
            for (int l = K.n_levels - 1; l >= 0; --l)
                if ((wet & K.level[l].node_mask) == K.level[l].node_mask) level = l;
        }
        znan[jz] = level < 0;
        const uint32_t mask = (level < 0 || jz >= nz) ? 0u : K.level[level].node_mask;
        use[jz] = mask;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) w[jz][m] = (m < K.n_nodes && (mask >> m & 1u)) ? node_weight(K, level, m, rxs, rys) : 0.0;
    }
    // vertical factors and the ghost point below the bottom (same rule as interp_value)
    double zw[2] = {1.0, 0.0};
    bool kill[2] = {false, false};
    if (K.vmode == SDK_MODE_LINEAR) {
        zw[0] = 1 - rz, zw[1] = rz;
        if (TWO && K.bottom_no_flux && znan[0]) {
            znan[0] = false, kill[0] = true;
            if (rz < 0.5) znan[NZ - 1] = false, kill[1] = true;
            zw[1] = 1.0;
        }
    } else if (K.vmode == SDK_MODE_DERIV) {
        zw[0] = -1.0, zw[1] = 1.0;
    }
    const int64_t o = p.out_index ? (int64_t)p.out_index[i] : i;
    for (int k = 0; k < a.n_fields; ++k) {
        const sdk_field &F = a.f[k];
        double total = 0.0;
        for (int jt = 0; jt < nt; ++jt) {
            const double tw = K.tmode == SDK_MODE_LINEAR ? (jt == 0 ? 1 - rt : rt) : (K.tmode == SDK_MODE_DERIV ? (jt == 0 ? -1.0 : 1.0) : 1.0);
            double part = 0.0;
#pragma unroll
            for (int jz = 0; jz < NZ; ++jz) {
                if (jz >= nz) continue;
                double acc = 0.0;
                if (!kill[jz]) {
#pragma unroll
                    for (int m = 0; m < MAXM; ++m) {
                        // (same expression as the single-field kernel: value times weight, summed over the stencil's nodes
                        // in node order; nodes outside the chosen stencil are not read)
                        if (m < K.n_nodes && (use[jz] >> m & 1u)) {
                            const double v = ns.packed[m] == SDK_NODE_RAISES ? NAN : field_value(a.g, F, ns.packed[m], kz0 - jz, kt0 + jt);
                            acc += v * w[jz][m];
                        }
                    }
                }
                part += znan[jz] ? NAN : acc * zw[jz];
            }
            total += part * tw;
        }
        a.out[k][o] = total;
    }
}

// K4: a dry cell's west / south wall is wet when the cell on the other side is wet; a W point is
// wet when the cell below or above it is (reference get_masks.py:12,46,79).
__global__ void build_masks_kernel(GridDev g, const uint8_t *maskC, int nz, uint8_t *maskU, uint8_t *maskV,
                                   uint8_t *maskW) {
    const int64_t plane = (int64_t)g.nface * g.ny * g.nx;
    const int64_t total = plane * nz;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(c / plane);
        const int64_t h = c - (int64_t)z * plane;
        const int f = (int)(h / (g.ny * g.nx));
        const int rem = (int)(h - (int64_t)f * g.ny * g.nx);
        const int iy = rem / g.nx, ix = rem - iy * g.nx;
        const uint8_t self = maskC[c] != 0;
        uint8_t mu = self, mv = self;
        if (!self) {
            for (int dir = 2; dir >= 1; --dir) {  // 2: left neighbour -> U, 1: lower neighbour -> V
                int nf = f, ny = iy, nx = ix, came;
                const int st = cell_move(g, dir, nf, ny, nx, came);
                uint8_t other = 0;
                if (st == 0) other = maskC[(int64_t)z * plane + ((int64_t)nf * g.ny + ny) * g.nx + nx] != 0;
                else if (st & SDK_ST_OUT) other = maskC[(int64_t)z * plane + plane - 1] != 0;  // numpy index -1,-1
                if (dir == 2) mu = other;
                else mv = other;
            }
        }
        maskU[c] = mu;
        maskV[c] = mv;
        maskW[c] = self || (z > 0 && maskC[c - plane] != 0);
    }
}

cudaError_t launch_interp(const InterpArgs &a, bool vector, cudaStream_t s) {
    if (a.pts.n == 0) return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.pts.n + threads - 1) / threads);
    const bool small = a.ku.n_nodes <= 9 && (!vector || a.kv.n_nodes <= 9);
    if (vector) {
        if (small) interp_vector_kernel<9><<<blocks, threads, 0, s>>>(a);
        else interp_vector_kernel<SDK_MAX_NODES><<<blocks, threads, 0, s>>>(a);
    } else {
        if (small) interp_scalar_kernel<9><<<blocks, threads, 0, s>>>(a);
        else interp_scalar_kernel<SDK_MAX_NODES><<<blocks, threads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_interp_multi(const InterpMultiArgs &a, cudaStream_t s) {
    if (a.pts.n == 0) return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.pts.n + threads - 1) / threads);
    const bool two = a.k.vmode != SDK_MODE_NEAREST || a.k.tmode != SDK_MODE_NEAREST;
    if (a.k.n_nodes <= 9) {
        if (two) interp_scalar_multi_kernel<9, true><<<blocks, threads, 0, s>>>(a);
        else interp_scalar_multi_kernel<9, false><<<blocks, threads, 0, s>>>(a);
    } else {
        if (two) interp_scalar_multi_kernel<SDK_MAX_NODES, true><<<blocks, threads, 0, s>>>(a);
        else interp_scalar_multi_kernel<SDK_MAX_NODES, false><<<blocks, threads, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_build_masks(const GridDev &g, const uint8_t *maskC, int nz, uint8_t *mu, uint8_t *mv, uint8_t *mw,
                               cudaStream_t s) {
    const int64_t total = (int64_t)g.nface * g.ny * g.nx * nz;
    if (total == 0) return cudaSuccess;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    build_masks_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, maskC, nz, mu, mv, mw);
    return cudaGetLastError();
}

}  // namespace sdk
