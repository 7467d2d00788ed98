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

// unroll factor of the loops over the stencil's nodes (1 = rolled; see the note above `struct Stencil`)
#ifndef SDK_INTERP_UNROLL
#define SDK_INTERP_UNROLL 1
#endif
#define SDK_STR2(x) #x
#define SDK_STR(x) SDK_STR2(x)
#define SDK_NODE_LOOP _Pragma(SDK_STR(unroll SDK_INTERP_UNROLL))

__device__ __forceinline__ int rim_slot(int ny, int nx, int r, int dense, int iy, int ix) {
    if (dense) return iy * nx + ix;
    if (iy >= ny - r) return (iy - (ny - r)) * nx + ix;
    if (iy < r) return r * nx + iy * nx + ix;
    if (ix < r) return 2 * r * nx + (iy - r) * r + ix;
    if (ix >= nx - r) return 2 * r * nx + (ny - 2 * r) * r + (iy - r) * r + (ix - (nx - r));
    return -1;
}

// The stencil of a point, node by node.  Away from the face edges a node is the cell plus the kernel's offset; cells closer
// than `reach` to an edge read the host-built rim table (index and, for vectors, the component / sign code).
//
// The kernels below loop over the nodes instead of unrolling them into register arrays.  Round 2's profile of the unrolled
// version (profiles/r02b_interp_*): 4000 static / 3460 executed warp instructions per 32 points, most of them index
// arithmetic repeated per node and per field, the SM instruction cache missing 16 % of its requests and the GPC
// instruction cache at 84 % of its request rate -- a kernel bound by instruction supply like K1, not by its gathers.
// Rolled, the body is a few hundred instructions, per-node state lives in a handful of registers (more warps per SM to
// hide the gathers' latency instead of more loads per thread), and the per-node work is done once for all the fields
// that share it.  Sums run over the nodes in ascending order exactly as before: results are bit-identical.
struct Stencil {
    int f, iy, ix;
    int slot;          // rim slot of the cell, -1 = interior
    int64_t rim_base;  // first entry of the cell's row in the rim table
};

__device__ __forceinline__ Stencil stencil_of(const GridDev &g, const sdk_kernel_desc &K, int f, int iy, int ix) {
    Stencil st;
    st.f = f, st.iy = iy, st.ix = ix;
    st.slot = K.rim ? rim_slot(g.ny, g.nx, K.reach, K.rim_dense, iy, ix) : -1;
    const int per_face = K.rim_dense ? g.ny * g.nx : 2 * K.reach * g.nx + 2 * K.reach * (g.ny - 2 * K.reach);
    st.rim_base = st.slot < 0 ? 0 : ((int64_t)f * per_face + st.slot) * K.n_nodes;
    return st;
}

// node m: SDK_PACK(face, iy, ix), SDK_NODE_RAISES or SDK_NODE_LAST; code bit0 = other component, bit1 = negative
__device__ __forceinline__ int node_of(const sdk_kernel_desc &K, const Stencil &st, int m, int &code) {
    if (st.slot < 0) {
        code = 0;
        return SDK_PACK(st.f, st.iy + K.dy[m], st.ix + K.dx[m]);
    }
    code = K.rim_code ? __ldg(K.rim_code + st.rim_base + m) : 0;
    return __ldg(K.rim + st.rim_base + m);
}

// SDK_NODE_LAST: the reference walked off a closed box and indexes [-1, -1], numpy's last element
__device__ __forceinline__ int64_t plane_index(int packed, int nface, int fny, int fnx) {
    if (packed == SDK_NODE_LAST) return (int64_t)nface * fny * fnx - 1;
    const int f = packed >> 26 & 0x1f, y = packed >> 13 & 0x1fff, x = packed & 0x1fff;
    return ((int64_t)f * fny + y) * fnx + x;
}

// element index of a node in a field at vertical / time level (kz, kt) -- the same for every field of that shape
__device__ __forceinline__ int64_t field_index(const GridDev &g, const sdk_field &F, int packed, int kz, int kt) {
    int64_t lev = 0;
    if (F.has_time) lev = wrap_index(kt - F.it0, F.nt);
    if (F.has_z) lev = lev * F.nz + wrap_index(kz, F.nz);
    return lev * ((int64_t)g.nface * F.ny * F.nx) + plane_index(packed, g.nface, F.ny, F.nx);
}

__device__ __forceinline__ double field_load(const sdk_field &F, int64_t idx) {
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
// coefficients) combined as the node's mode says.
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

// the largest stencil of the inheritance whose nodes are all wet (reference find_which_points_for_each_kernel :481); -1: none
__device__ __forceinline__ int pick_level(const sdk_kernel_desc &K, unsigned wet) {
    int level = -1;
    for (int l = K.n_levels - 1; l >= 0; --l)
        if ((wet & K.level[l].node_mask) == K.level[l].node_mask) level = l;
    return level;
}

// vertical factors of the two levels, and the ghost point below the bottom (reference kernel_weight.py:860-869):
// `kill[j]` drops level j's sum (it counts as zero), `znan[j]` makes it NaN
__device__ __forceinline__ void vertical_rule(const sdk_kernel_desc &K, double rz, bool two, double zw[2], bool znan[2], bool kill[2]) {
    zw[0] = 1.0, zw[1] = 0.0;
    kill[0] = kill[1] = false;
    if (K.vmode == SDK_MODE_LINEAR) {
        zw[0] = 1 - rz, zw[1] = rz;
        if (two && K.bottom_no_flux && znan[0]) {
            znan[0] = false, kill[0] = true;
            if (rz < 0.5) znan[1] = false, kill[1] = true;
            zw[1] = 1.0;
        }
    } else if (K.vmode == SDK_MODE_DERIV) {
        zw[0] = -1.0, zw[1] = 1.0;
    }
}

__device__ __forceinline__ double time_factor(const sdk_kernel_desc &K, int jt, double rt) {
    return K.tmode == SDK_MODE_LINEAR ? (jt == 0 ? 1 - rt : rt) : (K.tmode == SDK_MODE_DERIV ? (jt == 0 ? -1.0 : 1.0) : 1.0);
}

// One interpolated value.  P = field of the kernel's own component, O = the other component (vectors on multi-face
// grids; a node's wet flag and value come from the component the rim code names), both may be the same.
// (out of line: the vector kernel calls it twice, and one copy of the body is all the instruction cache should hold)
static __device__ __noinline__ double interp_value(const GridDev &g, const sdk_kernel_desc &K, const sdk_field &P, const sdk_field &O,
                                                    int f, int iy, int ix, double rxs, double rys, int kz0, double rz, int kt0, double rt) {
    const Stencil st = stencil_of(g, K, f, iy, ix);
    const int nz = K.vmode == SDK_MODE_NEAREST ? 1 : 2;
    const int nt = K.tmode == SDK_MODE_NEAREST ? 1 : 2;
    const bool use_mask = !K.ignore_mask && P.mask != nullptr;
    // pass 1: wet flags of all nodes on both vertical levels (masks do not depend on time) -> the stencil of each level
    int level[2] = {0, 0};
    if (use_mask) {
        unsigned wet[2] = {0u, 0u};
SDK_NODE_LOOP
        for (int m = 0; m < K.n_nodes; ++m) {
            int code;
            const int packed = node_of(K, st, m, code);
            const sdk_field &F = (code & 1) ? O : P;
            if (is_wet(g, F, packed, kz0)) wet[0] |= 1u << m;
            if (nz == 2 && is_wet(g, F, packed, kz0 - 1)) wet[1] |= 1u << m;
        }
        level[0] = pick_level(K, wet[0]);
        level[1] = nz == 2 ? pick_level(K, wet[1]) : 0;
    }
    bool znan[2] = {level[0] < 0, nz == 2 && level[1] < 0};
    const unsigned use[2] = {znan[0] ? 0u : K.level[level[0]].node_mask, (nz < 2 || znan[1]) ? 0u : K.level[level[1]].node_mask};
    // pass 2: value times weight, summed over the nodes of the level's stencil in node order, per (time, vertical) level
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // [jt][jz]
SDK_NODE_LOOP
    for (int m = 0; m < K.n_nodes; ++m) {
        if (!((use[0] | use[1]) >> m & 1u)) continue;
        int code;
        const int packed = node_of(K, st, m, code);
        const sdk_field &F = (code & 1) ? O : P;
#pragma unroll
        for (int jz = 0; jz < 2; ++jz) {
            if (!(use[jz] >> m & 1u)) continue;
            const double w = node_weight(K, level[jz], m, rxs, rys);
#pragma unroll
            for (int jt = 0; jt < 2; ++jt) {
                if (jt >= nt) continue;
                double v = NAN;
                if (packed != SDK_NODE_RAISES) {
                    v = field_load(F, field_index(g, F, packed, kz0 - jz, kt0 + jt));
                    if (code & 2) v = -v;
                }
                acc[jt][jz] += v * w;
            }
        }
    }
    double zw[2];
    bool kill[2];
    vertical_rule(K, rz, nz == 2, zw, znan, kill);
    double total = 0.0;
    for (int jt = 0; jt < nt; ++jt) {
        double part = znan[0] ? NAN : (kill[0] ? 0.0 : acc[jt][0]) * zw[0];
        if (nz == 2) part += znan[1] ? NAN : (kill[1] ? 0.0 : acc[jt][1]) * zw[1];
        total += part * time_factor(K, jt, rt);
    }
    return total;
}

#ifndef SDK_INTERP_MIN_BLOCKS
#define SDK_INTERP_MIN_BLOCKS 8
#endif

__global__ void __launch_bounds__(128, SDK_INTERP_MIN_BLOCKS) interp_scalar_kernel(const __grid_constant__ InterpArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const double rxs = a.fu.xstag ? p.rx[i] + 0.5 : p.rx[i];
    const double rys = a.fu.ystag ? p.ry[i] + 0.5 : p.ry[i];
    const double r = interp_value(a.g, a.ku, a.fu, a.fu, p.face ? p.face[i] : 0, p.iy[i], p.ix[i], rxs, rys, p.kz ? p.kz[i] : 0, p.rz ? p.rz[i] : 0.0, p.kt ? p.kt[i] : 0,
                                  p.rt ? p.rt[i] : 0.0);
    a.out_u[p.out_index ? (int64_t)p.out_index[i] : i] = r;
}

__global__ void __launch_bounds__(128, SDK_INTERP_MIN_BLOCKS) interp_vector_kernel(const __grid_constant__ InterpArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const int f = p.face ? p.face[i] : 0, iy = p.iy[i], ix = p.ix[i];
    const int kz = p.kz ? p.kz[i] : 0, kt = p.kt ? p.kt[i] : 0;
    const double rz = p.rz ? p.rz[i] : 0.0, rt = p.rt ? p.rt[i] : 0.0;
    const double rx = p.rx[i], ry = p.ry[i];
    double u = interp_value(a.g, a.ku, a.fu, a.fv, f, iy, ix, rx + 0.5, ry, kz, rz, kt, rt);
    double v = interp_value(a.g, a.kv, a.fv, a.fu, f, iy, ix, rx, ry + 0.5, kz, rz, kt, rt);
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

// Several scalar variables at once (T and S with one KnW, say): wet flags, the choice of the stencil, the Lagrange weight
// and the element index of a node depend on the points and the kernel only -- they are worked out once per node, and every
// field then costs one gather and one multiply-add.  TWO: vertical or temporal kernels with two nodes (otherwise the
// second set of sums does not exist and takes no registers).
template <bool TWO>
__global__ void __launch_bounds__(128, TWO ? 6 : SDK_INTERP_MIN_BLOCKS) interp_scalar_multi_kernel(const __grid_constant__ InterpMultiArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pts.n) return;
    const sdk_points &p = a.pts;
    const sdk_kernel_desc &K = a.k;
    const sdk_field &F0 = a.f[0];
    const Stencil st = stencil_of(a.g, K, p.face ? p.face[i] : 0, p.iy[i], p.ix[i]);
    const double rxs = F0.xstag ? p.rx[i] + 0.5 : p.rx[i];
    const double rys = F0.ystag ? p.ry[i] + 0.5 : p.ry[i];
    const int kz0 = p.kz ? p.kz[i] : 0, kt0 = p.kt ? p.kt[i] : 0;
    const double rz = p.rz ? p.rz[i] : 0.0, rt = p.rt ? p.rt[i] : 0.0;
    constexpr int N2 = TWO ? 2 : 1;
    const int nz = (TWO && K.vmode != SDK_MODE_NEAREST) ? 2 : 1;
    const int nt = (TWO && K.tmode != SDK_MODE_NEAREST) ? 2 : 1;
    const bool use_mask = !K.ignore_mask && F0.mask != nullptr;
    int level[2] = {0, 0};
    if (use_mask) {
        unsigned wet[2] = {0u, 0u};
SDK_NODE_LOOP
        for (int m = 0; m < K.n_nodes; ++m) {
            int code;
            const int packed = node_of(K, st, m, code);
            if (is_wet(a.g, F0, packed, kz0)) wet[0] |= 1u << m;
            if (TWO && nz == 2 && is_wet(a.g, F0, packed, kz0 - 1)) wet[1] |= 1u << m;
        }
        level[0] = pick_level(K, wet[0]);
        level[1] = nz == 2 ? pick_level(K, wet[1]) : 0;
    }
    bool znan[2] = {level[0] < 0, nz == 2 && level[1] < 0};
    const unsigned use[2] = {znan[0] ? 0u : K.level[level[0]].node_mask, (nz < 2 || znan[1]) ? 0u : K.level[level[1]].node_mask};
    double acc[SDK_MAX_MULTI_FIELDS][N2][N2];  // [field][jt][jz]
#pragma unroll
    for (int k = 0; k < SDK_MAX_MULTI_FIELDS; ++k)
#pragma unroll
        for (int jt = 0; jt < N2; ++jt)
#pragma unroll
            for (int jz = 0; jz < N2; ++jz) acc[k][jt][jz] = 0.0;
SDK_NODE_LOOP
    for (int m = 0; m < K.n_nodes; ++m) {
        if (!((use[0] | use[1]) >> m & 1u)) continue;
        int code;
        const int packed = node_of(K, st, m, code);
#pragma unroll
        for (int jz = 0; jz < N2; ++jz) {
            if (!(use[jz] >> m & 1u)) continue;
            const double w = node_weight(K, level[jz], m, rxs, rys);
#pragma unroll
            for (int jt = 0; jt < N2; ++jt) {
                if (jt >= nt) continue;
                // (the fields share shape, staggering and time / vertical layout: one index for all of them)
                const int64_t idx = packed == SDK_NODE_RAISES ? 0 : field_index(a.g, F0, packed, kz0 - jz, kt0 + jt);
#pragma unroll
                for (int k = 0; k < SDK_MAX_MULTI_FIELDS; ++k) {
                    if (k >= a.n_fields) continue;
                    const double v = packed == SDK_NODE_RAISES ? NAN : field_load(a.f[k], idx);
                    acc[k][jt][jz] += v * w;
                }
            }
        }
    }
    double zw[2];
    bool kill[2];
    vertical_rule(K, rz, TWO, zw, znan, kill);
    const int64_t o = p.out_index ? (int64_t)p.out_index[i] : i;
#pragma unroll
    for (int k = 0; k < SDK_MAX_MULTI_FIELDS; ++k) {
        if (k >= a.n_fields) continue;
        double total = 0.0;
#pragma unroll
        for (int jt = 0; jt < N2; ++jt) {
            if (jt >= nt) continue;
            double part = znan[0] ? NAN : (kill[0] ? 0.0 : acc[k][jt][0]) * zw[0];
            if (nz == 2) part += znan[1] ? NAN : (kill[1] ? 0.0 : acc[k][jt][N2 - 1]) * zw[1];
            total += part * time_factor(K, jt, rt);
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
    if (vector) interp_vector_kernel<<<blocks, threads, 0, s>>>(a);
    else interp_scalar_kernel<<<blocks, threads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_interp_multi(const InterpMultiArgs &a, cudaStream_t s) {
    if (a.pts.n == 0) return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.pts.n + threads - 1) / threads);
    if (a.k.vmode != SDK_MODE_NEAREST || a.k.tmode != SDK_MODE_NEAREST) interp_scalar_multi_kernel<true><<<blocks, threads, 0, s>>>(a);
    else interp_scalar_multi_kernel<false><<<blocks, threads, 0, s>>>(a);
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
