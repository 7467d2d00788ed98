// K6: the regular-grid stencils of the Eulerian budget (reference seaduck/eulerian_budget.py): halo
// ("buffer") gathers across LLC faces / periodic edges / vertical ends, and the 4-point schemes that
// turn a buffered tracer field into the concentration on cell walls -- third-order direct space-time
// (:477-532), second-order flux limiter with the superbee limiter (:345-445) and third-order upwind in
// the vertical (:448).  Pure fp64 arithmetic in the reference's operation order: bit-exact.
#include <cuda_runtime.h>

#include "sdk_device.cuh"

namespace sdk {

// out[o][k] = map[k] < 0 ? 0 : src[o * src_plane + map[k]]   (buffer_x_withface :214 etc.)
// I: uint32_t when every flat index fits (64-bit integer division costs ~10x more instructions)
template <typename I>
__global__ void gather_plane_kernel(const double *src, I n_outer, int64_t src_plane, const int64_t *map, I n_map,
                                    double *out) {
    const I total = n_outer * n_map;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const I o = i / n_map, k = i - o * n_map;
        const int64_t m = __ldg(map + k);
        out[i] = m < 0 ? 0.0 : __ldg(src + (int64_t)o * src_plane + m);
    }
}

__device__ __forceinline__ double sign_of(double x) { return (double)((x > 0) - (x < 0)); }

// _slope_ratio (:345) + _superbee_fluxlimiter (:359)
__device__ __forceinline__ double superbee_of_slope(double rjm, double rj, double rjp, double u, double not_z) {
    double cr = not_z * u > 0 ? rjm : rjp;
    const double cr_max = 1e6;
    if (fabs(rj) * cr_max <= fabs(cr)) cr = sign_of(cr) * sign_of(u) * cr_max;
    else cr = cr / rj;
    return fmax(0.0, fmax(fmin(1.0, 2 * cr), fmin(2.0, cr)));
}

// buf [n_outer][n_out + 3][n_inner], vel / out [n_outer][n_out][n_inner]; scheme 0 DST3, 1 limiter (x, y),
// 2 limiter (z)
template <typename I>
__global__ void wall_stencil_kernel(int scheme, const double *buf, const double *vel, double *out, I n_outer, I n_out,
                                    I n_inner) {
    const I total = n_outer * n_out * n_inner;
    const I slab = n_inner * n_out;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        // position inside buf: the same inner / axis position, in a slab that is 3 * n_inner longer
        const I o = i / slab;
        const double *b = buf + ((int64_t)i + (int64_t)o * 3 * n_inner);
        const double b0 = b[0], b1 = b[n_inner], b2 = b[2 * n_inner], b3 = b[3 * n_inner];
        const double rjm = nan_to_num(b1 - b0), rj = nan_to_num(b2 - b1), rjp = nan_to_num(b3 - b2);
        const double u = vel[i];
        double r;
        if (scheme == 0) {
            const double d0 = (2.0 - u) * (1.0 - u) / 6;
            const double d1 = (1.0 - u * u) / 6;
            r = 0.5 * ((1 + sign_of(u)) * (b1 + (d0 * rj + d1 * rjm)) + (1 - sign_of(u)) * (b2 - (d0 * rj + d1 * rjp)));
        } else {
            const double lim = superbee_of_slope(rjm, rj, rjp, u, scheme == 2 ? -1.0 : 1.0);
            const double mean = nan_to_num(b1 + b2) * 0.5;
            const double corr = sign_of(u) * ((1 - lim) + u * lim) * rj * 0.5;
            r = scheme == 2 ? mean + corr : mean - corr;
        }
        out[i] = r;
    }
}

// third_order_upwind_z (:448): s, out [nz][n_inner]; w [nz][n_inner] with w[0] := 0 first (as the reference does)
template <typename I>
__global__ void upwind3_z_kernel(const double *s, double *w, double *out, I nz_, I n_inner) {
    const I total = nz_ * n_inner;
    const int64_t nz = nz_;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const I k_ = i / n_inner;
        const int64_t in = i - k_ * n_inner, k = k_;
        auto S = [&](int64_t kk) { return s[kk * n_inner + in]; };
        auto RJ = [&](int64_t kk) { kk = (kk + nz) % nz; return kk == 0 ? 0.0 : nan_to_num(S(kk) - S(kk - 1)); };
        const double rj = RJ(k), rjp = RJ(k + 1), rjm = RJ(k - 1);  // np.roll: periodic in k
        const double rjjp = rjp - rj, rjjm = rj - rjm;
        const double ssum = k == 0 ? 0.0 : S(k) + S(k - 1);
        double wk = w[i];
        if (k == 0) {
            wk = 0.0;
            w[i] = 0.0;
        }
        out[i] = 0.5 * (ssum - 1.0 / 6 * (rjjp + rjjm) - sign_of(wk) * 1 / 6 * (rjjp - rjjm));
    }
}

// the grid-stride loop adds up to gridDim * blockDim to an index below `total` before comparing
static bool fits32(int64_t total) { return total + 148 * 16 * 256 < (int64_t)1 << 32; }

static unsigned blocks_for(int64_t total) {
    int64_t b = (total + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

cudaError_t launch_gather_plane(const double *src, int64_t n_outer, int64_t src_plane, const int64_t *map, int64_t n_map,
                                double *out, cudaStream_t s) {
    if (n_outer * n_map == 0) return cudaSuccess;
    if (fits32(n_outer * n_map))
        gather_plane_kernel<uint32_t><<<blocks_for(n_outer * n_map), 256, 0, s>>>(src, (uint32_t)n_outer, src_plane, map, (uint32_t)n_map, out);
    else
        gather_plane_kernel<int64_t><<<blocks_for(n_outer * n_map), 256, 0, s>>>(src, n_outer, src_plane, map, n_map, out);
    return cudaGetLastError();
}

cudaError_t launch_wall_stencil(int scheme, const double *buf, const double *vel, double *out, int64_t n_outer, int64_t n_out,
                                int64_t n_inner, cudaStream_t s) {
    if (n_outer * n_out * n_inner == 0) return cudaSuccess;
    const int64_t total = n_outer * n_out * n_inner;
    if (fits32(n_outer * (n_out + 3) * n_inner))
        wall_stencil_kernel<uint32_t><<<blocks_for(total), 256, 0, s>>>(scheme, buf, vel, out, (uint32_t)n_outer, (uint32_t)n_out, (uint32_t)n_inner);
    else
        wall_stencil_kernel<int64_t><<<blocks_for(total), 256, 0, s>>>(scheme, buf, vel, out, n_outer, n_out, n_inner);
    return cudaGetLastError();
}

cudaError_t launch_upwind3_z(const double *sfield, double *w, double *out, int64_t nz, int64_t n_inner, cudaStream_t s) {
    if (nz * n_inner == 0) return cudaSuccess;
    if (fits32(nz * n_inner))
        upwind3_z_kernel<uint32_t><<<blocks_for(nz * n_inner), 256, 0, s>>>(sfield, w, out, (uint32_t)nz, (uint32_t)n_inner);
    else
        upwind3_z_kernel<int64_t><<<blocks_for(nz * n_inner), 256, 0, s>>>(sfield, w, out, nz, n_inner);
    return cudaGetLastError();
}

}  // namespace sdk
