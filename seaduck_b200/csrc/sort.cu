// Cell order for particles: lanes of a warp that sit in neighbouring cells read neighbouring
// memory -- the difference between one 32-byte sector per 8-byte gather and a shared one.  The reference
// has no such step (numpy does not care); its nearest primitive is Position.subset / update_from_subset
// (seaduck/eulerian.py:319-384), a gather / scatter of every per-particle attribute, which is what
// sdk_particles_gather does for the device state.
//
// Key = ((level * nface + face) * ny + iy) * nx + ix, sorted with CUB's radix sort over the significant bits
// only; the value is the particle's index, so the result is the permutation itself.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cstdint>

#include "locate_args.cuh"

namespace sdk {

__global__ void __launch_bounds__(256) cell_key_kernel(int64_t n, int nface, int ny, int nx, const int32_t *face, const int32_t *iy,
                                                       const int32_t *ix, const int32_t *izl, uint64_t *key, int32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t f = face ? (uint64_t)face[i] : 0, z = izl ? (uint64_t)izl[i] : 0;
    key[i] = ((z * nface + f) * ny + (uint64_t)iy[i]) * nx + (uint64_t)ix[i];
    idx[i] = (int32_t)i;
}

static int bits_for(uint64_t top) {
    int bits = 1;
    while (bits < 64 && (top >> bits) != 0) ++bits;
    return bits;
}

static int key_bits(const GridDev &g) { return bits_for((uint64_t)(g.n_zl > 0 ? g.n_zl + 1 : 1) * g.nface * g.ny * g.nx); }

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cell_sort_scratch_bytes(int64_t n) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)n, 0, 64);
    return align256(tmp) + 2 * align256((size_t)n * 8) + align256((size_t)n * 4);
}

// the scratch buffer of a sort: keys in, keys out, indices in, CUB's workspace
struct SortScratch {
    uint64_t *key_in, *key_out;
    int32_t *idx_in;
    void *tmp;
    size_t tmp_bytes;
};

static SortScratch carve_sort(int64_t n, void *scratch, size_t scratch_bytes) {
    SortScratch w;
    char *cur = (char *)scratch;
    w.key_in = (uint64_t *)cur;
    cur += align256((size_t)n * 8);
    w.key_out = (uint64_t *)cur;
    cur += align256((size_t)n * 8);
    w.idx_in = (int32_t *)cur;
    cur += align256((size_t)n * 4);
    w.tmp = cur;
    w.tmp_bytes = scratch_bytes - (size_t)(cur - (char *)scratch);
    return w;
}

cudaError_t launch_cell_sort(const GridDev &g, int64_t n, const int32_t *face, const int32_t *iy, const int32_t *ix,
                             const int32_t *izl, int32_t *perm, void *scratch, size_t scratch_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const SortScratch w = carve_sort(n, scratch, scratch_bytes);
    cell_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, g.nface, g.ny, g.nx, face, iy, ix, izl, w.key_in, w.idx_in);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(w.tmp, const_cast<size_t &>(w.tmp_bytes), w.key_in, w.key_out, w.idx_in, perm, (int)n, 0, key_bits(g), s);
}

// Order for points that have not been located yet (sdk_track_host): the key is (vertical level, bucket of the
// locator's lat / lon grid) -- about one cell per bucket, so it is the cell order up to the walk of sdk_locate -- and it
// only needs the coordinates.  Sorting the 32 bytes of input of a point before locating it is cheaper than sorting
// the 220 bytes of state afterwards, and sdk_locate itself runs on coherent points.
cudaError_t launch_point_keys(const GridDev &g, const BucketGrid &b, int64_t n, const double *lon, const double *lat, const double *dep,
                              uint64_t *key, int32_t *idx, cudaStream_t s);  // locate.cu

cudaError_t launch_point_sort(const GridDev &g, const BucketGrid &b, int64_t n, const double *lon, const double *lat, const double *dep,
                              int32_t *perm, void *scratch, size_t scratch_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const SortScratch w = carve_sort(n, scratch, scratch_bytes);
    cudaError_t e = launch_point_keys(g, b, n, lon, lat, dep, w.key_in, w.idx_in, s);
    if (e != cudaSuccess) return e;
    const int bits = bits_for((uint64_t)(g.n_zl > 0 ? g.n_zl + 1 : 1) * (uint64_t)b.nlat * (uint64_t)b.nlon);
    return cub::DeviceRadixSort::SortPairs(w.tmp, const_cast<size_t &>(w.tmp_bytes), w.key_in, w.key_out, w.idx_in, perm, (int)n, 0, bits, s);
}

// out[k] = in[perm[k]] for the four coordinate arrays of a batch of points
__global__ void __launch_bounds__(256) points_gather_kernel(int64_t n, const int32_t *perm, const double *lon, const double *lat,
                                                            const double *dep, const double *t, double *olon, double *olat, double *odep,
                                                            double *ot) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = perm[k];
    olon[k] = lon[i];
    olat[k] = lat[i];
    if (dep) odep[k] = dep[i];
    if (t) ot[k] = t[i];
}

cudaError_t launch_points_gather(int64_t n, const int32_t *perm, const double *lon, const double *lat, const double *dep, const double *t,
                                 double *olon, double *olat, double *odep, double *ot, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    points_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, perm, lon, lat, dep, t, olon, olat, odep, ot);
    return cudaGetLastError();
}

// dst[k] = src[perm[k]] for every array both structures hold: one pass over the state, one thread per particle
__global__ void __launch_bounds__(256) particles_gather_kernel(sdk_particles src, sdk_particles dst, const int32_t *perm) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dst.n) return;
    const int64_t i = perm[k];
#define MOVE(m) \
    if (src.m && dst.m) dst.m[k] = src.m[i];
    MOVE(lon) MOVE(lat) MOVE(dep) MOVE(t)
    MOVE(face) MOVE(iy) MOVE(ix) MOVE(izl) MOVE(it)
    MOVE(rx) MOVE(ry) MOVE(rzl)
    MOVE(u) MOVE(v) MOVE(w) MOVE(du) MOVE(dv) MOVE(dw)
    MOVE(dx) MOVE(dy) MOVE(bzl) MOVE(dzl) MOVE(vol)
    MOVE(status) MOVE(niter) MOVE(ncross)
#undef MOVE
    if (src.px && dst.px) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst.px[j * dst.n + k] = src.px[j * src.n + i];
    }
    if (src.py && dst.py) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst.py[j * dst.n + k] = src.py[j * src.n + i];
    }
}

cudaError_t launch_particles_gather(const sdk_particles &src, const sdk_particles &dst, const int32_t *perm, cudaStream_t s) {
    if (dst.n == 0) return cudaSuccess;
    particles_gather_kernel<<<(unsigned)((dst.n + 255) / 256), 256, 0, s>>>(src, dst, perm);
    return cudaGetLastError();
}

}  // namespace sdk
