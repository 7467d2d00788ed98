// Arguments of the locate kernels (shared by locate.cu and capi.cu).
#pragma once

#include <climits>
#include <cuda_runtime.h>

#include "sdk_device.cuh"

namespace sdk {

struct BucketGrid {
    double lon0, lat0, span_lon, span_lat, inv_dlon, inv_dlat;
    int nlon, nlat, periodic;
    int *table;  // one nearby cell (flat index) per bucket
    // unit vector of every cell centre (3 doubles per cell, +inf for a NaN centre) as [nface][ny + 2][nx + 2][3]: the ring
    // around a face holds the centres of the cells across its edges, so the 8 neighbours of any cell are index offsets; NULL
    // when the table did not fit: a distance is then two sincos and the neighbours come from cell_move
    const double *cxyz;
    // (with cxyz) flat index of the cell at every halo position, -1 = none; per face: bottom row, top row (nx + 2 each),
    // left column, right column (ny each)
    const int *rim;
    // one byte per cell (with cxyz): 1 where the local grid vectors are sheared / stretched enough (or a neighbour is
    // missing) that a cell outside the 8 index neighbours can be the nearest one, so the walk has to look at the second ring
    const unsigned char *wide;
};

struct LocateArgs {
    GridDev g;
    BucketGrid b;
    int64_t n;
    const double *lon, *lat, *dep, *t, *ts;
    int n_t, max_walk;
    sdk_located out;
};

cudaError_t build_buckets(const GridDev &g, BucketGrid &b, cudaStream_t s);
cudaError_t launch_interleave_corners(const float *xg, const float *yg, float2 *out, int64_t n, cudaStream_t s);
cudaError_t launch_locate(const LocateArgs &a, cudaStream_t s);
cudaError_t launch_finish_stop(double *t, int32_t *it, const double *midp, int n_midp, double t_stop,
                               const unsigned long long *stats, unsigned long long *total, int64_t n, cudaStream_t s);
cudaError_t launch_time_index(const double *t, const double *midp, int n_midp, int32_t *it, int64_t n,
                              cudaStream_t s);
cudaError_t launch_status_or(const int32_t *status, int64_t n, unsigned int *flags, cudaStream_t s);

}  // namespace sdk
