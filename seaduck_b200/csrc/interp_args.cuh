// Arguments of the interpolation / mask kernels (shared by interp.cu and capi.cu).
#pragma once

#include <cuda_runtime.h>

#include "sdk_device.cuh"

namespace sdk {

struct InterpArgs {
    GridDev g;
    sdk_points pts;
    sdk_field fu, fv;
    sdk_kernel_desc ku, kv;
    int vec_transform;
    double *out_u, *out_v;
};

// several scalar variables that share the kernel, the grid position, the mask and the vertical / temporal layout
constexpr int kMaxMultiFields = 4;
struct InterpMultiArgs {
    GridDev g;
    sdk_points pts;
    sdk_kernel_desc k;
    int n_fields;
    sdk_field f[kMaxMultiFields];
    double *out[kMaxMultiFields];
};

cudaError_t launch_interp(const InterpArgs &a, bool vector, cudaStream_t s);
cudaError_t launch_interp_multi(const InterpMultiArgs &a, cudaStream_t s);
cudaError_t launch_build_masks(const GridDev &g, const uint8_t *maskC, int nz, uint8_t *mu, uint8_t *mv, uint8_t *mw,
                               cudaStream_t s);

}  // namespace sdk
