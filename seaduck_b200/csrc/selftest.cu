// Device self-test hooks: expose the branch-free math helpers of fastmath.cuh so that the tests can
// compare them with IEEE division and with 80-bit log / exp references (tests/test_gpu_math.py).
#include <cuda_runtime.h>

#include "../../include/seaduck_b200.h"
#include "fastmath.cuh"

namespace sdk {

__global__ void selftest_math_kernel(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool flag = false;
    double r;
    if (op == SDK_SELFTEST_DIV) r = fdiv(a[i], b[i], flag);
    else if (op == SDK_SELFTEST_LOG) r = flog(a[i], flag);
    else if (op == SDK_SELFTEST_EXP) r = fexp(a[i], flag);
    else if (op == SDK_SELFTEST_DIV_IEEE) r = a[i] / b[i];
    else if (op == SDK_SELFTEST_LOG_CUDA) r = log(a[i]);
    else r = exp(a[i]);
    out[i] = r;
    bad[i] = flag;
}

__global__ void pack_snapshot_kernel(int64_t n, int64_t n_pad, const double *lon, const double *lat, const double *dep,
                                     const int32_t *face, const int32_t *iy, const int32_t *ix, const int32_t *izl,
                                     double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const bool in = i < n;
    out[i] = in && lon ? lon[i] : 0.0;
    out[n_pad + i] = in && lat ? lat[i] : 0.0;
    out[2 * n_pad + i] = in && dep ? dep[i] : 0.0;
    // face (6 bits) | iy (13) | ix (13) | izl (8) in one word: 32 bytes per particle on the wire
    const unsigned long long f = in && face ? (unsigned)face[i] : 0u, y = in && iy ? (unsigned)iy[i] : 0u;
    const unsigned long long x = in && ix ? (unsigned)ix[i] : 0u, z = in && izl ? (unsigned)izl[i] : 0u;
    reinterpret_cast<unsigned long long *>(out + 3 * n_pad)[i] = (f & 63) << 34 | (y & 8191) << 21 | (x & 8191) << 8 | (z & 255);
}

cudaError_t launch_pack_snapshot(int64_t n, int64_t n_pad, const double *lon, const double *lat, const double *dep,
                                 const int32_t *face, const int32_t *iy, const int32_t *ix, const int32_t *izl, double *out,
                                 cudaStream_t s) {
    if (n_pad == 0) return cudaSuccess;
    pack_snapshot_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, s>>>(n, n_pad, lon, lat, dep, face, iy, ix, izl, out);
    return cudaGetLastError();
}

cudaError_t launch_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, n, a, b, out, bad);
    return cudaGetLastError();
}

}  // namespace sdk
