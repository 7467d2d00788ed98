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

cudaError_t launch_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, n, a, b, out, bad);
    return cudaGetLastError();
}

}  // namespace sdk
