// Device self-test hooks: expose the branch-free math helpers of fastmath.cuh so that the tests can
// compare them with IEEE division and with 80-bit log / exp references (tests/test_gpu_math.py).
#include <cuda_runtime.h>

#include "../../include/seaduck_b200.h"
#include "fastmath.cuh"
#include "sdk_device.cuh"

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

// One analytic in-cell step with the helpers the advection kernel uses (axis_wall_times -> first event ->
// axis_increment; the argmin is the one of advect.cu: step()): a = [n][10] (x0[3], u[3], du[3], tf),
// out = [n][8] (tend, t_event, newx[3], newu[3]).  Lets the tests feed the vectors of the reference's own
// unit tests (tests/test_particle.py:75-209) to the device code.
__global__ void selftest_step_kernel(int64_t n, const double *a, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *in = a + i * 10;
    const double x0[3] = {in[0], in[1], in[2]}, u[3] = {in[3], in[4], in[5]}, du[3] = {in[6], in[7], in[8]};
    const double tf = in[9];
    const double sg = sgn(tf);
    double ts[7];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const AxisTimes t = axis_wall_times(u[k], du[k], x0[k], sg);
        ts[2 * k] = t.x, ts[2 * k + 1] = t.y;
    }
    ts[6] = tf;
    int tend = 0;
    double best = INFINITY;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
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
        if (tend == k) te = ts[k];
    double *o = out + i * 8;
    o[0] = (double)tend;
    o[1] = te;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double move = axis_increment(te, u[k], du[k]);
        o[2 + k] = move + x0[k];
        o[5 + k] = u[k] + du[k] * move;
    }
}

__global__ void selftest_cell_move_kernel(GridDev g, int64_t n, const int32_t *in, int32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = in[4 * i], iy = in[4 * i + 1], ix = in[4 * i + 2], came;
    const int st = cell_move(g, in[4 * i + 3], f, iy, ix, came);
    out[5 * i] = f, out[5 * i + 1] = iy, out[5 * i + 2] = ix, out[5 * i + 3] = came, out[5 * i + 4] = st;
}

cudaError_t launch_selftest_cell_move(const GridDev &g, int64_t n, const int32_t *in, int32_t *out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    selftest_cell_move_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, n, in, out);
    return cudaGetLastError();
}

cudaError_t launch_selftest_math(int op, int64_t n, const double *a, const double *b, double *out, int32_t *bad,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if (op == SDK_SELFTEST_STEP) {
        selftest_step_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, a, out);
        return cudaGetLastError();
    }
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(op, n, a, b, out, bad);
    return cudaGetLastError();
}

}  // namespace sdk
