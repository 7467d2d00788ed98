// Arguments of the budget post-processing kernels (shared by budget.cu and capi.cu).
#pragma once

#include <cuda_runtime.h>

#include "sdk_device.cuh"

namespace sdk {

struct BudgetArgs {
    GridDev g;
    const sdk_log_record *rec;
    const int64_t *offset;
    int64_t n_particles, n_rec;
    int16_t *ind1, *ind2;  // [rows][n_rec]
    double *frac, *tres, *work;
    int rows;
};

cudaError_t launch_budget(const BudgetArgs &a, cudaStream_t s);

struct BudgetTermArgs {
    const sdk_log_record *rec;
    const int64_t *offset;
    int64_t n_particles, n_rec;
    const int16_t *ind1, *ind2;
    const double *frac, *tres;
    sdk_budget_fields f;
    double *conc, *contr;
    unsigned char *is_last;
};

cudaError_t launch_budget_terms(const BudgetTermArgs &a, cudaStream_t s);

struct WallValueArgs {
    GridDev g;
    const sdk_log_record *rec;
    int64_t n_rec;
    sdk_wall_fields f;
    double *c_list, *u_list, *lag_conv, *eul_conv;
};

cudaError_t launch_wall_values(const WallValueArgs &a, cudaStream_t s);

}  // namespace sdk
