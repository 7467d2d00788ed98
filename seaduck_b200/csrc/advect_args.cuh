// Arguments of the advection kernels and their host launchers (shared by advect.cu and capi.cu).
#pragma once

#include <cuda_runtime.h>

#include "sdk_device.cuh"

namespace sdk {

struct AdvectArgs {
    GridDev g;
    VelDev v;
    sdk_particles p;
    sdk_log log;
    double t_stop, tol, trim_tol;
    int max_iteration, kick_back, has_dep, has_log, commit, refill_threshold, keep_status;
    unsigned long long *queue;  // work queue head
    unsigned long long *stats;  // optional [4]: live at start, crossings, hit max_iteration, flagged
    // fused output exchange (sdk_advect_params.ship): the mappings, the step and the optional position table
    sdk_exchange x;
    unsigned long long ship_step;
    const int32_t *ship_index;
    int ship;
};

int advect_max_threads();
int advect_default_threads();
int advect_profile(const AdvectArgs &args, int threads);
int advect_default_blocks_per_sm(int profile);
cudaError_t launch_advect(const AdvectArgs &args, int blocks, int threads, cudaStream_t stream);
cudaError_t launch_get_u_du(const AdvectArgs &args, cudaStream_t stream);

}  // namespace sdk
