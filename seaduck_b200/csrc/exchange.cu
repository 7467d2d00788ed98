// Output records of a stop and their exchange between the GPUs of one node.
//
// What Particle.to_list_of_time hands back per stop (reference seaduck/lagrangian.py:804, the deepcopy of
// :725) is, per particle, a position and a cell: 32 bytes (sdk_snapshot_record).  This file packs them
//   * into a local buffer (optionally scattered back to the caller's particle order), and
//   * straight into the receive buffers of every rank of the node -- peer mappings of symmetric memory,
//     written over NVLink by the packing kernel itself (plain stores to each peer, or one multimem.st that
//     the NVSwitch replicates), followed by a release flag per peer.  No collective library call, no
//     copy-engine staging, no second pass over the data.
//
// Flow control between ranks uses two arrays of 64-bit counters in every rank's signal words:
//   arrived[r]  = last step whose records rank r has finished writing HERE
//   consumed[r] = last step whose records rank r has finished reading from ITS buffers
// A slot (step % n_slots) is overwritten only when every rank has consumed what it held.  Spins are bounded
// (kSpinLimitNs); a timeout raises the error word instead of hanging the GPU.
#include <cuda_runtime.h>

#include <cstdint>

#include "exchange_device.cuh"

namespace sdk {

__device__ __forceinline__ uint64_t cell_word(const sdk_particles &p, int64_t i) {
    const uint64_t f = p.face ? (uint64_t)(uint32_t)p.face[i] : 0;
    const uint64_t y = (uint64_t)(uint32_t)p.iy[i], x = (uint64_t)(uint32_t)p.ix[i];
    const uint64_t z = p.izl ? (uint64_t)(uint32_t)p.izl[i] : 0;
    return SDK_CELL_PACK(f, y, x, z);
}

__device__ __forceinline__ void load_record(const sdk_particles &p, int64_t i, int64_t n, double2 &a, double2 &b) {
    if (i < n) {
        a = make_double2(p.lon[i], p.lat[i]);
        b = make_double2(p.dep ? p.dep[i] : 0.0, __longlong_as_double((long long)cell_word(p, i)));
    } else {
        a = make_double2(0.0, 0.0);
        b = a;
    }
}

// records[out_index ? out_index[i] : i] = record of particle i
__global__ void __launch_bounds__(256) pack_records_kernel(sdk_particles p, int64_t n, const int32_t *out_index,
                                                           sdk_snapshot_record *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 a, b;
    load_record(p, i, n, a, b);
    double2 *dst = reinterpret_cast<double2 *>(out + (out_index ? (int64_t)out_index[i] : i));
    dst[0] = a;
    dst[1] = b;
}

struct ShipArgs {
    sdk_exchange x;
    sdk_particles p;
    int64_t n;
    uint64_t step;
    const int32_t *out_index;  // optional: record of particle i goes to position out_index[i] of the block
};

// Pack the records of this rank and store them into block [slot][rank] of every rank's receive buffer.
__global__ void __launch_bounds__(256) pack_ship_kernel(const __grid_constant__ ShipArgs a) {
    const sdk_exchange &x = a.x;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = ship_slot_free(x, a.step);
    __syncthreads();
    const int64_t base = ship_base(x, a.step);
    if (ok) {
        // (the padding behind the last particle is written too -- zeros -- unless the records are scattered)
        const int64_t count = a.out_index ? a.n : x.n_pad;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
            double2 ra, rb;
            load_record(a.p, i, a.n, ra, rb);
            ship_record(x, base + (a.out_index ? (int64_t)a.out_index[i] : i), ra, rb);
        }
    }
    ship_finish(x, a.step);
}

// All ranks' records of `step` are in this rank's buffer; optionally tell everyone that this rank is done
// with them (release).  One warp.
__global__ void exchange_wait_kernel(sdk_exchange x, uint64_t step, int wait, int release) {
    uint64_t *mine = x.signal[x.rank];
    if (threadIdx.x == 0 && wait) {
        if (!wait_all(mine + SDK_SIGNAL_ARRIVED, x.world, step)) atomicExch((unsigned long long *)(mine + SDK_SIGNAL_ERROR), 1ull);
    }
    __syncwarp();
    if (release && threadIdx.x < (unsigned)x.world) st_release_sys(x.signal[threadIdx.x] + SDK_SIGNAL_CONSUMED + x.rank, step);
}

cudaError_t launch_pack_records(const sdk_particles &p, int64_t n, const int32_t *out_index, sdk_snapshot_record *out,
                                cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    pack_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, out_index, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_ship(const sdk_exchange &x, const sdk_particles &p, int64_t n, const int32_t *out_index,
                             uint64_t step, int sm_count, cudaStream_t s) {
    ShipArgs a;
    a.x = x;
    a.p = p;
    a.n = n;
    a.step = step;
    a.out_index = out_index;
    // a small grid: the stores go out over NVLink, a few CTAs per SM on a part of the chip keep it busy and leave
    // the rest to the advection kernel of the next leg
    int64_t blocks = (x.n_pad + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pack_ship_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_exchange_wait(const sdk_exchange &x, uint64_t step, int wait, int release, cudaStream_t s) {
    exchange_wait_kernel<<<1, 32, 0, s>>>(x, step, wait, release);
    return cudaGetLastError();
}

}  // namespace sdk
