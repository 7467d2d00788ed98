// Device-side pieces of the record exchange between ranks (exchange.cu, and the fused write-back of advect.cu):
// system-scope flag loads / stores, the multimem store, and the bounded wait on a row of per-rank counters.
#pragma once

#include <cstdint>

#include "../../include/seaduck_b200.h"

namespace sdk {

constexpr unsigned long long kSpinLimitNs = 4000000000ull;  // 4 s

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void multimem_st16(double2 *dst, double2 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__double2loint(v.x)),
                 "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y))
                 : "memory");
}

// wait until every counter of `row` (one per rank) has reached `want`; false on timeout
__device__ __forceinline__ bool wait_all(const uint64_t *row, int world, uint64_t want) {
    const unsigned long long t0 = now_ns();
    for (int r = 0; r < world; ++r) {
        while (ld_acquire_sys(row + r) < want) {
            if (now_ns() - t0 > kSpinLimitNs) return false;
            __nanosleep(200);
        }
    }
    return true;
}

// store one 32-byte record (two 16-byte halves) at position `at` of every rank's receive buffer
__device__ __forceinline__ void ship_record(const sdk_exchange &x, int64_t at, double2 ra, double2 rb) {
    if (x.recv_multicast) {
        // one store, replicated into every rank's buffer by the switch (NVLS)
        // (multimem.st has no 64-bit vector form: the 16 bytes go as four 32-bit words; a store moves bits)
        double2 *dst = reinterpret_cast<double2 *>(x.recv_multicast + at);
        multimem_st16(dst, ra);
        multimem_st16(dst + 1, rb);
    } else {
#pragma unroll 1
        for (int k = 0; k < x.world; ++k) {
            const int r = (x.rank + k) % x.world;  // own copy first, then round the ring
            double2 *dst = reinterpret_cast<double2 *>(x.recv[r] + at);
            dst[0] = ra;
            dst[1] = rb;
        }
    }
}

// first record position of this rank's block in the slot of `step`
__device__ __forceinline__ int64_t ship_base(const sdk_exchange &x, unsigned long long step) {
    return ((int64_t)(step % (unsigned long long)x.n_slots) * x.world + x.rank) * x.n_pad;
}

// thread 0 of a CTA: the slot of `step` is free once every rank has consumed the step that last used it
__device__ __forceinline__ bool ship_slot_free(const sdk_exchange &x, unsigned long long step) {
    uint64_t *mine = x.signal[x.rank];
    const unsigned long long prev = step > (unsigned long long)x.n_slots ? step - x.n_slots : 0;
    const bool ok = prev == 0 || wait_all(mine + SDK_SIGNAL_CONSUMED, x.world, prev);
    if (!ok) atomicExch((unsigned long long *)(mine + SDK_SIGNAL_ERROR), 1ull);
    return ok;
}

// Called by every thread of a CTA after its last store; the CTA that finishes last posts `step` in every rank's
// arrived[rank].  One system-scope fence per CTA, not per thread: the barrier orders the CTA's stores before thread 0,
// whose fence (cumulative) then orders them before the ticket and the flags.  Contains __syncthreads().
__device__ __forceinline__ void ship_finish(const sdk_exchange &x, unsigned long long step) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(x.ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
#pragma unroll 1
            for (int r = 0; r < x.world; ++r) st_release_sys(x.signal[r] + SDK_SIGNAL_ARRIVED + x.rank, step);
            *x.ticket = 0;
        }
    }
}

}  // namespace sdk
