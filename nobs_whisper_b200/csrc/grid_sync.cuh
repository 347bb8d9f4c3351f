// Device-wide barrier over the G co-resident CTAs of one grid (decode_chain_sm100.cu), and its micro-benchmark variants.
// bar[0] = arrival counter, bar[1] = generation.  Self-resetting: the last arriver zeroes the counter before it publishes the
// new generation, so consecutive kernels on one stream can share the two words.  Every spin traps after ~2 s: a protocol error
// must surface as a launch failure, never as a hung GPU.
#pragma once
#include <cuda_runtime.h>

namespace nobs {

__device__ __forceinline__ unsigned int gs_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// V = 0: fence + relaxed atomic + polling with nanosleep + fence (cooperative-groups style)
// V = 1: the same without nanosleep
// V = 2: one acq_rel atomic to arrive, release-increment of the generation, acquire polling; no stand-alone fences
template <int V>
__device__ __forceinline__ void grid_sync_v(unsigned int* bar, unsigned int G, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int old;
        if (V == 2) {
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        } else {
            __threadfence();
            old = atomicAdd(&bar[0], 1u);
        }
        if (old == G - 1) {
            if (V == 2) {
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
            } else {
                bar[0] = 0;
                __threadfence();
                atomicAdd(&bar[1], 1u);
            }
        } else {
            const long long t0 = clock64();
            while (gs_ld_acquire(&bar[1]) == gen) {
                if (V == 0) __nanosleep(32);
                if (clock64() - t0 > 4000000000LL) __trap();
            }
        }
        if (V != 2) __threadfence();
    }
    gen += 1;
    __syncthreads();
}

}  // namespace nobs
