// Device-side helpers shared by the CUDA-core kernels (kernels.cu) and the tcgen05 kernels.
#pragma once
#include "kernels.cuh"

namespace nobs {

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor is still running; pdl_wait() blocks until the predecessor
// grid has completed and its writes are visible, pdl_launch_dependents() lets the successor start
// launching.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Kernel timeline (debugging aid, NOBS_WHISPER_TRACE): block (0,0) thread 0 of an instrumented kernel appends
// {kernel id, tag, start ns, end ns} (globaltimer) to a device buffer whose first word is the entry counter.
// One pointer per translation unit (no relocatable device code), all set to the same buffer by trace_set_*().
static __device__ unsigned long long* g_trace = nullptr;
static __device__ unsigned int g_trace_cap = 0;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ long long trace_begin(int kid, const void* tag) {
    unsigned long long* buf = g_trace;
    if (!buf || (blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) != 0) return -1;
    const unsigned long long i = atomicAdd(buf, 1ull);
    if (i >= g_trace_cap) return -1;
    unsigned long long* e = buf + 4 + 4 * i;
    e[0] = (unsigned long long)kid;
    e[1] = (unsigned long long)tag;
    e[2] = globaltimer_ns();
    e[3] = 0;
    return (long long)i;
}
__device__ __forceinline__ void trace_end(long long i) {
    if (i >= 0) g_trace[4 + 4 * i + 3] = globaltimer_ns();
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_tanh(float x) {
    // ggml GELU: 0.5*x*(1 + tanh(sqrt(2/pi)*x*(1 + 0.044715*x^2)))
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}

__device__ __forceinline__ int float_key(float f) {  // order-preserving float -> int
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide reductions (blockDim.x multiple of 32, <= 1024); `red` holds >= 32 elements
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T ident, Op op, T* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = lane < nw ? red[lane] : ident;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;  // every thread holds the result
}
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpAddF { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMinI { __device__ int operator()(int a, int b) const { return a < b ? a : b; } };


// GELU for the bf16 path: one MUFU tanh.approx (error ~2^-11, below bf16 resolution)
__device__ __forceinline__ float gelu_tanh_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
    return 0.5f * x * (1.0f + t);
}

__device__ __forceinline__ float apply_epilogue(float acc, int m, int n, const Epilogue& e) {
    float v = acc;
    if (e.bias) v += __ldg(e.bias + n);
    if (e.act == 1) v = gelu_tanh(v);
    if (e.res) {
        const int rm = e.res_mod > 0 ? m % e.res_mod : m;
        v += e.res[(size_t)rm * e.res_ld + n];
    }
    if (e.win_rows > 0 && (m % e.win_rows) >= e.valid_rows) v = 0.0f;
    return v;
}


}  // namespace nobs
