// Decoder cross-attention for single-token rows (bf16), the dominant HBM stream of a decoder step:
// every (row, head) reads one contiguous 1500 x 64 K panel and one V panel (192 KB each) of the
// head-major cross-KV pool exactly once.
//
// Persistent, one CTA per SM, deliberately small: 8 consumer warps + 1 producer warp, ~107 KB of shared
// memory, so that the latency-bound projection kernels of ANOTHER decode lane (engine.cu) stay resident
// on the same SMs while this kernel keeps HBM busy.
//
//   producer (1 thread): cp.async.bulk (TMA, non-tensor) of 16 KB panel chunks into a 6-deep shared-memory
//       ring, L2 evict_first (the stream is read once per step), running ahead across (row, head) items so
//       the softmax barriers of an item never drain the memory pipeline.  K chunks do not depend on the
//       predecessor kernel and are requested before the programmatic-dependent-launch wait.
//   consumers: 8 lanes x 16 B per key row (conflict-free 512-byte warp reads), scores to shared memory,
//       block softmax, P*V with the same mapping, deterministic cross-warp sum.
#include <cuda.h>

#include "device_utils.cuh"
#include "gemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

constexpr int CA_WARPS = 8;                       // consumer warps
constexpr int CA_THREADS = (CA_WARPS + 1) * 32;   // + producer warp
constexpr int CA_CHUNK_KEYS = 128;
constexpr int CA_ROW_BYTES = 64 * 2;
constexpr int CA_CHUNK_BYTES = CA_CHUNK_KEYS * CA_ROW_BYTES;  // 16 KB
constexpr int CA_STAGES = 6;
constexpr int CA_MAX_KEYS = kWinRows;
constexpr int CA_SMEM = CA_STAGES * CA_CHUNK_BYTES + CA_MAX_KEYS * 4 + CA_WARPS * 64 * 4 + 32 * 4 + 2 * CA_STAGES * 8 + 128;

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CA_WARPS * 32) : "memory"); }
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    // bf16 -> fp32 is a 16-bit shift
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

__global__ void __launch_bounds__(CA_THREADS, 1)
dec_cross_attention_sm100_kernel(const RowDesc* __restrict__ rows, int n_items, int n_head, const bf16* __restrict__ q, int ldq,
                                 const bf16* __restrict__ kc, const bf16* __restrict__ vc, bf16* __restrict__ out, int ldo, size_t slot_stride,
                                 size_t head_stride, int n_keys) {
    extern __shared__ __align__(128) uint8_t ca_smem[];
    uint8_t* ring = ca_smem;
    float* sc = reinterpret_cast<float*>(ring + CA_STAGES * CA_CHUNK_BYTES);
    float* part = sc + CA_MAX_KEYS;         // [CA_WARPS][64]
    float* red = part + CA_WARPS * 64;      // [2][16]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 32);
    uint64_t* empty = full + CA_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < CA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_chunks = (n_keys + CA_CHUNK_KEYS - 1) / CA_CHUNK_KEYS;

    if (warp == CA_WARPS) {
        // ===== producer =====
        if (lane == 0) {
            const uint64_t policy = l2_evict_first_policy();
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int r = item / n_head, h = item - r * n_head;
                const size_t base = (size_t)rows[r].audio_slot * slot_stride + (size_t)h * head_stride;
                for (int pass = 0; pass < 2; ++pass) {
                    const bf16* src = (pass ? vc : kc) + base;
                    for (int c = 0; c < n_chunks; ++c) {
                        const int keys = min(CA_CHUNK_KEYS, n_keys - c * CA_CHUNK_KEYS);
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], (uint32_t)keys * CA_ROW_BYTES);
                        bulk_load(ring + stage * CA_CHUNK_BYTES, src + (size_t)c * CA_CHUNK_KEYS * 64, (uint32_t)keys * CA_ROW_BYTES, &full[stage], policy);
                        if (++stage == CA_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            // every load of this CTA has been requested: let the next kernel of the stream start its prologue
            pdl_launch_dependents();
        }
        return;
    }

    // ===== consumers =====
    pdl_wait();  // q comes from the predecessor
    const int tid = threadIdx.x;
    const int sub = lane & 7, ks = lane >> 3;   // 8 lanes x 16 B per key row, 4 key rows per warp instruction
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int r = item / n_head, h = item - r * n_head;
        float qv[8];
        {
            const uint4 qraw = *reinterpret_cast<const uint4*>(q + (size_t)r * ldq + h * 64 + sub * 8);
            unpack8(qraw, qv);
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] *= 0.125f;
        }
        // ---- scores
        float lmax = -INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full[stage], phase);
            const uint8_t* cb = ring + stage * CA_CHUNK_BYTES;
            uint4 raw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) raw[i] = lds128(cb + ((i * CA_WARPS + warp) * 4 + ks) * CA_ROW_BYTES + sub * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = c * CA_CHUNK_KEYS + (i * CA_WARPS + warp) * 4 + ks;
                float kv[8];
                unpack8(raw[i], kv);
                float a = 0.0f;
#pragma unroll
                for (int e = 0; e < 8; ++e) a = fmaf(qv[e], kv[e], a);
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                if (j < n_keys) {
                    if (sub == 0) sc[j] = a;
                    lmax = fmaxf(lmax, a);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == CA_STAGES) { stage = 0; phase ^= 1; }
        }
        lmax = warp_max(lmax);
        if (lane == 0) red[warp] = lmax;
        consumer_sync();
        float mx = red[0];
#pragma unroll
        for (int w = 1; w < CA_WARPS; ++w) mx = fmaxf(mx, red[w]);
        float lsum = 0.0f;
        for (int j = tid; j < n_keys; j += CA_WARPS * 32) {
            const float p = expf(sc[j] - mx);
            sc[j] = p;
            lsum += p;
        }
        lsum = warp_sum(lsum);
        if (lane == 0) red[16 + warp] = lsum;
        consumer_sync();
        float total = 0.0f;
#pragma unroll
        for (int w = 0; w < CA_WARPS; ++w) total += red[16 + w];   // fixed order
        const float inv = 1.0f / total;
        // ---- P * V
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full[stage], phase);
            const uint8_t* cb = ring + stage * CA_CHUNK_BYTES;
            uint4 raw[4];
            float p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int jl = (i * CA_WARPS + warp) * 4 + ks, j = c * CA_CHUNK_KEYS + jl;
                raw[i] = lds128(cb + jl * CA_ROW_BYTES + sub * 16);
                p[i] = j < n_keys ? sc[j] : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float vv[8];
                unpack8(raw[i], vv);
                if (p[i] != 0.0f) {   // rows past n_keys of the last chunk hold stale bytes
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p[i], vv[e], acc[e]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == CA_STAGES) { stage = 0; phase ^= 1; }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        }
        if (ks == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[warp * 64 + sub * 8 + e] = acc[e];
        }
        consumer_sync();
        if (tid < 64) {
            float o = 0.0f;
#pragma unroll
            for (int w = 0; w < CA_WARPS; ++w) o += part[w * 64 + tid];
            out[(size_t)r * ldo + h * 64 + tid] = __float2bfloat16_rn(o * inv);
        }
        // the next item's first write to sc/red/part happens after its own barriers; the barrier above
        // already separates this item's reads of sc (P*V) from them, and part is rewritten only after two more.
    }
}

}  // namespace

int cross_attention_sm100_smem_bytes() { return CA_SMEM; }

bool launch_dec_cross_attention_sm100(const RowDesc* rows, int n_rows, const bf16* q, int ldq, const bf16* kbase, const bf16* vbase, bf16* out, int ldo,
                                      int n_head, size_t slot_stride, size_t head_stride, int n_keys, int max_ctas, cudaStream_t s) {
    if (n_rows <= 0) return true;
    if (n_keys <= 0 || n_keys > CA_MAX_KEYS || (ldq % 8) != 0) { sm100_set_error("cross attention: unsupported shape"); return false; }
    static bool configured = false;
    static int sms = 0;
    if (!configured) {
        if (cudaFuncSetAttribute(dec_cross_attention_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CA_SMEM) != cudaSuccess) {
            sm100_set_error("cudaFuncSetAttribute(cross attention smem) failed");
            return false;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        configured = true;
    }
    const int n_items = n_rows * n_head;
    int grid = n_items < sms ? n_items : sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    launch_kernel(dec_cross_attention_sm100_kernel, dim3(grid), dim3(CA_THREADS), (size_t)CA_SMEM, s, true, rows, n_items, n_head, q, ldq, kbase, vbase, out,
                  ldo, slot_stride, head_stride, n_keys);
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { sm100_set_error(std::string("cross attention launch: ") + cudaGetErrorString(err)); return false; }
    return true;
}

}  // namespace nobs
