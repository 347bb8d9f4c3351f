// Decoder projection chain for step batches (R <= 128 token rows, bf16), sm_100a only.
//
// Between two attention kernels a decoder layer is a CHAIN of small dependent projections
// (out-proj -> +residual -> LayerNorm -> FC1 -> GELU -> FC2 -> +residual -> LayerNorm -> next QKV ...).
// As separate launches every link costs a launch boundary (2-7 us when another decode lane shares the
// SMs) on top of ~5 us of work, and that chain, not HBM, bounded the decoder in round 1
// (profiles/r1_timeline_one_layer_*.txt).  This kernel runs a whole chain as ONE persistent grid:
//
//   step s:  swap-AB split-K tcgen05 GEMM (identical tiling / split order / numerics to
//            gemm_skinny_sm100_kernel: weight tile = 128-row UMMA A operand via TMA, token rows = B operand)
//            -> fp32 partial sums in the lane's workspace
//            -> device-wide barrier
//            -> fused epilogue (same arithmetic as skinny_reduce_kernel: fixed-order split sum, bias, GELU,
//               residual into the fp32 stream, next LayerNorm, KV-cache scatter), rows strided over the CTAs
//            -> device-wide barrier -> step s + 1
//
// The weight tiles of step s + 1 are requested while step s is still being reduced: they do not depend on
// anything computed here, so HBM latency of the weight stream disappears from the chain.  The last step may
// leave its partial sums to the attention kernel that follows (dec_self_attention_kernel /
// dec_cross_attention_tc_kernel finish them), exactly like the multi-launch path.
//
// Co-residency: the device-wide barrier needs every CTA of the grid resident at the same time.  The grid never
// exceeds the SM count and the launcher refuses configurations in which all decode lanes' chain CTAs could not
// sit on one SM together (see chain_fits): every other kernel of the decoder terminates without waiting on
// anything, so a chain CTA is at worst delayed, never starved.  All spins trap after ~2 s instead of hanging.
#include "gemm_sm100.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "device_utils.cuh"
#include "grid_sync.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

constexpr int BM = 128;      // UMMA M (weight rows per tile)
constexpr int BK = 64;       // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int CH_THREADS = 256;
constexpr int CH_MAXG = 5;     // float4 column groups per thread in the epilogue: N <= 5120
constexpr int CH_PLANES = 8;   // split-K planes summed per trip

__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct ChainMaps {
    CUtensorMap w[kChainMaxSteps];
    CUtensorMap x[kChainMaxSteps];
};

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) { return gs_ld_acquire(p); }

// barrier flavour (grid_sync.cuh); chosen from the micro-benchmark (whisper_b200_debug_grid_sync)
#ifndef NOBS_CHAIN_SYNC_VARIANT
#define NOBS_CHAIN_SYNC_VARIANT 2
#endif
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int G, unsigned int& gen) { grid_sync_v<NOBS_CHAIN_SYNC_VARIANT>(bar, G, gen); }

template <typename T> __device__ __forceinline__ void ch_store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void ch_store4<bf16>(bf16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// Epilogue of one token row: the arithmetic of skinny_reduce_kernel (kernels.cu), statement for statement, so that the
// chain and the multi-launch path give bit-identical results.
__device__ __forceinline__ void chain_reduce_row(const SkinnyEpilogue& e, int r, float* red) {
    using T = bf16;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)e.R * e.N;
    const float* p = e.partial + (size_t)r * e.N;
    const int n4 = e.N >> 2;
    float4 vals[CH_MAXG];
    RowDesc rd{};
    if (e.rows) rd = e.rows[r];
    float lsum = 0.0f;
#pragma unroll
    for (int g = 0; g < CH_MAXG; ++g) {
        const int c4 = tid + g * CH_THREADS;
        vals[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 >= n4) continue;
        const int n = c4 * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < e.splits; s0 += CH_PLANES) {   // fixed order: deterministic
            float4 t[CH_PLANES];
#pragma unroll
            for (int u = 0; u < CH_PLANES; ++u)
                t[u] = (s0 + u < e.splits) ? __ldcg(reinterpret_cast<const float4*>(p + (size_t)(s0 + u) * plane + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < CH_PLANES; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        if (e.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (e.act == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_tanh_fast(v[i]);
        }
        if (e.x) {
            float4* xp = reinterpret_cast<float4*>(e.x + (size_t)r * e.N + n);
            const float4 xv = __ldcg(xp);
            v[0] += xv.x; v[1] += xv.y; v[2] += xv.z; v[3] += xv.w;
            *xp = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (e.out) ch_store4<T>(static_cast<T*>(e.out) + (size_t)r * e.out_ld + n, v);
        if (e.rows && n >= e.d) {
            const int c = n - e.d, which = c >= e.d, i2 = which ? c - e.d : c;
            T* dst = static_cast<T*>(which ? e.vpanel : e.kpanel) + (size_t)rd.kv_slot * e.slot_stride + ((size_t)(i2 >> 6) * e.n_pos_cap + rd.pos) * 64 + (i2 & 63);
            ch_store4<T>(dst, v);
        }
        lsum += (v[0] + v[1]) + (v[2] + v[3]);
        vals[g] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (e.ln_g) {
        const float mean = block_reduce(lsum, 0.0f, OpAddF(), red) / e.N;
        float lvar = 0.0f;
#pragma unroll
        for (int g = 0; g < CH_MAXG; ++g) {
            if (tid + g * CH_THREADS < n4) {
                const float a = vals[g].x - mean, b = vals[g].y - mean, c = vals[g].z - mean, dd = vals[g].w - mean;
                lvar += (a * a + b * b) + (c * c + dd * dd);
            }
        }
        const float var = block_reduce(lvar, 0.0f, OpAddF(), red) / e.N;
        const float inv = rsqrtf(var + 1e-5f);
        T* y = static_cast<T*>(e.y) + (size_t)r * e.N;
#pragma unroll
        for (int g = 0; g < CH_MAXG; ++g) {
            const int c4 = tid + g * CH_THREADS;
            if (c4 < n4) {
                const int n = c4 * 4;
                const float4 gg = __ldg(reinterpret_cast<const float4*>(e.ln_g + n)), bb = __ldg(reinterpret_cast<const float4*>(e.ln_b + n));
                const float o[4] = {(vals[g].x - mean) * inv * gg.x + bb.x, (vals[g].y - mean) * inv * gg.y + bb.y,
                                    (vals[g].z - mean) * inv * gg.z + bb.z, (vals[g].w - mean) * inv * gg.w + bb.w};
                ch_store4<T>(y + n, o);
            }
        }
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(CH_THREADS, 3)
dec_chain_sm100_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainDesc cd) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t chain_smem_raw[];
    __shared__ float red[32];
    const uint32_t raw = smem_u32(chain_smem_raw);
    uint8_t* smem = chain_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const unsigned int G = gridDim.x;
    const int R = cd.R;
    const long long tr = trace_begin(8, cd.partial);

    // work item of this CTA in step s: weight tile m_blk, k-blocks [kb0, kb1)
    auto item = [&](int s, int& m_blk, int& kb0, int& kb1) -> bool {
        const ChainStep& st = cd.step[s];
        if (cta >= st.m_tiles * st.splits) return false;
        m_blk = cta % st.m_tiles;
        const int split = cta / st.m_tiles;
        const int num_k = (st.K + BK - 1) / BK;
        kb0 = split * st.kb_per_split;
        kb1 = min(num_k, kb0 + st.kb_per_split);
        return kb1 > kb0;
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < cd.n_steps; ++s) { tma_prefetch_desc(&maps.w[s]); tma_prefetch_desc(&maps.x[s]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint64_t w_policy = l2_evict_normal_policy();

    // ring positions: the producer's (warp 0) and the MMA issuer's (warp 1); every lane tracks them identically
    int stage = 0; uint32_t phase = 0;
    uint32_t acc_phase = 0;
    int pre = 0;   // weight tiles of the CURRENT step already requested (their activations are still to come)

    // The weights do not depend on the predecessor kernel: request the first tiles before the PDL wait.
    if (warp == 0) {
        int m_blk, kb0, kb1;
        if (item(0, m_blk, kb0, kb1)) {
            pre = min(STAGES, kb1 - kb0);
            if (lane == 0) {
                for (int i = 0; i < pre; ++i) {
                    mbar_expect_tx(&full[i], STAGE_BYTES);
                    tma_load_2d_hint(sA + i * A_BYTES, &maps.w[0], &full[i], (kb0 + i) * BK, m_blk * BM, w_policy);
                }
            }
        }
    }
    pdl_wait();
    trace_end(trace_begin(108, cd.partial));
    unsigned int gen = ld_acquire_u32(&cd.bar[1]);

    for (int s = 0; s < cd.n_steps; ++s) {
        const ChainStep& st = cd.step[s];
        // the successor (an attention kernel) may become resident while the last projection runs, not earlier:
        // a parked grid of 57-74 KB CTAs would hold shared memory the other decode lanes need
        if (s == cd.n_steps - 1) pdl_launch_dependents();
        int m_blk = 0, kb0 = 0, kb1 = 0;
        const bool has = item(s, m_blk, kb0, kb1);
        if (warp == 0) {
            // ===== TMA producer =====
            if (has) {
                if (lane == 0) {
                    fence_proxy_async_global();   // activations were written through the generic proxy by other CTAs
                    for (int i = 0; i < pre; ++i) {
                        const int j = stage + i, sj = j >= STAGES ? j - STAGES : j;
                        tma_load_2d(sB + sj * B_BYTES, &maps.x[s], &full[sj], (kb0 + i) * BK, 0);
                    }
                }
                __syncwarp();
                for (int i = 0; i < pre; ++i) { if (++stage == STAGES) { stage = 0; phase ^= 1; } }
                for (int kb = kb0 + pre; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        mbar_expect_tx(&full[stage], STAGE_BYTES);
                        tma_load_2d_hint(sA + stage * A_BYTES, &maps.w[s], &full[stage], kb * BK, m_blk * BM, w_policy);
                        tma_load_2d(sB + stage * B_BYTES, &maps.x[s], &full[stage], kb * BK, 0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // weight tiles of the next step: requested now, they land while this step is reduced
            pre = 0;
            if (s + 1 < cd.n_steps) {
                int m2, k0, k1;
                if (item(s + 1, m2, k0, k1)) {
                    pre = min(STAGES, k1 - k0);
                    for (int i = 0; i < pre; ++i) {
                        const int j = stage + i, sj = j >= STAGES ? j - STAGES : j;
                        const uint32_t pj = j >= STAGES ? phase ^ 1 : phase;
                        mbar_wait(&empty[sj], pj ^ 1);
                        if (lane == 0) {
                            mbar_expect_tx(&full[sj], STAGE_BYTES);
                            tma_load_2d_hint(sA + sj * A_BYTES, &maps.w[s + 1], &full[sj], (k0 + i) * BK, m2 * BM, w_policy);
                        }
                        __syncwarp();
                    }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            if (has) {
                constexpr uint32_t idesc = make_idesc(BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(sA + stage * A_BYTES), b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(tmem_base, make_smem_desc_kmajor(a_addr + k * UMMA_K * 2), make_smem_desc_kmajor(b_addr + k * UMMA_K * 2), idesc,
                                      (uint32_t)((kb > kb0) | (k != 0)));
                        umma_commit(&empty[stage]);
                        if (kb == kb1 - 1) umma_commit(tfull);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp >= 4) {
            // ===== partial sums: TMEM -> partial[split][r][n] (lane = output feature, TMEM column = token row) =====
            if (has) {
                const int q = warp & 3;
                const int n = m_blk * BM + q * 32 + lane;
                const int split = cta / st.m_tiles;
                mbar_wait(tfull, acc_phase);
                tc_fence_after();
                float* dst = cd.partial + (size_t)split * R * st.N + n;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (n < st.N) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < R) dst[(size_t)(c0 + j) * st.N] = __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
            }
        }
        if (has) acc_phase ^= 1;
        if (st.reduce) {
            if (g_trace) { __syncthreads(); trace_end(trace_begin(130 + s, cd.partial)); }   // this CTA's share of the GEMM is done
            grid_sync(cd.bar, G, gen);           // every partial sum of this step is in the workspace
            tc_fence_after();
            trace_end(trace_begin(110 + s, cd.partial));
            for (int r = cta; r < R; r += (int)G) chain_reduce_row(st.e, r, red);
            fence_proxy_async_global();          // the next step reads y / h through TMA (async proxy)
            if (g_trace) { __syncthreads(); trace_end(trace_begin(140 + s, cd.partial)); }   // this CTA's rows are reduced
            grid_sync(cd.bar, G, gen);
            trace_end(trace_begin(120 + s, cd.partial));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    trace_end(tr);
}

int num_sms_chain() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr int chain_smem_bytes(int bn, int stages) { return stages * (A_BYTES + bn * BK * 2) + 1024 + 256; }

template <int BN, int STAGES>
bool launch_chain_cfg(const ChainMaps& maps, const ChainDesc& cd, int grid, cudaStream_t s) {
    constexpr int SMEM = chain_smem_bytes(BN, STAGES);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dec_chain_sm100_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            sm100_set_error("cudaFuncSetAttribute(decoder chain smem) failed");
            return false;
        }
        configured = true;
    }
    launch_kernel(dec_chain_sm100_kernel<BN, STAGES>, dim3(grid), dim3(CH_THREADS), (size_t)SMEM, s, true, maps, cd);
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { sm100_set_error(std::string("decoder chain launch: ") + cudaGetErrorString(err)); return false; }
    return true;
}

}  // namespace

void trace_set_chain(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}

int chain_stages_for_lanes(int n_lanes) { return n_lanes <= 2 ? 3 : 2; }

// All decode lanes may sit in a chain kernel at the same time: their CTAs must fit on one SM together (shared memory,
// TMEM columns, threads; registers are capped by __launch_bounds__(256, 3) for up to 3 lanes), otherwise two grids could
// wait for each other's SM slots forever.  Worst-case tile (128 rows) is assumed.
bool chain_fits(int n_lanes, int stages) {
    if (n_lanes < 1 || n_lanes > 3) return false;
    if ((size_t)n_lanes * chain_smem_bytes(128, stages) > (size_t)220 * 1024) return false;
    if (n_lanes * 128 > 512) return false;
    return true;
}

bool launch_dec_chain_sm100(ChainDesc& cd, const bf16* const* W, const bf16* const* X, const int* ldx, int stages, cudaStream_t s) {
    if (cd.n_steps <= 0 || cd.n_steps > kChainMaxSteps || cd.R <= 0 || cd.R > 128 || !cd.partial || !cd.bar) { sm100_set_error("decoder chain: bad description"); return false; }
    const int sms = num_sms_chain();
    const int bn = cd.R <= 32 ? 32 : cd.R <= 64 ? 64 : 128;
    ChainMaps maps;
    int grid = cd.R;
    for (int i = 0; i < cd.n_steps; ++i) {
        ChainStep& st = cd.step[i];
        if (st.N <= 0 || st.K <= 0 || (st.N & 3) || st.N > 4 * CH_MAXG * CH_THREADS) { sm100_set_error("decoder chain: bad step shape"); return false; }
        if (!st.reduce && i != cd.n_steps - 1) { sm100_set_error("decoder chain: only the last step may leave its partial sums"); return false; }
        const int num_k = (st.K + BK - 1) / BK;
        st.m_tiles = (st.N + BM - 1) / BM;
        st.splits = skinny_gemm_splits(st.N, st.K);
        st.kb_per_split = (num_k + st.splits - 1) / st.splits;
        if (st.m_tiles * st.splits > sms) { sm100_set_error("decoder chain: step does not fit one wave"); return false; }
        grid = std::max(grid, st.m_tiles * st.splits);
        st.e.partial = cd.partial; st.e.splits = st.splits; st.e.R = cd.R; st.e.N = st.N;
        if (!get_tmap_bf16_2d(&maps.w[i], W[i], (uint64_t)st.K, (uint64_t)st.N, (uint64_t)st.K, BK, BM)) return false;
        if (!get_tmap_bf16_2d(&maps.x[i], X[i], (uint64_t)st.K, (uint64_t)cd.R, (uint64_t)ldx[i], BK, (uint32_t)bn)) return false;
    }
    for (int i = cd.n_steps; i < kChainMaxSteps; ++i) { maps.w[i] = maps.w[0]; maps.x[i] = maps.x[0]; }
    grid = std::min(grid, sms);
#define CHAIN_CASE(B, S) if (bn == B && stages == S) return launch_chain_cfg<B, S>(maps, cd, grid, s);
    CHAIN_CASE(32, 2) CHAIN_CASE(64, 2) CHAIN_CASE(128, 2) CHAIN_CASE(32, 3) CHAIN_CASE(64, 3) CHAIN_CASE(128, 3)
#undef CHAIN_CASE
    sm100_set_error("decoder chain: unsupported stage count");
    return false;
}

}  // namespace nobs
