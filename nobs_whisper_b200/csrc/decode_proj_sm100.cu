// Decoder-step projection with everything around it fused (sm_100a): ONE launch per projection of a step batch
// (R <= 128 token rows) instead of "split-K GEMM + reduce/epilogue kernel (+ LayerNorm)".
//
//   v[r][n] = sum_k X[r][k] * W[n][k] + bias[n]   (optionally GELU)
//   x[r][n] += v   and/or   out[r][n] = bf16(v)   and/or   K/V-cache append   and/or   y = bf16(LayerNorm(x))
//
// * swap-AB on tcgen05 as in gemm_skinny_sm100_kernel: the 128-row weight tile is the UMMA A operand, the token rows the
//   small B operand, fp32 accumulators D[128 features][BN rows] in TMEM.
// * split-K over a thread-block CLUSTER (1, 2, 4 or 8 CTAs): every CTA of a cluster owns the same 128 output features and a
//   slice of K.  The partial sums never go to global memory: each CTA parks its D tile in its own shared memory, and after a
//   cluster barrier CTA s sums features [s*128/S, (s+1)*128/S) of all S tiles through distributed shared memory in the fixed
//   order 0..S-1 (deterministic), then applies bias / GELU / residual add into the fp32 stream / bf16 store / KV-cache append.
// * LayerNorm of the updated residual stream rides on the same launch: every cluster emits, per row, (mean, M2) of its 128
//   features (rank partials combined with Chan's parallel-variance update), then takes a ticket; the cluster that draws the
//   last ticket knows the whole of x and every tile's statistics are in global memory, combines the tiles' statistics per row
//   and writes y = LayerNorm(x) * g + b as bf16 for the next projection's TMA loads.  Which cluster is last varies, what it
//   computes does not: results are bit-reproducible.
//
// One decoder layer of a step batch is 6 of these + self-attention + cross-attention = 8 launches (round 1: 12), and the
// dependent chain between two attention kernels loses its global-memory round trips (partials out / in, LayerNorm launch).
//
// 192 threads at <= 80 registers, <= 56 KB of shared memory at 64 rows: one such CTA fits beside the two persistent
// cross-attention CTAs (2 x 192 threads x 120 registers, 2 x 57 KB) another decode lane keeps on every SM — a projection that
// cannot become resident there would wait a whole attention launch.
// Warp roles: 0 = TMA producer (weights are requested BEFORE the programmatic-dependency wait), 1 = TMEM allocator + MMA
// issuer, 2..5 = TMEM -> shared memory; all 6 warps run the cluster reduction, the epilogue and the LayerNorm tail.
#include "gemm_sm100.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "device_utils.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int PJ_THREADS = 192;
constexpr int PJ_MAX_STAGES = 4;
constexpr int PJ_STAGE_LD = BM + 4;  // floats per row of the parked D tile: 16-byte aligned rows, conflict-free float4 reads
constexpr int PJ_PRE = 3;            // epilogue items (float4 of bias + residual) a thread requests before the cluster barrier

__host__ __device__ constexpr uint32_t pj_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// shared-memory layout (the ring depth is a launch parameter: deep K slices get 4 stages)
//   [0, region0)      operand ring: stages x (16 KB weight tile + BN x 128 B activation tile);  after the last MMA: the parked D tile
//   [region0, +stat)  rank 0: (mean, M2) received from the cluster ranks
//   [.., +bars)       mbarriers, TMEM slot, "this cluster drew the last ticket" flag
template <int BN> struct PjCfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int PARK_BYTES = BN * PJ_STAGE_LD * 4;
    static constexpr int STAT_BYTES = 8 * BN * 8;
    static constexpr int BAR_BYTES = 8 * (2 * PJ_MAX_STAGES + 1) + 16;
    static __host__ __device__ constexpr int region0(int stages) {
        return stages * (A_BYTES + B_BYTES) > PARK_BYTES ? stages * (A_BYTES + B_BYTES) : PARK_BYTES;
    }
    static __host__ __device__ constexpr int smem_bytes(int stages) { return region0(stages) + STAT_BYTES + BAR_BYTES; }
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_dsmem_f2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int BN>
__global__ void __maxnreg__(80)
dec_proj_cluster_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const ProjDesc a, const int stages) {
    using cfg = PjCfg<BN>;
    constexpr int B_BYTES = cfg::B_BYTES;
    constexpr uint32_t STAGE_TX = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) uint8_t smem[];              // swizzle atoms need 1024-byte alignment; same offsets in every CTA of the cluster
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const int region0 = cfg::region0(stages);
    uint8_t* sA = smem;
    uint8_t* sB = smem + stages * A_BYTES;
    float* park = reinterpret_cast<float*>(smem);                  // [BN][PJ_STAGE_LD], valid after the last MMA
    float2* statbuf = reinterpret_cast<float2*>(smem + region0);   // [8][BN]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + region0 + cfg::STAT_BYTES);
    uint64_t* empty = full + PJ_MAX_STAGES;
    uint64_t* tfull = empty + PJ_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    uint32_t* last_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int S = (int)cluster_size(), split = (int)cluster_rank();
    const int m_blk = blockIdx.x / S;
    const int R = a.R, N = a.N, K = a.K;
    const long long tr = trace_begin(9, a.W);
    const int num_k = K / BK;
    const int kb0 = (int)((long long)split * num_k / S), kb1 = (int)((long long)(split + 1) * num_k / S);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // epilogue geometry: this CTA finishes features [split*F, (split+1)*F) of the tile for all rows; thread -> (row, float4 group)
    const int F = BM / S, G = F >> 2;                 // G float4 groups per row (S = 8: 4 ... S = 1: 32), consecutive lanes of a warp
    float4 pre_b[PJ_PRE], pre_x[PJ_PRE];

    if (warp == 0) {
        // weights do not depend on the predecessor: fill the ring (and ask L2 for the rest of this CTA's tiles) before the wait
        const int pre = min(stages, nkb);
        const uint64_t w_policy = l2_evict_normal_policy();
        if (lane == 0) {
            for (int i = 0; i < pre; ++i) {
                mbar_expect_tx(&full[i], STAGE_TX);
                tma_load_2d_hint(sA + i * A_BYTES, &tmap_w, &full[i], (kb0 + i) * BK, m_blk * BM, w_policy);
            }
            for (int i = pre; i < nkb; ++i) tma_prefetch_l2_2d(&tmap_w, (kb0 + i) * BK, m_blk * BM);
        }
        pdl_wait();
        pdl_launch_dependents();   // successor prologue overlaps this kernel's work; never more than one kernel parked ahead
        trace_end(trace_begin(109, a.W));
        if (lane == 0) {
            for (int i = 0; i < pre; ++i) tma_load_2d(sB + i * B_BYTES, &tmap_x, &full[i], (kb0 + i) * BK, 0);
        }
        __syncwarp();
        int stage = pre % stages; uint32_t phase = pre == stages ? 1 : 0;
        for (int kb = kb0 + pre; kb < kb1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&full[stage], STAGE_TX);
                tma_load_2d_hint(sA + stage * A_BYTES, &tmap_w, &full[stage], kb * BK, m_blk * BM, w_policy);
                tma_load_2d(sB + stage * B_BYTES, &tmap_x, &full[stage], kb * BK, 0);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else {
        pdl_wait();
    }
    // bias and the residual values of this thread's first epilogue items: requested now, consumed after the cluster barrier
#pragma unroll
    for (int it = 0; it < PJ_PRE; ++it) {
        const int i = it * PJ_THREADS + tid, f4 = i % G, r = i / G;
        const int n = m_blk * BM + split * F + f4 * 4;
        pre_b[it] = pre_x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R && r < BN) {
            if (a.bias) pre_b[it] = __ldg(reinterpret_cast<const float4*>(a.bias + n));
            if (a.x) pre_x[it] = __ldcg(reinterpret_cast<const float4*>(a.x + (size_t)r * N + n));
        }
    }
    if (warp == 1) {
        constexpr uint32_t idesc = pj_idesc(BN);
        int stage = 0; uint32_t phase = 0;
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
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 2) {
        // ---- park this CTA's partial tile in shared memory: lane = output feature, TMEM column = token row
        const int q = warp & 3;   // TMEM lane quadrant a warp may read: warps 2, 3, 4, 5 -> quadrants 2, 3, 0, 1
        const int f = q * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (c0 >= R) break;
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < R) park[(c0 + j) * PJ_STAGE_LD + f] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (g_trace) trace_end(trace_begin(150, a.W));   // this CTA's partial tile is parked
    cluster_sync_all();
    if (g_trace) trace_end(trace_begin(151, a.W));   // the whole cluster's tiles are parked

    // ---- cluster reduction + epilogue
    const uint32_t park_addr = smem_u32(park);
    const uint32_t stat0_addr = map_to_rank(smem_u32(statbuf), 0);
    const int items = G * BN;
#pragma unroll 1
    for (int it = 0; it * PJ_THREADS < items; ++it) {
        const int i0 = it * PJ_THREADS;
        if (i0 / G >= R) break;                       // the whole pass is past the last row (uniform: G divides 192)
        const int i = i0 + tid;
        const int f4 = i % G, r = i / G;
        const bool valid = r < R && r < BN;
        const int fl = split * F + f4 * 4;            // feature inside the 128-wide tile
        const int n = m_blk * BM + fl;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), xv = bv;
            if (it < PJ_PRE) {
                bv = it == 0 ? pre_b[0] : it == 1 ? pre_b[1] : pre_b[2];
                xv = it == 0 ? pre_x[0] : it == 1 ? pre_x[1] : pre_x[2];
            } else {
                if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + n));
                if (a.x) xv = __ldcg(reinterpret_cast<const float4*>(a.x + (size_t)r * N + n));
            }
            const uint32_t off = (uint32_t)(r * PJ_STAGE_LD + fl) * 4u;
            float4 t[8];
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < S) t[s] = ld_dsmem_f4(map_to_rank(park_addr, (uint32_t)s) + off);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < S) { v[0] += t[s].x; v[1] += t[s].y; v[2] += t[s].z; v[3] += t[s].w; }   // fixed order: deterministic
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
            if (a.act == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = gelu_tanh_fast(v[e]);
            }
            if (a.x) {
                v[0] += xv.x; v[1] += xv.y; v[2] += xv.z; v[3] += xv.w;
                *reinterpret_cast<float4*>(a.x + (size_t)r * N + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
            uint2 pk;
            pk.x = pack_bf16x2(v[0], v[1]);
            pk.y = pack_bf16x2(v[2], v[3]);
            if (a.out) *reinterpret_cast<uint2*>(a.out + (size_t)r * a.out_ld + n) = pk;
            if (a.rows && n >= a.d) {   // QKV: 4 consecutive columns never straddle a 64-wide head block
                const RowDesc rd = a.rows[r];
                const int c = n - a.d, which = c >= a.d, i2 = which ? c - a.d : c;
                bf16* dst = (which ? a.vpanel : a.kpanel) + (size_t)rd.kv_slot * a.slot_stride + ((size_t)(i2 >> 6) * a.n_pos_cap + rd.pos) * 64 + (i2 & 63);
                *reinterpret_cast<uint2*>(dst) = pk;
            }
        }
        if (a.stats_out) {
            // (mean, M2) of this rank's F features of row r; the G lanes of a row are consecutive lanes of one warp
            float s1 = (v[0] + v[1]) + (v[2] + v[3]);
            for (int o = G >> 1; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            const float mean = s1 / (float)F;
            const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean, d3 = v[3] - mean;
            float m2 = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            for (int o = G >> 1; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            if (valid && f4 == 0) st_dsmem_f2(stat0_addr + (uint32_t)(split * BN + r) * 8u, mean, m2);
        }
    }
    if (g_trace) trace_end(trace_begin(152, a.W));   // this CTA's share is reduced and written
    if (a.y) __threadfence();   // this thread's x values are visible device-wide before the cluster's ticket is drawn
    cluster_sync_all();         // peers are done reading this CTA's tile; rank 0 has everybody's statistics
    if (a.stats_out && split == 0) {
        if (tid < R && tid < BN) {
            float msum = 0.0f;
            for (int s = 0; s < S; ++s) msum += statbuf[s * BN + tid].x;
            const float mean = msum / (float)S;
            float m2 = 0.0f;
            for (int s = 0; s < S; ++s) {
                const float2 st = statbuf[s * BN + tid];
                const float dm = st.x - mean;
                m2 += st.y + (float)F * dm * dm;
            }
            a.stats_out[(size_t)m_blk * 128 + tid] = make_float2(mean, m2);
        }
    }
    if (a.y) {
        // ---- LayerNorm tail: the cluster that draws the last ticket normalises every row
        if (split == 0) {
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const int n_clusters = N / BM;
                const int t = atomicAdd(a.ticket, 1);
                const uint32_t last = t == n_clusters - 1 ? 1u : 0u;
                if (last) *a.ticket = 0;   // every other cluster has drawn: re-armed for the next launch on this lane
                const uint32_t fa = smem_u32(last_flag);
                for (int s = 0; s < S; ++s) st_dsmem_u32(map_to_rank(fa, (uint32_t)s), last);
            }
        }
        cluster_sync_all();
        if (g_trace) trace_end(trace_begin(153, a.W));   // ticket drawn
        if (*reinterpret_cast<volatile uint32_t*>(last_flag)) {
            __threadfence();
            const int parts = N / BM;
            for (int r = split * (PJ_THREADS / 32) + warp; r < R; r += S * (PJ_THREADS / 32)) {   // one warp per row
                float msum = 0.0f;
                for (int p0 = 0; p0 < parts; p0 += 32)
                    if (p0 + lane < parts) msum += __ldcg(&a.stats_out[(size_t)(p0 + lane) * 128 + r]).x;
                const float mean = warp_sum(msum) / (float)parts;
                float m2 = 0.0f;
                for (int p0 = 0; p0 < parts; p0 += 32) {
                    if (p0 + lane < parts) {
                        const float2 st = __ldcg(&a.stats_out[(size_t)(p0 + lane) * 128 + r]);
                        const float dm = st.x - mean;
                        m2 += st.y + (float)BM * dm * dm;
                    }
                }
                const float rstd = rsqrtf(warp_sum(m2) / (float)N + 1e-5f);
                const float* xr = a.x + (size_t)r * N;
                bf16* yr = a.y + (size_t)r * N;
                for (int c = lane * 4; c < N; c += 128) {
                    const float4 xv = __ldcg(reinterpret_cast<const float4*>(xr + c));
                    const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_g + c)), b = __ldg(reinterpret_cast<const float4*>(a.ln_b + c));
                    uint2 pk;
                    pk.x = pack_bf16x2((xv.x - mean) * rstd * g.x + b.x, (xv.y - mean) * rstd * g.y + b.y);
                    pk.y = pack_bf16x2((xv.z - mean) * rstd * g.z + b.z, (xv.w - mean) * rstd * g.w + b.w);
                    *reinterpret_cast<uint2*>(yr + c) = pk;
                }
            }
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    trace_end(tr);
}

thread_local std::string g_perr;

int pj_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN>
bool launch_proj_cfg(const ProjDesc& a, int S, cudaStream_t s) {
    using cfg = PjCfg<BN>;
    CUtensorMap tw, tx;
    if (!get_tmap_bf16_2d(&tw, a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K, BK, BM)) { g_perr = sm100_last_error(); return false; }
    if (!get_tmap_bf16_2d(&tx, a.X, (uint64_t)a.K, (uint64_t)a.R, (uint64_t)a.ldx, BK, BN)) { g_perr = sm100_last_error(); return false; }
    // ring depth: up to 4 stages (the whole K slice of the short projections is requested before the dependency wait) while the
    // CTA still fits beside another lane's attention CTAs
    const int nkb = (a.K / BK + S - 1) / S;
    static const int forced = [] { const char* v = getenv("NOBS_WHISPER_PROJ_STAGES"); return (v && *v) ? atoi(v) : 0; }();
    int stages = std::max(2, std::min(nkb, PJ_MAX_STAGES));
    while (stages > 2 && cfg::smem_bytes(stages) > 104 * 1024) --stages;
    if (forced >= 2 && forced <= PJ_MAX_STAGES) stages = forced;
    const int smem = cfg::smem_bytes(stages);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dec_proj_cluster_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::smem_bytes(PJ_MAX_STAGES)) != cudaSuccess) {
            g_perr = "cudaFuncSetAttribute(proj smem) failed";
            return false;
        }
        configured = true;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)((a.N / BM) * S));
    lc.blockDim = dim3(PJ_THREADS);
    lc.dynamicSmemBytes = (size_t)smem;
    lc.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)S; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (g_use_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    lc.attrs = attr;
    lc.numAttrs = na;
    const cudaError_t err = cudaLaunchKernelEx(&lc, dec_proj_cluster_kernel<BN>, tw, tx, a, stages);
    count_launch();
    if (err != cudaSuccess) { g_perr = std::string("proj launch: ") + cudaGetErrorString(err); return false; }
    return true;
}

}  // namespace

// cluster size (= K splits) of one fused projection: as many CTAs as fit one wave, a power of two <= 8, never more splits than k-blocks
int dec_proj_splits(int N, int K) {
    const int m_tiles = N / BM, num_k = K / BK;
    int S = 8;
    while (S > 1 && (m_tiles * S > pj_num_sms() || S > num_k)) S >>= 1;
    static const int cap = [] { const char* v = getenv("NOBS_WHISPER_PROJ_MAX_SPLITS"); return (v && *v) ? atoi(v) : 0; }();
    if (cap > 0) while (S > cap) S >>= 1;
    return S;
}

bool dec_proj_supported(int R, int N, int K) { return R > 0 && R <= 128 && N > 0 && K > 0 && N % BM == 0 && K % BK == 0 && N / BM <= 40; }

bool launch_dec_proj_sm100(const ProjDesc& a, cudaStream_t s) {
    if (!dec_proj_supported(a.R, a.N, a.K)) { sm100_set_error("dec_proj: unsupported shape"); return false; }
    if (!a.W || !a.X || (a.y && (!a.x || !a.stats_out || !a.ticket || !a.ln_g || !a.ln_b))) { sm100_set_error("dec_proj: missing operand"); return false; }
    const int S = dec_proj_splits(a.N, a.K);
    bool ok;
    if (a.R <= 32) ok = launch_proj_cfg<32>(a, S, s);
    else if (a.R <= 64) ok = launch_proj_cfg<64>(a, S, s);
    else ok = launch_proj_cfg<128>(a, S, s);
    if (!ok) sm100_set_error(g_perr);
    return ok;
}

void trace_set_proj(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}

}  // namespace nobs
