// bf16 GEMM for sm_100a: C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 accumulation.
//
//   TMA (cp.async.bulk.tensor, 128B swizzle)  ->  shared-memory ring (STAGES x {A 128x64, W BNx64})
//   tcgen05.mma.cta_group::1.kind::f16 issued by one thread, accumulators in TMEM (2 x BN columns,
//   double buffered)  ->  tcgen05.ld by 4 epilogue warps  ->  fused bias / GELU / residual /
//   row-mask epilogue  ->  global (bf16 or fp32).
//
// Persistent: one CTA per SM walks a static tile schedule (m fastest, so CTAs that run together
// share the same weight tile in L2).  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM
// allocator, 4..7 = epilogue (warp % 4 selects the TMEM lane quadrant).
//
// The A operand may be an "overlapping rows" view (row stride smaller than the row length):
// that is how both stem convolutions run as GEMMs without materialising im2col
// (conv1: stride n_mels, K = 3*n_mels; conv2: stride 2d, K = 3d).
#include "gemm_sm100.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <string>

#include "device_utils.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

thread_local std::string g_err;

constexpr int BM = 128;      // UMMA M
constexpr int BK = 64;       // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N x 128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <typename TC> struct OutVec;
template <> struct OutVec<float> {
    static __device__ __forceinline__ void store8(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct OutVec<bf16> {
    static __device__ __forceinline__ void store8(bf16* p, const float* v) {
        uint4 u;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

template <int BN> struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32,64,128,256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN, typename TC>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, TC* C, int ldc, int M, int N,
                       int K, Epilogue e, int vec_ok) {
    using cfg = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);  // 1024-byte alignment for the 128B swizzle atoms
    uint8_t* sA = smem;
    uint8_t* sB = smem + cfg::STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE_BYTES);
    uint64_t* empty = full + cfg::STAGES;
    uint64_t* tfull = empty + cfg::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tr = trace_begin(5, C);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // everything above overlapped the previous kernel's tail
    pdl_launch_dependents();   // the successor's prologue overlaps this kernel's work (never more than one kernel parked ahead)

    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n, num_k = (K + BK - 1) / BK;

    if (warp == 0) {
        // ===== TMA producer =====
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&full[stage], cfg::STAGE_BYTES);
                    tma_load_2d(sA + stage * A_BYTES, &tmap_a, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sB + stage * cfg::B_BYTES, &tmap_b, &full[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(sA + stage * A_BYTES), b_addr = smem_u32(sB + stage * cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(tmem_d, make_smem_desc_kmajor(a_addr + k * UMMA_K * 2), make_smem_desc_kmajor(b_addr + k * UMMA_K * 2), idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[stage]);                    // smem slot is free once these MMAs retire
                    if (kb == num_k - 1) umma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias / GELU / residual / mask -> global =====
        const int q = warp & 3;              // TMEM lane quadrant this warp may read
        const int row = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int m = m_blk * BM + row;
            const bool row_ok = m < M;
            const bool row_zero = e.win_rows > 0 && (m % e.win_rows) >= e.valid_rows;
            const float* res_row = nullptr;
            if (e.res && row_ok) res_row = e.res + (size_t)(e.res_mod > 0 ? m % e.res_mod : m) * e.res_ld;
            TC* c_row = C + (size_t)m * ldc;
            size_t head_base = 0;
            int head_pos = 0;
            if (e.head_rows > 0) {  // head-major (KV-cache) output layout
                const int w = m / e.head_rows;
                head_pos = m - w * e.head_rows;
                head_base = (size_t)w * e.head_rows * ldc;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                tmem_ld_wait();
                const int n0 = n_blk * BN + c0;
                if (row_ok && n0 < N) {
                    float o[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + j;
                        float x = __uint_as_float(v[j]);
                        if (n < N) {
                            if (e.bias) x += __ldg(e.bias + n);
                            if (e.act == 1) x = gelu_tanh_fast(x);
                            if (res_row) x += res_row[n];
                        }
                        o[j] = row_zero ? 0.0f : x;
                    }
                    // 32 consecutive columns never straddle a 64-wide head block (n0 % 32 == 0)
                    TC* dst = e.head_rows > 0 ? C + head_base + ((size_t)(n0 >> 6) * e.head_rows + head_pos) * 64 + (n0 & 63) : c_row + n0;
                    if (vec_ok && n0 + 32 <= N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) OutVec<TC>::store8(dst + j, o + j);
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < N) dst[j] = from_f32<TC>(o[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, cfg::TMEM_COLS);
    }
    trace_end(tr);
}

// ------------------------------------------------------------------------------------------
// Skinny GEMM for decoder steps (R <= 128 token rows): swap-AB + split-K.
//   D^T[n, r] = sum_k W[n, k] * X[r, k]   — the weight tile is the 128-row UMMA "A" operand, the R token
//   rows are the (small) "B" operand, so a CTA streams 16 KB of weights per k-block and the activations
//   ride along from L2.  The K range is split across CTAs so that every SM pulls weights from HBM
//   (N_out/128 tiles alone would occupy 10-40 SMs); fp32 partial sums go to a workspace
//   [split][R][N_out] and skinny_reduce_kernel (kernels.cu) sums them in a fixed order and applies
//   bias / GELU / residual / LayerNorm / KV-cache scatter.
constexpr int kSkinnyStagesMax = 3;
int g_skinny_stages = kSkinnyStagesMax;

template <int BN, int STAGES>
__global__ void __launch_bounds__(256, 2)
gemm_skinny_sm100_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, float* __restrict__ partial, int R,
                         int N, int K, int kb_per_split, int w_keep) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, split = blockIdx.y;
    const long long tr = trace_begin(1, partial);
    const int num_k = (K + BK - 1) / BK;
    const int kb0 = split * kb_per_split, kb1 = min(num_k, kb0 + kb_per_split);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
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

    if (warp == 0) {
        // The weights do not depend on the previous kernel: fill the ring with weight tiles first, and only
        // then wait for the predecessor (PDL) before fetching the activations it produced.
        const int pre = min(STAGES, kb1 - kb0);
        // decoder weights are re-read by every decode lane within a fraction of a layer time: keep them in L2
        // (the cross-KV stream next to them is evict_first)
        const uint64_t w_policy = w_keep ? l2_evict_last_policy() : l2_evict_normal_policy();
        if (lane == 0) {
            for (int i = 0; i < pre; ++i) {
                mbar_expect_tx(&full[i], STAGE_BYTES);
                tma_load_2d_hint(sA + i * A_BYTES, &tmap_w, &full[i], (kb0 + i) * BK, m_blk * BM, w_policy);
            }
        }
        pdl_wait();
        pdl_launch_dependents();   // successor prologue overlaps this kernel's work; never more than one kernel parked ahead
        trace_end(trace_begin(101, partial));
        if (lane == 0) {
            for (int i = 0; i < pre; ++i) tma_load_2d(sB + i * B_BYTES, &tmap_x, &full[i], (kb0 + i) * BK, 0);
        }
        __syncwarp();
        int stage = pre % STAGES; uint32_t phase = pre == STAGES ? 1 : 0;
        for (int kb = kb0 + pre; kb < kb1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                tma_load_2d_hint(sA + stage * A_BYTES, &tmap_w, &full[stage], kb * BK, m_blk * BM, w_policy);
                tma_load_2d(sB + stage * B_BYTES, &tmap_x, &full[stage], kb * BK, 0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN);
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
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4 && kb1 > kb0) {
        // partial[split][r][n]: lane = output feature n, TMEM column = token row r -> coalesced over n
        const int q = warp & 3;
        const int n = m_blk * BM + q * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float* dst = partial + (size_t)split * R * N + n;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (n < N) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < R) dst[(size_t)(c0 + j) * N] = __uint_as_float(v[j]);
            }
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

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

void trace_set_gemm(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}

void sm100_set_error(const std::string& e) { g_err = e; }

// 2D bf16 tensor [rows][inner] with an arbitrary (16-byte multiple) row stride — smaller than the row
// length for the overlapping-row views; box = box_inner x box_rows, 128B swizzle, OOB reads as zero
bool make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_inner,
                       uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { g_err = "cuTensorMapEncodeTiled is not available"; return false; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * 2) % 16) { g_err = "TMA operand is not 16-byte aligned"; return false; }
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (inner " + std::to_string(inner) + ", rows " +
                std::to_string(rows) + ", stride " + std::to_string(row_stride_elems) + ")";
        return false;
    }
    return true;
}

namespace {
struct TmapKey {
    const void* ptr; uint64_t inner, rows, stride; uint32_t box_inner, box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && inner == o.inner && rows == o.rows && stride == o.stride && box_inner == o.box_inner && box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= (k.inner * 0xC2B2AE3D27D4EB4Full) ^ (k.rows << 17) ^ (k.stride << 29) ^ ((uint64_t)k.box_inner << 41) ^ ((uint64_t)k.box_rows << 52);
        return (size_t)(h ^ (h >> 31));
    }
};
}  // namespace

bool get_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_inner,
                      uint32_t box_rows) {
    thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, inner, rows, row_stride_elems, box_inner, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) { *map = it->second; return true; }
    if (!make_tmap_bf16_2d(map, ptr, inner, rows, row_stride_elems, box_inner, box_rows)) return false;
    if (cache.size() > 65536) cache.clear();   // a descriptor is a pure function of its key: dropping the cache is always safe
    cache.emplace(key, *map);
    return true;
}

namespace {

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, typename TC>
bool launch_cfg(const bf16* A, int lda, const bf16* W, int ldw, TC* C, int ldc, int M, int N, int K, const Epilogue& e, cudaStream_t s) {
    using cfg = Cfg<BN>;
    CUtensorMap ta, tb;
    if (!get_tmap_bf16_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return false;
    if (!get_tmap_bf16_2d(&tb, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BN)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_bf16_sm100_kernel<BN, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES) != cudaSuccess) {
            g_err = "cudaFuncSetAttribute(max dynamic smem) failed";
            return false;
        }
        configured = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    const int vec = (sizeof(TC) == 2 ? 8 : 4);
    const int vec_ok = (ldc % vec == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    launch_kernel(gemm_bf16_sm100_kernel<BN, TC>, dim3(grid), dim3(256), (size_t)cfg::SMEM_BYTES, s, true, ta, tb, C, ldc, M, N, K, e, vec_ok);
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { g_err = std::string("gemm launch: ") + cudaGetErrorString(err); return false; }
    return true;
}


template <int BN, int STAGES>
bool launch_skinny_cfg(const bf16* X, int ldx, const bf16* W, int ldw, float* partial, int R, int N, int K, int splits, int kb_per_split,
                       cudaStream_t s) {
    constexpr int SMEM = STAGES * (A_BYTES + BN * BK * 2) + 1024 + 256;
    CUtensorMap tw, tx;
    if (!get_tmap_bf16_2d(&tw, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, BK, BM)) return false;
    if (!get_tmap_bf16_2d(&tx, X, (uint64_t)K, (uint64_t)R, (uint64_t)ldx, BK, BN)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_skinny_sm100_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            g_err = "cudaFuncSetAttribute(skinny smem) failed";
            return false;
        }
        configured = true;
    }
    dim3 grid((N + BM - 1) / BM, splits);
    static const int w_keep = [] { const char* v = getenv("NOBS_WHISPER_W_EVICT_LAST"); return (v && *v == '1') ? 1 : 0; }();
    launch_kernel(gemm_skinny_sm100_kernel<BN, STAGES>, grid, dim3(256), (size_t)SMEM, s, true, tw, tx, partial, R, N, K, kb_per_split, w_keep);
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { g_err = std::string("skinny gemm launch: ") + cudaGetErrorString(err); return false; }
    return true;
}
// ring depth: 3 stages, or 2 (set_skinny_gemm_stages) to shrink the footprint so that two such CTAs fit beside the
// attention CTAs of other decode lanes
template <int BN>
bool launch_skinny_bn(const bf16* X, int ldx, const bf16* W, int ldw, float* partial, int R, int N, int K, int splits, int kb_per_split,
                      cudaStream_t s) {
    if (g_skinny_stages == 2) return launch_skinny_cfg<BN, 2>(X, ldx, W, ldw, partial, R, N, K, splits, kb_per_split, s);
    return launch_skinny_cfg<BN, 3>(X, ldx, W, ldw, partial, R, N, K, splits, kb_per_split, s);
}

}  // namespace

void set_skinny_gemm_stages(int stages) { g_skinny_stages = stages == 2 ? 2 : kSkinnyStagesMax; }

int skinny_gemm_splits(int N, int K) {
    const int m_tiles = (N + BM - 1) / BM, num_k = (K + BK - 1) / BK;
    int splits = num_sms() / m_tiles;                       // one wave: never more CTAs than SMs ...
    splits = std::max(1, std::min(splits, num_k / 2));      // ... and at least two k-blocks per CTA
    static const int cap = [] { const char* v = getenv("NOBS_WHISPER_SKINNY_MAX_SPLITS"); return (v && *v) ? atoi(v) : 0; }();
    if (cap > 0) splits = std::min(splits, cap);
    const int kb_per = (num_k + splits - 1) / splits;
    return (num_k + kb_per - 1) / kb_per;                   // no empty splits
}

bool launch_gemm_skinny_bf16_sm100(const bf16* X, int ldx, const bf16* W, int ldw, float* partial, int R, int N, int K, int* splits_out,
                                   cudaStream_t s) {
    if (R <= 0 || R > 128 || N <= 0 || K <= 0) { g_err = "skinny gemm: bad shape"; return false; }
    const int num_k = (K + BK - 1) / BK;
    const int splits = skinny_gemm_splits(N, K);
    const int kb_per = (num_k + splits - 1) / splits;
    *splits_out = splits;
    if (R <= 32) return launch_skinny_bn<32>(X, ldx, W, ldw, partial, R, N, K, splits, kb_per, s);
    if (R <= 64) return launch_skinny_bn<64>(X, ldx, W, ldw, partial, R, N, K, splits, kb_per, s);
    return launch_skinny_bn<128>(X, ldx, W, ldw, partial, R, N, K, splits, kb_per, s);
}

namespace {

template <typename TC>
bool dispatch(const bf16* A, int lda, const bf16* W, int ldw, TC* C, int ldc, int M, int N, int K, const Epilogue& e, cudaStream_t s) {
    const int sms = num_sms(), mt = (M + BM - 1) / BM;
    auto tiles = [&](int bn) { return mt * ((N + bn - 1) / bn); };
    // widest tile that still gives every SM work; skinny (decoder) problems fall to BN = 32 so that
    // the weight stream is spread over as many SMs as possible
    if (tiles(256) >= sms) {
        // wave quantisation: N = 1280 at M = 12288 is 480 tiles = 3.24 waves of 148 CTAs (81 % busy); the same problem
        // in 128-wide tiles is 6.49 waves (93 %) at ~8 % more operand traffic per FLOP
        static const int prefer128 = [] { const char* v = getenv("NOBS_WHISPER_GEMM_WAVE_BN128"); return (v && *v == '0') ? 0 : 1; }();
        auto eff = [&](int bn) { const int t = tiles(bn); return (double)t / (double)(((t + sms - 1) / sms) * sms); };
        if (prefer128 && eff(128) * 0.92 > eff(256)) return launch_cfg<128, TC>(A, lda, W, ldw, C, ldc, M, N, K, e, s);
        return launch_cfg<256, TC>(A, lda, W, ldw, C, ldc, M, N, K, e, s);
    }
    if (tiles(128) >= sms) return launch_cfg<128, TC>(A, lda, W, ldw, C, ldc, M, N, K, e, s);
    if (tiles(64) >= sms) return launch_cfg<64, TC>(A, lda, W, ldw, C, ldc, M, N, K, e, s);
    return launch_cfg<32, TC>(A, lda, W, ldw, C, ldc, M, N, K, e, s);
}

}  // namespace

bool launch_gemm_bf16_sm100(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, bool c_is_f32, int M, int N, int K, const Epilogue& e,
                            cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return true;
    if (c_is_f32) return dispatch<float>(A, lda, W, ldw, static_cast<float*>(C), ldc, M, N, K, e, s);
    return dispatch<bf16>(A, lda, W, ldw, static_cast<bf16*>(C), ldc, M, N, K, e, s);
}

const char* sm100_last_error() { return g_err.c_str(); }

}  // namespace nobs
