// Encoder self-attention for sm_100a (bf16 in/out, fp32 softmax and accumulation):
// non-causal, 1500 keys, head size 64, one CTA per (128-query tile, head, window).
//
//   TMA loads Q once and streams K/V blocks of 128 keys through a 3-stage ring (128B swizzle).
//   Per key block j, one elected thread issues
//        S_j = Q K_j^T        tcgen05.mma  M128 N128 K64   -> TMEM  (S double-buffered)
//        O_j = P_j V_j        tcgen05.mma  M128 N64  K128  -> TMEM  (V is the MN-major B operand, read
//                                                                    in place from the [key][dim] tile)
//   128 softmax threads each own one query row (TMEM lane): tcgen05.ld the row of S, running
//   max / exp2 / sum in registers (no shuffles), write P_j as bf16 into a 128B-swizzled K-major
//   tile in shared memory, and fold O_j into a register accumulator with the online-softmax
//   correction, so TMEM never needs rescaling.
//
//   Two builds of the same kernel (template parameter PER_SM):
//     1: one CTA per SM — S double-buffered in TMEM (512 columns allocated), 3-stage K/V ring, 145 KB of shared memory;
//     2: two CTAs per SM — S single-buffered (256 columns per CTA), 2-stage ring, 112 KB.  A CTA's softmax threads are a chain of
//        TMEM round trips and barrier waits (issue slots 44 % active with one CTA); the second CTA's chain fills those gaps.  S_{j+1}
//        is then issued right behind P_j V_j, i.e. once the softmax threads have read S_j.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "device_utils.cuh"
#include "gemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

constexpr int AQ = 128;        // queries per CTA
constexpr int AK = 128;        // keys per block
constexpr int DH = 64;         // head size
constexpr int TILE_BYTES = 128 * DH * 2;  // 16 KB: Q, K, V tiles and each 64-key half of P
// PER_SM = 2 has no alignment slack: two CTAs (2 x (112.25 KB + 1 KB reserved)) just fit the SM's 228 KB; the kernel checks the
// 1024-byte alignment the swizzled tiles need (a dynamic segment with no static shared memory in front of it starts aligned).
__host__ __device__ constexpr int att_kv_stages(int per_sm) { return per_sm == 1 ? 3 : 2; }
__host__ __device__ constexpr int att_smem(int per_sm) {
    return TILE_BYTES /*Q*/ + att_kv_stages(per_sm) * 2 * TILE_BYTES /*K,V*/ + 2 * TILE_BYTES /*P*/ + (per_sm == 1 ? 1024 : 0) + 256;
}

// instruction descriptors (kind::f16, D fp32, A/B bf16): S = Q K^T both K-major; O = P V with B MN-major
constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AK >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
constexpr uint32_t IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DH >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);

template <int PER_SM>
__global__ void __launch_bounds__(192, PER_SM)
enc_attention_sm100_kernel(const __grid_constant__ CUtensorMap tmap_qkv, bf16* __restrict__ out, int d, int n_valid) {
    constexpr int KV_STAGES = att_kv_stages(PER_SM);
    constexpr uint32_t TMEM_COLS_ATT = PER_SM == 1 ? 512 : 256;   // S: 2 x 128 (or 1 x 128), O: 2 x 64
    constexpr int S_BUFS = PER_SM == 1 ? 2 : 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    if (PER_SM != 1 && (raw & 1023u)) __trap();   // no slack to align with
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;                        // [stage][16 KB]
    uint8_t* sV = sK + KV_STAGES * TILE_BYTES;            // [stage][16 KB]
    uint8_t* sP = sV + KV_STAGES * TILE_BYTES;            // 2 x [128 rows][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
    uint64_t* q_full = bars;                   // 1
    uint64_t* kv_full = bars + 1;              // KV_STAGES
    uint64_t* kv_empty = kv_full + KV_STAGES;  // KV_STAGES
    uint64_t* s_full = kv_empty + KV_STAGES;   // 2   S_j landed in TMEM
    uint64_t* p_full = s_full + 2;             // 1   P_j written to smem (and S buffer j%2 drained)
    uint64_t* o_full = p_full + 1;             // 2   O_j landed in TMEM (and P buffer free again)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, w = blockIdx.z;
    const int row0 = w * kWinRows;                     // first row of this window in the qkv matrix
    const int n_blk = (n_valid + AK - 1) / AK;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&o_full[i], 1); }
        mbar_init(p_full, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS_ATT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;                 // columns [0, 128 * S_BUFS)
    const uint32_t tmem_O = tmem_base + S_BUFS * AK;   // 2 x 64 columns

    if (warp == 4) {
        // ===== TMA producer: Q once, then K/V blocks through the ring =====
        if (lane == 0) {
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_2d(sQ, &tmap_qkv, q_full, h * DH, row0 + qt * AQ);
        }
        int stage = 0; uint32_t phase = 0;
        for (int j = 0; j < n_blk; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&kv_full[stage], 2 * TILE_BYTES);
                tma_load_2d(sK + stage * TILE_BYTES, &tmap_qkv, &kv_full[stage], d + h * DH, row0 + j * AK);
                tma_load_2d(sV + stage * TILE_BYTES, &tmap_qkv, &kv_full[stage], 2 * d + h * DH, row0 + j * AK);
            }
            __syncwarp();
            if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
        auto issue_S = [&](int j) {  // S_j = Q K_j^T into S buffer j % 2
            const int stage = j % KV_STAGES;
            mbar_wait(&kv_full[stage], (uint32_t)((j / KV_STAGES) & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t k_addr = smem_u32(sK + stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    umma_bf16(tmem_S + (uint32_t)((j % S_BUFS) * AK), make_smem_desc_kmajor(q_addr + k * 32), make_smem_desc_kmajor(k_addr + k * 32), IDESC_S,
                              (uint32_t)(k != 0));
                umma_commit(&s_full[j % S_BUFS]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_S(0);
        if (S_BUFS == 2 && n_blk > 1) issue_S(1);
        for (int j = 0; j < n_blk; ++j) {
            mbar_wait(p_full, (uint32_t)(j & 1));   // P_j is in smem; the S buffer of block j has been read
            tc_fence_after();
            const int stage = j % KV_STAGES;
            if (lane == 0) {
                const uint32_t v_addr = smem_u32(sV + stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < AK / 16; ++k)   // 16 keys per MMA: A = P[:, 16k..], B = V[16k.., :] (MN-major)
                    umma_bf16(tmem_O + (uint32_t)((j & 1) * DH), make_smem_desc_kmajor(p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32),
                              make_smem_desc_mnmajor(v_addr + k * 16 * 128), IDESC_O, (uint32_t)(k != 0));
                umma_commit(&o_full[j & 1]);      // O_j complete (and P buffer reusable)
                umma_commit(&kv_empty[stage]);    // K_j / V_j consumed
            }
            __syncwarp();
            if (j + S_BUFS < n_blk) issue_S(j + S_BUFS);
        }
    } else {
        // ===== softmax + output: thread = one query row =====
        const int q = warp & 3;                      // TMEM lane quadrant (warps 0..3)
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        float m_run = -INFINITY, l_run = 0.0f;
        float O[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) O[i] = 0.0f;
        float corr_prev = 1.0f;
        uint8_t* p_row = sP + row * 128;
        const int sw = row & 7;
        for (int j = 0; j < n_blk; ++j) {
            mbar_wait(&s_full[j % S_BUFS], (uint32_t)((j / S_BUFS) & 1));
            tc_fence_after();
            // pass 1: row maximum of this block
            float bmax = -INFINITY;
            const int key0 = j * AK;
#pragma unroll 1
            for (int c0 = 0; c0 < AK; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_S + lane_addr + (uint32_t)((j % S_BUFS) * AK + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + c0 + i < n_valid) bmax = fmaxf(bmax, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m_run, bmax);
            const float corr = exp2f((m_run - m_new) * c);   // 0 on the first block
            const float mc = m_new * c;
            // O_{j-1} must be folded in (and its P buffer released) before P_j overwrites the buffer
            if (j > 0) {
                mbar_wait(&o_full[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < DH; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_O + lane_addr + (uint32_t)(((j - 1) & 1) * DH + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) O[c0 + i] = O[c0 + i] * corr_prev + __uint_as_float(v[i]);
                }
            }
            // pass 2: p = exp2(s*c - m*c), row sum, bf16 P into the swizzled K-major tile
            float bsum = 0.0f;
#pragma unroll 1
            for (int c0 = 0; c0 < AK; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_S + lane_addr + (uint32_t)((j % S_BUFS) * AK + c0), v);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = key0 + c0 + i < n_valid ? exp2f(__uint_as_float(v[i]) * c - mc) : 0.0f;
                    float p1 = key0 + c0 + i + 1 < n_valid ? exp2f(__uint_as_float(v[i + 1]) * c - mc) : 0.0f;
                    __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                    // accumulate the row sum from the rounded values so that P and l stay consistent
                    const float2 back = __bfloat1622float2(hh);
                    bsum += back.x + back.y;
                    packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                }
                // 32 keys = 64 bytes = four 16-byte chunks of this row inside K-half (c0 / 64)
                uint8_t* base = p_row + (c0 >> 6) * TILE_BYTES;
                const int chunk0 = (c0 & 63) >> 3;  // 0 or 4
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const uint4 val = make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
                    *reinterpret_cast<uint4*>(base + (((chunk0 + ch) ^ sw) << 4)) = val;
                }
            }
            l_run = l_run * corr + bsum;
            m_run = m_new;
            corr_prev = corr;
            // make the generic-proxy writes of P visible to the tensor core (async proxy), then signal
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(p_full);
        }
        {   // last block's O
            const int j = n_blk - 1;
            mbar_wait(&o_full[j & 1], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_O + lane_addr + (uint32_t)((j & 1) * DH + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) O[c0 + i] = O[c0 + i] * corr_prev + __uint_as_float(v[i]);
            }
        }
        const float inv = 1.0f / l_run;
        bf16* dst = out + (size_t)(row0 + qt * AQ + row) * d + h * DH;
#pragma unroll
        for (int i = 0; i < DH; i += 8) {
            uint4 u;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(O[i] * inv, O[i + 1] * inv), h1 = __floats2bfloat162_rn(O[i + 2] * inv, O[i + 3] * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(O[i + 4] * inv, O[i + 5] * inv), h3 = __floats2bfloat162_rn(O[i + 6] * inv, O[i + 7] * inv);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + i) = u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS_ATT);
    }
}

}  // namespace

bool launch_enc_attention_bf16_sm100(const bf16* qkv, bf16* out, int n_win, int n_head, int d, cudaStream_t s) {
    if (n_win <= 0) return true;
    CUtensorMap tm;
    // the whole qkv matrix [n_win*1536][3d]; one map serves Q, K and V tiles (64 columns x 128 rows)
    if (!make_tmap_bf16_2d(&tm, qkv, (uint64_t)3 * d, (uint64_t)n_win * kWinRows, (uint64_t)3 * d, 64, 128)) return false;
    // NOBS_WHISPER_ENC_ATT_PER_SM: 2 (default) = two CTAs per SM with a single S buffer each, 1 = one CTA per SM with S double-buffered
    static const int per_sm = [] { const char* v = getenv("NOBS_WHISPER_ENC_ATT_PER_SM"); return (v && *v == '1') ? 1 : 2; }();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(enc_attention_sm100_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, att_smem(1)) != cudaSuccess ||
            cudaFuncSetAttribute(enc_attention_sm100_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, att_smem(2)) != cudaSuccess) {
            sm100_set_error("cudaFuncSetAttribute(attention smem) failed");
            return false;
        }
        configured = true;
    }
    dim3 grid(kWinRows / AQ, n_head, n_win);
    if (per_sm == 1) {
        prefer_max_shared_carveout(reinterpret_cast<const void*>(&enc_attention_sm100_kernel<1>));
        enc_attention_sm100_kernel<1><<<grid, 192, att_smem(1), s>>>(tm, out, d, 1500);
    } else {
        prefer_max_shared_carveout(reinterpret_cast<const void*>(&enc_attention_sm100_kernel<2>));
        enc_attention_sm100_kernel<2><<<grid, 192, att_smem(2), s>>>(tm, out, d, 1500);
    }
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { sm100_set_error(std::string("attention launch: ") + cudaGetErrorString(err)); return false; }
    return true;
}

}  // namespace nobs
