// Decoder cross-attention for single-token rows (bf16), the dominant HBM stream of a decoder step:
// every (row, head) reads one contiguous 1500 x 64 K panel and one V panel (192 KB each) of the
// head-major cross-KV pool exactly once.  Both kernels here are persistent and deliberately small, so that the
// latency-bound projection kernels of ANOTHER decode lane (engine.cu) stay resident on the same SMs while the
// stream keeps HBM busy.
//
// 1. dec_cross_attention_tc_kernel (default, second half of this file): TMA tensor loads into a swizzled ring,
//    scores and P*V on tcgen05, 4 softmax warps.  14 % of the issue slots, 128 TMEM columns, ~57 KB per CTA.
// 2. dec_cross_attention_sm100_kernel (NOBS_WHISPER_CROSS_MODE=1, kept as the CUDA-core variant of the same
//    stream): a producer thread issues cp.async.bulk (TMA, non-tensor) copies of 16 KB panel chunks into a
//    shared-memory ring, L2 evict_first, running ahead across (row, head) items so the softmax barriers of an
//    item never drain the memory pipeline; K chunks do not depend on the predecessor kernel and are requested
//    before the programmatic-dependent-launch wait.  Consumer warps: 8 lanes x 16 B per key row (conflict-free
//    512-byte warp reads), scores to shared memory, block softmax, P*V with the same mapping, deterministic
//    cross-warp sum.  Ring depth / consumer warps / CTAs per SM are template parameters (sweep in profiles/).
#include <cuda.h>

#include <cstdlib>

#include "device_utils.cuh"
#include "gemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace nobs {

namespace {

constexpr int CA_CHUNK_KEYS = 128;
constexpr int CA_ROW_BYTES = 64 * 2;
constexpr int CA_CHUNK_BYTES = CA_CHUNK_KEYS * CA_ROW_BYTES;  // 16 KB
constexpr int CA_MAX_KEYS = kWinRows;
constexpr int CA_MAX_WARPS = 16;
constexpr int ca_smem_bytes(int stages, int warps) { return stages * CA_CHUNK_BYTES + CA_MAX_KEYS * 4 + warps * 64 * 4 + 2 * CA_MAX_WARPS * 4 + 2 * stages * 8 + 128; }

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
template <int THREADS>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
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

template <int CA_STAGES, int CA_WARPS>
__global__ void __launch_bounds__((CA_WARPS + 1) * 32, 1)
dec_cross_attention_sm100_kernel(const RowDesc* __restrict__ rows, int n_items, int n_head, const bf16* __restrict__ q, int ldq,
                                 const bf16* __restrict__ kc, const bf16* __restrict__ vc, bf16* __restrict__ out, int ldo, size_t slot_stride,
                                 size_t head_stride, int n_keys) {
    extern __shared__ __align__(128) uint8_t ca_smem[];
    uint8_t* ring = ca_smem;
    float* sc = reinterpret_cast<float*>(ring + CA_STAGES * CA_CHUNK_BYTES);
    float* part = sc + CA_MAX_KEYS;         // [CA_WARPS][64]
    float* red = part + CA_WARPS * 64;      // [2][CA_MAX_WARPS]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * CA_MAX_WARPS);
    constexpr int ITERS = CA_CHUNK_KEYS / (CA_WARPS * 4);
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
            uint4 raw[ITERS];
#pragma unroll
            for (int i = 0; i < ITERS; ++i) raw[i] = lds128(cb + ((i * CA_WARPS + warp) * 4 + ks) * CA_ROW_BYTES + sub * 16);
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
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
        consumer_sync<CA_WARPS * 32>();
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
        if (lane == 0) red[CA_MAX_WARPS + warp] = lsum;
        consumer_sync<CA_WARPS * 32>();
        float total = 0.0f;
#pragma unroll
        for (int w = 0; w < CA_WARPS; ++w) total += red[CA_MAX_WARPS + w];   // fixed order
        const float inv = 1.0f / total;
        // ---- P * V
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&full[stage], phase);
            const uint8_t* cb = ring + stage * CA_CHUNK_BYTES;
            uint4 raw[ITERS];
            float p[ITERS];
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int jl = (i * CA_WARPS + warp) * 4 + ks, j = c * CA_CHUNK_KEYS + jl;
                raw[i] = lds128(cb + jl * CA_ROW_BYTES + sub * 16);
                p[i] = j < n_keys ? sc[j] : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
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
        consumer_sync<CA_WARPS * 32>();
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


// ------------------------------------------------------------------------------------------
// Tensor-core variant: the same stream, but the shared-memory panels are read by tcgen05.mma instead of
// by 256 threads, so the kernel leaves the SM's issue slots to the co-resident projection kernels.
//
//   S^T[128 keys, 16]  = K_chunk[128 x 64] * qB[16 x 64]^T     (row 0 of qB = q, rows 1..15 are don't-care:
//                                                                 column n of D depends on row n of B only)
//   O  [128, 64 dims] += P_tile[128 x 16 keys] * V_chunk[16 keys x 64]   (row 0 of P_tile = p, rows 1..127
//                                                                 are whatever follows in shared memory:
//                                                                 lane m of D depends on row m of A only)
//   K/V chunks arrive through a TMA tensor map with the 128-byte swizzle the UMMA descriptors expect; V is
//   the MN-major B operand, read in place.  All 12 score chunks of an item sit in TMEM (60 columns at the
//   default pitch of 4, see SP below); 128 threads (thread = TMEM lane = key within chunk) do the softmax on 12
//   registers each and write p as bf16 into one linear row.  Warps 0-3 softmax/output (thread 0 also hands out
//   items from a device-wide counter), warp 4 TMA producer, warp 5 MMA issuer.  Two CTAs per SM, so that one
//   streams while the other is in its softmax.
// p of all 1536 keys for the GM rows of a group: [key / 8][GM][8] bf16 — for GM = 1 one linear row (3 KB used of 4)
__host__ __device__ constexpr int tc_p_bytes(int gm) { return gm == 1 ? 4096 : 1536 * 2 * gm; }
constexpr int TC_Q_BYTES = 2 * 2048;     // two q tiles (double buffered across items)
__host__ __device__ constexpr int tc_smem_bytes(int stages, int gm = 1) { return tc_p_bytes(gm) + TC_Q_BYTES + stages * CA_CHUNK_BYTES + 384 + 1024; }
// K-major operand WITHOUT swizzle whose 8x16-byte core matrices overlap at a 16-byte pitch (LBO = 16 B): row 0 of
// the tile is then a plain linear array; rows 1.. are the bytes that follow (don't-care rows).
// With GM rows per group the rows of a core matrix are real up to row GM - 1: p is stored [key / 8][GM][8 keys], i.e. the core matrix of
// a key block holds rows 0 .. GM-1 (then whatever follows), and the next key block's core matrix starts GM x 16 bytes later (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_p(uint32_t saddr, uint32_t gm) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)gm << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
constexpr uint32_t TC_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TC_IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ float tmem_ld_1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return __uint_as_float(v);
}

// SP: TMEM column pitch between the score blocks of consecutive chunks.  16 = disjoint blocks (256 columns
// allocated); 4 = overlapping blocks written in increasing order, each MMA's don't-care columns are overwritten by
// the following chunk's real column (128 columns allocated, two CTAs fit next to the projection GEMMs).
// MINB: minimum resident CTAs per SM the register allocation is sized for.  4 caps the kernel at 80 registers (114 without a cap,
// no spills either way): two of these CTAs then take 30.7 K of the SM's 64 K registers instead of 46 K, which is what lets the
// projection kernels of TWO other decode lanes (or a kernel and its programmatic successor) be resident beside them.
// GM: rows of ONE audio handled per item (1, 2 or 4).  Consecutive rows with the same audio slot (a temperature pass and its speculative
// successor, the beams of a beam search) attend over the same K / V panels: a group of up to GM of them is one item — their queries fill
// rows 0 .. g-1 of the 16-row q tile (the score MMA computes those columns anyway), their probabilities rows 0 .. g-1 of the P operand,
// and the panels are streamed from HBM / L2 into shared memory ONCE for the group.  Needs SP >= GM (score columns c * SP + row).  The
// groups come from the host (`groups[i]` = first row | size << 24, cross_attention_groups()); an item is then a (group, head) pair.
template <int TC_STAGES, int SP, int MINB, int GM>
__global__ void __launch_bounds__(192, MINB)
dec_cross_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_kv, const RowDesc* __restrict__ rows, int n_items, int n_head,
                              const bf16* __restrict__ q, int ldq, bf16* __restrict__ out, int ldo, long long k_row0, long long v_row0,
                              long long slot_rows, int n_keys, int* __restrict__ sched, CrossQPartials qp, const int* __restrict__ groups) {
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    uint8_t* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    constexpr uint32_t TC_TMEM_COLS = SP == 16 ? 256 : 128;
    constexpr uint32_t O_COL = SP == 16 ? 192 : 64;
    constexpr int QD = 4;                             // item queue depth (the scheduler runs two items ahead)
    static_assert(GM == 1 || GM == 2 || GM == 4, "group size");
    static_assert(SP >= GM, "score blocks of consecutive chunks would overwrite a group row's column");
    uint8_t* sP = smem;                               // [1536 / 8][GM][8] bf16
    constexpr int P_BYTES = tc_p_bytes(GM);
    uint8_t* sQ = sP + P_BYTES;                       // [2][2 KB]
    uint8_t* ring = sQ + TC_Q_BYTES;                  // [TC_STAGES][16 KB]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + TC_STAGES * CA_CHUNK_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* q_full = empty + TC_STAGES;             // 2
    uint64_t* s_full = q_full + 2;
    uint64_t* p_full = s_full + 1;
    uint64_t* o_full = p_full + 1;
    uint64_t* item_full = o_full + 1;                 // QD
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_full + QD);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);   // [8]
    int* item_q = reinterpret_cast<int*>(red + 8);          // [QD]: item | group size << 24, -1 ends the stream
    float* inv_s = reinterpret_cast<float*>(item_q + QD);   // [4]: 1 / sum of a group row's probabilities

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tr = trace_begin(4, out);
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_kv);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&q_full[0], GM == 1 ? 1 : 4); mbar_init(&q_full[1], GM == 1 ? 1 : 4);
        mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
        for (int i = 0; i < QD; ++i) mbar_init(&item_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + O_COL;
    const int n_chunks = (n_keys + CA_CHUNK_KEYS - 1) / CA_CHUNK_KEYS;

    // Items ((row, head) pairs) are handed out by a device-wide counter: CTAs that become resident late (another
    // lane's kernels hold the SM) simply take what is left, nothing is statically owned.  The scheduler thread
    // publishes item `it` in item_q[it % QD]; -1 ends the stream.
    auto next_item = [&](int it) -> int {
        mbar_wait(&item_full[it % QD], (uint32_t)((it / QD) & 1));
        return item_q[it % QD];
    };

    if (warp == 4) {
        // ===== TMA producer: runs ahead across items; nothing here depends on the predecessor kernel =====
        if (lane == 0) {
            const uint64_t policy = l2_evict_first_policy();
            int stage = 0; uint32_t phase = 0;
            for (int it = 0;; ++it) {
                const int packed = next_item(it);
                if (packed < 0) break;
                const int item = packed & 0xFFFFFF;
                const int r = item / n_head, h = item - r * n_head;
                const long long base = (long long)rows[r].audio_slot * slot_rows + (long long)h * kWinRows;
                for (int pass = 0; pass < 2; ++pass) {
                    const long long row0 = base + (pass ? v_row0 : k_row0);
                    for (int c = 0; c < n_chunks; ++c) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], CA_CHUNK_BYTES);
                        tma_load_2d_hint(ring + stage * CA_CHUNK_BYTES, &tmap_kv, &full[stage], 0, (int)(row0 + c * CA_CHUNK_KEYS), policy);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            pdl_launch_dependents();
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        int stage = 0; uint32_t phase = 0;
        for (int it = 0;; ++it) {
            if (next_item(it) < 0) break;
            const uint32_t qb_addr = smem_u32(sQ + (it & 1) * 2048);
            mbar_wait(&q_full[it & 1], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t k_addr = smem_u32(ring + stage * CA_CHUNK_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_S + (uint32_t)(c * SP), make_smem_desc_kmajor(k_addr + k * 32), make_smem_desc_kmajor(qb_addr + k * 32), TC_IDESC_S,
                                  (uint32_t)(k != 0));
                    umma_commit(&empty[stage]);
                    if (c == n_chunks - 1) umma_commit(s_full);
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            mbar_wait(p_full, (uint32_t)(it & 1));
            tc_fence_after();
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t v_addr = smem_u32(ring + stage * CA_CHUNK_BYTES), p_addr = smem_u32(sP + c * CA_CHUNK_KEYS * 2 * GM);
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // 16 keys per MMA
                        umma_bf16(tmem_O, make_smem_desc_p(p_addr + k * 32 * GM, GM), make_smem_desc_mnmajor(v_addr + k * 16 * 128),
                                  TC_IDESC_O, (uint32_t)((c | k) != 0));
                    umma_commit(&empty[stage]);
                    if (c == n_chunks - 1) umma_commit(o_full);
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== softmax + output: thread = TMEM lane = key within a chunk; thread 0 is also the item scheduler =====
        const int row = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        auto fetch = [&](int it) {   // thread 0 only
            int item = atomicAdd(&sched[0], 1), g = 1;
            if (item >= n_items) item = -1;
            else if (GM > 1) {
                // grouped launch: item = (group, head); the host's group table names the group's first row and its size
                const int grp = item / n_head, e = __ldg(groups + grp);
                g = e >> 24;
                item = (e & 0xFFFFFF) * n_head + (item - grp * n_head);
            }
            item_q[it % QD] = item < 0 ? -1 : (item | (g << 24));
            mbar_arrive(&item_full[it % QD]);   // release: the queue entry is visible to whoever completes the wait
        };
        if (threadIdx.x == 0) { fetch(0); fetch(1); }   // the counter and the row descriptors do not depend on the predecessor kernel
        pdl_wait();   // q does
        trace_end(trace_begin(104, out));
        // 128 bytes of q per group row -> rows 0 .. g-1 of the q tile (K-major, 128B swizzle: 16-byte chunk j of row i sits at chunk j ^ i;
        // row 0 is stored linearly).  Warp i writes row i.
        auto put_q = [&](int packed, int buf) {
            const int item = packed & 0xFFFFFF, g = GM == 1 ? 1 : packed >> 24;
            if (warp < g) {
                const int r = item / n_head + warp, h = item % n_head;
                uint8_t* qrow = sQ + buf * 2048 + warp * 128;
                if (qp.partial) {
                    // q is still the split-K partial sums of the query projection: finish it here (fixed split order,
                    // + bias) instead of in a separate epilogue launch; lane -> dims 2*lane, 2*lane + 1
                    const float* src = qp.partial + (size_t)r * qp.ld + h * 64 + 2 * lane;
                    float2 acc = make_float2(0.f, 0.f);
                    for (int sp = 0; sp < qp.splits; ++sp) {
                        const float2 t = __ldcg(reinterpret_cast<const float2*>(src + (size_t)sp * qp.plane));
                        acc.x += t.x; acc.y += t.y;
                    }
                    if (qp.bias) { acc.x += __ldg(qp.bias + h * 64 + 2 * lane); acc.y += __ldg(qp.bias + h * 64 + 2 * lane + 1); }
                    const __nv_bfloat162 hv = __floats2bfloat162_rn(acc.x, acc.y);
                    *reinterpret_cast<__nv_bfloat162*>(qrow + (((lane >> 2) ^ warp) << 4) + (lane & 3) * 4) = hv;
                } else if (lane < 8) {
                    const uint4 v = *reinterpret_cast<const uint4*>(q + (size_t)r * ldq + h * 64 + lane * 8);
                    *reinterpret_cast<uint4*>(qrow + ((lane ^ warp) << 4)) = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (GM == 1) {
                if (warp == 0) { __syncwarp(); if (lane == 0) mbar_arrive(&q_full[buf]); }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_full[buf]);   // one arrival per softmax warp, whether it had a row to write or not
            }
        };
        int packed = next_item(0);
        if (packed >= 0) put_q(packed, 0);
        for (int it = 0; packed >= 0; ++it) {
            const int item = packed & 0xFFFFFF, g = GM == 1 ? 1 : packed >> 24;
            const int r = item / n_head, h = item - r * n_head;
            mbar_wait(s_full, (uint32_t)(it & 1));
            tc_fence_after();
            int next = -1;
#pragma unroll 1
            for (int gi = 0; gi < g; ++gi) {
                float sv[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) sv[c] = c < n_chunks ? tmem_ld_1(tmem_S + lane_addr + (uint32_t)(c * SP + gi)) : 0.0f;
                tmem_ld_wait();
                float lmax = -INFINITY;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    sv[c] = (c * CA_CHUNK_KEYS + row < n_keys) ? sv[c] * 0.125f : -INFINITY;
                    lmax = fmaxf(lmax, sv[c]);
                }
                lmax = warp_max(lmax);
                if (lane == 0) red[warp] = lmax;
                consumer_sync<128>();
                const float mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
                float lsum = 0.0f;
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const bf16 pb = __float2bfloat16_rn(expf(sv[c] - mx));   // exp(-inf) = 0 for masked keys
                    lsum += __bfloat162float(pb);                            // the sum of what the tensor core will see
                    const int key = c * CA_CHUNK_KEYS + row;
                    if (c < n_chunks) reinterpret_cast<bf16*>(sP)[((key >> 3) * GM + gi) * 8 + (key & 7)] = pb;
                }
                if (gi == g - 1) {
                    // the last row's probabilities are in: hand P to the tensor core
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tc_fence_before();
                    mbar_arrive(p_full);
                }
                lsum = warp_sum(lsum);
                if (lane == 0) red[4 + warp] = lsum;
                if (gi == g - 1) {   // use the wait for P * V to prepare the next item
                    if (threadIdx.x == 0) fetch(it + 2);
                    next = next_item(it + 1);
                    if (next >= 0) put_q(next, (it + 1) & 1);
                }
                consumer_sync<128>();
                if (GM > 1 && threadIdx.x == 0) inv_s[gi] = 1.0f / ((red[4] + red[5]) + (red[6] + red[7]));
            }
            if (GM > 1) consumer_sync<128>();   // inv_s is visible to warp 0
            if (warp == 0) {
                mbar_wait(o_full, (uint32_t)(it & 1));
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(tmem_O, o0);
                tmem_ld_32x32(tmem_O + 32, o1);
                tmem_ld_wait();
                if (lane < g) {   // TMEM lane = group row
                    const float inv = GM > 1 ? inv_s[lane] : 1.0f / ((red[4] + red[5]) + (red[6] + red[7]));
                    bf16* dst = out + (size_t)(r + lane) * ldo + h * 64;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 u;
                        __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(o0[i]) * inv, __uint_as_float(o0[i + 1]) * inv);
                        __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv);
                        __nv_bfloat162 c2 = __floats2bfloat162_rn(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv);
                        __nv_bfloat162 d2 = __floats2bfloat162_rn(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv);
                        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
                        u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d2);
                        *reinterpret_cast<uint4*>(dst + i) = u;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 u;
                        __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(o1[i]) * inv, __uint_as_float(o1[i + 1]) * inv);
                        __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv);
                        __nv_bfloat162 c2 = __floats2bfloat162_rn(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv);
                        __nv_bfloat162 d2 = __floats2bfloat162_rn(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv);
                        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
                        u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d2);
                        *reinterpret_cast<uint4*>(dst + 32 + i) = u;
                    }
                }
                tc_fence_before();
            }
            // the next item's first barrier (after its s_full wait) orders warp 0's O read before anyone
            // arrives on the next p_full, i.e. before the next item's P*V can overwrite the accumulator
            packed = next;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
    trace_end(tr);
    // the last CTA to leave re-arms the counters for the next launch on this lane
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {
            sched[0] = 0;
            sched[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace

void trace_set_cross(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}

namespace {
int env_or(const char* name, int dflt);
template <int STAGES, int SP, int MINB, int GM>
bool launch_tc(const CUtensorMap& tm, const RowDesc* rows, int n_items, int n_head, const bf16* q, int ldq, bf16* out, int ldo, long long k_row0,
               long long v_row0, long long slot_rows, int n_keys, int* sched, const CrossQPartials& qp, const int* groups, int grid, cudaStream_t s) {
    constexpr int SMEM = tc_smem_bytes(STAGES, GM);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dec_cross_attention_tc_kernel<STAGES, SP, MINB, GM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            sm100_set_error("cudaFuncSetAttribute(cross attention tc smem) failed");
            return false;
        }
        configured = true;
    }
    launch_kernel(dec_cross_attention_tc_kernel<STAGES, SP, MINB, GM>, dim3(grid), dim3(192), (size_t)SMEM, s, true, tm, rows, n_items, n_head, q, ldq, out, ldo,
                  k_row0, v_row0, slot_rows, n_keys, sched, qp, groups);
    return true;
}
}  // namespace

int cross_attention_groups(const RowDesc* rows, int n_rows, int* groups) {
    // runs of consecutive rows with one audio slot, cut into pieces of at most 4; returns the kernel's group width (1: no row shares a slot
    // with its neighbour, `groups` is not needed), and leaves the number of groups in groups[n_rows] when the width is 2 or 4
    int n = 0, width = 1;
    for (int r = 0; r < n_rows;) {
        int g = 1;
        while (g < 4 && r + g < n_rows && rows[r + g].audio_slot == rows[r].audio_slot) ++g;
        groups[n++] = r | (g << 24);
        width = g > width ? g : width;
        r += g;
    }
    groups[n_rows] = n;
    return width <= 2 ? width : 4;
}

bool launch_dec_cross_attention_tc_sm100(const RowDesc* rows, int n_rows, const bf16* q, int ldq, const bf16* pool, size_t pool_elems, size_t k_off,
                                         size_t v_off, bf16* out, int ldo, int n_head, size_t slot_stride, int n_keys, int* sched, int max_ctas, cudaStream_t s,
                                         const CrossQPartials* qpart, const CrossGroups* grp) {
    if (n_rows <= 0) return true;
    const CrossQPartials qp = qpart ? *qpart : CrossQPartials{};
    if ((!q && !qp.partial) || (qp.partial && ((qp.ld & 1) || (qp.plane & 1)))) { sm100_set_error("cross attention (tc): no query"); return false; }
    if (!sched || n_keys <= 0 || n_keys > CA_MAX_KEYS || (ldq % 8) != 0 || (slot_stride % 64) || (k_off % 64) || (v_off % 64) || pool_elems / 64 >= (1ull << 31)) {
        sm100_set_error("cross attention (tc): unsupported shape");
        return false;
    }
    if (grp && grp->width > 1 && (!grp->groups || grp->n_groups <= 0 || grp->n_groups > n_rows || (grp->width != 2 && grp->width != 4))) {
        sm100_set_error("cross attention (tc): bad row groups");
        return false;
    }
    static int sms = 0, stages = 0, sp = 0, per_sm = 1, low_regs = 1;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        low_regs = env_or("NOBS_WHISPER_CROSS_LOW_REGS", 1);
        stages = env_or("NOBS_WHISPER_CROSS_STAGES", 3);
        sp = env_or("NOBS_WHISPER_CROSS_SPACING", 4);
        per_sm = env_or("NOBS_WHISPER_CROSS_PER_SM", 2);
    }
    // one tensor map over the whole pool viewed as [rows][64] (re-encoded when the pool moves or grows)
    thread_local const bf16* cached_pool = nullptr;
    thread_local size_t cached_elems = 0;
    thread_local CUtensorMap tm;
    if (cached_pool != pool || cached_elems != pool_elems) {
        if (!make_tmap_bf16_2d(&tm, pool, 64, pool_elems / 64, 64, 64, CA_CHUNK_KEYS)) return false;
        cached_pool = pool;
        cached_elems = pool_elems;
    }
    // grouped rows: only with the 4-column score pitch and the 80-register build (the default configuration); anything else streams row by row
    const int gm = (grp && grp->width > 1 && sp == 4 && low_regs && (stages == 2 || stages == 3 || stages == 4)) ? grp->width : 1;
    const int* groups = gm > 1 ? grp->groups : nullptr;
    const int n_items = (gm > 1 ? grp->n_groups : n_rows) * n_head;
    int grid = n_items < sms * per_sm ? n_items : sms * per_sm;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    // Balanced waves: 1200 items on 296 CTAs are four full waves plus 16 stragglers, and the launch ends one item time (~10 us of a
    // nearly idle HBM) after most CTAs have run dry.  The stream is bandwidth-bound, so a slightly smaller grid in which every CTA
    // takes the same number of items (1200 = 240 x 5) finishes with the bandwidth-bound time and no tail.
    static const int balance = env_or("NOBS_WHISPER_CROSS_BALANCE", 1);
    if (balance && n_items > grid) {
        const int waves = (n_items + grid - 1) / grid;
        grid = (n_items + waves - 1) / waves;
    }
    bool ok = false;
#define TC_ARGS tm, rows, n_items, n_head, q, ldq, out, ldo, (long long)(k_off / 64), (long long)(v_off / 64), (long long)(slot_stride / 64), n_keys, sched, qp, groups, grid, s
#define TC_CASE(S, P) if (stages == S && sp == P) ok = low_regs ? launch_tc<S, P, 4, 1>(TC_ARGS) : launch_tc<S, P, 1, 1>(TC_ARGS); else
#define TC_GROUPED(S) if (gm > 1 && stages == S) ok = gm == 2 ? launch_tc<S, 4, 4, 2>(TC_ARGS) : launch_tc<S, 4, 4, 4>(TC_ARGS); else
    TC_GROUPED(2) TC_GROUPED(3) TC_GROUPED(4)
    TC_CASE(4, 16) TC_CASE(6, 16) TC_CASE(8, 16) TC_CASE(10, 16) TC_CASE(2, 4) TC_CASE(3, 4) TC_CASE(4, 4) TC_CASE(5, 4) TC_CASE(6, 4) TC_CASE(8, 4)
    { sm100_set_error("cross attention (tc): unsupported NOBS_WHISPER_CROSS_STAGES / _SPACING"); return false; }
#undef TC_GROUPED
#undef TC_CASE
#undef TC_ARGS
    if (!ok) return false;
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { sm100_set_error(std::string("cross attention (tc) launch: ") + cudaGetErrorString(err)); return false; }
    return true;
}

namespace {
int env_or(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
template <int STAGES, int WARPS>
bool launch_ca(const RowDesc* rows, int n_items, int n_head, const bf16* q, int ldq, const bf16* kbase, const bf16* vbase, bf16* out, int ldo,
               size_t slot_stride, size_t head_stride, int n_keys, int grid, cudaStream_t s) {
    constexpr int SMEM = ca_smem_bytes(STAGES, WARPS);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dec_cross_attention_sm100_kernel<STAGES, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            sm100_set_error("cudaFuncSetAttribute(cross attention smem) failed");
            return false;
        }
        configured = true;
    }
    launch_kernel(dec_cross_attention_sm100_kernel<STAGES, WARPS>, dim3(grid), dim3((WARPS + 1) * 32), (size_t)SMEM, s, true, rows, n_items, n_head, q, ldq,
                  kbase, vbase, out, ldo, slot_stride, head_stride, n_keys);
    return true;
}
}  // namespace

bool launch_dec_cross_attention_sm100(const RowDesc* rows, int n_rows, const bf16* q, int ldq, const bf16* kbase, const bf16* vbase, bf16* out, int ldo,
                                      int n_head, size_t slot_stride, size_t head_stride, int n_keys, int max_ctas, cudaStream_t s) {
    if (n_rows <= 0) return true;
    if (n_keys <= 0 || n_keys > CA_MAX_KEYS || (ldq % 8) != 0) { sm100_set_error("cross attention: unsupported shape"); return false; }
    static int sms = 0, stages = 0, warps = 0, per_sm = 1;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        stages = env_or("NOBS_WHISPER_CROSS_SIMT_STAGES", 3);
        warps = env_or("NOBS_WHISPER_CROSS_SIMT_WARPS", 8);
        per_sm = env_or("NOBS_WHISPER_CROSS_SIMT_PER_SM", 2);
    }
    const int n_items = n_rows * n_head;
    int grid = n_items < sms * per_sm ? n_items : sms * per_sm;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    bool ok = false;
#define CA_CASE(S, W) if (stages == S && warps == W) ok = launch_ca<S, W>(rows, n_items, n_head, q, ldq, kbase, vbase, out, ldo, slot_stride, head_stride, n_keys, grid, s); else
    CA_CASE(3, 8) CA_CASE(4, 8) CA_CASE(6, 8) CA_CASE(8, 8) CA_CASE(10, 8) CA_CASE(12, 8) CA_CASE(4, 16) CA_CASE(6, 16) CA_CASE(8, 16) CA_CASE(12, 16) CA_CASE(3, 4) CA_CASE(4, 4)
    { sm100_set_error("cross attention: unsupported NOBS_WHISPER_CROSS_STAGES / _WARPS"); return false; }
#undef CA_CASE
    if (!ok) return false;
    count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { sm100_set_error(std::string("cross attention launch: ") + cudaGetErrorString(err)); return false; }
    return true;
}

}  // namespace nobs
