// tcgen05 / TMEM / TMA kernels for the bf16 production path (sm_100a only).
#pragma once
#include "kernels.cuh"

namespace nobs {

// C[M,N] = epi(A[M,K] * W[N,K]^T): bf16 operands (K-contiguous), fp32 accumulation in TMEM,
// C written as bf16 or fp32.  Returns false (see sm100_last_error) if a tensor map cannot be
// encoded or the launch fails.
bool launch_gemm_bf16_sm100(const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, bool c_is_f32, int M, int N, int K,
                            const Epilogue& e, cudaStream_t s);
// Decoder-step GEMM (R <= 128 token rows), swap-AB + split-K on tcgen05: writes fp32 partial sums
// partial[split][R][N] (the workspace must hold skinny_gemm_splits(N, K) * R * N floats); the caller
// finishes with launch_skinny_reduce (kernels.cuh).
int skinny_gemm_splits(int N, int K);
// shared-memory ring depth of the decoder-step GEMM: 3 (default) or 2 (engines with several decode lanes)
void set_skinny_gemm_stages(int stages);
bool launch_gemm_skinny_bf16_sm100(const bf16* X, int ldx, const bf16* W, int ldw, float* partial, int R, int N, int K, int* splits_out,
                                   cudaStream_t s);
// Non-causal encoder self-attention over 1500 keys per window, head size 64.
bool launch_enc_attention_bf16_sm100(const bf16* qkv, bf16* out, int n_win, int n_head, int d, cudaStream_t s);
// Decoder cross-attention over the head-major cross-KV panels for single-token rows (bf16): persistent,
// one small CTA per SM fed by a cp.async.bulk ring (cross_attention_sm100.cu).  max_ctas > 0 caps the grid.
bool launch_dec_cross_attention_sm100(const RowDesc* rows, int n_rows, const bf16* q, int ldq, const bf16* kbase, const bf16* vbase, bf16* out, int ldo,
                                      int n_head, size_t slot_stride, size_t head_stride, int n_keys, int max_ctas, cudaStream_t s);
// Tensor-core variant: panels go TMA -> 128B-swizzled shared memory -> tcgen05.mma (scores and P*V), the SM's
// issue slots stay free for co-resident kernels.  `pool` is the whole cross-KV pool (pool_elems bf16), k_off / v_off
// the element offsets of this layer's K / V panels inside an audio slot, head panels 1536 x 64 apart.  `sched`: two
// zero-initialised device ints owned by the calling stream (work counter + exit counter; the kernel re-arms them).
// Back-to-back launches on one stream must alternate between two such pairs (a launch's prologue may overlap its predecessor).
// Optional: the queries as the split-K partial sums of the query projection (partial[split][row][ld], + bias), summed by
// the attention kernel itself — saves the projection's epilogue launch on the decoder's latency chain.
struct CrossQPartials {
    const float* partial = nullptr;
    int splits = 0;
    size_t plane = 0;   // floats between consecutive splits (rows * ld)
    int ld = 0;
    const float* bias = nullptr;
};
// Optional: row groups.  Consecutive rows of ONE audio slot (a pass and its speculative successor, the beams of a beam search) attend over
// the same panels; as a group of up to 4 they are one work item whose panels are streamed once.  cross_attention_groups() (host) cuts the
// rows into such groups: groups[i] = first row | size << 24, groups[n_rows] = number of groups (the array holds n_rows + 1 ints); it returns
// the kernel's group width (1 = no neighbours share a slot: pass no groups).  The device copy of the table is what the launch takes.
struct CrossGroups {
    const int* groups = nullptr;   // device
    int n_groups = 0;
    int width = 1;                 // 1, 2 or 4
};
int cross_attention_groups(const RowDesc* rows, int n_rows, int* groups);
bool launch_dec_cross_attention_tc_sm100(const RowDesc* rows, int n_rows, const bf16* q, int ldq, const bf16* pool, size_t pool_elems, size_t k_off,
                                         size_t v_off, bf16* out, int ldo, int n_head, size_t slot_stride, int n_keys, int* sched, int max_ctas, cudaStream_t s,
                                         const CrossQPartials* qpart = nullptr, const CrossGroups* grp = nullptr);
// Decoder projection chain (decode_chain_sm100.cu): up to kChainMaxSteps dependent swap-AB split-K projections of one step
// batch (R <= 128 rows) in ONE persistent launch, separated by device-wide barriers instead of kernel boundaries.  Every
// step but the last must be reduced (`reduce` = 1: SkinnyEpilogue semantics, identical arithmetic to skinny_reduce_kernel);
// the last may leave its partial sums in `partial` for the attention kernel that follows.
constexpr int kChainMaxSteps = 4;
struct ChainStep {
    int N = 0, K = 0;                                // out[r][n] = sum_k X[r][k] * W[n][k]
    int m_tiles = 0, splits = 0, kb_per_split = 0;   // filled by the launcher (same tiling as launch_gemm_skinny_bf16_sm100)
    int reduce = 0;
    SkinnyEpilogue e;                                // partial / splits / R / N are filled by the launcher
};
struct ChainDesc {
    int n_steps = 0, R = 0;
    ChainStep step[kChainMaxSteps];
    float* partial = nullptr;        // split-K workspace of the lane
    unsigned int* bar = nullptr;     // 2 zero-initialised words owned by the lane (device-wide barrier state)
};
// W[i]: [N_i][K_i] weights, X[i]: [R][ldx[i]] activations of step i.  stages: shared-memory ring depth (2 or 3).
bool launch_dec_chain_sm100(ChainDesc& cd, const bf16* const* W, const bf16* const* X, const int* ldx, int stages, cudaStream_t s);
// Fused decoder-step projection (decode_proj_sm100.cu): v = X * W^T + bias for R <= 128 token rows in ONE launch — split-K over a
// thread-block cluster with the partial sums reduced through distributed shared memory; bias / GELU / residual / bf16 store /
// KV-cache append in the same kernel; and, for a projection that updates the residual stream, y = LayerNorm(x) for the next
// projection written by whichever cluster finishes last.  N must be a multiple of 128 (at most 40 tiles), K of 64.
struct ProjDesc {
    int R = 0, N = 0, K = 0;
    const bf16* W = nullptr;            // [N][K]
    const bf16* X = nullptr;            // bf16 activations [R][ldx]
    int ldx = 0;
    const float* bias = nullptr;        // [N]
    int act = 0;                        // 1: GELU
    float* x = nullptr;                 // residual stream [R][N]: x += v (v becomes the updated x)
    bf16* out = nullptr;                // [R][out_ld] = v
    int out_ld = 0;
    const RowDesc* rows = nullptr;      // if set (QKV): columns [d, 2d) / [2d, 3d) are also appended to the K / V cache panels
    bf16* kpanel = nullptr;
    bf16* vpanel = nullptr;
    size_t slot_stride = 0;
    int n_pos_cap = 0, d = 0;
    // LayerNorm of the updated residual stream (needs x): y [R][N] = LN(x) * ln_g + ln_b
    bf16* y = nullptr;
    const float* ln_g = nullptr;
    const float* ln_b = nullptr;
    float2* stats_out = nullptr;        // workspace [N / 128][128]: (mean, M2) per 128-feature tile and row
    int* ticket = nullptr;              // one zero-initialised int owned by the calling stream (re-armed by the kernel)
};
bool dec_proj_supported(int R, int N, int K);
int dec_proj_splits(int N, int K);
bool launch_dec_proj_sm100(const ProjDesc& a, cudaStream_t s);
void trace_set_proj(unsigned long long* buf, unsigned int cap);
int chain_stages_for_lanes(int n_lanes);
// true when n_lanes concurrent chain grids can always become co-resident (no mutual wait for SM slots)
bool chain_fits(int n_lanes, int stages);
void trace_set_chain(unsigned long long* buf, unsigned int cap);
const char* sm100_last_error();

}  // namespace nobs
