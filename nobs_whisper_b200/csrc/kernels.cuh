// Device kernels of the B200 Whisper hot path and their launchers.
// Stage map: K1 log-mel (a4), K2 conv stem + K3 encoder layers (a7),
// K4 cross-KV (a8), K5 decoder rows (a9), K6 logit filter / log-softmax / sampling (a10).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace nobs {

using bf16 = __nv_bfloat16;

constexpr int kWinRowsIn = 3072;   // padded mel frames per window (3000 valid)
constexpr int kWinRows = 1536;     // padded encoder positions per window (1500 valid)
constexpr int kNFft = 400;
constexpr int kHop = 160;
constexpr int kNFreq = 201;
constexpr int kMaxTopK = 8;

// ---------------------------------------------------------------- K1 log-mel
struct MelTables {           // device pointers, built once per context
    const float* hann;       // [400]
    const float2* tw;        // [400] (cos, sin)(2*pi*i/400)
    const float* filters;    // [n_mel][201]
    const int2* ranges;      // [n_mel] (first non-zero k, one past last)
    int n_mel;
};
struct MelJob {              // one audio
    const float* pcm;        // device, n_samples floats
    int n_samples;
    int n_frames;            // frames that overlap samples; later frames are exactly log10(1e-10)
    float* raw;              // [n_frames][n_mel] log10 mel, time-major
    int* max_key;            // ordered-int key of the running max (init INT_MIN)
};
void launch_mel_stft(const MelTables& t, const MelJob* jobs_dev, int n_jobs, int max_frames, cudaStream_t s);

struct PackJob {             // one window to encode
    const float* raw;        // job raw mel
    const int* max_key;
    int n_frames;            // computed frames
    int n_len;               // total frames of the audio (mel.n_len)
    int seek;                // first frame of the window
};
// dst: [(n_win * kWinRowsIn + 2)][n_mel]; writes rows 1 + w*kWinRowsIn + t
template <typename T>
void launch_pack_mel(const PackJob* jobs_dev, int n_win, int n_mel, T* dst, cudaStream_t s);
// normalised mel in upstream layout [n_mel][n_len] (inspection / parity hook)
void launch_export_mel(const float* raw, const int* max_key, int n_frames, int n_len, int n_mel, float* out, cudaStream_t s);

// ---------------------------------------------------------------- GEMM epilogue description
struct Epilogue {
    const float* bias = nullptr;  // [N]
    int act = 0;                  // 0 none, 1 GELU (tanh form, as ggml)
    const float* res = nullptr;   // fp32 residual, row stride res_ld; row index = m % res_mod (res_mod 0: m)
    int res_ld = 0;
    int res_mod = 0;
    int win_rows = 0;             // if > 0: rows with (m % win_rows) >= valid_rows are written as 0
    int valid_rows = 0;
    // if > 0: head-major output.  Row m = (w, pos) with pos = m % head_rows; column n = 64*blk + e.
    // Element goes to C[w*head_rows*ldc + (blk*head_rows + pos)*64 + e]: every 64-wide column block
    // (one attention head of K or V) becomes a contiguous [head_rows][64] panel (KV-cache layout).
    int head_rows = 0;
};

// C[M,N] = epi(A[M,K] * W[N,K]^T), all fp32, K-contiguous operands (fp32 parity mode)
void launch_gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K, const Epilogue& e,
                     cudaStream_t s);

// ---------------------------------------------------------------- LayerNorm
template <typename TO>
void launch_layernorm(const float* x, int ldx, const float* g, const float* b, TO* y, int ldy, int rows, int d, cudaStream_t s);
// gathered rows: y[i] = LN(x[idx[i]])
template <typename TO>
void launch_layernorm_gather(const float* x, int ldx, const int* idx, const float* g, const float* b, TO* y, int ldy, int rows, int d,
                             cudaStream_t s);

// ---------------------------------------------------------------- encoder attention (non-causal, 1500 keys)
// qkv: [n_win*kWinRows][3*d] (q | k | v), out: [n_win*kWinRows][d]
void launch_enc_attention_f32(const float* qkv, float* out, int n_win, int n_head, int d, cudaStream_t s);
// CUDA-core flash-style kernel (fp32 math), storage type T
template <typename T>
void launch_enc_attention_simt(const T* qkv, T* out, int n_win, int n_head, int d, cudaStream_t s);

// ---------------------------------------------------------------- decoder rows
struct RowDesc {
    int token, pos;
    int kv_slot;     // self-KV sequence slot
    int audio_slot;  // cross-KV slot
};
template <typename T>
void launch_embed(const RowDesc* rows, int n_rows, const T* tok_emb, const float* pos_emb, float* x, int d, cudaStream_t s);
// qkv [n_rows][3d] (q|k|v) -> head-major K/V panels of one layer:
// dst = panel0 + kv_slot*slot_stride + (head*n_pos_cap + pos)*64 + e
template <typename T>
void launch_scatter_kv(const RowDesc* rows, int n_rows, const T* qkv, T* kpanel0, T* vpanel0, size_t slot_stride, int n_pos_cap, int d,
                       cudaStream_t s);
// Optional input of the self-attention launch: q / k / v of the step rows as the split-K partial sums of the QKV
// projection (partial[split][row][ld = 3d], + bias [3d]); the kernel finishes them and appends k / v to the cache.
struct QkvPartials {
    const float* partial = nullptr;
    int splits = 0;
    size_t plane = 0;   // floats between consecutive splits (rows * ld)
    int ld = 0;
    const float* bias = nullptr;
};
// one (row, head) per block; key j of head h lives at base + slot*slot_stride + h*head_stride + j*64.
// cross == 0: slot = kv_slot, keys 0..pos (causal);  cross == 1: slot = audio_slot, keys 0..n_keys-1.
template <typename T>
void launch_dec_attention(const RowDesc* rows, int n_rows, const T* q, int ldq, const T* kbase, const T* vbase, T* out, int ldo, int n_head,
                          int cross, size_t slot_stride, size_t head_stride, int n_keys, cudaStream_t s, const QkvPartials* qkv_partials = nullptr);
// self-KV slot copy for beam search: the first n_pos rows of each of the n_panels [n_pos_cap][64] panels
struct KvCopy {
    int src, dst, n_pos;
};
template <typename T>
void launch_kv_copy(const KvCopy* pairs, int n_pairs, T* pool, size_t slot_stride, int n_panels, size_t panel_stride, cudaStream_t s);

// ---------------------------------------------------------------- skinny-GEMM epilogue (decoder steps)
// Sums the split-K partials of gemm_skinny (fixed order: deterministic) for one token row per block and
// applies everything that used to be separate launches: bias, GELU, residual add into the fp32 stream,
// the next LayerNorm, and the scatter of fresh K/V rows into the head-major self-KV panels.
struct SkinnyEpilogue {
    const float* partial = nullptr;  // [splits][R][N]
    int splits = 0, R = 0, N = 0;
    const float* bias = nullptr;     // [N]
    int act = 0;                     // 1: GELU
    float* x = nullptr;              // residual stream [R][N] fp32: x += v (v becomes the updated x)
    void* out = nullptr;             // T [R][out_ld]: out = v
    int out_ld = 0;
    const float* ln_g = nullptr;     // if set: y = LayerNorm(v) * g + b as T [R][N]
    const float* ln_b = nullptr;
    void* y = nullptr;
    const RowDesc* rows = nullptr;   // if set (QKV): columns [d,2d) / [2d,3d) also go to the K / V panels
    void* kpanel = nullptr;
    void* vpanel = nullptr;
    size_t slot_stride = 0;
    int n_pos_cap = 0, d = 0;
};
template <typename T>
void launch_skinny_reduce(const SkinnyEpilogue& e, cudaStream_t s);

// ---------------------------------------------------------------- K6 logits -> token
struct VocabIds {
    int n_vocab, eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, blank, n_lang;
};
struct SampleParams {
    float temperature;
    int is_initial, last_was_ts, penult_was_ts, has_ts;
    int suppress_blank, no_timestamps;
    int ts_initial_limit;  // first suppressed timestamp id on the initial step (n_vocab: none)
    int ts_min;            // timestamps below this id are suppressed when has_ts
    int mode;              // 0 argmax, 1 sample with u, 2 top-k
    int k;
    int want_nosp;         // compute softmax(raw logits)[nosp]
    double u;
};
struct SampleResult {
    int id, tid;
    float p, plog, pt, ptsum, no_speech_prob;
    int n_topk;
    int topk_id[kMaxTopK];
    float topk_plog[kMaxTopK];
    float topk_p[kMaxTopK];
};
// logits [n_rows][ld]; optional dumps of the filtered logprobs / probs (row-major [n_rows][n_vocab])
void launch_process_logits(const float* logits, int ld, const SampleParams* params, SampleResult* results, int n_rows, const VocabIds& v,
                           float* logprobs_out, float* probs_out, cudaStream_t s);
// language detection: softmax over the language-token logits of one row; probs_out [n_lang]
void launch_lang_probs(const float* logits, const VocabIds& v, float* probs_out, int* best, cudaStream_t s);

// ---------------------------------------------------------------- misc
template <typename TI, typename TO>
void launch_convert(const TI* in, TO* out, size_t n, cudaStream_t s);
// strided 2D copy with conversion: out[r][c] = in[r*ld_in + c]
template <typename TI, typename TO>
void launch_convert_2d(const TI* in, size_t ld_in, TO* out, size_t ld_out, int rows, int cols, cudaStream_t s);

// Launch with (or without) the programmatic-dependent-launch attribute.
extern bool g_use_pdl;
// Opt-in (NOBS_WHISPER_MAX_CARVEOUT=1): ask for the maximum shared-memory carveout for every kernel, so that SMs
// never have to drain to change their L1 / shared-memory split when kernels of different decode lanes share them.
void prefer_max_shared_carveout(const void* kernel);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_use_pdl) ? 1 : 0;
    prefer_max_shared_carveout(reinterpret_cast<const void*>(kernel));
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernel timeline buffer (device_utils.cuh); one setter per translation unit
void trace_set_kernels(unsigned long long* buf, unsigned int cap);
void trace_set_gemm(unsigned long long* buf, unsigned int cap);
void trace_set_cross(unsigned long long* buf, unsigned int cap);
long kernel_launch_count();  // process-wide count of kernels launched through these launchers
void count_launch();
void count_launches(long n);   // kernels replayed through a CUDA graph

}  // namespace nobs
