// GPU executor of the hot path: owns the weights in HBM, the KV / cross-KV slot pools, the
// activation workspaces and one stream.  Host control flow (full.cpp) drives it with batches
// of windows (encode) and batches of token rows (decode).
//
// HBM layout (T = float in fp32 parity mode, bf16 in production mode):
//   weights        one arena; per layer Wqkv [3d,d] (q|k|v fused, K has no bias), Wo [d,d],
//                  W1 [4d,d], W2 [d,4d]; conv weights re-ordered to [d][(tap, channel)];
//                  all decoder layers' cross K/V projections fused into one [2*L*d, d] matrix.
//   cross-KV pool  [audio_slot][L][K|V][head][1536 positions][64]  (written by ONE GEMM per encode batch
//                  through a head-major epilogue: each head's keys are one contiguous stream)
//   self-KV pool   [kv_slot][L][K|V][head][448 positions][64]
//   encoder out    [audio_slot][1536][d]
//   activations    encoder: rows = windows x 1536 (1500 valid), decoder: rows = token rows.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "model.h"

namespace nobs {

enum class Precision { FP32 = 1, BF16 = 2 };

// Raw log-mel of one audio, resident on the device (owned by a whisper_state).
struct DeviceMel {
    float* raw = nullptr;   // [n_frames][n_mel]
    int* max_key = nullptr; // running max key
    size_t raw_cap = 0;     // floats allocated
    int n_frames = 0;       // frames computed (overlapping samples)
    int n_len = 0;          // (n_samples + 480000) / 160
    int n_len_org = 0;      // 1 + (n_samples + 200 - 400) / 160
    int n_mel = 0;
};

struct MelRequest {
    const float* pcm;  // host
    int n_samples;
    DeviceMel* mel;
};
struct EncodeRequest {
    const DeviceMel* mel;
    int seek;
    int audio_slot;
};

struct EngineStats {
    double ms_mel = 0, ms_encode = 0, ms_decode = 0;
    // per-launch CUDA-event timing of the encoder's kernel classes (only when profiling is on)
    double ms_enc_gemm = 0, ms_enc_attn = 0, ms_dec_cross = 0;
    long n_enc_gemm = 0, n_enc_attn = 0, n_dec_cross = 0;
    double dec_cross_bytes = 0;  // algorithmic K/V panel bytes read by the timed cross-attention launches
    long n_launches = 0;
};

class Engine {
public:
    virtual ~Engine() {}
    static Engine* create(const HostModel& hm, int device, Precision prec, std::string& err);

    std::mutex mu;  // one `full` batch at a time per context
    bool profiling = false;  // record an event pair around every encoder GEMM / attention launch
    Precision precision() const { return prec_; }
    int device() const { return device_; }
    const std::string& last_error() const { return err_; }
    EngineStats stats;

    // slot pools (grow on demand, contents preserved)
    virtual int acquire_audio_slot() = 0;
    virtual void release_audio_slot(int s) = 0;
    virtual int acquire_kv_slot() = 0;
    virtual void release_kv_slot(int s) = 0;
    // make sure that many free slots exist (one pool growth instead of repeated doubling)
    virtual bool reserve_slots(int n_audio_free, int n_kv_free) = 0;
    virtual void free_mel(DeviceMel& m) = 0;

    virtual bool compute_mel(const std::vector<MelRequest>& reqs) = 0;
    // The encoder has its own stream.  encode_async queues the windows and returns a ticket, encode_wait blocks until the ticket's
    // encoder output and cross-KV panels are complete (tickets complete in the order they were handed out).
    virtual bool encode_async(const std::vector<EncodeRequest>& reqs, long* ticket) = 0;
    virtual bool encode_wait(long ticket) = 0;
    bool encode(const std::vector<EncodeRequest>& reqs) {
        long ticket = 0;
        const bool queued = encode_async(reqs, &ticket);
        return encode_wait(ticket) && queued;
    }
    // Decode lanes: independent (stream, workspace) pairs.  Jobs are partitioned over lanes by the host driver;
    // while one lane streams its cross-KV panels (HBM-bound) the projection chain of another lane
    // (latency-bound) runs on the same SMs, and the host prepares one lane's next round while the others compute.
    virtual int n_lanes() const = 0;
    // rows: token rows in order; sample_rows[i] indexes `rows`; one SampleParams per sample row.
    // logits_host (optional): receives [n_sample][n_vocab] raw logits.  Returns once the work is queued.
    // inject / inject_mask (optional, test hook): inject[i * n_vocab ..] replaces the logits of sample i on the device where inject_mask[i] != 0.
    virtual bool decode_submit(int lane, const std::vector<RowDesc>& rows, const std::vector<int>& sample_rows, const std::vector<SampleParams>& sp,
                               float* logits_host, const float* inject = nullptr, const unsigned char* inject_mask = nullptr) = 0;
    // waits for the lane's round; one SampleResult per sample row
    virtual bool decode_collect(int lane, std::vector<SampleResult>& results) = 0;
    bool decode(const std::vector<RowDesc>& rows, const std::vector<int>& sample_rows, const std::vector<SampleParams>& sp,
                std::vector<SampleResult>& results, float* logits_host) {
        return decode_submit(0, rows, sample_rows, sp, logits_host) && decode_collect(0, results);
    }
    // language probabilities from the raw logits of sample row `i` of the lane's last round
    virtual bool lang_probs(int lane, int sample_index, float* probs_host /*[100] or null*/, int* best) = 0;
    virtual bool kv_copy(int lane, const std::vector<KvCopy>& pairs) = 0;
    // stage-parity hook: run K6 on host-supplied logits
    virtual bool process_logits_host(const float* logits, const SampleParams& sp, SampleResult& out, float* logprobs, float* probs) = 0;

    virtual bool event_record(int slot) = 0;                 // on the engine's stream
    virtual double event_elapsed_ms(int a, int b) = 0;       // synchronises on event b

    virtual bool export_mel(const DeviceMel& m, float* out) = 0;                       // [n_mel][n_len]
    virtual bool export_encoder_output(int audio_slot, float* out) = 0;                 // [1500][d]
    virtual bool export_cross_kv(int audio_slot, int layer, float* k, float* v) = 0;    // [1500][d] each
    VocabIds vocab_ids{};

protected:
    Precision prec_ = Precision::FP32;
    int device_ = 0;
    std::string err_;
};

}  // namespace nobs
