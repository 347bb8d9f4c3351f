#include "engine.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>
#include <unordered_map>

#include "gemm_sm100.cuh"

extern char** environ;

namespace nobs {

// A profiler / sanitizer injects a library into the process and announces itself through the environment.
static bool cuda_tool_attached() {
    for (char** e = environ; e && *e; ++e) {
        if (!strncmp(*e, "CUDA_INJECTION64_PATH=", 22) || !strncmp(*e, "NV_NSIGHT_INJECTION", 19) || !strncmp(*e, "NV_COMPUTE_PROFILER", 19) ||
            !strncmp(*e, "NV_TPS_LAUNCH", 13) || !strncmp(*e, "NSYS_PROFILING_SESSION_ID=", 26) || !strncmp(*e, "NV_SANITIZER", 12))
            return true;
    }
    if (FILE* f = fopen("/proc/self/maps", "r")) {
        char line[512];
        bool found = false;
        while (!found && fgets(line, sizeof(line), f))
            found = strstr(line, "libcuda-injection") || strstr(line, "libToolsInjection") || strstr(line, "libInterceptorInjectionTarget") ||
                    strstr(line, "libsanitizer-collection") || strstr(line, "libTreeLauncherTargetInjection");
        fclose(f);
        if (found) return true;
    }
    return false;
}

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_err(std::string(#expr) + " failed: " + cudaGetErrorString(e_));                \
            return false;                                                                      \
        }                                                                                      \
    } while (0)

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoi(s);
}

// GEMM dispatch by storage type.  fp32 parity mode runs on CUDA cores; bf16 runs the
// tcgen05/TMA kernel (gemm_sm100.cu).
inline bool gemm(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K, const Epilogue& e, cudaStream_t s) {
    launch_gemm_f32(A, lda, W, ldw, C, ldc, M, N, K, e, s);
    return true;
}
inline bool gemm(const bf16* A, int lda, const bf16* W, int ldw, bf16* C, int ldc, int M, int N, int K, const Epilogue& e, cudaStream_t s) {
    return launch_gemm_bf16_sm100(A, lda, W, ldw, C, ldc, /*c_is_f32=*/false, M, N, K, e, s);
}
inline bool gemm(const bf16* A, int lda, const bf16* W, int ldw, float* C, int ldc, int M, int N, int K, const Epilogue& e, cudaStream_t s) {
    return launch_gemm_bf16_sm100(A, lda, W, ldw, C, ldc, /*c_is_f32=*/true, M, N, K, e, s);
}
inline bool enc_attention(const float* qkv, float* out, int n_win, int n_head, int d, cudaStream_t s) {
    launch_enc_attention_f32(qkv, out, n_win, n_head, d, s);
    return true;
}
inline bool enc_attention(const bf16* qkv, bf16* out, int n_win, int n_head, int d, cudaStream_t s) {
    return launch_enc_attention_bf16_sm100(qkv, out, n_win, n_head, d, s);
}

template <typename T>
struct Layer {
    float *ln1_g, *ln1_b;
    T* wqkv; float* bqkv;
    T* wo; float* bo;
    float *lnc_g, *lnc_b;  // decoder only
    T* wcq; float* bcq;
    T* wco; float* bco;
    float *ln2_g, *ln2_b;
    T* w1; float* b1;
    T* w2; float* b2;
};

template <typename T>
class EngineT final : public Engine {
public:
    EngineT(const HostModel& hm, int device, Precision prec) : hp_(hm.hp) {
        device_ = device;
        prec_ = prec;
    }
    ~EngineT() override {
        cudaSetDevice(device_);
        if (stream_) cudaStreamSynchronize(stream_);
        for (void* p : owned_) cudaFree(p);
        for (auto& m : mel_cache_) { cudaFree(m.raw); cudaFree(m.max_key); }
        if (pin_) cudaFreeHost(pin_);
        if (cross_pool_) cudaFree(cross_pool_);
        if (self_pool_) cudaFree(self_pool_);
        if (encout_pool_) cudaFree(encout_pool_);
        if (pcm_dev_) cudaFree(pcm_dev_);
        for (auto& e : ev_) if (e) cudaEventDestroy(e);
        for (auto& t : enc_ring_) { if (t.begin) cudaEventDestroy(t.begin); if (t.end) cudaEventDestroy(t.end); }
        for (auto& e : enc_pin_ev_) if (e) cudaEventDestroy(e);
        if (enc_pin_) cudaFreeHost(enc_pin_);
        if (env_int("NOBS_WHISPER_PROFILE_HOST", 0))
            fprintf(stderr, "[nobs profile] decode host: issuing launches %.1f ms, waiting for the GPU %.1f ms\n", host_issue_ms_, host_wait_ms_);
        if (detail_) {
            const char* names[8] = {"skinny_gemm", "skinny_reduce", "self_attn", "cross_attn", "layernorm", "logits+sample", "embed", "other"};
            for (int i = 0; i < 8; ++i)
                if (detail_n_[i]) fprintf(stderr, "[nobs profile] %-14s n=%8ld total=%10.2f ms avg=%8.2f us\n", names[i], detail_n_[i], detail_ms_[i], 1e3 * detail_ms_[i] / detail_n_[i]);
        }
        if (trace_buf_) {
            cudaDeviceSynchronize();
            unsigned long long n = 0;
            cudaMemcpy(&n, trace_buf_, 8, cudaMemcpyDeviceToHost);
            n = std::min<unsigned long long>(n, trace_cap_);
            const unsigned long long skip = std::min<unsigned long long>(n, (unsigned long long)std::max(0, env_int("NOBS_WHISPER_TRACE_SKIP", 0)));
            const unsigned long long cnt = std::min<unsigned long long>(n - skip, (unsigned long long)std::max(1, env_int("NOBS_WHISPER_TRACE_COUNT", 400000)));
            std::vector<unsigned long long> h(4 * cnt);
            cudaMemcpy(h.data(), trace_buf_ + 4 + 4 * skip, h.size() * 8, cudaMemcpyDeviceToHost);
            if (FILE* f = fopen(trace_path_.c_str(), "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
            fprintf(stderr, "[nobs trace] %llu entries recorded, wrote %llu from %llu to %s\n", n, cnt, skip, trace_path_.c_str());
            trace_set_kernels(nullptr, 0); trace_set_gemm(nullptr, 0); trace_set_cross(nullptr, 0); trace_set_chain(nullptr, 0); trace_set_proj(nullptr, 0);
            cudaFree(trace_buf_);
        }
        for (auto& e : main_marks_.pool) cudaEventDestroy(e);
        for (auto& L : lanes_) {
            if (L.stream) cudaStreamSynchronize(L.stream);
            for (auto& e : L.tm.pool) cudaEventDestroy(e);
            for (auto& e : L.ev) if (e) cudaEventDestroy(e);
            for (auto& g : L.graphs) cudaGraphExecDestroy(g.second.exec);
            if (L.pin) cudaFreeHost(L.pin);
            if (L.scratch) cudaFree(L.scratch);
            if (L.stream) cudaStreamDestroy(L.stream);
        }
        for (auto& e : user_ev_) if (e) cudaEventDestroy(e);
        if (stream_) cudaStreamDestroy(stream_);
    }

    bool init(const HostModel& hm) {
        int n_dev = 0;
        CUDA_OK(cudaGetDeviceCount(&n_dev));
        if (device_ < 0 || device_ >= n_dev) { set_err("invalid CUDA device " + std::to_string(device_)); return false; }
        CUDA_OK(cudaSetDevice(device_));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device_));
        if (prop.major != 10) { set_err(std::string("device '") + prop.name + "' is not sm_100 (this library has no other code path)"); return false; }
        CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        for (auto& e : ev_) CUDA_OK(cudaEventCreate(&e));
        d_ = hp_.n_audio_state;
        const Vocab& v = hm.vocab;
        vocab_ids = VocabIds{hp_.n_vocab, v.token_eot, v.token_sot, v.token_translate, v.token_transcribe, v.token_solm, v.token_prev,
                             v.token_nosp, v.token_not, v.token_beg, v.token_blank, kNumLangs};
        if (!upload_weights(hm)) return false;
        if (!alloc_workspace()) return false;
        if (const char* tp = getenv("NOBS_WHISPER_TRACE")) {   // kernel timeline (debugging aid)
            trace_path_ = tp;
            trace_cap_ = (unsigned)std::max(1024, env_int("NOBS_WHISPER_TRACE_CAP", 3000000));
            CUDA_OK(cudaMalloc(&trace_buf_, (4 + 4 * (size_t)trace_cap_) * 8));
            CUDA_OK(cudaMemset(trace_buf_, 0, (4 + 4 * (size_t)trace_cap_) * 8));
            trace_set_kernels(trace_buf_, trace_cap_);
            trace_set_gemm(trace_buf_, trace_cap_);
            trace_set_cross(trace_buf_, trace_cap_);
            trace_set_chain(trace_buf_, trace_cap_);
            trace_set_proj(trace_buf_, trace_cap_);
        }
        CUDA_OK(cudaStreamSynchronize(stream_));
        return true;
    }

    // ------------------------------------------------------------------ slots
    int acquire_audio_slot() override {
        if (audio_free_.empty() && !grow_audio(std::max(1, audio_cap_ * 2))) return -1;
        int s = audio_free_.back();
        audio_free_.pop_back();
        return s;
    }
    void release_audio_slot(int s) override { if (s >= 0) audio_free_.push_back(s); }
    int acquire_kv_slot() override {
        if (kv_free_.empty() && !grow_kv(std::max(1, kv_cap_ * 2))) return -1;
        int s = kv_free_.back();
        kv_free_.pop_back();
        return s;
    }
    void release_kv_slot(int s) override { if (s >= 0) kv_free_.push_back(s); }
    bool reserve_slots(int n_audio_free, int n_kv_free) override {
        CUDA_OK(cudaSetDevice(device_));
        if ((int)audio_free_.size() < n_audio_free && !grow_audio(audio_cap_ + n_audio_free - (int)audio_free_.size())) return false;
        if ((int)kv_free_.size() < n_kv_free && !grow_kv(kv_cap_ + n_kv_free - (int)kv_free_.size())) return false;
        std::sort(audio_free_.begin(), audio_free_.end(), std::greater<int>());  // pop ascending: consecutive slots per batch
        std::sort(kv_free_.begin(), kv_free_.end(), std::greater<int>());
        return true;
    }

    void free_mel(DeviceMel& m) override {
        if (m.raw) mel_cache_.push_back(m);
        m = DeviceMel();
    }

    // ------------------------------------------------------------------ K1
    bool compute_mel(const std::vector<MelRequest>& reqs) override {
        CUDA_OK(cudaSetDevice(device_));
        const int n = (int)reqs.size();
        if (n == 0) return true;
        size_t total = 0;
        std::vector<size_t> off(n);
        for (int i = 0; i < n; ++i) { off[i] = total; total += align_up((size_t)std::max(reqs[i].n_samples, 0), 4); }
        if (total > pcm_cap_) {
            if (pcm_dev_) CUDA_OK(cudaFree(pcm_dev_));
            pcm_dev_ = nullptr;
            pcm_cap_ = align_up(total + total / 4, 1024);
            CUDA_OK(cudaMalloc(&pcm_dev_, pcm_cap_ * sizeof(float)));
        }
        if (!ensure_pin(sizeof(MelJob) * n)) return false;
        MelJob* jobs = reinterpret_cast<MelJob*>(pin_);
        int max_frames = 0;
        CUDA_OK(cudaEventRecord(ev_[0], stream_));
        for (int i = 0; i < n; ++i) {
            const MelRequest& r = reqs[i];
            DeviceMel& m = *r.mel;
            const int ns = r.n_samples;
            m.n_mel = hp_.n_mels;
            m.n_len = (ns + 480000) / kHop;
            m.n_len_org = 1 + (ns + 200 - kNFft) / kHop;
            m.n_frames = std::min(m.n_len, (ns + 200 + kHop - 1) / kHop);
            const size_t need = (size_t)std::max(m.n_frames, 1) * m.n_mel;
            if (!m.raw || m.raw_cap < need) {
                if (m.raw) mel_cache_.push_back(m);
                m.raw = nullptr;
                if (!take_cached_mel(m, need)) {
                    m.raw_cap = align_up(need, 4096);
                    CUDA_OK(cudaMalloc(&m.raw, m.raw_cap * sizeof(float)));
                    CUDA_OK(cudaMalloc(&m.max_key, sizeof(int)));
                }
            }
            CUDA_OK(cudaMemcpyAsync(m.max_key, &init_key_, sizeof(int), cudaMemcpyHostToDevice, stream_));
            if (ns > 0) CUDA_OK(cudaMemcpyAsync(pcm_dev_ + off[i], r.pcm, (size_t)ns * sizeof(float), cudaMemcpyDefault, stream_));
            jobs[i] = MelJob{pcm_dev_ + off[i], ns, m.n_frames, m.raw, m.max_key};
            max_frames = std::max(max_frames, m.n_frames);
        }
        if (!ensure_dev_scratch(sizeof(MelJob) * n)) return false;
        CUDA_OK(cudaMemcpyAsync(dev_scratch_, jobs, sizeof(MelJob) * n, cudaMemcpyHostToDevice, stream_));
        launch_mel_stft(mel_tables_, reinterpret_cast<const MelJob*>(dev_scratch_), n, max_frames, stream_);
        CUDA_OK(cudaEventRecord(ev_[1], stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));  // the pinned job table and the caller's PCM are borrowed
        CUDA_OK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_[0], ev_[1]);
        stats.ms_mel += ms;
        return true;
    }

    // ------------------------------------------------------------------ K2-K4
    // Lanes may be driven by one host thread each (full.cpp): the encoder (one stream, one set of activations) is taken by one
    // thread at a time, while the other lanes keep decoding.
    // encode_async queues the windows on the encoder stream and returns a ticket; encode_wait blocks the calling thread until the
    // ticket's encoder output and cross-KV panels are complete.  A lane that has other audios to decode keeps decoding them while
    // the windows of the audios that became due are encoded (full.cpp), instead of standing still for the ~10 ms a window takes.
    bool encode_async(const std::vector<EncodeRequest>& reqs, long* ticket) override {
        std::lock_guard<std::mutex> enc_lock(enc_mu_);
        CUDA_OK(cudaSetDevice(device_));
        const long seq = enc_issued_ + 1;
        EncTicket& t = enc_ring_[seq % kEncRing];
        if (t.seq > 0 && !t.counted) {   // the ring slot's previous ticket (kEncRing calls ago) has to be complete before its events are re-recorded
            CUDA_OK(cudaEventSynchronize(t.end));
            account_ticket(t);
        }
        if (!t.begin) { CUDA_OK(cudaEventCreate(&t.begin)); CUDA_OK(cudaEventCreate(&t.end)); }
        CUDA_OK(cudaEventRecord(t.begin, stream_));
        bool ok = true;
        for (size_t b0 = 0; ok && b0 < reqs.size(); b0 += enc_batch_) {
            const int nb = (int)std::min<size_t>(enc_batch_, reqs.size() - b0);
            ok = encode_batch(&reqs[b0], nb);
        }
        CUDA_OK(cudaEventRecord(t.end, stream_));
        t.seq = seq; t.counted = false;
        enc_issued_ = seq;
        *ticket = seq;
        return ok;
    }
    bool encode_wait(long ticket) override {
        cudaEvent_t ev = nullptr;
        {
            std::lock_guard<std::mutex> enc_lock(enc_mu_);
            if (ticket <= enc_done_ || ticket > enc_issued_) return ticket <= enc_done_;
            const EncTicket& t = enc_ring_[ticket % kEncRing];
            if (t.seq == ticket) ev = t.end;   // else: the slot was recycled, which only happens after its ticket completed
        }
        CUDA_OK(cudaSetDevice(device_));
        if (ev) CUDA_OK(cudaEventSynchronize(ev));   // outside the lock: other lanes may queue their windows meanwhile
        std::lock_guard<std::mutex> enc_lock(enc_mu_);
        for (long q = enc_done_ + 1; q <= ticket; ++q) {   // the stream is in order: everything up to the ticket is complete
            EncTicket& t = enc_ring_[q % kEncRing];
            if (t.seq == q && !t.counted) account_ticket(t);
        }
        if (ticket > enc_done_) enc_done_ = ticket;
        if (enc_done_ == enc_issued_) {   // nothing queued behind it: every timing mark of the stream is complete
            std::lock_guard<std::mutex> stats_lock(stats_mu_);
            collect_marks();
        }
        CUDA_OK(cudaGetLastError());
        return true;
    }
    struct EncTicket {
        cudaEvent_t begin = nullptr, end = nullptr;
        long seq = 0;
        bool counted = true;
    };
    void account_ticket(EncTicket& t) {   // enc_mu_ held, t.end complete
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.begin, t.end) == cudaSuccess) {
            std::lock_guard<std::mutex> stats_lock(stats_mu_);
            stats.ms_encode += ms;
        }
        t.counted = true;
    }

    // ---- per-launch event timing (profiling mode); one context per stream
    struct Marks {
        cudaStream_t s = nullptr;
        std::vector<cudaEvent_t> pool;
        size_t used = 0;
        std::vector<std::pair<int, int>> marks;
    };
    void mark_begin(Marks& m, bool on) {
        if (!on) return;
        if (m.used + 2 > m.pool.size()) {
            for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); m.pool.push_back(e); }
        }
        cudaEventRecord(m.pool[m.used], m.s);
    }
    void mark_end(Marks& m, bool on, int cls) {
        if (!on) return;
        cudaEventRecord(m.pool[m.used + 1], m.s);
        m.marks.push_back({cls, (int)m.used});
        m.used += 2;
    }
    void collect_marks(Marks& m) {  // call after the stream has been synchronised
        for (const auto& k : m.marks) {
            float ms = 0;
            cudaEventElapsedTime(&ms, m.pool[k.second], m.pool[k.second + 1]);
            if (k.first >= 10) { detail_ms_[k.first - 10] += ms; detail_n_[k.first - 10]++; }
            else if (k.first == 0) { stats.ms_enc_gemm += ms; stats.n_enc_gemm++; }
            else if (k.first == 1) { stats.ms_enc_attn += ms; stats.n_enc_attn++; }
            else { stats.ms_dec_cross += ms; stats.n_dec_cross++; }
        }
        m.marks.clear();
        m.used = 0;
    }
    void mark_begin() { mark_begin(main_marks_, profiling); }
    void mark_end(int cls) { mark_end(main_marks_, profiling, cls); }
    void collect_marks() { collect_marks(main_marks_); }
    template <typename TA, typename TC>
    bool tgemm(const TA* A, int lda, const TA* W, int ldw, TC* C, int ldc, int M, int N, int K, const Epilogue& e) {
        mark_begin();
        const bool ok = gemm(A, lda, W, ldw, C, ldc, M, N, K, e, stream_);
        mark_end(0);
        return ok;
    }

    bool encode_batch(const EncodeRequest* reqs, int nb) {
        const int d = d_, nm = hp_.n_mels, M = nb * kWinRows;
        // window descriptors
        // window descriptors travel through a ring of pinned tables: a table is rewritten only after the copy that read it has completed
        // (an event per table), so queueing a batch never waits for the encoder stream to drain
        if (!enc_pin_) {
            CUDA_OK(cudaMallocHost(&enc_pin_, sizeof(PackJob) * (size_t)enc_batch_ * kEncPin));
            for (auto& e : enc_pin_ev_) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        const int pslot = (int)(enc_pin_next_++ % kEncPin);
        if (enc_pin_next_ > kEncPin) CUDA_OK(cudaEventSynchronize(enc_pin_ev_[pslot]));
        PackJob* pj = enc_pin_ + (size_t)pslot * enc_batch_;
        for (int w = 0; w < nb; ++w) {
            const DeviceMel& m = *reqs[w].mel;
            if (!m.raw || reqs[w].audio_slot < 0 || reqs[w].audio_slot >= audio_cap_) { set_err("encode: bad request"); return false; }
            pj[w] = PackJob{m.raw, m.max_key, m.n_frames, m.n_len, reqs[w].seek};
        }
        if (!ensure_dev_scratch(sizeof(PackJob) * nb)) return false;
        CUDA_OK(cudaMemcpyAsync(dev_scratch_, pj, sizeof(PackJob) * nb, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaEventRecord(enc_pin_ev_[pslot], stream_));
        if (!rezero_conv_pads(nb)) return false;
        launch_pack_mel<T>(reinterpret_cast<const PackJob*>(dev_scratch_), nb, nm, e_mel_, stream_);
        // conv1: rows (w, t) read the 3 neighbouring mel frames in place (row stride n_mels, K = 3*n_mels)
        {
            Epilogue e;
            e.bias = conv1_b_; e.act = 1; e.win_rows = kWinRowsIn; e.valid_rows = 3000;
            if (!tgemm(e_mel_, nm, conv1_w_, 3 * nm, e_h1_ + d, d, nb * kWinRowsIn, d, 3 * nm, e)) return gemm_fail();
        }
        // conv2 (stride 2): row (w, t') reads h1 rows 2t'-1..2t'+1 in place (row stride 2d, K = 3d); + positional embedding
        {
            Epilogue e;
            e.bias = conv2_b_; e.act = 1; e.res = enc_pos_; e.res_ld = d; e.res_mod = kWinRows;
            if (!tgemm(e_h1_, 2 * d, conv2_w_, 3 * d, e_x_, d, M, d, 3 * d, e)) return gemm_fail();
        }
        for (int l = 0; l < hp_.n_audio_layer; ++l) {
            const Layer<T>& L = enc_[l];
            launch_layernorm<T>(e_x_, d, L.ln1_g, L.ln1_b, e_y_, d, M, d, stream_);
            { Epilogue e; e.bias = L.bqkv; if (!tgemm(e_y_, d, L.wqkv, d, e_qkv_, 3 * d, M, 3 * d, d, e)) return gemm_fail(); }
            mark_begin();
            if (!enc_attention(e_qkv_, e_att_, nb, hp_.n_audio_head, d, stream_)) return gemm_fail();
            mark_end(1);
            { Epilogue e; e.bias = L.bo; e.res = e_x_; e.res_ld = d; if (!tgemm(e_att_, d, L.wo, d, e_x_, d, M, d, d, e)) return gemm_fail(); }
            launch_layernorm<T>(e_x_, d, L.ln2_g, L.ln2_b, e_y_, d, M, d, stream_);
            { Epilogue e; e.bias = L.b1; e.act = 1; if (!tgemm(e_y_, d, L.w1, d, e_h_, 4 * d, M, 4 * d, d, e)) return gemm_fail(); }
            { Epilogue e; e.bias = L.b2; e.res = e_x_; e.res_ld = d; if (!tgemm(e_h_, 4 * d, L.w2, 4 * d, e_x_, d, M, d, 4 * d, e)) return gemm_fail(); }
        }
        launch_layernorm<T>(e_x_, d, enc_lnp_g_, enc_lnp_b_, e_y_, d, M, d, stream_);
        // keep the encoder output per audio slot, then project every decoder layer's cross K/V
        // with one GEMM per run of consecutive slots: out rows land directly in the cross-KV pool.
        const int ldx = 2 * hp_.n_text_layer * d;
        int w = 0;
        while (w < nb) {
            int run = 1;
            while (w + run < nb && reqs[w + run].audio_slot == reqs[w].audio_slot + run) ++run;
            const int s0 = reqs[w].audio_slot;
            CUDA_OK(cudaMemcpyAsync(encout_pool_ + (size_t)s0 * kWinRows * d, e_y_ + (size_t)w * kWinRows * d,
                                    (size_t)run * kWinRows * d * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
            Epilogue e;
            e.bias = cross_b_;
            e.head_rows = kWinRows;  // head-major panels [slot][(layer, K|V, head)][1536][64]: contiguous key streams for the decoder
            if (!tgemm(e_y_ + (size_t)w * kWinRows * d, d, cross_w_, d, cross_pool_ + (size_t)s0 * kWinRows * ldx, ldx, run * kWinRows, ldx, d, e))
                return gemm_fail();
            w += run;
        }
        return true;
    }

    // ------------------------------------------------------------------ K5 + K6
    // One decode lane: a stream with its own activations, staging and timing context.
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        float* x = nullptr;
        T *y = nullptr, *qkv = nullptr, *att = nullptr, *h = nullptr, *ys = nullptr;
        float* logits = nullptr;
        float* probs = nullptr;           // [S][n_vocab] filtered probabilities of the last round (K6 scratch)
        float* partial = nullptr;
        unsigned int* bar = nullptr;      // device-wide barrier state of the fused projection chains (count, generation)
        float2* stats = nullptr;          // [40][128] per-tile (mean, M2) of the residual stream (workspace of the fused projections' LayerNorm tail)
        int* ticket = nullptr;            // "last cluster" ticket of the fused projections
        int* sched = nullptr;             // 2 x (work, exit) counters of the streaming cross-attention kernel: consecutive
        unsigned cross_seq = 0;           // launches alternate, a launch may start its prologue while the previous one drains
        char* pin = nullptr; size_t pin_cap = 0;
        char* scratch = nullptr; size_t scratch_cap = 0;
        Marks tm;
        bool inflight = false, has_events = false;
        // the chunk left in flight by decode_submit
        int pend_S = 0; size_t pend_off = 0; size_t pend_pin_off = 0;
        std::vector<SampleResult> results;
        int last_logit_rows = 0;
        size_t last_logit_base = 0;       // lane-wide index of the first sample whose logits the lane still holds (last chunk only)
        std::chrono::steady_clock::time_point t_submit;
        double issue_ms = 0, cross_bytes = 0;   // folded into the engine's counters when the round is collected
        struct GraphEntry { cudaGraphExec_t exec; long n_kernels; };
        std::unordered_map<uint64_t, GraphEntry> graphs;   // captured step rounds by shape (decode_chunk)
        std::unordered_map<uint64_t, int> graph_seen;
        unsigned graph_epoch = 0;
        CrossGroups cross_grp;                  // row groups of the chunk being queued (device table inside `scratch`)
        int cross_slots = 0;                    // distinct audio slots among the rows of the chunk being queued: rows of one audio (a pass and its
                                                // speculative successor, beams) stream the same cross-KV panels, which leave HBM once
    };
    int n_lanes() const override { return (int)lanes_.size(); }

    bool lane_pin(Lane& L, size_t bytes) {
        if (bytes <= L.pin_cap) return true;
        if (L.pin) { CUDA_OK(cudaStreamSynchronize(L.stream)); CUDA_OK(cudaFreeHost(L.pin)); L.pin = nullptr; }
        L.pin_cap = align_up(bytes * 2, 1 << 16);
        CUDA_OK(cudaMallocHost(&L.pin, L.pin_cap));
        return true;
    }
    bool lane_scratch(Lane& L, size_t bytes) {
        if (bytes <= L.scratch_cap) return true;
        if (L.scratch) { CUDA_OK(cudaStreamSynchronize(L.stream)); CUDA_OK(cudaFree(L.scratch)); L.scratch = nullptr; }
        L.scratch_cap = align_up(bytes * 2, 1 << 16);
        CUDA_OK(cudaMalloc(&L.scratch, L.scratch_cap));
        return true;
    }

    bool decode_submit(int lane, const std::vector<RowDesc>& rows, const std::vector<int>& sample_rows, const std::vector<SampleParams>& sp,
                       float* logits_host, const float* inject, const unsigned char* inject_mask) override {
        CUDA_OK(cudaSetDevice(device_));
        if (lane < 0 || lane >= (int)lanes_.size()) { set_err("decode: no such lane"); return false; }
        Lane& L = lanes_[lane];
        if (L.inflight) { set_err("decode: lane is busy"); return false; }
        if (sp.size() != sample_rows.size()) { set_err("decode: params/sample size mismatch"); return false; }
        L.results.assign(sample_rows.size(), SampleResult{});
        L.pend_S = 0;
        L.inflight = true;
        L.t_submit = std::chrono::steady_clock::now();
        L.has_events = !rows.empty();
        if (rows.empty()) return true;
        CUDA_OK(cudaEventRecord(L.ev[0], L.stream));
        size_t si = 0;
        for (size_t r0 = 0; r0 < rows.size();) {
            // a chunk holds at most dec_rows_ rows and dec_samples_ sample rows; a step-sized round that is a little too big for the
            // step kernels (R <= 128) is cut into equal chunks that all take them instead of one pass through the prefill path
            size_t limit = (size_t)dec_rows_;
            if (use_skinny_ && sizeof(T) == 2 && rows.size() > 128 && rows.size() <= 512) limit = (rows.size() + (rows.size() + 127) / 128 - 1) / ((rows.size() + 127) / 128);
            size_t r1 = std::min(rows.size(), r0 + limit);
            size_t sj = si;
            while (sj < sample_rows.size() && (size_t)sample_rows[sj] < r1) {
                if (sj - si == (size_t)dec_samples_) {
                    // the sample budget ends inside this chunk: cut in front of the row sample sj refers to.  Several samples
                    // may share that row (prefill with best_of / beam_size > 1 samples the prompt's last row n_cur times):
                    // all of them move to the next chunk, so every sample of a chunk indexes a row inside [r0, r1).
                    const size_t row = (size_t)sample_rows[sj];
                    while (sj > si && (size_t)sample_rows[sj - 1] == row) --sj;
                    r1 = row;
                    break;
                }
                ++sj;
            }
            if (r1 == r0) { set_err("decode: cannot make progress"); L.inflight = false; return false; }
            const bool last = r1 == rows.size();
            if (!decode_chunk(L, rows.data() + r0, (int)(r1 - r0), sample_rows.data() + si, (int)(sj - si), (int)r0, sp.data() + si, si,
                              logits_host ? logits_host + si * (size_t)hp_.n_vocab : nullptr, inject ? inject + si * (size_t)hp_.n_vocab : nullptr,
                              inject_mask ? inject_mask + si : nullptr)) { L.inflight = false; return false; }
            if (!last && !finish_chunk(L)) { L.inflight = false; return false; }
            si = sj;
            r0 = r1;
        }
        CUDA_OK(cudaEventRecord(L.ev[1], L.stream));
        return true;
    }

    bool decode_collect(int lane, std::vector<SampleResult>& results) override {
        CUDA_OK(cudaSetDevice(device_));
        if (lane < 0 || lane >= (int)lanes_.size()) { set_err("decode: no such lane"); return false; }
        Lane& L = lanes_[lane];
        if (!L.inflight) { set_err("decode: nothing in flight on this lane"); return false; }
        L.inflight = false;
        const auto t0 = std::chrono::steady_clock::now();
        if (!finish_chunk(L)) return false;
        const double waited = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        {
            std::lock_guard<std::mutex> stats_lock(stats_mu_);
            host_wait_ms_ += waited;
            host_issue_ms_ += L.issue_ms; L.issue_ms = 0;
            stats.dec_cross_bytes += L.cross_bytes; L.cross_bytes = 0;
            if (L.has_events) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, L.ev[0], L.ev[1]) == cudaSuccess) stats.ms_decode += ms;
            }
            collect_marks(L.tm);
        }
        results = L.results;
        return true;
    }

    // wait for the chunk in flight on a lane and take its sample results
    bool finish_chunk(Lane& L) {
        CUDA_OK(cudaStreamSynchronize(L.stream));
        CUDA_OK(cudaGetLastError());
        if (L.pend_S > 0) memcpy(L.results.data() + L.pend_off, L.pin + L.pend_pin_off, sizeof(SampleResult) * L.pend_S);
        L.pend_S = 0;
        return true;
    }

    // Decoder layers for a step batch of R <= 128 token rows (bf16): every projection is a swap-AB split-K
    // tcgen05 GEMM that streams its weights through all SMs, finished by one fused epilogue kernel
    // (bias / GELU / residual / next LayerNorm / KV scatter).  14 launches per layer.
    bool decode_layers_skinny(Lane& Ln, const RowDesc* drows, int R, bool distinct_slots) {
        const int d = d_, Ld = hp_.n_text_layer, ntc = hp_.n_text_ctx;
        cudaStream_t st = Ln.stream;
        const size_t self_head = (size_t)ntc * 64, self_kv = (size_t)ntc * d;
        const size_t self_slot = (size_t)Ld * 2 * self_kv;
        const size_t cross_head = (size_t)kWinRows * 64, cross_kv = (size_t)kWinRows * d;
        const size_t cross_slot = (size_t)Ld * 2 * cross_kv;
        const bf16* y = reinterpret_cast<const bf16*>(Ln.y);
        auto proj = [&](const void* X, int K, const T* W, int N, SkinnyEpilogue& e) -> bool {
            int splits = 0;
            mark_begin(Ln.tm, detail_);
            if (!launch_gemm_skinny_bf16_sm100(reinterpret_cast<const bf16*>(X), K, reinterpret_cast<const bf16*>(W), K, Ln.partial, R, N, K, &splits, st))
                return gemm_fail();
            mark_end(Ln.tm, detail_, 10);
            e.partial = Ln.partial; e.splits = splits; e.R = R; e.N = N;
            mark_begin(Ln.tm, detail_);
            launch_skinny_reduce<T>(e, st);
            mark_end(Ln.tm, detail_, 11);
            return true;
        };
        launch_layernorm<T>(Ln.x, d, dec_[0].ln1_g, dec_[0].ln1_b, Ln.y, d, R, d, st);
        for (int l = 0; l < Ld; ++l) {
            const Layer<T>& L = dec_[l];
            T* kc = self_pool_ + (size_t)l * 2 * self_kv;
            T* vc = kc + self_kv;
            if (fuse_qkv_ && distinct_slots) {
                // QKV projection: split-K GEMM only; the self-attention kernel sums the partials (+ bias), appends the new
                // key / value row to the cache and attends — one launch less on the chain.  Needs one row per KV slot.
                int splits = 0;
                mark_begin(Ln.tm, detail_);
                if (!launch_gemm_skinny_bf16_sm100(y, d, reinterpret_cast<const bf16*>(L.wqkv), d, Ln.partial, R, 3 * d, d, &splits, st)) return gemm_fail();
                mark_end(Ln.tm, detail_, 10);
                QkvPartials qp;
                qp.partial = Ln.partial; qp.splits = splits; qp.plane = (size_t)R * 3 * d; qp.ld = 3 * d; qp.bias = L.bqkv;
                mark_begin(Ln.tm, detail_);
                launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st, &qp);
                mark_end(Ln.tm, detail_, 12);
            } else {
                {   // QKV projection + KV-cache append
                    SkinnyEpilogue e;
                    e.bias = L.bqkv; e.out = Ln.qkv; e.out_ld = 3 * d;
                    e.rows = drows; e.kpanel = kc; e.vpanel = vc; e.slot_stride = self_slot; e.n_pos_cap = ntc; e.d = d;
                    if (!proj(y, d, L.wqkv, 3 * d, e)) return false;
                }
                mark_begin(Ln.tm, detail_);
                launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st);
                mark_end(Ln.tm, detail_, 12);
            }
            {   // out projection + residual + cross-attention LayerNorm
                SkinnyEpilogue e;
                e.bias = L.bo; e.x = Ln.x; e.ln_g = L.lnc_g; e.ln_b = L.lnc_b; e.y = Ln.y;
                if (!proj(Ln.att, d, L.wo, d, e)) return false;
            }
            const T* ck = cross_pool_ + (size_t)l * 2 * cross_kv;
            const T* cv = ck + cross_kv;
            const bool timed = profiling && (l % kCrossSample) == 0;
            if (cross_mode_ == 2 && fuse_cross_q_) {
                // cross-attention query: split-K GEMM only; the attention kernel sums the partials (+ bias) itself
                int splits = 0;
                mark_begin(Ln.tm, detail_);
                if (!launch_gemm_skinny_bf16_sm100(y, d, reinterpret_cast<const bf16*>(L.wcq), d, Ln.partial, R, d, d, &splits, st)) return gemm_fail();
                mark_end(Ln.tm, detail_, 10);
                CrossQPartials qp;
                qp.partial = Ln.partial; qp.splits = splits; qp.plane = (size_t)R * d; qp.ld = d; qp.bias = L.bcq;
                mark_begin(Ln.tm, timed || detail_);
                if (!launch_dec_cross_attention_tc_sm100(drows, R, nullptr, d, reinterpret_cast<const bf16*>(cross_pool_), (size_t)audio_cap_ * cross_slot,
                                                         (size_t)(ck - cross_pool_), (size_t)(cv - cross_pool_), reinterpret_cast<bf16*>(Ln.att), d, hp_.n_text_head,
                                                         cross_slot, hp_.n_audio_ctx, Ln.sched + ((Ln.cross_seq++ & 1u) << 1), cross_ctas_, st, &qp, &Ln.cross_grp))
                    return gemm_fail();
            } else {
                {   // cross-attention query
                    SkinnyEpilogue e;
                    e.bias = L.bcq; e.out = Ln.qkv; e.out_ld = d;
                    if (!proj(y, d, L.wcq, d, e)) return false;
                }
                mark_begin(Ln.tm, timed || detail_);
                if (!cross_attention(Ln, drows, R, Ln.qkv, d, ck, cv, Ln.att, cross_slot, cross_head)) return false;
            }
            if (timed) { mark_end(Ln.tm, true, 2); Ln.cross_bytes += 2.0 * Ln.cross_slots * hp_.n_audio_ctx * d * sizeof(T); } else mark_end(Ln.tm, detail_, 13);
            {   // cross out projection + residual + MLP LayerNorm
                SkinnyEpilogue e;
                e.bias = L.bco; e.x = Ln.x; e.ln_g = L.ln2_g; e.ln_b = L.ln2_b; e.y = Ln.y;
                if (!proj(Ln.att, d, L.wco, d, e)) return false;
            }
            if (fc1_fused_ && dec_proj_supported(R, 4 * d, d)) {
                // FC1 + GELU in ONE launch (decode_proj_sm100.cu): nothing after it needs whole rows, so the K split is reduced inside a
                // 2-CTA cluster and bias + GELU + the bf16 store happen in the same kernel — one dependent step less per layer
                ProjDesc p;
                p.R = R; p.N = 4 * d; p.K = d; p.W = reinterpret_cast<const bf16*>(L.w1); p.X = y; p.ldx = d; p.bias = L.b1; p.act = 1;
                p.out = reinterpret_cast<bf16*>(Ln.h); p.out_ld = 4 * d;
                mark_begin(Ln.tm, detail_);
                if (!launch_dec_proj_sm100(p, st)) return gemm_fail();
                mark_end(Ln.tm, detail_, 10);
            } else {   // FC1 + GELU
                SkinnyEpilogue e;
                e.bias = L.b1; e.act = 1; e.out = Ln.h; e.out_ld = 4 * d;
                if (!proj(y, d, L.w1, 4 * d, e)) return false;
            }
            {   // FC2 + residual + the next layer's first LayerNorm
                SkinnyEpilogue e;
                e.bias = L.b2; e.x = Ln.x;
                if (l + 1 < Ld) { e.ln_g = dec_[l + 1].ln1_g; e.ln_b = dec_[l + 1].ln1_b; e.y = Ln.y; }
                if (!proj(Ln.h, 4 * d, L.w2, d, e)) return false;
            }
        }
        return true;
    }

    // Step batches with the fused cluster projections (decode_proj_sm100.cu): per layer
    //   self-attention -> [out-proj + residual + LN] -> [cross-query] -> cross-attention -> [cross-out + residual + LN]
    //   -> [FC1 + GELU] -> [FC2 + residual + next LN] -> [next QKV + KV append]
    // = 8 launches; no split-K partials in global memory, no epilogue or LayerNorm kernel between two projections.
    bool decode_layers_proj(Lane& Ln, const RowDesc* drows, int R) {
        const int d = d_, Ld = hp_.n_text_layer, ntc = hp_.n_text_ctx;
        cudaStream_t st = Ln.stream;
        const size_t self_head = (size_t)ntc * 64, self_kv = (size_t)ntc * d;
        const size_t self_slot = (size_t)Ld * 2 * self_kv;
        const size_t cross_kv = (size_t)kWinRows * d;
        const size_t cross_slot = (size_t)Ld * 2 * cross_kv;
        bf16* const y = reinterpret_cast<bf16*>(Ln.y);
        bf16* const qkv = reinterpret_cast<bf16*>(Ln.qkv);
        bf16* const att = reinterpret_cast<bf16*>(Ln.att);
        bf16* const hbuf = reinterpret_cast<bf16*>(Ln.h);
        auto run = [&](ProjDesc& p) -> bool {
            p.R = R;
            mark_begin(Ln.tm, detail_);
            if (!launch_dec_proj_sm100(p, st)) return gemm_fail();
            mark_end(Ln.tm, detail_, 10);
            return true;
        };
        // x += X * W^T + b, then y = LayerNorm(x) * g + be
        auto residual = [&](const bf16* X, int K, const T* W, const float* b, const float* g, const float* be) -> bool {
            ProjDesc p;
            p.N = d; p.K = K; p.W = reinterpret_cast<const bf16*>(W); p.X = X; p.ldx = K; p.bias = b; p.x = Ln.x;
            if (g) { p.y = y; p.ln_g = g; p.ln_b = be; p.stats_out = Ln.stats; p.ticket = Ln.ticket; }
            return run(p);
        };
        auto qkv_proj = [&](int l) -> bool {
            const Layer<T>& L = dec_[l];
            bf16* kc = reinterpret_cast<bf16*>(self_pool_ + (size_t)l * 2 * self_kv);
            ProjDesc p;
            p.N = 3 * d; p.K = d; p.W = reinterpret_cast<const bf16*>(L.wqkv); p.X = y; p.ldx = d; p.bias = L.bqkv;
            p.out = qkv; p.out_ld = 3 * d;
            p.rows = drows; p.kpanel = kc; p.vpanel = kc + self_kv; p.slot_stride = self_slot; p.n_pos_cap = ntc; p.d = d;
            return run(p);
        };
        launch_layernorm<T>(Ln.x, d, dec_[0].ln1_g, dec_[0].ln1_b, Ln.y, d, R, d, st);   // the embedding kernel leaves x only
        if (!qkv_proj(0)) return false;
        for (int l = 0; l < Ld; ++l) {
            const Layer<T>& L = dec_[l];
            T* kc = self_pool_ + (size_t)l * 2 * self_kv;
            T* vc = kc + self_kv;
            mark_begin(Ln.tm, detail_);
            launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st);
            mark_end(Ln.tm, detail_, 12);
            if (!residual(att, d, L.wo, L.bo, L.lnc_g, L.lnc_b)) return false;
            {   // cross-attention query
                ProjDesc p;
                p.N = d; p.K = d; p.W = reinterpret_cast<const bf16*>(L.wcq); p.X = y; p.ldx = d; p.bias = L.bcq; p.out = qkv; p.out_ld = d;
                if (!run(p)) return false;
            }
            const T* ck = cross_pool_ + (size_t)l * 2 * cross_kv;
            const T* cv = ck + cross_kv;
            const bool timed = profiling && (l % kCrossSample) == 0;
            mark_begin(Ln.tm, timed || detail_);
            if (!launch_dec_cross_attention_tc_sm100(drows, R, qkv, d, reinterpret_cast<const bf16*>(cross_pool_), (size_t)audio_cap_ * cross_slot,
                                                     (size_t)(ck - cross_pool_), (size_t)(cv - cross_pool_), att, d, hp_.n_text_head, cross_slot, hp_.n_audio_ctx,
                                                     Ln.sched + ((Ln.cross_seq++ & 1u) << 1), cross_ctas_, st, nullptr, &Ln.cross_grp))
                return gemm_fail();
            if (timed) { mark_end(Ln.tm, true, 2); Ln.cross_bytes += 2.0 * Ln.cross_slots * hp_.n_audio_ctx * d * sizeof(T); } else mark_end(Ln.tm, detail_, 13);
            if (!residual(att, d, L.wco, L.bco, L.ln2_g, L.ln2_b)) return false;
            {   // FC1 + GELU
                ProjDesc p;
                p.N = 4 * d; p.K = d; p.W = reinterpret_cast<const bf16*>(L.w1); p.X = y; p.ldx = d; p.bias = L.b1; p.act = 1; p.out = hbuf; p.out_ld = 4 * d;
                if (!run(p)) return false;
            }
            const bool more = l + 1 < Ld;   // FC2 + residual (+ the next layer's first LayerNorm)
            if (!residual(hbuf, 4 * d, L.w2, L.b2, more ? dec_[l + 1].ln1_g : nullptr, more ? dec_[l + 1].ln1_b : nullptr)) return false;
            if (more && !qkv_proj(l + 1)) return false;
        }
        return true;
    }

    // The same layers with the projection chains fused (decode_chain_sm100.cu): per layer 4 launches instead of 12 —
    // [out-proj + LN + cross-query] -> cross-attention -> [cross-out + LN + FC1 + GELU + FC2 + LN + next QKV] -> self-attention.
    // Arithmetic and split order are those of decode_layers_skinny: results are bit-identical.
    bool decode_layers_chain(Lane& Ln, const RowDesc* drows, int R, bool distinct_slots) {
        const int d = d_, Ld = hp_.n_text_layer, ntc = hp_.n_text_ctx;
        cudaStream_t st = Ln.stream;
        const size_t self_head = (size_t)ntc * 64, self_kv = (size_t)ntc * d;
        const size_t self_slot = (size_t)Ld * 2 * self_kv;
        const size_t cross_kv = (size_t)kWinRows * d;
        const size_t cross_slot = (size_t)Ld * 2 * cross_kv;
        const bool fuse_qkv = fuse_qkv_ && distinct_slots;
        const bf16* y = reinterpret_cast<const bf16*>(Ln.y);
        const bf16* att = reinterpret_cast<const bf16*>(Ln.att);
        const bf16* h = reinterpret_cast<const bf16*>(Ln.h);
        struct Chain {
            ChainDesc cd;
            const bf16* W[kChainMaxSteps];
            const bf16* X[kChainMaxSteps];
            int ldx[kChainMaxSteps];
            int n = 0;
        };
        auto add = [&](Chain& c, const bf16* X, int K, const T* W, int N) -> SkinnyEpilogue& {
            ChainStep& s = c.cd.step[c.n];
            s = ChainStep();
            s.N = N; s.K = K; s.reduce = 1;
            c.W[c.n] = reinterpret_cast<const bf16*>(W); c.X[c.n] = X; c.ldx[c.n] = K;
            return c.cd.step[c.n++].e;
        };
        auto run = [&](Chain& c) -> bool {
            c.cd.n_steps = c.n; c.cd.R = R; c.cd.partial = Ln.partial; c.cd.bar = Ln.bar;
            mark_begin(Ln.tm, detail_);
            if (!launch_dec_chain_sm100(c.cd, c.W, c.X, c.ldx, chain_stages_, st)) return gemm_fail();
            mark_end(Ln.tm, detail_, 10);
            return true;
        };
        // QKV projection of layer l as the last step of a chain: either left as partial sums for the self-attention
        // kernel (one row per KV slot) or reduced here with the KV-cache append
        auto add_qkv = [&](Chain& c, int l) {
            const Layer<T>& L = dec_[l];
            SkinnyEpilogue& e = add(c, y, d, L.wqkv, 3 * d);
            if (fuse_qkv) { c.cd.step[c.n - 1].reduce = 0; return; }
            T* kc = self_pool_ + (size_t)l * 2 * self_kv;
            e.bias = L.bqkv; e.out = Ln.qkv; e.out_ld = 3 * d;
            e.rows = drows; e.kpanel = kc; e.vpanel = kc + self_kv; e.slot_stride = self_slot; e.n_pos_cap = ntc; e.d = d;
        };
        launch_layernorm<T>(Ln.x, d, dec_[0].ln1_g, dec_[0].ln1_b, Ln.y, d, R, d, st);
        {
            Chain c;
            add_qkv(c, 0);
            if (!run(c)) return false;
        }
        for (int l = 0; l < Ld; ++l) {
            const Layer<T>& L = dec_[l];
            T* kc = self_pool_ + (size_t)l * 2 * self_kv;
            T* vc = kc + self_kv;
            mark_begin(Ln.tm, detail_);
            if (fuse_qkv) {
                QkvPartials qp;
                qp.partial = Ln.partial; qp.splits = skinny_gemm_splits(3 * d, d); qp.plane = (size_t)R * 3 * d; qp.ld = 3 * d; qp.bias = L.bqkv;
                launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st, &qp);
            } else {
                launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st);
            }
            mark_end(Ln.tm, detail_, 12);
            {   // out projection + residual + cross-attention LayerNorm, then the cross-attention query (left as partial sums)
                Chain c;
                SkinnyEpilogue& e = add(c, att, d, L.wo, d);
                e.bias = L.bo; e.x = Ln.x; e.ln_g = L.lnc_g; e.ln_b = L.lnc_b; e.y = Ln.y;
                add(c, y, d, L.wcq, d);
                c.cd.step[c.n - 1].reduce = 0;
                if (!run(c)) return false;
            }
            const T* ck = cross_pool_ + (size_t)l * 2 * cross_kv;
            const T* cv = ck + cross_kv;
            const bool timed = profiling && (l % kCrossSample) == 0;
            CrossQPartials qp;
            qp.partial = Ln.partial; qp.splits = skinny_gemm_splits(d, d); qp.plane = (size_t)R * d; qp.ld = d; qp.bias = L.bcq;
            mark_begin(Ln.tm, timed || detail_);
            if (!launch_dec_cross_attention_tc_sm100(drows, R, nullptr, d, reinterpret_cast<const bf16*>(cross_pool_), (size_t)audio_cap_ * cross_slot,
                                                     (size_t)(ck - cross_pool_), (size_t)(cv - cross_pool_), reinterpret_cast<bf16*>(Ln.att), d, hp_.n_text_head,
                                                     cross_slot, hp_.n_audio_ctx, Ln.sched + ((Ln.cross_seq++ & 1u) << 1), cross_ctas_, st, &qp, &Ln.cross_grp))
                return gemm_fail();
            if (timed) { mark_end(Ln.tm, true, 2); Ln.cross_bytes += 2.0 * Ln.cross_slots * hp_.n_audio_ctx * d * sizeof(T); } else mark_end(Ln.tm, detail_, 13);
            {   // cross out-projection + residual + MLP LayerNorm, FC1 + GELU, FC2 + residual + next LayerNorm, next layer's QKV
                Chain c;
                SkinnyEpilogue& e1 = add(c, att, d, L.wco, d);
                e1.bias = L.bco; e1.x = Ln.x; e1.ln_g = L.ln2_g; e1.ln_b = L.ln2_b; e1.y = Ln.y;
                SkinnyEpilogue& e2 = add(c, y, d, L.w1, 4 * d);
                e2.bias = L.b1; e2.act = 1; e2.out = Ln.h; e2.out_ld = 4 * d;
                SkinnyEpilogue& e3 = add(c, h, 4 * d, L.w2, d);
                e3.bias = L.b2; e3.x = Ln.x;
                if (l + 1 < Ld) {
                    e3.ln_g = dec_[l + 1].ln1_g; e3.ln_b = dec_[l + 1].ln1_b; e3.y = Ln.y;
                    add_qkv(c, l + 1);
                }
                if (!run(c)) return false;
            }
        }
        return true;
    }

    // logits = ys * tok_emb^T.  For a step batch (S <= 128 sample rows, bf16) this is the same weight-streaming shape as
    // the layer projections: the swap-AB kernel with one K split writes fp32 [S][n_vocab] directly (406 weight tiles
    // over all SMs) instead of two ragged waves of 128 x 256 tiles.
    bool logits_gemm(Lane& Ln, int S) {
        const int d = d_;
        if (skinny_logits_ && sizeof(T) == 2 && S <= 128 && skinny_gemm_splits(hp_.n_vocab, d) == 1) {
            int splits = 0;
            if (!launch_gemm_skinny_bf16_sm100(reinterpret_cast<const bf16*>(Ln.ys), d, reinterpret_cast<const bf16*>(tok_emb_), d, Ln.logits, S, hp_.n_vocab, d,
                                               &splits, Ln.stream))
                return gemm_fail();
            return true;
        }
        Epilogue e;
        if (!gemm(Ln.ys, d, tok_emb_, d, Ln.logits, hp_.n_vocab, S, hp_.n_vocab, d, e, Ln.stream)) return gemm_fail();
        return true;
    }

    // cross-attention of R single-token rows over the head-major cross-KV panels of one layer
    bool cross_attention(Lane& Ln, const RowDesc* drows, int R, const float* q, int ldq, const float* ck, const float* cv, float* out, size_t cross_slot,
                         size_t cross_head) {
        launch_dec_attention<float>(drows, R, q, ldq, ck, cv, out, d_, hp_.n_text_head, /*cross=*/1, cross_slot, cross_head, hp_.n_audio_ctx, Ln.stream);
        return true;
    }
    bool cross_attention(Lane& Ln, const RowDesc* drows, int R, const bf16* q, int ldq, const bf16* ck, const bf16* cv, bf16* out, size_t cross_slot,
                         size_t cross_head) {
        if (cross_mode_ == 0) {
            launch_dec_attention<bf16>(drows, R, q, ldq, ck, cv, out, d_, hp_.n_text_head, /*cross=*/1, cross_slot, cross_head, hp_.n_audio_ctx, Ln.stream);
            return true;
        }
        if (cross_mode_ == 1) {
            if (!launch_dec_cross_attention_sm100(drows, R, q, ldq, ck, cv, out, d_, hp_.n_text_head, cross_slot, cross_head, hp_.n_audio_ctx, cross_ctas_, Ln.stream))
                return gemm_fail();
            return true;
        }
        if (!launch_dec_cross_attention_tc_sm100(drows, R, q, ldq, cross_pool_, (size_t)audio_cap_ * cross_slot, (size_t)(ck - cross_pool_), (size_t)(cv - cross_pool_),
                                                 out, d_, hp_.n_text_head, cross_slot, hp_.n_audio_ctx, Ln.sched + ((Ln.cross_seq++ & 1u) << 1), cross_ctas_, Ln.stream, nullptr, &Ln.cross_grp))
            return gemm_fail();
        return true;
    }

    // Queues one chunk of rows on the lane; its sample results land in the lane's pinned staging
    // (finish_chunk picks them up after the stream has drained).
    bool decode_chunk(Lane& Ln, const RowDesc* rows, int R, const int* samp, int S, int row_base, const SampleParams* sp, size_t res_off,
                      float* logits_host, const float* inject = nullptr, const unsigned char* inject_mask = nullptr) {
        const int d = d_, Ld = hp_.n_text_layer, ntc = hp_.n_text_ctx;
        cudaStream_t st = Ln.stream;
        for (int r = 0; r < R; ++r) {
            const RowDesc& rd = rows[r];
            if (rd.token < 0 || rd.token >= hp_.n_vocab || rd.pos < 0 || rd.pos >= ntc || rd.kv_slot < 0 || rd.kv_slot >= kv_cap_ ||
                rd.audio_slot < 0 || rd.audio_slot >= audio_cap_) { set_err("decode: bad row"); return false; }
        }
        // staging: rows | row groups of the cross-attention (R + 1 ints) | sample row indices | sampling parameters | results
        const size_t grp_off = align_up(sizeof(RowDesc) * R, 256);
        const size_t rows_bytes = grp_off + align_up(sizeof(int) * (R + 1), 256), idx_bytes = align_up(sizeof(int) * std::max(S, 1), 256);
        const size_t sp_bytes = align_up(sizeof(SampleParams) * std::max(S, 1), 256), res_bytes = align_up(sizeof(SampleResult) * std::max(S, 1), 256);
        if (!lane_pin(Ln, rows_bytes + idx_bytes + sp_bytes + res_bytes)) return false;
        if (!lane_scratch(Ln, rows_bytes + idx_bytes + sp_bytes + res_bytes)) return false;
        char* hp = Ln.pin;
        memcpy(hp, rows, sizeof(RowDesc) * R);
        {   // rows of one audio that follow each other (a pass and its speculative successor, beams) share their cross-KV panels
            int* hgrp = reinterpret_cast<int*>(hp + grp_off);
            Ln.cross_grp = CrossGroups{};
            const int width = cross_groups_ ? cross_attention_groups(rows, R, hgrp) : 1;
            if (width > 1) { Ln.cross_grp.groups = reinterpret_cast<const int*>(Ln.scratch + grp_off); Ln.cross_grp.n_groups = hgrp[R]; Ln.cross_grp.width = width; }
        }
        int* hidx = reinterpret_cast<int*>(hp + rows_bytes);
        for (int i = 0; i < S; ++i) hidx[i] = samp[i] - row_base;
        if (S) memcpy(hp + rows_bytes + idx_bytes, sp, sizeof(SampleParams) * S);
        const RowDesc* drows = reinterpret_cast<const RowDesc*>(Ln.scratch);
        const int* didx = reinterpret_cast<const int*>(Ln.scratch + rows_bytes);
        const SampleParams* dsp = reinterpret_cast<const SampleParams*>(Ln.scratch + rows_bytes + idx_bytes);
        SampleResult* dres = reinterpret_cast<SampleResult*>(Ln.scratch + rows_bytes + idx_bytes + sp_bytes);

        bool distinct = true;   // single-token steps: one row per KV slot (selects the kernels of the QKV / self-attention step)
        {
            std::vector<int> slots(R);
            for (int r = 0; r < R; ++r) slots[r] = rows[r].audio_slot;
            std::sort(slots.begin(), slots.end());
            Ln.cross_slots = (int)(std::unique(slots.begin(), slots.end()) - slots.begin());
            for (int r = 0; r < R; ++r) slots[r] = rows[r].kv_slot;
            std::sort(slots.begin(), slots.end());
            distinct = std::adjacent_find(slots.begin(), slots.end()) == slots.end();
        }
        const auto host_t0 = std::chrono::steady_clock::now();
        // ---- everything queued on the lane's stream for this chunk.  For small step batches (the single-utterance path: 1-8 rows)
        // the sequence depends on (R, S) and on buffer addresses only — tokens, positions, slots, sampling parameters travel through
        // the pinned staging — so it is captured once as a CUDA graph (programmatic-dependent-launch edges included) and replayed:
        // a round of a 4-layer decoder is ~60 launches, ~0.3 ms of host time that the GPU finishes faster than the host can issue.
        auto issue = [&]() -> bool {
        CUDA_OK(cudaMemcpyAsync(Ln.scratch, hp, rows_bytes + idx_bytes + sp_bytes, cudaMemcpyHostToDevice, st));
        launch_embed<T>(drows, R, tok_emb_, dec_pos_, Ln.x, d, st);
        // head-major KV panels: self [slot][layer][K|V][head][448][64], cross [slot][layer][K|V][head][1536][64]
        const size_t self_head = (size_t)ntc * 64, self_kv = (size_t)ntc * d;
        const size_t self_slot = (size_t)Ld * 2 * self_kv;
        const size_t cross_head = (size_t)kWinRows * 64, cross_kv = (size_t)kWinRows * d;
        const size_t cross_slot = (size_t)Ld * 2 * cross_kv;
        const bool skinny = use_skinny_ && sizeof(T) == 2 && R <= 128 && 4 * d <= 5120;
        if (skinny) {
            const bool chain = use_chain_ && cross_mode_ == 2 && fuse_cross_q_;
            const bool proj = use_proj_ && dec_proj_supported(R, d, d) && dec_proj_supported(R, 3 * d, d) && dec_proj_supported(R, 4 * d, d) &&
                              dec_proj_supported(R, d, 4 * d);
            if (proj) { if (!decode_layers_proj(Ln, drows, R)) return false; }
            else if (!(chain ? decode_layers_chain(Ln, drows, R, distinct) : decode_layers_skinny(Ln, drows, R, distinct))) return false;
        } else {
            for (int l = 0; l < Ld; ++l) {
                const Layer<T>& L = dec_[l];
                launch_layernorm<T>(Ln.x, d, L.ln1_g, L.ln1_b, Ln.y, d, R, d, st);
                { Epilogue e; e.bias = L.bqkv; if (!gemm(Ln.y, d, L.wqkv, d, Ln.qkv, 3 * d, R, 3 * d, d, e, st)) return gemm_fail(); }
                T* kc = self_pool_ + (size_t)l * 2 * self_kv;
                T* vc = kc + self_kv;
                launch_scatter_kv<T>(drows, R, Ln.qkv, kc, vc, self_slot, ntc, d, st);
                launch_dec_attention<T>(drows, R, Ln.qkv, 3 * d, kc, vc, Ln.att, d, hp_.n_text_head, /*cross=*/0, self_slot, self_head, 0, st);
                { Epilogue e; e.bias = L.bo; e.res = Ln.x; e.res_ld = d; if (!gemm(Ln.att, d, L.wo, d, Ln.x, d, R, d, d, e, st)) return gemm_fail(); }
                launch_layernorm<T>(Ln.x, d, L.lnc_g, L.lnc_b, Ln.y, d, R, d, st);
                { Epilogue e; e.bias = L.bcq; if (!gemm(Ln.y, d, L.wcq, d, Ln.qkv, d, R, d, d, e, st)) return gemm_fail(); }
                const T* ck = cross_pool_ + (size_t)l * 2 * cross_kv;
                const T* cv = ck + cross_kv;
                mark_begin(Ln.tm, profiling);
                launch_dec_attention<T>(drows, R, Ln.qkv, d, ck, cv, Ln.att, d, hp_.n_text_head, /*cross=*/1, cross_slot, cross_head, hp_.n_audio_ctx, st);
                mark_end(Ln.tm, profiling, 2);
                if (profiling) Ln.cross_bytes += 2.0 * Ln.cross_slots * hp_.n_audio_ctx * d * sizeof(T);
                { Epilogue e; e.bias = L.bco; e.res = Ln.x; e.res_ld = d; if (!gemm(Ln.att, d, L.wco, d, Ln.x, d, R, d, d, e, st)) return gemm_fail(); }
                launch_layernorm<T>(Ln.x, d, L.ln2_g, L.ln2_b, Ln.y, d, R, d, st);
                { Epilogue e; e.bias = L.b1; e.act = 1; if (!gemm(Ln.y, d, L.w1, d, Ln.h, 4 * d, R, 4 * d, d, e, st)) return gemm_fail(); }
                { Epilogue e; e.bias = L.b2; e.res = Ln.x; e.res_ld = d; if (!gemm(Ln.h, 4 * d, L.w2, 4 * d, Ln.x, d, R, d, 4 * d, e, st)) return gemm_fail(); }
            }
        }
        if (S > 0) {
            mark_begin(Ln.tm, detail_);
            launch_layernorm_gather<T>(Ln.x, d, didx, dec_ln_g_, dec_ln_b_, Ln.ys, d, S, d, st);
            if (!logits_gemm(Ln, S)) return false;
            if (inject && inject_mask) {   // scripted-logits test hook: pageable source, staged by the runtime before the call returns
                for (int i = 0; i < S; ++i)
                    if (inject_mask[i])
                        CUDA_OK(cudaMemcpyAsync(Ln.logits + (size_t)i * hp_.n_vocab, inject + (size_t)i * hp_.n_vocab, sizeof(float) * hp_.n_vocab,
                                                cudaMemcpyHostToDevice, st));
            }
            launch_process_logits(Ln.logits, hp_.n_vocab, dsp, dres, S, vocab_ids, nullptr, Ln.probs, st);
            mark_end(Ln.tm, detail_, 15);
            CUDA_OK(cudaMemcpyAsync(hp + rows_bytes + idx_bytes + sp_bytes, dres, sizeof(SampleResult) * S, cudaMemcpyDeviceToHost, st));
            if (logits_host)
                CUDA_OK(cudaMemcpyAsync(logits_host, Ln.logits, sizeof(float) * (size_t)S * hp_.n_vocab, cudaMemcpyDeviceToHost, st));
        }
        return true;
        };   // issue

        bool queued = false;
        const bool graphable = graph_rows_ > 0 && R <= graph_rows_ && !profiling && !detail_ && !trace_buf_ && !inject && !logits_host;
        if (graphable && Ln.graph_epoch != pool_epoch_) {   // a KV / cross-KV pool moved since these graphs were captured
            for (auto& g : Ln.graphs) cudaGraphExecDestroy(g.second.exec);
            Ln.graphs.clear();
            Ln.graph_seen.clear();
            Ln.graph_epoch = pool_epoch_;
        }
        if (graphable) {
            const uint64_t key = ((uint64_t)R << 40) ^ ((uint64_t)Ln.cross_grp.n_groups << 48) ^ ((uint64_t)Ln.cross_grp.width << 56) ^ ((uint64_t)S << 24) ^ ((uint64_t)(Ln.cross_seq & 1u) << 20) ^ ((uint64_t)(distinct ? 1 : 0) << 21) ^ (uint64_t)(reinterpret_cast<uintptr_t>(Ln.scratch) >> 4) ^
                                 ((uint64_t)(reinterpret_cast<uintptr_t>(Ln.pin) >> 4) << 7);
            auto it = Ln.graphs.find(key);
            if (it != Ln.graphs.end()) {
                CUDA_OK(cudaGraphLaunch(it->second.exec, st));
                count_launches(it->second.n_kernels);
                Ln.cross_seq += (unsigned)hp_.n_text_layer;
                queued = true;
            } else if (++Ln.graph_seen[key] >= 2 && Ln.graphs.size() < 64) {
                // second time this shape comes by (every first-use cudaFuncSetAttribute has happened): capture it
                const long k0 = kernel_launch_count();
                const unsigned seq0 = Ln.cross_seq;
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const bool ok = issue();
                    cudaGraph_t g = nullptr;
                    const cudaError_t e1 = cudaStreamEndCapture(st, &g);
                    cudaGraphExec_t exec = nullptr;
                    if (ok && e1 == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) {
                        Ln.graphs[key] = typename Lane::GraphEntry{exec, kernel_launch_count() - k0};
                        cudaGraphDestroy(g);
                        CUDA_OK(cudaGraphLaunch(exec, st));
                        queued = true;
                    } else {
                        if (g) cudaGraphDestroy(g);
                        cudaGetLastError();
                        Ln.cross_seq = seq0;
                        graph_rows_ = 0;   // capture is not possible here: direct launches from now on
                    }
                } else {
                    cudaGetLastError();
                    graph_rows_ = 0;
                }
            }
        }
        if (!queued && !issue()) return false;
        Ln.pend_S = 0;
        if (S > 0) {
            Ln.pend_S = S;
            Ln.pend_off = res_off;
            Ln.pend_pin_off = rows_bytes + idx_bytes + sp_bytes;
        }
        Ln.issue_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        Ln.last_logit_rows = S;
        Ln.last_logit_base = res_off;
        return true;
    }

    bool lang_probs(int lane, int sample_index, float* probs_host, int* best) override {
        CUDA_OK(cudaSetDevice(device_));
        if (lane < 0 || lane >= (int)lanes_.size()) { set_err("lang_probs: no such lane"); return false; }
        Lane& L = lanes_[lane];
        // only the last chunk's logits are still in the lane's buffer: index relative to it
        const long local = (long)sample_index - (long)L.last_logit_base;
        if (L.inflight || local < 0 || local >= L.last_logit_rows) { set_err("lang_probs: the logits of that sample row are no longer held by the lane"); return false; }
        sample_index = (int)local;
        CUDA_OK(cudaStreamSynchronize(L.stream));
        if (!lane_scratch(L, 1024)) return false;
        float* dp = reinterpret_cast<float*>(L.scratch);
        int* db = reinterpret_cast<int*>(L.scratch + 512);
        launch_lang_probs(L.logits + (size_t)sample_index * hp_.n_vocab, vocab_ids, dp, db, L.stream);
        float hp[kNumLangs];
        CUDA_OK(cudaMemcpyAsync(hp, dp, sizeof(hp), cudaMemcpyDeviceToHost, L.stream));
        CUDA_OK(cudaMemcpyAsync(best, db, sizeof(int), cudaMemcpyDeviceToHost, L.stream));
        CUDA_OK(cudaStreamSynchronize(L.stream));
        if (probs_host) memcpy(probs_host, hp, sizeof(hp));
        return true;
    }

    bool kv_copy(int lane, const std::vector<KvCopy>& pairs) override {
        CUDA_OK(cudaSetDevice(device_));
        if (pairs.empty()) return true;
        if (lane < 0 || lane >= (int)lanes_.size()) { set_err("kv_copy: no such lane"); return false; }
        Lane& L = lanes_[lane];
        if (L.inflight) { set_err("kv_copy: lane is busy"); return false; }
        const size_t bytes = sizeof(KvCopy) * pairs.size();
        if (!lane_pin(L, bytes) || !lane_scratch(L, bytes)) return false;
        CUDA_OK(cudaStreamSynchronize(L.stream));
        memcpy(L.pin, pairs.data(), bytes);
        CUDA_OK(cudaMemcpyAsync(L.scratch, L.pin, bytes, cudaMemcpyHostToDevice, L.stream));
        const int d = d_, ntc = hp_.n_text_ctx, Ld = hp_.n_text_layer;
        launch_kv_copy<T>(reinterpret_cast<const KvCopy*>(L.scratch), (int)pairs.size(), self_pool_, (size_t)Ld * 2 * ntc * d,
                          2 * Ld * hp_.n_text_head, (size_t)ntc * 64, L.stream);
        CUDA_OK(cudaStreamSynchronize(L.stream));
        return true;
    }

    bool process_logits_host(const float* logits, const SampleParams& sp, SampleResult& out, float* logprobs, float* probs) override {
        CUDA_OK(cudaSetDevice(device_));
        const int nv = hp_.n_vocab;
        const size_t lbytes = align_up(sizeof(float) * nv, 256);
        if (!ensure_dev_scratch(3 * lbytes + 1024)) return false;
        float* dl = reinterpret_cast<float*>(dev_scratch_);
        float* dlp = reinterpret_cast<float*>(dev_scratch_ + lbytes);
        float* dpr = reinterpret_cast<float*>(dev_scratch_ + 2 * lbytes);
        SampleParams* dsp = reinterpret_cast<SampleParams*>(dev_scratch_ + 3 * lbytes);
        SampleResult* dres = reinterpret_cast<SampleResult*>(dev_scratch_ + 3 * lbytes + 256);
        CUDA_OK(cudaMemcpyAsync(dl, logits, sizeof(float) * nv, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaMemcpyAsync(dsp, &sp, sizeof(sp), cudaMemcpyHostToDevice, stream_));
        launch_process_logits(dl, nv, dsp, dres, 1, vocab_ids, dlp, dpr, stream_);
        CUDA_OK(cudaMemcpyAsync(&out, dres, sizeof(out), cudaMemcpyDeviceToHost, stream_));
        if (logprobs) CUDA_OK(cudaMemcpyAsync(logprobs, dlp, sizeof(float) * nv, cudaMemcpyDeviceToHost, stream_));
        if (probs) CUDA_OK(cudaMemcpyAsync(probs, dpr, sizeof(float) * nv, cudaMemcpyDeviceToHost, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
        CUDA_OK(cudaGetLastError());
        return true;
    }

    bool event_record(int slot) override {
        CUDA_OK(cudaSetDevice(device_));
        if (slot < 0 || slot >= 8) return false;
        if (!user_ev_[slot]) CUDA_OK(cudaEventCreate(&user_ev_[slot]));
        CUDA_OK(cudaEventRecord(user_ev_[slot], stream_));
        return true;
    }
    double event_elapsed_ms(int a, int b) override {
        if (a < 0 || a >= 8 || b < 0 || b >= 8 || !user_ev_[a] || !user_ev_[b]) return -1.0;
        cudaSetDevice(device_);
        if (cudaEventSynchronize(user_ev_[b]) != cudaSuccess) return -1.0;
        float ms = -1.0f;
        if (cudaEventElapsedTime(&ms, user_ev_[a], user_ev_[b]) != cudaSuccess) return -1.0;
        return ms;
    }

    // ------------------------------------------------------------------ inspection
    bool export_mel(const DeviceMel& m, float* out) override {
        CUDA_OK(cudaSetDevice(device_));
        if (!m.raw) { set_err("no mel"); return false; }
        const size_t n = (size_t)m.n_len * m.n_mel;
        if (!ensure_dev_scratch(n * sizeof(float))) return false;
        launch_export_mel(m.raw, m.max_key, m.n_frames, m.n_len, m.n_mel, reinterpret_cast<float*>(dev_scratch_), stream_);
        CUDA_OK(cudaMemcpyAsync(out, dev_scratch_, n * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
        return true;
    }
    bool export_encoder_output(int slot, float* out) override {
        CUDA_OK(cudaSetDevice(device_));
        if (slot < 0 || slot >= audio_cap_) { set_err("bad slot"); return false; }
        const int n = hp_.n_audio_ctx;
        if (!ensure_dev_scratch((size_t)n * d_ * sizeof(float))) return false;
        launch_convert_2d<T, float>(encout_pool_ + (size_t)slot * kWinRows * d_, d_, reinterpret_cast<float*>(dev_scratch_), d_, n, d_, stream_);
        CUDA_OK(cudaMemcpyAsync(out, dev_scratch_, (size_t)n * d_ * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
        return true;
    }
    bool export_cross_kv(int slot, int layer, float* k, float* v) override {
        CUDA_OK(cudaSetDevice(device_));
        if (slot < 0 || slot >= audio_cap_ || layer < 0 || layer >= hp_.n_text_layer) { set_err("bad slot/layer"); return false; }
        const int n = hp_.n_audio_ctx;
        const size_t ldx = (size_t)2 * hp_.n_text_layer * d_;
        const size_t bytes = (size_t)n * d_ * sizeof(float);
        if (!ensure_dev_scratch(2 * bytes)) return false;
        const size_t cross_kv = (size_t)kWinRows * d_;
        const T* base = cross_pool_ + (size_t)slot * kWinRows * ldx + (size_t)layer * 2 * cross_kv;
        float* dk = reinterpret_cast<float*>(dev_scratch_);
        float* dv = reinterpret_cast<float*>(dev_scratch_ + bytes);
        for (int h = 0; h < hp_.n_text_head; ++h) {  // head-major panels -> [pos][d]
            launch_convert_2d<T, float>(base + (size_t)h * kWinRows * 64, 64, dk + h * 64, d_, n, 64, stream_);
            launch_convert_2d<T, float>(base + cross_kv + (size_t)h * kWinRows * 64, 64, dv + h * 64, d_, n, 64, stream_);
        }
        CUDA_OK(cudaMemcpyAsync(k, dk, bytes, cudaMemcpyDeviceToHost, stream_));
        CUDA_OK(cudaMemcpyAsync(v, dv, bytes, cudaMemcpyDeviceToHost, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
        return true;
    }

private:
    bool gemm_fail() {
        set_err(std::string("GEMM launch failed: ") + sm100_last_error(), /*keep_first=*/true);
        return false;
    }
    bool ensure_pin(size_t bytes) {
        if (bytes <= pin_cap_) return true;
        if (pin_) { CUDA_OK(cudaStreamSynchronize(stream_)); CUDA_OK(cudaFreeHost(pin_)); pin_ = nullptr; }
        pin_cap_ = align_up(bytes * 2, 1 << 16);
        CUDA_OK(cudaMallocHost(&pin_, pin_cap_));
        return true;
    }
    bool ensure_dev_scratch(size_t bytes) {
        if (bytes <= dev_scratch_cap_) return true;
        if (dev_scratch_) { CUDA_OK(cudaStreamSynchronize(stream_)); CUDA_OK(cudaFree(dev_scratch_)); dev_scratch_ = nullptr; }
        dev_scratch_cap_ = align_up(bytes * 2, 1 << 16);
        CUDA_OK(cudaMalloc(&dev_scratch_, dev_scratch_cap_));
        return true;
    }
    bool take_cached_mel(DeviceMel& m, size_t need) {
        for (size_t i = 0; i < mel_cache_.size(); ++i)
            if (mel_cache_[i].raw_cap >= need) {
                const DeviceMel c = mel_cache_[i];
                mel_cache_.erase(mel_cache_.begin() + i);
                m.raw = c.raw; m.max_key = c.max_key; m.raw_cap = c.raw_cap;
                return true;
            }
        return false;
    }
    template <typename P>
    bool grow_pool(P*& pool, int old_cap, int new_cap, size_t slot_elems) {
        P* np = nullptr;
        CUDA_OK(cudaDeviceSynchronize());  // every lane may be reading the pool
        CUDA_OK(cudaMalloc(&np, (size_t)new_cap * slot_elems * sizeof(P)));
        if (pool && old_cap > 0) CUDA_OK(cudaMemcpy(np, pool, (size_t)old_cap * slot_elems * sizeof(P), cudaMemcpyDeviceToDevice));
        if (pool) CUDA_OK(cudaFree(pool));
        pool = np;
        ++pool_epoch_;   // captured step rounds hold the old pool's address: they are dropped (decode_chunk)
        return true;
    }
    bool grow_audio(int new_cap) {
        const size_t cross_slot = (size_t)kWinRows * 2 * hp_.n_text_layer * d_;
        if (!grow_pool(cross_pool_, audio_cap_, new_cap, cross_slot)) return false;
        if (!grow_pool(encout_pool_, audio_cap_, new_cap, (size_t)kWinRows * d_)) return false;
        for (int s = new_cap - 1; s >= audio_cap_; --s) audio_free_.push_back(s);
        audio_cap_ = new_cap;
        return true;
    }
    bool grow_kv(int new_cap) {
        const size_t slot = (size_t)hp_.n_text_layer * 2 * hp_.n_text_ctx * d_;
        if (!grow_pool(self_pool_, kv_cap_, new_cap, slot)) return false;
        for (int s = new_cap - 1; s >= kv_cap_; --s) kv_free_.push_back(s);
        kv_cap_ = new_cap;
        return true;
    }

    // weights ------------------------------------------------------------------------------
    struct Arena {
        char* base = nullptr;
        size_t cap = 0, used = 0;
        void* take(size_t bytes) {
            used = align_up(used, 256);
            void* p = base ? base + used : nullptr;
            used += bytes;
            return p;
        }
    };
    template <typename P>
    P* put(Arena& a, const std::vector<float>& host) {  // upload as fp32 then convert to P on the device
        P* dst = reinterpret_cast<P*>(a.take(host.size() * sizeof(P)));
        if (!a.base) return dst;  // sizing pass
        upload_ok_ = upload_ok_ && upload_convert(host.data(), host.size(), dst);
        return dst;
    }
    bool upload_convert(const float* src, size_t n, float* dst) {
        CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
        return true;
    }
    bool upload_convert(const float* src, size_t n, bf16* dst) {
        const size_t chunk = (size_t)32 << 20;  // floats
        if (!ensure_dev_scratch(std::min(n, chunk) * sizeof(float))) return false;
        for (size_t o = 0; o < n; o += chunk) {
            const size_t c = std::min(chunk, n - o);
            CUDA_OK(cudaMemcpyAsync(dev_scratch_, src + o, c * sizeof(float), cudaMemcpyHostToDevice, stream_));
            launch_convert<float, bf16>(reinterpret_cast<const float*>(dev_scratch_), dst + o, c, stream_);
            CUDA_OK(cudaStreamSynchronize(stream_));
        }
        return true;
    }

    static std::vector<float> concat(std::initializer_list<const std::vector<float>*> parts) {
        std::vector<float> out;
        for (auto* p : parts) out.insert(out.end(), p->begin(), p->end());
        return out;
    }

    void build_weights(const HostModel& hm, Arena& a) {
        const int d = d_, nm = hp_.n_mels;
        auto W = [&](const std::string& n) -> const std::vector<float>& { return hm.get(n).data; };
        const std::vector<float> zeros_d(d, 0.0f);
        // conv weights [o][c][k] -> [o][(k, c)] so a row of the im2col-free operand (3 consecutive
        // time steps x channels) lines up with a weight row
        auto reorder_conv = [&](const std::vector<float>& w, int cin) {
            std::vector<float> r((size_t)d * 3 * cin);
            for (int o = 0; o < d; ++o)
                for (int c = 0; c < cin; ++c)
                    for (int k = 0; k < 3; ++k) r[((size_t)o * 3 + k) * cin + c] = w[((size_t)o * cin + c) * 3 + k];
            return r;
        };
        conv1_w_ = put<T>(a, reorder_conv(W("encoder.conv1.weight"), nm));
        conv1_b_ = put<float>(a, W("encoder.conv1.bias"));
        conv2_w_ = put<T>(a, reorder_conv(W("encoder.conv2.weight"), d));
        conv2_b_ = put<float>(a, W("encoder.conv2.bias"));
        {
            std::vector<float> pe((size_t)kWinRows * d, 0.0f);
            const auto& src = W("encoder.positional_embedding");
            std::copy(src.begin(), src.end(), pe.begin());
            enc_pos_ = put<float>(a, pe);
        }
        auto attn = [&](const std::string& p, Layer<T>& L, bool self) {
            if (self) {
                L.ln1_g = put<float>(a, W(p + "_ln.weight"));
                L.ln1_b = put<float>(a, W(p + "_ln.bias"));
                L.wqkv = put<T>(a, concat({&W(p + ".query.weight"), &W(p + ".key.weight"), &W(p + ".value.weight")}));
                L.bqkv = put<float>(a, concat({&W(p + ".query.bias"), &zeros_d, &W(p + ".value.bias")}));
                L.wo = put<T>(a, W(p + ".out.weight"));
                L.bo = put<float>(a, W(p + ".out.bias"));
            } else {
                L.lnc_g = put<float>(a, W(p + "_ln.weight"));
                L.lnc_b = put<float>(a, W(p + "_ln.bias"));
                L.wcq = put<T>(a, W(p + ".query.weight"));
                L.bcq = put<float>(a, W(p + ".query.bias"));
                L.wco = put<T>(a, W(p + ".out.weight"));
                L.bco = put<float>(a, W(p + ".out.bias"));
            }
        };
        auto mlp = [&](const std::string& p, Layer<T>& L) {
            L.ln2_g = put<float>(a, W(p + "mlp_ln.weight"));
            L.ln2_b = put<float>(a, W(p + "mlp_ln.bias"));
            L.w1 = put<T>(a, W(p + "mlp.0.weight"));
            L.b1 = put<float>(a, W(p + "mlp.0.bias"));
            L.w2 = put<T>(a, W(p + "mlp.2.weight"));
            L.b2 = put<float>(a, W(p + "mlp.2.bias"));
        };
        enc_.assign(hp_.n_audio_layer, Layer<T>());
        for (int i = 0; i < hp_.n_audio_layer; ++i) {
            const std::string p = "encoder.blocks." + std::to_string(i) + ".";
            attn(p + "attn", enc_[i], true);
            mlp(p, enc_[i]);
        }
        enc_lnp_g_ = put<float>(a, W("encoder.ln_post.weight"));
        enc_lnp_b_ = put<float>(a, W("encoder.ln_post.bias"));
        dec_.assign(hp_.n_text_layer, Layer<T>());
        std::vector<float> cw, cb;
        if (a.base) { cw.reserve((size_t)2 * hp_.n_text_layer * d * d); cb.reserve((size_t)2 * hp_.n_text_layer * d); }
        for (int i = 0; i < hp_.n_text_layer; ++i) {
            const std::string p = "decoder.blocks." + std::to_string(i) + ".";
            attn(p + "attn", dec_[i], true);
            attn(p + "cross_attn", dec_[i], false);
            mlp(p, dec_[i]);
            if (a.base) {
                const auto& k = W(p + "cross_attn.key.weight");
                const auto& v = W(p + "cross_attn.value.weight");
                cw.insert(cw.end(), k.begin(), k.end());
                cw.insert(cw.end(), v.begin(), v.end());
                cb.insert(cb.end(), zeros_d.begin(), zeros_d.end());
                const auto& vb = W(p + "cross_attn.value.bias");
                cb.insert(cb.end(), vb.begin(), vb.end());
            }
        }
        if (!a.base) { cw.resize((size_t)2 * hp_.n_text_layer * d * d); cb.resize((size_t)2 * hp_.n_text_layer * d); }
        cross_w_ = put<T>(a, cw);
        cross_b_ = put<float>(a, cb);
        dec_ln_g_ = put<float>(a, W("decoder.ln.weight"));
        dec_ln_b_ = put<float>(a, W("decoder.ln.bias"));
        dec_pos_ = put<float>(a, W("decoder.positional_embedding"));
        tok_emb_ = put<T>(a, W("decoder.token_embedding.weight"));
    }

    bool upload_weights(const HostModel& hm) {
        try {
            Arena sizing;
            build_weights(hm, sizing);
            Arena a;
            a.cap = align_up(sizing.used, 256) + 256;
            CUDA_OK(cudaMalloc(&a.base, a.cap));
            owned_.push_back(a.base);
            upload_ok_ = true;
            build_weights(hm, a);
            if (!upload_ok_) return false;
            weight_bytes_ = a.used;
        } catch (const std::exception& e) {
            set_err(e.what());
            return false;
        }
        // mel tables
        std::vector<float> hann(kNFft);
        std::vector<float2> tw(kNFft);
        for (int i = 0; i < kNFft; ++i) {
            const double th = 2.0 * M_PI * i / kNFft;
            hann[i] = (float)(0.5 * (1.0 - cos(th)));
            tw[i] = make_float2((float)cos(th), (float)sin(th));
        }
        std::vector<int2> ranges(hp_.n_mels);
        for (int m = 0; m < hp_.n_mels; ++m) {
            int lo = kNFreq, hi = 0;
            for (int k = 0; k < kNFreq; ++k)
                if (hm.filters[(size_t)m * kNFreq + k] != 0.0f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
            if (lo > hi) lo = hi = 0;
            ranges[m] = make_int2(lo, hi);
        }
        float *dh, *df; float2* dt; int2* dr;
        CUDA_OK(cudaMalloc(&dh, sizeof(float) * kNFft)); owned_.push_back(dh);
        CUDA_OK(cudaMalloc(&dt, sizeof(float2) * kNFft)); owned_.push_back(dt);
        CUDA_OK(cudaMalloc(&df, sizeof(float) * hm.filters.size())); owned_.push_back(df);
        CUDA_OK(cudaMalloc(&dr, sizeof(int2) * ranges.size())); owned_.push_back(dr);
        CUDA_OK(cudaMemcpy(dh, hann.data(), sizeof(float) * kNFft, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dt, tw.data(), sizeof(float2) * kNFft, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(df, hm.filters.data(), sizeof(float) * hm.filters.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dr, ranges.data(), sizeof(int2) * ranges.size(), cudaMemcpyHostToDevice));
        mel_tables_ = MelTables{dh, dt, df, dr, hp_.n_mels};
        return true;
    }

    bool alloc_workspace() {
        const int d = d_, nm = hp_.n_mels;
        const bool f32 = sizeof(T) == 4;
        enc_batch_ = std::max(1, env_int("NOBS_WHISPER_ENC_BATCH", f32 ? 2 : 24));   // 24 windows: M = 36864 rows per GEMM (measured: 8 -> 907 ms, 12 -> 883, 24 -> 853 per 120 windows)
        dec_rows_ = std::max(64, env_int("NOBS_WHISPER_DEC_ROWS", 4096));
        dec_samples_ = std::max(8, env_int("NOBS_WHISPER_DEC_SAMPLES", 1024));
        use_skinny_ = env_int("NOBS_WHISPER_SKINNY", 1) != 0;
        detail_ = env_int("NOBS_WHISPER_PROFILE_DECODE", 0) != 0;
        cross_mode_ = env_int("NOBS_WHISPER_CROSS_MODE", 2);
        fuse_cross_q_ = env_int("NOBS_WHISPER_FUSE_CROSS_Q", 1) != 0;
        skinny_logits_ = env_int("NOBS_WHISPER_SKINNY_LOGITS", 1) != 0;
        fuse_qkv_ = env_int("NOBS_WHISPER_FUSE_QKV", 1) != 0;
        cross_ctas_ = env_int("NOBS_WHISPER_CROSS_CTAS", 0);
        cross_groups_ = env_int("NOBS_WHISPER_CROSS_GROUPS", 1) != 0;
        const int n_lanes = std::min(8, std::max(1, env_int("NOBS_WHISPER_LANES", f32 ? 1 : 2)));
        // with several lanes the step GEMMs run a 2-deep ring (49 KB at 64 rows): two of them fit next to the two
        // attention CTAs (2 x 57 KB) an SM holds for another lane
        set_skinny_gemm_stages(env_int("NOBS_WHISPER_SKINNY_STAGES", n_lanes > 1 ? 2 : 3));
        // fused projection chains need every lane's chain grid co-resident (decode_chain_sm100.cu); otherwise the
        // multi-launch path runs
        chain_stages_ = env_int("NOBS_WHISPER_CHAIN_STAGES", chain_stages_for_lanes(n_lanes));
        // Off by default: measured on B200 (profiles/r2_chain_*.txt) a device-wide barrier costs 1.9 us — the same as a PDL launch
        // boundary — and the chain needs two per projection, so the fused grid is ~5 % SLOWER than twelve small launches.
        use_chain_ = env_int("NOBS_WHISPER_CHAIN", 0) != 0 && !f32 && chain_fits(n_lanes, chain_stages_);
        // Opt-in: one cluster launch per projection (decode_proj_sm100.cu), 8 launches per layer.  Measured slower in the 2-3 lane step than
        // the split-K GEMM + epilogue pairs (profiles/README.md, round 2): each fused kernel has the serial phases of two, and an
        // 8-CTA cluster is placed later than single CTAs while another lane's attention CTAs hold every SM.
        use_proj_ = env_int("NOBS_WHISPER_PROJ", 0) != 0 && !f32 && cross_mode_ == 2;
        fc1_fused_ = env_int("NOBS_WHISPER_FC1_FUSED", 1) != 0 && !f32;
        graph_rows_ = f32 ? 0 : std::max(0, env_int("NOBS_WHISPER_GRAPH_ROWS", 8));
        // Under a CUDA tool (Nsight Compute / Systems, compute-sanitizer) rounds are launched directly: capturing the programmatic-
        // dependent-launch sequence inside ncu's injection aborted the process ("free(): invalid pointer", observed with ncu 2025.2 on
        // smoke()), and a launch list of individual kernels is what a profile of this engine is taken for anyway.
        if (graph_rows_ > 0 && cuda_tool_attached()) graph_rows_ = 0;
        // The encoder and every decode lane own their activations: a lane may decode while the encoder
        // works on other windows and while other lanes decode.
        auto plan_enc = [&](Arena& a) {
            const size_t M = (size_t)enc_batch_ * kWinRows;
            e_mel_ = (T*)a.take(((size_t)enc_batch_ * kWinRowsIn + 2) * nm * sizeof(T));
            e_h1_ = (T*)a.take(((size_t)enc_batch_ * kWinRowsIn + 1) * d * sizeof(T));
            e_x_ = (float*)a.take(M * d * sizeof(float));
            e_y_ = (T*)a.take(M * d * sizeof(T));
            e_qkv_ = (T*)a.take(M * 3 * d * sizeof(T));
            e_att_ = (T*)a.take(M * d * sizeof(T));
            e_h_ = (T*)a.take(M * 4 * d * sizeof(T));
        };
        auto plan_dec = [&](Arena& a, Lane& L) {
            const size_t R = dec_rows_, S = dec_samples_;
            L.x = (float*)a.take(R * d * sizeof(float));
            L.y = (T*)a.take(R * d * sizeof(T));
            L.qkv = (T*)a.take(R * 3 * d * sizeof(T));
            L.att = (T*)a.take(R * d * sizeof(T));
            L.h = (T*)a.take(R * 4 * d * sizeof(T));
            L.partial = (float*)a.take((size_t)16 << 20);  // skinny-GEMM split-K partials: <= 4 splits x 128 rows x 5120 cols (FC1) fp32
            L.sched = (int*)a.take(256);
            L.bar = (unsigned int*)a.take(256);
            L.stats = (float2*)a.take(40 * 128 * sizeof(float2));
            L.ticket = (int*)a.take(256);
            L.ys = (T*)a.take(S * d * sizeof(T));
            L.logits = (float*)a.take(S * (size_t)hp_.n_vocab * sizeof(float));
            L.probs = (float*)a.take(S * (size_t)hp_.n_vocab * sizeof(float));
        };
        lanes_.assign(n_lanes, Lane());
        Arena sizing;
        plan_enc(sizing);
        for (auto& L : lanes_) plan_dec(sizing, L);
        const size_t bytes = sizing.used + 512;
        char* base = nullptr;
        CUDA_OK(cudaMalloc(&base, bytes));
        owned_.push_back(base);
        CUDA_OK(cudaMemset(base, 0, bytes));
        Arena a; a.base = base; a.cap = bytes;
        plan_enc(a);
        for (auto& L : lanes_) {
            plan_dec(a, L);
            {   // NOBS_WHISPER_DEC_PRIORITY=1: the lanes' (latency-bound) kernels are placed before pending encoder CTAs
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);
                if (env_int("NOBS_WHISPER_DEC_PRIORITY", 0) && hi < lo) CUDA_OK(cudaStreamCreateWithPriority(&L.stream, cudaStreamNonBlocking, hi));
                else CUDA_OK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
            }
            for (auto& e : L.ev) CUDA_OK(cudaEventCreate(&e));
            L.tm.s = L.stream;
        }
        main_marks_.s = stream_;
        ws_base_ = base;
        ws_bytes_ = bytes;
        return true;
    }

    // the conv operands rely on zero rows: rows 0 / last of e_mel_ (the last one moves with the batch size)
    // and row 0 of e_h1_; they are re-zeroed at the start of every encode batch.
public:
    bool rezero_conv_pads(int nb) {
        const int d = d_, nm = hp_.n_mels;
        CUDA_OK(cudaMemsetAsync(e_mel_, 0, (size_t)nm * sizeof(T), stream_));
        CUDA_OK(cudaMemsetAsync(e_mel_ + ((size_t)nb * kWinRowsIn + 1) * nm, 0, (size_t)nm * sizeof(T), stream_));
        CUDA_OK(cudaMemsetAsync(e_h1_, 0, (size_t)d * sizeof(T), stream_));
        return true;
    }

private:
    HParams hp_;
    int d_ = 0;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    std::vector<void*> owned_;
    size_t weight_bytes_ = 0;
    bool upload_ok_ = true;
    const int init_key_ = INT_MIN;

    // weights
    T *conv1_w_ = nullptr, *conv2_w_ = nullptr, *cross_w_ = nullptr, *tok_emb_ = nullptr;
    float *conv1_b_ = nullptr, *conv2_b_ = nullptr, *enc_pos_ = nullptr, *cross_b_ = nullptr, *dec_pos_ = nullptr;
    float *enc_lnp_g_ = nullptr, *enc_lnp_b_ = nullptr, *dec_ln_g_ = nullptr, *dec_ln_b_ = nullptr;
    std::vector<Layer<T>> enc_, dec_;
    MelTables mel_tables_{};

    // pools
    T *cross_pool_ = nullptr, *self_pool_ = nullptr, *encout_pool_ = nullptr;
    int audio_cap_ = 0, kv_cap_ = 0;
    std::vector<int> audio_free_, kv_free_;
    std::vector<DeviceMel> mel_cache_;

    // workspaces
    int enc_batch_ = 1, dec_rows_ = 0, dec_samples_ = 0;
    char* ws_base_ = nullptr;
    size_t ws_bytes_ = 0;
    T *e_mel_ = nullptr, *e_h1_ = nullptr, *e_y_ = nullptr, *e_qkv_ = nullptr, *e_att_ = nullptr, *e_h_ = nullptr;
    float* e_x_ = nullptr;
    std::vector<Lane> lanes_;
    Marks main_marks_;
    bool use_skinny_ = true;
    int cross_mode_ = 2;              // bf16 step rows: 2 tcgen05 streaming cross-attention, 1 SIMT streaming (cp.async.bulk ring), 0 block-per-head SIMT
    int cross_ctas_ = 0;              // > 0: cap that kernel's grid
    unsigned pool_epoch_ = 0;         // bumped whenever a slot pool is reallocated
    int graph_rows_ = 8;              // step batches of at most this many rows are replayed as CUDA graphs (0: off)
    bool fc1_fused_ = true;           // step batches: FC1 + GELU as one cluster launch instead of split-K GEMM + epilogue kernel
    bool use_proj_ = false;           // step batches: fused cluster projections (split-K reduced in DSMEM, LayerNorm split around the kernel boundary)
    bool use_chain_ = true;           // step batches: projection chains between the attention kernels as single persistent launches
    int chain_stages_ = 3;
    bool fuse_qkv_ = true;            // single-token steps: the self-attention kernel finishes the QKV projection's split-K sums
    bool skinny_logits_ = true;       // step batches: logits through the swap-AB weight-streaming GEMM
    bool cross_groups_ = true;        // consecutive rows of one audio are one cross-attention work item (their panels are streamed once)
    bool fuse_cross_q_ = true;        // the tcgen05 cross-attention sums the query projection's split-K partials itself
    bool detail_ = false;             // NOBS_WHISPER_PROFILE_DECODE=1: per-kernel-class event timing of decoder steps
    std::mutex enc_mu_, stats_mu_, err_mu_;
    static constexpr int kEncRing = 64, kEncPin = 16;
    EncTicket enc_ring_[kEncRing];
    long enc_issued_ = 0, enc_done_ = 0;      // tickets handed out / known complete (enc_mu_)
    PackJob* enc_pin_ = nullptr;              // kEncPin tables of enc_batch_ window descriptors
    cudaEvent_t enc_pin_ev_[kEncPin] = {};
    unsigned long enc_pin_next_ = 0;
    void set_err(const std::string& e, bool keep_first = false) {
        std::lock_guard<std::mutex> lock(err_mu_);
        if (!keep_first || err_.empty()) err_ = e;
    }
    double host_issue_ms_ = 0, host_wait_ms_ = 0;  // decode_chunk: time spent issuing launches vs waiting for the GPU
    double detail_ms_[8] = {};
    long detail_n_[8] = {};
    static constexpr int kCrossSample = 8;  // profiling: time the cross-attention of every 8th layer

    cudaEvent_t user_ev_[8] = {};
    unsigned long long* trace_buf_ = nullptr;
    unsigned trace_cap_ = 0;
    std::string trace_path_;
    char* pin_ = nullptr;
    size_t pin_cap_ = 0;
    char* dev_scratch_ = nullptr;
    size_t dev_scratch_cap_ = 0;
    float* pcm_dev_ = nullptr;
    size_t pcm_cap_ = 0;
};

}  // namespace

Engine* Engine::create(const HostModel& hm, int device, Precision prec, std::string& err) {
    if (prec == Precision::FP32) {
        auto* e = new EngineT<float>(hm, device, prec);
        if (!e->init(hm)) { err = e->last_error(); delete e; return nullptr; }
        return e;
    }
    auto* e = new EngineT<bf16>(hm, device, prec);
    if (!e->init(hm)) { err = e->last_error(); delete e; return nullptr; }
    return e;
}

}  // namespace nobs
