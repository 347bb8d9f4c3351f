// =====================================================================================
// whisper_oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference path).
//
// This file is the parity oracle for the B200 engine.  It is NOT part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load it.  The product library (nobs_whisper_b200/libnobswhisper_b200.so) never links,
// loads or calls anything in oracle/.
//
// What it restates: the path nobs-whisper reaches through whisper-rs
//   reference src-tauri/src/whisper.rs:36-52   WhisperContext::new_with_params  -> wo_load
//   reference src-tauri/src/whisper.rs:83-85   ctx.create_state()               -> (state is part of wo_ctx)
//   reference src-tauri/src/whisper.rs:88-124  FullParams                       -> wo_params
//   reference src-tauri/src/whisper.rs:127-129 state.full(params, audio)        -> wo_full
//   reference src-tauri/src/whisper.rs:132-141 full_n_segments / get_segment    -> wo_n_segments / wo_segment_text
// The arithmetic behind those calls lives in third-party crates that are NOT vendored in
// /root/reference and not present offline: whisper-rs 0.15.1 -> whisper-rs-sys 0.14.1
// (reference src-tauri/Cargo.lock:5642-5660), which wraps whisper.cpp/ggml (v1.7.x era).
// The algorithm below restates whisper.cpp's published algorithm (whisper_full_with_state,
// log_mel_spectrogram, whisper_encode_internal, whisper_decode_internal,
// whisper_process_logits, whisper_sample_token, whisper_sequence_score, tokenize) as
// summarised in SURVEY.md §8a rows a1, a4-a11, in "ideal fp32" numerics by default (no f16 GELU table,
// no f16 im2col / KV rounding); wo_set_ggml_faithful() switches those roundings on (see wo_ctx::faithful).
//
// PARITY UNPINNED: the reference's own tests hold no golden vector for mel / encoder /
// logits / tokens (only whisper.rs:272-305 pins NoModel + the hallucination
// filter).  The restatement is cross-checked against an independent implementation of the
// same published model (HF transformers Whisper, tests/golden/) — that validates the
// arithmetic, not whisper.cpp's control flow.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <regex>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int SAMPLE_RATE = 16000;
constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int CHUNK_SIZE = 30;
constexpr int N_FREQ = 201;

// ------------------------------------------------------------------ language table
// whisper.cpp g_lang order (id == index).  Reference UI exposes 8 of these
// (reference src/routes/+page.svelte:49-58); config default "auto" (config.rs:49).
const char* const LANGS[] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id",
    "hi", "fi", "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg",
    "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br",
    "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn", "yo", "so",
    "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt",
    "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"};
constexpr int N_LANGS = sizeof(LANGS) / sizeof(LANGS[0]);

int lang_id(const char* s) {
    for (int i = 0; i < N_LANGS; ++i)
        if (!strcmp(LANGS[i], s)) return i;
    return -1;
}

// ------------------------------------------------------------------ model
struct HParams {
    int n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
};

struct Vocab {
    int n_vocab = 0;
    std::vector<std::string> id_to_token;
    std::map<std::string, int> token_to_id;
    int token_eot = 50256, token_sot = 50257, token_translate = 50357, token_transcribe = 50358;
    int token_solm = 50359, token_prev = 50360, token_nosp = 50361, token_not = 50362, token_beg = 50363;
    bool is_multilingual() const { return n_vocab >= 51865; }
    int num_languages() const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
    int token_lang(int id) const { return token_sot + 1 + id; }
};

typedef std::vector<float> Vec;

struct Attn {
    Vec ln_w, ln_b, q_w, q_b, k_w, v_w, v_b, o_w, o_b;
};
struct Block {
    Attn attn, cross;
    Vec mlp_ln_w, mlp_ln_b, fc1_w, fc1_b, fc2_w, fc2_b;
};

struct Model {
    HParams hp;
    Vec filters;  // [n_mels][201]
    Vocab vocab;
    Vec e_pe, e_conv1_w, e_conv1_b, e_conv2_w, e_conv2_b, e_ln_w, e_ln_b;
    std::vector<Block> enc;
    Vec d_pe, d_te, d_ln_w, d_ln_b;
    std::vector<Block> dec;
};

float f16_to_f32(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, o;
    if (e == 0) {
        if (m == 0) o = s << 31;
        else {
            int sh = 0;
            while (!(m & 1024)) { m <<= 1; ++sh; }
            m &= 1023;
            o = (s << 31) | ((127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 31) o = (s << 31) | 0x7f800000u | (m << 13);
    else o = (s << 31) | ((e - 15 + 127) << 23) | (m << 13);
    float f;
    memcpy(&f, &o, 4);
    return f;
}

// restates whisper_model_load
bool load_model(const char* path, Model& m, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = "cannot open file"; return false; }
    auto rd = [&](void* p, size_t n) { return fread(p, 1, n, f) == n; };
    uint32_t magic = 0;
    if (!rd(&magic, 4) || magic != 0x67676d6c) { err = "bad magic"; fclose(f); return false; }
    if (!rd(&m.hp, sizeof(HParams))) { err = "bad hparams"; fclose(f); return false; }
    int32_t n_mel, n_fft;
    rd(&n_mel, 4); rd(&n_fft, 4);
    if (n_mel != m.hp.n_mels || n_fft != N_FREQ) { err = "bad filterbank"; fclose(f); return false; }
    m.filters.resize((size_t)n_mel * n_fft);
    rd(m.filters.data(), m.filters.size() * 4);
    int32_t n_tok = 0;
    rd(&n_tok, 4);
    Vocab& v = m.vocab;
    v.n_vocab = m.hp.n_vocab;
    v.id_to_token.resize(n_tok);
    for (int i = 0; i < n_tok; ++i) {
        uint32_t len;
        if (!rd(&len, 4) || len > (1u << 20)) { err = "bad vocab"; fclose(f); return false; }
        std::string w(len, '\0');
        if (len) rd(&w[0], len);
        v.id_to_token[i] = w;
        v.token_to_id[w] = i;
    }
    if (v.is_multilingual()) {
        v.token_eot++; v.token_sot++;
        const int dt = v.num_languages() - 98;
        v.token_translate += dt; v.token_transcribe += dt; v.token_solm += dt; v.token_prev += dt;
        v.token_nosp += dt; v.token_not += dt; v.token_beg += dt;
    }
    for (int i = n_tok; i < v.n_vocab; ++i) {
        std::string w;
        if (i > v.token_beg) w = "[_TT_" + std::to_string(i - v.token_beg) + "]";
        else if (i == v.token_eot) w = "[_EOT_]";
        else if (i == v.token_sot) w = "[_SOT_]";
        else if (i == v.token_translate) w = "[_TRANSLATE_]";
        else if (i == v.token_transcribe) w = "[_TRANSCRIBE_]";
        else if (i == v.token_solm) w = "[_SOLM_]";
        else if (i == v.token_prev) w = "[_PREV_]";
        else if (i == v.token_nosp) w = "[_NOSP_]";
        else if (i == v.token_not) w = "[_NOT_]";
        else if (i == v.token_beg) w = "[_BEG_]";
        else if (i > v.token_sot && i <= v.token_sot + v.num_languages()) w = std::string("[_LANG_") + LANGS[i - v.token_sot - 1] + "]";
        else w = "[_extra_token_" + std::to_string(i) + "]";
        v.id_to_token.push_back(w);
        v.token_to_id[w] = i;
    }
    std::map<std::string, Vec> T;
    while (true) {
        int32_t nd, nl, tt;
        if (!rd(&nd, 4)) break;
        rd(&nl, 4); rd(&tt, 4);
        // ggml block formats (32 weights per block): type -> bytes per block
        const int qbytes = tt == 2 ? 18 : tt == 3 ? 20 : tt == 6 ? 22 : tt == 7 ? 24 : tt == 8 ? 34 : 0;
        if (nd < 1 || nd > 4 || nl <= 0 || nl > 256 || (tt != 0 && tt != 1 && !qbytes)) { err = "bad tensor header"; fclose(f); return false; }
        size_t ne = 1;
        for (int i = 0; i < nd; ++i) { int32_t d; rd(&d, 4); ne *= (size_t)d; }
        std::string name(nl, '\0');
        rd(&name[0], nl);
        Vec d(ne);
        if (qbytes) {
            // dequantise as ggml's dequantize_row_q{4_0,4_1,5_0,5_1,8_0}: y = q * d (+ m), low nibbles first, fifth bits in qh
            std::vector<uint8_t> raw(ne / 32 * (size_t)qbytes);
            if (ne % 32 || !rd(raw.data(), raw.size())) { err = "truncated tensor"; fclose(f); return false; }
            for (size_t b = 0; b < ne / 32; ++b) {
                const uint8_t* p = raw.data() + b * qbytes;
                float* y = d.data() + b * 32;
                uint16_t h16; memcpy(&h16, p, 2); p += 2;
                const float dd = f16_to_f32(h16);
                float mm = 0.0f;
                if (tt == 3 || tt == 7) { memcpy(&h16, p, 2); p += 2; mm = f16_to_f32(h16); }
                if (tt == 8) { for (int j = 0; j < 32; ++j) y[j] = (float)(int8_t)p[j] * dd; continue; }
                uint32_t qh = 0;
                if (tt == 6 || tt == 7) { memcpy(&qh, p, 4); p += 4; }
                for (int j = 0; j < 16; ++j) {
                    int lo = p[j] & 15, hi = p[j] >> 4;
                    if (tt == 6 || tt == 7) { lo |= ((qh >> j) & 1) << 4; hi |= ((qh >> (j + 16)) & 1) << 4; }
                    if (tt == 2) { lo -= 8; hi -= 8; }
                    if (tt == 6) { lo -= 16; hi -= 16; }
                    y[j] = (float)lo * dd + mm;
                    y[j + 16] = (float)hi * dd + mm;
                }
            }
        }
        else if (tt == 0) { if (!rd(d.data(), ne * 4)) { err = "truncated tensor"; fclose(f); return false; } }
        else {
            std::vector<uint16_t> h(ne);
            if (!rd(h.data(), ne * 2)) { err = "truncated tensor"; fclose(f); return false; }
            for (size_t i = 0; i < ne; ++i) d[i] = f16_to_f32(h[i]);
        }
        T[name] = std::move(d);
    }
    fclose(f);
    bool ok = true;
    auto get = [&](const std::string& n, size_t want) -> Vec {
        auto it = T.find(n);
        if (it == T.end() || it->second.size() != want) { ok = false; err = "missing/mis-sized tensor " + n; return Vec(); }
        return std::move(it->second);
    };
    const HParams& hp = m.hp;
    const size_t d = hp.n_audio_state, dt = hp.n_text_state;
    m.d_pe = get("decoder.positional_embedding", (size_t)hp.n_text_ctx * dt);
    m.e_pe = get("encoder.positional_embedding", (size_t)hp.n_audio_ctx * d);
    m.d_te = get("decoder.token_embedding.weight", (size_t)hp.n_vocab * dt);
    m.e_conv1_w = get("encoder.conv1.weight", d * hp.n_mels * 3);
    m.e_conv1_b = get("encoder.conv1.bias", d);
    m.e_conv2_w = get("encoder.conv2.weight", d * d * 3);
    m.e_conv2_b = get("encoder.conv2.bias", d);
    m.e_ln_w = get("encoder.ln_post.weight", d);
    m.e_ln_b = get("encoder.ln_post.bias", d);
    m.d_ln_w = get("decoder.ln.weight", dt);
    m.d_ln_b = get("decoder.ln.bias", dt);
    auto load_attn = [&](const std::string& p, Attn& a, size_t n) {
        a.ln_w = get(p + "_ln.weight", n); a.ln_b = get(p + "_ln.bias", n);
        a.q_w = get(p + ".query.weight", n * n); a.q_b = get(p + ".query.bias", n);
        a.k_w = get(p + ".key.weight", n * n);
        a.v_w = get(p + ".value.weight", n * n); a.v_b = get(p + ".value.bias", n);
        a.o_w = get(p + ".out.weight", n * n); a.o_b = get(p + ".out.bias", n);
    };
    auto load_mlp = [&](const std::string& p, Block& b, size_t n) {
        b.mlp_ln_w = get(p + "mlp_ln.weight", n); b.mlp_ln_b = get(p + "mlp_ln.bias", n);
        b.fc1_w = get(p + "mlp.0.weight", 4 * n * n); b.fc1_b = get(p + "mlp.0.bias", 4 * n);
        b.fc2_w = get(p + "mlp.2.weight", 4 * n * n); b.fc2_b = get(p + "mlp.2.bias", n);
    };
    m.enc.resize(hp.n_audio_layer);
    for (int i = 0; i < hp.n_audio_layer; ++i) {
        std::string p = "encoder.blocks." + std::to_string(i) + ".";
        load_attn(p + "attn", m.enc[i].attn, d);
        load_mlp(p, m.enc[i], d);
    }
    m.dec.resize(hp.n_text_layer);
    for (int i = 0; i < hp.n_text_layer; ++i) {
        std::string p = "decoder.blocks." + std::to_string(i) + ".";
        load_attn(p + "attn", m.dec[i].attn, dt);
        load_attn(p + "cross_attn", m.dec[i].cross, dt);
        load_mlp(p, m.dec[i], dt);
    }
    return ok;
}

// ------------------------------------------------------------------ tokenizer
// restates whisper.cpp tokenize(): GPT-2 regex word split, then greedy longest match
// against the vocabulary (NOT BPE merges).  SURVEY.md §8a row a6.
std::vector<int> tokenize(const Vocab& vocab, const std::string& text) {
    std::vector<std::string> words;
    {
        std::string str = text;
        std::string pat = R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)";
        std::regex re(pat);
        std::smatch m;
        while (std::regex_search(str, m, re)) {
            for (auto x : m) words.push_back(x);
            str = m.suffix();
        }
    }
    std::vector<int> tokens;
    for (const auto& word : words) {
        if (word.empty()) continue;
        int i = 0, n = (int)word.size();
        while (i < n) {
            int j = n;
            bool found = false;
            while (j > i) {
                auto it = vocab.token_to_id.find(word.substr(i, j - i));
                if (it != vocab.token_to_id.end()) { tokens.push_back(it->second); i = j; found = true; break; }
                --j;
            }
            if (!found) ++i;
        }
    }
    return tokens;
}

// ------------------------------------------------------------------ linear algebra
// C[M,N] (ldc) = A[M,K] (lda) * W[N,K]^T (ldw) + bias[N]; fp32, K-contiguous operands.
void gemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N, int K) {
    // cache tiling: a 64-row W tile stays in L2 while 64-row A blocks stream past it; the
    // micro-kernel is 2 A rows x 4 W rows with K-vectorised accumulators.
    const int TM = 64, TN = 64;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int j0 = 0; j0 < N; j0 += TN) {
        for (int ib = 0; ib < M; ib += TM) {
            const int j1 = std::min(N, j0 + TN), ie = std::min(M, ib + TM);
            for (int i0 = ib; i0 < ie; i0 += 2) {
                const bool two = i0 + 1 < ie;
                const float* a0 = A + (size_t)i0 * lda;
                const float* a1 = two ? a0 + lda : a0;
                int j = j0;
                for (; j + 4 <= j1; j += 4) {
                    const float* w0 = W + (size_t)j * ldw;
                    const float* w1 = w0 + ldw; const float* w2 = w1 + ldw; const float* w3 = w2 + ldw;
                    float s00 = 0, s01 = 0, s02 = 0, s03 = 0, s10 = 0, s11 = 0, s12 = 0, s13 = 0;
#pragma omp simd reduction(+ : s00, s01, s02, s03, s10, s11, s12, s13)
                    for (int k = 0; k < K; ++k) {
                        const float x0 = a0[k], x1 = a1[k];
                        s00 += x0 * w0[k]; s01 += x0 * w1[k]; s02 += x0 * w2[k]; s03 += x0 * w3[k];
                        s10 += x1 * w0[k]; s11 += x1 * w1[k]; s12 += x1 * w2[k]; s13 += x1 * w3[k];
                    }
                    float* c0 = C + (size_t)i0 * ldc + j;
                    const float b0 = bias ? bias[j] : 0, b1 = bias ? bias[j + 1] : 0, b2 = bias ? bias[j + 2] : 0, b3 = bias ? bias[j + 3] : 0;
                    c0[0] = s00 + b0; c0[1] = s01 + b1; c0[2] = s02 + b2; c0[3] = s03 + b3;
                    if (two) { float* c1 = c0 + ldc; c1[0] = s10 + b0; c1[1] = s11 + b1; c1[2] = s12 + b2; c1[3] = s13 + b3; }
                }
                for (; j < j1; ++j) {
                    const float* w0 = W + (size_t)j * ldw;
                    float s0 = 0, s1 = 0;
#pragma omp simd reduction(+ : s0, s1)
                    for (int k = 0; k < K; ++k) { s0 += a0[k] * w0[k]; s1 += a1[k] * w0[k]; }
                    const float b = bias ? bias[j] : 0;
                    C[(size_t)i0 * ldc + j] = s0 + b;
                    if (two) C[(size_t)(i0 + 1) * ldc + j] = s1 + b;
                }
            }
        }
    }
}

void layer_norm(const float* x, const float* g, const float* b, float* y, int rows, int d) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const float* xr = x + (size_t)r * d;
        float* yr = y + (size_t)r * d;
        double mean = 0;
        for (int i = 0; i < d; ++i) mean += xr[i];
        mean /= d;
        double var = 0;
        for (int i = 0; i < d; ++i) { double t = xr[i] - mean; var += t * t; }
        var /= d;
        const float inv = (float)(1.0 / std::sqrt(var + 1e-5));
        for (int i = 0; i < d; ++i) yr[i] = (float)(xr[i] - mean) * inv * g[i] + b[i];
    }
}

// round-to-nearest-even fp32 -> IEEE half -> fp32 (what ggml's GGML_FP32_TO_FP16 / F16C does)
inline float round_f16(float x) { return (float)(_Float16)x; }
void round_f16_inplace(float* v, size_t n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) v[i] = round_f16(v[i]);
}

// whisper.cpp/ggml GELU: tanh approximation.  Default: ideal fp32 (no f16 table).  "erf" is only used to cross-check against
// HF transformers (which uses exact GELU).  f16_table: ggml's CPU path as it really runs — ggml_vec_gelu_f32 looks the value
// up in a table indexed by the HALF-precision bit pattern of x and stored in half precision: y = f16(gelu(f16(x))) for
// -10 < x < 10, 0 below, x above.
struct Gelu {
    bool erf_mode = false;
    bool f16_table = false;
    inline float operator()(float x) const {
        if (erf_mode) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        if (f16_table) {
            if (x <= -10.0f) return 0.0f;
            if (x >= 10.0f) return x;
            const float xh = round_f16(x);
            return round_f16(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh))));
        }
        return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
    }
};

// ------------------------------------------------------------------ log-mel
struct Mel {
    int n_len = 0, n_len_org = 0, n_mel = 0;
    Vec data;  // [n_mel][n_len]
};

struct FftTables {
    float s[N_FFT], c[N_FFT], hann[N_FFT];
    FftTables() {
        for (int i = 0; i < N_FFT; ++i) {
            double th = 2.0 * M_PI * i / N_FFT;
            s[i] = (float)sin(th);
            c[i] = (float)cos(th);
            hann[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * i / N_FFT)));
        }
    }
};
const FftTables g_fft;

void dft(const float* in, int N, float* out) {
    const int step = N_FFT / N;
    for (int k = 0; k < N; ++k) {
        float re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            int idx = (k * n * step) % N_FFT;
            re += in[n] * g_fft.c[idx];
            im -= in[n] * g_fft.s[idx];
        }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

// radix-2 decimation in time down to an odd length, then a direct DFT (400->200->100->50->25)
void fft(float* in, int N, float* out) {
    if (N == 1) { out[0] = in[0]; out[1] = 0; return; }
    const int half = N / 2;
    if (N - half * 2 == 1) { dft(in, N, out); return; }
    float* even = in + N;
    for (int i = 0; i < half; ++i) even[i] = in[2 * i];
    float* even_fft = out + 2 * N;
    fft(even, half, even_fft);
    float* odd = even;
    for (int i = 0; i < half; ++i) odd[i] = in[2 * i + 1];
    float* odd_fft = even_fft + N;
    fft(odd, half, odd_fft);
    const int step = N_FFT / N;
    for (int k = 0; k < half; ++k) {
        int idx = k * step;
        float re = g_fft.c[idx], im = -g_fft.s[idx];
        float re_odd = odd_fft[2 * k], im_odd = odd_fft[2 * k + 1];
        out[2 * k] = even_fft[2 * k] + re * re_odd - im * im_odd;
        out[2 * k + 1] = even_fft[2 * k + 1] + re * im_odd + im * re_odd;
        out[2 * (k + half)] = even_fft[2 * k] - re * re_odd + im * im_odd;
        out[2 * (k + half) + 1] = even_fft[2 * k + 1] - re * im_odd - im * re_odd;
    }
}

void log_mel(const float* samples, int n, const Model& m, Mel& mel) {
    const int pad1 = SAMPLE_RATE * CHUNK_SIZE, pad2 = N_FFT / 2;
    Vec padded((size_t)n + pad1 + 2 * pad2, 0.0f);
    std::copy(samples, samples + n, padded.begin() + pad2);
    // reflective pad at the beginning (right edge is zero padded)
    for (int i = 0; i < pad2 && 1 + i < n; ++i) padded[pad2 - 1 - i] = samples[1 + i];
    mel.n_mel = m.hp.n_mels;
    mel.n_len = (int)((padded.size() - N_FFT) / HOP);
    mel.n_len_org = 1 + (n + pad2 - N_FFT) / HOP;
    mel.data.assign((size_t)mel.n_mel * mel.n_len, 0.0f);
    const int n_samples = (int)padded.size();
#pragma omp parallel
    {
        Vec fin(N_FFT * 2), fout(N_FFT * 2 * 2 * 2);
#pragma omp for schedule(static)
        for (int i = 0; i < mel.n_len; ++i) {
            const int off = i * HOP;
            const int lim = std::min(N_FFT, n_samples - off);
            for (int j = 0; j < lim; ++j) fin[j] = g_fft.hann[j] * padded[off + j];
            for (int j = std::max(lim, 0); j < N_FFT; ++j) fin[j] = 0;
            fft(fin.data(), N_FFT, fout.data());
            for (int j = 0; j < N_FREQ; ++j) fout[j] = fout[2 * j] * fout[2 * j] + fout[2 * j + 1] * fout[2 * j + 1];
            for (int j = 0; j < mel.n_mel; ++j) {
                double sum = 0;
                const float* fl = m.filters.data() + (size_t)j * N_FREQ;
                for (int k = 0; k < N_FREQ; ++k) sum += fout[k] * fl[k];
                sum = log10(std::max(sum, 1e-10));
                mel.data[(size_t)j * mel.n_len + i] = (float)sum;
            }
        }
    }
    double mmax = -1e20;
    for (float v : mel.data) if (v > mmax) mmax = v;
    mmax -= 8.0;
    for (float& v : mel.data) {
        if (v < mmax) v = (float)mmax;
        v = (float)((v + 4.0) / 4.0);
    }
}

// ------------------------------------------------------------------ sequence bookkeeping
struct TokenData {
    int id = 0, tid = 0;
    float p = 0, plog = 0, pt = 0, ptsum = 0;
};
struct Sequence {
    std::vector<TokenData> tokens;
    int result_len = 0;
    double sum_logprobs_all = 0, sum_logprobs = -INFINITY, avg_logprobs = -INFINITY, entropy = 0, score = -INFINITY;
};
struct Decoder {
    Sequence sequence;
    int seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;
    Vec probs, logits, logprobs;
    std::mt19937 rng{0};
};
struct Segment {
    int64_t t0, t1;
    std::string text;
    float no_speech_prob;
    std::vector<TokenData> tokens;
};

}  // namespace

// ------------------------------------------------------------------ public C surface
extern "C" {

struct wo_params {
    int strategy;  // 0 greedy, 1 beam
    int best_of, beam_size;
    const char* language;        // NULL / "" / "auto" => detect
    const char* initial_prompt;  // NULL => none
    int translate, no_context, single_segment, no_timestamps, suppress_blank;
    float temperature, temperature_inc, max_initial_ts, length_penalty;
    float entropy_thold, logprob_thold, no_speech_thold;
    int n_max_text_ctx, max_tokens;
    int beam_sampled;  // 0: deterministic top-k candidates (north star); 1: upstream-style sampled candidates
};

struct wo_ctx {
    Model model;
    Gelu gelu;
    Mel mel;
    // encoder / cross-KV state
    Vec enc_out;                          // [n_ctx][d]
    std::vector<Vec> cross_k, cross_v;    // per decoder layer [n_ctx][d]
    // self KV per sequence id
    int n_seq = 0;
    std::vector<Vec> self_k, self_v;      // [layer*n_seq + seq][n_text_ctx][d]
    Vec logits;                           // logits of the last decoded token (per wo_decode call)
    std::vector<Decoder> decoders;
    std::vector<int> prompt_past;
    std::vector<Segment> result_all;
    float no_speech_prob = 0;
    int lang_id = 0;
    std::string err;
    // "ggml-faithful" numerics (SURVEY.md §7 step 1, §8c hazard 1): the roundings ggml's CPU backend applies on top of the ideal
    // fp32 graph — (1) mul_mat converts its activation operand to the weight's vec_dot type, i.e. to f16 for f16 weights,
    // (2) conv1d runs through an f16 im2col, (3) GELU goes through the f16 table, (4) self- and cross-KV are stored in f16, and the
    // attention products take f16 operands (Q and the softmax probabilities are converted because K / V are f16).
    // Off by default: the product's tolerances are stated against the ideal-fp32 mode; this mode measures how far apart the two are.
    bool faithful = false;
    // scripted-logits hook of the control-flow known-answer tests (tests/test_control_flow_kat.py): called after every decode
    // whose logits are sampled from; may overwrite them.  step = index of the token about to be sampled (0: from the prompt).
    int (*logits_hook)(void* user, int seek, int i_temp, int step, int decoder, int n_prompt, int n_vocab, float* logits) = nullptr;
    void* logits_hook_user = nullptr;
    // stats for the benchmark harness
    long n_encode = 0, n_decode_tokens = 0, n_decode_calls = 0, n_fail_p = 0, n_fail_h = 0;
};

wo_ctx* wo_load(const char* path) {
    wo_ctx* c = new wo_ctx();
    if (!load_model(path, c->model, c->err)) {
        fprintf(stderr, "[oracle] load failed: %s\n", c->err.c_str());
        delete c;
        return nullptr;
    }
    const HParams& hp = c->model.hp;
    c->n_seq = 16;
    c->self_k.assign((size_t)hp.n_text_layer * c->n_seq, Vec());
    c->self_v.assign((size_t)hp.n_text_layer * c->n_seq, Vec());
    c->cross_k.assign(hp.n_text_layer, Vec());
    c->cross_v.assign(hp.n_text_layer, Vec());
    c->decoders.resize(8);
    return c;
}
void wo_free(wo_ctx* c) { delete c; }
void wo_set_gelu_erf(wo_ctx* c, int on) { c->gelu.erf_mode = on != 0; }
void wo_set_ggml_faithful(wo_ctx* c, int on) { c->faithful = on != 0; c->gelu.f16_table = on != 0; }
float wo_gelu(wo_ctx* c, float x) { return c->gelu(x); }   // the activation in the context's current numeric mode (known-answer tests)
void wo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
    (void)n;
}
int wo_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void wo_hparams(wo_ctx* c, int* out11) { memcpy(out11, &c->model.hp, sizeof(HParams)); }
void wo_special_tokens(wo_ctx* c, int* out9) {
    const Vocab& v = c->model.vocab;
    int t[9] = {v.token_eot, v.token_sot, v.token_translate, v.token_transcribe, v.token_solm, v.token_prev, v.token_nosp, v.token_not, v.token_beg};
    memcpy(out9, t, sizeof(t));
}
int wo_tokenize(wo_ctx* c, const char* text, int* out, int cap) {
    auto t = tokenize(c->model.vocab, text);
    if ((int)t.size() > cap) return -(int)t.size();
    std::copy(t.begin(), t.end(), out);
    return (int)t.size();
}
const char* wo_token_str(wo_ctx* c, int id) {
    if (id < 0 || id >= (int)c->model.vocab.id_to_token.size()) return "";
    return c->model.vocab.id_to_token[id].c_str();
}
int wo_token_len(wo_ctx* c, int id) {
    if (id < 0 || id >= (int)c->model.vocab.id_to_token.size()) return 0;
    return (int)c->model.vocab.id_to_token[id].size();
}

// ---- stage 1: log-mel.  returns n_len; data pointer [n_mel][n_len]
int wo_pcm_to_mel(wo_ctx* c, const float* pcm, int n) {
    log_mel(pcm, n, c->model, c->mel);
    return c->mel.n_len;
}
int wo_mel_n_len(wo_ctx* c) { return c->mel.n_len; }
int wo_mel_n_len_org(wo_ctx* c) { return c->mel.n_len_org; }
const float* wo_mel_data(wo_ctx* c) { return c->mel.data.data(); }

// ---- stage 2: encoder + cross-KV
int wo_encode(wo_ctx* c, int mel_offset) {
    const Model& m = c->model;
    const HParams& hp = m.hp;
    const int n_ctx = hp.n_audio_ctx, d = hp.n_audio_state, n_head = hp.n_audio_head, dh = d / n_head;
    const int T = 2 * n_ctx, nm = hp.n_mels;
    if (c->mel.n_len == 0) return -1;
    c->n_encode++;
    // conv1 input im2col: A1[t][(ch,k)] = mel[ch][mel_offset + t + k - 1]
    Vec A1((size_t)T * nm * 3, 0.0f);
    const int i0 = std::min(mel_offset, c->mel.n_len), i1 = std::min(mel_offset + T, c->mel.n_len);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; ++t)
        for (int ch = 0; ch < nm; ++ch)
            for (int k = 0; k < 3; ++k) {
                int tt = t + k - 1;
                float v = 0;
                if (tt >= 0 && tt < T && i0 + tt < i1) v = c->mel.data[(size_t)ch * c->mel.n_len + i0 + tt];
                A1[((size_t)t * nm + ch) * 3 + k] = v;
            }
    // faithful mode: F = "this tensor is handed to ggml_mul_mat next to f16 data" -> rounded to f16 first
    const bool fa = c->faithful, fw = c->faithful && hp.ftype == 1;   // fw: the model's matrices are f16 (ggml-<id>.bin as shipped)
    auto F = [&](Vec& t, bool on) { if (on) round_f16_inplace(t.data(), t.size()); };
    F(A1, fa);   // ggml_conv_1d: im2col writes f16
    Vec h1((size_t)T * d);
    gemm_nt(A1.data(), nm * 3, m.e_conv1_w.data(), nm * 3, m.e_conv1_b.data(), h1.data(), d, T, d, nm * 3);
    for (auto& v : h1) v = c->gelu(v);
    // conv2 stride 2: A2[t'][(ch,k)] = h1[2t'+k-1][ch]
    Vec A2((size_t)n_ctx * d * 3);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < n_ctx; ++t)
        for (int ch = 0; ch < d; ++ch)
            for (int k = 0; k < 3; ++k) {
                int tt = 2 * t + k - 1;
                A2[((size_t)t * d + ch) * 3 + k] = (tt >= 0 && tt < T) ? h1[(size_t)tt * d + ch] : 0.0f;
            }
    F(A2, fa);
    Vec x((size_t)n_ctx * d);
    gemm_nt(A2.data(), d * 3, m.e_conv2_w.data(), d * 3, m.e_conv2_b.data(), x.data(), d, n_ctx, d, d * 3);
    for (size_t i = 0; i < x.size(); ++i) x[i] = c->gelu(x[i]) + m.e_pe[i];
    A1.clear(); A1.shrink_to_fit(); A2.clear(); A2.shrink_to_fit(); h1.clear(); h1.shrink_to_fit();

    Vec y((size_t)n_ctx * d), q((size_t)n_ctx * d), k((size_t)n_ctx * d), v((size_t)n_ctx * d), att((size_t)n_ctx * d);
    Vec hbuf((size_t)n_ctx * 4 * d);
    const float scale = 1.0f / sqrtf((float)dh);
    for (const Block& b : m.enc) {
        layer_norm(x.data(), b.attn.ln_w.data(), b.attn.ln_b.data(), y.data(), n_ctx, d);
        F(y, fw);
        gemm_nt(y.data(), d, b.attn.q_w.data(), d, b.attn.q_b.data(), q.data(), d, n_ctx, d, d);
        gemm_nt(y.data(), d, b.attn.k_w.data(), d, nullptr, k.data(), d, n_ctx, d, d);
        gemm_nt(y.data(), d, b.attn.v_w.data(), d, b.attn.v_b.data(), v.data(), d, n_ctx, d, d);
        F(k, fa); F(v, fa); F(q, fa);   // K and V are copied into f16 tensors; Q is converted because its partner K is f16
        // non-causal attention, one head at a time (threads across query rows)
        for (int h = 0; h < n_head; ++h) {
            Vec vt((size_t)dh * n_ctx);
            for (int j = 0; j < n_ctx; ++j)
                for (int e = 0; e < dh; ++e) vt[(size_t)e * n_ctx + j] = v[(size_t)j * d + h * dh + e];
            Vec s((size_t)n_ctx * n_ctx);
            gemm_nt(q.data() + h * dh, d, k.data() + h * dh, d, nullptr, s.data(), n_ctx, n_ctx, n_ctx, dh);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n_ctx; ++i) {
                float* r = s.data() + (size_t)i * n_ctx;
                float mx = -INFINITY;
                for (int j = 0; j < n_ctx; ++j) { r[j] *= scale; mx = std::max(mx, r[j]); }
                double sum = 0;
                for (int j = 0; j < n_ctx; ++j) { r[j] = expf(r[j] - mx); sum += r[j]; }
                const float inv = (float)(1.0 / sum);
                for (int j = 0; j < n_ctx; ++j) r[j] *= inv;
                if (fa) for (int j = 0; j < n_ctx; ++j) r[j] = round_f16(r[j]);   // mul_mat(V f16, softmax): probabilities become f16
            }
            gemm_nt(s.data(), n_ctx, vt.data(), n_ctx, nullptr, att.data() + h * dh, d, n_ctx, dh, n_ctx);
        }
        F(att, fw);
        gemm_nt(att.data(), d, b.attn.o_w.data(), d, b.attn.o_b.data(), y.data(), d, n_ctx, d, d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += y[i];
        layer_norm(x.data(), b.mlp_ln_w.data(), b.mlp_ln_b.data(), y.data(), n_ctx, d);
        F(y, fw);
        gemm_nt(y.data(), d, b.fc1_w.data(), d, b.fc1_b.data(), hbuf.data(), 4 * d, n_ctx, 4 * d, d);
        for (auto& t : hbuf) t = c->gelu(t);
        F(hbuf, fw);
        gemm_nt(hbuf.data(), 4 * d, b.fc2_w.data(), 4 * d, b.fc2_b.data(), y.data(), d, n_ctx, d, 4 * d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += y[i];
    }
    c->enc_out.resize((size_t)n_ctx * d);
    layer_norm(x.data(), m.e_ln_w.data(), m.e_ln_b.data(), c->enc_out.data(), n_ctx, d);
    // cross-KV projection per decoder layer (K has no bias)
    const int dt = hp.n_text_state;
    Vec enc_in = c->enc_out;
    F(enc_in, fw);
    for (int l = 0; l < hp.n_text_layer; ++l) {
        c->cross_k[l].resize((size_t)n_ctx * dt);
        c->cross_v[l].resize((size_t)n_ctx * dt);
        gemm_nt(enc_in.data(), d, m.dec[l].cross.k_w.data(), d, nullptr, c->cross_k[l].data(), dt, n_ctx, dt, d);
        gemm_nt(enc_in.data(), d, m.dec[l].cross.v_w.data(), d, m.dec[l].cross.v_b.data(), c->cross_v[l].data(), dt, n_ctx, dt, d);
        F(c->cross_k[l], fa); F(c->cross_v[l], fa);   // the cross-KV cache is f16
    }
    return 0;
}
const float* wo_encoder_out(wo_ctx* c) { return c->enc_out.data(); }
const float* wo_cross_k(wo_ctx* c, int layer) { return c->cross_k[layer].data(); }
const float* wo_cross_v(wo_ctx* c, int layer) { return c->cross_v[layer].data(); }

// ---- stage 3: decoder forward for one sequence.
// tokens[n] at positions n_past..n_past+n-1 of sequence `seq`; logits of the LAST token
// are left in c->logits.
int wo_decode(wo_ctx* c, const int* tokens, int n, int n_past, int seq) {
    const Model& m = c->model;
    const HParams& hp = m.hp;
    const int d = hp.n_text_state, n_head = hp.n_text_head, dh = d / n_head, n_ctx = hp.n_audio_ctx;
    if (n <= 0 || n_past + n > hp.n_text_ctx || seq < 0 || seq >= c->n_seq) return -1;
    if (c->enc_out.empty()) return -2;
    c->n_decode_calls++;
    c->n_decode_tokens += n;
    const bool fa = c->faithful, fw = c->faithful && hp.ftype == 1;   // see wo_encode
    auto F = [&](Vec& t, bool on) { if (on) round_f16_inplace(t.data(), t.size()); };
    Vec x((size_t)n * d), y((size_t)n * d), q((size_t)n * d), att((size_t)n * d), hbuf((size_t)n * 4 * d), tmp((size_t)n * d);
    for (int i = 0; i < n; ++i)
        for (int e = 0; e < d; ++e) x[(size_t)i * d + e] = m.d_te[(size_t)tokens[i] * d + e] + m.d_pe[(size_t)(n_past + i) * d + e];
    const float scale = 1.0f / sqrtf((float)dh);
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const Block& b = m.dec[l];
        Vec& K = c->self_k[(size_t)l * c->n_seq + seq];
        Vec& V = c->self_v[(size_t)l * c->n_seq + seq];
        if (K.empty()) { K.assign((size_t)hp.n_text_ctx * d, 0.0f); V.assign((size_t)hp.n_text_ctx * d, 0.0f); }
        // self-attention with KV append
        layer_norm(x.data(), b.attn.ln_w.data(), b.attn.ln_b.data(), y.data(), n, d);
        F(y, fw);
        gemm_nt(y.data(), d, b.attn.q_w.data(), d, b.attn.q_b.data(), q.data(), d, n, d, d);
        gemm_nt(y.data(), d, b.attn.k_w.data(), d, nullptr, K.data() + (size_t)n_past * d, d, n, d, d);
        gemm_nt(y.data(), d, b.attn.v_w.data(), d, b.attn.v_b.data(), V.data() + (size_t)n_past * d, d, n, d, d);
        if (fa) {   // the self-KV cache is f16; Q is converted next to it
            round_f16_inplace(K.data() + (size_t)n_past * d, (size_t)n * d);
            round_f16_inplace(V.data() + (size_t)n_past * d, (size_t)n * d);
            round_f16_inplace(q.data(), q.size());
        }
#pragma omp parallel for collapse(2) schedule(static)
        for (int i = 0; i < n; ++i)
            for (int h = 0; h < n_head; ++h) {
                const int nk = n_past + i + 1;
                float s[512];
                const float* qi = q.data() + (size_t)i * d + h * dh;
                float mx = -INFINITY;
                for (int j = 0; j < nk; ++j) {
                    const float* kj = K.data() + (size_t)j * d + h * dh;
                    float a = 0;
                    for (int e = 0; e < dh; ++e) a += qi[e] * kj[e];
                    s[j] = a * scale;
                    mx = std::max(mx, s[j]);
                }
                double sum = 0;
                for (int j = 0; j < nk; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
                const float inv = (float)(1.0 / sum);
                float* o = att.data() + (size_t)i * d + h * dh;
                for (int e = 0; e < dh; ++e) o[e] = 0;
                for (int j = 0; j < nk; ++j) {
                    const float pj = fa ? round_f16(s[j] * inv) : s[j] * inv;
                    const float* vj = V.data() + (size_t)j * d + h * dh;
                    for (int e = 0; e < dh; ++e) o[e] += pj * vj[e];
                }
            }
        F(att, fw);
        gemm_nt(att.data(), d, b.attn.o_w.data(), d, b.attn.o_b.data(), tmp.data(), d, n, d, d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += tmp[i];
        // cross-attention over the n_ctx encoder positions
        layer_norm(x.data(), b.cross.ln_w.data(), b.cross.ln_b.data(), y.data(), n, d);
        F(y, fw);
        gemm_nt(y.data(), d, b.cross.q_w.data(), d, b.cross.q_b.data(), q.data(), d, n, d, d);
        F(q, fa);
        const Vec& CK = c->cross_k[l];
        const Vec& CV = c->cross_v[l];
#pragma omp parallel for collapse(2) schedule(static)
        for (int i = 0; i < n; ++i)
            for (int h = 0; h < n_head; ++h) {
                float s[1500 + 36];
                const float* qi = q.data() + (size_t)i * d + h * dh;
                float mx = -INFINITY;
                for (int j = 0; j < n_ctx; ++j) {
                    const float* kj = CK.data() + (size_t)j * d + h * dh;
                    float a = 0;
                    for (int e = 0; e < dh; ++e) a += qi[e] * kj[e];
                    s[j] = a * scale;
                    mx = std::max(mx, s[j]);
                }
                double sum = 0;
                for (int j = 0; j < n_ctx; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
                const float inv = (float)(1.0 / sum);
                float* o = att.data() + (size_t)i * d + h * dh;
                for (int e = 0; e < dh; ++e) o[e] = 0;
                for (int j = 0; j < n_ctx; ++j) {
                    const float pj = fa ? round_f16(s[j] * inv) : s[j] * inv;
                    const float* vj = CV.data() + (size_t)j * d + h * dh;
                    for (int e = 0; e < dh; ++e) o[e] += pj * vj[e];
                }
            }
        F(att, fw);
        gemm_nt(att.data(), d, b.cross.o_w.data(), d, b.cross.o_b.data(), tmp.data(), d, n, d, d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += tmp[i];
        // MLP
        layer_norm(x.data(), b.mlp_ln_w.data(), b.mlp_ln_b.data(), y.data(), n, d);
        F(y, fw);
        gemm_nt(y.data(), d, b.fc1_w.data(), d, b.fc1_b.data(), hbuf.data(), 4 * d, n, 4 * d, d);
        for (auto& t : hbuf) t = c->gelu(t);
        F(hbuf, fw);
        gemm_nt(hbuf.data(), 4 * d, b.fc2_w.data(), 4 * d, b.fc2_b.data(), tmp.data(), d, n, d, 4 * d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += tmp[i];
    }
    Vec last(d);
    layer_norm(x.data() + (size_t)(n - 1) * d, m.d_ln_w.data(), m.d_ln_b.data(), last.data(), 1, d);
    F(last, fw);
    c->logits.resize(hp.n_vocab);
    gemm_nt(last.data(), d, m.d_te.data(), d, nullptr, c->logits.data(), hp.n_vocab, 1, hp.n_vocab, d);
    return 0;
}
const float* wo_logits(wo_ctx* c) { return c->logits.data(); }

static void kv_copy(wo_ctx* c, int src, int dst) {
    if (src == dst) return;
    for (int l = 0; l < c->model.hp.n_text_layer; ++l) {
        c->self_k[(size_t)l * c->n_seq + dst] = c->self_k[(size_t)l * c->n_seq + src];
        c->self_v[(size_t)l * c->n_seq + dst] = c->self_v[(size_t)l * c->n_seq + src];
    }
}

// ---- logits -> logprobs/probs with whisper's filters
static void process_logits(wo_ctx* c, Decoder& dec, const wo_params& p, float temperature) {
    const Vocab& vocab = c->model.vocab;
    const auto& cur = dec.sequence.tokens;
    const bool is_initial = cur.empty();
    const int n_logits = vocab.n_vocab;
    Vec& logits = dec.logits; Vec& logprobs = dec.logprobs; Vec& probs = dec.probs;
    logits = c->logits;
    if (temperature > 0.0f) for (auto& v : logits) v /= temperature;
    probs.resize(n_logits); logprobs.resize(n_logits);
    if (p.suppress_blank && is_initial) {
        logits[vocab.token_eot] = -INFINITY;
        auto it = vocab.token_to_id.find(" ");
        if (it != vocab.token_to_id.end()) logits[it->second] = -INFINITY;
    }
    logits[vocab.token_not] = -INFINITY;
    if (p.no_timestamps) for (int i = vocab.token_beg; i < n_logits; ++i) logits[i] = -INFINITY;
    logits[vocab.token_sot] = -INFINITY;
    logits[vocab.token_nosp] = -INFINITY;
    logits[vocab.token_solm] = -INFINITY;
    logits[vocab.token_translate] = -INFINITY;
    logits[vocab.token_transcribe] = -INFINITY;
    logits[vocab.token_prev] = -INFINITY;
    for (int i = 0; i < N_LANGS; ++i) logits[vocab.token_lang(i)] = -INFINITY;
    {
        const bool last_ts = !cur.empty() && cur.back().id >= vocab.token_beg;
        const bool pen_ts = cur.size() < 2 || cur[cur.size() - 2].id >= vocab.token_beg;
        if (last_ts) {
            if (pen_ts) for (int i = vocab.token_beg; i < n_logits; ++i) logits[i] = -INFINITY;
            else for (int i = 0; i < vocab.token_eot; ++i) logits[i] = -INFINITY;
        }
    }
    if (is_initial && p.max_initial_ts > 0.0f) {
        const float precision = float(CHUNK_SIZE) / c->model.hp.n_audio_ctx;
        const int tid0 = (int)std::round(p.max_initial_ts / precision);
        for (int i = vocab.token_beg + tid0 + 1; i < n_logits; ++i) logits[i] = -INFINITY;
    }
    if (dec.has_ts) {
        const int tid0 = dec.seek_delta / 2;
        for (int i = vocab.token_beg; i < vocab.token_beg + tid0 && i < n_logits; ++i) logits[i] = -INFINITY;
    }
    {
        const float logit_max = *std::max_element(logits.begin(), logits.end());
        float lse = 0.0f;
        for (int i = 0; i < n_logits; ++i) if (logits[i] > -INFINITY) lse += expf(logits[i] - logit_max);
        lse = logf(lse) + logit_max;
        for (int i = 0; i < n_logits; ++i) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
    }
    {
        float ts_logprob = -INFINITY;
        {
            float lse = 0.0f;
            const float lmax = *std::max_element(logprobs.begin() + vocab.token_beg, logprobs.end());
            for (int i = vocab.token_beg; i < n_logits; ++i) if (logprobs[i] > -INFINITY) lse += expf(logprobs[i] - lmax);
            if (lse > 0.0f) ts_logprob = logf(lse) + lmax;
        }
        const float max_text = *std::max_element(logprobs.begin(), logprobs.begin() + vocab.token_beg);
        if (ts_logprob > max_text)
            for (int i = 0; i < vocab.token_beg; ++i) { logits[i] = -INFINITY; logprobs[i] = -INFINITY; }
    }
    for (int i = 0; i < n_logits; ++i) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

static void ts_stats(const Vocab& vocab, const Vec& probs, TokenData& r) {
    double sum_ts = 0, max_ts = 0;
    for (int i = vocab.token_beg; i < vocab.n_vocab; ++i) {
        sum_ts += probs[i];
        if (max_ts < probs[i]) { max_ts = probs[i]; r.tid = i; }
    }
    r.pt = (float)(max_ts / (sum_ts + 1e-10));
    r.ptsum = (float)sum_ts;
}

static TokenData sample_token(wo_ctx* c, Decoder& dec, bool best) {
    const Vocab& vocab = c->model.vocab;
    TokenData r;
    ts_stats(vocab, dec.probs, r);
    if (best) {
        for (int i = 0; i < vocab.n_vocab; ++i)
            if (r.p < dec.probs[i]) { r.id = i; r.p = dec.probs[i]; r.plog = dec.logprobs[i]; }
    } else {
        std::discrete_distribution<> dist(dec.probs.begin(), dec.probs.end());
        r.id = dist(dec.rng);
        r.p = dec.probs[r.id];
        r.plog = dec.logprobs[r.id];
    }
    if (r.id >= vocab.token_beg) { r.tid = r.id; r.pt = r.p; }
    return r;
}

static std::vector<TokenData> sample_topk(wo_ctx* c, Decoder& dec, int k, bool sampled) {
    const Vocab& vocab = c->model.vocab;
    TokenData base;
    base.tid = vocab.token_beg;
    ts_stats(vocab, dec.probs, base);
    std::vector<TokenData> out;
    if (sampled) {
        std::discrete_distribution<> dist(dec.probs.begin(), dec.probs.end());
        for (int i = 0; i < k; ++i) {
            TokenData r = base;
            r.id = dist(dec.rng);
            r.p = dec.probs[r.id]; r.plog = dec.logprobs[r.id];
            if (r.id >= vocab.token_beg) { r.tid = r.id; r.pt = r.p; }
            out.push_back(r);
        }
    } else {
        // deterministic top-k by log-probability, ties -> lower id (BASELINE.json north star)
        std::vector<int> idx(vocab.n_vocab);
        for (int i = 0; i < vocab.n_vocab; ++i) idx[i] = i;
        k = std::min(k, vocab.n_vocab);
        std::partial_sort(idx.begin(), idx.begin() + k, idx.end(), [&](int a, int b) {
            if (dec.logprobs[a] != dec.logprobs[b]) return dec.logprobs[a] > dec.logprobs[b];
            return a < b;
        });
        for (int i = 0; i < k; ++i) {
            if (!(dec.logprobs[idx[i]] > -INFINITY)) break;
            TokenData r = base;
            r.id = idx[i];
            r.p = dec.probs[r.id]; r.plog = dec.logprobs[r.id];
            if (r.id >= vocab.token_beg) { r.tid = r.id; r.pt = r.p; }
            out.push_back(r);
        }
    }
    return out;
}

static void sequence_score(const wo_params& p, Sequence& s) {
    if (s.result_len == 0) return;
    double result = 0;
    for (int i = 0; i < s.result_len; ++i) result += s.tokens[i].plog;
    s.sum_logprobs = result;
    s.avg_logprobs = result / s.result_len;
    double penalty = s.result_len;
    if (p.length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, p.length_penalty);
    s.score = result / penalty;
    const int n = 32;
    int cnt = 0;
    double entropy = 0;
    std::map<int, int> counts;
    for (int i = std::max(0, s.result_len - n); i < s.result_len; ++i) { counts[s.tokens[i].id]++; cnt++; }
    for (auto& kv : counts) { const double q = kv.second / (double)cnt; entropy -= q * log(q); }
    s.entropy = entropy;
}

static bool seq_equal(const Sequence& a, const Sequence& b) {
    if (a.tokens.size() != b.tokens.size()) return false;
    for (int i = (int)a.tokens.size() - 1; i >= 0; --i) if (a.tokens[i].id != b.tokens[i].id) return false;
    return true;
}

// language auto-detect: encode @seek 0, decode [sot], softmax over
// the language tokens, argmax.
int wo_lang_detect(wo_ctx* c, float* probs_out) {
    const Vocab& vocab = c->model.vocab;
    if (wo_encode(c, 0) != 0) return -1;
    int sot = vocab.token_sot;
    if (wo_decode(c, &sot, 1, 0, 0) != 0) return -2;
    std::vector<std::pair<float, int>> li;
    for (int i = 0; i < N_LANGS; ++i) li.emplace_back(c->logits[vocab.token_lang(i)], i);
    std::stable_sort(li.begin(), li.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
    const float mx = li[0].first;
    double sum = 0;
    for (auto& kv : li) sum += expf(kv.first - mx);
    if (probs_out) for (auto& kv : li) probs_out[kv.second] = (float)(expf(kv.first - mx) / sum);
    return li[0].second;
}

// ---- whisper_full_with_state
constexpr int kDeltaMin = 10;  // frames (100 ms): minimum audio whisper_full still processes

int wo_full(wo_ctx* c, const wo_params* pp, const float* pcm, int n_samples) {
    wo_params p = *pp;
    const Vocab& vocab = c->model.vocab;
    const HParams& hp = c->model.hp;
    c->result_all.clear();
    if (n_samples > 0) log_mel(pcm, n_samples, c->model, c->mel);
    std::string lang = p.language ? p.language : "";
    if (lang.empty() || lang == "auto") {
        int id = wo_lang_detect(c, nullptr);
        if (id < 0) return -3;
        lang = LANGS[id];
    }
    const int seek_start = 0;
    const int seek_end = c->mel.n_len_org;
    // shorter than 100 ms (delta_min = 10 frames): nothing to do.  whisper.cpp v1.7.x (the snapshot whisper-rs-sys 0.14.1 vendors)
    // lowered this from 1 s; the reference app sends every clip longer than 1600 samples (state.rs:749)
    if (seek_end < seek_start + kDeltaMin) return 0;

    std::vector<float> temperatures;
    if (p.temperature_inc > 0.0f) for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) temperatures.push_back(t);
    else temperatures.push_back(p.temperature);

    int n_decoders = 1;
    if (p.strategy == 0) n_decoders = p.best_of; else n_decoders = std::max(p.best_of, p.beam_size);
    n_decoders = std::max(1, n_decoders);
    if (n_decoders > 8) return -4;
    for (auto& dcd : c->decoders) dcd.rng = std::mt19937(0);  // fresh state per call (whisper.rs:83-85)

    auto& prompt_past = c->prompt_past;
    prompt_past.clear();  // fresh state
    if (p.initial_prompt) {
        auto pt = tokenize(vocab, p.initial_prompt);
        if (!pt.empty()) {
            prompt_past.insert(prompt_past.end(), pt.begin(), pt.end());
            std::rotate(prompt_past.begin(), prompt_past.end() - pt.size(), prompt_past.end());
        }
    }
    std::vector<int> prompt_init = {vocab.token_sot};
    if (vocab.is_multilingual()) {
        int lid = lang_id(lang.c_str());
        if (lid < 0) return -5;
        c->lang_id = lid;
        prompt_init.push_back(vocab.token_lang(lid));
        prompt_init.push_back(p.translate ? vocab.token_translate : vocab.token_transcribe);
    }
    {
        const bool is_distil = hp.n_text_layer == 2 && hp.n_vocab != 51866;
        if (is_distil && !p.no_timestamps) p.no_timestamps = 1;
    }
    if (p.no_timestamps) prompt_init.push_back(vocab.token_not);

    int seek = seek_start;
    std::vector<int> prompt;
    struct Cand { int decoder_idx; int seek_delta; bool has_ts; Sequence sequence; };
    std::vector<std::vector<Cand>> bc_per_dec(n_decoders);
    std::vector<Cand> cands;

    while (true) {
        if (seek + kDeltaMin >= seek_end) break;
        if (wo_encode(c, seek) != 0) return -6;
        if (seek > seek_start && seek + 500 >= seek_end) prompt_past.clear();
        int best_decoder_id = 0;

        for (int it = 0; it < (int)temperatures.size(); ++it) {
            const float t_cur = temperatures[it];
            int n_cur = 1;
            if (p.strategy == 0) { if (t_cur > 0.0f) n_cur = p.best_of; }
            else { if (t_cur > 0.0f) n_cur = p.best_of; else n_cur = p.beam_size; }
            n_cur = std::max(1, n_cur);

            for (int j = 0; j < n_cur; ++j) {
                Decoder& d = c->decoders[j];
                d.sequence = Sequence();
                d.seek_delta = 100 * CHUNK_SIZE;
                d.failed = d.completed = d.has_ts = false;
            }
            {
                prompt.clear();
                if (!prompt_past.empty() && t_cur < 0.5f && p.n_max_text_ctx > 0) {
                    int n_take = std::min(std::min(p.n_max_text_ctx, hp.n_text_ctx / 2), (int)prompt_past.size());
                    prompt = {vocab.token_prev};
                    prompt.insert(prompt.begin() + 1, prompt_past.end() - n_take, prompt_past.end());
                }
                prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
                if (wo_decode(c, prompt.data(), (int)prompt.size(), 0, 0) != 0) return -7;
                if (c->logits_hook) c->logits_hook(c->logits_hook_user, seek, it, 0, 0, (int)prompt.size(), hp.n_vocab, c->logits.data());
                {
                    // no_speech_prob: softmax of the raw logits at the nosp token
                    const Vec& lg = c->logits;
                    const float mx = *std::max_element(lg.begin(), lg.end());
                    float lse = 0.0f;
                    for (float v : lg) lse += expf(v - mx);
                    lse = logf(lse) + mx;
                    c->no_speech_prob = expf(lg[vocab.token_nosp] - lse);
                }
                process_logits(c, c->decoders[0], p, t_cur);
                for (int j = 1; j < n_cur; ++j) {
                    kv_copy(c, 0, j);
                    c->decoders[j].probs = c->decoders[0].probs;
                    c->decoders[j].logits = c->decoders[0].logits;
                    c->decoders[j].logprobs = c->decoders[0].logprobs;
                }
            }
            const int n_max = hp.n_text_ctx / 2 - 4;
            for (int i = 0; i < n_max; ++i) {
                if (p.strategy == 1) for (auto& bc : bc_per_dec) bc.clear();
                for (int j = 0; j < n_cur; ++j) {
                    Decoder& d = c->decoders[j];
                    if (d.completed || d.failed) continue;
                    if (p.strategy == 0) {
                        d.sequence.tokens.push_back(sample_token(c, d, t_cur < 1e-6f));
                        d.sequence.sum_logprobs_all += d.sequence.tokens.back().plog;
                    } else {
                        auto toks = sample_topk(c, d, p.beam_size, p.beam_sampled != 0);
                        for (auto& t : toks) {
                            bc_per_dec[j].push_back({j, d.seek_delta, d.has_ts, d.sequence});
                            bc_per_dec[j].back().sequence.tokens.push_back(t);
                            bc_per_dec[j].back().sequence.sum_logprobs_all += t.plog;
                        }
                    }
                }
                if (p.strategy == 1) {
                    cands.clear();
                    for (auto& bc : bc_per_dec) cands.insert(cands.end(), bc.begin(), bc.end());
                    std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
                        if (a.sequence.sum_logprobs_all != b.sequence.sum_logprobs_all) return a.sequence.sum_logprobs_all > b.sequence.sum_logprobs_all;
                        if (a.decoder_idx != b.decoder_idx) return a.decoder_idx < b.decoder_idx;
                        return a.sequence.tokens.back().id < b.sequence.tokens.back().id;
                    });
                    size_t cur_c = 0;
                    std::vector<int> src(n_cur, -1);
                    for (int j = 0; j < n_cur; ++j) {
                        Decoder& d = c->decoders[j];
                        if (d.completed || d.failed) continue;
                        if (cands.empty()) { d.failed = true; continue; }
                        if (cur_c >= cands.size()) cur_c = 0;
                        Cand& cur = cands[cur_c++];
                        while (cands.size() > cur_c && seq_equal(cands[cur_c].sequence, cur.sequence) && i > 0) ++cur_c;
                        d.seek_delta = cur.seek_delta;
                        d.has_ts = cur.has_ts;
                        d.sequence = cur.sequence;
                        src[j] = cur.decoder_idx;
                    }
                    // KV reassignment through scratch sequence ids 8+j
                    for (int j = 0; j < n_cur; ++j) if (src[j] >= 0) kv_copy(c, src[j], 8 + j);
                    for (int j = 0; j < n_cur; ++j) if (src[j] >= 0) kv_copy(c, 8 + j, j);
                }
                for (int j = 0; j < n_cur; ++j) {
                    Decoder& d = c->decoders[j];
                    if (d.completed || d.failed) continue;
                    int& result_len = d.sequence.result_len;
                    const TokenData& tok = d.sequence.tokens.back();
                    if (tok.id > vocab.token_beg) {
                        const int sd_new = 2 * (tok.id - vocab.token_beg);
                        if (d.has_ts && d.seek_delta > sd_new && result_len < i) { d.failed = true; continue; }
                        d.seek_delta = sd_new;
                        result_len = i + 1;
                        d.has_ts = true;
                    }
                    if (tok.id == vocab.token_eot || (p.max_tokens > 0 && i >= p.max_tokens) ||
                        (d.has_ts && seek + d.seek_delta + kDeltaMin >= seek_end)) {
                        if (result_len == 0 && !p.no_timestamps) {
                            if (seek + d.seek_delta + kDeltaMin >= seek_end) result_len = i + 1;
                            else { d.failed = true; continue; }
                        }
                        if (p.single_segment || p.no_timestamps) { result_len = i + 1; d.seek_delta = 100 * CHUNK_SIZE; }
                        d.completed = true;
                        continue;
                    }
                    if (i == n_max - 1 && (result_len == 0 || d.seek_delta < 100 * CHUNK_SIZE / 2)) { d.failed = true; continue; }
                }
                {
                    bool all = true;
                    for (int j = 0; j < n_cur; ++j) if (!(c->decoders[j].completed || c->decoders[j].failed)) all = false;
                    if (all) break;
                }
                {
                    const int n_past = (int)prompt.size() + i;
                    for (int j = 0; j < n_cur; ++j) {
                        Decoder& d = c->decoders[j];
                        if (d.failed || d.completed) continue;
                        int tk = d.sequence.tokens.back().id;
                        if (wo_decode(c, &tk, 1, n_past, j) != 0) return -8;
                        if (c->logits_hook) c->logits_hook(c->logits_hook_user, seek, it, i + 1, j, (int)prompt.size(), hp.n_vocab, c->logits.data());
                        process_logits(c, d, p, t_cur);
                    }
                }
            }
            {
                double best_score = -INFINITY;
                for (int j = 0; j < n_cur; ++j) {
                    Decoder& d = c->decoders[j];
                    if (d.failed) continue;
                    d.sequence.tokens.resize(d.sequence.result_len);
                    sequence_score(p, d.sequence);
                    if (d.sequence.result_len > 32 && d.sequence.entropy < p.entropy_thold) { d.failed = true; c->n_fail_h++; continue; }
                    if (best_score < d.sequence.score) { best_score = d.sequence.score; best_decoder_id = j; }
                }
            }
            bool success = true;
            if (it != (int)temperatures.size() - 1) {
                const Decoder& d = c->decoders[best_decoder_id];
                if (d.failed || (d.sequence.avg_logprobs < p.logprob_thold && c->no_speech_prob < p.no_speech_thold)) {
                    success = false;
                    c->n_fail_p++;
                }
            }
            if (success) break;
        }
        {
            const Decoder& best = c->decoders[best_decoder_id];
            int seek_delta = best.seek_delta;
            const int result_len = best.sequence.result_len;
            const auto& tc = best.sequence.tokens;
            const bool is_no_speech = c->no_speech_prob > p.no_speech_thold && best.sequence.avg_logprobs < p.logprob_thold;
            prompt_past.clear();
            if (!prompt.empty() && prompt.front() == vocab.token_prev)
                prompt_past.insert(prompt_past.end(), prompt.begin() + 1, prompt.end() - prompt_init.size());
            for (int i = 0; i < result_len && !is_no_speech; ++i) prompt_past.push_back(tc[i].id);
            if (!tc.empty() && !is_no_speech) {
                int i0 = 0;
                int64_t t0 = seek + 2 * (tc.front().tid - vocab.token_beg);
                std::string text;
                for (int i = 0; i < (int)tc.size(); ++i) {
                    if (tc[i].id < vocab.token_eot) text += vocab.id_to_token[tc[i].id];
                    if (tc[i].id > vocab.token_beg && !p.single_segment) {
                        const int64_t t1 = seek + 2 * (tc[i].tid - vocab.token_beg);
                        if (!text.empty()) {
                            Segment s{t0, t1, text, c->no_speech_prob, {}};
                            for (int j = i0; j <= i; ++j) s.tokens.push_back(tc[j]);
                            c->result_all.push_back(std::move(s));
                        }
                        text.clear();
                        while (i < (int)tc.size() && tc[i].id > vocab.token_beg) ++i;
                        --i;
                        t0 = t1;
                        i0 = i + 1;
                    }
                }
                if (!text.empty()) {
                    Segment s{t0, (int64_t)seek + seek_delta, text, c->no_speech_prob, {}};
                    for (int j = i0; j < (int)tc.size(); ++j) s.tokens.push_back(tc[j]);
                    c->result_all.push_back(std::move(s));
                }
            }
            const bool single_ts_ending = tc.size() > 1 && tc[tc.size() - 2].id < vocab.token_beg && tc[tc.size() - 1].id > vocab.token_beg;
            if (single_ts_ending) seek_delta = std::min(seek_end - seek, CHUNK_SIZE * 100);
            seek += seek_delta;
        }
    }
    return 0;
}

int wo_n_segments(wo_ctx* c) { return (int)c->result_all.size(); }
const char* wo_segment_text(wo_ctx* c, int i) { return c->result_all[i].text.c_str(); }
int wo_segment_text_len(wo_ctx* c, int i) { return (int)c->result_all[i].text.size(); }
long wo_segment_t0(wo_ctx* c, int i) { return (long)c->result_all[i].t0; }
long wo_segment_t1(wo_ctx* c, int i) { return (long)c->result_all[i].t1; }
int wo_segment_n_tokens(wo_ctx* c, int i) { return (int)c->result_all[i].tokens.size(); }
int wo_segment_token_id(wo_ctx* c, int i, int j) { return c->result_all[i].tokens[j].id; }
int wo_segment_token_tid(wo_ctx* c, int i, int j) { return c->result_all[i].tokens[j].tid; }
float wo_segment_token_plog(wo_ctx* c, int i, int j) { return c->result_all[i].tokens[j].plog; }
float wo_no_speech_prob(wo_ctx* c) { return c->no_speech_prob; }
int wo_lang_id(wo_ctx* c) { return c->lang_id; }
void wo_stats(wo_ctx* c, long* out5) {
    out5[0] = c->n_encode; out5[1] = c->n_decode_calls; out5[2] = c->n_decode_tokens; out5[3] = c->n_fail_p; out5[4] = c->n_fail_h;
}
// single-step helper for stage parity: run the filters for an explicit history.
// hist[n_hist] = tokens sampled so far; returns logprobs/probs of decoder 0.
int wo_process_logits(wo_ctx* c, const wo_params* p, const int* hist, int n_hist, int has_ts, int seek_delta, float temperature,
                      float* logprobs_out, float* probs_out) {
    Decoder& d = c->decoders[0];
    d.sequence.tokens.clear();
    for (int i = 0; i < n_hist; ++i) { TokenData t; t.id = hist[i]; d.sequence.tokens.push_back(t); }
    d.has_ts = has_ts != 0;
    d.seek_delta = seek_delta;
    process_logits(c, d, *p, temperature);
    if (logprobs_out) memcpy(logprobs_out, d.logprobs.data(), d.logprobs.size() * 4);
    if (probs_out) memcpy(probs_out, d.probs.data(), d.probs.size() * 4);
    return 0;
}
void wo_set_logits_hook(wo_ctx* c, int (*hook)(void*, int, int, int, int, int, int, float*), void* user) {
    c->logits_hook = hook;
    c->logits_hook_user = user;
}
void wo_set_logits(wo_ctx* c, const float* lg) { c->logits.assign(lg, lg + c->model.hp.n_vocab); }
// draw `n` doubles exactly as std::discrete_distribution draws them from mt19937(0):
// std::generate_canonical<double,53>.  Used to pin the engine's host-side replica.
void wo_canonical_stream(double* out, int n) {
    std::mt19937 rng(0);
    for (int i = 0; i < n; ++i) out[i] = std::generate_canonical<double, 53>(rng);
}
// sample from explicit probabilities with a fresh mt19937(0), n draws
void wo_sample_stream(const float* probs, int n_probs, int* out, int n) {
    std::mt19937 rng(0);
    for (int i = 0; i < n; ++i) {
        std::discrete_distribution<> dist(probs, probs + n_probs);
        out[i] = dist(rng);
    }
}

}  // extern "C"
