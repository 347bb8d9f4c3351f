#include "model.h"

#include <cstdint>

#include <algorithm>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace nobs {

namespace {

struct LangEntry {
    const char* code;
    const char* name;
};
// id == index; the reference UI offers a subset (reference src/routes/+page.svelte:49-58),
// "auto" is the config default (reference src-tauri/src/config.rs:49).
const LangEntry kLangs[kNumLangs] = {
    {"en", "english"}, {"zh", "chinese"}, {"de", "german"}, {"es", "spanish"}, {"ru", "russian"}, {"ko", "korean"},
    {"fr", "french"}, {"ja", "japanese"}, {"pt", "portuguese"}, {"tr", "turkish"}, {"pl", "polish"}, {"ca", "catalan"},
    {"nl", "dutch"}, {"ar", "arabic"}, {"sv", "swedish"}, {"it", "italian"}, {"id", "indonesian"}, {"hi", "hindi"},
    {"fi", "finnish"}, {"vi", "vietnamese"}, {"he", "hebrew"}, {"uk", "ukrainian"}, {"el", "greek"}, {"ms", "malay"},
    {"cs", "czech"}, {"ro", "romanian"}, {"da", "danish"}, {"hu", "hungarian"}, {"ta", "tamil"}, {"no", "norwegian"},
    {"th", "thai"}, {"ur", "urdu"}, {"hr", "croatian"}, {"bg", "bulgarian"}, {"lt", "lithuanian"}, {"la", "latin"},
    {"mi", "maori"}, {"ml", "malayalam"}, {"cy", "welsh"}, {"sk", "slovak"}, {"te", "telugu"}, {"fa", "persian"},
    {"lv", "latvian"}, {"bn", "bengali"}, {"sr", "serbian"}, {"az", "azerbaijani"}, {"sl", "slovenian"}, {"kn", "kannada"},
    {"et", "estonian"}, {"mk", "macedonian"}, {"br", "breton"}, {"eu", "basque"}, {"is", "icelandic"}, {"hy", "armenian"},
    {"ne", "nepali"}, {"mn", "mongolian"}, {"bs", "bosnian"}, {"kk", "kazakh"}, {"sq", "albanian"}, {"sw", "swahili"},
    {"gl", "galician"}, {"mr", "marathi"}, {"pa", "punjabi"}, {"si", "sinhala"}, {"km", "khmer"}, {"sn", "shona"},
    {"yo", "yoruba"}, {"so", "somali"}, {"af", "afrikaans"}, {"oc", "occitan"}, {"ka", "georgian"}, {"be", "belarusian"},
    {"tg", "tajik"}, {"sd", "sindhi"}, {"gu", "gujarati"}, {"am", "amharic"}, {"yi", "yiddish"}, {"lo", "lao"},
    {"uz", "uzbek"}, {"fo", "faroese"}, {"ht", "haitian creole"}, {"ps", "pashto"}, {"tk", "turkmen"}, {"nn", "nynorsk"},
    {"mt", "maltese"}, {"sa", "sanskrit"}, {"lb", "luxembourgish"}, {"my", "myanmar"}, {"bo", "tibetan"}, {"tl", "tagalog"},
    {"mg", "malagasy"}, {"as", "assamese"}, {"tt", "tatar"}, {"haw", "hawaiian"}, {"ln", "lingala"}, {"ha", "hausa"},
    {"ba", "bashkir"}, {"jw", "javanese"}, {"su", "sundanese"}, {"yue", "cantonese"},
};

float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal: renormalise
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, sizeof f);
    return f;
}

struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
    bool read(void* dst, size_t n) { return n == 0 || fread(dst, 1, n, f) == n; }
};

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_alpha(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_other(unsigned char c) { return !is_space(c) && !is_alpha(c) && !is_digit(c); }

}  // namespace

int lang_id(const char* lang) {
    if (!lang) return -1;
    for (int i = 0; i < kNumLangs; ++i)
        if (!strcmp(kLangs[i].code, lang)) return i;
    for (int i = 0; i < kNumLangs; ++i)
        if (!strcmp(kLangs[i].name, lang)) return i;
    return -1;
}
const char* lang_str(int id) { return (id >= 0 && id < kNumLangs) ? kLangs[id].code : nullptr; }
const char* lang_str_full(int id) { return (id >= 0 && id < kNumLangs) ? kLangs[id].name : nullptr; }

const HostTensor& HostModel::get(const std::string& name) const {
    auto it = tensors.find(name);
    if (it == tensors.end()) throw std::runtime_error("tensor not found: " + name);
    return it->second;
}

// ggml block quantisation (32 weights per block), as written by whisper.cpp's `quantize` tool.  Type ids are ggml's:
// 2 Q4_0 {f16 d; u8 qs[16]}, 3 Q4_1 {f16 d, m; qs[16]}, 6 Q5_0 {f16 d; u8 qh[4]; qs[16]}, 7 Q5_1 {f16 d, m; qh[4]; qs[16]},
// 8 Q8_0 {f16 d; i8 qs[32]}.  Element j of the low nibbles is weight j, of the high nibbles weight j + 16; the fifth bit of
// weight i is bit i of qh.
static int quant_block_bytes(int ttype) {
    switch (ttype) {
        case 2: return 18;
        case 3: return 20;
        case 6: return 22;
        case 7: return 24;
        case 8: return 34;
        default: return 0;
    }
}
static void dequantize_blocks(int ttype, const uint8_t* src, size_t n_blocks, float* dst) {
    const int bs = quant_block_bytes(ttype);
    for (size_t b = 0; b < n_blocks; ++b, src += bs, dst += 32) {
        uint16_t hd, hm = 0;
        memcpy(&hd, src, 2);
        const float d = half_to_float(hd);
        const uint8_t* q = src + 2;
        float m = 0.0f;
        if (ttype == 3 || ttype == 7) { memcpy(&hm, src + 2, 2); m = half_to_float(hm); q += 2; }
        if (ttype == 8) {
            for (int j = 0; j < 32; ++j) dst[j] = (float)(int8_t)q[j] * d;
            continue;
        }
        uint32_t qh = 0;
        if (ttype == 6 || ttype == 7) { memcpy(&qh, q, 4); q += 4; }
        for (int j = 0; j < 16; ++j) {
            int x0 = q[j] & 0x0F, x1 = q[j] >> 4;
            if (ttype == 6 || ttype == 7) {
                x0 |= (int)(((qh >> j) << 4) & 0x10);
                x1 |= (int)((qh >> (j + 12)) & 0x10);
            }
            switch (ttype) {
                case 2: dst[j] = (float)(x0 - 8) * d; dst[j + 16] = (float)(x1 - 8) * d; break;
                case 3: dst[j] = (float)x0 * d + m; dst[j + 16] = (float)x1 * d + m; break;
                case 6: dst[j] = (float)(x0 - 16) * d; dst[j + 16] = (float)(x1 - 16) * d; break;
                default: dst[j] = (float)x0 * d + m; dst[j + 16] = (float)x1 * d + m; break;
            }
        }
    }
}

bool load_ggml_model(const std::string& path, HostModel& m, std::string& err) {
    File fp;
    fp.f = fopen(path.c_str(), "rb");
    if (!fp.f) { err = "failed to open '" + path + "'"; return false; }
    uint32_t magic = 0;
    if (!fp.read(&magic, 4) || magic != 0x67676d6cu) { err = "invalid model file '" + path + "' (bad magic)"; return false; }
    if (!fp.read(&m.hp, sizeof(HParams))) { err = "truncated header"; return false; }
    HParams& hp = m.hp;
    hp.ftype %= 1000;  // upper digits carry the quantisation version
    if (hp.n_vocab <= 0 || hp.n_audio_ctx <= 0 || hp.n_audio_state <= 0 || hp.n_audio_head <= 0 || hp.n_audio_layer <= 0 ||
        hp.n_text_ctx <= 0 || hp.n_text_state <= 0 || hp.n_text_head <= 0 || hp.n_text_layer <= 0 || hp.n_mels <= 0) {
        err = "invalid hyper-parameters";
        return false;
    }
    if (hp.n_audio_state % hp.n_audio_head || hp.n_text_state % hp.n_text_head || hp.n_audio_state / hp.n_audio_head != 64 ||
        hp.n_text_state / hp.n_text_head != 64) {
        err = "unsupported head size (expected 64)";
        return false;
    }
    if (hp.n_audio_ctx != 1500 || hp.n_text_ctx != 448 || hp.n_audio_state != hp.n_text_state) {
        err = "unsupported context sizes";
        return false;
    }
    {   // whisper.cpp file types: 0 f32, 1 f16, 2 q4_0, 3 q4_1, 7 q8_0, 8 q5_0, 9 q5_1 (+ 1000 * quantisation version)
        const int ft = hp.ftype % 1000;
        if (ft != 0 && ft != 1 && ft != 2 && ft != 3 && ft != 7 && ft != 8 && ft != 9) { err = "unsupported ftype " + std::to_string(hp.ftype); return false; }
    }
    switch (hp.n_audio_layer) {
        case 4: m.mtype = 1; break;
        case 6: m.mtype = 2; break;
        case 12: m.mtype = 3; break;
        case 24: m.mtype = 4; break;
        case 32: m.mtype = 5; break;
        default: m.mtype = 0;
    }
    int32_t n_mel = 0, n_fft = 0;
    if (!fp.read(&n_mel, 4) || !fp.read(&n_fft, 4) || n_mel != hp.n_mels || n_fft != 201) { err = "invalid mel filterbank header"; return false; }
    m.filters.resize((size_t)n_mel * n_fft);
    if (!fp.read(m.filters.data(), m.filters.size() * sizeof(float))) { err = "truncated mel filterbank"; return false; }

    int32_t n_file_tokens = 0;
    if (!fp.read(&n_file_tokens, 4) || n_file_tokens <= 0 || n_file_tokens > hp.n_vocab) { err = "invalid vocabulary size"; return false; }
    Vocab& v = m.vocab;
    v.n_vocab = hp.n_vocab;
    v.id_to_token.reserve(hp.n_vocab);
    std::string word;
    for (int i = 0; i < n_file_tokens; ++i) {
        uint32_t len = 0;
        if (!fp.read(&len, 4) || len > (1u << 16)) { err = "invalid vocabulary entry"; return false; }
        word.resize(len);
        if (!fp.read(len ? &word[0] : nullptr, len)) { err = "truncated vocabulary"; return false; }
        v.id_to_token.push_back(word);
        v.token_to_id[word] = i;
    }
    if (v.is_multilingual()) {
        v.token_eot += 1;
        v.token_sot += 1;
        const int dt = v.num_languages() - 98;
        v.token_translate += dt; v.token_transcribe += dt; v.token_solm += dt; v.token_prev += dt;
        v.token_nosp += dt; v.token_not += dt; v.token_beg += dt;
    }
    for (int i = n_file_tokens; i < hp.n_vocab; ++i) {
        if (i > v.token_beg) word = "[_TT_" + std::to_string(i - v.token_beg) + "]";
        else if (i == v.token_eot) word = "[_EOT_]";
        else if (i == v.token_sot) word = "[_SOT_]";
        else if (i == v.token_translate) word = "[_TRANSLATE_]";
        else if (i == v.token_transcribe) word = "[_TRANSCRIBE_]";
        else if (i == v.token_solm) word = "[_SOLM_]";
        else if (i == v.token_prev) word = "[_PREV_]";
        else if (i == v.token_nosp) word = "[_NOSP_]";
        else if (i == v.token_not) word = "[_NOT_]";
        else if (i == v.token_beg) word = "[_BEG_]";
        else if (i > v.token_sot && i <= v.token_sot + v.num_languages()) word = std::string("[_LANG_") + lang_str(i - v.token_sot - 1) + "]";
        else word = "[_extra_token_" + std::to_string(i) + "]";
        v.id_to_token.push_back(word);
        v.token_to_id[word] = i;
    }
    for (const auto& t : v.id_to_token) v.max_token_len = std::max(v.max_token_len, t.size());
    {
        auto it = v.token_to_id.find(" ");
        v.token_blank = it == v.token_to_id.end() ? -1 : it->second;
    }
    if (v.token_beg >= hp.n_vocab) { err = "vocabulary too small for timestamp tokens"; return false; }

    // tensors until EOF
    size_t file_size = 0;
    {
        const long here = ftell(fp.f);
        fseek(fp.f, 0, SEEK_END);
        file_size = (size_t)std::max(0L, ftell(fp.f));
        fseek(fp.f, here, SEEK_SET);
    }
    std::vector<uint16_t> half;
    std::vector<uint8_t> raw;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        if (fread(&n_dims, 1, 4, fp.f) != 4) break;  // clean EOF
        if (!fp.read(&name_len, 4) || !fp.read(&ttype, 4)) { err = "truncated tensor header"; return false; }
        if (n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 512) { err = "corrupt tensor header"; return false; }
        const int qblock = quant_block_bytes(ttype);   // ggml block formats: 32 weights per block
        if (ttype != 0 && ttype != 1 && qblock == 0) { err = "unsupported tensor type " + std::to_string(ttype); return false; }
        int32_t ne[4] = {1, 1, 1, 1};
        size_t count = 1;
        for (int i = 0; i < n_dims; ++i) {
            if (!fp.read(&ne[i], 4) || ne[i] <= 0) { err = "corrupt tensor dims"; return false; }
            count *= (size_t)ne[i];
        }
        std::string name((size_t)name_len, '\0');
        if (!fp.read(&name[0], name_len)) { err = "truncated tensor name"; return false; }
        // bound the element count against what the file still holds BEFORE allocating: a damaged header must come back as
        // an error, not as std::bad_alloc / std::length_error or a multi-gigabyte allocation
        {
            const size_t on_disk_max = file_size > (size_t)ftell(fp.f) ? file_size - (size_t)ftell(fp.f) : 0;
            const size_t bytes_needed = qblock ? count / 32 * (size_t)qblock : count * (ttype == 0 ? 4 : 2);
            bool overflow = false;
            size_t chk = 1;
            for (int i = 0; i < n_dims; ++i) { if (chk > SIZE_MAX / 8 / (size_t)ne[i]) overflow = true; chk *= (size_t)ne[i]; }
            if (overflow || bytes_needed > on_disk_max) { err = "tensor " + name + " is larger than the rest of the file"; return false; }
        }
        HostTensor t;
        t.ttype = ttype;
        for (int i = n_dims - 1; i >= 0; --i) t.shape.push_back(ne[i]);  // file stores innermost first
        t.data.resize(count);
        if (qblock) {
            // quantised matrix (the catalogue's q5_0 / q5_1 / q8_0 files, reference model.rs:155-186): widened on load
            if (ne[0] % 32 != 0) { err = "quantised tensor with a row length that is not a multiple of 32: " + name; return false; }
            raw.resize(count / 32 * (size_t)qblock);
            if (!fp.read(raw.data(), raw.size())) { err = "truncated tensor data: " + name; return false; }
            dequantize_blocks(ttype, raw.data(), count / 32, t.data.data());
        } else if (ttype == 0) {
            if (!fp.read(t.data.data(), count * 4)) { err = "truncated tensor data: " + name; return false; }
        } else {
            half.resize(count);
            if (!fp.read(half.data(), count * 2)) { err = "truncated tensor data: " + name; return false; }
            for (size_t i = 0; i < count; ++i) t.data[i] = half_to_float(half[i]);
        }
        m.tensors.emplace(std::move(name), std::move(t));
    }
    // presence / size check of everything the engine will upload
    const size_t d = hp.n_audio_state;
    auto need = [&](const std::string& n, size_t count) {
        auto it = m.tensors.find(n);
        if (it == m.tensors.end()) { if (err.empty()) err = "missing tensor " + n; return; }
        if (it->second.data.size() != count && err.empty()) err = "wrong size for tensor " + n;
    };
    need("decoder.positional_embedding", (size_t)hp.n_text_ctx * d);
    need("encoder.positional_embedding", (size_t)hp.n_audio_ctx * d);
    need("decoder.token_embedding.weight", (size_t)hp.n_vocab * d);
    need("encoder.conv1.weight", d * hp.n_mels * 3);
    need("encoder.conv1.bias", d);
    need("encoder.conv2.weight", d * d * 3);
    need("encoder.conv2.bias", d);
    need("encoder.ln_post.weight", d);
    need("encoder.ln_post.bias", d);
    need("decoder.ln.weight", d);
    need("decoder.ln.bias", d);
    auto need_attn = [&](const std::string& p) {
        need(p + "_ln.weight", d); need(p + "_ln.bias", d);
        need(p + ".query.weight", d * d); need(p + ".query.bias", d);
        need(p + ".key.weight", d * d);
        need(p + ".value.weight", d * d); need(p + ".value.bias", d);
        need(p + ".out.weight", d * d); need(p + ".out.bias", d);
    };
    auto need_mlp = [&](const std::string& p) {
        need(p + "mlp_ln.weight", d); need(p + "mlp_ln.bias", d);
        need(p + "mlp.0.weight", 4 * d * d); need(p + "mlp.0.bias", 4 * d);
        need(p + "mlp.2.weight", 4 * d * d); need(p + "mlp.2.bias", d);
    };
    for (int i = 0; i < hp.n_audio_layer; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i) + ".";
        need_attn(p + "attn");
        need_mlp(p);
    }
    for (int i = 0; i < hp.n_text_layer; ++i) {
        const std::string p = "decoder.blocks." + std::to_string(i) + ".";
        need_attn(p + "attn");
        need_attn(p + "cross_attn");
        need_mlp(p);
    }
    return err.empty();
}

// Word split equivalent to the pattern
//   's|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+
// (byte-wise, "C" locale classes), written as a scanner; then greedy longest match per word.
std::vector<int> tokenize(const Vocab& vocab, const std::string& text) {
    std::vector<int> out;
    const size_t n = text.size();
    const unsigned char* s = reinterpret_cast<const unsigned char*>(text.data());
    size_t i = 0;
    auto emit = [&](size_t b, size_t e) {
        size_t p = b;
        while (p < e) {
            size_t len = std::min(e - p, vocab.max_token_len);
            bool found = false;
            for (; len > 0; --len) {
                auto it = vocab.token_to_id.find(text.substr(p, len));
                if (it != vocab.token_to_id.end()) {
                    out.push_back(it->second);
                    p += len;
                    found = true;
                    break;
                }
            }
            if (!found) ++p;  // unknown byte: skipped
        }
    };
    auto run = [&](size_t b, bool (*cls)(unsigned char)) {
        size_t e = b;
        while (e < n && cls(s[e])) ++e;
        return e;
    };
    while (i < n) {
        const unsigned char c = s[i];
        if (c == '\'' && i + 1 < n) {
            const char* contr[] = {"s", "t", "re", "ve", "m", "ll", "d"};
            bool hit = false;
            for (const char* t : contr) {
                size_t l = strlen(t);
                if (i + 1 + l <= n && !memcmp(s + i + 1, t, l)) {
                    emit(i, i + 1 + l);
                    i += 1 + l;
                    hit = true;
                    break;
                }
            }
            if (hit) continue;
        }
        const size_t j = (c == ' ' && i + 1 < n) ? i + 1 : i;  // optional leading space
        const unsigned char h = s[j];
        if (j > i || !is_space(c)) {
            // ' ?class+' alternatives (with or without the leading space)
            if (is_alpha(h)) { size_t e = run(j, is_alpha); emit(i, e); i = e; continue; }
            if (is_digit(h)) { size_t e = run(j, is_digit); emit(i, e); i = e; continue; }
            if (is_other(h)) { size_t e = run(j, is_other); emit(i, e); i = e; continue; }
        }
        // whitespace run: \s+(?!\S) gives the run minus its last char when a non-space follows
        size_t e = run(i, is_space);
        if (e < n && e - i > 1) --e;
        emit(i, e);
        i = e;
    }
    return out;
}

}  // namespace nobs
