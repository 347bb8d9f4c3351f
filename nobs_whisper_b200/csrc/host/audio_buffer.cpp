#include "audio_buffer.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

#include "../../../include/whisper_b200.h"

namespace nobs {

namespace {
// audio.rs:9-15, 337-355
constexpr uint32_t kMaxBufferDurationS = 25;
constexpr uint32_t kChunkOverlapMs = 200;
constexpr float kSilenceThreshold = 0.01f;
constexpr uint32_t kMinSilenceDurationMs = 700;
constexpr size_t kNoiseFloorUpdateMaxFrames = 100;
constexpr float kAdaptiveThresholdNoiseFactor = 3.0f;
constexpr float kMinThresholdFactor = 0.5f;
constexpr float kNoiseFloorEmaDecay = 0.95f;
constexpr float kNoiseFloorUpdateThresholdFactor = 0.5f;
}  // namespace

float calculate_rms(const float* s, size_t n) {
    if (n == 0) return 0.0f;
    float sum = 0.0f;   // sequential float32 sum (built with -ffp-contract=off and without -ffast-math: no fusion, no reassociation)
    for (size_t i = 0; i < n; ++i) sum += s[i] * s[i];
    return std::sqrt(sum / (float)n);
}

AudioBuffer::AudioBuffer(uint32_t sample_rate) : sample_rate_(sample_rate), noise_floor_(kSilenceThreshold) {}

void AudioBuffer::push_samples(const float* samples, size_t n) {
    const size_t start_pos = samples_.size();
    samples_.insert(samples_.end(), samples, samples + n);
    const size_t w = sample_rate_ / 50;   // 20 ms windows
    if (w == 0) return;
    for (size_t i = 0, off = 0; off < n; ++i, off += w) {
        const float rms = calculate_rms(samples + off, std::min(w, n - off));   // slice::chunks: the last one may be short
        if (rms < noise_floor_ * kNoiseFloorUpdateThresholdFactor && noise_floor_frames_ < kNoiseFloorUpdateMaxFrames) {
            noise_floor_ = noise_floor_ * kNoiseFloorEmaDecay + rms * (1.0f - kNoiseFloorEmaDecay);
            ++noise_floor_frames_;
        }
        const float threshold = std::max(noise_floor_ * kAdaptiveThresholdNoiseFactor, kSilenceThreshold * kMinThresholdFactor);
        if (rms >= threshold) last_speech_pos_ = start_pos + (i + 1) * w;
    }
}

std::vector<float> AudioBuffer::take() {
    last_speech_pos_ = 0;
    overlap_.clear();
    std::vector<float> out;
    out.swap(samples_);
    return out;
}

bool AudioBuffer::has_silence_boundary() const {
    if (samples_.empty() || last_speech_pos_ == 0) return false;
    const size_t silence = samples_.size() > last_speech_pos_ ? samples_.size() - last_speech_pos_ : 0;
    return silence >= (size_t)(sample_rate_ * kMinSilenceDurationMs / 1000);
}

void AudioBuffer::cut(size_t split_point, std::vector<float>& out) {
    const size_t overlap_samples = (size_t)(sample_rate_ * kChunkOverlapMs / 1000);
    out.clear();
    out.reserve(overlap_.size() + split_point);
    out.insert(out.end(), overlap_.begin(), overlap_.end());
    out.insert(out.end(), samples_.begin(), samples_.begin() + split_point);
    const size_t overlap_start = split_point > overlap_samples ? split_point - overlap_samples : 0;
    overlap_.assign(samples_.begin() + overlap_start, samples_.begin() + split_point);
    samples_.erase(samples_.begin(), samples_.begin() + split_point);
}

bool AudioBuffer::take_chunk_at_silence(std::vector<float>& out) {
    if (!has_silence_boundary()) return false;
    if (last_speech_pos_ < (size_t)(sample_rate_ / 2)) return false;   // at least 0.5 s of speech
    const size_t silence_start = last_speech_pos_;
    const size_t split_point = silence_start + (samples_.size() - silence_start) / 2;
    cut(split_point, out);
    last_speech_pos_ = 0;
    return true;
}

bool AudioBuffer::take_forced_chunk(std::vector<float>& out) {
    const size_t max_samples = (size_t)sample_rate_ * kMaxBufferDurationS;
    if (samples_.size() <= max_samples) return false;
    const size_t search = (size_t)sample_rate_ * 5, w = sample_rate_ / 50;
    if (w == 0) return false;
    const size_t search_start = samples_.size() > search ? samples_.size() - search : 0;
    size_t quietest_pos = search_start;
    float quietest = FLT_MAX;
    for (size_t pos = search_start; pos + w <= samples_.size(); pos += w) {
        const float rms = calculate_rms(samples_.data() + pos, w);
        if (rms < quietest) { quietest = rms; quietest_pos = pos; }
    }
    const size_t split_point = std::min(quietest_pos + w / 2, samples_.size());
    if (split_point < (size_t)(sample_rate_ / 2)) return false;
    cut(split_point, out);
    last_speech_pos_ = last_speech_pos_ > split_point ? last_speech_pos_ - split_point : 0;
    return true;
}

}  // namespace nobs

// ---- C ABI (include/whisper_b200.h)
struct nobs_audio_buffer {
    nobs::AudioBuffer buf;
    std::vector<float> chunk;
    explicit nobs_audio_buffer(uint32_t sr) : buf(sr) {}
};

extern "C" {

float nobs_calculate_rms(const float* samples, size_t n) { return nobs::calculate_rms(samples, samples ? n : 0); }
struct nobs_audio_buffer* nobs_audio_buffer_new(uint32_t sample_rate) { return new nobs_audio_buffer(sample_rate); }
void nobs_audio_buffer_free(struct nobs_audio_buffer* b) { delete b; }
void nobs_audio_buffer_push_samples(struct nobs_audio_buffer* b, const float* samples, size_t n) { if (b && samples) b->buf.push_samples(samples, n); }
int nobs_audio_buffer_has_silence_boundary(const struct nobs_audio_buffer* b) { return b && b->buf.has_silence_boundary(); }
size_t nobs_audio_buffer_len(const struct nobs_audio_buffer* b) { return b ? b->buf.len() : 0; }
size_t nobs_audio_buffer_overlap_len(const struct nobs_audio_buffer* b) { return b ? b->buf.overlap_len() : 0; }
float nobs_audio_buffer_noise_floor(const struct nobs_audio_buffer* b) { return b ? b->buf.get_noise_floor() : 0.0f; }
static const float* hand_out(struct nobs_audio_buffer* b, bool ok, size_t* n) {
    if (!ok) { if (n) *n = 0; return nullptr; }
    if (n) *n = b->chunk.size();
    static const float kEmpty = 0.0f;
    return b->chunk.empty() ? &kEmpty : b->chunk.data();
}
const float* nobs_audio_buffer_take_chunk_at_silence(struct nobs_audio_buffer* b, size_t* n) {
    return hand_out(b, b && b->buf.take_chunk_at_silence(b->chunk), n);
}
const float* nobs_audio_buffer_take_forced_chunk(struct nobs_audio_buffer* b, size_t* n) {
    return hand_out(b, b && b->buf.take_forced_chunk(b->chunk), n);
}
const float* nobs_audio_buffer_take(struct nobs_audio_buffer* b, size_t* n) {
    if (!b) { if (n) *n = 0; return nullptr; }
    b->chunk = b->buf.take();
    return hand_out(b, true, n);
}

}  // extern "C"
