// C++ host-side mirror of the reference's streaming capture buffer (reference src-tauri/src/audio.rs:29-244):
// adaptive noise floor, silence-boundary detection, chunk extraction with 200 ms overlap, forced split of long
// continuous speech.  This is the caller immediately in front of the transcription path while recording
// (state.rs:586-605: push_samples -> take_chunk_at_silence / take_forced_chunk -> transcribe).  The pushes are tens
// of milliseconds of audio per callback, so this stays on the host (a kernel launch costs more than the work);
// the arithmetic is float32 in the reference's order, bit for bit.  The batch path over whole recordings
// (find_silence_boundaries) runs on the GPU: csrc/audio_chunker.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace nobs {

float calculate_rms(const float* samples, size_t n);   // audio.rs:364-370

class AudioBuffer {
public:
    explicit AudioBuffer(uint32_t sample_rate = 48000);   // audio.rs:44-57 (new() = 48 kHz)
    void push_samples(const float* samples, size_t n);    // audio.rs:59-86
    std::vector<float> take();                            // audio.rs:88-92
    bool has_silence_boundary() const;                    // audio.rs:96-105
    bool take_chunk_at_silence(std::vector<float>& out);  // audio.rs:110-158   (false == None)
    bool take_forced_chunk(std::vector<float>& out);      // audio.rs:163-227
    size_t len() const { return samples_.size(); }        // audio.rs:230-232
    bool is_empty() const { return samples_.empty(); }
    float get_noise_floor() const { return noise_floor_; }
    size_t overlap_len() const { return overlap_.size(); }
    size_t last_speech_pos() const { return last_speech_pos_; }

private:
    void cut(size_t split_point, std::vector<float>& out);
    std::vector<float> samples_;
    size_t last_speech_pos_ = 0;
    uint32_t sample_rate_;
    float noise_floor_;
    size_t noise_floor_frames_ = 0;
    std::vector<float> overlap_;
};

}  // namespace nobs
