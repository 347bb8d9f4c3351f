// Silence chunker of the reference (src-tauri/src/audio.rs:364-507), the step that cuts a long 16 kHz
// recording into the <= 30-s pieces the transcription path consumes (state.rs:757-780).  SURVEY.md §8f row N3.
//
//   GPU: RMS of every 20-ms window of the recording (HBM-bound: each sample is read once).  The reference adds
//        the squares of a window sequentially in float32 (`iter().map(|s| s * s).sum()`), so one thread owns one
//        window and adds in sample order with non-fused multiply / add: the values are bit-identical to the
//        reference's, which matters because they are compared against a threshold.  A thread's 16-byte loads
//        walk its window; the 128-byte lines they touch stay in L1 until consumed, so DRAM sees each byte once.
//   host: noise-floor estimate (10th percentile of the first 25 windows), adaptive threshold, the run-length
//        scan over the windows (180 k windows for an hour of audio) and the overlap-extended chunk ranges.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"
#include "kernels.cuh"

namespace nobs {
void set_last_error(const std::string& e);

namespace {

// audio.rs:337-360
constexpr float kSilenceThreshold = 0.01f;
constexpr uint32_t kMinSilenceDurationMs = 700;
constexpr uint32_t kMinChunkDurationMs = 1000;
constexpr float kAdaptiveThresholdNoiseFactor = 3.0f;
constexpr float kMinThresholdFactor = 0.5f;
constexpr size_t kNoiseFloorEstimationWindows = 25;
constexpr float kNoiseFloorPercentile = 0.1f;
constexpr float kMinNoiseFloorFactor = 0.3f;
constexpr uint32_t kChunkOverlapMs = 200;   // audio.rs:15

// audio.rs:364-370 for every full window [i*w, (i+1)*w)
__global__ void __launch_bounds__(128) window_rms_kernel(const float* __restrict__ pcm, size_t n_win, int w, float* __restrict__ rms, int vec_ok) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_win) return;
    const float* p = pcm + i * (size_t)w;
    float s = 0.0f;
    if (vec_ok) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int n4 = w >> 2;
#pragma unroll 4
        for (int k = 0; k < n4; ++k) {
            const float4 v = __ldg(p4 + k);
            s = __fadd_rn(s, __fmul_rn(v.x, v.x));
            s = __fadd_rn(s, __fmul_rn(v.y, v.y));
            s = __fadd_rn(s, __fmul_rn(v.z, v.z));
            s = __fadd_rn(s, __fmul_rn(v.w, v.w));
        }
    } else {
        for (int k = 0; k < w; ++k) {
            const float v = __ldg(p + k);
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
    }
    rms[i] = __fsqrt_rn(__fdiv_rn(s, (float)w));
}

struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
};

// RMS of every full window; `audio` may be host or device memory (cudaMemcpyDefault)
bool window_rms(const float* audio, size_t n, uint32_t w, std::vector<float>& out) {
    out.clear();
    if (w == 0) { set_last_error("window_rms: zero window"); return false; }
    const size_t n_win = n / w;
    if (n_win == 0) return true;
    cudaPointerAttributes attr{};
    const bool on_device = cudaPointerGetAttributes(&attr, audio) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    DevMem dpcm, drms;
    const float* src = audio;
    cudaStream_t s = nullptr;
    if (!on_device) {
        if (cudaMalloc(&dpcm.p, n_win * w * sizeof(float)) != cudaSuccess) { set_last_error("window_rms: cudaMalloc failed (is there a GPU?)"); return false; }
        if (cudaMemcpyAsync(dpcm.p, audio, n_win * w * sizeof(float), cudaMemcpyDefault, s) != cudaSuccess) { set_last_error("window_rms: H2D copy failed"); return false; }
        src = static_cast<const float*>(dpcm.p);
    }
    if (cudaMalloc(&drms.p, n_win * sizeof(float)) != cudaSuccess) { set_last_error("window_rms: cudaMalloc failed (is there a GPU?)"); return false; }
    const int vec_ok = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    window_rms_kernel<<<(unsigned)((n_win + 127) / 128), 128, 0, s>>>(src, n_win, (int)w, static_cast<float*>(drms.p), vec_ok);
    count_launch();
    out.resize(n_win);
    if (cudaMemcpyAsync(out.data(), drms.p, n_win * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess ||
        cudaGetLastError() != cudaSuccess) {
        set_last_error("window_rms: kernel or D2H copy failed");
        return false;
    }
    return true;
}

// audio.rs:373-395 on the window RMS values
float estimate_noise_floor(const std::vector<float>& rms) {
    std::vector<float> v(rms.begin(), rms.begin() + std::min(rms.size(), kNoiseFloorEstimationWindows));
    if (v.empty()) return kSilenceThreshold;
    std::sort(v.begin(), v.end());
    const size_t idx = (size_t)((float)v.size() * kNoiseFloorPercentile);
    const float nf = idx < v.size() ? v[idx] : kSilenceThreshold;
    return std::max(nf, kSilenceThreshold * kMinNoiseFloorFactor);
}

// audio.rs:400-463
void scan_boundaries(const std::vector<float>& rms, size_t n, uint32_t sample_rate, std::vector<size_t>& boundaries) {
    boundaries.clear();
    const size_t min_silence = (size_t)(sample_rate * kMinSilenceDurationMs / 1000);
    const size_t min_chunk = (size_t)(sample_rate * kMinChunkDurationMs / 1000);
    const size_t w = sample_rate / 50;
    const float threshold = std::max(estimate_noise_floor(rms) * kAdaptiveThresholdNoiseFactor, kSilenceThreshold * kMinThresholdFactor);
    size_t last_boundary = 0, silence_start = 0;
    bool in_silence = false;
    auto try_add = [&](size_t start, size_t end) {
        const size_t dur = end - start;
        if (dur >= min_silence) {
            const size_t split = start + dur / 2;
            if (split - last_boundary >= min_chunk) { boundaries.push_back(split); last_boundary = split; }
        }
    };
    size_t pos = 0;
    for (size_t i = 0; i < rms.size(); ++i, pos += w) {
        if (rms[i] < threshold) {
            if (!in_silence) { in_silence = true; silence_start = pos; }
        } else {
            if (in_silence) try_add(silence_start, pos);
            in_silence = false;
        }
    }
    if (in_silence) try_add(silence_start, n);
}

}  // namespace
}  // namespace nobs

using namespace nobs;

extern "C" {

int whisper_b200_window_rms(const float* audio, size_t n_samples, uint32_t window, float* rms_out, size_t cap, size_t* n_windows) {
    if ((!audio && n_samples) || !n_windows) return -1;
    std::vector<float> rms;
    if (!window_rms(audio, n_samples, window, rms)) return -100;
    *n_windows = rms.size();
    if (rms_out) std::copy(rms.begin(), rms.begin() + std::min(cap, rms.size()), rms_out);
    return 0;
}

int nobs_find_silence_boundaries(const float* audio, size_t n_samples, uint32_t sample_rate, size_t* boundaries, size_t cap, size_t* n_found) {
    if ((!audio && n_samples) || !n_found || sample_rate < 50) return -1;
    std::vector<float> rms;
    if (!window_rms(audio, n_samples, sample_rate / 50, rms)) return -100;
    std::vector<size_t> b;
    scan_boundaries(rms, n_samples, sample_rate, b);
    *n_found = b.size();
    if (boundaries) std::copy(b.begin(), b.begin() + std::min(cap, b.size()), boundaries);
    return 0;
}

// audio.rs:473-507: chunk k is audio[ranges[2k] .. ranges[2k+1]); at most n_boundaries + 1 chunks
int nobs_split_at_silences_with_overlap(size_t n_samples, const size_t* boundaries, size_t n_boundaries, uint32_t sample_rate, size_t* ranges, size_t* n_chunks) {
    if (!ranges || !n_chunks || (n_boundaries && !boundaries)) return -1;
    size_t k = 0;
    if (n_boundaries == 0) {
        ranges[0] = 0; ranges[1] = n_samples;
        *n_chunks = 1;
        return 0;
    }
    const size_t overlap = (size_t)(sample_rate * kChunkOverlapMs / 1000);
    size_t start = 0;
    for (size_t i = 0; i < n_boundaries; ++i) {
        const size_t b = boundaries[i];
        if (b > start && b < n_samples) {
            ranges[2 * k] = start > overlap ? start - overlap : 0;
            ranges[2 * k + 1] = b;
            ++k;
            start = b;
        }
    }
    if (start < n_samples) {
        ranges[2 * k] = start > overlap ? start - overlap : 0;
        ranges[2 * k + 1] = n_samples;
        ++k;
    }
    *n_chunks = k;
    return 0;
}

int nobs_split_at_silences(size_t n_samples, const size_t* boundaries, size_t n_boundaries, size_t* ranges, size_t* n_chunks) {
    return nobs_split_at_silences_with_overlap(n_samples, boundaries, n_boundaries, 16000, ranges, n_chunks);   // audio.rs:467-469
}

}  // extern "C"
