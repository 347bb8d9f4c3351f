// Sample-rate conversion to Whisper's 16 kHz (reference src-tauri/src/audio.rs:509-563 `resample_audio`, :329-334
// `resample_chunk`) — the step right before the transcription path on every recording.
//
// The reference calls rubato 0.15.0 (src-tauri/Cargo.lock:3896-3898) `FftFixedIn::<f32>::new(from, to, 1024, 2, 1)` with 1024-frame chunks (the last
// one zero-padded) and truncates to floor(n * to / from).  rubato is not vendored in the reference tree, so the
// algorithm below restates the crate's published design (PARITY UNPINNED beyond the reference's own length test,
// audio.rs:570-583; the CPU oracle oracle/audio_oracle.py::resample_audio restates the same design with numpy FFTs):
//   block sizes   fft_chunks = ceil(512 / (to / gcd)),  n_in = fft_chunks * from / gcd,  n_out = fft_chunks * to / gcd
//   filter        n_in-tap sinc low-pass, Blackman-Harris^2 window, unit DC gain, cutoff 0.4^(16 / n_in) (* n_out / n_in
//                 when decimating)
//   per block     zero-pad to 2 n_in, real FFT, multiply the first min(n_in + 1, n_out) bins by the filter spectrum,
//                 inverse real FFT of 2 n_out points, overlap-add the second half into the next block.
// B200 mapping: FFT -> multiply -> smaller inverse FFT of a block is ONE fixed linear map  y_blk[2 n_out] = M x_blk[n_in]
// with M[j][t] = g(j * from/gcd - t * to/gcd), g = the periodised impulse response of the filter on the common fine
// grid (one inverse FFT of length 2 n_in n_out / fft_chunks on the host, once per rate pair).  The recording viewed as
// [blocks][n_in] (in place, no copy) times M^T is a dense fp32 GEMM; an overlap-add kernel writes the 16 kHz stream.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"
#include "kernels.cuh"

namespace nobs {
void set_last_error(const std::string& e);

namespace {

using cd = std::complex<double>;

// in-place forward (sign = -1) / backward (sign = +1) DFT of any length: recursive mixed radix, O(n * sum of prime factors)
void fft_any(std::vector<cd>& a, int sign) {
    const size_t n = a.size();
    if (n <= 1) return;
    size_t p = 0;
    for (size_t f = 2; f * f <= n; ++f) if (n % f == 0) { p = f; break; }
    if (p == 0) {   // prime length: direct
        std::vector<cd> out(n);
        for (size_t k = 0; k < n; ++k) {
            cd s = 0;
            for (size_t t = 0; t < n; ++t) s += a[t] * std::polar(1.0, sign * 2.0 * M_PI * (double)((k * t) % n) / (double)n);
            out[k] = s;
        }
        a.swap(out);
        return;
    }
    const size_t m = n / p;
    std::vector<std::vector<cd>> sub(p, std::vector<cd>(m));
    for (size_t r = 0; r < p; ++r)
        for (size_t i = 0; i < m; ++i) sub[r][i] = a[i * p + r];
    for (size_t r = 0; r < p; ++r) fft_any(sub[r], sign);
    for (size_t k = 0; k < n; ++k) {
        cd s = 0;
        for (size_t r = 0; r < p; ++r) s += sub[r][k % m] * std::polar(1.0, sign * 2.0 * M_PI * (double)((r * k) % n) / (double)n);
        a[k] = s;
    }
}

struct Plan {
    uint32_t from = 0, to = 0;
    size_t n_in = 0, n_out = 0;
    size_t k_pad = 0;         // n_in rounded up to the GEMM's K step; the extra weight columns are zero
    float* M_dev = nullptr;   // [2 n_out][k_pad]
};

// rubato FftFixedIn::new(from, to, 1024, 2, 1) + FftResampler::new(n_in, n_out), folded into one matrix
bool build_plan(uint32_t from, uint32_t to, Plan& pl) {
    const uint32_t g = std::gcd(from, to);
    const size_t a = from / g, b = to / g;
    const size_t chunks = (size_t)std::ceil((float)(1024 / 2) / (float)b);
    const size_t n_in = chunks * a, n_out = chunks * b;
    // rate pairs with a tiny common divisor (44101 -> 16000) would need gigabyte tables; capture devices use the standard
    // rates (8 / 11.025 / 16 / 22.05 / 32 / 44.1 / 48 / 88.2 / 96 / 192 kHz), for which the block matrix stays below 40 MB
    if (2.0 * (double)n_out * (double)n_in > 4.0e7 || 2.0 * (double)a * (double)b * (double)chunks > 6.4e7) {
        set_last_error("resampler: unsupported rate pair " + std::to_string(from) + " -> " + std::to_string(to));
        return false;
    }
    const double cutoff = (double)std::pow(0.4f, 16.0f / (float)n_in) * (n_in > n_out ? (double)n_out / (double)n_in : 1.0);
    std::vector<double> h(n_in);
    double sum = 0.0;
    for (size_t x = 0; x < n_in; ++x) {
        const double xf = (double)x / (double)n_in;
        const double bh = 0.35875 - 0.48829 * cos(2 * M_PI * xf) + 0.14128 * cos(4 * M_PI * xf) - 0.01168 * cos(6 * M_PI * xf);
        const double u = ((double)x - (double)(n_in / 2)) * cutoff;
        const double sinc = u == 0.0 ? 1.0 : sin(M_PI * u) / (M_PI * u);
        h[x] = bh * bh * sinc;
        sum += h[x];
    }
    for (auto& v : h) v /= sum;
    // filter spectrum on the 2 n_in grid, bins 0 .. new_len-1
    const size_t new_len = n_in < n_out ? n_in + 1 : n_out;
    std::vector<cd> hf(2 * n_in, cd(0, 0));
    for (size_t x = 0; x < n_in; ++x) hf[x] = h[x] / (double)(2 * n_in);
    fft_any(hf, -1);
    // g(w) = Re( H[0] + 2 sum_{k>=1} H[k] e^{2 pi i k w / P} ),  P = 2 a b chunks  (w = j a - t b is the fine-grid offset)
    const size_t P = 2 * a * b * chunks;
    std::vector<cd> v(P, cd(0, 0));
    v[0] = cd(hf[0].real(), 0.0);
    for (size_t k = 1; k < new_len; ++k) v[k] = 2.0 * hf[k];
    fft_any(v, +1);
    const size_t k_pad = (n_in + 15) / 16 * 16;
    std::vector<float> M(2 * n_out * k_pad, 0.0f);
    for (size_t j = 0; j < 2 * n_out; ++j)
        for (size_t t = 0; t < n_in; ++t) {
            const long long w = (long long)(j * a) - (long long)(t * b);
            const size_t idx = (size_t)(((w % (long long)P) + (long long)P) % (long long)P);
            M[j * k_pad + t] = (float)v[idx].real();
        }
    float* dev = nullptr;
    if (cudaMalloc(&dev, M.size() * sizeof(float)) != cudaSuccess || cudaMemcpy(dev, M.data(), M.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("resampler: cannot upload the block matrix (is there a GPU?)");
        if (dev) cudaFree(dev);
        return false;
    }
    pl.from = from; pl.to = to; pl.n_in = n_in; pl.n_out = n_out; pl.k_pad = k_pad; pl.M_dev = dev;
    return true;
}

const Plan* get_plan(uint32_t from, uint32_t to) {
    static std::mutex mu;
    static std::map<std::tuple<int, uint32_t, uint32_t>, Plan> plans;   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(dev, from, to);
    auto it = plans.find(key);
    if (it != plans.end()) return &it->second;
    Plan pl;
    if (!build_plan(from, to, pl)) return nullptr;
    return &plans.emplace(key, pl).first->second;
}

// out[m * n_out + i] = C[m][i] + C[m-1][n_out + i]   (C rows are the 2 n_out-point block outputs)
__global__ void overlap_add_kernel(const float* __restrict__ C, size_t n_blocks, int n_out, float* __restrict__ out, size_t n_total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_total) return;
    const size_t m = idx / n_out;
    const int i = (int)(idx - m * n_out);
    float v = C[m * 2 * n_out + i];
    if (m > 0) v += C[(m - 1) * 2 * n_out + n_out + i];
    out[idx] = v;
}

struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
};

}  // namespace
}  // namespace nobs

using namespace nobs;

extern "C" {

int nobs_resample_audio(const float* audio, size_t n, uint32_t from_rate, uint32_t to_rate, float* out, size_t cap, size_t* n_out) {
    if ((!audio && n) || !n_out || from_rate == 0 || to_rate == 0) return -1;
    const Plan* pl = get_plan(from_rate, to_rate);
    if (!pl) return -100;
    const size_t padded = (n + 1023) / 1024 * 1024;                   // 1024-frame chunks, the last one zero-padded (audio.rs:531-538)
    const size_t n_blocks = padded / pl->n_in;                        // leftover frames are never flushed
    const size_t expected = (size_t)((double)n * ((double)to_rate / (double)from_rate));   // audio.rs:553-554
    const size_t produced = std::min(expected, n_blocks * pl->n_out);
    *n_out = produced;
    if (!out || produced == 0) return 0;
    if (cap < produced) return -1;
    DevMem x, c, y;
    cudaStream_t s = nullptr;
    // Rows of the recording are n_in frames apart.  When that pitch is 16-byte aligned the GEMM reads the recording in
    // place (K runs to k_pad over zero weights, + one zeroed K-step of slack after the last row); otherwise (e.g.
    // 22.05 kHz: 882-frame blocks) the rows are re-pitched to k_pad by a strided copy.
    const bool in_place = pl->n_in % 4 == 0;
    const size_t lda = in_place ? pl->n_in : pl->k_pad;
    const size_t x_elems = n_blocks * lda + 16;
    if (cudaMalloc(&x.p, x_elems * sizeof(float)) != cudaSuccess || cudaMalloc(&c.p, n_blocks * 2 * pl->n_out * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&y.p, produced * sizeof(float)) != cudaSuccess) {
        set_last_error("resampler: cudaMalloc failed");
        return -100;
    }
    const size_t used = std::min(n, n_blocks * pl->n_in);
    if (in_place) {
        cudaMemcpyAsync(x.p, audio, used * sizeof(float), cudaMemcpyDefault, s);
        cudaMemsetAsync(static_cast<float*>(x.p) + used, 0, (x_elems - used) * sizeof(float), s);
    } else {
        cudaMemsetAsync(x.p, 0, x_elems * sizeof(float), s);
        const size_t full_rows = used / pl->n_in, rest = used - full_rows * pl->n_in;
        if (full_rows) cudaMemcpy2DAsync(x.p, lda * sizeof(float), audio, pl->n_in * sizeof(float), pl->n_in * sizeof(float), full_rows, cudaMemcpyDefault, s);
        if (rest) cudaMemcpyAsync(static_cast<float*>(x.p) + full_rows * lda, audio + full_rows * pl->n_in, rest * sizeof(float), cudaMemcpyDefault, s);
    }
    Epilogue e;
    launch_gemm_f32(static_cast<const float*>(x.p), (int)lda, pl->M_dev, (int)pl->k_pad, static_cast<float*>(c.p), (int)(2 * pl->n_out), (int)n_blocks,
                    (int)(2 * pl->n_out), (int)pl->k_pad, e, s);
    overlap_add_kernel<<<(unsigned)((produced + 255) / 256), 256, 0, s>>>(static_cast<const float*>(c.p), n_blocks, (int)pl->n_out, static_cast<float*>(y.p), produced);
    count_launch();
    cudaMemcpyAsync(out, y.p, produced * sizeof(float), cudaMemcpyDefault, s);
    if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        set_last_error("resampler: kernel or copy failed");
        return -100;
    }
    return 0;
}

int nobs_resample_chunk(const float* audio, size_t n, uint32_t input_sample_rate, float* out, size_t cap, size_t* n_out) {
    if ((!audio && n) || !n_out) return -1;
    if (input_sample_rate == 16000) {   // audio.rs:330-332
        *n_out = n;
        if (!out) return 0;
        if (cap < n) return -1;
        return cudaMemcpy(out, audio, n * sizeof(float), cudaMemcpyDefault) == cudaSuccess ? 0 : -100;
    }
    return nobs_resample_audio(audio, n, input_sample_rate, 16000, out, cap, n_out);
}

// state.rs:590-594: mono = (sum of the frame's channels, left to right, float32) / channels
int nobs_mix_to_mono(const float* interleaved, size_t n_frames, uint32_t channels, float* out) {
    if (!interleaved || !out || channels == 0) return -1;
    for (size_t f = 0; f < n_frames; ++f) {
        float s = 0.0f;
        for (uint32_t c = 0; c < channels; ++c) s += interleaved[f * channels + c];
        out[f] = s / (float)channels;
    }
    return 0;
}

}  // extern "C"
