// CUDA-core kernels of the hot path (sm_100a): log-mel, fp32 parity-mode GEMM/attention,
// LayerNorm, decoder row kernels (embedding, KV scatter, KV-cache attention), and the fused
// logit filter / log-softmax / argmax / sampling / top-k kernel.
// The tcgen05 (tensor-core) kernels live in gemm_sm100.cu / attention_sm100.cu.
#include "kernels.cuh"
#include <cooperative_groups.h>

#include "device_utils.cuh"

#include <atomic>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <unordered_set>

namespace nobs {

bool g_use_pdl = [] { const char* e = getenv("NOBS_WHISPER_PDL"); return !(e && *e == '0'); }();
static std::atomic<long> g_launches{0};
long kernel_launch_count() { return g_launches.load(std::memory_order_relaxed); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void trace_set_kernels(unsigned long long* buf, unsigned int cap) {
    cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}

void prefer_max_shared_carveout(const void* kernel) {
    static std::mutex mu;
    static std::unordered_set<const void*> seen;
    static const bool enabled = [] { const char* v = getenv("NOBS_WHISPER_MAX_CARVEOUT"); return v && *v == '1'; }();  // measured: no effect on lane overlap, costs the encoder ~6 %
    if (!enabled) return;
    std::lock_guard<std::mutex> lock(mu);
    if (seen.insert(kernel).second) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
#define NOBS_COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

// ------------------------------------------------------------------------------------------
// K1: log-mel.  One warp per frame: windowed samples -> 16 x DFT-25 -> 4 radix-2 stages in
// shared memory -> power -> sparse mel filterbank (double accumulate) -> log10.
// HBM traffic: each PCM sample is fetched once from DRAM (neighbouring frames re-hit L1/L2),
// each output value written once.
// ------------------------------------------------------------------------------------------
constexpr int kMelWarps = 4;

__global__ void __launch_bounds__(kMelWarps * 32) mel_stft_kernel(MelTables t, const MelJob* __restrict__ jobs, int frames_per_block) {
    // Twiddles are kept as compact per-stage tables: indexing the 400-entry table with strides of 16 / 8 / 4 / 2
    // (as the factorisation does) lands every lane of a warp on the same one or two banks.  Same values, no conflicts.
    __shared__ float2 s_tw25[25];        // e^{-2 pi i j / 25}
    __shared__ float2 s_tws[375];        // radix-2 stage s (L = 25 << s): entries [25 * ((1 << s) - 1) + k] = e^{-2 pi i k / (2L)}
    __shared__ float s_hann[kNFft];
    __shared__ float s_x[kMelWarps][kNFft];
    __shared__ float2 s_a[kMelWarps][kNFft];
    __shared__ float2 s_b[kMelWarps][kNFft];

    const MelJob job = jobs[blockIdx.y];
    const int f_begin = blockIdx.x * frames_per_block;
    if (f_begin >= job.n_frames) return;
    for (int i = threadIdx.x; i < kNFft; i += blockDim.x) s_hann[i] = t.hann[i];
    for (int i = threadIdx.x; i < 25; i += blockDim.x) s_tw25[i] = t.tw[16 * i];
    for (int i = threadIdx.x; i < 375; i += blockDim.x) {
        const int stage = i < 25 ? 0 : i < 75 ? 1 : i < 175 ? 2 : 3;
        const int k = i - 25 * ((1 << stage) - 1);
        s_tws[i] = t.tw[k * (kNFft / (2 * (25 << stage)))];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* x = s_x[warp];
    float2* A = s_a[warp];
    float2* B = s_b[warp];
    const int f_end = min(f_begin + frames_per_block, job.n_frames);
    float wmax = -INFINITY;

    for (int f = f_begin + warp; f < f_end; f += kMelWarps) {
        const int base = f * kHop - kNFft / 2;  // sample index of window element 0
        if (base >= 0 && base + kNFft <= job.n_samples) {
            const float4* p = reinterpret_cast<const float4*>(job.pcm + base);  // base % 4 == 0, pcm 16B aligned
            for (int j = lane; j < kNFft / 4; j += 32) {
                const float4 v = __ldg(p + j);
                x[4 * j + 0] = v.x * s_hann[4 * j + 0];
                x[4 * j + 1] = v.y * s_hann[4 * j + 1];
                x[4 * j + 2] = v.z * s_hann[4 * j + 2];
                x[4 * j + 3] = v.w * s_hann[4 * j + 3];
            }
        } else {
            for (int j = lane; j < kNFft; j += 32) {
                const int s = base + j;
                float v = 0.0f;
                if (s < 0) { if (-s < job.n_samples) v = job.pcm[-s]; }   // reflect pad on the left edge
                else if (s < job.n_samples) v = job.pcm[s];                // zeros past the right edge
                x[j] = v * s_hann[j];
            }
        }
        __syncwarp();
        // 16 DFTs of length 25 over the residues n mod 16
        for (int o = lane; o < kNFft; o += 32) {
            const int r = o / 25, k = o - r * 25;
            float re = 0.0f, im = 0.0f;
#pragma unroll 5
            for (int m = 0; m < 25; ++m) {
                const float xv = x[16 * m + r];
                const float2 w = s_tw25[(m * k) % 25];
                re += xv * w.x;
                im -= xv * w.y;
            }
            A[o] = make_float2(re, im);
        }
        __syncwarp();
        // radix-2 combines: (R,L) = (16,25) -> (8,50) -> (4,100) -> (2,200) -> (1,400)
        float2* in = A;
        float2* out = B;
#pragma unroll
        for (int stage = 0; stage < 4; ++stage) {
            const int L = 25 << stage, halfR = 8 >> stage;
            const float2* tws = s_tws + 25 * ((1 << stage) - 1);
            for (int o = lane; o < kNFft / 2; o += 32) {
                const int r = o / L, k = o - r * L;
                const float2 e = in[r * L + k];
                const float2 od = in[(r + halfR) * L + k];
                const float2 w = tws[k];
                const float tr = w.x * od.x + w.y * od.y;
                const float ti = w.x * od.y - w.y * od.x;
                out[r * 2 * L + k] = make_float2(e.x + tr, e.y + ti);
                out[r * 2 * L + k + L] = make_float2(e.x - tr, e.y - ti);
            }
            __syncwarp();
            float2* tmp = in; in = out; out = tmp;
        }
        // power spectrum of bins 0..200 (x is free again)
        for (int k = lane; k < kNFreq; k += 32) {
            const float2 v = in[k];
            x[k] = v.x * v.x + v.y * v.y;
        }
        __syncwarp();
        for (int m = lane; m < t.n_mel; m += 32) {
            const int2 rg = t.ranges[m];
            const float* fl = t.filters + (size_t)m * kNFreq;
            double sum = 0.0;
            for (int k = rg.x; k < rg.y; ++k) sum += (double)__fmul_rn(x[k], __ldg(fl + k));
            const float v = (float)log10(fmax(sum, 1e-10));
            job.raw[(size_t)f * t.n_mel + m] = v;
            wmax = fmaxf(wmax, v);
        }
        __syncwarp();
    }
    wmax = warp_max(wmax);
    if (lane == 0 && wmax > -INFINITY) atomicMax(job.max_key, float_key(wmax));
}

void launch_mel_stft(const MelTables& t, const MelJob* jobs_dev, int n_jobs, int max_frames, cudaStream_t s) {
    if (n_jobs <= 0 || max_frames <= 0) return;
    const int frames_per_block = 32;
    dim3 grid((max_frames + frames_per_block - 1) / frames_per_block, n_jobs);
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&mel_stft_kernel));
    mel_stft_kernel<<<grid, kMelWarps * 32, 0, s>>>(t, jobs_dev, frames_per_block);
    NOBS_COUNT_LAUNCH();
}

// normalisation shared by pack/export: global max over all frames (frames past n_frames are
// exactly -10), clamp to max-8, (x+4)/4 — in double like the reference path.
__device__ __forceinline__ float mel_normalise(float raw, double mmax) {
    double v = raw;
    if (v < mmax) v = (double)(float)mmax;
    return (float)((v + 4.0) / 4.0);
}
__device__ __forceinline__ double mel_clamp_floor(const int* max_key, int n_frames, int n_len) {
    float gmax = key_float(*max_key);
    if (n_frames < n_len) gmax = fmaxf(gmax, -10.0f);
    return (double)gmax - 8.0;
}

template <typename T>
__global__ void pack_mel_kernel(const PackJob* __restrict__ jobs, int n_mel, T* __restrict__ dst) {
    const PackJob j = jobs[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kWinRowsIn * n_mel) return;
    const int t = idx / n_mel, c = idx - t * n_mel;
    float v = 0.0f;
    const int frame = j.seek + t;
    if (t < 3000 && frame < j.n_len) {
        const double mmax = mel_clamp_floor(j.max_key, j.n_frames, j.n_len);
        const float raw = frame < j.n_frames ? j.raw[(size_t)frame * n_mel + c] : -10.0f;
        v = mel_normalise(raw, mmax);
    }
    dst[((size_t)blockIdx.y * kWinRowsIn + 1 + t) * n_mel + c] = from_f32<T>(v);
}

template <typename T>
void launch_pack_mel(const PackJob* jobs_dev, int n_win, int n_mel, T* dst, cudaStream_t s) {
    if (n_win <= 0) return;
    dim3 grid((kWinRowsIn * n_mel + 255) / 256, n_win);
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&pack_mel_kernel<T>));
    pack_mel_kernel<T><<<grid, 256, 0, s>>>(jobs_dev, n_mel, dst);
    NOBS_COUNT_LAUNCH();
}
template void launch_pack_mel<float>(const PackJob*, int, int, float*, cudaStream_t);
template void launch_pack_mel<bf16>(const PackJob*, int, int, bf16*, cudaStream_t);

__global__ void export_mel_kernel(const float* __restrict__ raw, const int* __restrict__ max_key, int n_frames, int n_len, int n_mel,
                                  float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_len * n_mel) return;
    const int c = (int)(idx / n_len), i = (int)(idx - (size_t)c * n_len);
    const double mmax = mel_clamp_floor(max_key, n_frames, n_len);
    const float r = i < n_frames ? raw[(size_t)i * n_mel + c] : -10.0f;
    out[idx] = mel_normalise(r, mmax);
}
void launch_export_mel(const float* raw, const int* max_key, int n_frames, int n_len, int n_mel, float* out, cudaStream_t s) {
    const size_t n = (size_t)n_len * n_mel;
    export_mel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(raw, max_key, n_frames, n_len, n_mel, out);
    NOBS_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// fp32 GEMM (parity mode): C = epi(A * W^T), 128x128x16 tiles, 8x8 per thread.
// ------------------------------------------------------------------------------------------
constexpr int GB_M = 128, GB_N = 128, GB_K = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                       float* C, int ldc, int M, int N, int K, Epilogue e) {
    __shared__ __align__(16) float As[GB_K][GB_M + 4];
    __shared__ __align__(16) float Bs[GB_K][GB_N + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int idx = tid + p * 256;
            const int row = idx >> 2, kq = (idx & 3) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (m0 + row < M) va = *reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * lda + k0 + kq);
            if (n0 + row < N) vb = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + row) * ldw + k0 + kq));
            As[kq + 0][row] = va.x; As[kq + 1][row] = va.y; As[kq + 2][row] = va.z; As[kq + 3][row] = va.w;
            Bs[kq + 0][row] = vb.x; Bs[kq + 1][row] = vb.y; Bs[kq + 2][row] = vb.z; Bs[kq + 3][row] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GB_K; ++k) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n < N) {
                size_t idx = (size_t)m * ldc + n;
                if (e.head_rows > 0) {
                    const int w = m / e.head_rows, pos = m - w * e.head_rows;
                    idx = (size_t)w * e.head_rows * ldc + ((size_t)(n >> 6) * e.head_rows + pos) * 64 + (n & 63);
                }
                C[idx] = apply_epilogue(acc[i][j], m, n, e);
            }
        }
    }
}

void launch_gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K, const Epilogue& e,
                     cudaStream_t s) {
    if (M <= 0 || N <= 0) return;
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M);
    gemm_f32_kernel<<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, e);
    NOBS_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5), one warp per row, fp32 statistics
// ------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ idx,
                                                        const float* __restrict__ g, const float* __restrict__ b, TO* __restrict__ y,
                                                        int ldy, int rows, int d) {
    pdl_wait();
    pdl_launch_dependents();   // only one kernel ahead may sit resident: parked CTAs of a long dependency chain would hold SM resources another decode lane needs
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)(idx ? idx[row] : row) * ldx;
    float sum = 0.0f;
    for (int i = lane; i < d; i += 32) sum += xr[i];
    const float mean = warp_sum(sum) / d;
    float var = 0.0f;
    for (int i = lane; i < d; i += 32) { const float t = xr[i] - mean; var += t * t; }
    var = warp_sum(var) / d;
    const float inv = rsqrtf(var + 1e-5f);
    TO* yr = y + (size_t)row * ldy;
    for (int i = lane; i < d; i += 32) yr[i] = from_f32<TO>((xr[i] - mean) * inv * __ldg(g + i) + __ldg(b + i));
}

template <typename TO>
void launch_layernorm(const float* x, int ldx, const float* g, const float* b, TO* y, int ldy, int rows, int d, cudaStream_t s) {
    if (rows <= 0) return;
    launch_kernel(layernorm_kernel<TO>, dim3((rows + 7) / 8), dim3(256), 0, s, true, x, ldx, (const int*)nullptr, g, b, y, ldy, rows, d);
    NOBS_COUNT_LAUNCH();
}
template <typename TO>
void launch_layernorm_gather(const float* x, int ldx, const int* idx, const float* g, const float* b, TO* y, int ldy, int rows, int d,
                             cudaStream_t s) {
    if (rows <= 0) return;
    launch_kernel(layernorm_kernel<TO>, dim3((rows + 7) / 8), dim3(256), 0, s, true, x, ldx, idx, g, b, y, ldy, rows, d);
    NOBS_COUNT_LAUNCH();
}
template void launch_layernorm<float>(const float*, int, const float*, const float*, float*, int, int, int, cudaStream_t);
template void launch_layernorm<bf16>(const float*, int, const float*, const float*, bf16*, int, int, int, cudaStream_t);
template void launch_layernorm_gather<float>(const float*, int, const int*, const float*, const float*, float*, int, int, int, cudaStream_t);
template void launch_layernorm_gather<bf16>(const float*, int, const int*, const float*, const float*, bf16*, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------
// Encoder self-attention, fp32 parity mode: flash-style, 64 queries x 64 keys per step,
// 256 threads, each owning a 4x4 block of S and of O.
// ------------------------------------------------------------------------------------------
constexpr int EA_BQ = 64, EA_BK = 64, EA_DH = 64;

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(256) enc_attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int d, int n_valid) {
    extern __shared__ __align__(16) float ea_smem[];
    float (*Qs)[EA_DH + 4] = reinterpret_cast<float (*)[EA_DH + 4]>(ea_smem);                              // [64][68]
    float (*Kt)[EA_BK + 4] = reinterpret_cast<float (*)[EA_BK + 4]>(ea_smem + EA_BQ * (EA_DH + 4));          // [dh][64+4]
    float (*Vs)[EA_DH + 4] = reinterpret_cast<float (*)[EA_DH + 4]>(ea_smem + 2 * EA_BQ * (EA_DH + 4));      // [64][68]
    float (*Ps)[EA_BK + 4] = reinterpret_cast<float (*)[EA_BK + 4]>(ea_smem + 3 * EA_BQ * (EA_DH + 4));      // [64][68]

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int qt = blockIdx.x, h = blockIdx.y, w = blockIdx.z;
    const size_t ld = (size_t)3 * d;
    const T* base = qkv + (size_t)w * kWinRows * ld;
    const T* Qg = base + (size_t)qt * EA_BQ * ld + h * EA_DH;
    const T* Kg = base + d + h * EA_DH;
    const T* Vg = base + 2 * d + h * EA_DH;

    for (int i = tid; i < EA_BQ * EA_DH / 4; i += 256) {
        const int r = i >> 4, c = (i & 15) * 4;
        const float4 v = load4<T>(Qg + (size_t)r * ld + c);
        *reinterpret_cast<float4*>(&Qs[r][c]) = v;
    }
    float m_i[4], l_i[4], O[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_i[i] = -INFINITY;
        l_i[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) O[i][j] = 0.0f;
    }
    const int n_tiles = (n_valid + EA_BK - 1) / EA_BK;
    for (int kt = 0; kt < n_tiles; ++kt) {
        __syncthreads();  // previous tile fully consumed (also orders the Q store on the first pass)
        for (int i = tid; i < EA_BK * EA_DH / 4; i += 256) {
            const int r = i >> 4, c = (i & 15) * 4;
            const int key = kt * EA_BK + r;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (key < n_valid) {
                kv = load4<T>(Kg + (size_t)key * ld + c);
                vv = load4<T>(Vg + (size_t)key * ld + c);
            }
            Kt[c + 0][r] = kv.x; Kt[c + 1][r] = kv.y; Kt[c + 2][r] = kv.z; Kt[c + 3][r] = kv.w;
            *reinterpret_cast<float4*>(&Vs[r][c]) = vv;
        }
        __syncthreads();
        float S[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) S[i][j] = 0.0f;
#pragma unroll 8
        for (int e = 0; e < EA_DH; ++e) {
            const float4 kk = *reinterpret_cast<const float4*>(&Kt[e][tx * 4]);
            float q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = Qs[ty * 4 + i][e];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                S[i][0] = fmaf(q[i], kk.x, S[i][0]);
                S[i][1] = fmaf(q[i], kk.y, S[i][1]);
                S[i][2] = fmaf(q[i], kk.z, S[i][2]);
                S[i][3] = fmaf(q[i], kk.w, S[i][3]);
            }
        }
        // scale, mask padded keys, online softmax
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kt * EA_BK + tx * 4 + j;
                S[i][j] = key < n_valid ? S[i][j] * 0.125f : -INFINITY;
                rmax = fmaxf(rmax, S[i][j]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            const float m_new = fmaxf(m_i[i], rmax);
            const float corr = __expf(m_i[i] - m_new);  // 0 on the first tile (m_i = -inf)
            float rsum = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                S[i][j] = expf(S[i][j] - m_new);
                rsum += S[i][j];
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
            l_i[i] = l_i[i] * corr + rsum;
            m_i[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) O[i][j] *= corr;
            *reinterpret_cast<float4*>(&Ps[ty * 4 + i][tx * 4]) = make_float4(S[i][0], S[i][1], S[i][2], S[i][3]);
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < EA_BK; ++j) {
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p = Ps[ty * 4 + i][j];
                O[i][0] = fmaf(p, vv.x, O[i][0]);
                O[i][1] = fmaf(p, vv.y, O[i][1]);
                O[i][2] = fmaf(p, vv.z, O[i][2]);
                O[i][3] = fmaf(p, vv.w, O[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float inv = 1.0f / l_i[i];
        const size_t row = (size_t)w * kWinRows + qt * EA_BQ + ty * 4 + i;
        store4(out + row * d + h * EA_DH + tx * 4, make_float4(O[i][0] * inv, O[i][1] * inv, O[i][2] * inv, O[i][3] * inv));
    }
}

template <typename T>
void launch_enc_attention_simt(const T* qkv, T* out, int n_win, int n_head, int d, cudaStream_t s) {
    if (n_win <= 0) return;
    const int smem = 4 * EA_BQ * (EA_DH + 4) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(enc_attention_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    dim3 grid(kWinRows / EA_BQ, n_head, n_win);
    enc_attention_simt_kernel<T><<<grid, 256, smem, s>>>(qkv, out, d, 1500);
    NOBS_COUNT_LAUNCH();
}
template void launch_enc_attention_simt<float>(const float*, float*, int, int, int, cudaStream_t);
template void launch_enc_attention_simt<bf16>(const bf16*, bf16*, int, int, int, cudaStream_t);
void launch_enc_attention_f32(const float* qkv, float* out, int n_win, int n_head, int d, cudaStream_t s) {
    launch_enc_attention_simt<float>(qkv, out, n_win, n_head, d, s);
}

// ------------------------------------------------------------------------------------------
// Decoder row kernels
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_kernel(const RowDesc* __restrict__ rows, int n_rows, const T* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                             float* __restrict__ x, int d) {
    const int r = blockIdx.x;
    const RowDesc rd = rows[r];
    const T* te = tok_emb + (size_t)rd.token * d;
    const float* pe = pos_emb + (size_t)rd.pos * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[(size_t)r * d + i] = to_f32(te[i]) + pe[i];
}
template <typename T>
void launch_embed(const RowDesc* rows, int n_rows, const T* tok_emb, const float* pos_emb, float* x, int d, cudaStream_t s) {
    if (n_rows <= 0) return;
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&embed_kernel<T>));
    embed_kernel<T><<<n_rows, 128, 0, s>>>(rows, n_rows, tok_emb, pos_emb, x, d);
    NOBS_COUNT_LAUNCH();
}
template void launch_embed<float>(const RowDesc*, int, const float*, const float*, float*, int, cudaStream_t);
template void launch_embed<bf16>(const RowDesc*, int, const bf16*, const float*, float*, int, cudaStream_t);

template <typename T>
__global__ void scatter_kv_kernel(const RowDesc* __restrict__ rows, const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc,
                                  size_t slot_stride, int n_pos_cap, int d) {
    const int r = blockIdx.x;
    const RowDesc rd = rows[r];
    const T* src = qkv + (size_t)r * 3 * d;
    const size_t base = (size_t)rd.kv_slot * slot_stride + (size_t)rd.pos * 64;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const size_t dst = base + (size_t)(i >> 6) * n_pos_cap * 64 + (i & 63);
        kc[dst] = src[d + i];
        vc[dst] = src[2 * d + i];
    }
}
template <typename T>
void launch_scatter_kv(const RowDesc* rows, int n_rows, const T* qkv, T* kpanel0, T* vpanel0, size_t slot_stride, int n_pos_cap, int d,
                       cudaStream_t s) {
    if (n_rows <= 0) return;
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&scatter_kv_kernel<T>));
    scatter_kv_kernel<T><<<n_rows, 128, 0, s>>>(rows, qkv, kpanel0, vpanel0, slot_stride, n_pos_cap, d);
    NOBS_COUNT_LAUNCH();
}
template void launch_scatter_kv<float>(const RowDesc*, int, const float*, float*, float*, size_t, int, int, cudaStream_t);
template void launch_scatter_kv<bf16>(const RowDesc*, int, const bf16*, bf16*, bf16*, size_t, int, int, cudaStream_t);

// ---- decoder attention over a head-major KV panel: one (row, head) per block.
// A key / value row is 64 contiguous elements; LPK = 64*sizeof(T)/16 lanes cooperate on one row with a
// 16-byte load each, so a warp instruction reads 512 contiguous bytes and the whole panel is streamed
// exactly once for the scores and once for P*V (HBM-bound: the cross-attention panels are the
// dominant traffic of a decoder step).
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    typedef float4 Raw;
    static __device__ __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ Raw load(const float* p) {  // streaming: read once, do not keep in L1
        Raw r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[4]) { v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
};
template <> struct Vec16<bf16> {
    static constexpr int N = 8;
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ Raw load(const bf16* p) {
        Raw r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
};

constexpr int DA_THREADS = 256;
constexpr int DA_UNROLL = 8;

template <typename T>
__global__ void __launch_bounds__(DA_THREADS) dec_attention_kernel(const RowDesc* __restrict__ rows, const T* __restrict__ q, int ldq,
                                                                   const T* __restrict__ kc, const T* __restrict__ vc, T* __restrict__ out, int ldo,
                                                                   int cross, size_t slot_stride, size_t head_stride, int n_keys) {
    constexpr int VN = Vec16<T>::N;            // elements per 16-byte load
    constexpr int LPK = 64 / VN;               // lanes per key row (8 for bf16, 16 for fp32)
    constexpr int KPI = DA_THREADS / LPK;      // keys per block iteration (32 / 16)
    __shared__ __align__(16) float qs[64];
    __shared__ float sc[kWinRows];
    __shared__ float red[32];
    __shared__ float part[KPI][64 + 1];
    pdl_wait();
    pdl_launch_dependents();   // only one kernel ahead may sit resident: parked CTAs of a long dependency chain would hold SM resources another decode lane needs
    const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const RowDesc rd = rows[r];
    const int nk = cross ? n_keys : rd.pos + 1;
    const size_t base = (size_t)(cross ? rd.audio_slot : rd.kv_slot) * slot_stride + (size_t)h * head_stride;
    const T* K = kc + base;
    const T* V = vc + base;
    if (tid < 64) qs[tid] = to_f32(q[(size_t)r * ldq + h * 64 + tid]);
    __syncthreads();
    const int sub = tid % LPK, ks = tid / LPK;
    float qv[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) qv[i] = qs[sub * VN + i];

    // scores: DA_UNROLL key rows per thread in flight (16-byte loads issued back to back, converted at use)
    float lmax = -INFINITY;
    for (int j0 = 0; j0 < nk; j0 += DA_UNROLL * KPI) {
        typename Vec16<T>::Raw raw[DA_UNROLL];
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int j = j0 + u * KPI + ks;
            raw[u] = j < nk ? Vec16<T>::load(K + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int j = j0 + u * KPI + ks;
            float kvv[VN];
            Vec16<T>::unpack(raw[u], kvv);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < VN; ++i) acc = fmaf(qv[i], kvv[i], acc);
#pragma unroll
            for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (j < nk) {
                acc *= 0.125f;
                if (sub == 0) sc[j] = acc;
                lmax = fmaxf(lmax, acc);
            }
        }
    }
    const float mx = block_reduce(lmax, -INFINITY, OpMax(), red);  // its barriers also publish sc[]
    float lsum = 0.0f;
    for (int j = tid; j < nk; j += DA_THREADS) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        lsum += p;
    }
    const float total = block_reduce(lsum, 0.0f, OpAddF(), red);
    const float inv = 1.0f / total;

    // P * V with the same (key slot, 16-byte lane) mapping
    float acc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = 0.0f;
    for (int j0 = 0; j0 < nk; j0 += DA_UNROLL * KPI) {
        typename Vec16<T>::Raw raw[DA_UNROLL];
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int j = j0 + u * KPI + ks;
            raw[u] = j < nk ? Vec16<T>::load(V + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < DA_UNROLL; ++u) {
            const int j = j0 + u * KPI + ks;
            const float p = j < nk ? sc[j] : 0.0f;
            float vvv[VN];
            Vec16<T>::unpack(raw[u], vvv);
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[i] = fmaf(p, vvv[i], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) part[ks][sub * VN + i] = acc[i];
    __syncthreads();
    if (tid < 64) {
        float o = 0.0f;
#pragma unroll 8
        for (int k = 0; k < KPI; ++k) o += part[k][tid];
        out[(size_t)r * ldo + h * 64 + tid] = from_f32<T>(o * inv);
    }
}
// Causal self-attention over the (short) self-KV panel: one 4-warp block per (row, head), the keys interleaved
// over the warps (16-byte lanes as above), so the longest sequence of a step batch costs a quarter of a
// warp-per-head pass: this kernel sits on the latency chain of every decoder layer.
constexpr int SA_WARPS = 2;   // 128 rows x 20 heads x 64 threads at <= 64 registers is ONE resident wave; 4 warps were two waves plus a tail
// QkvPartials (kernels.cuh): when given, q and the new key / value row of this (row, head) are still the split-K
// partial sums of the QKV projection; the block finishes them (fixed split order, + bias), appends k / v to the cache
// panel and attends over keys 0..pos with the new row taken from shared memory — the projection's epilogue launch is
// gone from the latency chain.  Only valid when no two rows of the batch share a KV slot (single-token steps).
// PRE (single-token steps of a lane with <= 64 rows: half as many blocks per SM, twice the registers): the first trip of keys AND of
// values — the cache rows of earlier steps, which do not depend on the predecessor kernel — is requested BEFORE the programmatic
// dependency wait, so that after the wait only the partial-sum round trip stands between the block and its dot products.  This
// kernel sits on the latency chain of every decoder layer (profiles/README.md: 23 us per layer before, of which ~10 us were this
// block's eight dependent memory round trips).
template <typename T, int MINB, bool PRE>
__global__ void __launch_bounds__(SA_WARPS * 32, MINB) dec_self_attention_kernel(const RowDesc* __restrict__ rows, const T* __restrict__ q, int ldq,
                                                                          T* __restrict__ kc, T* __restrict__ vc, T* __restrict__ out,
                                                                          int ldo, int n_head, size_t slot_stride, size_t head_stride, QkvPartials qp) {
    constexpr int VN = Vec16<T>::N, LPK = 64 / VN, KPW = 32 / LPK, KPB = KPW * SA_WARPS, UNR = 8;
    __shared__ float sc[448];
    __shared__ float red[2 * SA_WARPS];
    __shared__ float part[SA_WARPS][64];
    __shared__ float qs[64];
    __shared__ __align__(16) T k_new[64];
    __shared__ __align__(16) T v_new[64];
    const long long tr = trace_begin(3, out);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int r = blockIdx.x, h = blockIdx.y;
    const RowDesc rd = rows[r];   // uploaded before the round's first kernel: safe to read ahead of the dependency wait
    const int nk = min(rd.pos + 1, 448);
    const size_t base = (size_t)rd.kv_slot * slot_stride + (size_t)h * head_stride;
    T* K = kc + base;
    T* V = vc + base;
    const bool fused = qp.partial != nullptr;
    const int j_new = fused ? nk - 1 : -1;    // the key row that only exists in shared memory so far
    const int sub = lane % LPK, ks = warp * KPW + lane / LPK;   // key slot of this thread within a block iteration
    constexpr int UNRV = 2;   // values: a quarter of a trip ahead (registers: 80 per thread keep 8 blocks of this kernel beside another lane's attention CTAs)
    typename Vec16<T>::Raw pre_k[PRE ? UNR : 1], pre_v[PRE ? UNRV : 1];
    if (PRE) {   // only used with fused == true: rows 0 .. nk-2 were written by earlier rounds
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = u * KPB + ks;
            pre_k[u] = j < nk - 1 ? Vec16<T>::load(K + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < UNRV; ++u) {
            const int j = u * KPB + ks;
            pre_v[u] = j < nk - 1 ? Vec16<T>::load(V + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
    }
    pdl_wait();
    pdl_launch_dependents();
    trace_end(trace_begin(103, out));
    if (fused) {
        const int d = n_head * 64;
        for (int e = tid; e < 64; e += SA_WARPS * 32) {   // element e of q, k and v: all split-K loads of a trip in flight together
            const float* src = qp.partial + (size_t)r * qp.ld + h * 64 + e;
            float aq = 0.0f, ak = 0.0f, av = 0.0f;
            for (int s0 = 0; s0 < qp.splits; s0 += 4) {
                float tq[4], tk[4], tv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = s0 + u < qp.splits;
                    const float* p = src + (size_t)(s0 + u) * qp.plane;
                    tq[u] = ok ? __ldcg(p) : 0.0f;
                    tk[u] = ok ? __ldcg(p + d) : 0.0f;
                    tv[u] = ok ? __ldcg(p + 2 * d) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { aq += tq[u]; ak += tk[u]; av += tv[u]; }   // fixed split order
            }
            if (qp.bias) { aq += __ldg(qp.bias + h * 64 + e); ak += __ldg(qp.bias + d + h * 64 + e); av += __ldg(qp.bias + 2 * d + h * 64 + e); }
            // q goes through the storage type exactly as on the path that materialises the projection (rounds that contain a
            // prefill take that path for every row): a row's result must not depend on what else is in its batch
            qs[e] = to_f32(from_f32<T>(aq));
            const T kb = from_f32<T>(ak), vb = from_f32<T>(av);
            k_new[e] = kb;
            v_new[e] = vb;
            K[(size_t)(nk - 1) * 64 + e] = kb;
            V[(size_t)(nk - 1) * 64 + e] = vb;
        }
    } else {
        for (int e = tid; e < 64; e += SA_WARPS * 32) qs[e] = to_f32(q[(size_t)r * ldq + h * 64 + e]);
    }
    __syncthreads();
    float qv[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) qv[i] = qs[sub * VN + i];
    float lmax = -INFINITY;
    for (int j0 = 0; j0 < nk; j0 += UNR * KPB) {
        typename Vec16<T>::Raw raw[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = j0 + u * KPB + ks;
            if (j == j_new) raw[u] = *reinterpret_cast<const typename Vec16<T>::Raw*>(k_new + sub * VN);
            else if (PRE && j0 == 0) raw[u] = pre_k[u];
            else raw[u] = j < nk ? Vec16<T>::load(K + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = j0 + u * KPB + ks;
            float kvv[VN];
            Vec16<T>::unpack(raw[u], kvv);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < VN; ++i) acc = fmaf(qv[i], kvv[i], acc);
#pragma unroll
            for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (j < nk) {
                acc *= 0.125f;
                if (sub == 0) sc[j] = acc;
                lmax = fmaxf(lmax, acc);
            }
        }
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float mx = red[0];
#pragma unroll
    for (int w = 1; w < SA_WARPS; ++w) mx = fmaxf(mx, red[w]);
    float lsum = 0.0f;
    for (int j = tid; j < nk; j += SA_WARPS * 32) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        lsum += p;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[SA_WARPS + warp] = lsum;
    __syncthreads();
    float total = 0.0f;
#pragma unroll
    for (int w = 0; w < SA_WARPS; ++w) total += red[SA_WARPS + w];
    const float inv = 1.0f / total;
    float acc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = 0.0f;
    for (int j0 = 0; j0 < nk; j0 += UNR * KPB) {
        typename Vec16<T>::Raw raw[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = j0 + u * KPB + ks;
            if (j == j_new) raw[u] = *reinterpret_cast<const typename Vec16<T>::Raw*>(v_new + sub * VN);
            else if (PRE && j0 == 0 && u < UNRV) raw[u] = pre_v[u < UNRV ? u : 0];
            else raw[u] = j < nk ? Vec16<T>::load(V + (size_t)j * 64 + sub * VN) : Vec16<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = j0 + u * KPB + ks;
            const float p = j < nk ? sc[j] : 0.0f;
            float vvv[VN];
            Vec16<T>::unpack(raw[u], vvv);
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[i] = fmaf(p, vvv[i], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
#pragma unroll
        for (int o = 16; o >= LPK; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane < LPK) {
#pragma unroll
        for (int i = 0; i < VN; ++i) part[warp][sub * VN + i] = acc[i];
    }
    __syncthreads();
    for (int e = tid; e < 64; e += SA_WARPS * 32) {
        float o = 0.0f;
#pragma unroll
        for (int w = 0; w < SA_WARPS; ++w) o += part[w][e];   // fixed order
        out[(size_t)r * ldo + h * 64 + e] = from_f32<T>(o * inv);
    }
    trace_end(tr);
}

template <typename T>
void launch_dec_attention(const RowDesc* rows, int n_rows, const T* q, int ldq, const T* kbase, const T* vbase, T* out, int ldo, int n_head,
                          int cross, size_t slot_stride, size_t head_stride, int n_keys, cudaStream_t s, const QkvPartials* qkv_partials) {
    if (n_rows <= 0) return;
    if (!cross) {
        dim3 grid(n_rows, n_head);
        const QkvPartials qp = qkv_partials ? *qkv_partials : QkvPartials{};
        static const bool pre_ok = [] { const char* v = getenv("NOBS_WHISPER_SA_PREFETCH"); return !(v && *v == '0'); }();
        if (pre_ok && qp.partial && sizeof(T) == 2 && n_rows * n_head <= 8 * 148)   // one wave at 8 blocks per SM
            launch_kernel(dec_self_attention_kernel<T, 12, true>, grid, dim3(SA_WARPS * 32), 0, s, true, rows, q, ldq, const_cast<T*>(kbase), const_cast<T*>(vbase),
                          out, ldo, n_head, slot_stride, head_stride, qp);
        else
            launch_kernel(dec_self_attention_kernel<T, 18, false>, grid, dim3(SA_WARPS * 32), 0, s, true, rows, q, ldq, const_cast<T*>(kbase), const_cast<T*>(vbase),
                          out, ldo, n_head, slot_stride, head_stride, qp);
        NOBS_COUNT_LAUNCH();
        return;
    }
    dim3 grid(n_rows, n_head);
    launch_kernel(dec_attention_kernel<T>, grid, dim3(DA_THREADS), 0, s, true, rows, q, ldq, kbase, vbase, out, ldo, cross, slot_stride, head_stride, n_keys);
    NOBS_COUNT_LAUNCH();
}
template void launch_dec_attention<float>(const RowDesc*, int, const float*, int, const float*, const float*, float*, int, int, int, size_t, size_t,
                                          int, cudaStream_t, const QkvPartials*);
template void launch_dec_attention<bf16>(const RowDesc*, int, const bf16*, int, const bf16*, const bf16*, bf16*, int, int, int, size_t, size_t, int,
                                         cudaStream_t, const QkvPartials*);

// beam search: copy the first n_pos rows of every [n_pos_cap][64] panel of one self-KV slot to another
template <typename T>
__global__ void kv_copy_kernel(const KvCopy* __restrict__ pairs, T* __restrict__ pool, size_t slot_stride, size_t panel_stride) {
    const KvCopy p = pairs[blockIdx.x];
    const size_t n = (size_t)p.n_pos * 64;
    const T* src = pool + (size_t)p.src * slot_stride + (size_t)blockIdx.y * panel_stride;
    T* dst = pool + (size_t)p.dst * slot_stride + (size_t)blockIdx.y * panel_stride;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
template <typename T>
void launch_kv_copy(const KvCopy* pairs, int n_pairs, T* pool, size_t slot_stride, int n_panels, size_t panel_stride, cudaStream_t s) {
    if (n_pairs <= 0) return;
    dim3 grid(n_pairs, n_panels);
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&kv_copy_kernel<T>));
    kv_copy_kernel<T><<<grid, 128, 0, s>>>(pairs, pool, slot_stride, panel_stride);
    NOBS_COUNT_LAUNCH();
}
template void launch_kv_copy<float>(const KvCopy*, int, float*, size_t, int, size_t, cudaStream_t);
template void launch_kv_copy<bf16>(const KvCopy*, int, bf16*, size_t, int, size_t, cudaStream_t);

// ------------------------------------------------------------------------------------------
// skinny-GEMM epilogue: one token row per block
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, const float (&v)[4]) {   // p is 8-byte aligned
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

constexpr int SR_THREADS = 256;
constexpr int SR_MAXG = 5;   // float4 column groups per thread: N <= 5120
constexpr int SR_PLANES = 8; // split-K planes summed per trip (independent 16-byte loads in flight)

// Register budget matters here: this kernel has to share SMs with the persistent cross-attention CTAs of
// another decode lane (2 x 288 threads x 64 registers), so it stays at <= 64 registers per thread.
template <typename T>
__global__ void __launch_bounds__(SR_THREADS, 4) skinny_reduce_kernel(SkinnyEpilogue e) {
    __shared__ float red[32];
    const long long tr = trace_begin(2, e.partial);
    pdl_wait();
    pdl_launch_dependents();   // only one kernel ahead may sit resident: parked CTAs of a long dependency chain would hold SM resources another decode lane needs
    trace_end(trace_begin(102, e.partial));
    const int r = blockIdx.x, tid = threadIdx.x;
    const size_t plane = (size_t)e.R * e.N;
    const float* p = e.partial + (size_t)r * e.N;
    const int n4 = e.N >> 2;                      // N is a multiple of 4 (model widths are multiples of 64)
    float4 vals[SR_MAXG];
    RowDesc rd{};
    if (e.rows) rd = e.rows[r];
    float lsum = 0.0f;
#pragma unroll
    for (int g = 0; g < SR_MAXG; ++g) {
        const int c4 = tid + g * SR_THREADS;
        vals[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 >= n4) continue;
        const int n = c4 * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < e.splits; s0 += SR_PLANES) {   // fixed order s0, s0+1, ...: deterministic
            float4 t[SR_PLANES];
#pragma unroll
            for (int u = 0; u < SR_PLANES; ++u)
                t[u] = (s0 + u < e.splits) ? __ldcg(reinterpret_cast<const float4*>(p + (size_t)(s0 + u) * plane + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < SR_PLANES; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        if (e.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (e.act == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_tanh_fast(v[i]);
        }
        if (e.x) {
            float4* xp = reinterpret_cast<float4*>(e.x + (size_t)r * e.N + n);
            const float4 xv = *xp;
            v[0] += xv.x; v[1] += xv.y; v[2] += xv.z; v[3] += xv.w;
            *xp = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (e.out) {
            store4<T>(static_cast<T*>(e.out) + (size_t)r * e.out_ld + n, v);
        }
        if (e.rows && n >= e.d) {   // 4 consecutive columns never straddle a 64-wide head block
            const int c = n - e.d, which = c >= e.d, i2 = which ? c - e.d : c;
            T* dst = static_cast<T*>(which ? e.vpanel : e.kpanel) + (size_t)rd.kv_slot * e.slot_stride + ((size_t)(i2 >> 6) * e.n_pos_cap + rd.pos) * 64 + (i2 & 63);
            store4<T>(dst, v);
        }
        lsum += (v[0] + v[1]) + (v[2] + v[3]);
        vals[g] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (e.ln_g) {
        const float mean = block_reduce(lsum, 0.0f, OpAddF(), red) / e.N;
        float lvar = 0.0f;
#pragma unroll
        for (int g = 0; g < SR_MAXG; ++g) {
            if (tid + g * SR_THREADS < n4) {
                const float a = vals[g].x - mean, b = vals[g].y - mean, c = vals[g].z - mean, dd = vals[g].w - mean;
                lvar += (a * a + b * b) + (c * c + dd * dd);
            }
        }
        const float var = block_reduce(lvar, 0.0f, OpAddF(), red) / e.N;
        const float inv = rsqrtf(var + 1e-5f);
        T* y = static_cast<T*>(e.y) + (size_t)r * e.N;
#pragma unroll
        for (int g = 0; g < SR_MAXG; ++g) {
            const int c4 = tid + g * SR_THREADS;
            if (c4 < n4) {
                const int n = c4 * 4;
                const float4 gg = __ldg(reinterpret_cast<const float4*>(e.ln_g + n)), bb = __ldg(reinterpret_cast<const float4*>(e.ln_b + n));
                const float o[4] = {(vals[g].x - mean) * inv * gg.x + bb.x, (vals[g].y - mean) * inv * gg.y + bb.y,
                                    (vals[g].z - mean) * inv * gg.z + bb.z, (vals[g].w - mean) * inv * gg.w + bb.w};
                store4<T>(y + n, o);
            }
        }
    }
    trace_end(tr);
}
template <typename T>
void launch_skinny_reduce(const SkinnyEpilogue& e, cudaStream_t s) {
    if (e.R <= 0) return;
    launch_kernel(skinny_reduce_kernel<T>, dim3(e.R), dim3(SR_THREADS), 0, s, true, e);
    NOBS_COUNT_LAUNCH();
}
template void launch_skinny_reduce<bf16>(const SkinnyEpilogue&, cudaStream_t);
template void launch_skinny_reduce<float>(const SkinnyEpilogue&, cudaStream_t);

// ------------------------------------------------------------------------------------------
// K6: logit filter + log-softmax + timestamp rule + argmax / sample / top-k, one block per row.
// Restates the reference path's per-step logit processing.
// ------------------------------------------------------------------------------------------
// The suppression rules of one sampling step, folded into ranges once per row (text ids are 97 % of the vocabulary
// and only need two comparisons).  Special tokens all sit at or above eot.
struct SuppressCtx {
    bool first_step_blank;   // suppress_blank on the first step: eot and " "
    bool text_forbidden;     // last token opened a timestamp pair: only a timestamp may follow
    bool ts_forbidden;       // no_timestamps, or the last two tokens were timestamps
    int ts_lo, ts_hi;        // timestamps outside [ts_lo, ts_hi) are suppressed
};
__device__ __forceinline__ SuppressCtx make_suppress_ctx(const SampleParams& p, const VocabIds& v) {
    SuppressCtx c;
    c.first_step_blank = p.suppress_blank && p.is_initial;
    c.text_forbidden = p.last_was_ts && !p.penult_was_ts;
    c.ts_forbidden = p.no_timestamps || (p.last_was_ts && p.penult_was_ts);
    c.ts_lo = p.has_ts ? max(v.beg, p.ts_min) : v.beg;
    c.ts_hi = p.is_initial ? min(v.n_vocab, max(p.ts_initial_limit, v.beg)) : v.n_vocab;
    return c;
}
__device__ __forceinline__ bool token_suppressed(int i, const SuppressCtx& c, const VocabIds& v) {
    if (i < v.eot) return c.text_forbidden || (c.first_step_blank && i == v.blank);
    if (i >= v.beg) return c.ts_forbidden || i < c.ts_lo || i >= c.ts_hi;
    if (c.first_step_blank && i == v.eot) return true;
    if (i == v.not_ || i == v.sot || i == v.nosp || i == v.solm || i == v.translate || i == v.transcribe || i == v.prev) return true;
    return i > v.sot && i <= v.sot + v.n_lang;
}

constexpr int PL_THREADS = 1024;

__global__ void __launch_bounds__(PL_THREADS) process_logits_kernel(const float* __restrict__ logits, int ld,
                                                                    const SampleParams* __restrict__ params, SampleResult* __restrict__ results,
                                                                    VocabIds v, float* __restrict__ logprobs_out, float* __restrict__ probs_out) {
    __shared__ float red_f[32];
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ double scan_d[32];
    __shared__ int s_pick;
    __shared__ int s_chosen[kMaxTopK];

    const long long tr = trace_begin(7, logits);
    pdl_wait();
    pdl_launch_dependents();
    trace_end(trace_begin(107, logits));   // only one kernel ahead may sit resident: parked CTAs of a long dependency chain would hold SM resources another decode lane needs
    const int row = blockIdx.x, tid = threadIdx.x;
    const SampleParams p = params[row];
    const SuppressCtx sc = make_suppress_ctx(p, v);
    const float* lg = logits + (size_t)row * ld;
    const int n = v.n_vocab;
    const bool scaled = p.temperature > 0.0f;
    // probabilities of pass D, re-read by the sampling passes instead of recomputing exp(logprob) three times
    const float* pr_cache = probs_out ? probs_out + (size_t)row * n : nullptr;

    // pass A: maxima
    float lmax = -INFINITY, rawmax = -INFINITY;
    for (int i = tid; i < n; i += PL_THREADS) {
        const float raw = lg[i];
        rawmax = fmaxf(rawmax, raw);
        if (!token_suppressed(i, sc, v)) lmax = fmaxf(lmax, scaled ? __fdiv_rn(raw, p.temperature) : raw);
    }
    lmax = block_reduce(lmax, -INFINITY, OpMax(), red_f);
    rawmax = block_reduce(rawmax, -INFINITY, OpMax(), red_f);
    // pass B: sums
    float lsum = 0.0f, rawsum = 0.0f;
    for (int i = tid; i < n; i += PL_THREADS) {
        const float raw = lg[i];
        if (p.want_nosp) rawsum += expf(raw - rawmax);
        if (!token_suppressed(i, sc, v)) lsum += expf((scaled ? __fdiv_rn(raw, p.temperature) : raw) - lmax);
    }
    lsum = block_reduce(lsum, 0.0f, OpAddF(), red_f);
    rawsum = block_reduce(rawsum, 0.0f, OpAddF(), red_f);
    const float logsumexp = logf(lsum) + lmax;
    auto logprob_of = [&](int i) -> float {
        if (token_suppressed(i, sc, v)) return -INFINITY;
        const float l = scaled ? __fdiv_rn(lg[i], p.temperature) : lg[i];
        return l - logsumexp;
    };
    // pass C: timestamp logsumexp vs best text token
    float ts_max = -INFINITY, text_max = -INFINITY;
    for (int i = tid; i < n; i += PL_THREADS) {
        const float lp = logprob_of(i);
        if (i >= v.beg) ts_max = fmaxf(ts_max, lp); else text_max = fmaxf(text_max, lp);
    }
    ts_max = block_reduce(ts_max, -INFINITY, OpMax(), red_f);
    text_max = block_reduce(text_max, -INFINITY, OpMax(), red_f);
    float ts_sum = 0.0f;
    for (int i = v.beg + tid; i < n; i += PL_THREADS) {
        const float lp = logprob_of(i);
        if (lp > -INFINITY) ts_sum += expf(lp - ts_max);
    }
    ts_sum = block_reduce(ts_sum, 0.0f, OpAddF(), red_f);
    const float ts_logprob = ts_sum > 0.0f ? logf(ts_sum) + ts_max : -INFINITY;
    const bool text_off = ts_logprob > text_max;

    auto final_logprob = [&](int i) -> float {
        if (text_off && i < v.beg) return -INFINITY;
        return logprob_of(i);
    };

    // pass D: probabilities, argmax (first maximum wins), timestamp statistics
    float best_p = 0.0f; int best_i = INT_MAX;
    float tsb_p = 0.0f; int tsb_i = INT_MAX;
    double sum_ts = 0.0, sum_all = 0.0;
    for (int i = tid; i < n; i += PL_THREADS) {
        const float lp = final_logprob(i);
        const float pr = lp > -INFINITY ? expf(lp) : 0.0f;
        if (logprobs_out) logprobs_out[(size_t)row * n + i] = lp;
        if (probs_out) probs_out[(size_t)row * n + i] = pr;
        if (pr > best_p) { best_p = pr; best_i = i; }          // strided ascending: keeps this thread's first max
        if (i >= v.beg) {
            sum_ts += pr;
            if (pr > tsb_p) { tsb_p = pr; tsb_i = i; }
        }
        sum_all += pr;
    }
    const float gbest_p = block_reduce(best_p, 0.0f, OpMax(), red_f);
    int cand = (best_p == gbest_p && best_i != INT_MAX) ? best_i : INT_MAX;
    const int gbest_i = block_reduce(cand, INT_MAX, OpMinI(), red_i);
    const float gts_p = block_reduce(tsb_p, 0.0f, OpMax(), red_f);
    cand = (tsb_p == gts_p && tsb_i != INT_MAX) ? tsb_i : INT_MAX;
    const int gts_i = block_reduce(cand, INT_MAX, OpMinI(), red_i);
    sum_ts = block_reduce(sum_ts, 0.0, OpAddD(), red_d);

    int pick = gbest_i == INT_MAX ? 0 : gbest_i;
    if (p.mode == 1) {
        // inverse-CDF sampling identical in structure to std::discrete_distribution:
        // q_i = p_i / sum, cp = inclusive prefix sums of q, pick = first i with cp[i] >= u.
        const int chunk = (n + PL_THREADS - 1) / PL_THREADS;
        const int i0 = tid * chunk, i1 = min(n, i0 + chunk);
        auto prob_of = [&](int i) -> double {
            if (pr_cache) return (double)__ldcg(pr_cache + i);   // written in pass D (other threads; ordered by the block reductions' barriers)
            const float lp = final_logprob(i);
            return lp > -INFINITY ? (double)expf(lp) : 0.0;
        };
        double tot = 0.0;
        for (int i = i0; i < i1; ++i) tot += prob_of(i);
        const double total = block_reduce(tot, 0.0, OpAddD(), red_d);
        double local = 0.0;
        for (int i = i0; i < i1; ++i) local += prob_of(i) / total;
        // exclusive scan of `local` over threads
        const int lane = tid & 31, warp = tid >> 5;
        double incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) scan_d[warp] = incl;
        if (tid == 0) s_pick = n - 1;  // cp.back() is forced to 1.0 by the reference distribution
        __syncthreads();
        double woff = 0.0;
        for (int w = 0; w < warp; ++w) woff += scan_d[w];
        double cp = woff + incl - local;
        int found = INT_MAX;
        if (!(cp >= p.u)) {   // otherwise the first index with cp >= u lies in an earlier thread's range
            for (int i = i0; i < i1; ++i) {
                cp += prob_of(i) / total;
                if (cp >= p.u) { found = i; break; }
            }
        }
        if (found != INT_MAX) atomicMin(&s_pick, found);
        __syncthreads();
        pick = s_pick;
    }

    int n_topk = 0;
    if (p.mode == 2) {
        // k rounds of block arg-max over log-probabilities, ties -> lower id
        const int k = min(p.k, kMaxTopK);
        for (int round = 0; round < k; ++round) {
            float bl = -INFINITY; int bi = INT_MAX;
            for (int i = tid; i < n; i += PL_THREADS) {
                bool taken = false;
                for (int c = 0; c < round; ++c) taken |= (s_chosen[c] == i);
                if (taken) continue;
                const float lp = final_logprob(i);
                if (lp > bl) { bl = lp; bi = i; }
            }
            const float gl = block_reduce(bl, -INFINITY, OpMax(), red_f);
            const int c2 = (bl == gl && bi != INT_MAX && gl > -INFINITY) ? bi : INT_MAX;
            const int gi = block_reduce(c2, INT_MAX, OpMinI(), red_i);
            if (gi == INT_MAX) break;  // uniform across the block
            if (tid == 0) s_chosen[round] = gi;
            __syncthreads();
            n_topk = round + 1;
        }
    }

    if (tid == 0) {
        SampleResult r;
        r.id = pick;
        const float lp = final_logprob(pick);
        r.plog = lp;
        r.p = lp > -INFINITY ? expf(lp) : 0.0f;
        r.tid = gts_i == INT_MAX ? -1 : gts_i;  // -1: no timestamp has probability > 0 (host picks the default)
        r.pt = (float)((double)gts_p / (sum_ts + 1e-10));
        r.ptsum = (float)sum_ts;
        if (pick >= v.beg) { r.tid = pick; r.pt = r.p; }
        r.no_speech_prob = p.want_nosp ? expf(lg[v.nosp] - (logf(rawsum) + rawmax)) : 0.0f;
        r.n_topk = n_topk;
        for (int c = 0; c < kMaxTopK; ++c) {
            if (c < n_topk) {
                const int id = s_chosen[c];
                const float l2 = final_logprob(id);
                r.topk_id[c] = id; r.topk_plog[c] = l2; r.topk_p[c] = expf(l2);
            } else { r.topk_id[c] = -1; r.topk_plog[c] = -INFINITY; r.topk_p[c] = 0.0f; }
        }
        results[row] = r;
    }
    trace_end(tr);
}

// ------------------------------------------------------------------------------------------
// K6 over a thread-block cluster: the same passes as process_logits_kernel, with one row spread over PLC_CL CTAs (contiguous
// vocabulary ranges) and every block-wide reduction finished through distributed shared memory — each CTA publishes its partial
// into a slot of every peer's shared memory, one cluster barrier (~0.2 us) per exchange, partials combined in rank order
// (deterministic).  One row per 1024-thread block made K6 150-220 us of every decoder round (one SM streaming 207 KB five times);
// eight CTAs per row bring the whole vocabulary pass onto 8 x rows SMs.
constexpr int PLC_CL = 8;
constexpr int PLC_THREADS = 512;

__global__ void __cluster_dims__(PLC_CL, 1, 1) __launch_bounds__(PLC_THREADS)
process_logits_cluster_kernel(const float* __restrict__ logits, int ld, const SampleParams* __restrict__ params, SampleResult* __restrict__ results, VocabIds v,
                              float* __restrict__ logprobs_out, float* __restrict__ probs_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red_f[32];
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ double scan_d[32];
    __shared__ int s_pick;
    __shared__ int s_chosen[kMaxTopK];
    // exchange slots, written by the peers: [exchange][source rank][value]
    __shared__ float xf[5 + kMaxTopK][PLC_CL][2];
    __shared__ double xd[3][PLC_CL];
    __shared__ int xi[2 + kMaxTopK][PLC_CL][2];

    const long long tr = trace_begin(7, logits);
    pdl_wait();
    pdl_launch_dependents();
    trace_end(trace_begin(107, logits));
    const int rank = (int)cluster.block_rank();
    const int row = blockIdx.y, tid = threadIdx.x;
    const SampleParams p = params[row];
    const SuppressCtx sc = make_suppress_ctx(p, v);
    const float* lg = logits + (size_t)row * ld;
    const int n = v.n_vocab;
    const int span = (n + PLC_CL - 1) / PLC_CL;
    const int lo = rank * span, hi = min(n, lo + span);     // this CTA's vocabulary range
    const bool scaled = p.temperature > 0.0f;
    const float* pr_cache = probs_out ? probs_out + (size_t)row * n : nullptr;

    auto put_f = [&](int x, float a, float b) {             // block-uniform a, b -> slot [x][rank] of every CTA of the cluster
        if (tid < PLC_CL) {
            float* dst = cluster.map_shared_rank(&xf[x][0][0], tid);
            dst[rank * 2] = a; dst[rank * 2 + 1] = b;
        }
    };
    auto put_d = [&](int x, double a) {
        if (tid < PLC_CL) cluster.map_shared_rank(&xd[x][0], tid)[rank] = a;
    };
    auto put_i = [&](int x, int a, int b) {
        if (tid < PLC_CL) {
            int* dst = cluster.map_shared_rank(&xi[x][0][0], tid);
            dst[rank * 2] = a; dst[rank * 2 + 1] = b;
        }
    };

    // pass A: maxima
    float lmax = -INFINITY, rawmax = -INFINITY;
    for (int i = lo + tid; i < hi; i += PLC_THREADS) {
        const float raw = lg[i];
        rawmax = fmaxf(rawmax, raw);
        if (!token_suppressed(i, sc, v)) lmax = fmaxf(lmax, scaled ? __fdiv_rn(raw, p.temperature) : raw);
    }
    lmax = block_reduce(lmax, -INFINITY, OpMax(), red_f);
    rawmax = block_reduce(rawmax, -INFINITY, OpMax(), red_f);
    put_f(0, lmax, rawmax);
    cluster.sync();
    lmax = -INFINITY; rawmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) { lmax = fmaxf(lmax, xf[0][c][0]); rawmax = fmaxf(rawmax, xf[0][c][1]); }
    // pass B: sums
    float lsum = 0.0f, rawsum = 0.0f;
    for (int i = lo + tid; i < hi; i += PLC_THREADS) {
        const float raw = lg[i];
        if (p.want_nosp) rawsum += expf(raw - rawmax);
        if (!token_suppressed(i, sc, v)) lsum += expf((scaled ? __fdiv_rn(raw, p.temperature) : raw) - lmax);
    }
    lsum = block_reduce(lsum, 0.0f, OpAddF(), red_f);
    rawsum = block_reduce(rawsum, 0.0f, OpAddF(), red_f);
    put_f(1, lsum, rawsum);
    cluster.sync();
    lsum = 0.0f; rawsum = 0.0f;
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) { lsum += xf[1][c][0]; rawsum += xf[1][c][1]; }   // rank order: deterministic
    const float logsumexp = logf(lsum) + lmax;
    auto logprob_of = [&](int i) -> float {
        if (token_suppressed(i, sc, v)) return -INFINITY;
        const float l = scaled ? __fdiv_rn(lg[i], p.temperature) : lg[i];
        return l - logsumexp;
    };
    // pass C: timestamp logsumexp vs best text token
    float ts_max = -INFINITY, text_max = -INFINITY;
    for (int i = lo + tid; i < hi; i += PLC_THREADS) {
        const float lp = logprob_of(i);
        if (i >= v.beg) ts_max = fmaxf(ts_max, lp); else text_max = fmaxf(text_max, lp);
    }
    ts_max = block_reduce(ts_max, -INFINITY, OpMax(), red_f);
    text_max = block_reduce(text_max, -INFINITY, OpMax(), red_f);
    put_f(2, ts_max, text_max);
    cluster.sync();
    ts_max = -INFINITY; text_max = -INFINITY;
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) { ts_max = fmaxf(ts_max, xf[2][c][0]); text_max = fmaxf(text_max, xf[2][c][1]); }
    float ts_sum = 0.0f;
    for (int i = max(lo, v.beg) + tid; i < hi; i += PLC_THREADS) {
        const float lp = logprob_of(i);
        if (lp > -INFINITY) ts_sum += expf(lp - ts_max);
    }
    ts_sum = block_reduce(ts_sum, 0.0f, OpAddF(), red_f);
    put_f(3, ts_sum, 0.0f);
    cluster.sync();
    ts_sum = 0.0f;
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) ts_sum += xf[3][c][0];
    const float ts_logprob = ts_sum > 0.0f ? logf(ts_sum) + ts_max : -INFINITY;
    const bool text_off = ts_logprob > text_max;
    auto final_logprob = [&](int i) -> float {
        if (text_off && i < v.beg) return -INFINITY;
        return logprob_of(i);
    };

    // pass D: probabilities, argmax (first maximum wins), timestamp statistics
    float best_p = 0.0f; int best_i = INT_MAX;
    float tsb_p = 0.0f; int tsb_i = INT_MAX;
    double sum_ts = 0.0;
    for (int i = lo + tid; i < hi; i += PLC_THREADS) {
        const float lp = final_logprob(i);
        const float pr = lp > -INFINITY ? expf(lp) : 0.0f;
        if (logprobs_out) logprobs_out[(size_t)row * n + i] = lp;
        if (probs_out) probs_out[(size_t)row * n + i] = pr;
        if (pr > best_p) { best_p = pr; best_i = i; }
        if (i >= v.beg) {
            sum_ts += pr;
            if (pr > tsb_p) { tsb_p = pr; tsb_i = i; }
        }
    }
    {
        const float gb = block_reduce(best_p, 0.0f, OpMax(), red_f);
        int cand = (best_p == gb && best_i != INT_MAX) ? best_i : INT_MAX;
        const int gbi = block_reduce(cand, INT_MAX, OpMinI(), red_i);
        const float gt = block_reduce(tsb_p, 0.0f, OpMax(), red_f);
        cand = (tsb_p == gt && tsb_i != INT_MAX) ? tsb_i : INT_MAX;
        const int gti = block_reduce(cand, INT_MAX, OpMinI(), red_i);
        sum_ts = block_reduce(sum_ts, 0.0, OpAddD(), red_d);
        put_f(4, gb, gt);
        put_i(0, gbi, gti);
        put_d(0, sum_ts);
    }
    cluster.sync();
    float gbest_p = 0.0f, gts_p = 0.0f;
    int gbest_i = INT_MAX, gts_i = INT_MAX;
    sum_ts = 0.0;
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) { gbest_p = fmaxf(gbest_p, xf[4][c][0]); gts_p = fmaxf(gts_p, xf[4][c][1]); sum_ts += xd[0][c]; }
#pragma unroll
    for (int c = 0; c < PLC_CL; ++c) {   // ranges ascend with the rank: the first CTA that holds the maximum holds its lowest index
        if (xf[4][c][0] == gbest_p && xi[0][c][0] != INT_MAX) gbest_i = min(gbest_i, xi[0][c][0]);
        if (xf[4][c][1] == gts_p && xi[0][c][1] != INT_MAX) gts_i = min(gts_i, xi[0][c][1]);
    }

    int pick = gbest_i == INT_MAX ? 0 : gbest_i;
    if (p.mode == 1) {
        // inverse-CDF sampling as in process_logits_kernel: q_i = p_i / sum, cp = inclusive prefix sums of q (double), pick = first i
        // with cp[i] >= u; threads own contiguous chunks of this CTA's range, CTAs contiguous ranges of the vocabulary
        const int len = max(hi - lo, 0);
        const int chunk = (len + PLC_THREADS - 1) / PLC_THREADS;
        const int i0 = lo + tid * chunk, i1 = min(hi, i0 + chunk);
        auto prob_of = [&](int i) -> double {
            if (pr_cache) return (double)__ldcg(pr_cache + i);
            const float lp = final_logprob(i);
            return lp > -INFINITY ? (double)expf(lp) : 0.0;
        };
        double tot = 0.0;
        for (int i = i0; i < i1; ++i) tot += prob_of(i);
        tot = block_reduce(tot, 0.0, OpAddD(), red_d);
        put_d(1, tot);
        cluster.sync();
        double total = 0.0;
#pragma unroll
        for (int c = 0; c < PLC_CL; ++c) total += xd[1][c];
        double local = 0.0;
        for (int i = i0; i < i1; ++i) local += prob_of(i) / total;
        const int lane = tid & 31, warp = tid >> 5;
        double incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) scan_d[warp] = incl;
        if (tid == 0) s_pick = INT_MAX;
        __syncthreads();
        double woff = 0.0, cta_sum = 0.0;
        for (int w = 0; w < PLC_THREADS / 32; ++w) { if (w < warp) woff += scan_d[w]; cta_sum += scan_d[w]; }
        put_d(2, cta_sum);
        cluster.sync();
        double cta_off = 0.0;
        for (int c = 0; c < rank; ++c) cta_off += xd[2][c];
        double cp = cta_off + woff + incl - local;
        int found = INT_MAX;
        if (!(cp >= p.u)) {
            for (int i = i0; i < i1; ++i) {
                cp += prob_of(i) / total;
                if (cp >= p.u) { found = i; break; }
            }
        }
        if (found != INT_MAX) atomicMin(&s_pick, found);
        __syncthreads();
        put_i(1, s_pick, 0);
        cluster.sync();
        int g = INT_MAX;
#pragma unroll
        for (int c = 0; c < PLC_CL; ++c) g = min(g, xi[1][c][0]);
        pick = g == INT_MAX ? n - 1 : g;   // cp.back() is forced to 1.0 by the reference distribution
    }

    int n_topk = 0;
    if (p.mode == 2) {
        const int k = min(p.k, kMaxTopK);
        for (int round = 0; round < k; ++round) {
            float bl = -INFINITY; int bi = INT_MAX;
            for (int i = lo + tid; i < hi; i += PLC_THREADS) {
                bool taken = false;
                for (int c = 0; c < round; ++c) taken |= (s_chosen[c] == i);
                if (taken) continue;
                const float lp = final_logprob(i);
                if (lp > bl) { bl = lp; bi = i; }
            }
            const float gl = block_reduce(bl, -INFINITY, OpMax(), red_f);
            const int c2 = (bl == gl && bi != INT_MAX && gl > -INFINITY) ? bi : INT_MAX;
            const int gi = block_reduce(c2, INT_MAX, OpMinI(), red_i);
            put_f(5 + round, gl, 0.0f);
            put_i(2 + round, gi, 0);
            cluster.sync();
            float cl = -INFINITY; int ci = INT_MAX;
#pragma unroll
            for (int c = 0; c < PLC_CL; ++c) cl = fmaxf(cl, xf[5 + round][c][0]);
#pragma unroll
            for (int c = 0; c < PLC_CL; ++c)
                if (xf[5 + round][c][0] == cl && xi[2 + round][c][0] != INT_MAX) ci = min(ci, xi[2 + round][c][0]);
            if (ci == INT_MAX || !(cl > -INFINITY)) break;   // uniform across the cluster
            if (tid == 0) s_chosen[round] = ci;
            __syncthreads();
            n_topk = round + 1;
        }
    }

    if (rank == 0 && tid == 0) {
        SampleResult r;
        r.id = pick;
        const float lp = final_logprob(pick);
        r.plog = lp;
        r.p = lp > -INFINITY ? expf(lp) : 0.0f;
        r.tid = gts_i == INT_MAX ? -1 : gts_i;
        r.pt = (float)((double)gts_p / (sum_ts + 1e-10));
        r.ptsum = (float)sum_ts;
        if (pick >= v.beg) { r.tid = pick; r.pt = r.p; }
        r.no_speech_prob = p.want_nosp ? expf(lg[v.nosp] - (logf(rawsum) + rawmax)) : 0.0f;
        r.n_topk = n_topk;
        for (int c = 0; c < kMaxTopK; ++c) {
            if (c < n_topk) {
                const int id = s_chosen[c];
                const float l2 = final_logprob(id);
                r.topk_id[c] = id; r.topk_plog[c] = l2; r.topk_p[c] = expf(l2);
            } else { r.topk_id[c] = -1; r.topk_plog[c] = -INFINITY; r.topk_p[c] = 0.0f; }
        }
        results[row] = r;
    }
    cluster.sync();   // no CTA leaves while a peer may still write into its shared memory
    trace_end(tr);
}

void launch_process_logits(const float* logits, int ld, const SampleParams* params, SampleResult* results, int n_rows, const VocabIds& v,
                           float* logprobs_out, float* probs_out, cudaStream_t s) {
    if (n_rows <= 0) return;
    static const bool use_cluster = [] { const char* e = getenv("NOBS_WHISPER_K6_CLUSTER"); return !(e && *e == '0'); }();
    if (use_cluster)   // one row per cluster of PLC_CL CTAs (the cluster shape is a compile-time attribute of the kernel)
        launch_kernel(process_logits_cluster_kernel, dim3(PLC_CL, n_rows), dim3(PLC_THREADS), 0, s, true, logits, ld, params, results, v, logprobs_out, probs_out);
    else
        launch_kernel(process_logits_kernel, dim3(n_rows), dim3(PL_THREADS), 0, s, true, logits, ld, params, results, v, logprobs_out, probs_out);
    NOBS_COUNT_LAUNCH();
}

__global__ void lang_probs_kernel(const float* __restrict__ logits, VocabIds v, float* __restrict__ probs_out, int* __restrict__ best) {
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    const int tid = threadIdx.x;
    const float l = tid < v.n_lang ? logits[v.sot + 1 + tid] : -INFINITY;
    const float mx = block_reduce(l, -INFINITY, OpMax(), red_f);
    const float e = tid < v.n_lang ? expf(l - mx) : 0.0f;
    const float sum = block_reduce(e, 0.0f, OpAddF(), red_f);
    if (tid < v.n_lang && probs_out) probs_out[tid] = e / sum;
    const int cand = (tid < v.n_lang && l == mx) ? tid : INT_MAX;
    const int b = block_reduce(cand, INT_MAX, OpMinI(), red_i);
    if (tid == 0) *best = b;
}
void launch_lang_probs(const float* logits, const VocabIds& v, float* probs_out, int* best, cudaStream_t s) {
    prefer_max_shared_carveout(reinterpret_cast<const void*>(&lang_probs_kernel));
    lang_probs_kernel<<<1, 128, 0, s>>>(logits, v, probs_out, best);
    NOBS_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = from_f32<TO>(to_f32(in[i]));
}
template <typename TI, typename TO>
void launch_convert(const TI* in, TO* out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    convert_kernel<TI, TO><<<blocks, 256, 0, s>>>(in, out, n);
    NOBS_COUNT_LAUNCH();
}
template void launch_convert<float, bf16>(const float*, bf16*, size_t, cudaStream_t);
template void launch_convert<bf16, float>(const bf16*, float*, size_t, cudaStream_t);
template void launch_convert<float, float>(const float*, float*, size_t, cudaStream_t);

template <typename TI, typename TO>
__global__ void convert_2d_kernel(const TI* __restrict__ in, size_t ld_in, TO* __restrict__ out, size_t ld_out, int rows, int cols) {
    const int r = blockIdx.y;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
        out[(size_t)r * ld_out + c] = from_f32<TO>(to_f32(in[(size_t)r * ld_in + c]));
}
template <typename TI, typename TO>
void launch_convert_2d(const TI* in, size_t ld_in, TO* out, size_t ld_out, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((cols + 255) / 256, rows);
    convert_2d_kernel<TI, TO><<<grid, 256, 0, s>>>(in, ld_in, out, ld_out, rows, cols);
    NOBS_COUNT_LAUNCH();
}
template void launch_convert_2d<float, float>(const float*, size_t, float*, size_t, int, int, cudaStream_t);
template void launch_convert_2d<bf16, float>(const bf16*, size_t, float*, size_t, int, int, cudaStream_t);

}  // namespace nobs
