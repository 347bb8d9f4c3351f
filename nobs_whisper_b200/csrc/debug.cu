// Test hooks that run single kernels of the bf16 path on caller-supplied data (C ABI, used by
// tests/test_gpu_kernels_bf16.py).  Not on any product path.
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"
#include "gemm_sm100.cuh"

namespace nobs { void set_last_error(const std::string& e); }
using namespace nobs;

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16) == cudaSuccess; }
};
}  // namespace

extern "C" int whisper_b200_debug_gemm_bf16(int M, int N, int K, int lda, const float* A, size_t a_elems, const float* W, const float* bias, int act,
                                            const float* res, int res_mod, int win_rows, int valid_rows, int out_f32, float* C_out) {
    if (M <= 0 || N <= 0 || K <= 0 || !A || !W || !C_out) return -1;
    const size_t w_elems = (size_t)N * K, c_elems = (size_t)M * N;
    const int res_rows = res_mod > 0 ? res_mod : M;
    DevBuf dAf, dWf, dA, dW, dBias, dRes, dC, dCf;
    if (!dAf.alloc(a_elems * 4) || !dWf.alloc(w_elems * 4) || !dA.alloc(a_elems * 2) || !dW.alloc(w_elems * 2) || !dBias.alloc((size_t)N * 4) ||
        !dRes.alloc((size_t)res_rows * N * 4) || !dC.alloc(c_elems * 4) || !dCf.alloc(c_elems * 4)) {
        set_last_error("debug_gemm: cudaMalloc failed");
        return -2;
    }
    cudaStream_t s = nullptr;
    cudaMemcpy(dAf.p, A, a_elems * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dWf.p, W, w_elems * 4, cudaMemcpyHostToDevice);
    launch_convert<float, bf16>((const float*)dAf.p, (bf16*)dA.p, a_elems, s);
    launch_convert<float, bf16>((const float*)dWf.p, (bf16*)dW.p, w_elems, s);
    Epilogue e;
    if (bias) { cudaMemcpy(dBias.p, bias, (size_t)N * 4, cudaMemcpyHostToDevice); e.bias = (const float*)dBias.p; }
    if (res) { cudaMemcpy(dRes.p, res, (size_t)res_rows * N * 4, cudaMemcpyHostToDevice); e.res = (const float*)dRes.p; e.res_ld = N; e.res_mod = res_mod; }
    e.act = act;
    e.win_rows = win_rows;
    e.valid_rows = valid_rows;
    cudaMemset(dC.p, 0xff, c_elems * 4);
    if (!launch_gemm_bf16_sm100((const bf16*)dA.p, lda, (const bf16*)dW.p, K, dC.p, N, out_f32 != 0, M, N, K, e, s)) {
        set_last_error(std::string("debug_gemm: ") + sm100_last_error());
        return -3;
    }
    if (out_f32) {
        cudaMemcpyAsync(C_out, dC.p, c_elems * 4, cudaMemcpyDeviceToHost, s);
    } else {
        launch_convert<bf16, float>((const bf16*)dC.p, (float*)dCf.p, c_elems, s);
        cudaMemcpyAsync(C_out, dCf.p, c_elems * 4, cudaMemcpyDeviceToHost, s);
    }
    const cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) {
        set_last_error(std::string("debug_gemm: ") + cudaGetErrorString(err));
        return -4;
    }
    return 0;
}

// Encoder attention test hook: qkv fp32 host [n_win*1536][3*d] (q|k|v, rounded to bf16 on the device),
// out fp32 host [n_win*1536][d].  use_simt != 0 runs the CUDA-core kernel instead of the tcgen05 one.
extern "C" int whisper_b200_debug_enc_attention(int n_win, int n_head, const float* qkv, float* out, int use_simt) {
    if (n_win <= 0 || n_head <= 0 || !qkv || !out) return -1;
    const int d = n_head * 64;
    const size_t rows = (size_t)n_win * kWinRows, in_elems = rows * 3 * d, out_elems = rows * d;
    DevBuf dInF, dIn, dOut, dOutF;
    if (!dInF.alloc(in_elems * 4) || !dIn.alloc(in_elems * 2) || !dOut.alloc(out_elems * 2) || !dOutF.alloc(out_elems * 4)) return -2;
    cudaStream_t s = nullptr;
    cudaMemcpy(dInF.p, qkv, in_elems * 4, cudaMemcpyHostToDevice);
    launch_convert<float, bf16>((const float*)dInF.p, (bf16*)dIn.p, in_elems, s);
    cudaMemset(dOut.p, 0, out_elems * 2);
    if (use_simt) launch_enc_attention_simt<bf16>((const bf16*)dIn.p, (bf16*)dOut.p, n_win, n_head, d, s);
    else if (!launch_enc_attention_bf16_sm100((const bf16*)dIn.p, (bf16*)dOut.p, n_win, n_head, d, s)) {
        set_last_error(std::string("debug_enc_attention: ") + sm100_last_error());
        return -3;
    }
    launch_convert<bf16, float>((const bf16*)dOut.p, (float*)dOutF.p, out_elems, s);
    cudaMemcpyAsync(out, dOutF.p, out_elems * 4, cudaMemcpyDeviceToHost, s);
    const cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) { set_last_error(std::string("debug_enc_attention: ") + cudaGetErrorString(err)); return -4; }
    return 0;
}
