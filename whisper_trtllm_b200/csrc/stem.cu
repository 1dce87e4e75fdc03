// Encoder stem helpers.  conv1 (80 -> d, k=3, pad 1) runs as a GEMM over an im2col'd log-mel
// (K = 3*80 = 240, zero padded to 256); conv2 (d -> d, k=3, stride 2, pad 1) needs no gather at all because
// conv1's output is written time-major with one zero row in front of every utterance, so row t' of the
// conv2 operand is the contiguous window h1[2t' .. 2t'+2][:] (lda = 2d).  See runtime.cu.
// Reference: oracle WhisperEncoder.forward conv1/conv2 (modeling_whisper.py:934-935, 992-993);
// TRT-LLM used Conv2d with a (1,3) kernel (models/whisper/model.py:77-79, 96-100).
#include "wb_internal.h"

namespace wb {

namespace {
constexpr int TT = 32;  // time steps per block

template <typename T>
__global__ void __launch_bounds__(256) im2col_conv1_kernel(const float* __restrict__ mel, T* __restrict__ out, int n_mels,
                                                           int Tlen, int kpad) {
    extern __shared__ float s[];  // [n_mels][TT + 2]
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const float* mb = mel + (size_t)b * n_mels * Tlen;
    for (int i = threadIdx.x; i < n_mels * (TT + 2); i += blockDim.x) {
        const int c = i / (TT + 2), tt = i - c * (TT + 2);
        const int t = t0 + tt - 1;
        s[i] = (t >= 0 && t < Tlen) ? mb[(size_t)c * Tlen + t] : 0.f;
    }
    __syncthreads();
    const int K = 3 * n_mels;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        const int tap = k / n_mels, c = k - tap * n_mels;
        for (int tt = 0; tt < TT; ++tt) {
            const int t = t0 + tt;
            if (t >= Tlen) break;
            const float v = (k < K) ? s[c * (TT + 2) + tt + tap] : 0.f;
            out[((size_t)b * Tlen + t) * kpad + k] = from_f32<T>(v);
        }
    }
}

template <typename TIn, typename TOut>
__global__ void cast_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = from_f32<TOut>(to_f32(in[i]));
}
}  // namespace

void im2col_conv1(const float* mel, void* out, int out_dtype, int B, int n_mels, int T, int kpad, cudaStream_t stream) {
    WB_REQUIRE(kpad >= 3 * n_mels, "kpad too small");
    dim3 grid(ceil_div(T, TT), B), block(256);
    const size_t smem = (size_t)n_mels * (TT + 2) * sizeof(float);
    if (out_dtype == F32) im2col_conv1_kernel<float><<<grid, block, smem, stream>>>(mel, (float*)out, n_mels, T, kpad);
    else im2col_conv1_kernel<bf16><<<grid, block, smem, stream>>>(mel, (bf16*)out, n_mels, T, kpad);
    WB_CHECK_LAUNCH();
}

void cast(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream) {
    if (n == 0) return;
    const int block = 256;
    const int grid = (int)std::min<long long>((n + block - 1) / block, 148 * 16);
    if (in_dtype == F32 && out_dtype == BF16) cast_kernel<float, bf16><<<grid, block, 0, stream>>>((const float*)in, (bf16*)out, n);
    else if (in_dtype == BF16 && out_dtype == F32) cast_kernel<bf16, float><<<grid, block, 0, stream>>>((const bf16*)in, (float*)out, n);
    else if (in_dtype == F32 && out_dtype == F32) cast_kernel<float, float><<<grid, block, 0, stream>>>((const float*)in, (float*)out, n);
    else cast_kernel<bf16, bf16><<<grid, block, 0, stream>>>((const bf16*)in, (bf16*)out, n);
    WB_CHECK_LAUNCH();
}

}  // namespace wb
