// Small-batch decode step (B <= 16 utterances per GPU): weight-streaming GEMV kernels with the LayerNorm fused in front.
//
// At small batch the decode step is launch / latency bound (profiles/r01_decode_step_us.md: 268 kernels at ~8 us, 14x above
// the weight-streaming floor): the tcgen05 path pays a TMEM allocation, mbarrier set-up and a 128-row MMA tile for one to
// sixteen rows, plus a separate LayerNorm launch in front of three of the six GEMMs of a layer.  Here
//   ln_gemv_kernel    out[m, n] = act( LN(x[m, :]) . W[n, :] + bias[n] )      LN1+qkv, LN2+cross-q, LN3+fc1, final LN + LM head
//   gemv_res_kernel   x[m, n]  += bias[n] + a[m, :] . W[n, :]                 out-proj, cross-out, fc2 (residual in place)
// so a decoder layer is 8 launches instead of 11 and no kernel touches tensor memory.
// One CTA = 4 warps; the CTA first stages the (normalised) activation rows as bf16 in shared memory, then every warp
// streams the weight rows of 2 output columns with 16-byte L1-bypassing loads and keeps M x 2 fp32 accumulators.
// Semantics as the reference layers: LayerNorm eps 1e-5 (layers/normalization.py:6-30), ColumnLinear / RowLinear
// (layers/linear.py:38-139), erf GELU; activations are rounded to bf16 exactly where the tcgen05 path rounds them.
#include "wb_internal.h"

namespace wb {

namespace {
constexpr int GV_WARPS = 4, GV_THREADS = GV_WARPS * 32, GV_COLS = 2, GV_MMAX = 16;

// dot products of this warp's GV_COLS weight rows with the M staged activation rows; a_s: [M][K] bf16 in shared memory
template <int M_T>
__device__ __forceinline__ void gemv_accumulate(const bf16* __restrict__ W, long long ldw, int n0, int N, int K,
                                                const bf16* __restrict__ a_s, int lane, float (&acc)[M_T][GV_COLS]) {
#pragma unroll
    for (int m = 0; m < M_T; ++m)
#pragma unroll
        for (int c = 0; c < GV_COLS; ++c) acc[m][c] = 0.f;
    for (int k0 = lane * 8; k0 < K; k0 += 32 * 8) {
        float wf[GV_COLS][8];
#pragma unroll
        for (int c = 0; c < GV_COLS; ++c) {
            const int n = min(n0 + c, N - 1);
            ld16_stream(W + (size_t)n * ldw + k0).unpack(wf[c]);
        }
#pragma unroll
        for (int m = 0; m < M_T; ++m) {
            float af[8];
            ld16(a_s + (size_t)m * K + k0).unpack(af);
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[m][c] = fmaf(af[i], wf[c][i], acc[m][c]);
        }
    }
#pragma unroll
    for (int m = 0; m < M_T; ++m)
#pragma unroll
        for (int c = 0; c < GV_COLS; ++c) acc[m][c] = warp_sum(acc[m][c]);
}

// LayerNorm of row m of x (fp32, d <= 1024) into a_s[m][:] as bf16; executed by one warp
__device__ __forceinline__ void ln_row_to_smem(const float* __restrict__ xr, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, bf16* __restrict__ dst, int d, float eps, int lane) {
    const int nvec = d >> 2;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            v[i] = reinterpret_cast<const float4*>(xr)[idx];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / (float)d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + e * e);
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 g = reinterpret_cast<const float4*>(gamma)[idx], b = reinterpret_cast<const float4*>(beta)[idx];
            __nv_bfloat162 p0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            __nv_bfloat162 p1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(dst + idx * 4) = u;
        }
    }
}

template <int M_T, typename TOut>
__global__ void __launch_bounds__(GV_THREADS) ln_gemv_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, const bf16* __restrict__ W,
                                                             long long ldw, const float* __restrict__ bias, TOut* __restrict__ out,
                                                             long long ldo, int M, int N, int d, int act,
                                                             const int* __restrict__ active) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    bf16* a_s = reinterpret_cast<bf16*>(gv_smem);            // [M_T][d]
    pdl_wait();
    pdl_trigger();
    const bool run = (active == nullptr) || (*active != 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = warp; m < M_T; m += GV_WARPS) {
        if (m < M) {
            ln_row_to_smem(x + (size_t)m * d, gamma, beta, a_s + (size_t)m * d, d, eps, lane);
        } else {
            for (int i = lane; i < d / 8; i += 32) reinterpret_cast<uint4*>(a_s + (size_t)m * d)[i] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    const int n0 = (blockIdx.x * GV_WARPS + warp) * GV_COLS;
    if (n0 >= N) return;
    float acc[M_T][GV_COLS];
    gemv_accumulate<M_T>(W, ldw, n0, N, d, a_s, lane, acc);
    if (!run) return;
    // lane m finishes row m (all lanes hold the reduced sums)
#pragma unroll
    for (int m = 0; m < M_T; ++m) {
        if (lane == m && m < M) {
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c) {
                const int n = n0 + c;
                if (n < N) {
                    float v = acc[m][c] + (bias != nullptr ? bias[n] : 0.f);
                    if (act == 1) v = gelu_erf_fast(v);
                    out[(size_t)m * ldo + n] = from_f32<TOut>(v);
                }
            }
        }
    }
}

template <int M_T>
__global__ void __launch_bounds__(GV_THREADS) gemv_res_kernel(const bf16* __restrict__ a, long long lda, const bf16* __restrict__ W,
                                                              long long ldw, const float* __restrict__ bias, float* __restrict__ x,
                                                              long long ldx, int M, int N, int K, const int* __restrict__ active) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    bf16* a_s = reinterpret_cast<bf16*>(gv_smem);            // [M_T][K]
    pdl_wait();
    pdl_trigger();
    const bool run = (active == nullptr) || (*active != 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vec_per_row = K / 8;
    for (int i = threadIdx.x; i < M_T * vec_per_row; i += GV_THREADS) {
        const int m = i / vec_per_row, j = i - m * vec_per_row;
        reinterpret_cast<uint4*>(a_s + (size_t)m * K)[j] =
            m < M ? reinterpret_cast<const uint4*>(a + (size_t)m * lda)[j] : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int n0 = (blockIdx.x * GV_WARPS + warp) * GV_COLS;
    if (n0 >= N) return;
    float acc[M_T][GV_COLS];
    gemv_accumulate<M_T>(W, ldw, n0, N, K, a_s, lane, acc);
    if (!run) return;
#pragma unroll
    for (int m = 0; m < M_T; ++m) {
        if (lane == m && m < M) {
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c) {
                const int n = n0 + c;
                if (n < N) x[(size_t)m * ldx + n] += acc[m][c] + (bias != nullptr ? bias[n] : 0.f);   // this thread owns (m, n)
            }
        }
    }
}

template <int M_T, typename TOut>
void launch_ln_gemv(const float* x, const float* gamma, const float* beta, float eps, const bf16* W, long long ldw, const float* bias,
                    TOut* out, long long ldo, int M, int N, int d, int act, const int* active, cudaStream_t st) {
    const size_t smem = (size_t)M_T * d * sizeof(bf16);
    static bool configured = false;
    if (!configured) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(ln_gemv_kernel<M_T, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, GV_MMAX * 1024 * 2));
        configured = true;
    }
    const int grid = ceil_div(N, GV_WARPS * GV_COLS);
    launch_kernel(ln_gemv_kernel<M_T, TOut>, dim3(grid), dim3(GV_THREADS), smem, st, true, x, gamma, beta, eps, W, ldw, bias, out, ldo, M,
                  N, d, act, active);
}

template <int M_T>
void launch_gemv_res(const bf16* a, long long lda, const bf16* W, long long ldw, const float* bias, float* x, long long ldx, int M, int N,
                     int K, const int* active, cudaStream_t st) {
    const size_t smem = (size_t)M_T * K * sizeof(bf16);
    static bool configured = false;
    if (!configured) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(gemv_res_kernel<M_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, GV_MMAX * 4096 * 2));
        configured = true;
    }
    const int grid = ceil_div(N, GV_WARPS * GV_COLS);
    launch_kernel(gemv_res_kernel<M_T>, dim3(grid), dim3(GV_THREADS), smem, st, true, a, lda, W, ldw, bias, x, ldx, M, N, K, active);
}
}  // namespace

bool skinny_gemv_supported(int M, int K, int dtype) {
    return dtype == BF16 && M >= 1 && M <= GV_MMAX && K % 8 == 0 && K <= 4096;
}

// out (bf16 or fp32) = act(LN(x) W^T + bias); x fp32 [M, d], d <= 1024
void ln_gemv(const float* x, const float* gamma, const float* beta, float eps, const void* W, long long ldw, const float* bias, void* out,
             long long ldo, int out_dtype, int M, int N, int d, int act, const int* active, cudaStream_t st) {
    WB_REQUIRE(skinny_gemv_supported(M, d, BF16) && d <= 1024 && ldw % 8 == 0, "ln_gemv: M <= 16, d % 8 == 0, d <= 1024");
    const bf16* w = reinterpret_cast<const bf16*>(W);
#define WB_LN_GEMV(MT)                                                                                                              \
    do {                                                                                                                            \
        if (out_dtype == F32) launch_ln_gemv<MT, float>(x, gamma, beta, eps, w, ldw, bias, (float*)out, ldo, M, N, d, act, active, st); \
        else launch_ln_gemv<MT, bf16>(x, gamma, beta, eps, w, ldw, bias, (bf16*)out, ldo, M, N, d, act, active, st);                 \
    } while (0)
    if (M <= 1) WB_LN_GEMV(1);
    else if (M <= 4) WB_LN_GEMV(4);
    else if (M <= 8) WB_LN_GEMV(8);
    else WB_LN_GEMV(16);
#undef WB_LN_GEMV
}

// x[m, n] += bias[n] + a[m, :] . W[n, :]; a bf16 [M, K], x fp32 [M, N] updated in place
void gemv_residual(const void* a, long long lda, const void* W, long long ldw, const float* bias, float* x, long long ldx, int M, int N,
                   int K, const int* active, cudaStream_t st) {
    WB_REQUIRE(skinny_gemv_supported(M, K, BF16) && lda % 8 == 0 && ldw % 8 == 0, "gemv_residual: M <= 16, K % 8 == 0, K <= 4096");
    const bf16* ap = reinterpret_cast<const bf16*>(a);
    const bf16* w = reinterpret_cast<const bf16*>(W);
    if (M <= 1) launch_gemv_res<1>(ap, lda, w, ldw, bias, x, ldx, M, N, K, active, st);
    else if (M <= 4) launch_gemv_res<4>(ap, lda, w, ldw, bias, x, ldx, M, N, K, active, st);
    else if (M <= 8) launch_gemv_res<8>(ap, lda, w, ldw, bias, x, ldx, M, N, K, active, st);
    else launch_gemv_res<16>(ap, lda, w, ldw, bias, x, ldx, M, N, K, active, st);
}

}  // namespace wb
