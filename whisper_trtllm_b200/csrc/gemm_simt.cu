// CUDA-core GEMM with true fp32 FFMA accumulation: the exactness path (token-identical to the fp32 oracle,
// SURVEY.md §7 "Token-exact fp32 parity") and the fallback when the tcgen05 kernel does not apply.
// C[M,N] = A[M,K] * W[N,K]^T, both operands K-major (nn.Linear layout; the reference's ColumnLinear /
// RowLinear, layers/linear.py:38-139, and the oracle's F.linear).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, register-staged double buffering.
#include "wb_epilogue.cuh"

namespace wb {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename T> __device__ __forceinline__ void load4(const T* p, float* f);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float* f) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long long lda, const T* __restrict__ W,
                                                        long long ldw, int K, EpiParams ep,
                                                        const int* __restrict__ active) {
    pdl_wait();
    pdl_trigger();
    if (active != nullptr && *active == 0) return;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // loader mapping: each thread brings 4 consecutive k of one row of A and of W
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int am = min(m0 + lrow, ep.M - 1), wn = min(n0 + lrow, ep.N - 1);
    const T* ap = A + (long long)am * lda + lk;
    const T* wp = W + (long long)wn * ldw + lk;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    load4<T>(ap, ra);
    load4<T>(wp, rb);
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[0][lk + i][lrow] = ra[i]; Bs[0][lk + i][lrow] = rb[i]; }
    __syncthreads();

    const int nk = K / BK;
    for (int kb = 0; kb < nk; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nk) {
            load4<T>(ap + (long long)(kb + 1) * BK, ra);
            load4<T>(wp + (long long)(kb + 1) * BK, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { As[cur ^ 1][lk + i][lrow] = ra[i]; Bs[cur ^ 1][lk + i][lrow] = rb[i]; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const EpiRow r = epi_row(ep, m0 + ty * 4 + i);
        epi_store<4, false>(ep, r, n0 + tx * 4, acc[i]);
    }
}
}  // namespace

void gemm_simt(const GemmArgs& a, cudaStream_t stream) {
    validate_gemm_common(a);
    WB_REQUIRE(a.lda % 4 == 0 && a.ldw % 4 == 0, "lda/ldw must be multiples of 4");
    EpiParams ep = make_epi(a);
    dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM)), block(256);
    WB_REQUIRE(grid.y <= 65535, "M too large for the SIMT GEMM grid");
    if (a.in_dtype == F32)
        launch_kernel(gemm_simt_kernel<float>, grid, block, 0, stream, true, (const float*)a.A, a.lda, (const float*)a.W, a.ldw, a.K, ep, a.active);
    else
        launch_kernel(gemm_simt_kernel<bf16>, grid, block, 0, stream, true, (const bf16*)a.A, a.lda, (const bf16*)a.W, a.ldw, a.K, ep, a.active);
}

}  // namespace wb
