// LayerNorm (eps 1e-5, fp32 statistics) over the fp32 residual stream.
// Reference semantics: nn.LayerNorm in the oracle (modeling_whisper.py:608,613,675-686,940,1060);
// TRT-LLM LayerNorm eps=1e-5 (layers/normalization.py:6-30).  HBM-bound: one warp per row, the whole
// row lives in registers (d <= 1024 -> 8 float4 per lane), 16-byte coalesced loads and stores.
#include "wb_internal.h"

namespace wb {

// kPreAdd: x[r,:] += bias + sum_s parts[s][r,:] first (consumer side of a split-K GEMM, fixed summation order),
// the updated residual row is written back in place.
template <typename TOut, bool kPreAdd>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, LnPreAdd pre, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, TOut* __restrict__ out,
                                                        float* __restrict__ out2, int rows, int d, float eps,
                                                        const int* __restrict__ active) {
    pdl_wait();
    pdl_trigger();
    if (active != nullptr && *active == 0) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nvec = d >> 2;
    float4* xr = reinterpret_cast<float4*>(x + (size_t)warp * d);
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            v[i] = xr[idx];
            if constexpr (kPreAdd) {
                float4 t = pre.bias != nullptr ? reinterpret_cast<const float4*>(pre.bias)[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < pre.n_parts; ++p) {
                    const float4 q = reinterpret_cast<const float4*>(pre.parts + (size_t)p * pre.part_stride + (size_t)warp * d)[idx];
                    t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
                }
                v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
                xr[idx] = v[i];
            }
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
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 g = g4[idx], b = b4[idx];
            float y[4];
            y[0] = (v[i].x - mean) * rstd * g.x + b.x;
            y[1] = (v[i].y - mean) * rstd * g.y + b.y;
            y[2] = (v[i].z - mean) * rstd * g.z + b.z;
            y[3] = (v[i].w - mean) * rstd * g.w + b.w;
            TOut* o = out + (size_t)warp * d + idx * 4;
            if constexpr (sizeof(TOut) == 4) {
                *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(y[2], y[3]);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&p0);
                u.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(o) = u;
            }
            if (out2 != nullptr)
                *reinterpret_cast<float4*>(out2 + (size_t)warp * d + idx * 4) = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
}

// Few rows (the decode step: one row per utterance): ONE CTA PER ROW, one float4 per thread, so that the row, the bias
// and every split-K slab are requested in a single round trip (the warp-per-row kernel above walks them one after
// the other: 12-15 us for 256 rows under ncu).  Block reduction through shared memory, two passes like above.
template <typename TOut, bool kPreAdd>
__global__ void __launch_bounds__(256) layernorm_row_kernel(float* __restrict__ x, LnPreAdd pre, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, TOut* __restrict__ out, int d,
                                                            float eps, const int* __restrict__ active) {
    constexpr int MAXP = 8;
    __shared__ float red[2][8];
    pdl_wait();
    pdl_trigger();
    // the stop flag is only needed before the stores: its load overlaps the row loads instead of preceding them
    const bool run = (active == nullptr) || (*active != 0);
    const int row = blockIdx.x, t = threadIdx.x, nvec = d >> 2;
    const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    const bool on = t < nvec;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
        v = reinterpret_cast<const float4*>(x + (size_t)row * d)[t];
        if constexpr (kPreAdd) {
            float4 b = pre.bias != nullptr ? reinterpret_cast<const float4*>(pre.bias)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 q[MAXP];
#pragma unroll
            for (int p = 0; p < MAXP; ++p)   // all slabs in flight together
                q[p] = p < pre.n_parts ? reinterpret_cast<const float4*>(pre.parts + (size_t)p * pre.part_stride + (size_t)row * d)[t]
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < MAXP; ++p) { b.x += q[p].x; b.y += q[p].y; b.z += q[p].z; b.w += q[p].w; }   // fixed order
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            if (run) reinterpret_cast<float4*>(x + (size_t)row * d)[t] = v;
        }
    }
    const float4 g = on ? reinterpret_cast<const float4*>(gamma)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 be = on ? reinterpret_cast<const float4*>(beta)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += red[0][w];
    const float mean = tot / (float)d;
    const float a = v.x - mean, b2 = v.y - mean, c = v.z - mean, e = v.w - mean;
    float ss = on ? (a * a + b2 * b2) + (c * c + e * e) : 0.f;
    ss = warp_sum(ss);
    if (lane == 0) red[1][warp] = ss;
    __syncthreads();
    float tot2 = 0.f;
    for (int w = 0; w < nwarps; ++w) tot2 += red[1][w];
    const float rstd = rsqrtf(tot2 / (float)d + eps);
    if (on && run) {
        const float y0 = a * rstd * g.x + be.x, y1 = b2 * rstd * g.y + be.y, y2 = c * rstd * g.z + be.z, y3 = e * rstd * g.w + be.w;
        TOut* o = out + (size_t)row * d + t * 4;
        if constexpr (sizeof(TOut) == 4) {
            *reinterpret_cast<float4*>(o) = make_float4(y0, y1, y2, y3);
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(y2, y3);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(o) = u;
        }
    }
}

constexpr int LN_ROW_KERNEL_MAX_ROWS = 2048;

template <bool kPreAdd>
static void launch_ln_rows(float* x, const LnPreAdd& pre, const float* gamma, const float* beta, void* out, int out_dtype, int rows,
                           int d, float eps, const int* active, cudaStream_t stream) {
    const int threads = ((d / 4) + 31) / 32 * 32;
    if (out_dtype == F32)
        launch_kernel(layernorm_row_kernel<float, kPreAdd>, dim3(rows), dim3(threads), 0, stream, true, x, pre, gamma, beta, (float*)out, d, eps, active);
    else
        launch_kernel(layernorm_row_kernel<bf16, kPreAdd>, dim3(rows), dim3(threads), 0, stream, true, x, pre, gamma, beta, (bf16*)out, d, eps, active);
}

void layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_dtype, float* out2,
               int rows, int d, float eps, const int* active, cudaStream_t stream) {
    WB_REQUIRE(d % 4 == 0 && d <= 1024, "layernorm supports d % 4 == 0, d <= 1024");
    if (rows == 0) return;
    if (rows <= LN_ROW_KERNEL_MAX_ROWS && out2 == nullptr) {
        launch_ln_rows<false>(const_cast<float*>(x), LnPreAdd(), gamma, beta, out, out_dtype, rows, d, eps, active, stream);
        return;
    }
    const int warps_per_block = 8;
    dim3 grid(ceil_div(rows, warps_per_block)), block(warps_per_block * 32);
    LnPreAdd none;
    float* xm = const_cast<float*>(x);   // never written without kPreAdd
    if (out_dtype == F32)
        launch_kernel(layernorm_kernel<float, false>, grid, block, 0, stream, true, xm, none, gamma, beta, (float*)out, out2, rows, d, eps, active);
    else
        launch_kernel(layernorm_kernel<bf16, false>, grid, block, 0, stream, true, xm, none, gamma, beta, (bf16*)out, out2, rows, d, eps, active);
}

void layernorm_preadd(float* x, const LnPreAdd& pre, const float* gamma, const float* beta, void* out, int out_dtype,
                      int rows, int d, float eps, const int* active, cudaStream_t stream) {
    WB_REQUIRE(d % 4 == 0 && d <= 1024, "layernorm supports d % 4 == 0, d <= 1024");
    WB_REQUIRE(pre.n_parts >= 0 && (pre.n_parts == 0 || pre.parts != nullptr) && pre.part_stride % 4 == 0, "bad pre-add slabs");
    WB_REQUIRE(pre.n_parts <= 8, "at most 8 split-K slabs");
    if (rows == 0) return;
    if (rows <= LN_ROW_KERNEL_MAX_ROWS) {
        launch_ln_rows<true>(x, pre, gamma, beta, out, out_dtype, rows, d, eps, active, stream);
        return;
    }
    const int warps_per_block = 4;
    dim3 grid(ceil_div(rows, warps_per_block)), block(warps_per_block * 32);
    if (out_dtype == F32)
        launch_kernel(layernorm_kernel<float, true>, grid, block, 0, stream, true, x, pre, gamma, beta, (float*)out, (float*)nullptr, rows, d, eps, active);
    else
        launch_kernel(layernorm_kernel<bf16, true>, grid, block, 0, stream, true, x, pre, gamma, beta, (bf16*)out, (float*)nullptr, rows, d, eps, active);
}

}  // namespace wb
