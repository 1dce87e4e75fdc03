// GEMM epilogue shared by the SIMT (fp32 exactness path) and tcgen05 (bf16 speed path) kernels:
// bias -> erf-GELU -> fp32 residual -> typed store with row remap / head-split scatter (see wb_internal.h).
#pragma once
#include "wb_internal.h"

namespace wb {

struct EpiParams {
    const float* bias;
    const float* res; long long ldres; int res_periodic;
    void* out; long long ldo; int out_dtype;
    void* out2; long long ldo2;
    int M, N, act;
    int m_period_in, m_valid, m_period_out, m_out_offset;
    int out_mode, hs_heads, hs_batch, hs_b0;
    float* am_val; int* am_idx; int am_stride;                      // fused logits processors + argmax (see GemmArgs)
    const unsigned char* am_mask; const StepState* am_state; const int* am_row_len; int am_begin;
};

inline EpiParams make_epi(const GemmArgs& a) {
    EpiParams p;
    p.bias = a.bias; p.res = a.res; p.ldres = a.ldres; p.res_periodic = a.res_periodic;
    p.out = a.out; p.ldo = a.ldo; p.out_dtype = a.out_dtype; p.out2 = a.out2; p.ldo2 = a.ldo2;
    p.M = a.M; p.N = a.N; p.act = a.act;
    p.m_period_in = a.m_period_in; p.m_valid = a.m_valid; p.m_period_out = a.m_period_out;
    p.m_out_offset = a.m_out_offset;
    p.out_mode = a.out_mode; p.hs_heads = a.hs_heads; p.hs_batch = a.hs_batch; p.hs_b0 = a.hs_b0;
    p.am_val = a.am_val; p.am_idx = a.am_idx; p.am_stride = a.am_stride;
    p.am_mask = a.am_mask; p.am_state = a.am_state; p.am_row_len = a.am_row_len; p.am_begin = a.am_begin;
    return p;
}

struct EpiRow {
    bool valid;
    int g, r_in;
    long long out_row, res_row;
};

__device__ __forceinline__ EpiRow epi_row(const EpiParams& p, int m) {
    EpiRow r;
    r.valid = m < p.M;
    if (p.m_period_in > 0) {
        r.g = m / p.m_period_in;
        r.r_in = m - r.g * p.m_period_in;
        r.valid = r.valid && (r.r_in < p.m_valid);
        r.out_row = (long long)r.g * p.m_period_out + r.r_in + p.m_out_offset;
    } else {
        r.g = 0;
        r.r_in = m;
        r.out_row = m;
    }
    r.res_row = p.res_periodic ? r.r_in : r.out_row;
    return r;
}

// Same as epi_row for a row m >= row0 close to row0: g0 = row0 / m_period_in is computed once per tile, no division per row.
__device__ __forceinline__ EpiRow epi_row_near(const EpiParams& p, int m, int g0) {
    EpiRow r;
    r.valid = m < p.M;
    if (p.m_period_in > 0) {
        r.g = g0;
        r.r_in = m - g0 * p.m_period_in;
        while (r.r_in >= p.m_period_in) { r.r_in -= p.m_period_in; r.g += 1; }   // at most 32 / period + 1 trips
        r.valid = r.valid && (r.r_in < p.m_valid);
        r.out_row = (long long)r.g * p.m_period_out + r.r_in + p.m_out_offset;
    } else {
        r.g = 0;
        r.r_in = m;
        r.out_row = m;
    }
    r.res_row = p.res_periodic ? r.r_in : r.out_row;
    return r;
}

// typed store of NV finished values (row remap / head-split scatter, optional fp32 copy)
template <int NV>
__device__ __forceinline__ void epi_write(const EpiParams& p, const EpiRow& r, int n0, const float* v) {
    if (!r.valid || n0 >= p.N) return;
    long long off;
    if (p.out_mode == 0) {
        off = r.out_row * p.ldo + n0;
    } else {
        const int dm = p.hs_heads * 64;
        const int kv = n0 / dm;
        const int c = n0 - kv * dm;
        const int h = c >> 6, j = c & 63;
        off = ((((long long)kv * p.hs_batch + p.hs_b0 + r.g) * p.hs_heads + h) * p.m_valid + r.r_in) * 64 + j;
    }
    if (p.out_dtype == F32) {
        float* o = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
        bf16* o = reinterpret_cast<bf16*>(p.out) + off;
        uint32_t w[NV / 2];
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        if constexpr (NV == 8) *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
        else *reinterpret_cast<uint2*>(o) = make_uint2(w[0], w[1]);
    }
    if (p.out2 != nullptr) {
        float* o = reinterpret_cast<float*>(p.out2) + r.out_row * p.ldo2 + n0;
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
}

// NV consecutive columns starting at n0 (n0 % NV == 0); v holds the raw accumulators.
template <int NV, bool kFastGelu>
__device__ __forceinline__ void epi_store(const EpiParams& p, const EpiRow& r, int n0, float* v) {
    static_assert(NV == 4 || NV == 8, "NV");
    if (!r.valid || n0 >= p.N) return;
    if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
    }
    if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = kFastGelu ? gelu_erf_fast(v[i]) : gelu_erf(v[i]);
    }
    if (p.res != nullptr) {
        const float* rp = p.res + r.res_row * p.ldres + n0;
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(rp + i);
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
    }
    epi_write<NV>(p, r, n0, v);
}

inline void validate_gemm_common(const GemmArgs& a) {
    WB_REQUIRE(a.A && a.W && a.out, "null pointer");
    WB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM");
    WB_REQUIRE(a.N % 8 == 0, "N must be a multiple of 8");
    WB_REQUIRE(a.K % 16 == 0, "K must be a multiple of 16");
    WB_REQUIRE(a.ldo % 8 == 0 || a.out_mode == 1, "ldo must be a multiple of 8");
    WB_REQUIRE(a.res == nullptr || a.ldres % 4 == 0, "ldres must be a multiple of 4");
    if (a.m_period_in > 0) WB_REQUIRE(a.m_valid > 0 && a.m_valid <= a.m_period_in && a.m_period_out > 0, "bad row remap");
    if (a.out_mode == 1) WB_REQUIRE(a.hs_heads > 0 && a.m_period_in > 0 && a.N % (a.hs_heads * 64) == 0, "bad head split");
}

}  // namespace wb
