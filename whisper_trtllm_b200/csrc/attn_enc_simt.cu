// Encoder self-attention, CUDA-core flash-style kernel with fp32 arithmetic throughout.
// Exactness path for the fp32 configs (and the fallback for bf16 when the tcgen05 kernel is disabled).
// Semantics: oracle WhisperEncoderAttention.forward (modeling_whisper.py:569-593): softmax(q k^T) v over all
// S = 1500 keys, no mask, q pre-scaled by head_dim^-0.5 (the scale is folded into the packed q weights).
// The S x S score matrix the reference materialises never leaves the SM.
#include "wb_internal.h"

namespace wb {

namespace {
constexpr int TQ = 64, TK = 64, DH = 64, PADW = 4;

struct SmemLayout {
    float Qt[DH][TQ + PADW];   // Qt[dd][row]
    float Kt[DH][TK + PADW];   // Kt[dd][key]
    float Vs[TK][DH + PADW];   // Vs[key][col]
    float Pt[TK][TQ + PADW];   // Pt[key][row]
};

template <typename T>
__global__ void __launch_bounds__(256) enc_attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int S, int H) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int d = H * DH, ld = 3 * d;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const T* base = qkv + (size_t)b * S * ld + h * DH;

    // load the Q tile transposed (rows beyond S are clamped; they are never stored)
    for (int i = tid; i < TQ * (DH / 4); i += 256) {
        const int r = i / (DH / 4), c4 = (i - r * (DH / 4)) * 4;
        const T* p = base + (size_t)min(q0 + r, S - 1) * ld + c4;
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.Qt[c4 + j][r] = to_f32(p[j]);
    }
    float m_run[4], l_run[4], acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < S; k0 += TK) {
        __syncthreads();  // previous tile fully consumed (also orders the Q stores before first use)
        for (int i = tid; i < TK * (DH / 4); i += 256) {
            const int r = i / (DH / 4), c4 = (i - r * (DH / 4)) * 4;
            const size_t row = (size_t)min(k0 + r, S - 1) * ld;
            const T* pk = base + row + d + c4;
            const T* pv = base + row + 2 * d + c4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sm.Kt[c4 + j][r] = to_f32(pk[j]);
                sm.Vs[r][c4 + j] = to_f32(pv[j]);
            }
        }
        __syncthreads();
        // scores: rows ty*4+i, keys tx*4+j
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int dd = 0; dd < DH; ++dd) {
            const float4 a = *reinterpret_cast<const float4*>(&sm.Qt[dd][ty * 4]);
            const float4 k = *reinterpret_cast<const float4*>(&sm.Kt[dd][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, kv[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], kv[j], s[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + tx * 4 + j >= S) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i][j] = -INFINITY;
            }
        // online softmax; the 16 threads sharing a row are 16 consecutive lanes
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m_run[i], mx);
            const float corr = expf(m_run[i] - m_new);  // exp(-inf) = 0 on the first tile
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - m_new);
                s[i][j] = p;
                ps += p;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            l_run[i] = l_run[i] * corr + ps;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] *= corr;
                sm.Pt[tx * 4 + j][ty * 4 + i] = s[i][j];
            }
        }
        __syncthreads();
        // O[rows ty*4+i][cols tx*4+j] += sum_key P[row][key] * V[key][col]
#pragma unroll 8
        for (int kk = 0; kk < TK; ++kk) {
            const float4 p = *reinterpret_cast<const float4*>(&sm.Pt[kk][ty * 4]);
            const float4 v = *reinterpret_cast<const float4*>(&sm.Vs[kk][tx * 4]);
            const float pv[4] = {p.x, p.y, p.z, p.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pv[i], vv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = q0 + ty * 4 + i;
        if (row < S) {
            const float inv = 1.0f / l_run[i];
            T* o = out + ((size_t)b * S + row) * d + h * DH + tx * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = from_f32<T>(acc[i][j] * inv);
        }
    }
}
}  // namespace

void encoder_attention_simt(const void* qkv, void* out, int dtype, int B, int S, int H, cudaStream_t stream) {
    dim3 grid(ceil_div(S, TQ), H, B), block(256);
    const size_t smem = sizeof(SmemLayout);
    if (dtype == F32) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enc_attn_simt_kernel<float><<<grid, block, smem, stream>>>((const float*)qkv, (float*)out, S, H);
    } else {
        WB_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enc_attn_simt_kernel<bf16><<<grid, block, smem, stream>>>((const bf16*)qkv, (bf16*)out, S, H);
    }
    WB_CHECK_LAUNCH();
}

static int g_attn_backend = 0;
void encoder_attention(const void* qkv, void* out, int dtype, int B, int S, int H, cudaStream_t stream) {
    WB_REQUIRE(qkv && out && B > 0 && S > 0 && H > 0, "bad encoder attention arguments");
    if (dtype == BF16 && g_attn_backend == 0) encoder_attention_tc(qkv, out, B, S, H, stream);
    else encoder_attention_simt(qkv, out, dtype, B, S, H, stream);
}
void set_attn_backend(int backend) { g_attn_backend = backend; }
int get_attn_backend() { return g_attn_backend; }

}  // namespace wb
