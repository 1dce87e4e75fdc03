// One-query attention for the greedy decode step: softmax(q K^T) V per (utterance, head), head_dim 64.
//
//  * cross attention: K/V were projected ONCE per utterance into [B, H, 1500, 64] and are only read here
//    (the reference re-copies the whole 1500-frame K/V block three times per step: identity-plugin memcpy
//    models/whisper/model.py:287-288, output concat :461-462, .clone() run.py:145-146).
//  * self attention: the new key/value row is appended IN PLACE into a paged cache (page = 64 tokens x all
//    heads) and the kernel attends over cur_len keys, replacing the oracle's torch.cat growth
//    (modeling_whisper.py:494-495) and TRT-LLM's slice + concat (model.py:276-281).
//  Oracle semantics (modeling_whisper.py:468-526): no mask, q pre-scaled (folded into the q weights), fp32 softmax.
//
// HBM-bound streaming kernel: every key row (128 B in bf16) is read by 8 consecutive lanes with one
// 16-byte load each, so a warp instruction covers 4 complete rows = 512 contiguous bytes.
#include "wb_internal.h"

namespace wb {

namespace {
constexpr int DH = 64, THREADS = 256, WARPS = THREADS / 32;

template <typename T>
__global__ void __launch_bounds__(THREADS) decode_attn_kernel(DecAttnArgs a) {
    constexpr int VEC = Vec16<T>::N;       // elements per 16-byte load: 8 (bf16) / 4 (fp32)
    constexpr int LPK = DH / VEC;          // lanes per key row: 8 / 16
    constexpr int KPW = 32 / LPK;          // key rows per warp instruction: 4 / 2
    constexpr int KPB = KPW * WARPS;       // key rows per block iteration
    extern __shared__ float sc[];          // [n] scores, then probabilities
    __shared__ float red[WARPS];
    __shared__ float ored[WARPS][DH];

    int n = a.n_keys;
    if (a.state != nullptr) {
        if (a.state->active == 0) return;
        n = a.state->cur_len;
    }
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane % LPK, grp = lane / LPK;
    const bool paged = a.k_pages != nullptr;

    float qf[VEC];
    ld16(reinterpret_cast<const T*>(a.q) + (size_t)b * a.q_stride + h * DH + sub * VEC).unpack(qf);

    const T* kbase = nullptr;
    const T* vbase = nullptr;
    const int* pt = nullptr;
    if (paged) {
        pt = a.page_table + (size_t)b * a.pages_per_seq;
        if (a.k_new != nullptr && warp == 0 && grp == 0) {  // in-place append at slot n-1
            const int s = n - 1;
            const size_t off = (((size_t)pt[s / a.page_tokens] * a.H + h) * a.page_tokens + (s % a.page_tokens)) * DH + sub * VEC;
            const size_t src = (size_t)b * a.new_stride + h * DH + sub * VEC;
            st16(reinterpret_cast<T*>(a.k_pages) + off, ld16(reinterpret_cast<const T*>(a.k_new) + src));
            st16(reinterpret_cast<T*>(a.v_pages) + off, ld16(reinterpret_cast<const T*>(a.v_new) + src));
        }
        __syncthreads();  // the appended row is read below by other warps of this block
    } else {
        const size_t o = (size_t)b * a.kv_bstride + (size_t)h * a.kv_hstride;
        kbase = reinterpret_cast<const T*>(a.k) + o;
        vbase = reinterpret_cast<const T*>(a.v) + o;
    }
    auto row_ptr = [&](const T* contiguous, const void* pages, int s) -> const T* {
        if (!paged) return contiguous + (size_t)s * DH + sub * VEC;
        return reinterpret_cast<const T*>(pages) +
               (((size_t)pt[s / a.page_tokens] * a.H + h) * a.page_tokens + (s % a.page_tokens)) * DH + sub * VEC;
    };

    // ---- phase 1: scores ----
    float lmax = -INFINITY;
#pragma unroll 4
    for (int s0 = warp * KPW; s0 < n; s0 += KPB) {
        const int s = s0 + grp;
        const bool valid = s < n;
        float kf[VEC];
        (paged ? ld16(row_ptr(kbase, a.k_pages, valid ? s : n - 1))
               : ld16_stream(row_ptr(kbase, a.k_pages, valid ? s : n - 1))).unpack(kf);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) dot = fmaf(qf[i], kf[i], dot);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (valid && sub == 0) sc[s] = dot;
        if (valid) lmax = fmaxf(lmax, dot);
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float gmax = red[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) gmax = fmaxf(gmax, red[w]);
    __syncthreads();  // red is reused below

    // ---- phase 2: exp + sum ----
    float lsum = 0.f;
    for (int s = tid; s < n; s += THREADS) {
        const float p = expf(sc[s] - gmax);
        sc[s] = p;
        lsum += p;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float gsum = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) gsum += red[w];

    // ---- phase 3: out = P V ----
    float of[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) of[i] = 0.f;
#pragma unroll 4
    for (int s0 = warp * KPW; s0 < n; s0 += KPB) {
        const int s = s0 + grp;
        const bool valid = s < n;
        float vf[VEC];
        (paged ? ld16(row_ptr(vbase, a.v_pages, valid ? s : n - 1))
               : ld16_stream(row_ptr(vbase, a.v_pages, valid ? s : n - 1))).unpack(vf);
        const float p = valid ? sc[s] : 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) of[i] = fmaf(p, vf[i], of[i]);
    }
    // reduce over the key groups of the warp (lanes with equal sub), then over warps
#pragma unroll
    for (int i = 0; i < VEC; ++i)
#pragma unroll
        for (int o = LPK; o < 32; o <<= 1) of[i] += __shfl_xor_sync(0xffffffffu, of[i], o);
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) ored[warp][sub * VEC + i] = of[i];
    }
    __syncthreads();
    if (tid < DH) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += ored[w][tid];
        reinterpret_cast<T*>(a.out)[(size_t)b * a.out_stride + h * DH + tid] = from_f32<T>(v / gsum);
    }
}
}  // namespace

void decode_attention(const DecAttnArgs& a, cudaStream_t stream) {
    WB_REQUIRE(a.q && a.out && a.B > 0 && a.H > 0, "bad decode attention arguments");
    const bool paged = a.k_pages != nullptr;
    WB_REQUIRE(paged || (a.k && a.v), "missing K/V");
    WB_REQUIRE(!paged || (a.page_table && a.v_pages && a.pages_per_seq > 0 && a.page_tokens > 0), "bad paged cache");
    const int max_keys = paged ? a.pages_per_seq * a.page_tokens : a.n_keys;
    WB_REQUIRE(max_keys > 0 && max_keys <= 8192, "key count out of range");
    dim3 grid(a.H, a.B), block(THREADS);
    const size_t smem = (size_t)max_keys * sizeof(float);
    if (a.dtype == F32) decode_attn_kernel<float><<<grid, block, smem, stream>>>(a);
    else decode_attn_kernel<bf16><<<grid, block, smem, stream>>>(a);
    WB_CHECK_LAUNCH();
}

}  // namespace wb
