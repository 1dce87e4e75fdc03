// Persistent whole-step decoder kernel for small batches (B <= 16 utterances per GPU, bf16).
//
// At small batch the decode step is bound by kernel boundaries, not by bytes: 24 layers x 11 launches at ~8 us each is 14x above
// the weight-streaming floor (profiles/r01_decode_step_us.md).  Here ONE cooperative kernel (one 256-thread CTA per SM) runs the
// whole decoder for one token: 24 x {LN1+qkv -> paged self-attention -> out-proj(+res) -> LN2+cross-q -> cross-attention ->
// cross-out(+res) -> LN3+fc1(GELU) -> fc2(+res)}, final LN + LM head = 193 phases.  The embedding of the token that is fed is
// written into the residual stream by the greedy kernel of the step that chose it (greedy.cu), so a step is two launches.
// The phases are separated by a grid barrier (one release-add + acquire-poll in L2) instead of a kernel boundary, and
//   * the kernel is an interpreter over a table of phase descriptors resolved once per session on the host (PhaseDesc): one
//     copy of the staging code, of each tile geometry and of each attention flavour - the step has to stay small for the
//     instruction cache, and every instruction per warp on the critical path counts (8 warps per SM, latency bound);
//   * between arriving at the barrier and waiting on it, every warp loads the first weight tile of the NEXT linear phase into
//     registers and requests its LayerNorm parameters / bias / first attention unit into L2: weight latency is hidden behind
//     the barrier and the staging of the activations;
//   * linear layers run "swap-AB" on mma.sync m16n8k16: the 16-row MMA dimension walks the WEIGHT rows (streamed straight from
//     global memory into the A fragment with 16-byte loads, K permuted consistently in both operands), the 8-column dimension
//     holds the (<= 8 / <= 16) utterances staged once per CTA as bf16 in shared memory; fp32 accumulation;
//     K is split over the warps of a CTA and reduced through shared memory in a fixed order (deterministic);
//   * attention items are split over CTAs along the keys (cross: 1500 frames / split) so that a single utterance still uses
//     every SM; partials (max, sum, acc[64]) are merged in a fixed order by the last CTA to arrive at the item's counter.
// Semantics and rounding points are those of the multi-kernel step (runtime.cu decode_step_large):
// LayerNorm eps 1e-5 in fp32 (layers/normalization.py:6-30), bf16 activations into every Linear (layers/linear.py:38-139),
// fp32 residual stream, erf GELU, q pre-scaled in the packed weights, fp32 softmax, no mask (model.py:240-304),
// position = cur_len - 1 (model.py:423-425), tied LM head without bias (modeling_whisper.py:1335,1433).
// (tcgen05 needs M = 128 row tiles and a TMEM round trip per phase: at <= 16 rows mma.sync from registers is the right tool.)
// Measurements of every build of this file: profiles/r01_kernel_variants.md; per-phase clock stamps: tools/step_trace.py.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "wb_runtime.h"

namespace wb {

namespace {
constexpr int MG_THREADS = 256, MG_WARPS = 8;
constexpr int MG_MAX_LAYERS = 64;     // (any depth works: the phase table lives in global memory; this only bounds the trace buffer)
constexpr int MG_MAX_SPLITS = 16;
constexpr int MG_PART = 72;          // floats per attention partial: [0] max, [1] sum, [8..72) acc
constexpr int MG_RS = 20;            // reduction buffer: floats between consecutive utterance rows (bank-conflict free)
constexpr int MG_RED_BYTES = 2 * MG_WARPS * 16 * MG_RS * 4;   // two parities x 8 warps x 16 rows
constexpr int MG_ACT_PAD = 32;       // bf16 elements of padding per staged row (row stride = 64 mod 128 bytes)
constexpr unsigned MG_SPIN_LIMIT = 1u << 26;
constexpr int MG_ATT_UNROLL = 8;
// linear layers of a decoder layer: tiles of 8 weight rows, K split over the 8 warps of the CTA (one tile per CTA and round);
// wide ones (qkv, fc1 at d = 1024: more than two such rounds): tiles of 16 rows, K split over 4 warps, two tiles per CTA and
// round - every round costs a block barrier and an epilogue; LM head (51864 rows): tiles of 16 rows, one per warp, no K split
constexpr int MG_HEAD_KS = 1, MG_LAYER_KS = 8, MG_WIDE_KS = 4;

enum { PH_LINEAR = 0, PH_SELF_ATTN = 1, PH_CROSS_ATTN = 2, PH_HEAD = 3, PH_LINEAR_WIDE = 4 };
enum { ST_LN_X = 0, ST_COPY = 2 };
enum { EP_QKV = 0, EP_RESIDUAL = 1, EP_BF16 = 2, EP_GELU_BF16 = 3, EP_F32 = 4 };

// One phase of the step.  The table of all 8 * layers + 1 phases is built once per session on the host (everything that needs
// a division or a pointer chase is resolved there) and lives in global memory; the kernel is an interpreter over it: warp 0
// copies the NEXT descriptor into shared memory while the CTA waits at the grid barrier.  The phases are latency-bound
// instruction streams (8 warps per SM, ~5 cycles per dependent instruction): the instruction count per warp on the critical
// path is what counts, and the whole interpreter has to stay small for the instruction cache (profiles/r01_step_trace_*.md).
struct alignas(16) PhaseDesc {
    const bf16* W; const float* bias; const float* gamma; const float* beta;   // linear: weights [N, K], bias, LayerNorm in front
    const bf16* src; bf16* out_bf16; float* out_f32;                            // rows to copy when there is no LayerNorm; outputs
    bf16* k_pages; bf16* v_pages;      // qkv epilogue + self-attention: this layer's pages; cross-attention: this layer's K / V
    int kind, N, K, stage, epi;
    int nsc;                           // super-chunks (8 x 16-byte loads per lane) per tile
    int rounds_base, rounds_rem;       // rounds of CTA c = rounds_base + (c * tiles_per_round < rounds_rem)
    unsigned vpr_rcp;                  // ceil(2^32 / (K / 8)): row of a 16-byte vector without a division
    int pad[5];
};
static_assert(sizeof(PhaseDesc) == 128, "PhaseDesc is copied as 16 x 8 bytes");

struct MegaParams {
    int M, d, H, ffn, vocab, n_phases, n_ctx, cross_splits, pages_per_seq;
    long long cross_bstride;
    const StepState* state; const int* unfinished; const int* page_table;
    float* x; bf16 *q, *ctx; float* part;
    const PhaseDesc* table;
    long long* trace;  // optional (tools/step_trace.py): SM clock stamps of CTA 0, 8 slots per phase (see the kernel body)
    unsigned* sync;    // [0] grid barrier counter, [32 ..) per-item arrival counters; all zero between launches
};

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {   // read-only for the whole kernel (weights, cross K/V)
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_cg16(const void* p) {   // written by other CTAs of this kernel: L2 is the point of coherence
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float4 ldg_cg_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ldg_cg_f(const float* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// D(16 x 8, fp32) += A(16 x 16, bf16, row) * B(16 x 8, bf16, col)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// All CTAs of the (co-resident, cooperative) grid meet here.  Writes made before the barrier by any thread of any CTA are
// visible to every thread after it: bar.sync orders the CTA's writes before thread 0's gpu-scope release, the acquire poll
// plus bar.sync orders the other CTAs' writes before this CTA's later reads.  A lost CTA traps instead of hanging the GPU.
// Split in two so that the requests for the next phase's weights are issued between the arrival and the wait.
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
    __syncthreads();
    if (threadIdx.x == 0) red_release_gpu(counter, 1u);
}
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned& epoch) {
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_gpu(counter) < epoch) {
            if (++spins > MG_SPIN_LIMIT) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- activation staging: the (normalised) rows of all M utterances as bf16 in shared memory, row stride K + MG_ACT_PAD
// Rows come from the fp32 residual stream; the embedding of the token that is fed (x = E[token] + P[position], model.py:423-425)
// was written there by the previous step's greedy kernel (greedy.cu, GreedyArgs::embed_x) or by decode_begin.
// The phases are latency-bound instruction streams: what counts is the number of dependent instructions per warp.
//  * M <= 2: a row is split over the whole CTA (thread t owns float4 t, d <= 1024): 3 loads per thread in one round trip, two
//    block reductions, ~60 instructions per row and warp;
//  * M > 2: one warp per row (rows warp, warp + 8), the 8 x-loads of a lane issued together, gamma / beta (requested into L2
//    before the grid barrier) in two half-batches: no block barrier, the rows run in parallel on the 8 warps.
// (Measured alternatives, profiles/r01_step_trace_*.md: guarded per-element loop - ptxas serialised the gamma / beta loads
// into 8 dependent round trips, 6.4 K cycles; every thread computing all 8 / 16 rows - 4.2 K / 17 K cycles.)
__device__ __forceinline__ void stage_layernorm_wide(const MegaParams& p, bf16* act_s, float* red_s, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta) {
    constexpr int MR = 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = p.d, nvec = d >> 2, astride = d + MG_ACT_PAD, M = p.M;
    const bool on = tid < nvec;
    const float inv_d = 1.0f / (float)d;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4, v[MR];
    if (on) {
        g4 = __ldg(reinterpret_cast<const float4*>(gamma) + tid);
        b4 = __ldg(reinterpret_cast<const float4*>(beta) + tid);
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        v[m] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on && m < M) v[m] = ldg_cg_f4(p.x + (size_t)m * d + tid * 4);
    }
    float* rs = red_s;                  // [MR][8] partial sums, then [MR][8] partial sums of squares
    float* rq = red_s + MR * MG_WARPS;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        const float s = warp_sum((v[m].x + v[m].y) + (v[m].z + v[m].w));     // threads past the row hold zeros
        if (lane == 0) rs[m * MG_WARPS + warp] = s;
    }
    __syncthreads();
    float mean[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        const float4 a = *reinterpret_cast<const float4*>(rs + m * MG_WARPS), b = *reinterpret_cast<const float4*>(rs + m * MG_WARPS + 4);
        mean[m] = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * inv_d;
        const float c0 = v[m].x - mean[m], c1 = v[m].y - mean[m], c2 = v[m].z - mean[m], c3 = v[m].w - mean[m];
        const float ss = warp_sum(on ? (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3) : 0.f);
        if (lane == 0) rq[m * MG_WARPS + warp] = ss;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        const float4 a = *reinterpret_cast<const float4*>(rq + m * MG_WARPS), b = *reinterpret_cast<const float4*>(rq + m * MG_WARPS + 4);
        const float rstd = rsqrtf((((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * inv_d + 1e-5f);
        if (on && m < M) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn((v[m].x - mean[m]) * rstd * g4.x + b4.x, (v[m].y - mean[m]) * rstd * g4.y + b4.y);
            __nv_bfloat162 p1 = __floats2bfloat162_rn((v[m].z - mean[m]) * rstd * g4.z + b4.z, (v[m].w - mean[m]) * rstd * g4.w + b4.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(act_s + (size_t)m * astride + tid * 4) = u;
        }
    }
}

__device__ __forceinline__ void stage_layernorm_rows(const MegaParams& p, bf16* act_s, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.d, nvec = d >> 2, astride = d + MG_ACT_PAD;
    const float inv_d = 1.0f / (float)d;
    for (int m = warp; m < p.M; m += MG_WARPS) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = lane + 32 * i;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < nvec) v[i] = ldg_cg_f4(p.x + (size_t)m * d + idx * 4);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);      // lanes past the row hold zeros
        const float mean = warp_sum(s) * inv_d;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (lane + 32 * i < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                ss += (a * a + b * b) + (c * c + e * e);
            }
        }
        const float rstd = rsqrtf(warp_sum(ss) * inv_d + 1e-5f);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 g4[4], b4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * (half * 4 + j);
                g4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                b4[j] = g4[j];
                if (idx < nvec) {
                    g4[j] = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
                    b4[j] = __ldg(reinterpret_cast<const float4*>(beta) + idx);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = half * 4 + j, idx = lane + 32 * i;
                if (idx < nvec) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn((v[i].x - mean) * rstd * g4[j].x + b4[j].x, (v[i].y - mean) * rstd * g4[j].y + b4[j].y);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn((v[i].z - mean) * rstd * g4[j].z + b4[j].z, (v[i].w - mean) * rstd * g4[j].w + b4[j].w);
                    uint2 u;
                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(act_s + (size_t)m * astride + idx * 4) = u;
                }
            }
        }
    }
}

// bf16 rows of the previous phase -> shared memory; the requests of a thread are issued together (no clamped duplicates:
// 148 SMs x 256 threads asking for one line serialise in its L2 slice)
__device__ __forceinline__ void stage_copy(bf16* act_s, const bf16* src, int M, int K, unsigned vpr_rcp) {
    const int vpr = K >> 3, astride = K + MG_ACT_PAD, total = M * vpr;
    for (int base = 0; base < total; base += MG_THREADS * 8) {
        uint4 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MG_THREADS + (int)threadIdx.x;
            r[u] = make_uint4(0, 0, 0, 0);
            if (i < total) r[u] = ldg_cg16(src + (size_t)i * 8);      // rows are contiguous: element offset = i * 8
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MG_THREADS + (int)threadIdx.x;
            if (i < total) {
                const int m = (int)__umulhi((unsigned)i, vpr_rcp), j = i - m * vpr;   // vpr_rcp = ceil(2^32 / vpr), exact for i < 2^16
                *reinterpret_cast<uint4*>(act_s + (size_t)m * astride + j * 8) = r[u];
            }
        }
    }
}

// called exactly once per output element (n, m) by exactly one thread of the grid
// kPre: pre_b / pre_x already hold bias[n] / x[m, n] (residual epilogue), loaded while the weights were in flight
template <bool kPre>
__device__ __forceinline__ void linear_epilogue(const MegaParams& p, const PhaseDesc& D, int pos, int n, int m, float v, float pre_b,
                                                float pre_x) {
    if (n >= D.N || m >= p.M) return;
    if (kPre) v += pre_b;
    else if (D.bias != nullptr) v += __ldg(D.bias + n);
    const int d = p.d;
    switch (D.epi) {
        case EP_QKV: {   // q -> activation buffer; k / v rows straight into the paged cache at slot cur_len - 1
            const int which = (n >= d ? 1 : 0) + (n >= 2 * d ? 1 : 0), c = n - which * d;
            if (which == 0) {
                D.out_bf16[(size_t)m * d + c] = __float2bfloat16_rn(v);
            } else {
                const int page = p.page_table[(size_t)m * p.pages_per_seq + (pos >> 6)];
                const size_t off = (((size_t)page * p.H + (c >> 6)) * 64 + (pos & 63)) * 64 + (c & 63);
                (which == 1 ? D.k_pages : D.v_pages)[off] = __float2bfloat16_rn(v);
            }
            break;
        }
        case EP_RESIDUAL: {   // fp32 residual stream, updated in place: this thread owns (m, n)
            float* xp = p.x + (size_t)m * d + n;
            *xp = (kPre ? pre_x : ldg_cg_f(xp)) + v;
            break;
        }
        case EP_BF16: D.out_bf16[(size_t)m * D.N + n] = __float2bfloat16_rn(v); break;
        case EP_GELU_BF16: D.out_bf16[(size_t)m * D.N + n] = __float2bfloat16_rn(gelu_erf_fast(v)); break;
        default: D.out_f32[(size_t)m * D.N + n] = v; break;
    }
}

// out[m, n] = epilogue( sum_k act[m, k] * W[n, k] ) for all n < N, m < 8 * NM; the activations are already staged and the
// first weight tile is already in registers.
// KS warps split K for one tile of RT = 8 / 16 weight rows; a CTA works on 8 / KS tiles per round; tiles (r * grid + cta) * TPR + ..
template <int NM, int KS, bool R16>
__device__ __forceinline__ void linear_phase(const MegaParams& p, const PhaseDesc& D, int pos, uint8_t* smem, long long* tr,
                                             const uint4 (&first)[8]) {
    constexpr int TPR = MG_WARPS / KS, RT_SHIFT = R16 ? 4 : 3, RT = 1 << RT_SHIFT, CPS = R16 ? 4 : 8;
    float* red_s = reinterpret_cast<float*>(smem);
    bf16* act_s = reinterpret_cast<bf16*>(smem + MG_RED_BYTES);
    const bf16* __restrict__ W = D.W;
    const int N = D.N, K = D.K, C = K >> 5, astride = K + MG_ACT_PAD, nsc = D.nsc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int tl = warp / KS, ks_id = warp % KS;
    const int c0 = (ks_id * C) / KS, c1 = ((ks_id + 1) * C) / KS;
    const int rounds = D.rounds_base + ((int)blockIdx.x * TPR < D.rounds_rem ? 1 : 0);
    const int total = rounds * nsc;
    auto tile_of = [&](int r, int t_local) { return (r * (int)gridDim.x + (int)blockIdx.x) * TPR + t_local; };

    // (r, sc) = (round, super-chunk) of the tile being requested / computed; advanced incrementally, no divisions
    auto issue = [&](int r, int sc, uint4(&buf)[8]) __attribute__((always_inline)) {
        const int n0 = tile_of(r, tl) << RT_SHIFT;
        const bf16* pa = W + (size_t)min(n0 + g, N - 1) * K + tq * 8;
        if constexpr (R16) {
            const bf16* pb = W + (size_t)min(n0 + g + 8, N - 1) * K + tq * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = c0 + sc * 4 + j;
                buf[2 * j] = make_uint4(0, 0, 0, 0);
                buf[2 * j + 1] = make_uint4(0, 0, 0, 0);
                if (chunk < c1) {
                    buf[2 * j] = ldg_nc16(pa + chunk * 32);
                    buf[2 * j + 1] = ldg_nc16(pb + chunk * 32);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int chunk = c0 + sc * 8 + j;
                buf[j] = make_uint4(0, 0, 0, 0);
                if (chunk < c1) buf[j] = ldg_nc16(pa + chunk * 32);
            }
        }
    };

    // K-split tiles: thread o < OUTS owns output element (n_l, m) of tile t2 in every round; its bias (and the residual it adds
    // to) are loaded when the round's weights are requested, not after the reduction
    constexpr int M_CNT = 8 * NM, OUTS = KS == 1 ? 0 : TPR * M_CNT * RT, OPT = OUTS == 0 ? 1 : (OUTS + MG_THREADS - 1) / MG_THREADS;
    // output element e of this thread: o = tid + e * 256 -> (n_l, m, t2)
    auto out_nl = [&](int e) { return (tid + e * MG_THREADS) & (RT - 1); };
    auto out_m = [&](int e) { return ((tid + e * MG_THREADS) >> RT_SHIFT) % M_CNT; };
    auto out_t2 = [&](int e) { return ((tid + e * MG_THREADS) >> RT_SHIFT) / M_CNT; };
    float2 pre_cur[OPT], pre_nxt[OPT];     // (bias, residual); no arithmetic until the epilogue
#pragma unroll
    for (int e = 0; e < OPT; ++e) pre_cur[e] = pre_nxt[e] = make_float2(0.f, 0.f);
    auto preload = [&](int r, float2(&dst)[OPT]) __attribute__((always_inline)) {
#pragma unroll
        for (int e = 0; e < OPT; ++e) {
            dst[e] = make_float2(0.f, 0.f);
            if (tid + e * MG_THREADS < OUTS) {
                const int n = (tile_of(r, out_t2(e)) << RT_SHIFT) + out_nl(e), m = out_m(e);
                if (n < N && m < p.M) {
                    if (D.bias != nullptr) dst[e].x = __ldg(D.bias + n);
                    if (D.epi == EP_RESIDUAL) dst[e].y = ldg_cg_f(p.x + (size_t)m * p.d + n);
                }
            }
        }
    };

    float acc[NM][4];
    auto finish = [&](int r) __attribute__((always_inline)) {
        if constexpr (KS == 1) {
            const int n0 = tile_of(r, tl) << RT_SHIFT;
#pragma unroll
            for (int mb = 0; mb < NM; ++mb) {
#pragma unroll 1
                for (int e = 0; e < (R16 ? 4 : 2); ++e) {   // rolled: one copy of the epilogue code
                    const float v = e == 0 ? acc[mb][0] : e == 1 ? acc[mb][1] : e == 2 ? acc[mb][2] : acc[mb][3];
                    linear_epilogue<false>(p, D, pos, n0 + g + (e >> 1) * 8, mb * 8 + 2 * tq + (e & 1), v, 0.f, 0.f);
                }
            }
        } else {
            constexpr int WSTRIDE = 16 * MG_RS;   // floats per warp slab: 16 utterance rows x MG_RS
            float* rw = red_s + ((r & 1) * MG_WARPS + warp) * WSTRIDE;
#pragma unroll
            for (int mb = 0; mb < NM; ++mb) {
                rw[(mb * 8 + 2 * tq) * MG_RS + g] = acc[mb][0];
                rw[(mb * 8 + 2 * tq + 1) * MG_RS + g] = acc[mb][1];
                if constexpr (R16) {
                    rw[(mb * 8 + 2 * tq) * MG_RS + g + 8] = acc[mb][2];
                    rw[(mb * 8 + 2 * tq + 1) * MG_RS + g + 8] = acc[mb][3];
                }
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < OPT; ++e) {
                if (tid + e * MG_THREADS < OUTS) {
                    const int t2 = out_t2(e), m = out_m(e), n_l = out_nl(e);
                    const float* rr = red_s + ((r & 1) * MG_WARPS + t2 * KS) * WSTRIDE + m * MG_RS + n_l;
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < KS; ++k) v += rr[k * WSTRIDE];     // fixed order: deterministic
                    linear_epilogue<true>(p, D, pos, (tile_of(r, t2) << RT_SHIFT) + n_l, m, v, pre_cur[e].x, pre_cur[e].y);
                }
            }
        }
    };
    auto compute = [&](int r, int sc, const uint4(&buf)[8]) __attribute__((always_inline)) {
        if (sc == 0) {
#pragma unroll
            for (int mb = 0; mb < NM; ++mb)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mb][i] = 0.f;
        }
        const bf16* arow = act_s + (size_t)g * astride + tq * 8;
        // K is permuted identically in both operands: lane tq supplies elements [8 tq, 8 tq + 8) of every 32-element chunk
#pragma unroll
        for (int j = 0; j < CPS; ++j) {
            const int chunk = c0 + sc * CPS + j;
            if (chunk < c1) {   // warp-uniform
#pragma unroll
                for (int mb = 0; mb < NM; ++mb) {
                    const uint4 a = *reinterpret_cast<const uint4*>(arow + (size_t)mb * 8 * astride + chunk * 32);
                    if constexpr (R16) {
                        mma_16816(acc[mb], buf[2 * j].x, buf[2 * j + 1].x, buf[2 * j].y, buf[2 * j + 1].y, a.x, a.y);
                        mma_16816(acc[mb], buf[2 * j].z, buf[2 * j + 1].z, buf[2 * j].w, buf[2 * j + 1].w, a.z, a.w);
                    } else {       // 8-row tiles: MMA rows 8..15 are zero
                        mma_16816(acc[mb], buf[j].x, 0u, buf[j].y, 0u, a.x, a.y);
                        mma_16816(acc[mb], buf[j].z, 0u, buf[j].w, 0u, a.z, a.w);
                    }
                }
            }
        }
        if (sc == nsc - 1) finish(r);
    };

    // `first` = super-chunk 0 of round 0, requested before the grid barrier (prefetch_linear): it has been in flight during the
    // wait and the staging
    uint4 cur[8], nxt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = first[i];
    if (total > 0) {
        if constexpr (KS > 1) preload(0, pre_cur);
    }
    if (tr != nullptr) tr[3] = clock64();
    int r = 0, sc = 0;
#pragma unroll 1
    for (int idx = 0; idx < total; ++idx) {   // trip counts are CTA-uniform (finish() contains a block barrier)
        int r1 = r, sc1 = sc + 1;
        if (sc1 == nsc) { sc1 = 0; ++r1; }
        if (idx + 1 < total) {
            issue(r1, sc1, nxt);
            if constexpr (KS > 1) { if (r1 != r) preload(r1, pre_nxt); }
        }
        compute(r, sc, cur);
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        if (r1 != r) {
#pragma unroll
            for (int e = 0; e < OPT; ++e) pre_cur[e] = pre_nxt[e];
        }
        r = r1;
        sc = sc1;
    }
}

// ---- one-query attention, one CTA per (item, key split) unit; 8 lanes per key row, 4 rows per warp instruction
template <bool kTC>
__device__ __forceinline__ void attention_phase(const MegaParams& p, const bool kPaged, const bf16* __restrict__ kbase, const bf16* __restrict__ vbase,
                                                int n_keys, int splits, uint8_t* smem, unsigned* item_cnt) {
    float* pm = reinterpret_cast<float*>(smem);
    float* pl = pm + MG_WARPS;
    float* po = pl + MG_WARPS;                 // [8][64]
    int* flag = reinterpret_cast<int*>(po + MG_WARPS * 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & 7, grp = lane >> 3;
    const int H = p.H, d = p.d, n_items = p.M * H, units = n_items * splits;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int item = unit / splits, sp = unit - item * splits;
        const int b = item / H, h = item - b * H;
        if (p.unfinished[b] == 0) continue;    // CTA-uniform; every split of the item skips, the counter stays 0
        const int s_beg = (int)((long long)sp * n_keys / splits), s_end = (int)((long long)(sp + 1) * n_keys / splits);
        if constexpr (kTC) {
            // ---- tensor-core variant: 16 keys per mma.sync block, 2 blocks (32 keys) per warp iteration.
            //  scores = K q:  A = K block (rows = keys; lane (g, tq) loads dims [8 tq, 8 tq + 8) and [32 + 8 tq, ..) of rows g and g + 8
            //                 with two 16-byte requests each, K permuted identically in q), B = q in column 0 (lanes g == 0);
            //  out += p V:    A = p in row 0 (lanes g == 0; keys 4 tq .. 4 tq + 3, bf16), B = V block: lane (g, tq) loads dims
            //                 [8 g, 8 g + 8) of keys 4 tq .. 4 tq + 3 (16 bytes each) and interleaves key pairs with PRMT; MMA i covers
            //                 dims 8 n + i of column n, so lane (0, tq) ends with out[16 tq + i] and out[16 tq + 8 + i].
            // ~7 warp instructions per key instead of ~19 for the 8-lanes-per-key dot products below - and measured slower (see
            // mega_attention_tc): the default is the CUDA-core flavour.
            const int g = lane >> 2, tq = lane & 3;
            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
            if (g == 0) {
                const bf16* qp = p.q + (size_t)b * d + h * 64;
                q0 = ldg_cg16(qp + 8 * tq);
                q1 = ldg_cg16(qp + 32 + 8 * tq);
            }
            int my_page = 0;
            if (kPaged) my_page = lane < p.pages_per_seq ? p.page_table[(size_t)b * p.pages_per_seq + lane] : 0;
            auto row_base = [&](int s) -> size_t {
                if (kPaged) {   // CTA-uniform
                    const int page = __shfl_sync(0xffffffffu, my_page, s >> 6);
                    return (((size_t)page * H + h) * 64 + (s & 63)) * 64;
                } else {
                    return (size_t)b * p.cross_bstride + ((size_t)h * n_keys + s) * 64;
                }
            };
            float m_run = -INFINITY, l_part = 0.f, acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
            for (int blk = s_beg + warp * 32; blk < s_end; blk += MG_WARPS * 32) {   // warp-uniform trip count
                uint4 ka[2][2], kb[2][2], vv[2][4];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int base = blk + h2 * 16;
                    const size_t oa = row_base(min(base + g, s_end - 1)) + 8 * tq;
                    const size_t ob = row_base(min(base + g + 8, s_end - 1)) + 8 * tq;
                    ka[h2][0] = ldg_cg16(kbase + oa);
                    ka[h2][1] = ldg_cg16(kbase + oa + 32);
                    kb[h2][0] = ldg_cg16(kbase + ob);
                    kb[h2][1] = ldg_cg16(kbase + ob + 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) vv[h2][j] = ldg_cg16(vbase + row_base(min(base + 4 * tq + j, s_end - 1)) + 8 * g);
                }
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int base = blk + h2 * 16;
                    if (base < s_end) {   // warp-uniform
                        float c[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_16816(c, ka[h2][0].x, kb[h2][0].x, ka[h2][0].y, kb[h2][0].y, q0.x, q0.y);
                        mma_16816(c, ka[h2][0].z, kb[h2][0].z, ka[h2][0].w, kb[h2][0].w, q0.z, q0.w);
                        mma_16816(c, ka[h2][1].x, kb[h2][1].x, ka[h2][1].y, kb[h2][1].y, q1.x, q1.y);
                        mma_16816(c, ka[h2][1].z, kb[h2][1].z, ka[h2][1].w, kb[h2][1].w, q1.z, q1.w);
                        // column 0 lives in the lanes tq == 0: scores of keys g (c[0]) and g + 8 (c[2])
                        float s0 = __shfl_sync(0xffffffffu, c[0], lane & ~3), s1 = __shfl_sync(0xffffffffu, c[2], lane & ~3);
                        if (base + g >= s_end) s0 = -INFINITY;
                        if (base + g + 8 >= s_end) s1 = -INFINITY;
                        float mb = fmaxf(s0, s1);
                        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 4));
                        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 8));
                        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 16));
                        const float m_new = fmaxf(m_run, mb);          // finite: key `base` is valid
                        const float scale = __expf(m_run - m_new);     // exp(-inf) = 0 while l / acc are still 0
                        l_part *= scale;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { acc[i][0] *= scale; acc[i][1] *= scale; }
                        m_run = m_new;
                        // lane (0, tq) needs p of keys 4 tq .. 4 tq + 3: key k sits in s0 (k < 8) / s1 of the lanes with g = k & 7
                        const float mine = tq >= 2 ? s1 : s0;
                        float pj[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float v = __shfl_sync(0xffffffffu, mine, ((((tq & 1) << 2) + j) << 2) + tq);
                            pj[j] = (g == 0 && v != -INFINITY) ? __expf(v - m_new) : 0.f;
                        }
                        const __nv_bfloat162 p01 = __floats2bfloat162_rn(pj[0], pj[1]), p23 = __floats2bfloat162_rn(pj[2], pj[3]);
                        // the sum uses the rounded probabilities, like the products below
                        l_part += (__low2float(p01) + __high2float(p01)) + (__low2float(p23) + __high2float(p23));
                        const uint32_t pa = *reinterpret_cast<const uint32_t*>(&p01), pb = *reinterpret_cast<const uint32_t*>(&p23);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
                            const uint32_t r0 = (i >> 1) == 0 ? vv[h2][0].x : (i >> 1) == 1 ? vv[h2][0].y : (i >> 1) == 2 ? vv[h2][0].z : vv[h2][0].w;
                            const uint32_t r1 = (i >> 1) == 0 ? vv[h2][1].x : (i >> 1) == 1 ? vv[h2][1].y : (i >> 1) == 2 ? vv[h2][1].z : vv[h2][1].w;
                            const uint32_t r2 = (i >> 1) == 0 ? vv[h2][2].x : (i >> 1) == 1 ? vv[h2][2].y : (i >> 1) == 2 ? vv[h2][2].z : vv[h2][2].w;
                            const uint32_t r3 = (i >> 1) == 0 ? vv[h2][3].x : (i >> 1) == 1 ? vv[h2][3].y : (i >> 1) == 2 ? vv[h2][3].z : vv[h2][3].w;
                            mma_16816(acc[i], pa, 0u, pb, 0u, prmt(r0, r1, sel), prmt(r2, r3, sel));
                        }
                    }
                }
            }
            float l_run = l_part;
            l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
            l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
            if (g == 0) {
                if (tq == 0) { pm[warp] = m_run; pl[warp] = l_run; }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    po[warp * 64 + 16 * tq + i] = acc[i][0];
                    po[warp * 64 + 16 * tq + 8 + i] = acc[i][1];
                }
            }
        } else {
            float qf[8];
            unpack8(ldg_cg16(p.q + (size_t)b * d + h * 64 + sub * 8), qf);
            int my_page = 0;
            if (kPaged) my_page = lane < p.pages_per_seq ? p.page_table[(size_t)b * p.pages_per_seq + lane] : 0;
            auto row_off = [&](int s) -> size_t {
                if (kPaged) {   // CTA-uniform
                    const int page = __shfl_sync(0xffffffffu, my_page, s >> 6);
                    return (((size_t)page * H + h) * 64 + (s & 63)) * 64 + sub * 8;
                } else {
                    return (size_t)b * p.cross_bstride + ((size_t)h * n_keys + s) * 64 + sub * 8;
                }
            };
            float m_run = -INFINITY, l_run = 0.f, acc[8];
    #pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            for (int sb = s_beg + warp * 4; sb < s_end; sb += 32 * MG_ATT_UNROLL) {   // warp-uniform trip count
                uint4 kr[MG_ATT_UNROLL], vr[MG_ATT_UNROLL];
    #pragma unroll
                for (int u = 0; u < MG_ATT_UNROLL; ++u) {
                    const size_t off = row_off(min(sb + grp + u * 32, s_end - 1));
                    kr[u] = ldg_cg16(kbase + off);          // L2-coherent: the newest self-attention row was written by another CTA
                    vr[u] = ldg_cg16(vbase + off);          // in the previous phase (cross K/V are read-only; same L1 bypass)
                }
                float sc[MG_ATT_UNROLL], mb = -INFINITY;
    #pragma unroll
                for (int u = 0; u < MG_ATT_UNROLL; ++u) {
                    float kf[8];
                    unpack8(kr[u], kf);
                    float dot = 0.f;
    #pragma unroll
                    for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[i], dot);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    sc[u] = (sb + grp + u * 32 < s_end) ? dot : -INFINITY;
                    mb = fmaxf(mb, sc[u]);
                }
                const float m_new = fmaxf(m_run, mb);
                const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
                const float scale = __expf(m_run - m_use);
                l_run *= scale;
    #pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] *= scale;
    #pragma unroll
                for (int u = 0; u < MG_ATT_UNROLL; ++u) {
                    const float pr = __expf(sc[u] - m_use);
                    l_run += pr;
                    float vf[8];
                    unpack8(vr[u], vf);
    #pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, vf[i], acc[i]);
                }
                m_run = m_new;
            }
    #pragma unroll
            for (int o = 8; o < 32; o <<= 1) {      // merge the four key groups of the warp
                const float om = __shfl_xor_sync(0xffffffffu, m_run, o);
                const float ol = __shfl_xor_sync(0xffffffffu, l_run, o);
                const float mm = fmaxf(m_run, om);
                const float s1 = (m_run == -INFINITY) ? 0.f : __expf(m_run - mm);
                const float s2 = (om == -INFINITY) ? 0.f : __expf(om - mm);
                l_run = l_run * s1 + ol * s2;
    #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
                    acc[i] = acc[i] * s1 + oa * s2;
                }
                m_run = mm;
            }
            if (grp == 0) {
                if (sub == 0) { pm[warp] = m_run; pl[warp] = l_run; }
    #pragma unroll
                for (int i = 0; i < 8; ++i) po[warp * 64 + sub * 8 + i] = acc[i];
            }
        }
        __syncthreads();
        if (tid < 64) {                         // merge the warps (fixed order)
            float mm = pm[0];
#pragma unroll
            for (int w = 1; w < MG_WARPS; ++w) mm = fmaxf(mm, pm[w]);
            float l = 0.f, o = 0.f;
#pragma unroll
            for (int w = 0; w < MG_WARPS; ++w) {
                const float s = (pm[w] == -INFINITY) ? 0.f : __expf(pm[w] - mm);
                l = fmaf(pl[w], s, l);
                o = fmaf(po[w * 64 + tid], s, o);
            }
            if (splits == 1) {
                p.ctx[(size_t)b * d + h * 64 + tid] = __float2bfloat16_rn(o / l);
            } else {
                float* pp = p.part + (size_t)(item * splits + sp) * MG_PART;
                pp[8 + tid] = o;
                if (tid == 0) { pp[0] = mm; pp[1] = l; }
            }
        }
        if (splits > 1) {                       // the last split to arrive merges the item's partials in split order
            __threadfence();
            __syncthreads();
            if (tid == 0) *flag = (atomicAdd(&item_cnt[item], 1u) == (unsigned)(splits - 1));
            __syncthreads();
            if (*flag) {                        // CTA-uniform
                __threadfence();
                // all partials of the item in ONE round trip: 256 threads x <= 5 independent loads -> shared memory
                float* pbuf = reinterpret_cast<float*>(smem) + 1024;
                const float* pp = p.part + (size_t)item * splits * MG_PART;
                const int n_f = splits * MG_PART;
                float r[(MG_MAX_SPLITS * MG_PART + MG_THREADS - 1) / MG_THREADS];
#pragma unroll
                for (int u = 0; u < (MG_MAX_SPLITS * MG_PART + MG_THREADS - 1) / MG_THREADS; ++u)
                    r[u] = ldg_cg_f(pp + min(u * MG_THREADS + tid, n_f - 1));
#pragma unroll
                for (int u = 0; u < (MG_MAX_SPLITS * MG_PART + MG_THREADS - 1) / MG_THREADS; ++u)
                    if (u * MG_THREADS + tid < n_f) pbuf[u * MG_THREADS + tid] = r[u];
                __syncthreads();
                if (tid < 64) {
                    float ms[MG_MAX_SPLITS], mm = -INFINITY;
#pragma unroll
                    for (int s = 0; s < MG_MAX_SPLITS; ++s) {
                        ms[s] = s < splits ? pbuf[s * MG_PART] : -INFINITY;
                        mm = fmaxf(mm, ms[s]);
                    }
                    float l = 0.f, o = 0.f;
#pragma unroll
                    for (int s = 0; s < MG_MAX_SPLITS; ++s) {
                        if (s < splits) {
                            const float w = __expf(ms[s] - mm);
                            l = fmaf(pbuf[s * MG_PART + 1], w, l);
                            o = fmaf(pbuf[s * MG_PART + 8 + tid], w, o);
                        }
                    }
                    p.ctx[(size_t)b * d + h * 64 + tid] = __float2bfloat16_rn(o / l);
                }
                if (tid == 0) item_cnt[item] = 0u;
            }
        }
        __syncthreads();                        // shared partials are reused by the next unit
    }
}

// ---- requests for what the NEXT phase reads first, issued between the arrival at the grid barrier and the wait (none of it
// depends on the other CTAs): the first weight tile of this warp, LayerNorm parameters, bias; the first attention unit
__device__ __forceinline__ void prefetch_linear(const MegaParams& p, const PhaseDesc* nd, int kind, uint4 (&first)[8]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const bf16* W = nd->W;
    const float* gamma = nd->gamma;
    const float* beta = nd->beta;
    const float* bias = nd->bias;
    const int N = nd->N, K = nd->K, C = K >> 5;
    if (gamma != nullptr) {
        const int lines = p.d >> 5;                   // 128-byte lines per fp32 vector of d elements (<= 32)
        if (tid < lines) prefetch_l2(gamma + tid * 32);
        else if (tid < 2 * lines) prefetch_l2(beta + (tid - lines) * 32);
    }
    const bool rows16 = kind != PH_LINEAR;
    const int ks_shift = kind == PH_HEAD ? 0 : kind == PH_LINEAR_WIDE ? 2 : 3, rt = rows16 ? 16 : 8, tpr = MG_WARPS >> ks_shift;
    const int tl = warp >> ks_shift, ks_id = warp & ((1 << ks_shift) - 1);
    const int c0 = (ks_id * C) >> ks_shift, c1 = ((ks_id + 1) * C) >> ks_shift;
    const int n0 = ((int)blockIdx.x * tpr + tl) * rt;
#pragma unroll
    for (int j = 0; j < 8; ++j) first[j] = make_uint4(0, 0, 0, 0);
    if (nd->rounds_base == 0 && (int)blockIdx.x * tpr >= nd->rounds_rem) return;   // this CTA has no round in that phase
    if (bias != nullptr && lane == 31 && n0 < N) prefetch_l2(bias + n0);
    // super-chunk 0 of round 0 of this warp, straight into the registers linear_phase starts from (same layout as its issue())
    const bf16* pa = W + (size_t)min(n0 + g, N - 1) * K + tq * 8;
    if (rows16) {
        const bf16* pb = W + (size_t)min(n0 + g + 8, N - 1) * K + tq * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j < c1) {
                first[2 * j] = ldg_nc16(pa + (c0 + j) * 32);
                first[2 * j + 1] = ldg_nc16(pb + (c0 + j) * 32);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c0 + j < c1) first[j] = ldg_nc16(pa + (c0 + j) * 32);
    }
    // the second super-chunk (fc2, LM head) into L2
    const int cps = rows16 ? 4 : 8;
    if (c0 + cps < c1) {
        const int row = rows16 ? lane >> 1 : lane >> 2, line = rows16 ? lane & 1 : lane & 3;
        if (c0 + cps + line * 2 < c1) prefetch_l2(W + (size_t)min(n0 + row, N - 1) * K + (c0 + cps + line * 2) * 32);
    }
}

__device__ __forceinline__ void prefetch_cross(const MegaParams& p, const bf16* kbase, const bf16* vbase) {
    const int splits = p.cross_splits, units = p.M * p.H * splits;
    const int unit = blockIdx.x;
    if (unit >= units) return;
    const int item = unit / splits, sp = unit - item * splits, b = item / p.H, h = item - b * p.H;
    const int s_beg = (int)((long long)sp * p.n_ctx / splits), s_end = (int)((long long)(sp + 1) * p.n_ctx / splits);
    const size_t off = (size_t)b * p.cross_bstride + ((size_t)h * p.n_ctx + s_beg) * 64;
    for (int i = threadIdx.x; i < s_end - s_beg; i += MG_THREADS) {   // one 128-byte row per request
        prefetch_l2(kbase + off + (size_t)i * 64);
        prefetch_l2(vbase + off + (size_t)i * 64);
    }
}

// pages of utterance b, head h (the newest row is still being written)
__device__ __forceinline__ void prefetch_self(const MegaParams& p, const bf16* kbase, const bf16* vbase, int n_keys) {
    const int item = blockIdx.x;
    if (item >= p.M * p.H) return;
    const int b = item / p.H, h = item - b * p.H;
    for (int s = threadIdx.x; s < n_keys - 1; s += MG_THREADS) {
        const int page = p.page_table[(size_t)b * p.pages_per_seq + (s >> 6)];
        const size_t off = (((size_t)page * p.H + h) * 64 + (s & 63)) * 64;
        prefetch_l2(kbase + off);
        prefetch_l2(vbase + off);
    }
}

template <int NM, bool kTC>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_step_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t mg_smem[];
    __shared__ __align__(16) PhaseDesc sdesc[2];
    if (p.state->active == 0) return;          // the loop has stopped: grid-uniform, nobody touches the barrier
    const int cur_len = p.state->cur_len, pos = cur_len - 1;
    const int tid = threadIdx.x;
    unsigned* bar = p.sync;
    unsigned* item_cnt = p.sync + 32;
    unsigned epoch = 0;
    const int n_phases = p.n_phases;           // 8 per layer + final LayerNorm / LM head
    long long* const trace = (p.trace != nullptr && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;
    if (tid < 16) reinterpret_cast<unsigned long long*>(&sdesc[0])[tid] = reinterpret_cast<const unsigned long long*>(p.table)[tid];
    __syncthreads();
    uint4 first[8];                            // first weight tile of the next linear phase, requested before its grid barrier
    prefetch_linear(p, &sdesc[0], sdesc[0].kind, first);
    if (trace != nullptr) trace[0] = clock64();
#pragma unroll 1
    for (int ph = 0; ph < n_phases; ++ph) {
        const PhaseDesc& D = sdesc[ph & 1];
        const int kind = D.kind;
        // the next descriptor is requested now and parked in a register: its round trip overlaps this phase's work
        unsigned long long nd_word = 0ull;
        if (tid < 16 && ph + 1 < n_phases) nd_word = __ldg(reinterpret_cast<const unsigned long long*>(p.table + ph + 1) + tid);
        if (kind != PH_SELF_ATTN && kind != PH_CROSS_ATTN) {
            // activations of a linear layer -> shared memory (ONE copy of this code for the three tile geometries below; the
            // first weight tile was requested into L2 before the grid barrier)
            bf16* act_s = reinterpret_cast<bf16*>(mg_smem + MG_RED_BYTES);
            if (D.stage == ST_COPY) stage_copy(act_s, D.src, p.M, D.K, D.vpr_rcp);
            else if (p.M <= 2) stage_layernorm_wide(p, act_s, reinterpret_cast<float*>(mg_smem), D.gamma, D.beta);
            else stage_layernorm_rows(p, act_s, D.gamma, D.beta);
            if (trace != nullptr) trace[8 * ph + 1] = clock64();
            __syncthreads();
            if (trace != nullptr) trace[8 * ph + 2] = clock64();
        }
        if (kind == PH_LINEAR) {
            linear_phase<NM, MG_LAYER_KS, false>(p, D, pos, mg_smem, trace != nullptr ? trace + 8 * ph : nullptr, first);
        } else if (kind == PH_LINEAR_WIDE) {
            linear_phase<NM, MG_WIDE_KS, true>(p, D, pos, mg_smem, trace != nullptr ? trace + 8 * ph : nullptr, first);
        } else if (kind == PH_HEAD) {
            linear_phase<NM, MG_HEAD_KS, true>(p, D, pos, mg_smem, trace != nullptr ? trace + 8 * ph : nullptr, first);
        } else if (kind == PH_SELF_ATTN) {   // cached self-attention over cur_len keys (newest row appended by the qkv epilogue)
            attention_phase<kTC>(p, true, D.k_pages, D.v_pages, cur_len, 1, mg_smem, item_cnt);
        } else {                             // cross-attention over the encoder K/V projected once per utterance
            attention_phase<kTC>(p, false, D.k_pages, D.v_pages, p.n_ctx, p.cross_splits, mg_smem, item_cnt);
        }
        if (trace != nullptr) trace[8 * ph + 4] = clock64();
        if (ph + 1 == n_phases) break;
        // arrive at the grid barrier, request what the next phase reads first, then wait
        if (tid < 16) reinterpret_cast<unsigned long long*>(&sdesc[(ph + 1) & 1])[tid] = nd_word;
        grid_arrive(bar);                         // (its block barrier publishes the descriptor to the CTA)
        const PhaseDesc* nd = &sdesc[(ph + 1) & 1];
        const int k2 = nd->kind;
        if (k2 == PH_CROSS_ATTN || k2 == PH_SELF_ATTN) {
#pragma unroll
            for (int j = 0; j < 8; ++j) first[j] = make_uint4(0, 0, 0, 0);      // dead across the attention phase
            if (k2 == PH_CROSS_ATTN) prefetch_cross(p, nd->k_pages, nd->v_pages);
            else prefetch_self(p, nd->k_pages, nd->v_pages, cur_len);
        } else {
            prefetch_linear(p, nd, k2, first);
        }
        if (trace != nullptr) trace[8 * ph + 5] = clock64();
        grid_wait(bar, epoch);
        if (trace != nullptr) trace[8 * ph + 8] = clock64();   // = slot 0 of the next phase
    }
    // leave the barrier counter at zero for the next launch: the last CTA to get here resets it (nobody polls it any more)
    __syncthreads();
    if (tid == 0) {
        const unsigned old = atomicAdd(bar, 1u);
        if (old == epoch + gridDim.x - 1) atomicExch(bar, 0u);
    }
}

int pick_cross_splits(int items, int n_keys, int grid) {
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= MG_MAX_SPLITS; ++s) {
        if (n_keys / s < 64) break;
        const int units = items * s, rounds = (units + grid - 1) / grid;
        const double cost = rounds * ((double)n_keys / s + 96.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    return best;
}

size_t mega_smem_bytes(int nm, int ffn) { return (size_t)MG_RED_BYTES + (size_t)8 * nm * (ffn + MG_ACT_PAD) * sizeof(bf16); }

void fill_geometry(PhaseDesc& o, int ks, bool rows16, int grid) {
    const int rt = rows16 ? 16 : 8, cps = rows16 ? 4 : 8, tpr = MG_WARPS / ks, C = o.K / 32;
    const int n_tiles = (o.N + rt - 1) / rt, stride = grid * tpr;
    o.nsc = ((C + ks - 1) / ks + cps - 1) / cps;
    // CTA c owns tiles (r * grid + c) * tpr + [0, tpr): it has a round r iff (r * grid + c) * tpr < n_tiles
    o.rounds_base = n_tiles / stride;
    o.rounds_rem = n_tiles % stride;
    o.vpr_rcp = (unsigned)((0x100000000ull + (unsigned)(o.K / 8) - 1) / (unsigned)(o.K / 8));
}
}  // namespace

// attention flavour of the whole-step kernel: 8 lanes per key with CUDA-core dot products (default) / mma.sync blocks of 16 keys.
// Measured (profiles/r01_kernel_variants.md): the mma.sync flavour needs 2.7x fewer instructions per key and is still 4-13 %
// SLOWER per step (B = 1 / 8 / 16: 1013 / 1386 / 2047 us vs 979 / 1277 / 1801) - it stays as a tested, switchable variant.
std::atomic<bool>& mega_attention_tc() {
    static std::atomic<bool> on{false};
    return on;
}
void set_mega_attention_tc(bool on) { mega_attention_tc() = on; }

std::atomic<long long*>& step_trace_ptr() {
    static std::atomic<long long*> ptr{nullptr};
    return ptr;
}
void set_step_trace(long long* dev_ptr) { step_trace_ptr() = dev_ptr; }

size_t mega_part_bytes(int max_batch, int heads) {
    return (size_t)std::min(max_batch, 16) * heads * MG_MAX_SPLITS * MG_PART * sizeof(float);
}
size_t mega_sync_bytes(int max_batch, int heads) { return (size_t)(32 + std::min(max_batch, 16) * heads) * sizeof(unsigned); }
size_t mega_table_bytes(int dec_layers) { return (size_t)(8 * dec_layers + 1) * sizeof(PhaseDesc); }

bool Session::mega_supported() const {
    const ModelConfig& g = m->cfg;
    return mega_table != nullptr && m->dtype == BF16 && batch >= 1 && batch <= 16 && g.d_model % 64 == 0 && g.d_model <= 1024 &&
           g.n_heads * 64 == g.d_model && g.ffn % 64 == 0 && g.ffn <= 4096 && g.dec_layers <= MG_MAX_LAYERS && g.vocab % 4 == 0 &&
           pages_per_seq <= 32 && mega_smem_bytes(batch <= 8 ? 1 : 2, g.ffn) <= 200 * 1024;
}

// the phase table of this session: every pointer and every division resolved once (called by the Session constructor)
void Session::build_mega_table() {
    const ModelConfig& g = m->cfg;
    mega_grid = 0;
    if (m->dtype != BF16 || mega_table == nullptr) return;
    int dev = 0;
    WB_CHECK_CUDA(cudaGetDevice(&dev));
    WB_CHECK_CUDA(cudaDeviceGetAttribute(&mega_grid, cudaDevAttrMultiProcessorCount, dev));
    const int d = g.d_model;
    std::vector<PhaseDesc> t((size_t)8 * g.dec_layers + 1);
    const size_t per_kv = (size_t)max_batch * g.n_heads * g.n_ctx * 64;
    auto linear = [&](PhaseDesc& o, const Linear& l, int stage, const LNorm* ln, const void* src, int epi, void* out_bf16) {
        std::memset(&o, 0, sizeof(o));
        o.kind = PH_LINEAR; o.W = (const bf16*)l.w; o.bias = l.b; o.N = l.n; o.K = l.k; o.stage = stage; o.epi = epi;
        if (ln != nullptr) { o.gamma = ln->g; o.beta = ln->b; }
        o.src = (const bf16*)src; o.out_bf16 = (bf16*)out_bf16;
        if ((l.n + 7) / 8 > 2 * mega_grid) {      // more than two rounds of 8-row tiles: 16-row tiles, two per CTA and round
            o.kind = PH_LINEAR_WIDE;
            fill_geometry(o, MG_WIDE_KS, true, mega_grid);
        } else {
            fill_geometry(o, MG_LAYER_KS, false, mega_grid);
        }
    };
    for (int l = 0; l < g.dec_layers; ++l) {
        const DecLayer& L = m->dec[l];
        PhaseDesc* o = &t[(size_t)8 * l];
        bf16* sk = (bf16*)self_k + (size_t)l * self_layer_elems();
        bf16* sv = (bf16*)self_v + (size_t)l * self_layer_elems();
        bf16* ck = (bf16*)cross + (size_t)l * cross_layer_elems();
        linear(o[0], L.qkv, ST_LN_X, &L.ln1, nullptr, EP_QKV, dq);
        o[0].k_pages = sk; o[0].v_pages = sv;
        std::memset(&o[1], 0, sizeof(PhaseDesc));
        o[1].kind = PH_SELF_ATTN; o[1].k_pages = sk; o[1].v_pages = sv;
        linear(o[2], L.out, ST_COPY, nullptr, datt, EP_RESIDUAL, nullptr);
        linear(o[3], L.cq, ST_LN_X, &L.ln2, nullptr, EP_BF16, dq);
        std::memset(&o[4], 0, sizeof(PhaseDesc));
        o[4].kind = PH_CROSS_ATTN; o[4].k_pages = ck; o[4].v_pages = ck + per_kv;
        linear(o[5], L.cout, ST_COPY, nullptr, datt, EP_RESIDUAL, nullptr);
        linear(o[6], L.fc1, ST_LN_X, &L.ln3, nullptr, EP_GELU_BF16, dffn);
        linear(o[7], L.fc2, ST_COPY, nullptr, dffn, EP_RESIDUAL, nullptr);
    }
    PhaseDesc& h = t.back();   // final LayerNorm + LM head (weights shared with the embedding table, no bias) -> fp32 logits
    std::memset(&h, 0, sizeof(h));
    h.kind = PH_HEAD; h.W = (const bf16*)m->emb; h.N = g.vocab; h.K = d; h.stage = ST_LN_X; h.epi = EP_F32;
    h.gamma = m->dec_ln.g; h.beta = m->dec_ln.b; h.out_f32 = logits;
    fill_geometry(h, MG_HEAD_KS, true, mega_grid);
    WB_CHECK_CUDA(cudaMemcpy(mega_table, t.data(), t.size() * sizeof(PhaseDesc), cudaMemcpyHostToDevice));
}

// one token for the whole batch in ONE cooperative launch (the logits processors / argmax kernel follows, decode_step)
void Session::decode_step_mega(cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    const int B = batch;
    WB_REQUIRE(mega_grid > 0, "whole-step kernel: no phase table");
    const int nm = B <= 8 ? 1 : 2;
    const size_t smem = mega_smem_bytes(nm, g.ffn);
    const bool tc = mega_attention_tc();
    auto kernel = nm == 1 ? (tc ? decode_step_mega_kernel<1, true> : decode_step_mega_kernel<1, false>)
                          : (tc ? decode_step_mega_kernel<2, true> : decode_step_mega_kernel<2, false>);
    {   // per device and kernel instance: the largest dynamic shared-memory size requested so far
        static std::mutex mu;
        static size_t configured[WB_MAX_DEVICES][3][2] = {};
        const int dev = current_device();
        std::lock_guard<std::mutex> lk(mu);
        if (configured[dev][nm][tc] < smem) {
            WB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev][nm][tc] = smem;
        }
    }
    int per_sm = 0;
    WB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MG_THREADS, smem));
    WB_REQUIRE(per_sm >= 1, "whole-step kernel does not fit on an SM");

    MegaParams p{};
    p.M = B; p.d = g.d_model; p.H = g.n_heads; p.ffn = g.ffn; p.vocab = g.vocab; p.n_phases = 8 * g.dec_layers + 1; p.n_ctx = g.n_ctx;
    p.cross_splits = pick_cross_splits(B * g.n_heads, g.n_ctx, mega_grid);
    p.pages_per_seq = pages_per_seq;
    p.cross_bstride = (long long)g.n_heads * g.n_ctx * 64;
    p.state = state; p.unfinished = unfinished; p.page_table = page_table;
    p.x = dx; p.q = (bf16*)dq; p.ctx = (bf16*)datt;
    p.part = mega_part; p.sync = mega_sync; p.trace = step_trace_ptr();
    p.table = reinterpret_cast<const PhaseDesc*>(mega_table);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(mega_grid);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident, or the launch fails: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    WB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
}

}  // namespace wb
