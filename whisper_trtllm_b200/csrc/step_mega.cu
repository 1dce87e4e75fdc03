// Persistent whole-step decoder kernel for small batches (B <= 16 utterances per GPU, bf16).
//
// At small batch the decode step is bound by kernel boundaries, not by bytes: 24 layers x 8..11 launches at ~7 us each is
// 10x above the weight-streaming floor (profiles/r01_decode_step_us.md).  Here ONE cooperative kernel (one 256-thread CTA per
// SM) runs the whole decoder for one token: embedding, 24 x {LN1+qkv -> paged self-attention -> out-proj(+res) -> LN2+cross-q
// -> cross-attention -> cross-out(+res) -> LN3+fc1(GELU) -> fc2(+res)}, final LN + LM head.  The phases are separated by a
// grid barrier (one release-add + acquire-poll in L2) instead of a kernel boundary, and
//   * the weights of the NEXT phase are requested before the barrier (L2 prefetch) and its first tile is loaded into
//     registers while the activations are being staged (LayerNorm) - weight latency never sits on the critical path;
//   * linear layers run "swap-AB" on mma.sync m16n8k16: the 16-row MMA dimension walks the WEIGHT rows (streamed straight from
//     global memory into the A fragment with 16-byte loads, K permuted consistently in both operands), the 8-column dimension
//     holds the (<= 8 / <= 16) utterances staged once per CTA as bf16 in shared memory; fp32 accumulation;
//     small layers split K over the 8 warps of a CTA and reduce through shared memory in a fixed order (deterministic);
//   * attention items are split over CTAs along the keys (cross: 1500 frames / split) so that a single utterance still uses
//     every SM; partials (max, sum, acc[64]) are merged in a fixed order by the last CTA to arrive at the item's counter.
// Semantics and rounding points are those of the multi-kernel paths (runtime.cu decode_step_small / decode_step_large):
// LayerNorm eps 1e-5 in fp32 (layers/normalization.py:6-30), bf16 activations into every Linear (layers/linear.py:38-139),
// fp32 residual stream, erf GELU, q pre-scaled in the packed weights, fp32 softmax, no mask (model.py:240-304),
// position = cur_len - 1 (model.py:423-425), tied LM head without bias (modeling_whisper.py:1335,1433).
// (tcgen05 needs M = 128 row tiles and a TMEM round trip per phase: at <= 16 rows mma.sync from registers is the right tool.)
#include <algorithm>

#include "wb_runtime.h"

namespace wb {

namespace {
constexpr int MG_THREADS = 256, MG_WARPS = 8;
constexpr int MG_MAX_LAYERS = 32, MG_MAX_SPLITS = 16;
constexpr int MG_PART = 72;          // floats per attention partial: [0] max, [1] sum, [8..72) acc
constexpr int MG_RS = 20;            // reduction buffer: floats between consecutive utterance rows (bank-conflict free)
constexpr int MG_RED_BYTES = 2 * MG_WARPS * 16 * MG_RS * 4;   // two parities x 8 warps x 16 rows
constexpr int MG_ACT_PAD = 32;       // bf16 elements of padding per staged row (row stride = 64 mod 128 bytes)
constexpr unsigned MG_SPIN_LIMIT = 1u << 26;
constexpr int MG_ATT_UNROLL = 4;

struct MegaLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    const bf16 *qkv_w, *out_w, *cq_w, *cout_w, *fc1_w, *fc2_w;
    const float *qkv_b, *out_b, *cq_b, *cout_b, *fc1_b, *fc2_b;
    bf16 *self_k, *self_v;
    const bf16 *cross_k, *cross_v;
};

// work split of one linear layer over the grid: tiles of 8 or 16 weight rows, K split over `ks` warps of a CTA
// (everything that needs a division is computed once on the host: the phases are latency-bound instruction streams)
struct GemvCfg {
    int ks, ks_shift, rows16;
    int C;                 // 32-element chunks along K
    int cps, nsc;          // chunks per super-chunk (8 x 16-byte loads per lane), super-chunks per tile
    int n_tiles, tpr;      // tiles of 8 / 16 weight rows, tiles per CTA and round
    int rounds_base, rounds_rem;   // rounds of CTA c = rounds_base + (c * tpr < rounds_rem)
};

struct MegaParams {
    int M, d, H, ffn, vocab, n_layers, n_ctx, cross_splits, pages_per_seq, tokens_stride;
    long long cross_bstride;
    GemvCfg c_qkv, c_dd, c_fc1, c_fc2, c_head;
    const int* tokens; const StepState* state; const int* unfinished; const int* page_table;
    const bf16 *emb, *pos; const float *lnf_g, *lnf_b;
    float* x; bf16 *q, *ctx, *ffn_act; float* logits; float* part;
    long long* trace;  // optional (tools/step_trace.py): SM clock stamps of CTA 0, 8 slots per phase (see the kernel body)
    unsigned* sync;    // [0] grid barrier counter, [32 ..) per-item arrival counters; all zero between launches
    MegaLayer layer[MG_MAX_LAYERS];
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
// LayerNorm staging, one warp per row (rows m = warp, warp + 8): the 24 loads of a row (x, gamma, beta: 8 float4 each per lane
// at d = 1024) are all issued before the first use - one round trip - and the statistics need only warp shuffles.  The phases
// are latency-bound instruction streams (2 warps per scheduler): what counts is the number of instructions per warp on the
// critical path.  (First version: guarded per-element loop, ptxas serialised the gamma / beta loads into 8 dependent round
// trips; second version: every thread touched every row - 1300 instructions per warp, 6.4 K cycles, profiles/r01_step_trace*.)
// Rows come from the fp32 residual stream, or (first layer) straight from the embedding tables: x = E[token] + P[position]
// with token = ids[m, cur_len - 1] (model.py:423-425); CTA 0 then also writes the residual stream.
template <bool kFull>   // kFull: d == 1024, every lane owns 8 float4 of the row (no predicates)
__device__ __forceinline__ void stage_layernorm(const MegaParams& p, bf16* act_s, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, bool embed, int pos) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.d, nvec = d >> 2, astride = d + MG_ACT_PAD;
    const float inv_d = 1.0f / (float)d;
    for (int m = warp; m < p.M; m += MG_WARPS) {
        float4 v[8];
        if (embed) {
            const int tok = p.tokens[(size_t)m * p.tokens_stride + pos];
            const uint2* erow = reinterpret_cast<const uint2*>(p.emb + (size_t)tok * d);
            const uint2* prow = reinterpret_cast<const uint2*>(p.pos + (size_t)pos * d);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = lane + 32 * i;
                const uint2 e = (kFull || idx < nvec) ? __ldg(erow + idx) : make_uint2(0u, 0u);
                const uint2 q = (kFull || idx < nvec) ? __ldg(prow + idx) : make_uint2(0u, 0u);
                v[i].x = __uint_as_float(e.x << 16) + __uint_as_float(q.x << 16);
                v[i].y = __uint_as_float(e.x & 0xffff0000u) + __uint_as_float(q.x & 0xffff0000u);
                v[i].z = __uint_as_float(e.y << 16) + __uint_as_float(q.y << 16);
                v[i].w = __uint_as_float(e.y & 0xffff0000u) + __uint_as_float(q.y & 0xffff0000u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = lane + 32 * i;
                v[i] = (kFull || idx < nvec) ? ldg_cg_f4(p.x + (size_t)m * d + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);      // lanes past the row hold zeros
        const float mean = warp_sum(s) * inv_d;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kFull || lane + 32 * i < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                ss += (a * a + b * b) + (c * c + e * e);
            }
        }
        const float rstd = rsqrtf(warp_sum(ss) * inv_d + 1e-5f);
        // gamma / beta were requested into L2 before the grid barrier: two half-batches of 8 loads (an L2 round trip each) keep
        // the register peak at v[8] + 8 float4 next to the weight tile that is already in flight
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 g4[4], b4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * (half * 4 + j);
                g4[j] = (kFull || idx < nvec) ? __ldg(reinterpret_cast<const float4*>(gamma) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                b4[j] = (kFull || idx < nvec) ? __ldg(reinterpret_cast<const float4*>(beta) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = half * 4 + j, idx = lane + 32 * i;
                if (kFull || idx < nvec) {
                    if (embed && blockIdx.x == 0) *reinterpret_cast<float4*>(p.x + (size_t)m * d + idx * 4) = v[i];
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

// bf16 rows of the previous phase -> shared memory; 8 independent 16-byte requests per thread and trip
__device__ __forceinline__ void stage_copy(bf16* act_s, const bf16* src, int M, int K) {
    const int vpr = K >> 3, astride = K + MG_ACT_PAD, total = M * vpr;
    for (int base = 0; base < total; base += MG_THREADS * 8) {
        uint4 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MG_THREADS + (int)threadIdx.x;   // no clamped duplicates: 148 SMs x 256 threads asking for one
            r[u] = i < total ? ldg_cg16(src + (size_t)i * 8) : make_uint4(0, 0, 0, 0);   // line serialise in its L2 slice
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MG_THREADS + (int)threadIdx.x;
            if (i < total) {
                const int m = i / vpr, j = i - m * vpr;
                *reinterpret_cast<uint4*>(act_s + (size_t)m * astride + j * 8) = r[u];
            }
        }
    }
}

// ---- geometry of one linear layer on this CTA / warp
struct GemvGeom {
    int rt_shift, rt, tpr, tl, ks_id, c0, c1, cps, nsc, rounds;
    __device__ __forceinline__ GemvGeom(const GemvCfg& cfg) {
        const int warp = threadIdx.x >> 5;
        rt_shift = cfg.rows16 ? 4 : 3;
        rt = 1 << rt_shift;
        tpr = cfg.tpr;
        tl = warp >> cfg.ks_shift;
        ks_id = warp & (cfg.ks - 1);
        c0 = (ks_id * cfg.C) >> cfg.ks_shift;
        c1 = ((ks_id + 1) * cfg.C) >> cfg.ks_shift;
        cps = cfg.cps;
        nsc = cfg.nsc;
        rounds = cfg.rounds_base + ((int)blockIdx.x * tpr < cfg.rounds_rem ? 1 : 0);
    }
    __device__ __forceinline__ int tile_of(int r, int t_local) const { return (r * (int)gridDim.x + (int)blockIdx.x) * tpr + t_local; }
};

// One linear layer of the step, described at run time.  The kernel body is a small interpreter over these descriptors with ONE
// copy of the linear-layer code and one of each attention flavour: the whole step has to stay inside the instruction cache
// (a first version that inlined the eight phases of a layer was 20 K instructions per layer iteration and ran 1.6x SLOWER
// than the multi-kernel path: every phase was an instruction-cache miss streak).
enum { ST_LN_X = 0, ST_LN_EMBED = 1, ST_COPY = 2 };
enum { EP_QKV = 0, EP_RESIDUAL = 1, EP_BF16 = 2, EP_GELU_BF16 = 3, EP_F32 = 4 };
struct LinearPhase {
    const bf16* W; const float* bias; int N, K; GemvCfg cfg;
    int stage; const float *gamma, *beta; const bf16* src;      // LayerNorm parameters / bf16 rows to copy
    int epi; bf16* out_bf16; float* out_f32; bf16 *k_pages, *v_pages;
};

// phase k of layer l: 0 LN1+qkv, 2 out-proj, 3 LN2+cross-q, 5 cross-out, 6 LN3+fc1, 7 fc2, 8 final LN + LM head
__device__ __forceinline__ void make_linear_phase(const MegaParams& p, int l, int k, LinearPhase& o) {
    const MegaLayer& L = p.layer[l < p.n_layers ? l : 0];
    o.src = nullptr; o.gamma = nullptr; o.beta = nullptr; o.out_bf16 = nullptr; o.out_f32 = nullptr; o.k_pages = nullptr; o.v_pages = nullptr;
    o.K = p.d; o.N = p.d; o.cfg = p.c_dd; o.stage = ST_LN_X; o.epi = EP_RESIDUAL;
    switch (k) {
        case 0:
            o.W = L.qkv_w; o.bias = L.qkv_b; o.N = 3 * p.d; o.cfg = p.c_qkv; o.stage = l == 0 ? ST_LN_EMBED : ST_LN_X;
            o.gamma = L.ln1_g; o.beta = L.ln1_b; o.epi = EP_QKV; o.out_bf16 = p.q; o.k_pages = L.self_k; o.v_pages = L.self_v;
            break;
        case 2: o.W = L.out_w; o.bias = L.out_b; o.stage = ST_COPY; o.src = p.ctx; break;
        case 3:
            o.W = L.cq_w; o.bias = L.cq_b; o.gamma = L.ln2_g; o.beta = L.ln2_b; o.epi = EP_BF16; o.out_bf16 = p.q;
            break;
        case 5: o.W = L.cout_w; o.bias = L.cout_b; o.stage = ST_COPY; o.src = p.ctx; break;
        case 6:
            o.W = L.fc1_w; o.bias = L.fc1_b; o.N = p.ffn; o.cfg = p.c_fc1; o.gamma = L.ln3_g; o.beta = L.ln3_b;
            o.epi = EP_GELU_BF16; o.out_bf16 = p.ffn_act;
            break;
        case 7: o.W = L.fc2_w; o.bias = L.fc2_b; o.K = p.ffn; o.cfg = p.c_fc2; o.stage = ST_COPY; o.src = p.ffn_act; break;
        default:
            o.W = p.emb; o.bias = nullptr; o.N = p.vocab; o.cfg = p.c_head; o.gamma = p.lnf_g; o.beta = p.lnf_b;
            o.epi = EP_F32; o.out_f32 = p.logits;
            break;
    }
}

// request what the next linear layer reads first into L2: this warp's first weight tile, the LayerNorm parameters, the bias
// of the tile (issued between the arrival at the grid barrier and the wait: none of it depends on the other CTAs)
__device__ __forceinline__ void prefetch_linear(const MegaParams& p, const LinearPhase& ph) {
    const GemvGeom gm(ph.cfg);
    const int tid = threadIdx.x, lane = tid & 31;
    if (ph.gamma != nullptr) {
        const int lines = p.d >> 5;                   // 128-byte lines per fp32 vector of d elements (<= 32)
        if (tid < lines) prefetch_l2(ph.gamma + tid * 32);
        else if (tid < 2 * lines) prefetch_l2(ph.beta + (tid - lines) * 32);
    }
    if (gm.rounds == 0) return;
    const int n0 = gm.tile_of(0, gm.tl) << gm.rt_shift;
    if (ph.bias != nullptr && lane == 31) prefetch_l2(ph.bias + min(n0, ph.N - 1));
    const int lines = ((gm.c1 - gm.c0) * 64 + 127) >> 7;      // <= 32 for K <= 4096
    if (lane < lines) {
        const bf16* base = ph.W + gm.c0 * 32 + lane * 64;
        for (int r = 0; r < gm.rt; ++r) prefetch_l2(base + (size_t)min(n0 + r, ph.N - 1) * ph.K);
    }
}

// called exactly once per output element (n, m) by exactly one thread of the grid
__device__ __forceinline__ void linear_epilogue(const MegaParams& p, const LinearPhase& ph, int pos, int n, int m, float v) {
    if (n >= ph.N || m >= p.M) return;
    if (ph.bias != nullptr) v += __ldg(ph.bias + n);
    const int d = p.d;
    switch (ph.epi) {
        case EP_QKV: {   // q -> activation buffer; k / v rows straight into the paged cache at slot cur_len - 1
            const int which = (n >= d ? 1 : 0) + (n >= 2 * d ? 1 : 0), c = n - which * d;
            if (which == 0) {
                ph.out_bf16[(size_t)m * d + c] = __float2bfloat16_rn(v);
            } else {
                const int page = p.page_table[(size_t)m * p.pages_per_seq + (pos >> 6)];
                const size_t off = (((size_t)page * p.H + (c >> 6)) * 64 + (pos & 63)) * 64 + (c & 63);
                (which == 1 ? ph.k_pages : ph.v_pages)[off] = __float2bfloat16_rn(v);
            }
            break;
        }
        case EP_RESIDUAL: {   // fp32 residual stream, updated in place: this thread owns (m, n)
            float* xp = p.x + (size_t)m * d + n;
            *xp = ldg_cg_f(xp) + v;
            break;
        }
        case EP_BF16: ph.out_bf16[(size_t)m * ph.N + n] = __float2bfloat16_rn(v); break;
        case EP_GELU_BF16: ph.out_bf16[(size_t)m * ph.N + n] = __float2bfloat16_rn(gelu_erf_fast(v)); break;
        default: ph.out_f32[(size_t)m * ph.N + n] = v; break;
    }
}

// out[m, n] = epilogue( sum_k act[m, k] * W[n, k] ) for all n < N, m < 8 * NM
template <int NM>
__device__ __forceinline__ void linear_phase(const MegaParams& p, const LinearPhase& ph, int pos, uint8_t* smem, long long* tr) {
    float* red_s = reinterpret_cast<float*>(smem);
    bf16* act_s = reinterpret_cast<bf16*>(smem + MG_RED_BYTES);
    const bf16* __restrict__ W = ph.W;
    const int N = ph.N, K = ph.K;
    const GemvCfg cfg = ph.cfg;
    const int astride = K + MG_ACT_PAD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const GemvGeom gm(cfg);
    const int total = gm.rounds * gm.nsc;
    const bool rows16 = cfg.rows16 != 0;

    // (r, sc) = (round, super-chunk) of the tile being requested / computed; advanced incrementally, no divisions
    auto issue = [&](int idx, int r, int sc, uint4(&buf)[8]) __attribute__((always_inline)) {
        const int n0 = gm.tile_of(r, gm.tl) << gm.rt_shift;
        const bf16* pa = W + (size_t)min(n0 + g, N - 1) * K + tq * 8;
        if (rows16) {
            const bf16* pb = W + (size_t)min(n0 + g + 8, N - 1) * K + tq * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = gm.c0 + sc * 4 + j;
                const bool in = chunk < gm.c1;
                buf[2 * j] = in ? ldg_nc16(pa + chunk * 32) : make_uint4(0, 0, 0, 0);
                buf[2 * j + 1] = in ? ldg_nc16(pb + chunk * 32) : make_uint4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int chunk = gm.c0 + sc * 8 + j;
                buf[j] = chunk < gm.c1 ? ldg_nc16(pa + chunk * 32) : make_uint4(0, 0, 0, 0);
            }
        }
        if (idx + 3 < total) {   // keep the stream ahead of the two register buffers: one L2 prefetch per lane, 3 super-chunks on
            int r3 = r, sc3 = sc + 3;
            while (sc3 >= gm.nsc) { sc3 -= gm.nsc; ++r3; }
            const int n3 = gm.tile_of(r3, gm.tl) << gm.rt_shift;
            const int row = rows16 ? lane >> 1 : lane >> 2, line = rows16 ? lane & 1 : lane & 3;   // 256 / 512 bytes per row
            const int chunk = gm.c0 + sc3 * gm.cps + line * 2;
            if (chunk < gm.c1) prefetch_l2(W + (size_t)min(n3 + row, N - 1) * K + chunk * 32);
        }
    };

    float acc[NM][4];
    auto finish = [&](int r) __attribute__((always_inline)) {
        const int n0 = gm.tile_of(r, gm.tl) << gm.rt_shift;
        if (cfg.ks == 1) {
#pragma unroll
            for (int mb = 0; mb < NM; ++mb) {
                linear_epilogue(p, ph, pos, n0 + g, mb * 8 + 2 * tq, acc[mb][0]);
                linear_epilogue(p, ph, pos, n0 + g, mb * 8 + 2 * tq + 1, acc[mb][1]);
                if (rows16) {
                    linear_epilogue(p, ph, pos, n0 + g + 8, mb * 8 + 2 * tq, acc[mb][2]);
                    linear_epilogue(p, ph, pos, n0 + g + 8, mb * 8 + 2 * tq + 1, acc[mb][3]);
                }
            }
        } else {
            constexpr int WSTRIDE = 16 * MG_RS;   // floats per warp slab: 16 utterance rows x MG_RS
            float* rw = red_s + ((r & 1) * MG_WARPS + warp) * WSTRIDE;
#pragma unroll
            for (int mb = 0; mb < NM; ++mb) {
                rw[(mb * 8 + 2 * tq) * MG_RS + g] = acc[mb][0];
                rw[(mb * 8 + 2 * tq + 1) * MG_RS + g] = acc[mb][1];
                rw[(mb * 8 + 2 * tq) * MG_RS + g + 8] = acc[mb][2];
                rw[(mb * 8 + 2 * tq + 1) * MG_RS + g + 8] = acc[mb][3];
            }
            __syncthreads();
            const int m_cnt = 8 * NM;
            const int outs = (gm.tpr * m_cnt) << gm.rt_shift;
            for (int o = tid; o < outs; o += MG_THREADS) {
                const int n_l = o & (gm.rt - 1);
                const int rest = o >> gm.rt_shift;
                const int m = rest % m_cnt, t2 = rest / m_cnt;
                const float* rr = red_s + ((r & 1) * MG_WARPS + t2 * cfg.ks) * WSTRIDE + m * MG_RS + n_l;
                float v = 0.f;
                for (int k = 0; k < cfg.ks; ++k) v += rr[k * WSTRIDE];     // fixed order: deterministic
                linear_epilogue(p, ph, pos, (gm.tile_of(r, t2) << gm.rt_shift) + n_l, m, v);
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
        if (rows16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = gm.c0 + sc * 4 + j;
                if (chunk < gm.c1) {   // warp-uniform
#pragma unroll
                    for (int mb = 0; mb < NM; ++mb) {
                        const uint4 a = *reinterpret_cast<const uint4*>(arow + (size_t)mb * 8 * astride + chunk * 32);
                        mma_16816(acc[mb], buf[2 * j].x, buf[2 * j + 1].x, buf[2 * j].y, buf[2 * j + 1].y, a.x, a.y);
                        mma_16816(acc[mb], buf[2 * j].z, buf[2 * j + 1].z, buf[2 * j].w, buf[2 * j + 1].w, a.z, a.w);
                    }
                }
            }
        } else {               // 8-row tiles: MMA rows 8..15 are zero
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int chunk = gm.c0 + sc * 8 + j;
                if (chunk < gm.c1) {
#pragma unroll
                    for (int mb = 0; mb < NM; ++mb) {
                        const uint4 a = *reinterpret_cast<const uint4*>(arow + (size_t)mb * 8 * astride + chunk * 32);
                        mma_16816(acc[mb], buf[j].x, 0u, buf[j].y, 0u, a.x, a.y);
                        mma_16816(acc[mb], buf[j].z, 0u, buf[j].w, 0u, a.z, a.w);
                    }
                }
            }
        }
        if (sc == gm.nsc - 1) finish(r);
    };

    uint4 cur[8], nxt[8];
    if (total > 0) {
        issue(0, 0, 0, cur);     // in flight while the activations are staged
        if (tr != nullptr) tr[1] = clock64();
        if (ph.stage == ST_COPY) stage_copy(act_s, ph.src, p.M, K);
        else if (p.d == 1024) stage_layernorm<true>(p, act_s, ph.gamma, ph.beta, ph.stage == ST_LN_EMBED, pos);
        else stage_layernorm<false>(p, act_s, ph.gamma, ph.beta, ph.stage == ST_LN_EMBED, pos);
    }
    if (tr != nullptr) tr[2] = clock64();
    __syncthreads();
    if (tr != nullptr) tr[3] = clock64();
    int r = 0, sc = 0;
#pragma unroll 1
    for (int idx = 0; idx < total; ++idx) {   // trip counts are CTA-uniform (finish() contains a block barrier)
        int r1 = r, sc1 = sc + 1;
        if (sc1 == gm.nsc) { sc1 = 0; ++r1; }
        if (idx + 1 < total) issue(idx + 1, r1, sc1, nxt);
        compute(r, sc, cur);
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        r = r1;
        sc = sc1;
    }
}

// ---- one-query attention, one CTA per (item, key split) unit; 8 lanes per key row, 4 rows per warp instruction
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
        // software pipeline: the 16 requests per lane of batch i + 1 are in flight while batch i is reduced (one CTA per SM
        // and 8 warps in lock step: without it the SM alternates between waiting for HBM and computing, 13 B/clk)
        auto issue = [&](int sb, uint4(&kr)[MG_ATT_UNROLL], uint4(&vr)[MG_ATT_UNROLL]) __attribute__((always_inline)) {
#pragma unroll
            for (int u = 0; u < MG_ATT_UNROLL; ++u) {
                const size_t off = row_off(min(sb + grp + u * 32, s_end - 1));
                kr[u] = ldg_cg16(kbase + off);          // L2-coherent: the newest self-attention row was written by another CTA
                vr[u] = ldg_cg16(vbase + off);          // in the previous phase (cross K/V are read-only; same L1 bypass)
            }
        };
        uint4 kr[MG_ATT_UNROLL], vr[MG_ATT_UNROLL], kn[MG_ATT_UNROLL], vn[MG_ATT_UNROLL];
        const int sb0 = s_beg + warp * 4;
        if (sb0 < s_end) issue(sb0, kr, vr);
        for (int sb = sb0; sb < s_end; sb += 32 * MG_ATT_UNROLL) {   // warp-uniform trip count
            const bool more = sb + 32 * MG_ATT_UNROLL < s_end;
            if (more) issue(sb + 32 * MG_ATT_UNROLL, kn, vn);
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
            if (more) {
#pragma unroll
                for (int u = 0; u < MG_ATT_UNROLL; ++u) { kr[u] = kn[u]; vr[u] = vn[u]; }
            }
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
                    float mm = -INFINITY;
                    for (int s = 0; s < splits; ++s) mm = fmaxf(mm, pbuf[s * MG_PART]);
                    float l = 0.f, o = 0.f;
                    for (int s = 0; s < splits; ++s) {
                        const float w = __expf(pbuf[s * MG_PART] - mm);
                        l = fmaf(pbuf[s * MG_PART + 1], w, l);
                        o = fmaf(pbuf[s * MG_PART + 8 + tid], w, o);
                    }
                    p.ctx[(size_t)b * d + h * 64 + tid] = __float2bfloat16_rn(o / l);
                }
                if (tid == 0) item_cnt[item] = 0u;
            }
        }
        __syncthreads();                        // shared partials are reused by the next unit
    }
}

// request the first cross-attention unit of this CTA into L2 (before the barrier in front of the phase)
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

// same for the first self-attention item of this CTA (pages of utterance b, head h; the newest row is still being written)
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

template <int NM>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_step_mega_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t mg_smem[];
    if (p.state->active == 0) return;          // the loop has stopped: grid-uniform, nobody touches the barrier
    const int cur_len = p.state->cur_len, pos = cur_len - 1;
    unsigned* bar = p.sync;
    unsigned* item_cnt = p.sync + 32;
    unsigned epoch = 0;
    const int n_phases = 8 * p.n_layers + 1;   // 8 per layer + final LayerNorm / LM head
    long long* const trace = (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) ? p.trace : nullptr;
    if (trace != nullptr) trace[0] = clock64();
#pragma unroll 1
    for (int ph = 0; ph < n_phases; ++ph) {
        const int l = ph >> 3, k = (ph == n_phases - 1) ? 8 : (ph & 7);
        if (k == 1) {          // cached self-attention over cur_len keys (the newest row was appended by phase 0's epilogue)
            attention_phase(p, true, p.layer[l].self_k, p.layer[l].self_v, cur_len, 1, mg_smem, item_cnt);
        } else if (k == 4) {   // cross-attention over the encoder K/V projected once per utterance
            attention_phase(p, false, p.layer[l].cross_k, p.layer[l].cross_v, p.n_ctx, p.cross_splits, mg_smem, item_cnt);
        } else {
            LinearPhase lp;
            make_linear_phase(p, l, k, lp);
            linear_phase<NM>(p, lp, pos, mg_smem, trace != nullptr ? trace + 8 * ph : nullptr);
        }
        if (trace != nullptr) trace[8 * ph + 4] = clock64();
        if (ph + 1 == n_phases) break;
        // arrive at the grid barrier, request what the next phase reads first, then wait for the other CTAs
        grid_arrive(bar);
        const int l2 = (ph + 1) >> 3, k2 = (ph + 2 == n_phases) ? 8 : ((ph + 1) & 7);
        if (k2 == 4) {
            prefetch_cross(p, p.layer[l2].cross_k, p.layer[l2].cross_v);
        } else if (k2 == 1) {
            prefetch_self(p, p.layer[l2].self_k, p.layer[l2].self_v, cur_len);
        } else {
            LinearPhase np;
            make_linear_phase(p, l2, k2, np);
            prefetch_linear(p, np);
        }
        if (trace != nullptr) trace[8 * ph + 5] = clock64();
        grid_wait(bar, epoch);
        if (trace != nullptr) trace[8 * ph + 8] = clock64();   // = slot 0 of the next phase
    }
    // leave the barrier counter at zero for the next launch: the last CTA to get here resets it (nobody polls it any more)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(bar, 1u);
        if (old == epoch + gridDim.x - 1) atomicExch(bar, 0u);
    }
}


// rounds x (serial 16-byte loads per lane + fixed per-round cost): smallest wins, ties go to fewer K splits
GemvCfg pick_gemv_cfg(int N, int K, int grid) {
    GemvCfg best{};
    double best_cost = 1e30;
    const int C = K / 32;
    for (int rows16 = 1; rows16 >= 0; --rows16) {
        for (int ks = 1, shift = 0; ks <= 8; ks *= 2, ++shift) {
            if (C < ks) break;
            const int rt = rows16 ? 16 : 8, tpr = MG_WARPS / ks;
            const int n_tiles = (N + rt - 1) / rt;
            const int stride = grid * tpr;
            const int rounds = (n_tiles + stride - 1) / stride;
            const double cost = rounds * ((double)((C + ks - 1) / ks) * (rows16 ? 1.0 : 0.5) + 2.0);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                GemvCfg c{};
                c.ks = ks; c.ks_shift = shift; c.rows16 = rows16; c.C = C;
                c.cps = rows16 ? 4 : 8;
                c.nsc = ((C + ks - 1) / ks + c.cps - 1) / c.cps;
                c.n_tiles = n_tiles; c.tpr = tpr;
                // CTA c owns tiles (r * grid + c) * tpr + [0, tpr): it has a round r iff (r * grid + c) * tpr < n_tiles
                c.rounds_base = n_tiles / stride;
                c.rounds_rem = n_tiles % stride;
                best = c;
            }
        }
    }
    return best;
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
}  // namespace

long long*& step_trace_ptr() {
    static long long* ptr = nullptr;
    return ptr;
}
void set_step_trace(long long* dev_ptr) { step_trace_ptr() = dev_ptr; }

size_t mega_part_bytes(int max_batch, int heads) {
    return (size_t)std::min(max_batch, 16) * heads * MG_MAX_SPLITS * MG_PART * sizeof(float);
}
size_t mega_sync_bytes(int max_batch, int heads) { return (size_t)(32 + std::min(max_batch, 16) * heads) * sizeof(unsigned); }

bool Session::mega_supported() const {
    const ModelConfig& g = m->cfg;
    return m->dtype == BF16 && batch >= 1 && batch <= 16 && g.d_model % 64 == 0 && g.d_model <= 1024 && g.n_heads * 64 == g.d_model &&
           g.ffn % 64 == 0 && g.dec_layers <= MG_MAX_LAYERS && g.vocab % 4 == 0 && pages_per_seq <= 32 &&
           mega_smem_bytes(batch <= 8 ? 1 : 2, g.ffn) <= 220 * 1024;
}

// one token for the whole batch in ONE cooperative launch (the logits processors / argmax kernel follows, decode_step)
void Session::decode_step_mega(cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, B = batch;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        WB_CHECK_CUDA(cudaGetDevice(&dev));
        WB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int nm = B <= 8 ? 1 : 2;
    const size_t smem = mega_smem_bytes(nm, g.ffn);
    auto kernel = nm == 1 ? decode_step_mega_kernel<1> : decode_step_mega_kernel<2>;
    static size_t configured[3] = {0, 0, 0};
    if (configured[nm] < smem) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[nm] = smem;
    }
    int per_sm = 0;
    WB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MG_THREADS, smem));
    WB_REQUIRE(per_sm >= 1, "whole-step kernel does not fit on an SM");
    const int grid = sms;

    MegaParams p{};
    p.M = B; p.d = d; p.H = g.n_heads; p.ffn = g.ffn; p.vocab = g.vocab; p.n_layers = g.dec_layers; p.n_ctx = g.n_ctx;
    p.cross_splits = pick_cross_splits(B * g.n_heads, g.n_ctx, grid);
    p.pages_per_seq = pages_per_seq; p.tokens_stride = g.max_tgt;
    p.cross_bstride = (long long)g.n_heads * g.n_ctx * 64;
    p.c_qkv = pick_gemv_cfg(3 * d, d, grid);
    p.c_dd = pick_gemv_cfg(d, d, grid);
    p.c_fc1 = pick_gemv_cfg(g.ffn, d, grid);
    p.c_fc2 = pick_gemv_cfg(d, g.ffn, grid);
    p.c_head = pick_gemv_cfg(g.vocab, d, grid);
    p.tokens = tokens; p.state = state; p.unfinished = unfinished; p.page_table = page_table;
    p.emb = (const bf16*)m->emb; p.pos = (const bf16*)m->dec_pos; p.lnf_g = m->dec_ln.g; p.lnf_b = m->dec_ln.b;
    p.x = dx; p.q = (bf16*)dq; p.ctx = (bf16*)datt; p.ffn_act = (bf16*)dffn; p.logits = logits;
    p.part = mega_part; p.sync = mega_sync; p.trace = step_trace_ptr();
    const size_t per_kv = (size_t)max_batch * g.n_heads * g.n_ctx * 64;
    for (int l = 0; l < g.dec_layers; ++l) {
        const DecLayer& L = m->dec[l];
        MegaLayer& o = p.layer[l];
        o.ln1_g = L.ln1.g; o.ln1_b = L.ln1.b; o.ln2_g = L.ln2.g; o.ln2_b = L.ln2.b; o.ln3_g = L.ln3.g; o.ln3_b = L.ln3.b;
        o.qkv_w = (const bf16*)L.qkv.w; o.out_w = (const bf16*)L.out.w; o.cq_w = (const bf16*)L.cq.w;
        o.cout_w = (const bf16*)L.cout.w; o.fc1_w = (const bf16*)L.fc1.w; o.fc2_w = (const bf16*)L.fc2.w;
        o.qkv_b = L.qkv.b; o.out_b = L.out.b; o.cq_b = L.cq.b; o.cout_b = L.cout.b; o.fc1_b = L.fc1.b; o.fc2_b = L.fc2.b;
        o.self_k = (bf16*)self_k + (size_t)l * self_layer_elems();
        o.self_v = (bf16*)self_v + (size_t)l * self_layer_elems();
        o.cross_k = (const bf16*)cross + (size_t)l * cross_layer_elems();
        o.cross_v = o.cross_k + per_kv;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
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
