// Fused decode-step chains for large batches (17 <= B <= 512 utterances per GPU, bf16).
//
// Between two attention kernels the decode step is a chain of skinny GEMMs (M = B rows) and LayerNorms:
//     out-proj -> (+res) LN2 -> cross-q            |  cross-out -> (+res) LN3 -> fc1 (GELU) -> fc2 -> (+res) LN1' -> qkv'
// As separate launches (runtime.cu decode_step_large: 11 kernels per layer) every link costs a kernel boundary plus the
// prologue of a tcgen05 kernel (barrier init, TMEM allocation, descriptor fetch, pipeline fill): 8-14 us per link for 0.3 us
// of weight bytes (ncu, profiles/r01_ncu_decode_layer_v15.md) - 1.55 ms of a 7.9 ms step at B = 256 and 1.6 ms of a 2.5 ms
// step at B = 32 (the per-rank batch of the 8-GPU strong-scaling configuration).
//
// Here a whole chain is ONE persistent cooperative kernel (one CTA per SM, warp-specialised like gemm_tc.cu):
//   warp 0      TMA producer: weight tiles of the NEXT phase are requested BEFORE the grid barrier (they do not depend on it),
//               activation tiles right after it
//   warp 1      tcgen05.mma issuer (UMMA 128 x BN x 16, BN chosen per phase on the host, fp32 accumulators in TMEM, two
//               accumulator buffers so that the epilogue of one tile overlaps the main loop of the next)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue (tcgen05.ld -> transpose through swizzled smem -> bias / erf-GELU -> 128-byte row stores), and the
//               workers of the LayerNorm phases: x += bias + split-K slabs (fixed order, deterministic), LayerNorm, bf16 out
// The phases of a chain are separated by a grid barrier (release-add / acquire-poll on one L2 word) instead of a kernel
// boundary; the TMEM allocation, the mbarriers and the smem ring live across phases.  The kernel is an interpreter over a
// table of phase descriptors (tensor maps included) built once per (session, batch) on the host.
// Semantics and rounding points are exactly those of decode_step_large (fp32 residual stream, LayerNorm eps 1e-5 in fp32,
// bf16 activations into every Linear, split-K partials summed in slab order): models/whisper/model.py:321-369,
// layers/normalization.py:6-30, modeling_whisper.py:710-751.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wb_runtime.h"
#include "wb_ptx.cuh"

namespace wb {

CUtensorMap make_tmap_bf16_kgroups(const void* ptr, long long ld_elems, int rows, int cols, int box_rows, int box_k);   // gemm_tc.cu
std::atomic<long long*>& step_trace_ptr();                                                                            // step_mega.cu

namespace {
constexpr int CH_BM = 128, CH_BK = 64, CH_MAX_BN = 128;
// One stage of the ring holds a GROUP of kgroup consecutive k-blocks: [kgroup activation slices | kgroup weight slices], each
// fetched by ONE TMA operation (3-D tensor maps: {64 k, rows, k-block}).  The TMA unit of an SM completes roughly one operation
// per ~200 cycles whatever its box size (two per k-block made every k-loop run at 395 cycles per k-block, B and tile width
// regardless: profiles/r02_chain_trace_*_v3_uniform_issue.md), so the boxes are made as deep as a 48 KB stage allows.
constexpr int CH_STAGES = 4;
constexpr uint32_t CH_STAGE_BYTES = 48 * 1024;
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_THREADS = (4 + CH_EPI_WARPS) * 32;                // 384
constexpr uint32_t CH_STAGING_BYTES = CH_EPI_WARPS * 32 * 32 * 4;  // one 32x32 fp32 transpose tile per epilogue warp
constexpr uint32_t CH_TMEM_COLS = 2 * CH_MAX_BN;                   // two accumulator buffers
constexpr uint32_t CH_SMEM_BYTES = CH_STAGES * CH_STAGE_BYTES + CH_STAGING_BYTES + 512 /*barriers, LN scratch*/ + 1024 /*alignment*/;
constexpr int CH_MAX_PARTS = MAX_K_SPLITS;

enum { CH_GEMM = 0, CH_LN = 1, CH_SELF = 2, CH_CROSS = 3 };
enum { CH_EPI_PARTIAL = 0, CH_EPI_BF16 = 1, CH_EPI_GELU_BF16 = 2 };

// One phase of a chain.  GEMM: out = epi(A[M, K] W[N, K]^T) on 128 x bn tiles, k_splits slabs; LN: the consumer side of a
// split-K GEMM fused with the LayerNorm in front of the next Linear.
struct alignas(128) ChainPhase {
    CUtensorMap tmA;            // activations [M, K] bf16 as {64, M, K / 64}, box {64, a_rows, kgroup}
    CUtensorMap tmW;            // weights [N, K] bf16 as {64, N, K / 64}, box {64, bn, kgroup}
    int kind;
    int bn, K, N, n_tiles_n, k_splits, epi;
    int n_parts;                // LN: slabs to add to x (0: plain LayerNorm of x)
    int a_bytes;                // GEMM: bytes of one activation box ({64, a_rows, kgroup}: only the rows that exist are loaded)
    int kgroup;                 // GEMM: k-blocks per TMA operation / ring stage
    int w_off;                  // GEMM: byte offset of the weight slices inside a stage (= a_bytes)
    long long split_stride;     // GEMM (partial epilogue) / LN: floats between consecutive slabs
    long long ldo;              // GEMM: elements between output rows
    const float* bias;          // GEMM: epilogue bias (bf16 epilogues); LN: bias of the split-K GEMM that produced the slabs
    void* out;                  // GEMM: fp32 slabs or bf16 rows; LN: bf16 normalised rows [M, d]
    const float* parts;         // LN: slabs
    const float* gamma; const float* beta;
    float* x;                   // LN: fp32 residual stream [M, d], updated in place
    // attention phases (small batches: the attention kernels become phases of the layer's launch).  SELF: this layer's K / V pages;
    // CROSS: this layer's K / V of the encoder frames; q of CROSS = bias + n_parts slabs in `parts` (the cross-q GEMM's split-K output)
    const bf16* kbase; const bf16* vbase;
};
static_assert(sizeof(ChainPhase) == 384, "ChainPhase layout");

constexpr int CH_MAX_PHASES = 11;  // phases of one launch (a whole decoder layer with its two attention phases): kernel parameters

// The phase descriptors of ONE launch live in the kernel parameters (constant bank): every per-phase scalar is a uniform
// load, the loops they control are provably warp-uniform, and the tensor maps sit where the TMA unit fetches them fastest.
// (A table in global memory cost a cold ~1.5 us descriptor + tensor-map fetch per launch and, worse, made ptxas treat the
// k-loop state as divergent: every tcgen05.mma / TMA operand was moved lane -> uniform register one by one, ~100 cycles per
// instruction, profiles/r02_chain_trace_*.md.)
struct ChainParams {
    ChainPhase ph[CH_MAX_PHASES];
    int n_ph;                   // phases of this launch
    int ph0;                    // index of ph[0] in the step's phase list (trace slots only)
    int M, d;
    float eps;
    // attention phases
    int H, n_ctx, pages_per_seq;
    long long cross_bstride;    // elements between the cross K (or V) of consecutive utterances
    const bf16* qkv;            // fused q | k | v rows of the step [M, 3 d] (SELF: q, and the row to append)
    bf16* attn_out;             // [M, d]
    const int* unfinished; const int* row_len; const int* page_table;
    const StepState* state;
    unsigned* sync;             // [0] grid-barrier arrivals, [1] exits: zero between launches
    long long* trace;           // optional (tools/chain_trace.py): SM clock stamps of CTA 0, 8 slots per phase
};

__device__ __forceinline__ unsigned ch_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ch_red_release(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes (st.global by other CTAs) -> async-proxy reads (TMA) of the same global addresses
__device__ __forceinline__ void ch_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float4 ch_ld_cg_f4(const float* p) {   // written by other CTAs of this kernel: read at L2
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
// Converged-warp TMA forms (see ptx::umma_f16_elect): every lane executes the call with warp-uniform operands, the elected
// lane issues; operands stay in uniform registers.
__device__ __forceinline__ void ch_tma_load_3d_elect(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// TMA prefetch of one box into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void ch_tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void ch_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ch_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void ch_named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ bool ch_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait on an mbarrier given by its shared-memory address (a protocol bug traps instead of hanging the GPU)
__device__ __forceinline__ void ch_mbar_wait(uint32_t bar, uint32_t parity) {
    if (ch_mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!ch_mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            printf("wb: chain mbarrier timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// every CTA has arrived `target` times: all writes of the phases before are visible.  A lost CTA traps instead of hanging.
__device__ __forceinline__ void ch_grid_poll(const unsigned* counter, unsigned target) {
    if (ch_ld_acquire(counter) >= target) return;
    const long long t0 = clock64();
    while (ch_ld_acquire(counter) < target) {
        if (clock64() - t0 > 4000000000ll) {
            printf("wb: chain grid barrier timeout (block %d thread %d target %u)\n", blockIdx.x, threadIdx.x, target);
            __trap();
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// Attention phases (batches whose (utterance, head) items fit the worker warps of one launch: B x H <= 8 x SMs).  At these
// sizes the two attention kernels of a layer were launch-latency bound (13 + 35 us of kernel for 4 + 28 us of bytes at B = 32,
// plus a kernel boundary each); as phases of the layer's launch they cost a grid barrier instead.  One warp PAIR per item, the
// keys interleaved between the two warps in blocks, merged through shared memory; 8 lanes per key row, 16-byte loads, online
// softmax per lane group (same arithmetic as attn_dec.cu: fp32 scores and statistics, __expf, q pre-scaled in the weights).
// ------------------------------------------------------------------------------------------------------------------------
struct ChainAttnAcc {
    float m, l, acc[8];
};

__device__ __forceinline__ void ch_attn_block(ChainAttnAcc& st, const float (&qf)[8], const Vec16<bf16>* kr, const Vec16<bf16>* vr,
                                              int first_key, int key_step, int n, int nu) {
    float sc[8];
    float mb = -INFINITY;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < nu) {
            float kf[8];
            kr[u].unpack(kf);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[i], dot);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            sc[u] = (first_key + u * key_step < n) ? dot : -INFINITY;
            mb = fmaxf(mb, sc[u]);
        }
    }
    const float m_new = fmaxf(st.m, mb);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float scale = __expf(st.m - m_use);
    st.l *= scale;
#pragma unroll
    for (int i = 0; i < 8; ++i) st.acc[i] *= scale;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (u < nu) {
            const float pr = __expf(sc[u] - m_use);
            st.l += pr;
            float vf[8];
            vr[u].unpack(vf);
#pragma unroll
            for (int i = 0; i < 8; ++i) st.acc[i] = fmaf(pr, vf[i], st.acc[i]);
        }
    }
    st.m = m_new;
}

// merge the four 8-lane key groups of a warp (lanes with equal `sub`): afterwards lanes 0-7 hold the warp's state
__device__ __forceinline__ void ch_attn_merge_groups(ChainAttnAcc& st) {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, st.m, o);
        const float ol = __shfl_xor_sync(0xffffffffu, st.l, o);
        const float mm = fmaxf(st.m, om);
        const float s1 = (st.m == -INFINITY) ? 0.f : __expf(st.m - mm);
        const float s2 = (om == -INFINITY) ? 0.f : __expf(om - mm);
        st.l = st.l * s1 + ol * s2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float oa = __shfl_xor_sync(0xffffffffu, st.acc[i], o);
            st.acc[i] = st.acc[i] * s1 + oa * s2;
        }
        st.m = mm;
    }
}

// the second warp of a pair hands its state over through shared memory, the first merges (fixed order) and writes the row
__device__ __forceinline__ void ch_attn_finish(ChainAttnAcc& st, float* pair_smem, int half, int pair_in_cta, int lane, bf16* out_row) {
    const int sub = lane & 7;
    if (half == 1 && lane < 8) {
        if (sub == 0) { pair_smem[0] = st.m; pair_smem[1] = st.l; }
#pragma unroll
        for (int i = 0; i < 8; ++i) pair_smem[8 + sub * 8 + i] = st.acc[i];
    }
    ch_named_bar(4 + pair_in_cta, 64);
    if (half == 0 && lane < 8) {
        const float om = pair_smem[0], ol = pair_smem[1];
        const float mm = fmaxf(st.m, om);
        const float s1 = (st.m == -INFINITY) ? 0.f : __expf(st.m - mm);
        const float s2 = (om == -INFINITY) ? 0.f : __expf(om - mm);
        const float inv = 1.0f / (st.l * s1 + ol * s2);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (st.acc[i] * s1 + pair_smem[8 + sub * 8 + i] * s2) * inv;
        Vec16<bf16> v;
        v.pack(o);
        st16(out_row + sub * 8, v);
    }
}

// cached self-attention of the step: append the new K / V row in place, attend over the row's length
__device__ __forceinline__ void ch_self_attention(const ChainParams& p, const ChainPhase& D, int wwarp, int lane, float* scratch) {
    const int sub = lane & 7, grp = lane >> 3;
    const int half = wwarp & 1, pair_in_cta = wwarp >> 1;
    const int items = p.M * p.H, d = p.d;
    int parity = 0;
    for (int item = (int)blockIdx.x * 4 + pair_in_cta; item < items; item += (int)gridDim.x * 4) {   // pair-uniform trip count
        const int b = item / p.H, h = item - b * p.H;
        if (p.unfinished[b] == 0) continue;                  // finished utterance (pair-uniform)
        const int n = p.row_len != nullptr ? p.row_len[b] : p.state->cur_len;
        float* pair_smem = scratch + (pair_in_cta * 2 + parity) * 72;
        parity ^= 1;
        const int my_page = lane < p.pages_per_seq ? p.page_table[(size_t)b * p.pages_per_seq + lane] : 0;
        const bf16* qrow = p.qkv + (size_t)b * 3 * d + h * 64 + sub * 8;
        float qf[8];
        ld16(qrow).unpack(qf);
        const Vec16<bf16> k_new = ld16(qrow + d), v_new = ld16(qrow + 2 * d);
        auto row_off = [&](int s) -> size_t {
            const int page = __shfl_sync(0xffffffffu, my_page, s >> 6);
            return (((size_t)page * p.H + h) * 64 + (s & 63)) * 64 + sub * 8;
        };
        {   // in-place append at slot n - 1 (this step reads that row from the registers above)
            const size_t off = row_off(n - 1);
            if (half == 0 && grp == 0) {
                st16(const_cast<bf16*>(D.kbase) + off, k_new);
                st16(const_cast<bf16*>(D.vbase) + off, v_new);
            }
        }
        ChainAttnAcc st;
        st.m = -INFINITY; st.l = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) st.acc[i] = 0.f;
        for (int sb = half * 16; sb < n; sb += 32) {         // blocks of 16 keys, alternating between the two warps
            Vec16<bf16> kr[8], vr[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = min(sb + grp + u * 4, n - 1);
                const size_t off = row_off(s);
                kr[u] = ld16_issue(D.kbase + off);
                vr[u] = ld16_issue(D.vbase + off);
                if (s == n - 1) { kr[u] = k_new; vr[u] = v_new; }
            }
            ch_attn_block(st, qf, kr, vr, sb + grp, 4, n, 4);
        }
        ch_attn_merge_groups(st);
        ch_attn_finish(st, pair_smem, half, pair_in_cta, lane, p.attn_out + (size_t)b * d + h * 64);
    }
}

// cross-attention over the encoder K / V projected once per utterance; q = bias + split-K slabs of the cross-q GEMM
__device__ __forceinline__ void ch_cross_attention(const ChainParams& p, const ChainPhase& D, int wwarp, int lane, float* scratch) {
    const int sub = lane & 7, grp = lane >> 3;
    const int half = wwarp & 1, pair_in_cta = wwarp >> 1;
    const int items = p.M * p.H, d = p.d, n = p.n_ctx;
    int parity = 0;
    for (int item = (int)blockIdx.x * 4 + pair_in_cta; item < items; item += (int)gridDim.x * 4) {
        const int b = item / p.H, h = item - b * p.H;
        if (p.unfinished[b] == 0) continue;
        float* pair_smem = scratch + (pair_in_cta * 2 + parity) * 72;
        parity ^= 1;
        float qf[8];
        {
            const int c0 = h * 64 + sub * 8;
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
                float4 t = D.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(D.bias + c0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = 0; s < D.n_parts; ++s) {
                    const float4 u = ch_ld_cg_f4(D.parts + (size_t)s * D.split_stride + (size_t)b * d + c0 + i);
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                qf[i] = t.x; qf[i + 1] = t.y; qf[i + 2] = t.z; qf[i + 3] = t.w;
            }
        }
        const size_t base = (size_t)b * p.cross_bstride + (size_t)h * n * 64 + sub * 8;
        const bf16* kb = D.kbase + base;
        const bf16* vb = D.vbase + base;
        ChainAttnAcc st;
        st.m = -INFINITY; st.l = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) st.acc[i] = 0.f;
        for (int sb = half * 32; sb < n; sb += 64) {         // blocks of 32 keys, alternating between the two warps
            Vec16<bf16> kr[8], vr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const size_t off = (size_t)min(sb + grp + u * 4, n - 1) * 64;
                kr[u] = ld16_stream(kb + off);
                vr[u] = ld16_stream(vb + off);
            }
            ch_attn_block(st, qf, kr, vr, sb + grp, 4, n, 8);
        }
        ch_attn_merge_groups(st);
        ch_attn_finish(st, pair_smem, half, pair_in_cta, lane, p.attn_out + (size_t)b * d + h * 64);
    }
}

__global__ void __launch_bounds__(CH_THREADS, 1) decode_chain_kernel(const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    if (p.state->active == 0) return;   // the loop has stopped: grid-uniform, nobody touches the barrier
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t skew = (1024u - (raw_addr & 1023u)) & 1023u;          // SWIZZLE_128B tiles need 1024-byte aligned bases
    uint8_t* smem = smem_raw + skew;
    const uint32_t smem_a = raw_addr + skew;                             // shared-memory address of the ring (warp-uniform)
    constexpr uint32_t BAR_OFF = CH_STAGES * CH_STAGE_BYTES + CH_STAGING_BYTES;
    float* staging = reinterpret_cast<float*>(smem + CH_STAGES * CH_STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* empty_bar = full_bar + CH_STAGES;
    uint64_t* tmem_full_bar = empty_bar + CH_STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* ln_red = reinterpret_cast<float*>(tmem_ptr_smem + 4);        // [2 groups][2 reductions][4 warps]
    // the same barriers by shared-memory address
    const uint32_t full_a = smem_a + BAR_OFF, empty_a = full_a + CH_STAGES * 8;
    const uint32_t tfull_a = empty_a + CH_STAGES * 8, tempty_a = tfull_a + 16;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform by construction
    const int lane = threadIdx.x & 31;
    // stamps (CTA 0, lane 0 of each role): [0] phase start, [1] epilogue leader past the grid barrier, [2] producer past it,
    // [3] activation loads issued, [4] first k-block landed, [5] first accumulator complete, [6] this CTA's share written,
    // [7] arrived at the next grid barrier
    const bool tracing = p.trace != nullptr && blockIdx.x == 0;         // uniform
    const int n_mt = (p.M + CH_BM - 1) / CH_BM;
    const int n_ph = p.n_ph;
    const unsigned grid = gridDim.x;
    unsigned* const bar = p.sync;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < CH_STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 2);    // the two producer warps
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], CH_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<CH_TMEM_COLS>(tmem_ptr_smem);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer of the ACTIVATION tiles (converged warp, elected lane issues) =====================
        // Two producer warps: a TMA instruction costs its issuing warp ~150-300 cycles whatever the box size, and a k-block of
        // these skinny tiles is worth ~150 cycles of MMA issue: one warp issuing both loads of a stage was the pace-maker of
        // every k-loop (394 cycles per k-block, profiles/r02_chain_trace_*.md).  Each warp posts its own bytes on the stage's
        // full barrier (initialised with two arrivals).
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_ph; ++i) {
            const ChainPhase& D = p.ph[i];
            if (D.kind != CH_GEMM) continue;
            const int nt = D.n_tiles_n, ksp = D.k_splits, kg = D.kgroup;
            const int num_tiles = n_mt * nt * ksp, nk = D.K / CH_BK / ksp, ng = nk / kg;
            const uint32_t a_bytes = (uint32_t)D.a_bytes;
            bool need_wait = i > 0;    // the activations were written by the previous phase of this launch
            for (int tile = blockIdx.x; tile < num_tiles; tile += grid) {
                const int mn = tile / ksp, ks = tile - mn * ksp;
                const int m_blk = mn / nt;
                if (need_wait) {
                    if (lane == 0) {
                        ch_grid_poll(bar, (unsigned)i * grid);
                        if (tracing) p.trace[8 * (p.ph0 + i) + 2] = clock64();
                        ch_fence_proxy_async_all();
                    }
                    __syncwarp();
                    need_wait = false;
                }
                for (int g = 0; g < ng; ++g) {
                    ch_mbar_wait(empty_a + stage * 8, phase ^ 1);
                    ch_expect_tx_elect(full_a + stage * 8, a_bytes);
                    ch_tma_load_3d_elect(smem_a + stage * CH_STAGE_BYTES, &D.tmA, full_a + stage * 8, 0, m_blk * CH_BM, ks * nk + g * kg);
                    if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
                }
                if (tracing && lane == 0 && tile == (int)blockIdx.x) p.trace[8 * (p.ph0 + i) + 3] = clock64();
            }
        }
    } else if (warp == 2) {
        // ===================== TMA producer of the WEIGHT tiles =====================
        // Weights do not depend on the previous phase: this warp never looks at the grid barrier, it runs ahead as far as the
        // ring lets it (the weight tiles of the next phase are in flight while the CTA waits for the others).
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_ph; ++i) {
            const ChainPhase& D = p.ph[i];
            if (D.kind != CH_GEMM) continue;
            const int bn = D.bn, nt = D.n_tiles_n, ksp = D.k_splits, kg = D.kgroup;
            const int num_tiles = n_mt * nt * ksp, nk = D.K / CH_BK / ksp, ng = nk / kg;
            const uint32_t w_bytes = (uint32_t)(bn * kg) * CH_BK * 2, w_off = (uint32_t)D.w_off;
            for (int tile = blockIdx.x; tile < num_tiles; tile += grid) {
                const int mn = tile / ksp, ks = tile - mn * ksp;
                const int n_blk = mn % nt;
                for (int g = 0; g < ng; ++g) {
                    ch_mbar_wait(empty_a + stage * 8, phase ^ 1);
                    ch_expect_tx_elect(full_a + stage * 8, w_bytes);
                    ch_tma_load_3d_elect(smem_a + stage * CH_STAGE_BYTES + w_off, &D.tmW, full_a + stage * 8, 0, n_blk * bn, ks * nk + g * kg);
                    if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== L2 prefetcher =====================
        // The weights of a decode step are read once per step and never hit L2 on their own (tens of GB of K/V stream through
        // it in between).  Every weight box this CTA will load in ANY phase of the launch is requested into L2 right away: the
        // launch streams its weights at HBM speed while the first phases run, the later k-loops run at L2 latency.
        // LayerNorm / bias vectors likewise.
        if (lane < n_ph && p.ph[lane].kind == CH_GEMM) {
            ptx::prefetch_tensormap(&p.ph[lane].tmA);
            ptx::prefetch_tensormap(&p.ph[lane].tmW);
        }
        for (int i = 0; i < n_ph; ++i) {
            const ChainPhase& D = p.ph[i];
            if (D.kind == CH_GEMM) {
                const int bn = D.bn, nt = D.n_tiles_n, ksp = D.k_splits, kg = D.kgroup;
                const int num_tiles = n_mt * nt * ksp, nk = D.K / CH_BK / ksp, ng = nk / kg;
                for (int tile = blockIdx.x; tile < num_tiles; tile += grid) {   // (row tiles sharing a weight box: an L2 hit)
                    const int mn = tile / ksp, ks = tile - mn * ksp;
                    const int n_blk = mn % nt;
                    for (int g = lane; g < ng; g += 32) ch_tma_prefetch_3d(&D.tmW, 0, n_blk * bn, ks * nk + g * kg);
                }
                if (D.bias != nullptr)
                    for (int j = lane * 32; j < D.N; j += 32 * 32) ch_prefetch_l2(D.bias + j);
            } else if (D.kind == CH_LN) {
                for (int j = lane * 32; j < p.d; j += 32 * 32) {
                    ch_prefetch_l2(D.gamma + j);
                    ch_prefetch_l2(D.beta + j);
                    if (D.bias != nullptr) ch_prefetch_l2(D.bias + j);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, elected lane issues) =====================
        // tcgen05.mma / tcgen05.commit take their operands from UNIFORM registers: the loop state below is derived from kernel
        // parameters and uniform counters only, so the descriptors are computed on the uniform datapath.
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = 0; i < n_ph; ++i) {
            const ChainPhase& D = p.ph[i];
            if (D.kind != CH_GEMM) continue;
            const int bn = D.bn, ksp = D.k_splits, kg = D.kgroup;
            const int num_tiles = n_mt * D.n_tiles_n * ksp, nk = D.K / CH_BK / ksp, ng = nk / kg;
            const uint32_t idesc = ptx::make_idesc_bf16(CH_BM, (uint32_t)bn, 0, 0);
            const uint32_t a_slice = (uint32_t)D.a_bytes / (uint32_t)kg, w_slice = (uint32_t)bn * CH_BK * 2, w_off = (uint32_t)D.w_off;
            for (int tile = blockIdx.x; tile < num_tiles; tile += grid) {
                ch_mbar_wait(tempty_a + acc * 8, acc_phase ^ 1);   // epilogue drained this accumulator
                ptx::tcgen05_fence_after();
                const uint32_t d_tmem = tmem_u + acc * CH_MAX_BN;
                for (int g = 0; g < ng; ++g) {
                    ch_mbar_wait(full_a + stage * 8, phase);
                    ptx::tcgen05_fence_after();
                    if (tracing && lane == 0) {
                        if (g == 0 && tile == (int)blockIdx.x) p.trace[8 * (p.ph0 + i) + 4] = clock64();
                        if (p.ph0 + i == 16 && g < 32) p.trace[4096 + 2 * g] = clock64();   // (dev) ring-stage timeline of fc1, layer 1
                    }
                    const uint32_t sa = smem_a + stage * CH_STAGE_BYTES;
                    for (int sl = 0; sl < kg; ++sl) {
                        const uint64_t da = ptx::make_smem_desc_sw128(sa + sl * a_slice, 1024, 16);
                        const uint64_t db = ptx::make_smem_desc_sw128(sa + w_off + sl * w_slice, 1024, 16);
                        ptx::umma_f16_elect(d_tmem, da, db, idesc, (g | sl) != 0);
                        ptx::umma_f16_elect(d_tmem, da + 2, db + 2, idesc, 1);
                        ptx::umma_f16_elect(d_tmem, da + 4, db + 4, idesc, 1);
                        ptx::umma_f16_elect(d_tmem, da + 6, db + 6, idesc, 1);
                    }
                    ptx::umma_commit_elect(empty_a + stage * 8);
                    if (g == ng - 1) ptx::umma_commit_elect(tfull_a + acc * 8);
                    if (tracing && lane == 0 && p.ph0 + i == 16 && g < 32) p.trace[4096 + 2 * g + 1] = clock64();
                    if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue / LayerNorm workers =====================
        const int et = threadIdx.x - 128;       // 0..255
        const int q = warp & 3;                 // TMEM lane quarter of this warp
        const int half = (warp - 4) >> 2;       // column-slab share (GEMM) / row group (LN)
        float4* st4 = reinterpret_cast<float4*>(staging + (warp - 4) * 32 * 32);
        const int rrow = lane >> 3, rchunk = lane & 7;
        long long* const trace = (tracing && et == 0) ? p.trace + 8 * p.ph0 : nullptr;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = 0; i < n_ph; ++i) {
            const ChainPhase& D = p.ph[i];
            // The barrier word is ONE monotonic counter: it is only meaningful if no CTA arrives for phase i before every CTA
            // has arrived for phase i - 1.  So the thread that arrives for this CTA waits for the previous barrier at the start
            // of EVERY phase, also in CTAs that have no tile of it (they must not run ahead and be counted twice).
            if (trace != nullptr) trace[8 * i + 0] = clock64();
            if (i > 0 && et == 0) ch_grid_poll(bar, (unsigned)i * grid);
            if (trace != nullptr) trace[8 * i + 1] = clock64();
            if (D.kind == CH_GEMM) {
                const int bn = D.bn, nt = D.n_tiles_n, ksp = D.k_splits, N = D.N, epi = D.epi;
                const int num_tiles = n_mt * nt * ksp;
                const int slabs = bn >> 5, spw = (slabs + 1) >> 1;
                const float* __restrict__ bias = D.bias;
                const long long ldo = D.ldo, sstride = D.split_stride;
                void* const outp = D.out;
                for (int tile = blockIdx.x; tile < num_tiles; tile += grid) {
                    const int mn = tile / ksp, ks = tile - mn * ksp;
                    const int m_blk = mn / nt, n_blk = mn - m_blk * nt;
                    ch_mbar_wait(tfull_a + acc * 8, acc_phase);
                    ptx::tcgen05_fence_after();
                    if (trace != nullptr && tile == (int)blockIdx.x) trace[8 * i + 5] = clock64();
                    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * CH_MAX_BN;
                    const int row0 = m_blk * CH_BM + q * 32;
#pragma unroll 1
                    for (int ci = 0; ci < spw; ++ci) {
                        const int c = half * spw + ci;
                        if (c >= slabs) break;
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(t_row + c * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        __syncwarp();
                        const int n0 = n_blk * bn + c * 32 + rchunk * 4;
                        const bool col_ok = n0 < N;
                        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (epi != CH_EPI_PARTIAL && bias != nullptr && col_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + n0));
#pragma unroll
                        for (int r8 = 0; r8 < 8; ++r8) {
                            const int rr = r8 * 4 + rrow;
                            const int row = row0 + rr;
                            const float4 f = st4[rr * 8 + (rchunk ^ (rr & 7))];
                            if (row < p.M && col_ok) {
                                if (epi == CH_EPI_PARTIAL) {
                                    float* o = reinterpret_cast<float*>(outp) + (long long)ks * sstride + (long long)row * ldo + n0;
                                    *reinterpret_cast<float4*>(o) = f;
                                } else {
                                    float o0 = f.x + bb.x, o1 = f.y + bb.y, o2 = f.z + bb.z, o3 = f.w + bb.w;
                                    if (epi == CH_EPI_GELU_BF16) {
                                        o0 = gelu_erf_fast(o0); o1 = gelu_erf_fast(o1); o2 = gelu_erf_fast(o2); o3 = gelu_erf_fast(o3);
                                    }
                                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
                                    uint2 u;
                                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                                    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(outp) + (long long)row * ldo + n0) = u;
                                }
                            }
                        }
                        __syncwarp();   // the staging tile is rewritten by the next slab
                    }
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            } else if (D.kind == CH_SELF || D.kind == CH_CROSS) {
                if (i > 0) ch_named_bar(1, CH_EPI_WARPS * 32);   // thread 0 of the group has seen the barrier (above)
                if (D.kind == CH_SELF) ch_self_attention(p, D, warp - 4, lane, staging);
                else ch_cross_attention(p, D, warp - 4, lane, staging);
            } else {
                // ---- LayerNorm phase: x[r] += bias + sum of slabs (fixed order); out[r] = LN(x[r]) as bf16
                if (i > 0) ch_named_bar(1, CH_EPI_WARPS * 32);   // thread 0 of the group has seen the barrier (above)
                const int d = p.d, nvec = d >> 2;
                const int gt = et & 127, gw = gt >> 5;              // thread / warp inside the 128-thread row group
                const int n_parts = D.n_parts;
                const long long pstride = D.split_stride;
                float* const xp = D.x;
                const float* const partp = D.parts;
                const float4* const biasp = reinterpret_cast<const float4*>(D.bias);
                const float4* const gammap = reinterpret_cast<const float4*>(D.gamma);
                const float4* const betap = reinterpret_cast<const float4*>(D.beta);
                bf16* const lnout = reinterpret_cast<bf16*>(D.out);
                float* red0 = ln_red + half * 8;
                float* red1 = red0 + 4;
                const float inv_d = 1.0f / (float)d;
                for (int row = (int)blockIdx.x * 2 + half; row < p.M; row += (int)grid * 2) {   // group-uniform trip count
                    float4 v[2], g[2], be[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int t = gt + j * 128;
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        g[j] = v[j]; be[j] = v[j];
                        if (t < nvec) {
                            float* xr = xp + (size_t)row * d + t * 4;
                            v[j] = ch_ld_cg_f4(xr);
                            if (n_parts > 0) {
                                float4 b = biasp != nullptr ? __ldg(biasp + t) : make_float4(0.f, 0.f, 0.f, 0.f);
                                float4 pq[CH_MAX_PARTS];
#pragma unroll
                                for (int s = 0; s < CH_MAX_PARTS; ++s)   // all slabs in flight together
                                    pq[s] = s < n_parts ? ch_ld_cg_f4(partp + (size_t)s * pstride + (size_t)row * d + t * 4)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                for (int s = 0; s < CH_MAX_PARTS; ++s) { b.x += pq[s].x; b.y += pq[s].y; b.z += pq[s].z; b.w += pq[s].w; }
                                v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;
                                *reinterpret_cast<float4*>(xr) = v[j];
                            }
                            g[j] = __ldg(gammap + t);
                            be[j] = __ldg(betap + t);
                        }
                    }
                    float s = warp_sum(((v[0].x + v[0].y) + (v[0].z + v[0].w)) + ((v[1].x + v[1].y) + (v[1].z + v[1].w)));
                    if (trace != nullptr) trace[8 * i + 3] = clock64();     // (LN phases) loads landed, row sum of this warp
                    if (lane == 0) red0[gw] = s;
                    ch_named_bar(2 + half, 128);
                    const float mean = ((red0[0] + red0[1]) + (red0[2] + red0[3])) * inv_d;
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (gt + j * 128 < nvec) {
                            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
                            ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
                        }
                    }
                    ss = warp_sum(ss);
                    if (lane == 0) red1[gw] = ss;
                    ch_named_bar(2 + half, 128);
                    const float rstd = rsqrtf(((red1[0] + red1[1]) + (red1[2] + red1[3])) * inv_d + p.eps);
                    if (trace != nullptr) trace[8 * i + 4] = clock64();     // statistics done
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int t = gt + j * 128;
                        if (t < nvec) {
                            const float y0 = v[j].x * rstd * g[j].x + be[j].x, y1 = v[j].y * rstd * g[j].y + be[j].y;
                            const float y2 = v[j].z * rstd * g[j].z + be[j].z, y3 = v[j].w * rstd * g[j].w + be[j].w;
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
                            uint2 u;
                            u.x = *reinterpret_cast<uint32_t*>(&p0);
                            u.y = *reinterpret_cast<uint32_t*>(&p1);
                            *reinterpret_cast<uint2*>(lnout + (size_t)row * d + t * 4) = u;
                        }
                    }
                }
            }
            // ---- this CTA's share of the phase is written: arrive at the grid barrier (the last phase needs none)
            if (trace != nullptr) trace[8 * i + 6] = clock64();
            if (i + 1 < n_ph) {
                ch_named_bar(1, CH_EPI_WARPS * 32);
                if (et == 0) {
                    // bar.sync ordered the group's writes before this thread; the release below is cumulative at gpu scope
                    ch_fence_proxy_async_all();   // the next phase reads these rows through TMA
                    ch_red_release(bar, 1u);
                    if (trace != nullptr) trace[8 * i + 7] = clock64();
                }
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<CH_TMEM_COLS>(tmem_base);
    // leave the counters at zero for the next launch: the last CTA to get here resets them (nobody polls any more: a CTA gets
    // here only after all of its warps have passed every poll of the launch)
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(bar + 1, 1u);   // exit counter: separate word, exits must not be mistaken for arrivals
        if (old == grid - 1) {
            atomicExch(bar, 0u);
            atomicExch(bar + 1, 0u);
        }
    }
}

// tile width / split-K of one chain GEMM: same cost model as gemm_tc.cu pick_config (skinny M: bytes a CTA has to pull, one
// wave, the partial slabs its consumer re-reads), tile width limited to CH_MAX_BN
void pick_chain_config(int M, int a_rows, int N, int K, int max_splits, int sms, int& bn_out, int& splits_out) {
    const int mt = ceil_div(M, CH_BM), nkb = K / CH_BK;
    double best = 1e30;
    bn_out = 32; splits_out = 1;
    // (dev) WB_CHAIN_BN / WB_CHAIN_SPLITS: force the tile width / cap the split count
    static const int force_bn = std::getenv("WB_CHAIN_BN") ? std::atoi(std::getenv("WB_CHAIN_BN")) : 0;
    static const int cap_splits = std::getenv("WB_CHAIN_SPLITS") ? std::atoi(std::getenv("WB_CHAIN_SPLITS")) : 0;
    if (cap_splits > 0) max_splits = std::min(max_splits, cap_splits);
    for (int bn : {32, 64, 128}) {
        if (force_bn > 0 && bn != force_bn) continue;
        for (int s = 1; s <= max_splits; s *= 2) {
            if (nkb % s != 0) continue;
            const long long ctas = (long long)mt * ceil_div(N, bn) * s;
            const double waves = (double)((ctas + sms - 1) / sms);
            const double ingest = (double)(a_rows + bn) * (K / s) * 2.0;
            const double partial = s > 1 ? (double)M * N * 4.0 * s * 2.0 / 6000.0 : 0.0;
            const double t = waves * (ingest / 40.0 + 1500.0) + partial;
            if (t < best) { best = t; bn_out = bn; splits_out = s; }
        }
    }
}
}  // namespace

size_t chain_table_bytes(int dec_layers) { return (size_t)(2 + 9 * dec_layers) * sizeof(ChainPhase); }   // host table
size_t chain_sync_bytes() { return 256; }

static std::atomic<bool>& chain_enabled() {
    static std::atomic<bool> on{true};
    return on;
}
void set_chain_path(bool on) { chain_enabled() = on; }
bool chain_path_enabled() { return chain_enabled(); }

bool Session::chain_supported() const {
    const ModelConfig& g = m->cfg;
    return chain_sync != nullptr && m->dtype == BF16 && batch >= 1 && g.d_model % 64 == 0 && g.d_model <= 1024 && g.ffn % 64 == 0 &&
           chain_grid > 0;
}

// The phase table of this session for the current batch.  Phase indices:
//   0 LN1(0), 1 qkv(0); layer l at base = 2 + 9 l: +0 out-proj (slabs), +1 LN2, +2 cross-q (slabs) | +3 cross-out (slabs), +4 LN3,
//   +5 fc1 (GELU), +6 fc2 (slabs), +7 LN1 of layer l + 1 (final LayerNorm after the last layer), +8 qkv of layer l + 1
void Session::build_chain_table() {
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, B = batch, L = g.dec_layers;
    WB_REQUIRE(chain_grid > 0, "fused chains are not available on this device");
    chain_host.assign(chain_table_bytes(L), 0);
    ChainPhase* t = reinterpret_cast<ChainPhase*>(chain_host.data());
    const long long part_stride = (long long)B * d;
    // Only the activation rows that exist travel: a CTA's k-loop is bound by the boxes it pulls (~600 cycles per 24 KB stage
    // with every SM pulling: the chip-wide L2 throughput cap, tools/chain_trace.py), and rows past the batch are never stored.
    // (The MMA still covers 128 rows: the rest of the tile holds stale shared memory, rows are independent.)
    const int a_rows = B >= CH_BM ? CH_BM : (B + 7) / 8 * 8;
    auto gemm_phase = [&](ChainPhase& o, const void* A, int K, const Linear& l, int epi, void* out, long long ldo) -> int {
        WB_REQUIRE(l.k == K && K % CH_BK == 0 && l.n % 8 == 0, "fused chains: unsupported Linear shape");
        int bn = 64, splits = 1;
        pick_chain_config(B, a_rows, l.n, K, epi == CH_EPI_PARTIAL ? MAX_K_SPLITS : 1, chain_grid, bn, splits);
        o.kind = CH_GEMM; o.bn = bn; o.K = K; o.N = l.n; o.n_tiles_n = ceil_div(l.n, bn); o.k_splits = splits; o.epi = epi;
        o.split_stride = part_stride; o.ldo = ldo; o.bias = epi == CH_EPI_PARTIAL ? nullptr : l.b; o.out = out;
        // k-blocks per TMA operation: as many as a stage holds, dividing the k-blocks of a tile
        const int nk = K / CH_BK / splits;
        int kg = 1;
        for (int c : {2, 4})
            if (nk % c == 0 && (size_t)c * (a_rows + bn) * CH_BK * 2 <= CH_STAGE_BYTES) kg = c;
        WB_REQUIRE((size_t)kg * (a_rows + bn) * CH_BK * 2 <= CH_STAGE_BYTES, "fused chains: tile does not fit a ring stage");
        o.kgroup = kg;
        o.tmA = make_tmap_bf16_kgroups(A, K, B, K, a_rows, kg);
        o.a_bytes = kg * a_rows * CH_BK * 2;
        o.w_off = o.a_bytes;
        o.tmW = make_tmap_bf16_kgroups(l.w, l.k, l.n, l.k, bn, kg);
        return splits;
    };
    auto ln_phase = [&](ChainPhase& o, const LNorm& n, int n_parts, const float* bias) {
        o.kind = CH_LN; o.n_parts = n_parts; o.split_stride = part_stride; o.bias = bias; o.parts = dpart;
        o.gamma = n.g; o.beta = n.b; o.x = dx; o.out = dln;
    };
    chain_q_parts.assign((size_t)L, 1);
    ln_phase(t[0], m->dec[0].ln1, 0, nullptr);
    gemm_phase(t[1], dln, d, m->dec[0].qkv, CH_EPI_BF16, dqkv, 3 * d);
    for (int l = 0; l < L; ++l) {
        const DecLayer& Y = m->dec[l];
        ChainPhase* o = &t[(size_t)2 + 9 * l];
        int s = gemm_phase(o[0], datt, d, Y.out, CH_EPI_PARTIAL, dpart, d);
        ln_phase(o[1], Y.ln2, s, Y.out.b);
        chain_q_parts[l] = gemm_phase(o[2], dln, d, Y.cq, CH_EPI_PARTIAL, dpart, d);
        s = gemm_phase(o[3], datt, d, Y.cout, CH_EPI_PARTIAL, dpart, d);
        ln_phase(o[4], Y.ln3, s, Y.cout.b);
        gemm_phase(o[5], dln, d, Y.fc1, CH_EPI_GELU_BF16, dffn, g.ffn);
        s = gemm_phase(o[6], dffn, g.ffn, Y.fc2, CH_EPI_PARTIAL, dpart, d);
        if (l + 1 < L) {
            ln_phase(o[7], m->dec[l + 1].ln1, s, Y.fc2.b);
            gemm_phase(o[8], dln, d, m->dec[l + 1].qkv, CH_EPI_BF16, dqkv, 3 * d);
        } else {
            ln_phase(o[7], m->dec_ln, s, Y.fc2.b);
        }
    }
    // Small batches, opt-in (see chain_merge_ok): the two attention kernels of a layer become phases of ONE launch per layer
    //   SELF, out-proj, LN2, cross-q, CROSS, cross-out, LN3, fc1, fc2, LN1' (final LayerNorm), qkv'
    chain_merged.clear();
    if (chain_merge_ok()) {
        std::vector<ChainPhase> mlist;
        mlist.push_back(t[0]);
        mlist.push_back(t[1]);
        const size_t per_kv = (size_t)max_batch * g.n_heads * g.n_ctx * 64;
        for (int l = 0; l < L; ++l) {
            const ChainPhase* o = &t[(size_t)2 + 9 * l];
            ChainPhase a;
            std::memset(&a, 0, sizeof(a));
            a.kind = CH_SELF;
            a.kbase = (const bf16*)self_k + (size_t)l * self_layer_elems();
            a.vbase = (const bf16*)self_v + (size_t)l * self_layer_elems();
            mlist.push_back(a);
            for (int k = 0; k < 3; ++k) mlist.push_back(o[k]);
            std::memset(&a, 0, sizeof(a));
            a.kind = CH_CROSS;
            a.kbase = (const bf16*)cross + (size_t)l * cross_layer_elems();
            a.vbase = a.kbase + per_kv;
            a.parts = dpart; a.n_parts = chain_q_parts[l]; a.split_stride = part_stride; a.bias = m->dec[l].cq.b;
            mlist.push_back(a);
            for (int k = 3; k < (l + 1 < L ? 9 : 8); ++k) mlist.push_back(o[k]);
        }
        chain_merged.resize(mlist.size() * sizeof(ChainPhase));
        std::memcpy(chain_merged.data(), mlist.data(), chain_merged.size());
    }
    chain_batch = B;
}

// the attention phases need one warp pair per (utterance, head) item in at most two rounds
// OPT-IN (wb_session_set_option "merge_attention" = 1, or WB_CHAIN_MERGE=1): measured SLOWER than the separate attention kernels -
// us per step at t = 128, same call: B = 17 2359 vs 1826, B = 32 2496 vs 2208, B = 64 3714 vs 2982 (profiles/r02_kernel_variants.md).
// One CTA of 384 threads per SM gives the attention phases 8 warps x 8 KB of loads in flight per SM where the stand-alone kernels
// keep 16 warps x 8 KB: the launch gaps it removes (~10 us per layer) are smaller than what the streaming loses.
bool Session::chain_merge_ok() const {
    static const bool env_on = std::getenv("WB_CHAIN_MERGE") != nullptr;   // (dev)
    return (opt_merge_attention == 1 || (opt_merge_attention < 0 && env_on)) && chain_grid > 0 && pages_per_seq <= 32 &&
           batch * m->cfg.n_heads <= 8 * chain_grid;
}

void Session::init_chain() {
    chain_grid = 0;
    chain_batch = -1;
    if (m->dtype != BF16 || chain_sync == nullptr) return;
    int dev = 0, sms = 0;
    WB_CHECK_CUDA(cudaGetDevice(&dev));
    WB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WB_CHECK_CUDA(cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM_BYTES));
    int per_sm = 0;
    WB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_chain_kernel, CH_THREADS, CH_SMEM_BYTES));
    if (per_sm < 1) return;   // does not fit on this device: the multi-kernel step stays
    chain_grid = sms;
    WB_CHECK_CUDA(cudaMemset(chain_sync, 0, chain_sync_bytes()));
}

void Session::launch_chain(int ph_begin, int ph_end, cudaStream_t st, bool merged) {
    static const bool serial = std::getenv("WB_CHAIN_SERIAL") != nullptr;   // (dev) one launch per phase: kernel boundaries instead of grid barriers
    if (serial && ph_end - ph_begin > 1) {
        for (int ph = ph_begin; ph < ph_end; ++ph) launch_chain(ph, ph + 1, st, merged);
        return;
    }
    const std::vector<unsigned char>& table = merged ? chain_merged : chain_host;
    WB_REQUIRE(ph_end > ph_begin && ph_end - ph_begin <= CH_MAX_PHASES && (size_t)ph_end * sizeof(ChainPhase) <= table.size(),
               "fused chains: bad phase range");
    const ModelConfig& g = m->cfg;
    ChainParams p;
    std::memset(&p, 0, sizeof(p));
    std::memcpy(p.ph, table.data() + (size_t)ph_begin * sizeof(ChainPhase), (size_t)(ph_end - ph_begin) * sizeof(ChainPhase));
    p.n_ph = ph_end - ph_begin; p.ph0 = ph_begin; p.M = batch; p.d = g.d_model; p.eps = 1e-5f;
    p.H = g.n_heads; p.n_ctx = g.n_ctx; p.pages_per_seq = pages_per_seq;
    p.cross_bstride = (long long)g.n_heads * g.n_ctx * 64;
    p.qkv = (const bf16*)dqkv; p.attn_out = (bf16*)datt;
    p.unfinished = unfinished; p.row_len = ragged_len(); p.page_table = page_table;
    p.state = state; p.sync = chain_sync; p.trace = step_trace_ptr();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(chain_grid);
    cfg.blockDim = dim3(CH_THREADS);
    cfg.dynamicSmemBytes = CH_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident, or the launch fails: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    WB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_chain_kernel, p));
}

// one token for the whole batch: 4 launches per layer (self-attention, chain, cross-attention, chain) instead of 11
void Session::decode_step_chain(cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, dt = m->dtype, B = batch, L = g.dec_layers;
    WB_REQUIRE(chain_batch == B, "fused chains: the phase table was built for another batch (prepare_step was not called)");
    const int* active = &state->active;
    if (!chain_merged.empty()) {
        // small batches: one launch per decoder layer, the attention kernels are phases of it (2 + L launches + LM head + argmax)
        { ProfScope ps(this, PROF_DEC_GEMM, st); launch_chain(0, 2, st, true); }
        int ph = 2;
        for (int l = 0; l < L; ++l) {
            const int n = l + 1 < L ? 11 : 10;
            ProfScope ps(this, PROF_DEC_GEMM, st);
            launch_chain(ph, ph + n, st, true);
            ph += n;
        }
    } else {
    { ProfScope ps(this, PROF_DEC_GEMM, st); launch_chain(0, 2, st); }
    for (int l = 0; l < L; ++l) {
        const int base = 2 + 9 * l;
        {
            DecAttnArgs a;
            a.dtype = dt; a.q = dqkv; a.q_stride = 3 * d; a.out = datt; a.out_stride = d; a.B = B; a.H = g.n_heads;
            a.state = state; a.row_active = unfinished; a.row_len = ragged_len();
            a.k_new = (uint8_t*)dqkv + (size_t)d * 2; a.v_new = (uint8_t*)dqkv + (size_t)2 * d * 2; a.new_stride = 3 * d;
            a.k_pages = (uint8_t*)self_k + (size_t)l * self_layer_elems() * 2;
            a.v_pages = (uint8_t*)self_v + (size_t)l * self_layer_elems() * 2;
            a.page_table = page_table; a.pages_per_seq = pages_per_seq; a.page_tokens = PAGE_TOKENS;
            ProfScope ps(this, PROF_SELF_ATTN, st);
            decode_attention(a, st);
        }
        { ProfScope ps(this, PROF_DEC_GEMM, st); launch_chain(base, base + 3, st); }
        {
            DecAttnArgs a;
            a.dtype = dt; a.out = datt; a.out_stride = d; a.B = B; a.H = g.n_heads;
            a.q_parts = dpart; a.q_n_parts = chain_q_parts[l]; a.q_part_stride = (long long)B * d; a.q_bias = m->dec[l].cq.b;
            a.state = nullptr; a.n_keys = g.n_ctx; a.active = active; a.row_active = unfinished;
            const size_t per_kv = (size_t)max_batch * g.n_heads * g.n_ctx * 64;
            a.k = (uint8_t*)cross + (size_t)l * cross_layer_elems() * 2;
            a.v = (uint8_t*)cross + ((size_t)l * cross_layer_elems() + per_kv) * 2;
            a.kv_bstride = (long long)g.n_heads * g.n_ctx * 64; a.kv_hstride = (long long)g.n_ctx * 64;
            ProfScope ps(this, PROF_CROSS_ATTN, st);
            decode_attention(a, st);
        }
        { ProfScope ps(this, PROF_DEC_GEMM, st); launch_chain(base + 3, l + 1 < L ? base + 9 : base + 8, st); }
    }
    }
    lm_head(st);
}

}  // namespace wb
